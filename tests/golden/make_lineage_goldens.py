"""Goldens for the lineage marker-set selection (checkm/treeParser.py:468-553 and helpers).  Run in
the build container only (imports the reference read-only):

    python tests/golden/make_lineage_goldens.py

The reference's `checkm.treeParser` imports `dendropy`, which is not installed here.  Its code needs only a handful of
tree-walking names, so this script puts a STAND-IN `dendropy` module into `sys.modules` (a small recursive Newick reader
written here, independent of `checkm_b200/util/newick.py`) and then runs the REFERENCE's own `TreeParser` methods --
`getBinMarkerSets`, `readLineageMetadata`, `getBinTaxonomy`, `getInsertionBranchId` -- over a synthetic placed tree, a
synthetic `genome_tree.metadata.tsv` / `missing_duplicate_genes_50.tsv`, and the reference's own `ResultsParser` reduction
of the e2e goldens' phylogenetic search tables.  The selection logic that is frozen is therefore the reference's; only the
file reader underneath it is a stand-in.

What it freezes under tests/golden/lineage/:
  tree/concatenated.tre                    the placed tree: reference genomes IMG_*, labelled internal nodes `UID|taxonomy|`,
                                           bins inserted by "pplacer" under unlabelled nodes (nested twice for binD/binE),
                                           one bin on the branch between the root and the bacterial domain node (binB)
  data/genome_tree/*.tsv                   node metadata and lineage-specific missing / duplicated genes
  data/selected_marker_sets.tsv            uid -> uid of the marker set `checkm analyze` selects (markerSets.py:95-121)
  data/pfam/Pfam-A.hmm.dat                 the e2e goldens' clan file (the reduction of the search tables reads it)
  (the `hmmer.tree.txt` of the seven bins are the e2e goldens' tables: binD := binA's, binE := binC's, binF := binB's, binG's is empty)
  expected.json['wf']                      `lineage_wf` in miniature on the e2e bins binA / binB / binC: tree search tables ->
                                           the reference's reduction -> getBinMarkerSets -> per-bin subsets (marker genes + clan
                                           mates) -> the oracle's domtblout -> the reference's ResultsParser: QA tables 1 and 2
  expected.json                            per option variant: {bin: [[uid, lineage, numGenomes, [sorted set, ...]], ...]}, the
                                           unique / multi-copy counts the selection saw, and the three look-up tables
"""
import json
import os
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, 'lineage')
CPR = os.path.join(HERE, 'cpr_43_markers.hmm')
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')


# ------------------------------------------------------------------ stand-in dendropy (test infrastructure) ----
class _Taxon(object):
    def __init__(self, label):
        self.label = label


class _Node(object):
    def __init__(self):
        self.label = None
        self.taxon = None
        self.parent_node = None
        self.kids = []

    def child_nodes(self):
        return list(self.kids)

    def is_internal(self):
        return bool(self.kids)

    def leaf_nodes(self):
        if not self.kids:
            return [self]
        out = []
        for k in self.kids:
            out.extend(k.leaf_nodes())
        return out

    def sister_nodes(self):
        return [k for k in self.parent_node.kids if k is not self] if self.parent_node else []

    def walk(self):
        yield self
        for k in self.kids:
            for n in k.walk():
                yield n


class _Tree(object):
    def __init__(self, root):
        self.seed_node = root

    @classmethod
    def get_from_path(cls, path, schema, rooting=None, preserve_underscores=False):
        assert schema == 'newick' and preserve_underscores
        text = open(path).read().strip()
        pos = [0]

        def label():
            if pos[0] < len(text) and text[pos[0]] == "'":
                end = text.index("'", pos[0] + 1)
                s = text[pos[0] + 1:end]
                pos[0] = end + 1
                return s
            start = pos[0]
            while pos[0] < len(text) and text[pos[0]] not in '(),:;':
                pos[0] += 1
            return text[start:pos[0]]

        def clade(parent):
            node = _Node()
            node.parent_node = parent
            if text[pos[0]] == '(':
                pos[0] += 1
                while True:
                    node.kids.append(clade(node))
                    if text[pos[0]] == ',':
                        pos[0] += 1
                        continue
                    assert text[pos[0]] == ')', text[pos[0]:pos[0] + 20]
                    pos[0] += 1
                    break
                lab = label()
                node.label = lab if lab else None
            else:
                node.taxon = _Taxon(label())
            if pos[0] < len(text) and text[pos[0]] == ':':
                pos[0] += 1
                while text[pos[0]] not in '(),;':
                    pos[0] += 1
            return node

        root = clade(None)
        assert text[pos[0]] == ';'
        return cls(root)

    def find_node(self, filter_fn):
        for n in self.seed_node.walk():
            if filter_fn(n):
                return n
        return None

    def find_node_with_taxon_label(self, label):
        for n in self.seed_node.walk():
            if n.taxon is not None and n.taxon.label == label:
                return n
        return None


def install_dendropy_stand_in():
    mod = types.ModuleType('dendropy')
    mod.Tree = _Tree
    sys.modules['dendropy'] = mod


# ------------------------------------------------------------------ the synthetic placed tree ----
# binA: deep inside Firmicutes, next to a node with a single genome (skipped at numGenomesMarkers = 2) and one without
#       taxonomy (named after its next named ancestor); binD / binE: two bins inserted at the same place (two unlabelled
#       nodes between them and Proteobacteria); binB: on the branch between the root and the bacterial domain node (the
#       `bRoot` case: the walk starts below the domain node); binC: Archaea; binG: in the tree but with an empty search
#       table (too few unique markers -> domain set forced); binF: has a bin directory but was not placed.
TREE = ("(("
        "binB:0.2,"
        "((((('IMG_1001':0.1,binA:0.05):0.01,IMG_1002:0.1)'UID30||':0.1,IMG_1003:0.2)'UID20|c__Bacilli;o__Lactobacillales|':0.3,"
        "((IMG_1004:0.1,binG:0.3):0.02,IMG_1005:0.1)'UID21|c__Clostridia|':0.2)'UID10|p__Firmicutes|':0.5,"
        "((IMG_1006:0.1,((binD:0.1,binE:0.1):0.01,IMG_1008:0.2):0.02)'UID22||':0.1,IMG_1007:0.3)'UID11|p__Proteobacteria|':0.4"
        ")'UID2|k__Bacteria|':0.6"
        "):0.1,"
        "((IMG_2001:0.1,(IMG_2002:0.2,binC:0.4):0.1)'UID12|p__Euryarchaeota|':0.3,IMG_2003:0.5)'UID3|k__Archaea|':0.7"
        ")'UID1|root|';\n")


def metadata(accs):
    pf = [a for a in accs if a.startswith('PF')]
    tg = [a for a in accs if a.startswith('TIGR')]
    rows = [
        # uid, #genomes, taxonomy, bootstrap, marker set
        ('UID1', 5656, 'root', 'NA', [set(tg[0:6]), set(tg[6:12]), set(pf[0:2])]),
        ('UID2', 5449, 'k__Bacteria', '100', [set(pf[0:3]), set(pf[3:5] + tg[0:2]), set(tg[2:7]), set(tg[7:12]), {pf[6]}]),
        ('UID3', 207, 'k__Archaea', '100', [set(tg[25:31]), set(pf[7:9]), set(tg[16:20])]),
        ('UID10', 120, 'k__Bacteria;p__Firmicutes', '95', [set(pf[0:1] + tg[0:3]), set(tg[3:9]), {pf[3]}, set(tg[20:25])]),
        ('UID11', 88, 'k__Bacteria;p__Proteobacteria', '55', [set(tg[10:14]), set([pf[9], pf[11]] + tg[14:16]), {pf[2]}]),
        ('UID12', 60, 'k__Archaea;p__Euryarchaeota', '80', [set(tg[25:28]), set(pf[7:8] + tg[16:18])]),
        ('UID20', 30, 'k__Bacteria;p__Firmicutes;c__Bacilli;o__Lactobacillales', '72', [set(tg[3:6] + pf[0:1]), set(tg[20:23]), {pf[3], pf[5]}]),
        ('UID21', 12, 'k__Bacteria;p__Firmicutes;c__Clostridia', '40', [set(tg[1:5]), set(pf[4:6])]),
        ('UID22', 7, '', '65', [set(tg[10:12]), {pf[9]}, set(tg[14:16])]),
        ('UID30', 1, '', '60', [set(tg[3:5])]),
    ]
    missing_dup = {
        'UID1': (set(), set()),
        'UID2': ({'pfam%s' % pf[1][2:7]}, set()),
        'UID3': (set(), {tg[17]}),
        'UID10': ({tg[4]}, {'pfam%s' % pf[3][2:7]}),
        'UID11': ({tg[11], tg[14]}, set()),
        'UID12': ({tg[26]}, {tg[25]}),
        'UID20': ({tg[5], tg[21]}, {tg[0]}),
        'UID21': ({tg[2]}, {'pfam%s' % pf[4][2:7]}),
        'UID22': ({tg[10]}, {tg[15], 'pfam%s' % pf[9][2:7]}),
        'UID30': ({tg[3]}, {tg[7]}),
    }
    return rows, missing_dup


VARIANTS = {
    # name: (numGenomesMarkers, bootstrap, bNoLineageSpecificRefinement, bForceDomain, bRequireTaxonomy, minUnique, maxMulti)
    'lineage_wf_defaults': (2, 0, False, False, False, 10, 10),
    'no_refinement': (2, 0, True, False, False, 10, 10),
    'force_domain': (2, 0, False, True, False, 10, 10),
    'strict_nodes': (50, 70, False, False, True, 10, 10),
    'few_markers_allowed': (2, 0, False, False, False, 0, 1000),
    'many_unique_needed': (2, 0, True, False, False, 25, 0),
}
BIN_TABLE_FROM = {'binA': 'binA', 'binB': 'binB', 'binC': 'binC', 'binD': 'binA', 'binE': 'binC', 'binF': 'binB', 'binG': None}


def parse_marker_file(path):
    out = {}
    with open(path) as f:
        assert f.readline().rstrip('\n') == '# [Lineage Marker File]'
        for line in f:
            fields = line.rstrip('\n').split('\t')
            sets = []
            for i in range(int(fields[1])):
                uid, lineage, n, text = fields[2 + 4 * i:6 + 4 * i]
                sets.append([uid, lineage, int(n), [sorted(s) for s in eval(text)]])
            out[fields[0]] = sets
    return out


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    for d in ('tree', 'data/genome_tree', 'data/pfam'):
        os.makedirs(os.path.join(OUT, d))
    shutil.copyfile(os.path.join(HERE, 'e2e', 'data', 'pfam', 'Pfam-A.hmm.dat'), os.path.join(OUT, 'data', 'pfam', 'Pfam-A.hmm.dat'))
    os.environ['CHECKM_DATA_PATH'] = os.path.join(OUT, 'data')
    install_dendropy_stand_in()
    from checkm.defaultValues import DefaultValues
    from checkm.hmmerModelParser import HmmModelParser
    from checkm.resultsParser import ResultsParser
    from checkm.treeParser import TreeParser
    assert DefaultValues.GENOME_TREE_DIR.startswith(OUT)

    models = HmmModelParser(CPR).models()
    accs = list(models.keys())
    rows, missing_dup = metadata(accs)
    with open(os.path.join(OUT, 'tree', 'concatenated.tre'), 'w') as f:
        f.write(TREE)
    with open(os.path.join(OUT, 'data', 'genome_tree', 'genome_tree.metadata.tsv'), 'w') as f:
        f.write('UID\t# genomes\ttaxonomy\tbootstrap\tgc mean\tgc std\tgenome size mean\tgenome size std\tgene count mean\tgene count std\tmarker set\n')
        for i, (uid, n, tax, boot, sets) in enumerate(rows):
            f.write('%s\t%d\t%s\t%s\t%.2f\t%.2f\t%d\t%d\t%.1f\t%.1f\t%s\n' % (
                uid, n, tax, boot, 40.0 + i, 2.5 + 0.1 * i, 3000000 + 100000 * i, 500000 + 1000 * i, 2900.0 + 10 * i, 300.0 + i, str(sets)))
    with open(os.path.join(OUT, 'data', 'genome_tree', 'missing_duplicate_genes_50.tsv'), 'w') as f:
        for uid, (missing, dup) in missing_dup.items():
            f.write('%s\t%s\t%s\n' % (uid, str(missing) if missing else 'set()', str(dup) if dup else 'set()'))

    with open(os.path.join(OUT, 'data', 'selected_marker_sets.tsv'), 'w') as f:      # checkm data: uid -> uid of the set to use
        for uid, _, _, _, _ in rows:
            f.write('%s\t%s\n' % (uid, {'UID30': 'UID20', 'UID22': 'UID11', 'UID21': 'UID10'}.get(uid, uid)))

    e2e = json.load(open(os.path.join(HERE, 'e2e', 'expected.json')))['hmm']['domtblout']

    # the run directory of `checkm tree`: bins/<id>/hmmer.tree.txt, storage/tree/concatenated.tre, storage/bin_stats.tree.tsv
    work = tempfile.mkdtemp(prefix='lineage_gold_')
    os.makedirs(os.path.join(work, 'storage', 'tree'))
    shutil.copyfile(os.path.join(OUT, 'tree', 'concatenated.tre'), os.path.join(work, 'storage', 'tree', 'concatenated.tre'))
    binIds = sorted(BIN_TABLE_FROM)
    with open(os.path.join(work, 'storage', 'bin_stats.tree.tsv'), 'w') as f:
        for binId in binIds:
            os.makedirs(os.path.join(work, 'bins', binId))
            src = BIN_TABLE_FROM[binId]
            with open(os.path.join(work, 'bins', binId, 'hmmer.tree.txt'), 'w') as t:
                t.write('#' + ' ' * 70 + '--- full sequence --- -------------- this domain -------------\n')
                for line in (e2e[src] if src else []):
                    t.write(line + '\n')
            f.write("%s\t{'GC': 0.5, 'Genome size': 2000000, '# predicted genes': 150}\n" % binId)
    RP = ResultsParser({b: models for b in binIds})
    RP.analyseResults(work, 'bin_stats.tree.tsv', 'hmmer.tree.txt')

    expected = {'unique_multi': {b: list(RP.results[b].countUniqueHits()) for b in binIds}, 'variants': {}}
    tp = TreeParser()
    for name, (ng, boot, noref, force, reqtax, minu, maxm) in VARIANTS.items():
        mf = os.path.join(work, name + '.ms')
        tp.getBinMarkerSets(work, mf, ng, boot, noref, force, reqtax, RP, minu, maxm)
        expected['variants'][name] = {'options': [ng, boot, noref, force, reqtax, minu, maxm], 'bins': parse_marker_file(mf)}
        print(name, {b: [s[0] for s in v] for b, v in sorted(expected['variants'][name]['bins'].items())})
    meta = tp.readLineageMetadata(work, binIds)
    expected['lineage_metadata'] = {b: {k: (v if k != 'marker set' or v == 'NA' else [sorted(s) for s in eval(v)]) for k, v in d.items()}
                                    for b, d in meta.items()}
    expected['bin_taxonomy'] = tp.getBinTaxonomy(work, binIds)
    expected['insertion_uid'] = tp.getInsertionBranchId(work, binIds)
    print(expected['bin_taxonomy'])
    print(expected['insertion_uid'])
    expected['wf'] = workflow(models, accs)
    with open(os.path.join(OUT, 'expected.json'), 'w') as f:
        json.dump(expected, f, indent=0, sort_keys=True)
    shutil.rmtree(work)


BIN_STATS = ("{'GC': %r, 'GC std': 0.0213, 'Genome size': %d, '# ambiguous bases': 0, '# scaffolds': 4, '# contigs': 4, "
             "'Longest scaffold': 90000, 'Longest contig': 90000, 'N50 (scaffolds)': 60000, 'N50 (contigs)': 60000, "
             "'Mean scaffold length': 45000.5, 'Mean contig length': 45000.5, 'Coding density': 0.8812, 'Translation table': 11, "
             "'# predicted genes': %d}")


class _AAI:
    aaiMeanBinHetero = {}


def read_faa(path):
    import gzip
    opener = gzip.open if path.endswith('.gz') else open
    names, descs, seqs = [], [], []
    with opener(path, 'rt') as f:
        for line in f:
            line = line.rstrip('\n')
            if line.startswith('>'):
                head = line[1:].split(None, 1)
                names.append(head[0])
                descs.append(head[1] if len(head) > 1 else '')
                seqs.append([])
            elif line:
                seqs[-1].append(line)
    return names, descs, [''.join(s) for s in seqs]


def workflow(models, accs):
    """tree -> lineage_set -> analyze -> qa (checkm/main.py:181-343,424-457) with the oracle standing in for hmmsearch and the
    placed tree standing in for pplacer; every reduction, selection and report is the reference's code."""
    import io
    from contextlib import redirect_stdout
    import numpy as np
    from oracle import pyoracle as po
    from checkm.defaultValues import DefaultValues
    from checkm.hmmerModelParser import HmmModelParser
    from checkm.markerSets import MarkerSetParser
    from checkm.resultsParser import ResultsParser
    from checkm.treeParser import TreeParser
    from checkm.util.pfam import PFAM
    e2e_dir = os.path.join(HERE, 'e2e')
    e2e = json.load(open(os.path.join(e2e_dir, 'expected.json')))['hmm']['domtblout']
    binFiles = {'binA': 'binA.faa', 'binB': 'binB.faa.gz', 'binC': 'binC.faa'}
    binIds = sorted(binFiles)
    work = tempfile.mkdtemp(prefix='lineage_wf_gold_')
    os.makedirs(os.path.join(work, 'storage', 'tree'))
    shutil.copyfile(os.path.join(OUT, 'tree', 'concatenated.tre'), os.path.join(work, 'storage', 'tree', 'concatenated.tre'))
    bins = {}
    for binId in binIds:
        os.makedirs(os.path.join(work, 'bins', binId))
        with open(os.path.join(work, 'bins', binId, 'hmmer.tree.txt'), 'w') as f:
            for line in e2e[binId]:
                f.write(line + '\n')
        names, descs, seqs = read_faa(os.path.join(e2e_dir, 'bins', binFiles[binId]))
        dsq = [po.digitize(s) for s in seqs]
        offsets = np.zeros(len(dsq) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(d) for d in dsq])
        bins[binId] = (names, descs, np.concatenate(dsq), offsets)
    for name in ('bin_stats.tree.tsv', 'bin_stats.analyze.tsv'):
        with open(os.path.join(work, 'storage', name), 'w') as f:
            for i, binId in enumerate(binIds):
                f.write(binId + '\t' + BIN_STATS % (0.41 + 0.07 * i, 180000 + 1111 * i, len(bins[binId][0])) + '\n')
    # lineage_set (main.py:240-268)
    RPt = ResultsParser({b: models for b in binIds})
    RPt.analyseResults(work, 'bin_stats.tree.tsv', 'hmmer.tree.txt')
    mf = os.path.join(work, 'lineage.ms')
    TreeParser().getBinMarkerSets(work, mf, 2, 0, False, False, False, RPt, 10, 10)
    entry = {'marker_sets': parse_marker_file(mf), 'unique_multi': {b: list(RPt.results[b].countUniqueHits()) for b in binIds},
             'subset': {}, 'domtblout': {}}
    # analyze (main.py:325-343): per-bin subsets, the oracle as hmmsearch
    hf = po.HmmFile(CPR)
    recs, cur = [], []
    for line in open(CPR):
        cur.append(line)
        if line.startswith('//'):
            recs.append(''.join(cur))
            cur = []
    msp = MarkerSetParser()
    binIdToModels = {}
    for binId in binIds:
        genes = msp.parseLineageMarkerSetFile(mf)[binId].getMarkerGenes()
        want = genes | PFAM(DefaultValues.PFAM_CLAN_FILE).genesInSameClan(genes)
        idx = [i for i, a in enumerate(accs) if a in want]
        entry['subset'][binId] = [accs[i] for i in idx]
        sub = os.path.join(work, 'sub.hmm')
        with open(sub, 'w') as f:
            f.write(''.join(recs[i] for i in idx))
        binIdToModels[binId] = HmmModelParser(sub).models()
        names, descs, residues, offsets = bins[binId]
        rp = po.search(hf, residues, offsets, nthreads=8, models=idx)
        table = os.path.join(work, 'bins', binId, 'hmmer.analyze.txt')
        po.write_domtblout(rp, hf, names, descs, table, models=idx)
        po.free_results(rp)
        entry['domtblout'][binId] = [l.rstrip('\n') for l in open(table) if l.strip() and not l.startswith('#')]
    # qa (main.py:424-457)
    bms = msp.getMarkerSets(work, binIds, mf)
    RP = ResultsParser(binIdToModels)
    RP.analyseResults(work, 'bin_stats.analyze.tsv', 'hmmer.analyze.txt')
    entry['selected_uid'] = {b: str(bms[b].selectedMarkerSet().UID) for b in binIds}
    entry['counts'] = {b: RP.results[b].geneCountsForSelectedMarkerSet(bms[b], False) for b in binIds}
    entry['tables'] = {}
    for fmt in (1, 2, 3):
        buf = io.StringIO()
        with redirect_stdout(buf):
            RP.printSummary(fmt, _AAI(), bms, False, None, True, '', work)
        entry['tables'][str(fmt)] = buf.getvalue()
    print('wf', entry['selected_uid'], {b: entry['counts'][b][6:] for b in binIds}, {b: len(v) for b, v in entry['subset'].items()})
    shutil.rmtree(work)
    return entry


if __name__ == '__main__':
    main()
