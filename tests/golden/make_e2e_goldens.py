"""Goldens for the end-to-end entry-point test (BASELINE.json configs[0]: taxonomy_wf-style plumbing on 3 bins, plus the
HMM-file and lineage branches).  Run in the build container only (imports the reference read-only):

    python tests/golden/make_e2e_goldens.py

What it freezes under tests/golden/e2e/:
  bins/binA.faa, binB.faa.gz, binC.faa   three seeded synthetic protein bins (Prodigal naming, planted + split + tandem homologs)
  data/pfam/Pfam-A.hmm.dat, data/selected_marker_sets.tsv   a CheckM data root in miniature (hmms/checkm.hmm := cpr_43_markers.hmm,
                                                            copied in at test time)
  markers/taxon.ms, markers/lineage.ms    a taxon marker file over 24 accessions and a lineage file with DIFFERENT sets per bin
  expected.json                           per mode (hmm | taxon | lineage):
      subset[bin]        accessions searched for the bin = marker genes + Pfam clan mates (checkm/markerSets.py:443-454), db order
      domtblout[bin]     data lines of the ORACLE's domtblout for that bin x subset (what `hmmsearch --domtblout` would hold)
      models[bin]        {acc: [name, leng, ga, tc, nc]} as the REFERENCE's HmmModelParser reads the `hmmfetch -f` output
      the REFERENCE's ResultsParser output on those tables: marker hits, printSummary formats 1-9, bin_stats_ext.tsv,
      marker_gene_stats.tsv (checkm/resultsParser.py:50-143,275-319,567-968; checkm/main.py:325-343,424-457)
The GPU test (tests/test_find_e2e_gpu.py) drives checkm_b200.MarkerGeneFinder.find -> ResultsParser on the same inputs.
"""
import ast
import gzip
import io
import json
import os
import shutil
import sys
import tempfile
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
E2E = os.path.join(HERE, 'e2e')
CPR = os.path.join(HERE, 'cpr_43_markers.hmm')
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

CLAN_FILE = """# STOCKHOLM 1.0
#=GF ID   Ribosomal_L23
#=GF AC   PF00276.21
#=GF CL   CL0001
//
# STOCKHOLM 1.0
#=GF ID   Ribosomal_L5
#=GF AC   PF00281.20
#=GF CL   CL0001
//
# STOCKHOLM 1.0
#=GF ID   Ribosomal_S17
#=GF AC   PF00366.21
#=GF CL   CL0002
#=GF NE   Ribosomal_S9
//
# STOCKHOLM 1.0
#=GF ID   Ribosomal_S9
#=GF AC   PF00380.20
#=GF CL   CL0002
//
# STOCKHOLM 1.0
#=GF ID   Ribosomal_S11
#=GF AC   PF00411.20
#=GF CL   CL0002
//
# STOCKHOLM 1.0
#=GF ID   tRNA-synt_2d
#=GF AC   PF01409.21
#=GF CL   CL0040
//
# STOCKHOLM 1.0
#=GF ID   tRNA-synt_2c
#=GF AC   PF01411.20
#=GF CL   CL0040
//
# STOCKHOLM 1.0
#=GF ID   tRNA-synt_His
#=GF AC   PF13393.7
#=GF CL   CL0040
//
# STOCKHOLM 1.0
#=GF ID   Ribosomal_S8
#=GF AC   PF00410.20
//
"""

# uid -> selected uid (checkm data: selected_marker_sets.tsv).  UID7 -> UID3 -> UID1 forces the walk-up of
# BinMarkerSets.setLineageSpecificSelectedMarkerSet (markerSets.py:95-121) for a bin that carries UID7 and UID1 only.
SELECTED = [('UID1', 'UID1'), ('UID3', 'UID1'), ('UID5', 'UID5'), ('UID7', 'UID3'), ('UID9', 'UID9')]

BIN_STATS = ("{'GC': %r, 'GC std': 0.0213, 'Genome size': %d, '# ambiguous bases': 0, '# scaffolds': 4, '# contigs': 4, "
             "'Longest scaffold': 90000, 'Longest contig': 90000, 'N50 (scaffolds)': 60000, 'N50 (contigs)': 60000, "
             "'Mean scaffold length': 45000.5, 'Mean contig length': 45000.5, 'Coding density': 0.8812, 'Translation table': 11, "
             "'# predicted genes': %d}")


def data_lines(path):
    return [l.rstrip('\n') for l in open(path) if l.strip() and not l.startswith('#')]


def split_models(path):
    """Raw text records of an HMMER3 file, in order."""
    recs, cur = [], []
    for line in open(path):
        cur.append(line)
        if line.startswith('//'):
            recs.append(''.join(cur))
            cur = []
    return recs


class _AAI:
    aaiMeanBinHetero = {}


def main():
    import numpy as np
    os.environ['CHECKM_DATA_PATH'] = os.path.join(E2E, 'data')
    from oracle import pyoracle as po
    from tools import synth

    shutil.rmtree(E2E, ignore_errors=True)
    for d in ('bins', 'data/pfam', 'markers'):
        os.makedirs(os.path.join(E2E, d))
    with open(os.path.join(E2E, 'data', 'pfam', 'Pfam-A.hmm.dat'), 'w') as f:
        f.write(CLAN_FILE)
    with open(os.path.join(E2E, 'data', 'selected_marker_sets.tsv'), 'w') as f:
        for a, b in SELECTED:
            f.write('%s\t%s\n' % (a, b))
    # the reference reads DefaultValues.HMM_MODELS only through hmmfetch; nothing to place under data/hmms here

    from checkm.hmmerModelParser import HmmModelParser
    from checkm.markerSets import MarkerSetParser
    from checkm.resultsParser import ResultsParser
    from checkm.util.pfam import PFAM
    from checkm.defaultValues import DefaultValues
    assert DefaultValues.PFAM_CLAN_FILE.startswith(E2E), DefaultValues.PFAM_CLAN_FILE

    hf = po.HmmFile(CPR)
    accs = hf.accs()
    recs = split_models(CPR)
    assert len(recs) == len(accs) == 43
    hm = synth.read_hmms(CPR)

    # ---- bins ----
    bins = {}
    for binId, seed, kw in (('binA', 501, dict(n_orfs=150, split_prob=0.35, tandem_prob=0.15)),
                            ('binB', 502, dict(n_orfs=200, split_prob=0.2, tandem_prob=0.0, copies=(0, 0, 1, 1, 2, 3))),
                            ('binC', 503, dict(n_orfs=170, split_prob=0.0, tandem_prob=0.4, copies=(0, 1, 1, 2)))):
        b = synth.make_bin(binId, hm, seed=seed, max_len=700, sharpen=0.45, **kw)
        bins[binId] = b
        text = b.fasta()
        # 60-column FASTA like Prodigal writes
        out = []
        for rec in text.strip().split('\n>'):
            head, seq = rec.lstrip('>').split('\n', 1)
            seq = seq.replace('\n', '')
            out.append('>' + head + '\n' + '\n'.join(seq[i:i + 60] for i in range(0, len(seq), 60)) + '\n')
        text = ''.join(out)
        if binId == 'binB':
            with open(os.path.join(E2E, 'bins', 'binB.faa.gz'), 'wb') as raw:
                with gzip.GzipFile(filename='', mode='wb', fileobj=raw, mtime=0) as gz:
                    gz.write(text.encode())
        else:
            with open(os.path.join(E2E, 'bins', binId + '.faa'), 'w') as f:
                f.write(text)
    binIds = sorted(bins)

    # ---- marker files ----
    pf = [a for a in accs if a.startswith('PF')]
    tg = [a for a in accs if a.startswith('TIGR')]
    taxon_sets = [set(pf[0:3]), set(pf[3:5] + tg[0:2]), set(tg[2:7]), set(tg[7:12]), {pf[6]}, set(tg[12:18] + [pf[9]])]
    with open(os.path.join(E2E, 'markers', 'taxon.ms'), 'w') as f:
        f.write('# [Taxon Marker File]\n')
        f.write('Bacteria\t2\t42\tk__Bacteria\t5449\t%s\t0\troot\t5656\t%s\n' % (str(taxon_sets), str([set(pf[0:2]), set(tg[0:4])])))
    lineage = {
        'binA': [('UID5', 'k__Bacteria;p__Firmicutes', 120, [set(pf[0:1] + tg[0:3]), set(tg[3:9]), {pf[3]}, set(tg[20:25])]),
                 ('UID1', 'root', 5656, [set(tg[0:6]), set(tg[6:12])])],
        'binB': [('UID7', 'k__Bacteria;p__Proteobacteria;c__Gamma', 33, [set(tg[10:14]), set([pf[9], pf[11]] + tg[14:16])]),
                 ('UID1', 'root', 5656, [set(tg[0:6]), set(tg[6:12]), set(pf[4:6])])],
        'binC': [('UID9', 'k__Archaea', 207, [set(tg[25:31]), set(pf[7:9]), set(tg[16:20])])],
    }
    with open(os.path.join(E2E, 'markers', 'lineage.ms'), 'w') as f:
        f.write('# [Lineage Marker File]\n')
        for binId in binIds:
            f.write(binId + '\t' + str(len(lineage[binId])))
            for uid, lin, ng, sets in lineage[binId]:
                f.write('\t%s\t%s\t%d\t%s' % (uid, lin, ng, str(sets)))
            f.write('\n')

    expected = {}
    work = tempfile.mkdtemp(prefix='e2e_gold_')
    for mode, markerFile in (('hmm', CPR), ('taxon', os.path.join(E2E, 'markers', 'taxon.ms')),
                             ('lineage', os.path.join(E2E, 'markers', 'lineage.ms'))):
        out = os.path.join(work, mode)
        os.makedirs(os.path.join(out, 'storage'))
        msp = MarkerSetParser()
        kind = msp.markerFileType(markerFile)
        entry = {'subset': {}, 'domtblout': {}, 'models': {}}
        binIdToModels = {}
        with open(os.path.join(out, 'storage', 'bin_stats.analyze.tsv'), 'w') as f:
            for i, binId in enumerate(binIds):
                f.write(binId + '\t' + BIN_STATS % (0.41 + 0.07 * i, 180000 + 1111 * i, bins[binId].nseq) + '\n')
        for binId in binIds:
            b = bins[binId]
            bdir = os.path.join(out, 'bins', binId)
            os.makedirs(bdir)
            with open(os.path.join(bdir, 'genes.faa'), 'w') as f:
                f.write(b.fasta())
            # the bin's model subset exactly as markerSets.py:443-454 forms it (reference code), then database order
            if kind == 3:
                want = set(accs)
            else:
                bms = msp.parseTaxonomicMarkerSetFile(markerFile) if kind == 1 else msp.parseLineageMarkerSetFile(markerFile)[binId]
                genes = bms.getMarkerGenes()
                want = genes | PFAM(DefaultValues.PFAM_CLAN_FILE).genesInSameClan(genes)
            idx = [i for i, a in enumerate(accs) if a in want]
            assert len(idx) == len([a for a in want if a in accs])
            entry['subset'][binId] = [accs[i] for i in idx]
            # what `hmmfetch -f checkm.hmm keys` leaves behind, read by the reference's header parser (markerGeneFinder.py:160-163)
            sub = os.path.join(work, 'sub.hmm')
            with open(sub, 'w') as f:
                f.write(''.join(recs[i] for i in idx))
            models = HmmModelParser(sub).models()
            binIdToModels[binId] = models
            entry['models'][binId] = {a: [m.name, m.leng, m.ga, m.tc, m.nc] for a, m in models.items()}
            # the search itself: the oracle on this bin x subset, written as domtblout text
            rp = po.search(hf, b.residues, b.offsets, nthreads=8, models=idx)
            table = os.path.join(bdir, 'hmmer.analyze.txt')
            po.write_domtblout(rp, hf, b.names, b.descs, table, models=idx)
            po.free_results(rp)
            entry['domtblout'][binId] = data_lines(table)
        # ---- the reference's own reduction + reports on those tables (main.py:424-457) ----
        bms_all = msp.getMarkerSets(out, binIds, markerFile)
        RP = ResultsParser(binIdToModels)
        RP.analyseResults(out, 'bin_stats.analyze.tsv', 'hmmer.analyze.txt')
        entry['hits'] = {b: [[acc, [[h.target_name, h.target_length, h.hmm_from, h.hmm_to, h.ali_from, h.ali_to, h.env_from, h.env_to,
                                      h.dom_score, h.full_score, h.full_e_value, h.i_evalue] for h in hits]]
                             for acc, hits in RP.results[b].markerHits.items()] for b in binIds}
        entry['selected_uid'] = {b: str(bms_all[b].selectedMarkerSet().UID) for b in binIds}
        entry['counts'] = {b: {'colloc': RP.results[b].geneCountsForSelectedMarkerSet(bms_all[b], False),
                               'indiv': RP.results[b].geneCountsForSelectedMarkerSet(bms_all[b], True),
                               'unique': list(RP.results[b].countUniqueHits())} for b in binIds}
        tables = {}
        for fmt in range(1, 10):
            for tab in (True, False):
                if not tab and fmt not in (1, 2, 3):
                    continue
                buf = io.StringIO()
                with redirect_stdout(buf):
                    RP.printSummary(fmt, _AAI(), bms_all, False, None, tab, '', out)
                tables['%d%s' % (fmt, 't' if tab else 'p')] = buf.getvalue()
        entry['tables'] = tables
        RP.cacheResults(out, bms_all, False)
        for name in ('bin_stats_ext.tsv', 'marker_gene_stats.tsv'):
            d = {}
            for line in open(os.path.join(out, 'storage', name)):
                k, v = line.rstrip('\n').split('\t', 1)
                d[k] = ast.literal_eval(v)
            entry[name] = d
        expected[mode] = entry
        print(mode, {b: (len(entry['subset'][b]), len(entry['domtblout'][b]), entry['counts'][b]['colloc']) for b in binIds})
    with open(os.path.join(E2E, 'expected.json'), 'w') as f:
        json.dump(expected, f, indent=0, sort_keys=True)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == '__main__':
    main()
