"""End to end through the real entry point (BASELINE.json configs[0]):

    MarkerGeneFinder(1).find(binFiles, outDir, 'hmmer.analyze.txt', ..., markerFile, bCalledGenes=True)
      -> bins/<id>/genes.faa, domtblout (+ side-car)              checkm/markerGeneFinder.py:45-144
    MarkerSetParser.getMarkerSets / writeBinModels / loadBinModels  checkm/markerSets.py:248-285,524-540
    ResultsParser.analyseResults -> printSummary(1..9) -> cacheResults    checkm/resultsParser.py:50-143,275-319

three ways: an HMM file, a taxon marker file, and a LINEAGE marker file whose bins carry different marker sets (per-bin
query subsets = marker genes + Pfam clan mates, ckm_models_select + ckm_search_per_bin).  Expected values
(tests/golden/e2e/expected.json, made by tests/golden/make_e2e_goldens.py): the oracle's domtblout for the same
bin x subset, pushed through the REFERENCE's own HmmModelParser / MarkerSetParser / ResultsParser.
Everything is compared exactly: domtblout data lines as text, hit tables, float reprs, printed tables."""
import ast
import gzip
import io
import json
import os
import shutil
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import CPR_HMM, GOLDEN

pytestmark = pytest.mark.gpu
E2E = os.path.join(GOLDEN, 'e2e')
BINFILES = [os.path.join(E2E, 'bins', f) for f in ('binA.faa', 'binB.faa.gz', 'binC.faa')]
BIN_IDS = ['binA', 'binB', 'binC']
MARKER = {'hmm': CPR_HMM, 'taxon': os.path.join(E2E, 'markers', 'taxon.ms'), 'lineage': os.path.join(E2E, 'markers', 'lineage.ms')}


class _AAI:
    aaiMeanBinHetero = {}


@pytest.fixture(scope='module')
def expected():
    with open(os.path.join(E2E, 'expected.json')) as f:
        return json.load(f)


@pytest.fixture(scope='module')
def dataroot(tmp_path_factory, engine):
    """A CheckM data root in miniature: hmms/checkm.hmm is the reference's CPR fixture."""
    from checkm_b200.defaultValues import DefaultValues
    root = str(tmp_path_factory.mktemp('checkm_data'))
    shutil.copytree(os.path.join(E2E, 'data'), root, dirs_exist_ok=True)
    os.makedirs(os.path.join(root, 'hmms'))
    shutil.copyfile(CPR_HMM, os.path.join(root, 'hmms', 'checkm.hmm'))
    saved = DefaultValues.CHECKM_DATA_DIR
    DefaultValues.set_data_root(root)
    yield root
    DefaultValues.set_data_root(saved)


def _data_lines(path):
    return [l.rstrip('\n') for l in open(path) if l.strip() and not l.startswith('#')]


def _dump(rm):
    return [[acc, [[h.target_name, h.target_length, h.hmm_from, h.hmm_to, h.ali_from, h.ali_to, h.env_from, h.env_to,
                    h.dom_score, h.full_score, h.full_e_value, h.i_evalue] for h in hits]] for acc, hits in rm.markerHits.items()]


def _fmt4(text):
    """Format 4 lists the marker genes of a set in Python-set order (hash-seed dependent): compare as mappings."""
    out = {}
    lines = [l for l in text.split('\n') if l != '']
    for head, cnt in zip(lines[0::2], lines[1::2]):
        h, c = head.split('\t'), cnt.split('\t')
        out[(h[0], c[0])] = dict(zip(h[1:], c[1:]))
    return out


BIN_STATS = ("{'GC': %r, 'GC std': 0.0213, 'Genome size': %d, '# ambiguous bases': 0, '# scaffolds': 4, '# contigs': 4, "
             "'Longest scaffold': 90000, 'Longest contig': 90000, 'N50 (scaffolds)': 60000, 'N50 (contigs)': 60000, "
             "'Mean scaffold length': 45000.5, 'Mean contig length': 45000.5, 'Coding density': 0.8812, 'Translation table': 11, "
             "'# predicted genes': %d}")


@pytest.mark.parametrize('mode,batch_residues', [('hmm', None), ('taxon', None), ('lineage', None), ('lineage', '30000'), ('hmm', '1')])
def test_find_then_qa(mode, batch_residues, expected, dataroot, tmp_path, monkeypatch):
    """batch_residues: the default searches the three bins as one batch; '30000' / '1' force two / three batches through the
    reader -> searchers -> writer pipeline of MarkerGeneFinder (one bin per batch, both engines busy)."""
    if batch_residues is not None:
        monkeypatch.setenv('CKM_BATCH_RESIDUES', batch_residues)
    from checkm_b200.markerGeneFinder import MarkerGeneFinder
    from checkm_b200.markerSets import MarkerSetParser
    from checkm_b200.resultsParser import ResultsParser
    exp = expected[mode]
    out = str(tmp_path / 'out')
    os.makedirs(os.path.join(out, 'storage'))
    markerFile = MARKER[mode]

    # ---- analyze: find marker genes (main.py:325-343) ----
    binIdToModels = MarkerGeneFinder(1).find(BINFILES, out, 'hmmer.analyze.txt', 'hmmer.analyze.ali.txt', markerFile, False, False, True)
    assert sorted(binIdToModels.keys()) == BIN_IDS
    mismatched = []
    for binFile, binId in zip(BINFILES, BIN_IDS):
        bdir = os.path.join(out, 'bins', binId)
        opener = gzip.open if binFile.endswith('.gz') else open
        with opener(binFile, 'rt') as f:
            assert open(os.path.join(bdir, 'genes.faa')).read() == f.read()           # markerGeneFinder.py:118-127
        got = _data_lines(os.path.join(bdir, 'hmmer.analyze.txt'))
        want = exp['domtblout'][binId]
        assert len(got) == len(want), (mode, binId, len(got), len(want))
        mismatched += [(binId, g, w) for g, w in zip(got, want) if g.split() != w.split()]
        # the HmmModel dict of the bin = what HmmModelParser reads from the per-bin `hmmfetch -f` output (markerGeneFinder.py:160-163)
        models = binIdToModels[binId]
        assert list(models.keys()) == list(exp['models'][binId].keys())
        for acc, (name, leng, ga, tc, nc) in exp['models'][binId].items():
            m = models[acc]
            assert (m.name, m.leng) == (name, leng)
            assert [None if v is None else list(v) for v in (m.ga, m.tc, m.nc)] == [ga, tc, nc], acc
    print('%s: domtblout rows %d, rows whose text differs from the oracle\'s: %d' %
          (mode, sum(len(v) for v in exp['domtblout'].values()), len(mismatched)))

    msp = MarkerSetParser(1)
    binIdToBinMarkerSets = msp.getMarkerSets(out, BIN_IDS, markerFile)
    info = os.path.join(out, 'storage', 'checkm_hmm_info.pkl.gz')
    msp.writeBinModels(binIdToModels, info)
    with open(os.path.join(out, 'storage', 'bin_stats.analyze.tsv'), 'w') as f:
        for i, binId in enumerate(BIN_IDS):
            nseq = sum(1 for l in open(os.path.join(out, 'bins', binId, 'genes.faa')) if l.startswith('>'))
            f.write(binId + '\t' + BIN_STATS % (0.41 + 0.07 * i, 180000 + 1111 * i, nseq) + '\n')

    # ---- qa (main.py:424-457), twice: from the binary side-cars, then from the domtblout text alone ----
    loaded = MarkerSetParser(1).loadBinModels(info)
    assert {b: sorted(m.keys()) for b, m in loaded.items()} == {b: sorted(m.keys()) for b, m in binIdToModels.items()}
    for use_text in (False, True):
        if use_text:
            for binId in BIN_IDS:
                os.remove(os.path.join(out, 'bins', binId, 'hmmer.analyze.txt.ckm.npz'))
        RP = ResultsParser(loaded)
        RP.analyseResults(out, 'bin_stats.analyze.tsv', 'hmmer.analyze.txt')
        for b in BIN_IDS:
            assert _dump(RP.results[b]) == exp['hits'][b], (mode, b, use_text)
            assert str(binIdToBinMarkerSets[b].selectedMarkerSet().UID) == exp['selected_uid'][b]
            rm = RP.results[b]
            assert [repr(v) for v in rm.geneCountsForSelectedMarkerSet(binIdToBinMarkerSets[b], False)] == [repr(v) for v in exp['counts'][b]['colloc']]
            assert [repr(v) for v in rm.geneCountsForSelectedMarkerSet(binIdToBinMarkerSets[b], True)] == [repr(v) for v in exp['counts'][b]['indiv']]
            assert list(rm.countUniqueHits()) == exp['counts'][b]['unique']
        for key, want in exp['tables'].items():
            fmt, tab = int(key[0]), key[1] == 't'
            buf = io.StringIO()
            with redirect_stdout(buf):
                RP.printSummary(fmt, _AAI(), binIdToBinMarkerSets, False, None, tab, '', out)
            if fmt == 4:
                assert _fmt4(buf.getvalue()) == _fmt4(want), (mode, key)
            else:
                assert buf.getvalue() == want, (mode, key, use_text)
        RP.cacheResults(out, binIdToBinMarkerSets, False)
        for name in ('bin_stats_ext.tsv', 'marker_gene_stats.tsv'):
            got = {}
            for line in open(os.path.join(out, 'storage', name)):
                k, v = line.rstrip('\n').split('\t', 1)
                got[k] = ast.literal_eval(v)
            assert list(got.keys()) == list(exp[name].keys())
            for b in got:
                if name == 'bin_stats_ext.tsv':      # the GCN lists follow the bin's model-dict order: identical
                    assert {k: (repr(v) if isinstance(v, float) else v) for k, v in got[b].items()} == \
                           {k: (repr(v) if isinstance(v, float) else v) for k, v in exp[name][b].items()}, (mode, name, b)
                else:
                    assert got[b] == exp[name][b], (mode, name, b)
        # the cached files parse back (resultsParser.py:161-189)
        assert sorted(RP.parseBinStatsExt(out).keys()) == BIN_IDS and sorted(RP.parseMarkerGeneStats(out).keys()) == BIN_IDS
    assert not mismatched, mismatched[:3]


def test_subsets_and_fetch(expected, dataroot, tmp_path, oracle):
    """Per-bin HMM subsets (markerSets.py:326-343,443-476): marker genes + clan mates, and the file `hmmfetch -f` would
    have written -- here ckm_models_select + ckm_models_write -- read back by the oracle's full parser."""
    from checkm_b200 import runtime
    from checkm_b200.defaultValues import DefaultValues
    from checkm_b200.hmmer import HMMERRunner
    from checkm_b200.hmmerModelParser import HmmModelParser
    from checkm_b200.markerSets import MarkerSetParser
    msp = MarkerSetParser(1)
    full = oracle.HmmFile(CPR_HMM)
    full_accs = full.accs()
    for mode in ('taxon', 'lineage'):
        for binId in BIN_IDS:
            tmp = msp.createHmmModelFile(binId, MARKER[mode])
            try:
                sub = oracle.HmmFile(tmp)
                assert sub.accs() == expected[mode]['subset'][binId], (mode, binId)
                for i, acc in enumerate(sub.accs()):
                    j = full_accs.index(acc)
                    M = sub.headers[i].M
                    assert M == full.headers[j].M
                    for field, n in (('mat', (M + 1) * 20), ('ins', (M + 1) * 20), ('t', (M + 1) * 7)):
                        a = np.ctypeslib.as_array(getattr(sub.headers[i], field), shape=(n,))
                        b = np.ctypeslib.as_array(getattr(full.headers[j], field), shape=(n,))
                        assert np.array_equal(a, b), (acc, field)
                    assert list(sub.headers[i].evparam) == list(full.headers[j].evparam)
                got = HmmModelParser(tmp).models()
                assert list(got.keys()) == list(expected[mode]['models'][binId].keys())
            finally:
                os.remove(tmp)
    # HMMERRunner.fetch with a key file / a single key (hmmer.py:97-129)
    keys = tmp_path / 'keys.txt'
    keys.write_text('TIGR00029\nPF00276.21\nRibosomal_S9\n')
    outp = str(tmp_path / 'fetched.hmm')
    HMMERRunner(mode='fetch').fetch(DefaultValues.HMM_MODELS, str(keys), outp, bKeyFile=True)
    assert oracle.HmmFile(outp).accs() == ['PF00276.21', 'PF00380.20', 'TIGR00029']       # database order
    HMMERRunner(mode='fetch').index(outp)
    HMMERRunner(mode='fetch').fetch(DefaultValues.HMM_MODELS, 'TIGR00422', outp)
    assert oracle.HmmFile(outp).accs() == ['TIGR00422']
    assert runtime.models_for(DefaultValues.HMM_MODELS).n == 43


def test_find_with_an_empty_gene_file(dataroot, tmp_path):
    """A bin whose gene file is empty (no ORFs called) goes through find and qa like any other: empty table, zero counts
    (the reference's "processing empty gene files" case)."""
    from checkm_b200.markerGeneFinder import MarkerGeneFinder
    from checkm_b200.markerSets import MarkerSetParser
    from checkm_b200.resultsParser import ResultsParser
    empty = tmp_path / 'nothing.faa'
    empty.write_text('')
    out = str(tmp_path / 'out')
    os.makedirs(os.path.join(out, 'storage'))
    for files in ([str(empty)], [str(empty), BINFILES[0]]):
        models = MarkerGeneFinder(1).find(files, out, 'hmmer.analyze.txt', 'hmmer.analyze.ali.txt', CPR_HMM, False, False, True)
        ids = ['nothing'] + (['binA'] if len(files) == 2 else [])
        assert sorted(models.keys()) == sorted(ids)
        assert _data_lines(os.path.join(out, 'bins', 'nothing', 'hmmer.analyze.txt')) == []
        with open(os.path.join(out, 'storage', 'bin_stats.analyze.tsv'), 'w') as f:
            for b in ids:
                f.write("%s\t{'GC': 0.5}\n" % b)
        RP = ResultsParser(models)
        RP.analyseResults(out, 'bin_stats.analyze.tsv', 'hmmer.analyze.txt')
        ms = MarkerSetParser(1).getMarkerSets(out, ids, CPR_HMM)
        assert RP.results['nothing'].markerHits == {}
        assert RP.results['nothing'].geneCountsForSelectedMarkerSet(ms['nothing'], False) == [43, 0, 0, 0, 0, 0, 0.0, 0.0]


def test_lineage_wf_in_miniature(dataroot, tmp_path):
    """`checkm lineage_wf` minus Prodigal and pplacer (BASELINE.json configs[2]'s workflow):
    tree search (main.py:181-222) -> reduction -> TreeParser.getBinMarkerSets on the placed tree (treeParser.py:468-553) ->
    find with the lineage marker file it wrote (per-bin subsets) -> qa.  Expected values: tests/golden/lineage/expected.json['wf'],
    every selection / reduction / report step of which is the REFERENCE's code (tests/golden/make_lineage_goldens.py)."""
    from checkm_b200.defaultValues import DefaultValues
    from checkm_b200.markerGeneFinder import MarkerGeneFinder
    from checkm_b200.markerSets import MarkerSetParser, _parse_set_list
    from checkm_b200.resultsParser import ResultsParser
    from checkm_b200.treeParser import TreeParser
    lin = os.path.join(GOLDEN, 'lineage')
    with open(os.path.join(lin, 'expected.json')) as f:
        exp = json.load(f)['wf']
    # the data root of the e2e tests plus the genome-tree metadata
    shutil.copytree(os.path.join(lin, 'data', 'genome_tree'), os.path.join(dataroot, 'genome_tree'), dirs_exist_ok=True)
    saved_selected = open(DefaultValues.SELECTED_MARKER_SETS).read()
    shutil.copyfile(os.path.join(lin, 'data', 'selected_marker_sets.tsv'), DefaultValues.SELECTED_MARKER_SETS)
    try:
        out = str(tmp_path / 'out')
        os.makedirs(os.path.join(out, 'storage', 'tree'))
        shutil.copyfile(os.path.join(lin, 'tree', 'concatenated.tre'), os.path.join(out, 'storage', 'tree', 'concatenated.tre'))
        # ---- tree: the phylogenetic markers (the 43 CPR models stand in for phylo.hmm) ----
        phylo = MarkerGeneFinder(1).find(BINFILES, out, 'hmmer.tree.txt', 'hmmer.tree.ali.txt', CPR_HMM, False, False, True)
        with open(os.path.join(out, 'storage', 'bin_stats.tree.tsv'), 'w') as f:
            for i, binId in enumerate(BIN_IDS):
                nseq = sum(1 for l in open(os.path.join(out, 'bins', binId, 'genes.faa')) if l.startswith('>'))
                f.write(binId + '\t' + BIN_STATS % (0.41 + 0.07 * i, 180000 + 1111 * i, nseq) + '\n')
        shutil.copyfile(os.path.join(out, 'storage', 'bin_stats.tree.tsv'), os.path.join(out, 'storage', 'bin_stats.analyze.tsv'))
        # ---- lineage_set ----
        RPt = ResultsParser(phylo)
        RPt.analyseResults(out, 'bin_stats.tree.tsv', 'hmmer.tree.txt')
        assert {b: list(RPt.results[b].countUniqueHits()) for b in BIN_IDS} == exp['unique_multi']
        mf = os.path.join(out, 'lineage.ms')
        TreeParser().getBinMarkerSets(out, mf, 2, 0, False, False, False, RPt, 10, 10)
        got = {}
        for line in list(open(mf))[1:]:
            fields = line.rstrip('\n').split('\t')
            got[fields[0]] = [[fields[2 + 4 * i], fields[3 + 4 * i], int(fields[4 + 4 * i]),
                               [sorted(s) for s in _parse_set_list(fields[5 + 4 * i])]] for i in range(int(fields[1]))]
        assert got == exp['marker_sets']
        # ---- analyze + qa with the file just written ----
        models = MarkerGeneFinder(1).find(BINFILES, out, 'hmmer.analyze.txt', 'hmmer.analyze.ali.txt', mf, False, False, True)
        for binId in BIN_IDS:
            assert list(models[binId].keys()) == exp['subset'][binId]
            rows = _data_lines(os.path.join(out, 'bins', binId, 'hmmer.analyze.txt'))
            assert [r.split() for r in rows] == [r.split() for r in exp['domtblout'][binId]], binId
        bms = MarkerSetParser(1).getMarkerSets(out, BIN_IDS, mf)
        RP = ResultsParser(models)
        RP.analyseResults(out, 'bin_stats.analyze.tsv', 'hmmer.analyze.txt')
        for b in BIN_IDS:
            assert str(bms[b].selectedMarkerSet().UID) == exp['selected_uid'][b]
            assert [repr(v) for v in RP.results[b].geneCountsForSelectedMarkerSet(bms[b], False)] == [repr(v) for v in exp['counts'][b]]
        for fmt, want in exp['tables'].items():
            buf = io.StringIO()
            with redirect_stdout(buf):
                RP.printSummary(int(fmt), _AAI(), bms, False, None, True, '', out)
            assert buf.getvalue() == want, fmt
    finally:
        with open(DefaultValues.SELECTED_MARKER_SETS, 'w') as f:
            f.write(saved_selected)
