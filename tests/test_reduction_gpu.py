"""Device reduction (checkm_b200/csrc/reduce.cu via checkm_b200.resultsParser) against goldens produced by the
REFERENCE's own ResultsParser / MarkerSetParser / PFAM code (tests/golden/make_reduction_goldens.py).
Bit-exact: marker hit tables (names, merged names, coordinates, order), copy-number histograms, completeness and
contamination (float64 repr), and the printed QA tables."""
import io
import json
import os
from contextlib import redirect_stdout

import pytest

from conftest import CPR_HMM, GOLDEN

pytestmark = pytest.mark.gpu
RED = os.path.join(GOLDEN, 'reduction')


class _AAI:
    aaiMeanBinHetero = {}


def _dump(rm):
    return [[acc, [[h.target_name, h.target_length, h.hmm_from, h.hmm_to, h.ali_from, h.ali_to, h.env_from, h.env_to,
                    h.dom_score, h.full_score, h.full_e_value, h.i_evalue] for h in hits]] for acc, hits in rm.markerHits.items()]


@pytest.fixture(scope='module')
def golden():
    with open(os.path.join(RED, 'reduction_goldens.json')) as f:
        return json.load(f)


@pytest.fixture(scope='module')
def setup(engine):
    from checkm_b200.defaultValues import DefaultValues
    from checkm_b200.hmmerModelParser import HmmModelParser
    DefaultValues.set_data_root(os.path.join(RED, 'data'))
    return HmmModelParser(CPR_HMM).models()


KW = {'default': {}, 'noadj': {'bSkipAdjCorrection': True}, 'nopseudo': {'bSkipPseudoGeneCorrection': True},
      'ignore': {'bIgnoreThresholds': True}}


@pytest.mark.parametrize('case', ['kat1', 'kat2', 'synth', 'stress'])
@pytest.mark.parametrize('label', ['default', 'noadj', 'nopseudo', 'ignore'])
def test_reduction_matches_reference(case, label, golden, setup):
    from checkm_b200.markerSets import MarkerSetParser
    from checkm_b200.resultsParser import ResultsParser
    models = setup
    g = golden[case][label]
    cdir = os.path.join(RED, case)
    binIds = sorted(g['hits'].keys())
    rp = ResultsParser({b: models for b in binIds})
    rp.analyseResults(cdir, 'bin_stats.analyze.tsv', 'hmmer.analyze.txt', **KW[label])
    for b in binIds:
        assert _dump(rp.results[b]) == g['hits'][b], (case, label, b)
    for which, mfile in (('hmm', CPR_HMM), ('taxon', os.path.join(RED, 'taxon.ms'))):
        ms = MarkerSetParser().getMarkerSets(cdir, binIds, mfile)
        for b in binIds:
            rm = rp.results[b]
            exp = g[which]['bins'][b]
            got_c = rm.geneCountsForSelectedMarkerSet(ms[b], False)
            got_i = rm.geneCountsForSelectedMarkerSet(ms[b], True)
            assert [repr(v) for v in got_c] == [repr(v) for v in exp['counts_colloc']], (case, label, which, b)
            assert [repr(v) for v in got_i] == [repr(v) for v in exp['counts_indiv']], (case, label, which, b)
            assert list(rm.countUniqueHits()) == exp['unique']
        for fmt in (1, 5, 6, 8):
            buf = io.StringIO()
            with redirect_stdout(buf):
                rp.printSummary(fmt, _AAI(), ms, False, None, True, '', None)
            assert buf.getvalue() == g[which]['table%d' % fmt], (case, label, which, fmt)


def test_sidecar_equals_text_path(engine, setup, tmp_path):
    """Search -> domtblout + side-car; the reduction must give the same table from either."""
    import shutil
    import numpy as np
    from tools import synth
    from checkm_b200.hmmer import HMMERRunner
    from checkm_b200.resultsParser import ResultsParser
    hm = synth.read_hmms(CPR_HMM)
    b = synth.make_bin('sc', hm, seed=77, n_orfs=200, split_prob=0.5, max_len=800)
    out = tmp_path / 'run'
    (out / 'bins' / 'sc').mkdir(parents=True)
    (out / 'storage').mkdir()
    faa = str(out / 'bins' / 'sc' / 'genes.faa')
    with open(faa, 'w') as f:
        f.write(b.fasta())
    (out / 'storage' / 'bin_stats.analyze.tsv').write_text("sc\t{'GC': 0.5}\n")
    table = str(out / 'bins' / 'sc' / 'hmmer.analyze.txt')
    HMMERRunner().search(CPR_HMM, faa, table, '/dev/null', '--cpu 1 --notextw -E 0.1 --domE 0.1 --noali', False)
    assert os.path.exists(table + '.ckm.npz')
    rp1 = ResultsParser({'sc': setup})
    rp1.analyseResults(str(out), 'bin_stats.analyze.tsv', 'hmmer.analyze.txt')
    os.remove(table + '.ckm.npz')
    rp2 = ResultsParser({'sc': setup})
    rp2.analyseResults(str(out), 'bin_stats.analyze.tsv', 'hmmer.analyze.txt')
    assert _dump(rp1.results['sc']) == _dump(rp2.results['sc'])
    assert len(rp1.results['sc'].markerHits) > 5
