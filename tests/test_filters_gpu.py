"""Stages 2-4 parity on the GPU (checkm_b200/csrc/kernels_filters.cu) against the oracle (orc_filters):
pass/fail decisions identical; bias-filter and Viterbi scores bit-exact (integer / libm-exact arithmetic);
Forward scores within 1e-3 bits (north_star tolerance; fp32 summation order differs)."""
import numpy as np
import pytest

from tools import synth
from conftest import CPR_HMM

pytestmark = pytest.mark.gpu
LN2 = 0.69314718055994529


def _check(engine, models, ohf, b, oracle):
    db = engine.seqdb(b.residues, b.offsets)
    fs, vs, fw, ps = engine.filter_scores(models, db)
    st = engine.stats()
    cnt = np.zeros(4, int)
    worst_fwd = 0.0
    for m in range(models.n):
        for s in range(b.nseq):
            d = b.seq(s)
            if len(d) == 0:
                assert ps[m, s] == 0
                continue
            r = oracle.filters(ohf, m, d)
            exp = r.passed_msv | (r.passed_bias << 1) | (r.passed_vit << 2) | (r.passed_fwd << 3)
            assert ps[m, s] == exp, (m, s, len(d), int(ps[m, s]), exp, r.msv_sc, r.filtersc, r.vit_sc, r.fwd_sc, fs[m, s], vs[m, s], fw[m, s])
            cnt += [r.passed_msv, r.passed_bias, r.passed_vit, r.passed_fwd]
            if r.passed_msv:
                assert fs[m, s] == np.float32(r.filtersc), (m, s, fs[m, s], r.filtersc)
            if r.passed_bias and not np.isnan(r.vit_sc):
                assert vs[m, s] == np.float32(r.vit_sc), (m, s, vs[m, s], r.vit_sc)
            if r.passed_vit:
                err = abs(float(fw[m, s]) - float(r.fwd_sc)) / LN2
                worst_fwd = max(worst_fwd, err)
                assert err < 1e-3, (m, s, fw[m, s], r.fwd_sc)
    db.close()
    assert (st.n_past_msv, st.n_past_bias, st.n_past_vit, st.n_past_fwd) == tuple(cnt)
    return cnt, worst_fwd


def test_filters_cpr43(engine, cpr_models, cpr_oracle, oracle):
    hm = synth.read_hmms(CPR_HMM)
    b = synth.make_bin('b0', hm, seed=21, n_orfs=220, tandem_prob=0.15, max_len=1200)
    cnt, worst = _check(engine, cpr_models, cpr_oracle, b, oracle)
    assert cnt[3] >= 40
    print('passed per stage', cnt, 'worst fwd err (bits)', worst)


def test_filters_long_models(engine, oracle, tmp_path):
    p = str(tmp_path / 'long.hmm')
    ms = synth.make_model_db(p, CPR_HMM, [33, 64, 65, 300, 700, 1100], seed=5)
    ohf = oracle.HmmFile(p)
    models = engine.load_models(p)
    b = synth.make_bin('b1', ms, seed=6, n_orfs=40, copies=(1,), max_len=2500, split_prob=0.0)
    _check(engine, models, ohf, b, oracle)
    models.close()
