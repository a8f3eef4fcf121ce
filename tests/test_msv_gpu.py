"""Stage-1 parity on the GPU: SSV pre-filter + exact MSV (checkm_b200/csrc/kernels_msv.cu) against the oracle's
8-bit MSV (oracle/hmmer_oracle.c: orc_msv).  Bit-exact: every pair the real filter passes must come back with
the oracle's xJ byte, and every pair the GPU scores must agree with the oracle."""
import numpy as np
import pytest

from tools import synth
from conftest import CPR_HMM

pytestmark = pytest.mark.gpu


def _check(engine, models, ohf, b, oracle, model_idx=None):
    db = engine.seqdb(b.residues, b.offsets)
    xj = engine.msv_scores(models, db, model_idx)
    st = engine.stats()
    idx = range(models.n) if model_idx is None else model_idx
    n_pass = n_cand = 0
    for row, m in enumerate(idx):
        ev = ohf.headers[m].evparam
        for s in range(b.nseq):
            d = b.seq(s)
            if len(d) == 0:
                assert xj[row, s] == -1
                continue
            sc, oxj = oracle.msv(ohf, m, d)
            null = oracle.lib().orc_null1(len(d))
            bits = (np.float32(sc) - np.float32(null)) / np.float32(0.69314718055994529)
            y = ev[1] * (float(bits) - ev[0])
            P = 1.0 - np.exp(-np.exp(-y)) if np.isfinite(sc) else 0.0
            if xj[row, s] >= 0:
                n_cand += 1
                assert xj[row, s] == oxj, (m, s, len(d), xj[row, s], oxj)
            if P <= 0.02:
                n_pass += 1
                assert xj[row, s] == oxj, "filter-passing pair missing from the GPU candidates: %r" % ((m, s, len(d), xj[row, s], oxj),)
    db.close()
    assert st.n_past_msv == n_pass
    assert st.n_ssv_cand == n_cand
    return n_pass, n_cand, st


@pytest.mark.parametrize('resolve', ['1', '0'])
def test_msv_parity_cpr43(engine, cpr_models, cpr_oracle, oracle, resolve, monkeypatch):
    """resolve=1: pairs whose J state cannot have fired are scored in the SSV epilogue (kernels_msv.cu); resolve=0: every firing
    pair goes to the exact MSV kernels, as in round 1.  Same bytes, same pass set either way."""
    monkeypatch.setenv('CKM_SSV_RESOLVE', resolve)
    hm = synth.read_hmms(CPR_HMM)
    b = synth.make_bin('b0', hm, seed=11, n_orfs=260, tandem_prob=0.1, max_len=1500)
    n_pass, n_cand, st = _check(engine, cpr_models, cpr_oracle, b, oracle)
    assert n_pass > 40 and n_cand >= n_pass
    assert st.n_cells > 0
    assert (st.n_msv_exact < n_cand // 4) if resolve == '1' else (st.n_msv_exact == n_cand)


def test_msv_edge_lengths(engine, cpr_models, cpr_oracle, oracle):
    """Empty, 1-residue, all-'*', all-X and very long sequences."""
    rng = np.random.default_rng(5)
    seqs = [np.zeros(0, np.uint8), np.array([27], np.uint8), np.array([3], np.uint8), np.full(40, 27, np.uint8),
            np.full(33, 26, np.uint8), rng.choice(20, size=15, p=synth.BG).astype(np.uint8),
            rng.choice(20, size=16, p=synth.BG).astype(np.uint8), rng.choice(20, size=17, p=synth.BG).astype(np.uint8),
            rng.choice(20, size=6000, p=synth.BG).astype(np.uint8)]
    hm = synth.read_hmms(CPR_HMM)
    seqs.append(np.concatenate([synth.emit_homolog(hm[10], rng) for _ in range(3)]))   # strong multi-hit: J state + overflow territory
    off = np.zeros(len(seqs) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    b = synth.Bin('e', np.concatenate(seqs), off, ['s%d' % i for i in range(len(seqs))], [''] * len(seqs), [])
    _check(engine, cpr_models, cpr_oracle, b, oracle)


@pytest.mark.parametrize('policy', [None, 'auto', '8'])
def test_msv_long_models_and_subset(engine, oracle, tmp_path, policy, monkeypatch):
    """Chained tiles (M >= 64 J) and a query subset, under every tile-width policy: the default J = 32 tiles, CKM_SSV_J=auto
    (J = 4 / 8 / 16 by model length) and one fixed narrow width (the policy is read when the database is loaded)."""
    if policy is None:
        monkeypatch.delenv('CKM_SSV_J', raising=False)
    else:
        monkeypatch.setenv('CKM_SSV_J', policy)
    p = str(tmp_path / 'long.hmm')
    ms = synth.make_model_db(p, CPR_HMM, [30, 57, 130, 255, 256, 300, 511, 512, 700, 1023, 1024, 1100, 2500], seed=3)
    ohf = oracle.HmmFile(p)
    models = engine.load_models(p)
    b = synth.make_bin('b1', ms, seed=4, n_orfs=60, copies=(1,), max_len=3500, split_prob=0.0)
    _check(engine, models, ohf, b, oracle)
    _check(engine, models, ohf, b, oracle, model_idx=[12, 3, 10])
    models.close()


def test_msv_chunked_kernel_on_short_models(engine, cpr_models, cpr_oracle, oracle, monkeypatch):
    """CKM_BLK=0 sends every SSV candidate to msv_exact_kernel (shared-memory rows), which production uses only for models
    without a lane-block class: same bytes, same pass set."""
    monkeypatch.setenv('CKM_BLK', '0')
    hm = synth.read_hmms(CPR_HMM)
    b = synth.make_bin('b0', hm, seed=12, n_orfs=150, tandem_prob=0.1, max_len=1200)
    n_pass, n_cand, _ = _check(engine, cpr_models, cpr_oracle, b, oracle)
    assert n_pass > 20
