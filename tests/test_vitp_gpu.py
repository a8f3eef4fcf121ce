"""ViterbiFilter on the GPU, every pair (ckm_viterbi_scores): the packed int16x2 kernel with its int32 redo list
(checkm_b200/csrc/kernels_vitp.cu) against the oracle's orc_vitfilter (oracle/hmmer_oracle.c) and against the int32 kernels.
Scores are a function of the final int16 xC only, so the bar is bit-exact float32 equality.

What the cases are for (the guard conditions in the header of kernels_vitp.cu):
  * background ORFs                      -> the common path: lazy-F rows, a few full D->D rows;
  * planted homologs, partial homologs   -> long full-D->D stretches, row maxima at the int16 ceiling (redo list);
  * '*', '-', '~', X, B.. anywhere       -> rows whose match emissions are all -inf / degenerate (clamped table entries);
  * L = 0, 1, 2; L not a multiple of 4/16 -> loop tails;
  * a 20,000-residue ORF x long models   -> condition C1' fails (full D->D asks for the int32 kernel), C1 still holds.
"""
import numpy as np
import pytest

from tools import synth
from conftest import CPR_HMM

pytestmark = pytest.mark.gpu


def _special_bin(hm, rng, n_bg=60):
    seqs = []
    for L in (0, 1, 2, 3, 4, 5, 15, 16, 17, 31, 33, 63, 64, 65):
        seqs.append(rng.choice(20, size=L, p=synth.BG).astype(np.uint8))
    for L in synth.random_lengths(rng, n_bg, hi=1500):
        seqs.append(rng.choice(20, size=int(L), p=synth.BG).astype(np.uint8))
    for h in hm:                                   # full and partial homologs, some flanked, some with odd symbols inside
        s = synth.emit_homolog(h, rng)
        seqs.append(s)
        k0 = int(rng.integers(1, max(2, h.M // 2)))
        part = synth.emit_homolog(h, rng, k_from=k0, k_to=min(h.M, k0 + max(10, h.M // 3)))
        fl = rng.choice(20, size=int(rng.integers(5, 200)), p=synth.BG).astype(np.uint8)
        seqs.append(np.concatenate([fl, part, fl[::-1]]))
        t = s.copy()
        for code in (27, 20, 28, 26, 21, 25):      # '*', '-', '~', 'X', 'B', 'U'
            if len(t) > 8:
                t[int(rng.integers(0, len(t)))] = code
        seqs.append(t)
    seqs.append(np.full(40, 27, np.uint8))                           # nothing but '*'
    star_first = rng.choice(20, size=120, p=synth.BG).astype(np.uint8)
    star_first[0] = 27
    seqs.append(star_first)
    star_last = rng.choice(20, size=200, p=synth.BG).astype(np.uint8)
    star_last[-1] = 27
    seqs.append(star_last)
    res = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
    off = np.zeros(len(seqs) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return res, off


def _compare(engine, models, ohf, oracle, res, off, model_idx=None, oracle_pairs=None):
    db = engine.seqdb(res, off)
    vp = engine.viterbi_scores(models, db, model_idx=model_idx)
    n_redo = engine.stats().n_vit_redo
    v32 = engine.viterbi_scores(models, db, model_idx=model_idx, int32_only=True)
    db.close()
    assert vp.tobytes() == v32.tobytes(), 'packed kernel (+redo) differs from the int32 kernels at %s' % (np.argwhere(vp != v32)[:5],)
    mi = list(range(models.n)) if model_idx is None else list(model_idx)
    nseq = len(off) - 1
    rng = np.random.default_rng(1)
    pairs = [(a, s) for a in range(len(mi)) for s in range(nseq)]
    if oracle_pairs is not None and len(pairs) > oracle_pairs:
        pairs = [pairs[i] for i in rng.choice(len(pairs), size=oracle_pairs, replace=False)]
    for a, s in pairs:
        exp = np.float32(oracle.vitfilter(ohf, mi[a], res[off[s]:off[s + 1]]))
        got = vp[a, s]
        assert got == exp or (np.isinf(got) and np.isinf(exp) and np.sign(got) == np.sign(exp)), (mi[a], s, off[s + 1] - off[s], got, exp)
    return n_redo, vp


def test_chunked_viterbi_kernel_on_short_models(engine, cpr_models, cpr_oracle, oracle):
    """ckm_viterbi_scores mode 2: every pair through vit_kernel (shared-memory int32 rows), the kernel production keeps for
    M > 1024 -- bit-identical to the lane-blocked kernels and to the oracle on the 43 short models."""
    hm = synth.read_hmms(CPR_HMM)
    res, off = _special_bin(hm, np.random.default_rng(78), n_bg=30)
    db = engine.seqdb(res, off)
    v32 = engine.viterbi_scores(cpr_models, db, int32_only=True)
    vch = engine.viterbi_scores(cpr_models, db, chunked_only=True)
    db.close()
    assert v32.tobytes() == vch.tobytes(), np.argwhere(v32 != vch)[:5]
    rng = np.random.default_rng(2)
    for _ in range(1500):
        a, s = int(rng.integers(cpr_models.n)), int(rng.integers(len(off) - 1))
        exp = np.float32(oracle.vitfilter(cpr_oracle, a, res[off[s]:off[s + 1]]))
        assert vch[a, s] == exp or (np.isinf(vch[a, s]) and np.isinf(exp) and np.sign(vch[a, s]) == np.sign(exp)), (a, s)


def test_vitp_cpr43_all_pairs(engine, cpr_models, cpr_oracle, oracle):
    hm = synth.read_hmms(CPR_HMM)
    res, off = _special_bin(hm, np.random.default_rng(77))
    n_redo, vp = _compare(engine, cpr_models, cpr_oracle, oracle, res, off, oracle_pairs=4000)
    assert np.isinf(vp).any() and (vp == np.inf).any()          # homologs overflow int16 exactly as the reference filter does
    assert 0 < n_redo < 0.2 * vp.size                           # ... and they are what the redo list is for
    print('pairs', vp.size, 'redo', n_redo)


def test_vitp_long_models_long_orf(engine, oracle, tmp_path):
    p = str(tmp_path / 'long.hmm')
    ms = synth.make_model_db(p, CPR_HMM, [64, 65, 129, 500, 897, 1000, 1024, 1100], seed=9)
    ohf = oracle.HmmFile(p)
    models = engine.load_models(p)
    rng = np.random.default_rng(5)
    seqs = [rng.choice(20, size=20000, p=synth.BG).astype(np.uint8)]
    big = rng.choice(20, size=9000, p=synth.BG).astype(np.uint8)
    hom = synth.emit_homolog(ms[5], rng, k_from=200, k_to=520)       # a weak partial hit inside a long ORF
    big[4000:4000 + len(hom)] = hom
    seqs.append(big)
    for h in ms:
        seqs.append(synth.emit_homolog(h, rng, k_from=1, k_to=max(20, h.M // 6)))
    res = np.concatenate(seqs)
    off = np.zeros(len(seqs) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    n_redo, vp = _compare(engine, models, ohf, oracle, res, off)
    assert n_redo >= len(seqs)          # at least the class-less 1100-position model goes through the redo list
    models.close()
