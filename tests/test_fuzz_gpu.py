"""Differential fuzzing of the whole search against the oracle on inputs the i.i.d.-background bins do not reach:
compositionally biased ORFs (skewed Dirichlet compositions: the regime the bias filter and the null2 correction exist for),
low-complexity repeats and homopolymer runs, ORFs dense in degenerate codes (B J Z O U X), internal stop symbols, homologs
planted inside biased flanks, random model subsets.  Every row of the device's domtblout must equal the oracle's as text
and every float bit for bit (tests/test_text_parity_gpu.py holds the comparison); the bar is 0 mismatches."""
import numpy as np
import pytest

from conftest import CPR_HMM
from tools import synth
from test_text_parity_gpu import _both_tables, _compare

pytestmark = pytest.mark.gpu


def _rebuild(b, seqs):
    chunks = [np.concatenate([np.asarray(s, dtype=np.uint8), np.array([27], dtype=np.uint8)]) for s in seqs]
    offsets = np.zeros(len(chunks) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(c) for c in chunks])
    return synth.Bin(b.bin_id, np.concatenate(chunks), offsets, b.names, b.descs, b.planted)


def fuzz_bin(tag, hm, seed, n_orfs=220):
    rng = np.random.default_rng(seed)
    b = synth.make_bin(tag, hm, seed=seed, n_orfs=n_orfs, max_len=int(rng.integers(600, 2200)), tandem_prob=float(rng.uniform(0, 0.4)),
                       split_prob=float(rng.uniform(0, 0.3)), sharpen=float(rng.uniform(0, 0.5)), copies=(0, 1, 1, 2),
                       degenerate_prob=float(rng.choice([0.0, 0.002, 0.02])))
    planted = {o for _, o, _ in b.planted} | {o + 1 for _, o, kind in b.planted if kind == 'split'}
    seqs = []
    for i in range(b.nseq):
        s = b.seq(i)[:-1].copy()                       # without the trailing '*'
        L = len(s)
        kind = rng.random()
        keep = i in planted and rng.random() < 0.6     # most homologs keep their sequence, flanks may still be rewritten below
        if not keep:
            if kind < 0.25:                            # a skewed composition over the whole ORF
                comp = rng.dirichlet(np.full(20, float(rng.choice([0.05, 0.2, 0.6]))))
                s = rng.choice(20, size=L, p=comp).astype(np.uint8)
            elif kind < 0.40:                          # a short motif repeated over a window
                motif = rng.integers(0, 20, size=int(rng.integers(1, 8))).astype(np.uint8)
                w0 = int(rng.integers(0, max(1, L - 20)))
                w1 = min(L, w0 + int(rng.integers(20, 300)))
                s[w0:w1] = np.resize(motif, w1 - w0)
            elif kind < 0.47:                          # dense in degenerate codes
                m = rng.random(L) < 0.3
                s[m] = rng.choice([21, 22, 23, 24, 25, 26], size=int(m.sum()))
        elif rng.random() < 0.5 and L > 80:            # a homolog between biased flanks
            comp = rng.dirichlet(np.full(20, 0.1))
            f = int(rng.integers(10, 40))
            s[:f] = rng.choice(20, size=f, p=comp)
            s[L - f:] = rng.choice(20, size=f, p=comp)
        if rng.random() < 0.03 and L > 10:             # an internal stop symbol
            s[int(rng.integers(1, L - 1))] = 27
        seqs.append(s)
    return _rebuild(b, seqs)


@pytest.mark.parametrize('seed', [7001, 7002, 7003, 7004, 7005, 7006, 7007, 7008])
def test_fuzzed_bins_equal_the_oracle(seed, engine, cpr_models, cpr_oracle, oracle, tmp_path):
    hm = synth.read_hmms(CPR_HMM)
    rng = np.random.default_rng(seed + 1)
    b = fuzz_bin('z%d' % seed, hm, seed)
    idx = None
    if seed % 2 == 0:                                  # every other case searches a random subset of the models
        idx = sorted(int(x) for x in rng.choice(len(hm), size=int(rng.integers(5, 30)), replace=False))
    g, o, hits, rows = _both_tables(engine, cpr_models, cpr_oracle, oracle, b, tmp_path, 'fuzz%d' % seed, model_idx=idx)
    stats = []
    bad = _compare(g, o, hits, rows, 'fuzz %d (%d ORFs, %s models)' % (seed, b.nseq, 'all' if idx is None else len(idx)), stats)
    assert not bad, bad[:3]
