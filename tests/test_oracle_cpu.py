"""The oracle (oracle/hmmer_oracle.c) against everything that can pin it without a HMMER binary:
the reference's HMM fixture (parser KAT), the STATS lines real HMMER 3.1b2 calibrated into that fixture (score
statistics), internal identities (SSV == MSV when J is idle, Forward == Backward), and its own frozen output."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import CPR_HMM, GOLDEN
from checkm_b200 import synth


def test_parser_kat(cpr_oracle):
    hf = cpr_oracle
    assert hf.n == 43 and sum(h.M for h in hf.headers) == 8926
    h = hf.headers[0]
    assert (h.name, h.acc, h.M) == (b'Ribosomal_L23', b'PF00276.21', 86)
    assert [round(x, 5) for x in h.evparam] == [-9.2208, 0.71847, -10.0492, 0.71847, -3.8645, 0.71847]
    assert abs(h.ga[0] - 30.8) < 1e-5 and abs(h.nc[1] - 30.7) < 1e-5
    # first match emission of node 1 is 2.88296 nats for 'A'; B->M1 0.01076... wait that is node 1's m->m
    assert abs(-np.log(h.mat[1 * 20 + 0]) - 2.88296) < 1e-5
    assert abs(-np.log(h.t[0 * 7 + 0]) - 0.11535) < 1e-5 and h.t[0 * 7 + 6] == 0.0       # '*' is probability 0
    assert sorted(hh.M for hh in hf.headers)[:2] == [57, 68] and max(hh.M for hh in hf.headers) == 863


@pytest.mark.parametrize('m', [0, 1, 10])
def test_msv_and_viterbi_mu_match_hmmer_calibration(cpr_oracle, oracle, m):
    """Random i.i.d. sequences of length 200 (HMMER's calibration length): the ML Gumbel location with the model's
    lambda must reproduce the mu that HMMER 3.1b2 itself wrote into the HMM file."""
    hf = cpr_oracle
    ev = list(hf.headers[m].evparam)
    rng = np.random.default_rng(100 + m)
    L = oracle.lib()
    msv, vit = [], []
    null = L.orc_null1(200)
    for _ in range(500):
        d = rng.choice(20, size=200, p=synth.BG).astype(np.uint8)
        sc = C.c_float()
        xj = C.c_int()
        L.orc_msv(hf.prof_ptrs[m], d.ctypes.data, 200, C.byref(sc), C.byref(xj))
        msv.append((sc.value - null) / np.log(2))
        L.orc_vitfilter(hf.prof_ptrs[m], d.ctypes.data, 200, C.byref(sc))
        vit.append((sc.value - null) / np.log(2))

    def fit(x, lam):
        return -np.log(np.mean(np.exp(-lam * np.asarray(x)))) / lam
    assert abs(fit(msv, ev[1]) - ev[0]) < 0.35, (fit(msv, ev[1]), ev[0])
    assert abs(fit(vit, ev[3]) - ev[2]) < 0.35, (fit(vit, ev[3]), ev[2])


def test_ssv_equals_msv_when_j_idle_and_forward_equals_backward(cpr_oracle, oracle):
    hf = cpr_oracle
    rng = np.random.default_rng(3)
    L = oracle.lib()
    hm = synth.read_hmms(CPR_HMM)
    for m in (2, 7, 20):
        p = hf.profiles[m]
        for t in range(30):
            d = rng.choice(20, size=int(rng.integers(20, 400)), p=synth.BG).astype(np.uint8)
            if t % 5 == 0:
                d = np.concatenate([d, synth.emit_homolog(hm[m], rng)]).astype(np.uint8)
            sc, xj = oracle.msv(hf, m, d)
            xe = oracle.ssv_xe(hf, m, d)
            if xe - p.tec_b <= p.base_b and xj != 256:
                assert xj == max(xe - p.tec_b, 0)
            f, b = C.c_float(), C.c_float()
            L.orc_forward_parser(hf.prof_ptrs[m], d.ctypes.data, len(d), C.byref(f))
            L.orc_backward_parser(hf.prof_ptrs[m], d.ctypes.data, len(d), C.byref(b))
            assert abs(f.value - b.value) < 2e-3 * max(1.0, abs(f.value) / 50)


def test_oracle_output_is_frozen(cpr_oracle, oracle):
    with open(os.path.join(GOLDEN, 'oracle_hits.json')) as f:
        golden = json.load(f)
    hm = synth.read_hmms(CPR_HMM)
    for key, kw in (('seed7', dict(n_orfs=150, max_len=800, tandem_prob=0.2)), ('seed8', dict(n_orfs=150, max_len=800, split_prob=0.4))):
        b = synth.make_bin('g' + key[4:], hm, seed=int(key[4:]), **kw)
        rp = oracle.search(cpr_oracle, b.residues, b.offsets, nthreads=4)
        rows = oracle.hits_table(rp)
        oracle.free_results(rp)
        keep = ('seqidx', 'model', 'tlen', 'dom', 'ndom', 'hmm_from', 'hmm_to', 'ali_from', 'ali_to', 'env_from', 'env_to')
        got = [[int(r[k]) for k in keep] + ['%.1f' % r['full_score'], '%.1f' % r['dom_score'], '%.2g' % r['full_E']] for r in rows]
        assert got == golden[key]
        assert len(got) > 20


def test_planted_homologs_are_found(cpr_oracle, oracle):
    hm = synth.read_hmms(CPR_HMM)
    b = synth.make_bin('p', hm, seed=5, n_orfs=150, max_len=700, split_prob=0.0)
    rp = oracle.search(cpr_oracle, b.residues, b.offsets, nthreads=4)
    rows = oracle.hits_table(rp)
    oracle.free_results(rp)
    found = set((r['model'], r['seqidx']) for r in rows)
    hit = sum((m, o) in found for m, o, _ in b.planted)
    assert hit >= 0.9 * len(b.planted), (hit, len(b.planted))      # a sampled homolog can legitimately fall below E = 0.1
