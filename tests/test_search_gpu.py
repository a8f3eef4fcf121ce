"""Full-pipeline parity on the GPU: ckm_search (all stages) against the oracle's orc_search on the same seeded bins.
Hit table rows (target, query, domain index, hmm/ali/env coordinates) must be identical; bit scores within 1e-3 bits
(north_star tolerance); E-values to the corresponding relative precision."""
import numpy as np
import pytest

from tools import synth
from conftest import CPR_HMM

pytestmark = pytest.mark.gpu


def compare(rows, hits, exact=True):
    """exact=True (models with a blocked class, M <= 1024: the oracle evaluates the fp32 sums in the engine's canonical order):
    every float of the row must be the oracle's float, bit for bit.  exact=False (the chunked kernels, whose row sums are
    associated differently): scores within the north-star tolerance of 1e-3 bits plus the fp32 noise of the bias sums."""
    key_o = [(r['model'], r['seqidx'], r['dom'], r['ndom'], r['hmm_from'], r['hmm_to'], r['ali_from'], r['ali_to'], r['env_from'], r['env_to'], r['tlen']) for r in rows]
    key_g = [(int(h['model']), int(h['seq']), int(h['dom']), int(h['ndom']), int(h['hmm_from']), int(h['hmm_to']), int(h['ali_from']), int(h['ali_to']), int(h['env_from']), int(h['env_to']), int(h['tlen'])) for h in hits]
    assert len(key_o) == len(key_g), (len(key_o), len(key_g), sorted(set(key_o) ^ set(key_g))[:10])
    assert key_o == key_g, [(a, b) for a, b in zip(key_o, key_g) if a != b][:10]
    worst = 0.0
    for r, h in zip(rows, hits):
        slack = 0.0 if exact else 1e-3 + 1e-5 * abs(float(r['full_score'])) + 1e-4 * abs(float(r['full_bias']))
        for a, b in ((r['full_score'], h['full_score']), (r['dom_score'], h['dom_score']), (r['full_bias'], h['full_bias']), (r['dom_bias'], h['dom_bias'])):
            d = abs(float(np.float32(a)) - float(b))
            worst = max(worst, d)
            assert d <= slack, (r, h)
        assert abs(float(np.float32(r['acc'])) - float(h['acc'])) <= (1e-6 if exact else 1e-3)
        for a, b in ((r['full_E'], h['full_evalue']), (r['c_E'], h['c_evalue']), (r['i_E'], h['i_evalue'])):
            # E = exp(lnP) * Z in double on both sides; the two exp() implementations may differ in the last place
            assert abs(np.log(max(a, 1e-300)) - np.log(max(float(b), 1e-300))) <= 1e-12 + slack * 1.5, (a, b)
    return worst


def run_bin(engine, models, ohf, oracle, b, **kw):
    db = engine.seqdb(b.residues, b.offsets)
    hits = engine.search(models, db, **kw)
    rp = oracle.search(ohf, b.residues, b.offsets, nthreads=8, models=kw.get('model_idx'))
    rows = oracle.hits_table(rp)
    if kw.get('model_idx') is not None:
        for r in rows:
            r['model'] = kw['model_idx'][r['model']]
    nclu = sum(1 for h in range(rp.contents.nhits) if rp.contents.hits[h].nclustered > 0)
    oracle.free_results(rp)
    db.close()
    return rows, hits, nclu


def test_search_single_domain(engine, cpr_models, cpr_oracle, oracle):
    hm = synth.read_hmms(CPR_HMM)
    b = synth.make_bin('b0', hm, seed=31, n_orfs=300, max_len=1200)
    rows, hits, nclu = run_bin(engine, cpr_models, cpr_oracle, oracle, b)
    assert len(rows) >= 30
    worst = compare(rows, hits)
    st = engine.stats()
    print('rows', len(rows), 'worst score diff (bits)', worst, 'clustered regions in oracle', nclu,
          'stats', st.n_past_msv, st.n_past_bias, st.n_past_vit, st.n_past_fwd, st.n_hits_seq, st.n_domains, st.n_reported)


def test_search_tandem_domains(engine, cpr_models, cpr_oracle, oracle):
    """Two copies of a family inside one ORF: multi-domain regions resolved by the stochastic-trace ensemble."""
    hm = synth.read_hmms(CPR_HMM)
    b = synth.make_bin('b1', hm, seed=32, n_orfs=200, tandem_prob=0.5, max_len=1500)
    rows, hits, nclu = run_bin(engine, cpr_models, cpr_oracle, oracle, b)
    assert nclu >= 1
    print('tandem: rows', len(rows), 'worst', compare(rows, hits), 'clustered', nclu)


def test_search_subset_and_order(engine, cpr_models, cpr_oracle, oracle):
    hm = synth.read_hmms(CPR_HMM)
    b = synth.make_bin('b2', hm, seed=33, n_orfs=150, max_len=900)
    idx = [5, 0, 17, 42, 9]
    rows, hits, _ = run_bin(engine, cpr_models, cpr_oracle, oracle, b, model_idx=idx)
    compare(rows, hits)
    assert [int(h['model']) for h in hits] == [r['model'] for r in rows]


def test_search_two_bins_Z(engine, cpr_models, cpr_oracle, oracle):
    """E-values use Z = number of ORFs of the hit's own bin."""
    hm = synth.read_hmms(CPR_HMM)
    b1 = synth.make_bin('x', hm, seed=34, n_orfs=120, max_len=800)
    b2 = synth.make_bin('y', hm, seed=35, n_orfs=260, max_len=800)
    res = np.concatenate([b1.residues, b2.residues])
    off = np.concatenate([b1.offsets, b2.offsets[1:] + b1.offsets[-1]])
    binof = np.concatenate([np.zeros(b1.nseq, np.int32), np.ones(b2.nseq, np.int32)])
    db = engine.seqdb(res, off, binof, 2)
    hits = engine.search(cpr_models, db)
    db.close()
    for bi, b in enumerate((b1, b2)):
        rp = oracle.search(cpr_oracle, b.residues, b.offsets, nthreads=8)
        rows = oracle.hits_table(rp)
        oracle.free_results(rp)
        sub = hits[hits['bin'] == bi].copy()
        sub['seq'] -= 0 if bi == 0 else b1.nseq
        compare(rows, sub)


def test_search_with_chunked_kernels(engine, cpr_models, cpr_oracle, oracle, monkeypatch):
    """CKM_BLK=0: MSV, Viterbi, Forward and the domain stage all on the chunked shared-memory kernels (production: M > 1024)."""
    monkeypatch.setenv('CKM_BLK', '0')
    hm = synth.read_hmms(CPR_HMM)
    b = synth.make_bin('b5', hm, seed=41, n_orfs=160, max_len=900, tandem_prob=0.2)
    rows, hits, _ = run_bin(engine, cpr_models, cpr_oracle, oracle, b)
    assert len(rows) >= 20
    compare(rows, hits, exact=False)


def test_queue_overflow_is_retried(engine, cpr_models, cpr_oracle, oracle):
    """A candidate-dense batch (every ORF carries a homolog of every queried model) overflows the default SSV/MSV queues,
    which hold a sixth / a twelfth of the pairs (+64k): the cascade is re-run with larger queues instead of failing."""
    hm = synth.read_hmms(CPR_HMM)
    rng = np.random.default_rng(9)
    idx = [0, 1, 2]
    distinct = []
    for _ in range(200):
        parts = []
        for m in idx:
            parts += [synth.emit_homolog(hm[m], rng, sharpen=0.6), rng.choice(20, size=12, p=synth.BG).astype(np.uint8)]
        distinct.append(np.concatenate(parts + [np.array([27], np.uint8)]))
    n = 40000                                  # 120,000 pairs, all past every filter; queues: 85,536 and 75,536
    seqs = [distinct[i % 200] for i in range(n)]
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    res = np.concatenate(seqs)
    db = engine.seqdb(res, off)
    hits = engine.search(cpr_models, db, model_idx=idx)
    st = engine.stats()
    db.close()
    assert st.n_queue_retries >= 1
    assert st.n_past_fwd == 3 * n
    # the first 200 ORFs against the oracle (same Z is not needed for coordinates)
    rp = oracle.search(cpr_oracle, res[:off[200]], off[:201], nthreads=8, models=idx)
    rows = [r for r in oracle.hits_table(rp)]
    oracle.free_results(rp)
    got = sorted((int(h['model']), int(h['seq']), int(h['hmm_from']), int(h['hmm_to']), int(h['ali_from']), int(h['ali_to'])) for h in hits if h['seq'] < 200)
    want = sorted((idx[r['model']], r['seqidx'], r['hmm_from'], r['hmm_to'], r['ali_from'], r['ali_to']) for r in rows)
    assert got == want and len(want) >= 600


def test_search_wide_classes_and_long_models(engine, oracle, tmp_path):
    """Lane-block classes the fixture's models do not reach (Q = 24: M = 700, Q = 32: M = 1000): bit-exact like the rest.
    M = 1100 and 2500 have no blocked class (chunked kernels; M = 2500 is a chained SSV tile): coordinates exact, scores within the
    stated tolerance."""
    p = str(tmp_path / 'wide.hmm')
    ms = synth.make_model_db(p, CPR_HMM, [700, 1000, 1100, 2500], seed=21)
    ohf = oracle.HmmFile(p)
    models = engine.load_models(p)
    b = synth.make_bin('w', ms, seed=22, n_orfs=40, copies=(1, 2), max_len=3000, split_prob=0.0, tandem_prob=0.2)
    db = engine.seqdb(b.residues, b.offsets)
    hits = engine.search(models, db)
    db.close()
    rp = oracle.search(ohf, b.residues, b.offsets, nthreads=8)
    rows = oracle.hits_table(rp)
    oracle.free_results(rp)
    assert len(rows) == len(hits) and len(rows) >= 6
    blocked = [(r, h) for r, h in zip(rows, hits) if r['model'] < 2]
    chunked = [(r, h) for r, h in zip(rows, hits) if r['model'] >= 2]
    assert blocked and chunked
    compare([r for r, _ in blocked], np.array([h for _, h in blocked], dtype=hits.dtype), exact=True)
    compare([r for r, _ in chunked], np.array([h for _, h in chunked], dtype=hits.dtype), exact=False)
    models.close()


def test_search_giant_and_tiny_sequences(engine, cpr_models, cpr_oracle, oracle):
    """One 30,000-residue protein carrying many domains of many families (the worst case of every per-sequence buffer: regions,
    envelopes, ensemble traces), next to sequences of 1, 2 and 5 residues and ordinary ORFs."""
    hm = synth.read_hmms(CPR_HMM)
    b = synth.make_bin('g', hm, seed=41, n_orfs=120, max_len=1500, tandem_prob=0.2)
    rng = np.random.default_rng(42)
    parts = []
    planted = [i for i in range(b.nseq) if b.offsets[i + 1] - b.offsets[i] > 150]
    while sum(len(p) for p in parts) < 30000:
        parts.append(b.seq(int(rng.choice(planted))))
        parts.append(rng.integers(0, 20, size=int(rng.integers(5, 120))).astype(np.uint8))
    giant = np.concatenate(parts)[:30000]
    tiny = [np.array([10], dtype=np.uint8), np.array([0, 19], dtype=np.uint8), np.array([3, 3, 3, 3, 3], dtype=np.uint8)]
    seqs = [b.seq(i) for i in range(40)] + [giant] + tiny + [b.seq(i) for i in range(40, 60)]
    residues = np.concatenate(seqs)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    db = engine.seqdb(residues, offsets)
    hits = engine.search(cpr_models, db)
    db.close()
    rp = oracle.search(cpr_oracle, residues, offsets, nthreads=8)
    rows = oracle.hits_table(rp)
    oracle.free_results(rp)
    on_giant = [r for r in rows if r['seqidx'] == 40]
    assert len(on_giant) >= 20, len(on_giant)
    print('giant sequence: %d domain rows of %d families; worst score diff %g' % (len(on_giant), len({r['model'] for r in on_giant}), compare(rows, hits)))


def test_region_with_more_domains_than_slots(engine, cpr_models, cpr_oracle, oracle):
    """150 internal fragments of a 57-position family back to back: two regions of ambiguous boundaries that the trace ensemble
    resolves into 150 domains -- more than the 32 slots, the 4,096 sampled segments and the 64 segments per trace a region
    starts with.  The regions report what they need and the domain phase is repeated; the table equals the oracle's, which
    has no such limits."""
    hm = synth.read_hmms(CPR_HMM)
    fam = min(hm, key=lambda h: h.M)
    rng = np.random.default_rng(103)
    repeats = np.concatenate([synth.emit_homolog(fam, rng, k_from=int(rng.integers(20, 25)), k_to=int(rng.integers(45, 50)), sharpen=0.6)
                              for _ in range(150)])
    b = synth.make_bin('r', hm, seed=78, n_orfs=120, max_len=900)
    seqs = [b.seq(i) for i in range(15)] + [repeats] + [b.seq(i) for i in range(15, 30)]
    residues = np.concatenate(seqs)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    db = engine.seqdb(residues, offsets)
    hits = engine.search(cpr_models, db)
    st = engine.stats()
    db.close()
    rp = oracle.search(cpr_oracle, residues, offsets, nthreads=8)
    rows = oracle.hits_table(rp)
    regions = [(rp.contents.hits[h].nregions, rp.contents.hits[h].nclustered, rp.contents.hits[h].nenvelopes) for h in range(rp.contents.nhits)
               if rp.contents.hits[h].nenvelopes > 64]
    oracle.free_results(rp)
    print('repeat protein (regions, clustered, envelopes):', regions, 'passes repeated:', st.n_queue_retries, 'worst score diff', compare(rows, hits))
    assert regions and regions[0][2] > 64 * regions[0][0] // 2
    assert st.n_queue_retries >= 1


def test_search_model_beyond_the_ssv_tiles(engine, oracle, tmp_path):
    """M = 3,300: the chain of SSV tiles of such a model does not fit shared memory, so it gets none and every pair of it goes
    straight to the exact MSV kernel (ssv_bypass_kernel); the rest of the cascade runs on the chunked kernels.  Searched next to
    an ordinary model, as the whole database and as a per-call subset; a model beyond the DP rows (M > 4,608) is refused by name."""
    p = str(tmp_path / 'long.hmm')
    ms = synth.make_model_db(p, CPR_HMM, [300, 3300], seed=31)
    ohf = oracle.HmmFile(p)
    models = engine.load_models(p)
    b = synth.make_bin('l', ms, seed=32, n_orfs=40, copies=(2, 3), max_len=4000, split_prob=0.0, tandem_prob=0.0)
    db = engine.seqdb(b.residues, b.offsets)
    for idx in (None, [1]):
        hits = engine.search(models, db) if idx is None else engine.search(models, db, model_idx=idx)
        rp = oracle.search(ohf, b.residues, b.offsets, nthreads=8, models=idx)
        rows = oracle.hits_table(rp)
        oracle.free_results(rp)
        if idx is not None:
            for r in rows:
                r['model'] = idx[r['model']]
        assert any(r['model'] == 1 for r in rows)
        short = [(r, h) for r, h in zip(rows, hits) if r['model'] == 0]
        long_ = [(r, h) for r, h in zip(rows, hits) if r['model'] == 1]
        assert len(rows) == len(hits)
        compare([r for r, _ in short], np.array([h for _, h in short], dtype=hits.dtype), exact=True)
        compare([r for r, _ in long_], np.array([h for _, h in long_], dtype=hits.dtype), exact=False)
    db.close()
    models.close()
    q = str(tmp_path / 'giant.hmm')
    synth.make_model_db(q, CPR_HMM, [4700], seed=33)
    from checkm_b200._lib import CkmError
    with pytest.raises(CkmError, match='4608'):
        engine.load_models(q)
