"""Full-size properties (BASELINE.json configs[2] stand-in: 3 Mb bins x the 5,000-model database bench.py uses).

The oracle cannot score 58 M pairs in a test, so at this size the checks are properties that do not need it:
  * replica invariance -- the database holds each of the 43 real models 116 times under different names, packed into
    different SSV tiles / slots and interleaved with other models in every work list; every replica must report exactly
    the rows of the original (same ORFs, coordinates, float bits of every score), and the per-bin E-values (Z = ORFs of
    the bin, domZ = hits of that model in that bin) must agree too;
  * batch-split invariance -- bins searched together give the rows of the bins searched one at a time;
  * determinism -- the same search twice gives the same bytes (queues are filled by atomics and re-sorted);
  * the cascade passes the fractions its thresholds define (MSV P <= 0.02 on a null-dominated workload).
The same 43 models x the same ORFs against the oracle is what tests/test_search_gpu.py does at small size."""
import numpy as np
import pytest

import bench
from tools import synth

pytestmark = pytest.mark.gpu

FIELDS = ('seq', 'tlen', 'qlen', 'dom', 'ndom', 'hmm_from', 'hmm_to', 'ali_from', 'ali_to', 'env_from', 'env_to',
          'full_score', 'full_bias', 'dom_score', 'dom_bias', 'acc', 'full_evalue', 'c_evalue', 'i_evalue')


@pytest.fixture(scope='module')
def big(engine):
    db_path = bench.model_db(bench.N_MODELS)
    models = engine.load_models(db_path)
    hm = synth.read_hmms(bench.CPR)
    bins = [synth.make_bin('fs%d' % i, hm, seed=4200 + i, n_orfs=bench.CFG[3]['orfs'], copies=(0, 1, 1, 1, 2)) for i in range(2)]
    yield models, bins
    models.close()


def _search(engine, models, bins):
    res = np.concatenate([b.residues for b in bins])
    lens = np.concatenate([np.diff(b.offsets) for b in bins])
    off = np.zeros(len(lens) + 1, np.int64)
    off[1:] = np.cumsum(lens)
    binof = np.concatenate([np.full(b.nseq, i, np.int32) for i, b in enumerate(bins)])
    db = engine.seqdb(res, off, binof, len(bins))
    hits = engine.search(models, db)
    st = engine.stats()
    db.close()
    return hits, st


def test_fullsize_replicas_split_determinism(engine, big):
    models, bins = big
    hits, st = _search(engine, models, bins)
    assert st.n_pairs == 2 * bench.CFG[3]['orfs'] * models.n
    # cascade fractions: the MSV filter lets through P <= 0.02 of a null-dominated workload (plus the planted homologs)
    assert 0.015 < st.n_past_msv / st.n_pairs < 0.03, st.n_past_msv / st.n_pairs
    assert st.n_past_vit / st.n_pairs < 0.004 and st.n_past_fwd / st.n_pairs < 0.001
    assert len(hits) > 5000
    # determinism
    hits2, _ = _search(engine, models, bins)
    assert hits.tobytes() == hits2.tobytes()
    # replica invariance: model index = replica * 43 + base model (bench.model_db writes the 43 models round after round)
    nbase = 43
    per = {}
    for h in hits:
        per.setdefault((int(h['bin']), int(h['model'])), []).append(tuple(h[f].item() for f in FIELDS))
    nrep = models.n // nbase
    checked = 0
    for b in range(len(bins)):
        for m in range(nbase):
            ref = per.get((b, m), [])
            for r in range(1, nrep):
                if r * nbase + m >= models.n:
                    continue
                assert per.get((b, r * nbase + m), []) == ref, (b, m, r)
                checked += 1
    assert checked > 9000
    # batch-split invariance
    for bi, b in enumerate(bins):
        solo, _ = _search(engine, models, [b])
        sub = hits[hits['bin'] == bi].copy()
        sub['seq'] -= 0 if bi == 0 else bins[0].nseq
        sub['bin'] = 0
        assert solo.tobytes() == sub.tobytes(), bi
    print('full size: %d rows, %d replica groups identical, cascade %d/%d/%d/%d of %d pairs' %
          (len(hits), checked, st.n_past_msv, st.n_past_bias, st.n_past_vit, st.n_past_fwd, st.n_pairs))
