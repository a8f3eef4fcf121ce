"""Stage-1 (SSV + exact MSV) throughput probe at config-#2-like and config-#3-like shapes.  Development tool."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import synth
from checkm_b200.engine import Engine

CPR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'cpr_43_markers.hmm')


def bins(hm, nb, n_orfs, seed0=0):
    bs = [synth.make_bin('b%d' % i, hm, seed=seed0 + i, n_orfs=n_orfs) for i in range(nb)]
    res = np.concatenate([b.residues for b in bs])
    off = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(b.offsets) for b in bs]))]).astype(np.int64)
    binof = np.concatenate([np.full(b.nseq, i, np.int32) for i, b in enumerate(bs)])
    return res, off, binof


def main():
    e = Engine(0)
    print(e.device_name())
    hm = synth.read_hmms(CPR)
    t0 = time.time()
    res, off, binof = bins(hm, 8, 1900)
    print('made bins', time.time() - t0, len(off) - 1, 'seqs', len(res), 'residues')
    m = e.load_models(CPR)
    db = e.seqdb(res, off, binof, 8)
    for rep in range(3):
        t0 = time.time()
        xj = e.msv_scores(m, db)
        st = e.stats()
        print('cfg2-like: wall %.3f s  ssv %.2f ms  msv %.2f ms  cells %.3e  GCUPS(ssv) %.1f  cand %d (%.2f%%) pass %d' % (
            time.time() - t0, st.ms_ssv, st.ms_msv, st.n_cells, st.n_cells / st.ms_ssv / 1e6, st.n_ssv_cand,
            100.0 * st.n_ssv_cand / st.n_pairs, st.n_past_msv))
    # real cells (sum L*M) for the same workload
    L = np.diff(off)
    print('sum L*M = %.3e' % (float(L.sum()) * sum(h.M for h in hm)))
    m.close()
    # config-#3-like: 600 synthetic models
    rng = np.random.default_rng(0)
    lens = synth.perturbed_model_lengths(rng, 600)
    p = '/tmp/syn600.hmm'
    synth.make_model_db(p, CPR, lens, seed=1)
    m = e.load_models(p)
    db2 = e.seqdb(res[:off[3800]], off[:3801], binof[:3800], 2)
    for rep in range(3):
        t0 = time.time()
        xj = e.msv_scores(m, db2)
        st = e.stats()
        print('cfg3-like: wall %.3f s  ssv %.2f ms  msv %.2f ms  cells %.3e  GCUPS(ssv) %.1f  cand %d (%.2f%%) pass %d' % (
            time.time() - t0, st.ms_ssv, st.ms_msv, st.n_cells, st.n_cells / st.ms_ssv / 1e6, st.n_ssv_cand,
            100.0 * st.n_ssv_cand / st.n_pairs, st.n_past_msv))
    print('sum L*M = %.3e' % (float(np.diff(off[:3801]).sum()) * float(lens.sum())))


if __name__ == '__main__':
    main()
