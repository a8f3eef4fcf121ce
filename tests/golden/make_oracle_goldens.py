"""Freezes the oracle's own output on seeded inputs (regression pin of oracle/hmmer_oracle.c):
    python tests/golden/make_oracle_goldens.py
writes tests/golden/oracle_hits.json (hit-table rows + filter counters for two seeded synthetic bins)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po          # noqa: E402
from tools import synth              # noqa: E402

CPR = os.path.join(HERE, 'cpr_43_markers.hmm')


def rows_for(seed, **kw):
    hf = po.HmmFile(CPR)
    hm = synth.read_hmms(CPR)
    b = synth.make_bin('g%d' % seed, hm, seed=seed, **kw)
    rp = po.search(hf, b.residues, b.offsets, nthreads=8)
    rows = po.hits_table(rp)
    po.free_results(rp)
    keep = ('seqidx', 'model', 'tlen', 'dom', 'ndom', 'hmm_from', 'hmm_to', 'ali_from', 'ali_to', 'env_from', 'env_to')
    return [[int(r[k]) for k in keep] + ['%.1f' % r['full_score'], '%.1f' % r['dom_score'], '%.2g' % r['full_E']] for r in rows]


if __name__ == '__main__':
    out = {'seed7': rows_for(7, n_orfs=150, max_len=800, tandem_prob=0.2),
           'seed8': rows_for(8, n_orfs=150, max_len=800, split_prob=0.4)}
    with open(os.path.join(HERE, 'oracle_hits.json'), 'w') as f:
        json.dump(out, f)
    print({k: len(v) for k, v in out.items()})
