#!/bin/bash
# Settles SURVEY.md 8(c) wherever a real HMMER exists: runs `hmmsearch` with CheckM's flags (checkm/markerGeneFinder.py:141) on
# the golden bins, and diffs its domtblout data lines against (a) the oracle's and (b) the engine's (when a GPU is present).
#   tools/validate_against_hmmer.sh [outdir]
# Exit status: 0 = every row identical as text, 1 = differences (listed), 2 = no hmmsearch on PATH.
set -u
cd "$(dirname "$0")/.." || exit 2
OUT=${1:-/tmp/ckm_validate_hmmer}
if ! command -v hmmsearch >/dev/null 2>&1; then echo "hmmsearch: not on PATH -- nothing to validate against"; exit 2; fi
mkdir -p "$OUT"
hmmsearch -h | sed -n 2p
python - "$OUT" <<'PY'
import gzip, os, subprocess, sys
sys.path.insert(0, '.')
out = sys.argv[1]
from oracle import pyoracle as po
from checkm_b200.seqio import read_fasta
hmm = 'tests/golden/cpr_43_markers.hmm'
hf = po.HmmFile(hmm)
bad = 0
def data(path):
    return [l.split() for l in open(path) if l.strip() and not l.startswith('#')]
for name in ('binA.faa', 'binB.faa.gz', 'binC.faa'):
    src = os.path.join('tests/golden/e2e/bins', name)
    faa = os.path.join(out, name.replace('.gz', ''))
    with (gzip.open(src, 'rt') if src.endswith('.gz') else open(src)) as f, open(faa, 'w') as g:
        g.write(f.read())
    ref = os.path.join(out, name + '.hmmsearch.txt')
    subprocess.check_call(['hmmsearch', '--domtblout', ref, '--noali', '--notextw', '-E', '0.1', '--domE', '0.1', '--cpu', '1', hmm, faa], stdout=subprocess.DEVNULL)
    names, descs, res, off = read_fasta(faa)
    rp = po.search(hf, res, off, nthreads=os.cpu_count() or 1)
    orc = os.path.join(out, name + '.oracle.txt')
    po.write_domtblout(rp, hf, names, descs, orc)
    po.free_results(rp)
    a, b = data(ref), data(orc)
    diff = [(x, y) for x, y in zip(a, b) if x[:22] != y[:22]]
    print('%s: hmmsearch %d rows, oracle %d rows, %d rows differ in the 22 data columns' % (name, len(a), len(b), len(diff) + abs(len(a) - len(b))))
    for x, y in diff[:5]:
        print('   hmmsearch:', ' '.join(x[:22])); print('   oracle   :', ' '.join(y[:22]))
    bad += len(diff) + abs(len(a) - len(b))
    try:
        from checkm_b200.hmmer import HMMERRunner
        gpu = os.path.join(out, name + '.gpu.txt')
        HMMERRunner().search(hmm, faa, gpu, '/dev/null', '--cpu 1 --notextw -E 0.1 --domE 0.1 --noali', False)
        c = data(gpu)
        d2 = [(x, y) for x, y in zip(a, c) if x[:22] != y[:22]]
        print('   engine: %d rows, %d differ from hmmsearch' % (len(c), len(d2) + abs(len(a) - len(c))))
        bad += len(d2) + abs(len(a) - len(c))
    except SystemExit:
        print('   engine: no GPU here, skipped')
sys.exit(1 if bad else 0)
PY
