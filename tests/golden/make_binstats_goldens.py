"""Fixture bins for the bin-statistics row and the dictionaries the REFERENCE writes for them.

Run in the build container (needs /root/reference):  python tests/golden/make_binstats_goldens.py
Writes tests/golden/binstats/{bins/*.fna[.gz], out/bins/<id>/genes.{gff,faa}, bin_stats.tsv, sequence_stats.json}.
The nucleotide files exercise what the device scan has to get right: runs of N of length 1..25 and 100 at the start, in the
middle and at the end of scaffolds, runs lying across the 64-byte and 16 KB boundaries of the scan, lower-case bases and n,
U, IUPAC codes, scaffolds below the 1000-base GC-std cut, CRLF line ends, a last line without newline, blank lines, a
repeated id, gzip."""
import gzip
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, 'binstats')
sys.path.insert(0, '/root/reference')


def random_dna(rng, n, gc=0.5, lower=0.0):
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    s = rng.choice(np.frombuffer(b'ACGT', dtype=np.uint8), size=n, p=p)
    if lower > 0:
        low = rng.random(n) < lower
        s = np.where(low, s + 32, s)
    return bytearray(s.astype(np.uint8).tobytes())


def put(seq, at, text):
    seq[at:at + len(text)] = text.encode()


def wrap(seq, width, eol='\n'):
    s = seq.decode()
    return eol.join(s[i:i + width] for i in range(0, len(s), width)) + eol


def main():
    rng = np.random.default_rng(20260923)
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(os.path.join(OUT, 'bins'))

    # ---- bin1: long scaffolds, runs placed on the scan's boundaries ----
    a = random_dna(rng, 70000, 0.42, lower=0.05)
    for at, r in ((0, 12), (50, 9), (60, 10), (16380, 10), (16384 * 2 - 3, 25), (16384 * 3 - 10, 10), (16384 * 3 + 1, 9),
                  (40000, 100), (40200, 19), (40300, 20), (40400, 11), (40500, 1), (50000 - 5, 10), (69990, 10)):
        put(a, at, 'N' * r)
    put(a, 45000, 'nnnnnnnnnnnn')                       # lower-case n: ambiguous, never a break
    put(a, 45100, 'NNNNNnNNNNN')                        # broken by a lower-case n: two runs of 5
    put(a, 45200, 'RYKMSWBDHVUuXx-*')
    b = random_dna(rng, 33000, 0.61)
    put(b, 32768 - 64, 'N' * 64)                        # exactly one 64-byte chunk
    put(b, 20000, 'N' * 128)
    c = random_dna(rng, 999, 0.3)                       # below the GC-std cut
    d = random_dna(rng, 1001, 0.7)
    e = bytearray(b'N' * 40)                            # nothing but a break: no contig at all
    f = random_dna(rng, 64, 0.5)
    with open(os.path.join(OUT, 'bins', 'bin1.fna'), 'w') as fh:
        fh.write('>scaf_a first scaffold\n' + wrap(a, 60))
        fh.write('\n   \n>scaf_b\n' + wrap(b, 80))
        fh.write('>scaf_c\n' + wrap(c, 70) + '>scaf_d\n' + wrap(d, 1001) + '>scaf_e\n' + wrap(e, 60) + '>scaf_f\n' + wrap(f, 61))

    # ---- bin2: gzip, CRLF, repeated id, last line without newline ----
    g = random_dna(rng, 20000, 0.55)
    put(g, 5000, 'N' * 10)
    put(g, 19999, 'N')
    h = random_dna(rng, 3000, 0.5)
    h2 = random_dna(rng, 2500, 0.35)
    text = '>c1 x\r\n' + wrap(g, 60, '\r\n') + '>dup\r\n' + wrap(h, 60, '\r\n') + '>c3\r\n' + wrap(random_dna(rng, 1500, 0.45), 60, '\r\n') + \
           '>dup again\r\n' + wrap(h2, 60, '\r\n')
    with gzip.open(os.path.join(OUT, 'bins', 'bin2.fna.gz'), 'wb') as fh:
        fh.write(text.encode())
    # ---- bin3: no final newline (the reference drops the last base), genes called ----
    k = random_dna(rng, 30000, 0.48)
    put(k, 10000, 'N' * 30)
    m = random_dna(rng, 12000, 0.52)
    with open(os.path.join(OUT, 'bins', 'bin3.fna'), 'w') as fh:
        fh.write('>k\n' + wrap(k, 60) + '>m\n' + wrap(m, 60)[:-1])
    gdir = os.path.join(OUT, 'out', 'bins', 'bin3')
    os.makedirs(gdir)
    genes = {'k': [(3, 900), (850, 2000), (2100, 2900), (2500, 2600), (9000, 11000), (29000, 30000)], 'm': [(1, 300), (301, 1200), (5000, 5001)]}
    with open(os.path.join(gdir, 'genes.gff'), 'w') as fh, open(os.path.join(gdir, 'genes.faa'), 'w') as fa:
        fh.write('##gff-version  3\n')
        for sid, spans in genes.items():
            fh.write('# Sequence Data: seqnum=1;seqlen=1;seqhdr="%s"\n' % sid)
            fh.write('# Model Data: version=Prodigal.v2.6.3;run_type=Single;model="Ab initio";gc_cont=48.00;transl_table=11;uses_sd=1\n')
            for i, (s, e2) in enumerate(spans):
                fh.write('%s\tProdigal_v2.6.3\tCDS\t%d\t%d\t50.0\t+\t0\tID=1_%d;partial=00\n' % (sid, s, e2, i + 1))
                fa.write('>%s_%d # %d # %d # 1\n%s*\n' % (sid, i + 1, s, e2, 'M' + 'A' * ((e2 - s + 1) // 3 - 1)))

    # ---- the reference ----
    os.environ['CHECKM_DATA_PATH'] = tempfile.mkdtemp()
    from checkm.binStatistics import BinStatistics
    files = [os.path.join(OUT, 'bins', f) for f in ('bin1.fna', 'bin2.fna.gz', 'bin3.fna')]
    os.makedirs(os.path.join(OUT, 'out', 'storage'))
    BinStatistics(1).calculate(files, os.path.join(OUT, 'out'), 'bin_stats.tsv')
    lines = sorted(open(os.path.join(OUT, 'out', 'storage', 'bin_stats.tsv')).read().splitlines())
    with open(os.path.join(OUT, 'bin_stats.tsv'), 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    shutil.rmtree(os.path.join(OUT, 'out', 'storage'))
    for b in ('bin1', 'bin2'):
        shutil.rmtree(os.path.join(OUT, 'out', 'bins', b), ignore_errors=True)
    seqstats = {os.path.basename(f): BinStatistics(1).sequenceStats(os.path.join(OUT, 'out'), f) for f in files[2:]}
    # the dictionary-taking methods on the first bin
    from checkm.util.seqUtils import readFasta
    seqs = readFasta(files[0])
    seqs.pop('scaf_e')
    gc = BinStatistics(1).calculateGC(seqs)
    ss = BinStatistics(1).calculateSeqStats(seqs)
    with open(os.path.join(OUT, 'sequence_stats.json'), 'w') as fh:
        json.dump({'sequenceStats': seqstats, 'calculateGC_bin1_without_e': [repr(v) for v in gc],
                   'calculateSeqStats_bin1_without_e': [repr(v) for v in ss]}, fh, indent=1, sort_keys=True)
    print(open(os.path.join(OUT, 'bin_stats.tsv')).read())


if __name__ == '__main__':
    main()
