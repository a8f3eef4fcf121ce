"""Bin statistics on the device: `BinStatistics` writes, character for character, the dictionaries the
reference wrote for the fixture bins; the scan itself is held to the oracle on adversarial layouts and to a numpy
run-length restatement at 64 MB."""
import json
import os
import shutil

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

BS = os.path.join(GOLDEN, 'binstats')
FILES = [os.path.join(BS, 'bins', f) for f in ('bin1.fna', 'bin2.fna.gz', 'bin3.fna')]


def _layout(seqs, pad_byte=0):
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    padded = (lens + 63) // 64 * 64
    starts = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    data = np.full(int(padded.sum()) + 64, pad_byte, dtype=np.uint8)
    for s, at in zip(seqs, starts):
        data[at:at + len(s)] = np.frombuffer(s, dtype=np.uint8)
    return data, starts, lens


def _check_against_oracle(engine, seqs, pad_byte=0):
    from oracle import binstats_oracle as bo
    data, starts, lens = _layout(seqs, pad_byte)
    stats, cscaf, clen, _ = engine.scaffold_stats(data, starts, lens)
    for i, raw in enumerate(seqs):
        s = raw.decode('latin-1')
        a, c, g, t = bo.base_counts(s)
        want = bo.contig_lengths(s)
        assert list(stats[i, :6]) == [a, c, g, t, s.count('N'), s.count('n')], i
        assert sorted(clen[cscaf == i].tolist()) == sorted(want), (i, len(s))
        assert (stats[i, 6], stats[i, 7]) == (len(want), sum(want)), i


def test_calculate_writes_the_reference_dictionaries(engine, tmp_path):
    from checkm_b200.binStatistics import BinStatistics
    out = str(tmp_path / 'out')
    shutil.copytree(os.path.join(BS, 'out'), out)
    BinStatistics(4).calculate(FILES, out, 'bin_stats.analyze.tsv')
    got = open(os.path.join(out, 'storage', 'bin_stats.analyze.tsv')).read()
    assert got == open(os.path.join(BS, 'bin_stats.tsv')).read()
    # the file is what ResultsParser.analyseResults reads back (resultsParser.py:63,121-131)
    from checkm_b200.resultsParser import ResultsParser
    parsed = ResultsParser({}).parseBinStats(out, 'bin_stats.analyze.tsv')
    assert parsed['bin3']['# predicted genes'] == 9 and parsed['bin1']['# contigs'] == 16


def test_dictionary_methods_and_sequence_stats(engine, tmp_path):
    from checkm_b200.binStatistics import BinStatistics
    from oracle import binstats_oracle as bo
    want = json.load(open(os.path.join(BS, 'sequence_stats.json')))
    seqs = bo.read_fasta(FILES[0])
    seqs.pop('scaf_e')
    bs = BinStatistics(1)
    assert [repr(v) for v in bs.calculateGC(seqs)] == want['calculateGC_bin1_without_e']
    assert [repr(v) for v in bs.calculateSeqStats(seqs)] == want['calculateSeqStats_bin1_without_e']
    out = str(tmp_path / 'out')
    shutil.copytree(os.path.join(BS, 'out'), out)
    got = bs.sequenceStats(out, FILES[2])
    ref = want['sequenceStats']['bin3.fna']
    assert list(got.keys()) == list(ref.keys())
    for seqId in ref:
        assert {k: repr(v) for k, v in got[seqId].items()} == {k: repr(v) for k, v in ref[seqId].items()}, seqId
    assert bs.calculateGC({}) == (0.0, 0.0)


def test_scan_on_boundaries(engine):
    """Runs of N of every length 1..25 ending at every offset around the 64-byte chunk and the 16 KB tile edges."""
    rng = np.random.default_rng(5)
    seqs = []
    for edge in (64, 128, 16384, 32768, 131072, 262144):            # chunk, tile and segment edges
        for r in list(range(1, 26)) + [64, 65, 200]:
            for end in (edge - 1, edge, edge + 1, edge + 9, edge + 10):
                s = bytearray(rng.choice(np.frombuffer(b'ACGTacgtn', dtype=np.uint8), size=edge + 300).tobytes())
                s[max(end - r, 0):end] = b'N' * (end - max(end - r, 0))
                seqs.append(bytes(s))
    for n in (0, 1, 9, 10, 11, 63, 64, 65, 16383, 16384, 16385, 131071, 131072, 131073, 400000):
        seqs.append(b'N' * n)
        seqs.append(b'A' * n)
        seqs.append(b'G' * n)
        seqs.append(b'TG' * (n // 2))
        seqs.append(b'C' * n)
        seqs.append((b'N' * 10 + b'C') * (n // 11) + b'N' * (n % 11))
    _check_against_oracle(engine, seqs)
    _check_against_oracle(engine, seqs[::7], pad_byte=ord('N'))         # whatever lies in the padding is not sequence


def test_scan_random_bins(engine):
    rng = np.random.default_rng(11)
    seqs = []
    for _ in range(1500):                                                # more scaffolds than CTAs: the work queue
        n = int(rng.choice([0, 1, 50, 500, 3000, 20000, 70000], p=[0.02, 0.03, 0.2, 0.35, 0.3, 0.08, 0.02]))
        s = rng.choice(np.frombuffer(b'ACGTUacgtuNnRYX', dtype=np.uint8), size=n,
                       p=[0.2, 0.2, 0.2, 0.2, 0.01, 0.03, 0.03, 0.03, 0.03, 0.01, 0.02, 0.01, 0.01, 0.01, 0.01]).copy()
        for _ in range(int(rng.integers(0, 6))):
            if n > 0:
                at, r = int(rng.integers(0, n)), int(rng.geometric(0.08))
                s[at:at + r] = ord('N')
        seqs.append(s.tobytes())
    _check_against_oracle(engine, seqs)


def test_more_contigs_than_first_allowed_for(engine):
    s = (b'A' + b'N' * 10) * 6000 + b'ACGT'
    _check_against_oracle(engine, [s, b'ACGT' * 100])


def test_scan_64mb_against_run_lengths(engine):
    """One 48 MB scaffold and many small ones; expectation from numpy run-length arithmetic."""
    rng = np.random.default_rng(3)
    sizes = [48 << 20] + [int(v) for v in rng.integers(1000, 200000, size=160)]
    seqs = []
    for n in sizes:
        s = rng.choice(np.frombuffer(b'ACGTN', dtype=np.uint8), size=n, p=[0.25, 0.25, 0.25, 0.2499, 0.0001]).copy()
        for at in rng.integers(0, n, size=max(n // 20000, 1)):
            s[at:at + int(rng.integers(1, 40))] = ord('N')
        seqs.append(s)
    data, starts, lens = _layout([s.tobytes() for s in seqs])
    stats, cscaf, clen, ms = engine.scaffold_stats(data, starts, lens)
    print('scan of %.1f MB: %.3f ms = %.0f GB/s' % (lens.sum() / 1e6, ms, lens.sum() / ms / 1e6))
    for i, s in enumerate(seqs):
        isn = (s == ord('N')).astype(np.int8)
        assert list(stats[i, :6]) == [int((s == ord(ch)).sum()) for ch in 'ACGT'] + [int(isn.sum()), 0]
        edges = np.flatnonzero(np.diff(np.concatenate([[0], isn, [0]])))
        run_start, run_end = edges[0::2], edges[1::2]
        long_runs = (run_end - run_start) >= 10
        cuts = np.concatenate([[0], run_end[long_runs], [len(s)]])         # contig k = bytes cuts[k]..cuts[k+1] that are not N
        nonn = np.concatenate([[0], np.cumsum(1 - isn)])
        want = np.diff(nonn[cuts])
        want = np.sort(want[want > 0])
        assert np.array_equal(np.sort(clen[cscaf == i]), want), i
