"""Bin statistics, CPU side: the oracle against the dictionaries the reference wrote for the fixture bins, and
the library's nucleotide FASTA reader (host code, no device) against the oracle's reader."""
import ast
import json
import os

import numpy as np

from conftest import GOLDEN

BS = os.path.join(GOLDEN, 'binstats')
FILES = {'bin1': 'bin1.fna', 'bin2': 'bin2.fna.gz', 'bin3': 'bin3.fna'}


def _golden_rows():
    rows = {}
    for line in open(os.path.join(BS, 'bin_stats.tsv')):
        k, v = line.rstrip('\n').split('\t', 1)
        rows[k] = v
    return rows


def test_oracle_reproduces_the_reference_dictionaries():
    from oracle import binstats_oracle as bo
    rows = _golden_rows()
    for binId, fname in FILES.items():
        scaffolds = bo.read_fasta(os.path.join(BS, 'bins', fname))
        gdir = os.path.join(BS, 'out', 'bins', binId)
        has_genes = os.path.exists(os.path.join(gdir, 'genes.gff'))
        d = bo.bin_statistics(scaffolds, os.path.join(gdir, 'genes.gff') if has_genes else None, os.path.join(gdir, 'genes.faa') if has_genes else None)
        assert str(d) == rows[binId], binId
    want = json.load(open(os.path.join(BS, 'sequence_stats.json')))['sequenceStats']['bin3.fna']
    got = bo.sequence_statistics(bo.read_fasta(os.path.join(BS, 'bins', 'bin3.fna')))
    for seqId, stats in got.items():
        for k, v in stats.items():
            assert repr(want[seqId][k]) == repr(v), (seqId, k)


def test_library_reader_matches_the_reference_reader(tmp_path):
    from oracle import binstats_oracle as bo
    from checkm_b200 import seqio
    cases = {f: os.path.join(BS, 'bins', f) for f in FILES.values()}
    odd = tmp_path / 'odd.fna'
    odd.write_bytes(b'\n>a desc\nACGT\rNNNN\r\n\n \t\n>b\nAC GT\n>a\nTTTT\nGG')      # lone CR, blanks, inner blank, repeated id, no last newline
    cases['odd'] = str(odd)
    empty = tmp_path / 'empty.fna'
    empty.write_bytes(b'')
    cases['empty'] = str(empty)
    for name, path in cases.items():
        want = bo.read_fasta(path)
        ids, data, starts, lens = seqio.scan_nt_fasta(seqio.read_bytes(path))
        assert ids == list(want.keys()), name
        assert all(s % 64 == 0 for s in starts)
        for i, s, n in zip(ids, starts, lens):
            assert data[s:s + n].tobytes().decode('latin-1') == want[i], (name, i)
            pad = (s + n + 63) // 64 * 64
            assert not data[s + n:pad].any()
