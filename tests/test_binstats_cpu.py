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


def test_reader_fuzz_against_the_reference_reader(tmp_path):
    """Random mixtures of line ends ("\\n", "\\r\\n", lone "\\r"), blank and whitespace-only lines, repeated ids, descriptions, and a
    last line with or without a newline: the library's scan and the oracle's restatement of readFasta agree byte for byte."""
    from oracle import binstats_oracle as bo
    from checkm_b200 import seqio
    rng = np.random.default_rng(2024)
    alphabet = np.frombuffer(b'ACGTNacgtnRYKM-*xX', dtype=np.uint8)
    for case in range(60):
        eols = [b'\n', b'\r\n', b'\r'] if case % 3 == 0 else ([b'\n'] if case % 3 == 1 else [b'\r\n'])
        out = []
        nrec = int(rng.integers(1, 8))
        for r in range(nrec):
            name = b'id%d' % (int(rng.integers(0, 4)) if case % 5 == 0 else r)
            desc = b' some text' if rng.random() < 0.4 else b''
            out.append(b'>' + name + desc + eols[int(rng.integers(len(eols)))])
            for _ in range(int(rng.integers(0, 6))):
                kind = rng.random()
                if kind < 0.1:
                    line = b''
                elif kind < 0.2:
                    line = b' \t '[:int(rng.integers(1, 4))]
                else:
                    line = rng.choice(alphabet, size=int(rng.integers(1, 90))).tobytes()
                out.append(line + eols[int(rng.integers(len(eols)))])
        text = b''.join(out)
        if case % 4 == 0:
            text = text.rstrip(b'\r\n')                       # no final newline: the reference drops the last character
        path = tmp_path / ('f%d.fna' % case)
        path.write_bytes(text)
        want = bo.read_fasta(str(path))
        ids, data, starts, lens = seqio.scan_nt_fasta(text)
        assert ids == list(want.keys()), (case, text)
        for i, s, n in zip(ids, starts, lens):
            assert data[s:s + n].tobytes().decode('latin-1') == want[i], (case, i, text)


def test_gene_features_and_n50_against_the_oracle(tmp_path):
    from oracle import binstats_oracle as bo
    from checkm_b200.binStatistics import _GeneFeatures, _n50
    rng = np.random.default_rng(7)
    for case in range(20):
        path = tmp_path / ('g%d.gff' % case)
        with open(path, 'w') as f:
            f.write('##gff-version  3\n')
            for s in range(int(rng.integers(1, 5))):
                sid = 'contig_%d' % s
                f.write('# Sequence Data: seqnum=%d;seqlen=100000;seqhdr="%s"\n' % (s + 1, sid))
                f.write('# Model Data: version=Prodigal.v2.6.3;run_type=Single;model="Ab initio";gc_cont=50.00;transl_table=%d;uses_sd=1\n' % (11 if case % 2 else 4))
                for g in range(int(rng.integers(0, 30))):
                    a = int(rng.integers(1, 5000))
                    b = a + int(rng.integers(0, 900))
                    f.write('%s\tProdigal_v2.6.3\tCDS\t%d\t%d\t10.0\t%s\t0\tID=%d_%d;partial=00\n' % (sid, a, b, '+-'[g % 2], s + 1, g + 1))
        table, covered = bo.coding_bases(str(path))
        mine = _GeneFeatures(str(path))
        assert mine.translationTable == table
        for sid, n in covered.items():
            assert mine.codingBases(sid) == float(n), (case, sid)
        assert mine.codingBases('absent') == 0.0
    for _ in range(200):
        lens = [int(v) for v in rng.integers(1, 1000, size=int(rng.integers(1, 40)))]
        assert _n50(lens) == bo.n50(list(lens))
