"""CPU restatement of CheckM's bin statistics (checkm/binStatistics.py:99-139,176-253) -- TEST INFRASTRUCTURE ONLY.

Only tests/ and bench legs that time a CPU baseline may import this; the product (checkm_b200/binStatistics.py) scans the
bases on the device and never comes here.  Pinned: tests/test_binstats_cpu.py holds it to the dictionaries the reference's own
BinStatistics wrote for the fixture bins (tests/golden/binstats/, made by tests/golden/make_binstats_goldens.py).

Plain string operations, one function per reference function."""
import gzip
import math

import numpy as np

CONTIG_BREAK = 'N' * 10          # checkm/defaultValues.py:100
GC_STD_MIN_LEN = 1000            # checkm/defaultValues.py:104


def read_fasta(path):
    """util/seqUtils.py:180-211: text-mode lines, blank ones skipped, id = first token of the header, the last character of
    every sequence line taken to be its newline."""
    handle = gzip.open(path, 'rt') if path.endswith('.gz') else open(path, 'rt')
    pieces, current = {}, None
    with handle:
        for line in handle:
            if line.strip() == '':
                continue
            if line.startswith('>'):
                current = line[1:].split(None, 1)[0]
                pieces[current] = []
            else:
                pieces[current].append(line[:-1])
    return {k: ''.join(v) for k, v in pieces.items()}


def base_counts(seq):
    """util/seqUtils.py:279-286."""
    s = seq.upper()
    return s.count('A'), s.count('C'), s.count('G'), s.count('T') + s.count('U')


def contig_lengths(scaffold):
    """binStatistics.py:214-222: pieces between CONTIG_BREAKs, their other N's not counted, empty pieces dropped."""
    out = []
    for piece in scaffold.split(CONTIG_BREAK):
        n = len(piece) - piece.count('N')
        if n > 0:
            out.append(n)
    return out


def n50(lengths):
    """util/seqUtils.py:289-301."""
    half = sum(lengths) / 2.0
    run = 0
    for v in sorted(lengths, reverse=True):
        run += v
        if run >= half:
            return v


def gc_stats(scaffolds):
    """binStatistics.py:176-206."""
    gc_all = at_all = 0
    per_seq = []
    for seq in scaffolds.values():
        a, c, g, t = base_counts(seq)
        gc_all += g + c
        at_all += a + t
        frac = float(g + c) / (g + c + a + t) if (g + c + a + t) > 0 else 0.0
        if len(seq) > GC_STD_MIN_LEN:
            per_seq.append(frac)
    GC = float(gc_all) / (gc_all + at_all) if (gc_all + at_all) > 0 else 0.0
    var = np.mean([(x - GC) ** 2 for x in per_seq]) if len(per_seq) > 1 else 0
    return GC, math.sqrt(var)


def coding_bases(gff_path):
    """prodigal.py:202-273: translation table, and per sequence the bases covered by a gene (a 0/1 mask, so overlaps count once)."""
    table, spans, n = None, {}, 0
    for line in open(gff_path):
        if line.startswith('# Model Data') and not table:
            for token in line.split(';'):
                if 'transl_table' in token:
                    table = int(token[token.find('=') + 1:])
        if line[0] == '#' or line.strip() == '"':
            continue
        f = line.split('\t')
        if f[0] not in spans:
            n = 0
            spans[f[0]] = {}
        spans[f[0]][n] = (int(f[3]), int(f[4]))
        n += 1
    covered = {}
    for seq_id, genes in spans.items():
        mask = np.zeros(max(e for _, e in genes.values()))
        for s, e in genes.values():
            mask[s - 1:e] = 1
        covered[seq_id] = mask.sum()
    return table, covered


def bin_statistics(scaffolds, gff_path=None, aa_path=None):
    """binStatistics.py:99-139: the dictionary written to bin_stats.analyze.tsv for one bin."""
    d = {}
    d['GC'], d['GC std'] = gc_stats(scaffolds)
    scaffold_lens = [len(s) for s in scaffolds.values()]
    contigs = [n for s in scaffolds.values() for n in contig_lengths(s)]
    d['Genome size'] = sum(scaffold_lens)
    d['# ambiguous bases'] = sum(s.count('N') + s.count('n') for s in scaffolds.values())
    d['# scaffolds'] = len(scaffolds)
    d['# contigs'] = len(contigs)
    d['Longest scaffold'] = max(scaffold_lens)
    d['Longest contig'] = max(contigs)
    d['N50 (scaffolds)'] = n50(scaffold_lens)
    d['N50 (contigs)'] = n50(contigs)
    d['Mean scaffold length'] = float(np.mean(scaffold_lens))
    d['Mean contig length'] = float(np.mean(contigs))
    if gff_path is None:
        d['Coding density'], d['Translation table'], d['# predicted genes'] = -1, -1, -1
    else:
        table, covered = coding_bases(gff_path)
        total = 0
        for seq_id in scaffolds:
            total += covered.get(seq_id, 0)
        d['Coding density'] = float(total) / d['Genome size']
        d['Translation table'] = table
        d['# predicted genes'] = len(read_fasta(aa_path))
    return d


def sequence_statistics(scaffolds):
    """binStatistics.py:263-275 (without the ORF columns): GC, length, contig total and count per sequence."""
    out = {}
    for seq_id, seq in scaffolds.items():
        a, c, g, t = base_counts(seq)
        lens = contig_lengths(seq)
        out[seq_id] = {'GC': float(g + c) / (g + c + a + t) if (g + c + a + t) > 0 else 0.0, 'Length': len(seq),
                       'Total contig length': sum(lens), '# contigs': len(lens)}
    return out
