"""Protein FASTA -> digitised residue stream (the `genes.faa` reader that sat inside hmmsearch)."""
import gzip

import numpy as np

from .engine import digitize


def read_fasta(path):
    """Returns names, descriptions (header text after the first blank), and the sequences as one bytes object + offsets."""
    opener = gzip.open if path.endswith('.gz') else open
    names, descs, chunks, lens = [], [], [], []
    cur = []
    with opener(path, 'rt') as f:
        for line in f:
            if not line:
                continue
            if line[0] == '>':
                if names:
                    s = ''.join(cur)
                    chunks.append(s)
                    lens.append(len(s))
                    cur = []
                header = line[1:].rstrip('\n').rstrip('\r')
                parts = header.split(None, 1)
                names.append(parts[0] if parts else '')
                descs.append(parts[1] if len(parts) > 1 else '')
            else:
                cur.append(line.strip())
    if names:
        s = ''.join(cur)
        chunks.append(s)
        lens.append(len(s))
    offsets = np.zeros(len(lens) + 1, dtype=np.int64)
    if lens:
        offsets[1:] = np.cumsum(lens)
    residues = digitize(''.join(chunks)) if lens else np.zeros(0, dtype=np.uint8)
    return names, descs, residues, offsets
