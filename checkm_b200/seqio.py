"""Protein FASTA -> digitised residue stream (the `genes.faa` reader that sat inside hmmsearch).

One pass in the library (`ckm_fasta_parse`): residue codes, CSR offsets and the header lines; names and descriptions are
split out of the header lines here."""
import ctypes as C
import gzip

import numpy as np

from . import _lib


def read_fasta(path):
    """Returns names, descriptions (header text after the first blank), the digitised residues and the CSR offsets."""
    opener = gzip.open if path.endswith('.gz') else open
    with opener(path, 'rb') as f:
        raw = f.read()
    return parse_fasta(raw)


def parse_fasta(raw):
    n = len(raw)
    max_rec = raw.count(b'>')
    residues = np.empty(max(n, 1), dtype=np.uint8)
    offsets = np.zeros(max_rec + 1, dtype=np.int64)
    headers = C.create_string_buffer(max(n, 1))
    nrec, nres, hb = C.c_int32(), C.c_int64(), C.c_int64()
    _lib.check(_lib.lib().ckm_fasta_parse(raw, n, residues.ctypes.data, offsets.ctypes.data, max_rec, headers, max(n, 1),
                                          C.byref(nrec), C.byref(nres), C.byref(hb)))
    if nrec.value == 0:
        return [], [], np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64)
    names, descs = [], []
    for line in headers.raw[:hb.value].decode('utf-8', 'replace').split('\n'):
        parts = line.split(None, 1)
        names.append(parts[0] if parts else '')
        descs.append(parts[1] if len(parts) > 1 else '')
    return names, descs, residues[:nres.value].copy(), offsets[:nrec.value + 1].copy()
