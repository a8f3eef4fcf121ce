"""HMMER surface of CheckM (mirror of checkm/hmmer.py) with the search running in libckm.so on the GPU.

`HMMERRunner.search` keeps the reference's signature and contract (checkm/hmmer.py:61-74): it takes an HMM file and a
protein FASTA, writes a domtblout file at `tableOut` (and a short report at `hmmerOut`), and on failure logs and calls
`sys.exit(rtn)`.  No `hmmsearch` process is spawned and there is no CPU fallback.  `fetch`/`index` replace
`hmmfetch` (hmmer.py:97-129) through the engine's model database; `align` replaces `hmmalign` (hmmer.py:76-95; row f3 of
SURVEY.md section 8) with ckm_align; `press` has nothing to do.

`HMMERParser` / `HmmerHitDOM` / `HmmerHitTBL` read the tabular text back exactly like hmmer.py:140-311."""
import logging
import re
import sys

import numpy as np

from . import runtime
from ._lib import CkmError
from .seqio import read_fasta


class FormatError(BaseException):
    pass


class HMMERError(BaseException):
    pass


class HMMMERModeError(BaseException):
    pass


_MODES = {'dom': 'domtblout', 'tbl': 'tblout', 'align': 'align', 'fetch': 'fetch'}


def _option_value(options, flag, default):
    m = re.search(r'(?:^|\s)' + re.escape(flag) + r'\s+(\S+)', options or '')
    return float(m.group(1)) if m else default


class HMMERRunner(object):
    def __init__(self, mode="dom"):
        self.logger = logging.getLogger('timestamp')
        self.checkForHMMER()
        if mode not in _MODES:
            raise HMMMERModeError("Mode %s not understood" % mode)
        self.mode = _MODES[mode]

    def checkForHMMER(self):
        """The reference probes `hmmsearch -h` (hmmer.py:131-137); here the probe is the CUDA engine itself."""
        try:
            runtime.engine()
        except (CkmError, ImportError) as err:
            self.logger.error("The B200 search engine is not available: %s" % err)
            sys.exit(1)

    def search(self, db, query, tableOut, hmmerOut, cmdlineOptions='', bKeepOutput=True):
        if self.mode not in ('domtblout', 'tblout'):
            raise HMMMERModeError("Mode %s not compatible with search" % self.mode)
        if self.mode == 'tblout':
            self.logger.error('tblout output is not produced by the B200 engine (CheckM reads domtblout)')
            sys.exit(1)
        E = _option_value(cmdlineOptions, '-E', 10.0)
        domE = _option_value(cmdlineOptions, '--domE', 10.0)
        try:
            eng = runtime.engine()
            models = runtime.models_for(db)
            names, descs, residues, offsets = read_fasta(query)
            sdb = eng.seqdb(residues, offsets)
            try:
                hits = eng.search(models, sdb, E=E, domE=domE)
            finally:
                sdb.close()
            write_domtblout(models, hits, 0, 0, names, descs, tableOut)
            write_sidecar(tableOut, hits, names, descs, models)
            if bKeepOutput and hmmerOut and hmmerOut != '/dev/null':
                st = eng.stats()
                with open(hmmerOut, 'w') as f:
                    f.write('# checkm_b200 search report (alignment display is not produced; see SURVEY.md 8f2)\n')
                    f.write('# query HMM file: %s\n# target sequence database: %s\n' % (db, query))
                    f.write('# pairs %d; past MSV %d; past bias %d; past Vit %d; past Fwd %d; rows %d\n' %
                            (st.n_pairs, st.n_past_msv, st.n_past_bias, st.n_past_vit, st.n_past_fwd, st.n_reported))
        except CkmError as err:
            self.logger.error('search engine exited with code: %d (%s)' % (err.code, err))
            sys.exit(err.code)

    def align(self, db, query, outputFile, writeMode='>', outputFormat='PSIBLAST', trim=True):
        """`hmmalign [--trim] --outformat <fmt> db query > outputFile` (hmmer.py:76-95): every sequence of `query` aligned to the
        single model of `db` by optimal accuracy on the GPU (ckm_align); the alignment is written with all consensus columns."""
        if self.mode != 'align':
            raise HMMMERModeError("Mode %s not compatible with align" % self.mode)
        try:
            eng = runtime.engine()
            models = runtime.models_for(db)
            if models.n != 1:
                self.logger.error('hmmalign needs an HMM file with exactly one model; %s holds %d' % (db, models.n))
                sys.exit(1)
            names, descs, residues, offsets = read_fasta(query)
            sdb = eng.seqdb(residues, offsets)
            try:
                state, _oasc = eng.align(models, sdb, 0)
            finally:
                sdb.close()
            text = format_alignment(names, descs, residues, offsets, state, int(models.info()[0].M), outputFormat, trim)
            with open(outputFile, 'a' if writeMode.strip() == '>>' else 'w') as f:
                f.write(text)
        except CkmError as err:
            self.logger.error('hmmalign engine exited with code: %d (%s)' % (err.code, err))
            sys.exit(err.code)

    def fetch(self, db, key, fetchFileName, bKeyFile=False):
        if self.mode != 'fetch':
            raise HMMMERModeError("Mode %s not compatible with fetch" % self.mode)
        try:
            models = runtime.models_for(db)
            if bKeyFile:
                with open(key) as f:
                    keys = [line.strip() for line in f if line.strip() and not line.startswith('#')]
            else:
                keys = [key]
            models.write(models.select(keys), fetchFileName)
        except CkmError as err:
            self.logger.error('model fetch exited with code: %d (%s)' % (err.code, err))
            sys.exit(err.code)

    def press(self, hmmModelFile):
        """hmmpress builds binary indices for hmmscan; the engine needs none."""
        return None

    def index(self, hmmModelFile):
        """`hmmfetch --index` wrote an .ssi file; the engine looks models up in memory, so this only validates the file."""
        if self.mode != 'fetch':
            raise HMMMERModeError("Mode %s not compatible with fetch" % self.mode)
        try:
            runtime.models_for(hmmModelFile)
        except CkmError as err:
            self.logger.error('model index exited with code: %d (%s)' % (err.code, err))
            sys.exit(err.code)


_LETTERS = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"


def format_alignment(names, descs, residues, offsets, state, M, outputFormat='Pfam', trim=False):
    """The multiple alignment `hmmalign` builds from per-sequence traces: one column per consensus (match) position, plus
    insert columns wide enough for the longest insertion at each position (residues of an insertion split in half: left
    half flush left, right half flush right; the N-terminal flank flush right, the C-terminal flank flush left; `--trim`
    drops the flanks).  Match residues upper case, inserted residues lower case, '-' deletions, '.' padding.
    Formats: Pfam / Stockholm (one block, `#=GS <name> DE`, `#=GC RF` with x on consensus columns), afa, PSIBLAST."""
    nseq = len(names)
    per = []                       # per sequence: (match letter per k, insert strings per k = 0..M)
    width = [0] * (M + 1)
    for s in range(nseq):
        st = state[offsets[s]:offsets[s + 1]]
        seq = ''.join(_LETTERS[c] for c in residues[offsets[s]:offsets[s + 1]])
        match = ['-'] * (M + 1)
        ins = [''] * (M + 1)
        aligned = [i for i in range(len(st)) if st[i] != 0]
        if aligned:
            a0, a1 = aligned[0], aligned[-1]
            for i in range(a0, a1 + 1):
                k = int(st[i])
                if k > 0:
                    match[k] = seq[i].upper()
                elif k < 0:
                    ins[-k] += seq[i].lower()
            if not trim:
                ins[0], ins[M] = seq[:a0].lower(), ins[M] + seq[a1 + 1:].lower()
        elif not trim:
            ins[0] = seq.lower()
        for k in range(M + 1):
            width[k] = max(width[k], len(ins[k]))
        per.append((match, ins))
    rows = []
    for match, ins in per:
        out = ['.' * (width[0] - len(ins[0])) + ins[0]]
        for k in range(1, M + 1):
            out.append(match[k])
            w, t = width[k], ins[k]
            if k == M:
                out.append(t + '.' * (w - len(t)))
            else:
                h = len(t) // 2
                out.append(t[:h] + '.' * (w - len(t)) + t[h:])
        rows.append(''.join(out))
    rf = '.' * width[0] + ''.join('x' + '.' * width[k] for k in range(1, M + 1))
    fmt = (outputFormat or 'Pfam').lower()
    if fmt == 'afa':
        return ''.join('>%s%s\n%s\n' % (n, (' ' + d) if d else '', r) for n, d, r in zip(names, descs, rows))
    if fmt == 'psiblast':
        pad = max([len(n) for n in names] + [1])
        return ''.join('%-*s  %s\n' % (pad, n, r.replace('.', '-')) for n, r in zip(names, rows))
    pad = max([len(n) for n in names] + [len('#=GC RF')])
    lines = ['# STOCKHOLM 1.0', '']
    gs = ['#=GS %-*s DE %s' % (max(len(n) for n in names) if names else 1, n, d) for n, d in zip(names, descs) if d]
    if gs:
        lines += gs + ['']
    for n, r in zip(names, rows):
        lines.append('%-*s %s' % (pad, n, r))
    lines.append('%-*s %s' % (pad, '#=GC RF', rf))
    lines.append('//')
    return '\n'.join(lines) + '\n'


def write_domtblout(models, hits, bin_index, seq_base, names, descs, path):
    """domtblout text for one bin of a search (include/ckm.h: ckm_write_domtblout)."""
    import ctypes as C
    from . import _lib
    n = len(names)
    c_names = (C.c_char_p * max(n, 1))(*[s.encode() for s in names])
    c_descs = (C.c_char_p * max(n, 1))(*[s.encode() for s in descs])
    arr = np.ascontiguousarray(hits)
    ptr = arr.ctypes.data_as(C.POINTER(_lib.Hit))
    _lib.check(_lib.lib().ckm_write_domtblout(models._h, ptr, len(arr), int(bin_index), int(seq_base), c_names, c_descs,
                                              path.encode()))


def write_sidecar(table_path, hits, names, descs, models):
    """Binary companion of a domtblout file: the same rows as ckm_hit records, plus the names/descriptions of the target
    sequences and the (name, accession) of the queries that occur in them (rows re-indexed into those two tables)."""
    info = models.info()
    used = np.unique(hits['model']) if len(hits) else np.zeros(0, dtype=np.int32)
    useq = np.unique(hits['seq']) if len(hits) else np.zeros(0, dtype=np.int32)
    rows = hits.copy()
    if len(rows):
        rows['model'] = np.searchsorted(used, rows['model']).astype(np.int32)
        rows['seq'] = np.searchsorted(useq, rows['seq']).astype(np.int32)

    def blob(items):
        return np.frombuffer('\n'.join(items).encode('utf-8', 'replace'), dtype=np.uint8)
    with open(table_path + '.ckm.npz', 'wb') as f:
        np.savez(f, hits=rows, names=blob(names[int(i)] for i in useq), descs=blob(descs[int(i)].replace('\n', ' ') for i in useq),
                 qnames=blob(info[int(m)].name.decode() for m in used), qaccs=blob(info[int(m)].acc.decode() for m in used))


def read_sidecar(table_path):
    def lines(a, n):
        return a.tobytes().decode('utf-8', 'replace').split('\n') if n else []
    with np.load(table_path + '.ckm.npz') as z:
        hits = z['hits']
        nseq = int(hits['seq'].max()) + 1 if len(hits) else 0
        nq = int(hits['model'].max()) + 1 if len(hits) else 0
        names, descs = lines(z['names'], nseq), lines(z['descs'], nseq)
        qids = [(n, a if a else '-') for n, a in zip(lines(z['qnames'], nq), lines(z['qaccs'], nq))]
    return hits, names, descs, qids


class HMMERParser(object):
    """Iterates over the data lines of a tblout / domtblout file."""

    def __init__(self, fileHandle, mode='dom'):
        self.handle = fileHandle
        if mode == 'dom':
            self.mode = 'domtblout'
        elif mode == 'tbl':
            self.mode = 'tblout'
        else:
            raise HMMERError("Mode %s not understood, please use 'dom' or 'tbl'" % mode)

    def next(self):
        hit = self.readHitsDOM() if self.mode == 'domtblout' else self.readHitsTBL()
        return None if hit == {} else hit

    def _next_fields(self, minimum):
        while True:
            raw = self.handle.readline()
            if raw == '':
                return None
            line = raw.rstrip()
            if len(line) == 0:
                return None                      # the reference stops at the first blank line (IndexError path)
            if line[0] == '#':
                continue
            fields = re.split(r'\s+', line)
            if len(fields) < minimum:
                raise FormatError("Error processing line:\n%s" % (line))
            return fields

    def readHitsTBL(self):
        fields = self._next_fields(19)
        if fields is None:
            return {}
        return HmmerHitTBL(fields[0:18] + [" ".join(fields[18:])])

    def readHitsDOM(self):
        fields = self._next_fields(23)
        if fields is None:
            return {}
        return HmmerHitDOM(fields[0:22] + [" ".join(fields[22:])])


class HmmerHitTBL(object):
    _FIELDS = ['target_name', 'target_accession', 'query_name', 'query_accession', 'full_e_value', 'full_score',
               'full_bias', 'best_e_value', 'best_score', 'best_bias', 'exp', 'reg', 'clu', 'ov', 'env', 'dom', 'rep',
               'inc', 'target_description']

    def __init__(self, values):
        if len(values) == 19:
            for i, name in enumerate(self._FIELDS):
                v = values[i]
                if 4 <= i <= 10:
                    v = float(v)
                elif 11 <= i <= 17:
                    v = int(v)
                setattr(self, name, v)
            if self.query_accession == '-':
                self.query_accession = self.query_name

    def __str__(self):
        return "\t".join(str(getattr(self, f)) for f in self._FIELDS)


class HmmerHitDOM(object):
    """One domtblout row; attribute names are the ones CheckM's callers reach for (hmmer.py:259-285)."""
    _FIELDS = ['target_name', 'target_accession', 'target_length', 'query_name', 'query_accession', 'query_length',
               'full_e_value', 'full_score', 'full_bias', 'dom', 'ndom', 'c_evalue', 'i_evalue', 'dom_score', 'dom_bias',
               'hmm_from', 'hmm_to', 'ali_from', 'ali_to', 'env_from', 'env_to', 'acc', 'target_description']
    _INT = {2, 5, 9, 10, 15, 16, 17, 18, 19, 20}
    _FLOAT = {6, 7, 8, 11, 12, 13, 14, 21}

    def __init__(self, values):
        if len(values) == 23:
            for i, name in enumerate(self._FIELDS):
                v = values[i]
                if i in self._INT:
                    v = int(v)
                elif i in self._FLOAT:
                    v = float(v)
                setattr(self, name, v)
            if self.query_accession == '-':
                self.query_accession = self.query_name

    def __str__(self):
        return "\t".join(str(getattr(self, f)) for f in self._FIELDS)
