"""ctypes wrapper over oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  Nothing in checkm_b200/ does.  Filters pinned by HMMER's own calibration numbers, domain definition PARITY
UNPINNED (see hmmer_oracle.h).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Domain(C.Structure):
    _fields_ = [("ienv", C.c_int), ("jenv", C.c_int), ("hmmfrom", C.c_int), ("hmmto", C.c_int),
                ("sqfrom", C.c_int), ("sqto", C.c_int), ("envsc", C.c_float), ("domcorrection", C.c_float),
                ("dombias", C.c_float), ("oasc", C.c_float), ("bitscore", C.c_float), ("lnP", C.c_double),
                ("is_reported", C.c_int)]


class Hit(C.Structure):
    _fields_ = [("seqidx", C.c_int), ("model", C.c_int), ("L", C.c_int), ("pre_score", C.c_float),
                ("score", C.c_float), ("sum_score", C.c_float), ("lnP", C.c_double), ("ndom", C.c_int),
                ("nreported", C.c_int), ("nregions", C.c_int), ("nclustered", C.c_int), ("nenvelopes", C.c_int),
                ("is_reported", C.c_int), ("dcl", C.POINTER(Domain))]


class FilterResult(C.Structure):
    _fields_ = [("msv_xJ", C.c_int), ("msv_sc", C.c_float), ("nullsc", C.c_float), ("filtersc", C.c_float),
                ("vit_sc", C.c_float), ("fwd_sc", C.c_float), ("passed_msv", C.c_int), ("passed_bias", C.c_int),
                ("passed_vit", C.c_int), ("passed_fwd", C.c_int)]


class Results(C.Structure):
    _fields_ = [("nhits", C.c_int), ("hits", C.POINTER(Hit)), ("Z", C.c_double), ("domZ", C.POINTER(C.c_double)),
                ("nmodels", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_hmmfile_read.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        L.orc_hmm_at.restype = C.c_void_p
        L.orc_hmm_at.argtypes = [C.c_void_p, C.c_int]
        L.orc_profile_create.restype = C.c_void_p
        L.orc_profile_create.argtypes = [C.c_void_p]
        L.orc_profile_free.argtypes = [C.c_void_p]
        L.orc_hmms_free.argtypes = [C.c_void_p, C.c_int]
        L.orc_null1.restype = C.c_float
        L.orc_msv.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.orc_ssv_xe.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_biasfilter.restype = C.c_float
        L.orc_biasfilter.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_vitfilter.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.orc_forward_parser.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.orc_backward_parser.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.orc_filters.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(FilterResult)]
        L.orc_pipeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Hit)]
        L.orc_search.restype = C.POINTER(Results)
        L.orc_search.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                 C.c_double, C.c_int]
        L.orc_results_free.argtypes = [C.POINTER(Results)]
        L.orc_write_domtblout.argtypes = [C.POINTER(Results), C.POINTER(C.c_void_p), C.POINTER(C.c_char_p),
                                          C.POINTER(C.c_char_p), C.c_char_p]
        L.orc_digitize.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        L.orc_profile_enable_simd.argtypes = [C.c_void_p]
        L.orc_msv_simd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.orc_vitfilter_simd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.orc_striped_create.restype = C.c_void_p
        L.orc_striped_create.argtypes = [C.c_void_p]
        L.orc_striped_free.argtypes = [C.c_void_p]
        L.orc_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_float)]
        L.orc_stage_scores.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


AMINO = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"


def digitize(seq):
    s = seq.encode() if isinstance(seq, str) else seq
    out = np.empty(len(s), dtype=np.uint8)
    lib().orc_digitize(s, len(s), out.ctypes.data)
    return out


class HmmHeader(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("acc", C.c_char * 64), ("desc", C.c_char * 256), ("M", C.c_int),
                ("mat", C.POINTER(C.c_float)), ("ins", C.POINTER(C.c_float)), ("t", C.POINTER(C.c_float)),
                ("compo", C.c_float * 20), ("has_compo", C.c_int), ("ga", C.c_float * 2), ("tc", C.c_float * 2),
                ("nc", C.c_float * 2), ("has_ga", C.c_int), ("has_tc", C.c_int), ("has_nc", C.c_int),
                ("evparam", C.c_float * 6), ("has_stats", C.c_int)]


class Profile(C.Structure):
    _fields_ = [("M", C.c_int), ("hmm", C.c_void_p), ("tsc", C.POINTER(C.c_float)), ("bm", C.POINTER(C.c_float)),
                ("msc", C.POINTER(C.c_float)), ("rbv", C.POINTER(C.c_uint8)), ("tbm_b", C.c_uint8),
                ("tec_b", C.c_uint8), ("base_b", C.c_uint8), ("bias_b", C.c_uint8), ("scale_b", C.c_float),
                ("rwv", C.POINTER(C.c_int16)), ("twv", C.POINTER(C.c_int16)), ("base_w", C.c_int16),
                ("xw_e_loop", C.c_int16), ("xw_e_move", C.c_int16), ("ddbound_w", C.c_int16), ("scale_w", C.c_float),
                ("rfv", C.POINTER(C.c_float)), ("tfv", C.POINTER(C.c_float))]


class HmmFile:
    """All models of one HMMER3/f file, with profiles configured."""

    def __init__(self, path):
        L = lib()
        self.path = path
        self._h = C.c_void_p()
        n = C.c_int()
        st = L.orc_hmmfile_read(path.encode(), C.byref(self._h), C.byref(n))
        if st != 0:
            raise IOError("oracle failed to read %s (status %d)" % (path, st))
        self.n = n.value
        self.hmm_ptrs = [L.orc_hmm_at(self._h, i) for i in range(self.n)]
        self.headers = [C.cast(p, C.POINTER(HmmHeader)).contents for p in self.hmm_ptrs]
        self.prof_ptrs = [L.orc_profile_create(p) for p in self.hmm_ptrs]
        self.profiles = [C.cast(p, C.POINTER(Profile)).contents for p in self.prof_ptrs]

    def enable_simd(self):
        """Switch the MSV / Viterbi filters of every profile to the SSE2 striped versions (CPU baseline of bench.py)."""
        for p in self.prof_ptrs:
            lib().orc_profile_enable_simd(p)

    def names(self):
        return [h.name.decode() for h in self.headers]

    def accs(self):
        return [h.acc.decode() for h in self.headers]

    def prof_array(self, idx=None):
        idx = range(self.n) if idx is None else idx
        arr = (C.c_void_p * len(idx))(*[self.prof_ptrs[i] for i in idx])
        return arr


def filters(hf, m, dsq):
    r = FilterResult()
    d = np.ascontiguousarray(dsq, dtype=np.uint8)
    lib().orc_filters(hf.prof_ptrs[m], d.ctypes.data, len(d), C.byref(r))
    return r


def vitfilter(hf, m, dsq):
    """ViterbiFilter score (nats) of one pair; +inf on int16 overflow."""
    d = np.ascontiguousarray(dsq, dtype=np.uint8)
    sc = C.c_float()
    lib().orc_vitfilter(hf.prof_ptrs[m], d.ctypes.data, len(d), C.byref(sc))
    return sc.value


def msv(hf, m, dsq):
    sc = C.c_float()
    xj = C.c_int()
    d = np.ascontiguousarray(dsq, dtype=np.uint8)
    lib().orc_msv(hf.prof_ptrs[m], d.ctypes.data, len(d), C.byref(sc), C.byref(xj))
    return sc.value, xj.value


def ssv_xe(hf, m, dsq):
    d = np.ascontiguousarray(dsq, dtype=np.uint8)
    return lib().orc_ssv_xe(hf.prof_ptrs[m], d.ctypes.data, len(d))


def align(hf, m, dsq):
    """hmmalign of one sequence to model m: (state per residue: +k match / -k insert / 0 flank, optimal-accuracy score)."""
    d = np.ascontiguousarray(dsq, dtype=np.uint8)
    state = np.zeros(len(d), dtype=np.int32)
    sc = C.c_float()
    rc = lib().orc_align(hf.prof_ptrs[m], d.ctypes.data, len(d), state.ctypes.data, C.byref(sc))
    return state, sc.value, rc


def stage_scores(hf, residues, offsets, models=None, msv=True, vit=True, fwd=True, nthreads=1):
    """Raw MSV / Viterbi / Forward filter scores (nats) of every (model, sequence) pair: dict of [nmodels, nseq] arrays."""
    res = np.ascontiguousarray(residues, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    arr = hf.prof_array(models)
    n = len(off) - 1
    out = {k: np.empty((len(arr), n), dtype=np.float32) for k, on in (('msv', msv), ('vit', vit), ('fwd', fwd)) if on}
    lib().orc_stage_scores(arr, len(arr), res.ctypes.data, off.ctypes.data, n,
                           out['msv'].ctypes.data if msv else None, out['vit'].ctypes.data if vit else None,
                           out['fwd'].ctypes.data if fwd else None, nthreads)
    return out


def search(hf, residues, offsets, E=0.1, domE=0.1, nthreads=1, models=None):
    """Returns the ctypes Results pointer (free with free_results)."""
    res = np.ascontiguousarray(residues, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    arr = hf.prof_array(models)
    return lib().orc_search(arr, len(arr), res.ctypes.data, off.ctypes.data, len(off) - 1, E, domE, nthreads)


def hits_table(rp):
    """Flatten Results into a list of dict rows (one per reported domain), domtblout order."""
    r = rp.contents
    rows = []
    for h in range(r.nhits):
        hit = r.hits[h]
        if not hit.is_reported:
            continue
        nd = 0
        for d in range(hit.ndom):
            dom = hit.dcl[d]
            if not dom.is_reported:
                continue
            nd += 1
            rows.append(dict(seqidx=hit.seqidx, model=hit.model, tlen=hit.L,
                             full_E=float(np.exp(hit.lnP) * r.Z), full_score=hit.score,
                             full_bias=hit.pre_score - hit.score, dom=nd, ndom=hit.nreported,
                             c_E=float(np.exp(dom.lnP) * r.domZ[hit.model]), i_E=float(np.exp(dom.lnP) * r.Z),
                             dom_score=dom.bitscore, dom_bias=dom.dombias / np.log(2.0),
                             hmm_from=dom.hmmfrom, hmm_to=dom.hmmto, ali_from=dom.sqfrom, ali_to=dom.sqto,
                             env_from=dom.ienv, env_to=dom.jenv,
                             acc=dom.oasc / (1.0 + abs(dom.jenv - dom.ienv))))
    return rows


def free_results(rp):
    lib().orc_results_free(rp)


def write_domtblout(rp, hf, seqnames, seqdescs, path, models=None):
    arr = hf.prof_array(models)
    n = len(seqnames)
    names = (C.c_char_p * n)(*[s.encode() for s in seqnames])
    descs = (C.c_char_p * n)(*[s.encode() for s in seqdescs])
    return lib().orc_write_domtblout(rp, arr, names, descs, path.encode())
