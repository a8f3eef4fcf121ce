#!/usr/bin/env python
"""bench.py -- genomes/hour of the marker-gene search hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (config["workload"]): BASELINE.json configs[2] stand-in -- synthetic 3 Mb bins (2,900 ORFs, ~0.9 M residues,
SURVEY.md 8d) x a 5,000-model HMM database (the 43 real, HMMER-calibrated CPR marker HMMs x 116 replicas under
distinct accessions, sum M = 1.04 M).  One "step" = the whole hot path (SSV/MSV -> bias -> Viterbi -> Forward -> domain definition ->
hit table -> marker-set reduction) over one batch of `bins_per_step` bins.  With N > 1 every rank searches its own bins
(weak scaling, no data-path collective) and the per-bin QA rows are all-gathered over NCCL at the end of each step.

value : inputs (digitised ORFs, models) resident in HBM before the timed region.
e2e   : the same through the public API with HOST buffers -- H2D of the step's residues and D2H of its hit table and QA
        rows inside the timed region.
--impl reference : the CPU restatement of HMMER3's pipeline (oracle/, "port") on all host cores, a bounded sample/step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CPR = os.path.join(ROOT, 'tests', 'golden', 'cpr_43_markers.hmm')

N_MODELS = 5000
ORFS_PER_BIN = 2900
BINS_PER_STEP = 16


def rank_info():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def model_db(tag=''):
    """5,000-model database file, written once per box under /tmp: every one of the 43 real, HMMER-calibrated CPR marker
    HMMs repeated under distinct names/accessions until there are 5,000 models (sum M = 1.04 M).  Replicas are separate
    models to the engine (own tiles, own tables, own hits); using real models keeps the STATS lines -- and with them the
    filter pass rates of the cascade (2% / 0.1% / 1e-5) -- those of a real search, which model rows stitched at random
    do not (their Viterbi/Forward tails are several bits off any fitted calibration)."""
    path = '/tmp/ckm_bench_db_%d%s.hmm' % (N_MODELS, tag)
    if not os.path.exists(path):
        recs = [r + '//\n' for r in open(CPR).read().split('//\n') if r.strip()]
        tmp = path + '.%d.tmp' % os.getpid()
        with open(tmp, 'w') as out:
            n = 0
            rep = 0
            while n < N_MODELS:
                for r in recs:
                    if n >= N_MODELS:
                        break
                    if rep == 0:
                        out.write(r)
                    else:
                        lines = r.split('\n')
                        for i, ln in enumerate(lines[:6]):
                            if ln.startswith('NAME '):
                                lines[i] = ln + '_r%d' % rep
                            elif ln.startswith('ACC '):
                                acc = ln.split()[1]
                                lines[i] = 'ACC   %s' % ((acc.split('.')[0] + 'r%d.' % rep + acc.split('.')[1]) if '.' in acc else acc + 'r%d' % rep)
                        out.write('\n'.join(lines))
                    n += 1
                rep += 1
        os.replace(tmp, path)
    return path, CPR


def make_bins(plant_path, n, seed0):
    from tools import synth
    hm = synth.read_hmms(plant_path)
    bins = [synth.make_bin('bin%d' % (seed0 + i), hm, seed=seed0 + i, n_orfs=ORFS_PER_BIN, copies=(0, 1, 1, 1, 2)) for i in range(n)]
    res = np.concatenate([b.residues for b in bins])
    lens = np.concatenate([np.diff(b.offsets) for b in bins])
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    binof = np.concatenate([np.full(b.nseq, i, np.int32) for i, b in enumerate(bins)])
    return bins, res, off, binof


class ClockSampler(threading.Thread):
    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured'
    return {"hbm_gbs": 6650.0}, 'fallback'


def ssv_traffic():
    """DRAM bytes per SSV launch from the committed `ncu --set full` capture (profiles/r1_ssv32_traffic.json), or None."""
    p = os.path.join(ROOT, 'profiles', 'r1_ssv32_traffic.json')
    try:
        return json.load(open(p))
    except Exception:
        return None


def oracle_sample(db_path, bins, nthreads, n_models=64):
    """CPU restatement on a bounded sample: bin 0 x the first `n_models` models; returns (seconds, sum M of the sample)."""
    from oracle import pyoracle as po
    sub = '/tmp/ckm_bench_cpu_sample_%d.hmm' % n_models
    if not os.path.exists(sub):
        with open(db_path) as f, open(sub + '.tmp', 'w') as out:
            n = 0
            for line in f:
                out.write(line)
                if line.startswith('//'):
                    n += 1
                    if n >= n_models:
                        break
        os.replace(sub + '.tmp', sub)
    hf = po.HmmFile(sub)
    b = bins[0]
    t0 = time.perf_counter()
    rp = po.search(hf, b.residues, b.offsets, nthreads=nthreads)
    dt = time.perf_counter() - t0
    nh = rp.contents.nhits
    po.free_results(rp)
    return dt, sum(h.M for h in hf.headers), nh


def total_model_positions(db_path):
    tot = 0
    with open(db_path) as f:
        for line in f:
            if line.startswith('LENG'):
                tot += int(line.split()[1])
    return tot


def workload_name(sumM_all):
    return ("configs[2] stand-in: synthetic 3 Mb bins (2,900 ORFs, ~0.9 M residues, 0-2 planted homologs per CPR family) x 5,000 HMMs "
            "(the 43 real HMMER-calibrated CPR models x 116 replicas under distinct accessions, sum M = %d)" % sumM_all)


def run_reference(args):
    rank, local, world = rank_info()
    if rank != 0:
        return
    db_path, plant = model_db()
    bins, _, _, _ = make_bins(plant, 1, 1000)
    cores = os.cpu_count() or 1
    sumM_all = total_model_positions(db_path)
    for _ in range(max(args.warmup, 0) and 1):
        oracle_sample(db_path, bins, cores, n_models=16)
    t = 0.0
    sumM = 0
    for _ in range(args.steps):
        dt, sumM, _ = oracle_sample(db_path, bins, cores, n_models=400)
        t += dt
    per_bin = (t / args.steps) * (sumM_all / float(sumM))        # seconds to search one bin against all 5,000 models
    gph = 3600.0 / per_bin
    sample = "1 bin (%d ORFs) x first 400 of %d models per step, scaled by model positions (%d of %d)" % (ORFS_PER_BIN, N_MODELS, sumM, sumM_all)
    line = {"metric": "genomes/hour", "value": gph, "unit": "genomes/hour", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16/f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(sumM_all), "bins_per_step": args.bins_per_step, "orfs_per_bin": ORFS_PER_BIN, "n_models": N_MODELS,
                       "per_gpu_bins_per_step": args.bins_per_step, "parallelism": "CPU: one thread per ORF block, all host cores",
                       "sample_per_step": sample},
            "cpu_baseline": {"value": gph, "unit": "genomes/hour", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": gph, "unit": "genomes/hour", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ckm')
    ap.add_argument('--bins-per-step', type=int, default=BINS_PER_STEP)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--pipeline', type=int, default=2, help='batches in flight per GPU (one engine + host thread each)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
        return
    rank, local, world = rank_info()
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    os.environ['CKM_DEVICE'] = str(local)
    from checkm_b200 import _lib, runtime
    from checkm_b200.resultsParser import QA_DTYPE
    B = args.bins_per_step
    if rank == 0:
        db_path, plant = model_db()
    if world > 1:
        dist.barrier()
    db_path, plant = model_db()
    NP = max(1, args.pipeline)
    engs = runtime.engines(NP)
    eng = engs[0]
    t0 = time.perf_counter()
    models = runtime.models_for(db_path)
    t_load = time.perf_counter() - t0
    info = models.info()
    sumM_all = sum(int(mi.M) for mi in info)
    # two alternating batches of bins per rank, distinct across ranks; every pipeline slot keeps its own resident copy
    batches = []
    for z in range(2):
        bins, res, off, binof = make_bins(plant, B, 10000 * (rank + 1) + 100 * z)
        batches.append(dict(bins=bins, res=res, off=off, binof=binof, db=[e_.seqdb(res, off, binof, B) for e_ in engs]))
    # reduction metadata: one marker set per bin = all models (HMM-file semantics), no clans
    nm = models.n
    acc_is_tigr = np.asarray([1 if b'TIGR' in mi.acc else 0 for mi in info], dtype=np.uint8)
    is_pfam = np.asarray([1 if mi.acc.startswith(b'PF') else 0 for mi in info], dtype=np.uint8)
    clan = np.full(nm, -1, dtype=np.int32)
    nest_off = np.zeros(nm + 1, dtype=np.int64)
    has = np.zeros((nm, 3), dtype=np.int32)
    cut = np.zeros((nm, 6), dtype=np.float64)
    for i, mi in enumerate(info):
        has[i] = (mi.has_ga, mi.has_tc, mi.has_nc)
        cut[i] = (mi.ga_d[0], mi.ga_d[1], mi.tc_d[0], mi.tc_d[1], mi.nc_d[0], mi.nc_d[1])
    opts = _lib.ReduceOpts()
    opts.evalue_threshold, opts.evalue_exp10, opts.evalue_mant = 1e-10, -10, 10.0
    opts.length_threshold, opts.pseudogene_length = 0.7, 0.3
    bin_set_off = np.arange(B + 1, dtype=np.int64)
    set_marker_off = (np.arange(B + 1, dtype=np.int64) * nm)
    set_marker_idx = np.tile(np.arange(nm, dtype=np.int32), B)

    def reduce_hits(batch, hits, eng=eng):
        names = [n for b in batch['bins'] for n in b.names]
        key = id(batch)
        if key not in reduce_hits.cache:
            scaf, num, rank_ = [], [], []
            base = 0
            for bi, b in enumerate(batch['bins']):
                order = {n: r for r, n in enumerate(sorted(b.names))}
                for n in b.names:
                    c = n.rfind('_')
                    scaf.append(hash((bi, n[:c])) & 0x7fffffff)
                    num.append(int(n[c + 1:]))
                    rank_.append(order[n])
            reduce_hits.cache[key] = tuple(np.asarray(a, dtype=np.int32) for a in (scaf, num, rank_))
        scaf, num, rank_ = reduce_hits.cache[key]
        meta = _lib.ReduceMeta()
        meta.is_pfam, meta.is_tigr, meta.clan = is_pfam.ctypes.data, acc_is_tigr.ctypes.data, clan.ctypes.data
        meta.nest_off, meta.nest_idx = nest_off.ctypes.data, None
        meta.has_cut, meta.cutoffs = has.ctypes.data, cut.ctypes.data
        meta.scaffold_id, meta.orf_num, meta.name_rank = scaf.ctypes.data, num.ctypes.data, rank_.ctypes.data
        meta.bin_set_off, meta.set_marker_off, meta.set_marker_idx = bin_set_off.ctypes.data, set_marker_off.ctypes.data, set_marker_idx.ctypes.data
        qa = C.POINTER(_lib.QaRow)()
        nqa = C.c_int32()
        mh = C.POINTER(_lib.MarkerHit)()
        nmh = C.c_int64()
        harr = np.ascontiguousarray(hits)
        _lib.check(_lib.lib().ckm_reduce(eng._h, nm, len(names), B, harr.ctypes.data_as(C.POINTER(_lib.Hit)), len(harr), C.byref(opts),
                                         C.byref(meta), C.byref(qa), C.byref(nqa), C.byref(mh), C.byref(nmh)))
        buf = (C.c_char * (nqa.value * C.sizeof(_lib.QaRow))).from_address(C.addressof(qa.contents))
        rows = np.frombuffer(buf, dtype=QA_DTYPE).copy()
        _lib.lib().ckm_free(qa)
        _lib.lib().ckm_free(mh)
        return rows, nmh.value
    reduce_hits.cache = {}

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    gather_buf = torch.empty((world, B * QA_DTYPE.itemsize), dtype=torch.uint8, device='cuda') if world > 1 else None

    def gather(rows):
        if world == 1:
            return rows
        mine = torch.from_numpy(rows.view(np.uint8).copy()).cuda()
        dist.all_gather_into_tensor(gather_buf.view(-1), mine)        # NCCL all-gather of the fixed-width QA rows (config #4)
        return gather_buf

    host_ms = {'search': 0.0, 'reduce': 0.0, 'gather': 0.0}
    gather_lock = threading.Lock()

    def step_resident(i, w):
        """One step on pipeline slot w: inputs resident in HBM."""
        e_ = engs[w]
        batch = batches[(i // NP) % 2]
        flush_buf.zero_()                              # 256 MiB write, asynchronous to the engine streams
        t1 = time.perf_counter()
        hits = e_.search(models, batch['db'][w])
        st = e_.stats()
        t2 = time.perf_counter()
        rows, nmh = reduce_hits(batch, hits, e_)
        t3 = time.perf_counter()
        with gather_lock:
            gather(rows)
        t4 = time.perf_counter()
        return hits, st, rows, (1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3))

    def step_e2e(i, w):
        """The same through the public API with HOST buffers: H2D of the residues, D2H of the hit table and QA rows."""
        e_ = engs[w]
        batch = batches[(i // NP) % 2]
        flush_buf.zero_()
        db = e_.seqdb(batch['res'], batch['off'], batch['binof'], B)       # host buffers -> HBM
        try:
            hits = e_.search(models, db)                                    # hit table back on the host
            st = e_.stats()
        finally:
            db.close()
        rows, nmh = reduce_hits(batch, hits, e_)
        with gather_lock:
            gather(rows)
        return hits, st, rows, (0.0, 0.0, 0.0)

    def run_steps(nsteps, fn):
        """nsteps steps, NP in flight: slot w takes steps w, w+NP, ...  Returns the per-step records in step order."""
        out = [None] * nsteps
        errs = []

        def worker(w):
            try:
                torch.cuda.set_device(local)           # the current device is per host thread
                for i in range(w, nsteps, NP):
                    out[i] = fn(i, w)
            except BaseException as ex:       # surfaced on the main thread
                errs.append(ex)
        if NP == 1:
            worker(0)
        else:
            ths = [threading.Thread(target=worker, args=(w,)) for w in range(NP)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        if errs:
            raise errs[0]
        return out

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(nsteps, fn):
        """nsteps steps between two device-wide synchronisations (+ barrier), timed on the device with CUDA events recorded
        right after the first and right after the second synchronisation (the engines launch on their own streams, so the
        events bracket the region rather than ride one stream); the host wall clock is kept as a cross-check."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        ev0.record()
        t0 = time.perf_counter()
        recs_ = run_steps(nsteps, fn)
        sync()
        ev1.record()
        ev1.synchronize()
        return recs_, ev0.elapsed_time(ev1) / 1e3, time.perf_counter() - t0

    run_steps(max(args.warmup, 0), step_resident)
    sampler = ClockSampler(local)
    sampler.start()
    recs, t_res, t_res_wall = timed(args.steps, step_resident)
    ssv_ms = msv_ms = other_ms = 0.0
    launches = cells = pairs = 0
    for hits, st, rows, hm in recs:
        ssv_ms += st.ms_ssv
        msv_ms += st.ms_msv
        other_ms += st.ms_bias + st.ms_vit + st.ms_fwd + st.ms_domdef
        launches += st.kernel_launches + 5
        cells += st.n_cells
        pairs += st.n_pairs
        for k_, v_ in zip(('search', 'reduce', 'gather'), hm):
            host_ms[k_] += v_
    last = recs[-1][:3]
    recs_e, t_e2e, t_e2e_wall = timed(args.steps, step_e2e)
    # one batch at a time on one engine: the stage times of an undisturbed search (the SSV roofline is quoted on both)
    iso = []
    for i in range(2):
        iso.append(step_resident(i * NP, 0)[1])
    sync()
    iso_ssv_ms = sum(s_.ms_ssv for s_ in iso) / len(iso)
    iso_st = iso[-1]
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        tt = torch.tensor([t_res, t_e2e], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hits, st, rows = last
    genomes = B * args.steps * world
    value = genomes / t_res * 3600.0
    e2e = genomes / t_e2e * 3600.0
    # roofline of the dominant kernel (SSV pre-filter): algorithmic bytes = sum over pairs of (L + 4)
    resid = float(sum(len(b['res']) for b in batches[:1]))
    alg_bytes_per_step = resid * nm + 4.0 * (pairs / args.steps)
    peaks, peak_kind = measured_peaks()
    traffic = ssv_traffic()
    ssv_s = (ssv_ms / args.steps) / 1000.0
    achieved = alg_bytes_per_step / ssv_s / 1e9
    real_cells = resid * sumM_all
    h2d = int(batches[0]['db'][0].residues.nbytes + batches[0]['off'].nbytes + batches[0]['binof'].nbytes)
    d2h = int(hits.nbytes + rows.nbytes)
    line = {"metric": "genomes/hour", "value": value, "unit": "genomes/hour", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * t_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16 (SSV) / u8 (MSV) / int16 (Viterbi) / f32 (Forward, domain definition)", "data": "synthetic",
            "config": {"workload": workload_name(sumM_all),
                       "bins_per_step": B, "orfs_per_bin": ORFS_PER_BIN, "n_models": nm, "per_gpu_bins_per_step": B, "parallelism": "bins sharded, 1 process/GPU",
                       "batches_in_flight_per_gpu": NP,
                       "l2": "256 MiB flush write issued before every step; the model tables alone (> 200 MB) exceed L2", "model_load_s": t_load},
            "e2e": {"value": e2e, "unit": "genomes/hour", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": achieved / peaks.get("hbm_gbs"),
                         "traffic": (traffic["dram_bytes_per_bin"] * B) if traffic else None, "traffic_source": (traffic or {}).get("source"), "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == 'measured' else 'fallback 6650',
                         "kernel": "ssv_kernel<J> (SSV pre-filter, all pairs)", "kernel_ms_per_step": ssv_ms / args.steps,
                         "isolated": {"kernel_ms": iso_ssv_ms, "achieved": alg_bytes_per_step / (iso_ssv_ms / 1e3) / 1e9,
                                      "frac": alg_bytes_per_step / (iso_ssv_ms / 1e3) / 1e9 / peaks.get("hbm_gbs"),
                                      "what": "same kernel, one batch in flight (no other stream on the SMs)"},
                         "note": "the stage is DP-cell bound, not HBM bound (SURVEY.md 8d): see gcups"},
            "gcups": {"real_cells_per_step": real_cells, "tile_cells_per_step": cells / args.steps, "ssv_gcups_real": real_cells / ssv_s / 1e9,
                      "ssv_gcups_tile": cells / args.steps / ssv_s / 1e9, "ssv_gcups_real_isolated": real_cells / (iso_ssv_ms / 1e3) / 1e9, "smem_bound_gcups": 148 * (2048.0 / 31.0) * 1.965,
                      "smem_bound_note": "J=32 tile row = 7 LDS.128 (1 int8 chunk + 6 int16 quads = 28 wavefronts) + 1 SHFL (3) per 2048 cells at 1 wavefront/clk/SM, 148 SMs, 1.965 GHz",
                      "ssv_gcups_tile_isolated": cells / args.steps / (iso_ssv_ms / 1e3) / 1e9,
                      "ssv_frac_of_smem_bound": (cells / args.steps / (iso_ssv_ms / 1e3) / 1e9) / (148 * (2048.0 / 31.0) * 1.965),
                      "stage_ms_per_step": {"ssv": ssv_ms / args.steps, "msv_exact": msv_ms / args.steps, "bias+vit+fwd+domdef": other_ms / args.steps,
                                            "wall_ms_per_step": {k_: v_ / args.steps for k_, v_ in host_ms.items()},
                                            "isolated_step": {"ssv": iso_st.ms_ssv, "msv_exact": iso_st.ms_msv, "bias": iso_st.ms_bias, "vit": iso_st.ms_vit, "fwd": iso_st.ms_fwd,
                                                              "domdef": iso_st.ms_domdef, "total": iso_st.ms_total}}},
            "cascade": {"pairs": int(st.n_pairs), "ssv_cand": int(st.n_ssv_cand), "past_msv": int(st.n_past_msv), "past_bias": int(st.n_past_bias),
                        "past_vit": int(st.n_past_vit), "past_fwd": int(st.n_past_fwd), "rows": int(st.n_reported), "vit_int32_redo": int(st.n_vit_redo)},
            "timing": {"how": "CUDA events around the K timed steps (after barrier + device synchronize on both sides), max over ranks", "host_wall_s": t_res_wall, "host_wall_e2e_s": t_e2e_wall},
            "clocks": sampler.summary()}
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        dt, sumM, _ = oracle_sample(db_path, batches[0]['bins'], cores, n_models=400)
        per_bin = dt * (sumM_all / float(sumM))
        line["cpu_baseline"] = {"value": 3600.0 / per_bin, "unit": "genomes/hour", "cores": cores, "kind": "port",
                                "sample": "1 bin x first 400 of %d models (%d of %d model positions), %.1f s, scaled" % (nm, sumM, sumM_all, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
