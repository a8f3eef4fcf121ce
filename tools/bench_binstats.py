"""Throughput of the bin-statistics scan (checkm_b200/csrc/ntstats.cu) against HBM bandwidth, with the
reference's algorithm (oracle/binstats_oracle.py, one core) timed beside it on a bounded sample.

    python tools/bench_binstats.py [--bins 256] [--scaffolds 100] [--mean-len 30000] [--steps 5]

One JSON line: `value` = scaffold bytes / kernel time (CUDA events, inputs resident), `e2e` = the same through
Engine.scaffold_stats with host buffers (pageable H2D copy, results copied back), `roofline` against MEASURED_PEAKS.json."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_scaffolds(nscaf, mean_len, seed=7):
    """Lengths log-normal around mean_len; bases drawn once (64 MB) and reused at random offsets; a few N-gaps per scaffold."""
    rng = np.random.default_rng(seed)
    lens = np.maximum(rng.lognormal(np.log(mean_len) - 0.5, 1.0, size=nscaf).astype(np.int64), 200)
    pool = rng.choice(np.frombuffer(b'ACGT', dtype=np.uint8), size=(64 << 20) + int(lens.max()))
    padded = (lens + 63) // 64 * 64
    starts = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    data = np.zeros(int(padded.sum()) + 64, dtype=np.uint8)
    for s, n in zip(starts, lens):
        at = int(rng.integers(0, 64 << 20))
        data[s:s + n] = pool[at:at + n]
        for g in rng.integers(0, n, size=int(n // 50000) + (rng.random() < 0.3)):
            r = int(rng.choice([1, 5, 10, 50, 100]))
            data[s + g:min(s + g + r, s + n)] = ord('N')
    return data, starts, lens


def plugin_path(nbins, data, starts, lens, per_bin, bo):
    """Files on disk -> BinStatistics.calculate -> bin_stats.analyze.tsv (what `checkm analyze` runs, main.py:370-372), against the
    oracle port doing the same per bin on one core (a bounded sample of the bins)."""
    import shutil
    import tempfile
    from checkm_b200.binStatistics import BinStatistics
    root = tempfile.mkdtemp(prefix='ckm_binstats_')
    try:
        files, nbytes = [], 0
        for b in range(nbins):
            path = os.path.join(root, 'bin%04d.fna' % b)
            with open(path, 'wb') as f:
                for i in range(b * per_bin, (b + 1) * per_bin):
                    seq = data[starts[i]:starts[i] + lens[i]].tobytes()
                    f.write(b'>scaffold_%d\n' % i)
                    f.write(b'\n'.join(seq[k:k + 60] for k in range(0, len(seq), 60)) + b'\n')
                    nbytes += len(seq)
            files.append(path)
        out = os.path.join(root, 'out')
        os.makedirs(os.path.join(out, 'storage'))
        bs = BinStatistics(1)
        bs.calculate(files[:2], out, 'warm.tsv')
        t0 = time.perf_counter()
        bs.calculate(files, out, 'bin_stats.analyze.tsv')
        t_gpu = time.perf_counter() - t0
        got = dict(line.rstrip('\n').split('\t', 1) for line in open(os.path.join(out, 'storage', 'bin_stats.analyze.tsv')))
        t0, k = time.perf_counter(), 0
        while k < nbins and time.perf_counter() - t0 < 10.0:
            want = str(bo.bin_statistics(bo.read_fasta(files[k])))
            assert got['bin%04d' % k] == want, 'bin %d differs from the oracle' % k
            k += 1
        t_cpu = time.perf_counter() - t0
        return {"bins": nbins, "MB": nbytes / 1e6, "seconds": t_gpu, "ms_per_bin": 1e3 * t_gpu / nbins, "MB_per_s": nbytes / 1e6 / t_gpu,
                "oracle_ms_per_bin": 1e3 * t_cpu / k, "oracle_bins_timed": k, "identical_text": True,
                "what": "FASTA files on disk -> BinStatistics.calculate -> bin_stats.analyze.tsv; the oracle port (one core) on the same files, dictionaries compared as text"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bins', type=int, default=256)
    ap.add_argument('--scaffolds', type=int, default=100)
    ap.add_argument('--mean-len', type=int, default=30000)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--plugin-bins', type=int, default=64, help='bins written to disk for the BinStatistics.calculate measurement (0: skip)')
    a = ap.parse_args()
    from checkm_b200 import runtime
    data, starts, lens = synthetic_scaffolds(a.bins * a.scaffolds, a.mean_len)
    total = int(lens.sum())
    eng = runtime.engine()
    eng.scaffold_stats(data, starts, lens)                   # warm-up: workspace allocation
    kernel_ms, wall_ms = [], []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        stats, cscaf, clen, ms = eng.scaffold_stats(data, starts, lens)
        wall_ms.append((time.perf_counter() - t0) * 1e3)
        kernel_ms.append(ms)
    k, w = float(np.median(kernel_ms)), float(np.median(wall_ms))
    peaks = {}
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        peaks = json.load(open(p))
    peak = float(peaks.get('hbm_gbs', 6650.0))
    # the reference's algorithm on one core, bounded
    from oracle import binstats_oracle as bo
    t0, done, i = time.perf_counter(), 0, 0
    while time.perf_counter() - t0 < a.cpu_seconds and i < len(lens):
        s = data[starts[i]:starts[i] + lens[i]].tobytes().decode('latin-1')
        bo.base_counts(s)
        bo.contig_lengths(s)
        s.count('N') + s.count('n')
        done += int(lens[i])
        i += 1
    cpu = time.perf_counter() - t0
    plugin = None
    if a.plugin_bins > 0:
        plugin = plugin_path(a.plugin_bins, data, starts, lens, a.scaffolds, bo)
    print(json.dumps({
        "metric": "scaffold bytes scanned per second", "value": total / k / 1e6, "unit": "GB/s", "kernel_ms": k, "steps": a.steps,
        "config": {"workload": "%d bins x %d scaffolds, log-normal lengths (mean %d), %.2f GB; input larger than L2" % (a.bins, a.scaffolds, a.mean_len, total / 1e9)},
        "contigs": int(len(clen)),
        "e2e": {"value": total / w / 1e6, "unit": "GB/s", "ms": w, "h2d_bytes_per_step": int(data.size + 20 * len(lens)), "d2h_bytes_per_step": int(stats.nbytes + 8 * len(clen))},
        "roofline": {"bound": "hbm", "achieved": total / k / 1e6, "peak": peak, "unit": "GB/s", "frac": total / k / 1e6 / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650", "traffic": None,
                     "algorithmic_bytes": "1 byte read per base; results are 64 B per scaffold + 8 B per contig"},
        "plugin": plugin,
        "cpu_baseline": {"value": done / cpu / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                         "sample": "%d scaffolds (%.1f MB) through oracle/binstats_oracle.py base_counts + contig_lengths + N counts" % (i, done / 1e6)},
    }))


if __name__ == '__main__':
    main()
