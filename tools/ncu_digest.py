"""Digest of one ncu capture exported as CSV (development tool).

    python tools/ncu_digest.py details.csv [sass.csv]

Prints the speed-of-light / scheduler / occupancy lines of the details page, and from the SASS page the instructions per
iteration of the hottest loop grouped into runs of equal execution frequency, with their share of the stall samples."""
import csv
import sys

KEEP = ('GPU Speed Of Light Throughput', 'Compute Workload Analysis', 'Memory Workload Analysis', 'Scheduler Statistics',
        'Warp State Statistics', 'Occupancy', 'Launch Statistics', 'Instruction Statistics')
WANT = ('Duration', 'Memory Throughput', 'DRAM Throughput', 'Compute (SM) Throughput', 'Issue Slots Busy', 'Executed Ipc Active', 'L1/TEX Hit Rate',
        'Mem Pipes Busy', 'One or More Eligible', 'Active Warps Per Scheduler', 'Eligible Warps Per Scheduler', 'Warp Cycles Per Issued Instruction',
        'Avg. Active Threads Per Warp', 'Executed Instructions', 'Grid Size', 'Registers Per Thread', 'Dynamic Shared Memory Per Block',
        'Theoretical Occupancy', 'Achieved Occupancy', 'Block Limit Registers', 'Block Limit Shared Mem', 'L2 Hit Rate')


def details(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        if d.get('Section Name') in KEEP and d.get('Metric Name') in WANT:
            print('%-34s %-38s %s %s' % (d['Section Name'], d['Metric Name'], d['Metric Value'], d['Metric Unit']))


def sass(path):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    ex = [int(r[ix['Instructions Executed']] or 0) for r in data]
    sm = [int(r[ix['# Samples']] or 0) for r in data]
    loop = max(ex)
    S = max(sum(sm), 1)
    print('hottest instruction executed %d times; instructions per such iteration: %.1f' % (loop, sum(ex) / loop))
    stall_cols = [h for h in hdr if h.startswith('stall_')]
    tot = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
    print('stall samples:', ', '.join('%s %.1f%%' % (h[6:], 100.0 * v / max(sum(tot.values()), 1)) for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
    i = 0
    while i < len(data):
        j, ratio = i, ex[i] / loop
        while j < len(data) and abs(ex[j] / loop - ratio) < 0.06:
            j += 1
        if sum(ex[i:j]) / loop >= 1.0 or sum(sm[i:j]) / S >= 0.01:
            print('%5d..%5d  n=%4d  freq=%.2f  instr/iter=%6.1f  samples=%5.1f%%   %s' % (i, j - 1, j - i, ratio, sum(ex[i:j]) / loop, 100.0 * sum(sm[i:j]) / S, data[i][ix['Source']].strip()[:60]))
        i = j


if __name__ == '__main__':
    details(sys.argv[1])
    if len(sys.argv) > 2:
        sass(sys.argv[2])
