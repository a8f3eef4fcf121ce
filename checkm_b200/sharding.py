"""Multi-GPU plumbing for the bin-parallel search: bins are independent (E-values use Z = ORFs of the
bin itself), so ranks take disjoint bins and nothing crosses GPUs on the data path; the per-bin QA rows (64 B each) are
all-gathered at the end -- NCCL on GPUs, gloo in the CPU tests."""
import numpy as np


def partition_bins(costs, world):
    """Greedy longest-processing-time assignment of bins to ranks by cost (residues x model positions).
    Returns a list of index arrays, one per rank; deterministic (ties by bin index)."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    owner = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[r].append(i)
        load[r] += float(costs[i])
    return [np.asarray(sorted(o), dtype=np.int64) for o in owner]


def gather_rows(rows, dist, device=None):
    """All-gather a structured numpy array of fixed-width rows whose count differs per rank.
    Returns the concatenation in rank order.  `dist` is torch.distributed (already initialised)."""
    import torch
    world = dist.get_world_size()
    if world == 1:
        return rows
    dev = device if device is not None else ('cuda' if dist.get_backend() == 'nccl' else 'cpu')
    n = torch.tensor([len(rows)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    width = rows.dtype.itemsize
    cap = max(counts) if counts else 0
    mine = torch.zeros(max(cap, 1) * width, dtype=torch.uint8, device=dev)
    if len(rows):
        mine[:len(rows) * width] = torch.from_numpy(np.ascontiguousarray(rows).view(np.uint8).reshape(-1).copy()).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    blob = b''.join(o.cpu().numpy().tobytes()[:c * width] for o, c in zip(out, counts))      # byte-wise: keeps struct padding intact
    return np.frombuffer(bytearray(blob), dtype=rows.dtype)
