"""Frame sharding across GPUs.  Every (stage, time) frame is independent, so
rank r owns a contiguous block of frames and runs the same batch job on them; there is no
data-path collective.  The only exchange is the gather of the small fixed-width row tables
(per-ROI / per-adhesion) to rank 0, which writes the CSV in (frame, roi) order.

torch.distributed supplies the plumbing: NCCL between GPUs of one box (NVLink 5 / NVSwitch; the
tables are KBs per frame, so the fabric is never the limit), gloo in the CPU tests.
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one, order preserved."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_tables(table, dist=None, device=None, dst=0):
    """Gathers one numpy structured (or plain) 1-D table per rank to `dst`.  Returns the list of
    per-rank tables on `dst` (rank order == frame order for contiguous shards), None elsewhere.
    Two collectives: row counts, then the padded byte tables."""
    table = np.ascontiguousarray(table)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [table]
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([table.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    row = table.dtype.itemsize
    cap = max(max(counts), 1) * row
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if table.shape[0]:
        buf[: table.nbytes] = torch.from_numpy(table.view(np.uint8).reshape(-1).copy()).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank != dst:
        return None
    return [b[: c * row].cpu().numpy().view(table.dtype).copy() for b, c in zip(bufs, counts)]


def rows_to_table(rows, columns):
    """Fixed-width float64 table of row dicts (None / missing -> NaN) for gather_tables."""
    out = np.full((len(rows), len(columns)), np.nan, dtype=np.float64)
    for i, r in enumerate(rows):
        for j, c in enumerate(columns):
            v = r.get(c)
            if v is not None:
                out[i, j] = float(v)
    return out


def gather_rows(rows, columns, dist=None, device=None, dst=0):
    """Row dicts of every rank, concatenated in rank order on `dst` (numeric columns only)."""
    tabs = gather_tables(rows_to_table(rows, columns).reshape(-1), dist, device, dst)
    if tabs is None:
        return None
    merged = []
    for t in tabs:
        for vals in t.reshape(-1, len(columns)):
            merged.append({c: (None if np.isnan(v) else v) for c, v in zip(columns, vals)})
    return merged
