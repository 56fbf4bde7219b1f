"""Frame sharding across GPUs.  Every (stage, time) frame is independent, so
rank r owns a contiguous block of frames and runs the same batch job on them; there is no
data-path collective.  The only exchange is the gather of the small fixed-width row tables
(per-ROI / per-adhesion) to rank 0, which writes the CSV in (frame, roi) order.

torch.distributed supplies the plumbing: NCCL between GPUs of one box (NVLink 5 / NVSwitch; the
tables are KBs per frame, so the fabric is never the limit), gloo in the CPU tests.
"""
import numpy as np


def bind_near_gpu(device_index, apply=True):
    """Pins the calling process to the CPU cores of the GPU's NUMA node (sysfs: the PCI device's
    numa_node and the node's cpulist, intersected with the cores the process may use), so that the
    pinned staging buffers it allocates afterwards are first-touched on that node and its H2D copies
    do not cross the socket link.  On an 8-GPU box every rank's default placement is node 0, whose
    DRAM then feeds all eight PCIe links (profiles/r2_h2d_matrix.json).  Returns what it found and
    did; never raises (containers may hide sysfs or forbid the affinity call)."""
    import os
    info = {"device": int(device_index), "numa_node": None, "cpus": None, "bound": False}
    try:
        import torch
        pr = torch.cuda.get_device_properties(int(device_index))
        if hasattr(pr, "pci_bus_id"):
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:                                                # older torch: ask the runtime
            from cuda import cudart
            _, raw = cudart.cudaDeviceGetPCIBusId(32, int(device_index))
            bdf = raw.decode().rstrip("\x00").lower()
            bdf = bdf[-12:] if len(bdf) > 12 else bdf          # sysfs uses a 4-digit domain
        info["pci"] = bdf
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus |= set(range(int(a), int(b or a) + 1))
        mine = cpus & set(os.sched_getaffinity(0))
        info["cpus"] = len(mine)
        if mine and apply:
            os.sched_setaffinity(0, mine)
            info["bound"] = True
    except Exception as e:
        info["error"] = f"{type(e).__name__}: {e}"[:160]
    return info


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one, order preserved."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_many(tables, dist=None, device=None, dst=0):
    """Gathers a LIST of numpy 1-D tables per rank to `dst` with two collectives in total: the
    byte sizes of every table of every rank, then one padded byte blob per rank.  Returns, on
    `dst`, a list (one entry per table) of per-rank arrays in rank order (== frame order for
    contiguous shards); None on the other ranks."""
    tables = [np.ascontiguousarray(t) for t in tables]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [[t] for t in tables]
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    sizes = torch.tensor([t.nbytes for t in tables], dtype=torch.int64, device=dev)
    all_sizes = torch.zeros(world * len(tables), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes)
    all_sizes = all_sizes.cpu().numpy().reshape(world, len(tables))
    cap = max(int(all_sizes.sum(axis=1).max()), 16)
    blob = np.zeros(cap, dtype=np.uint8)
    off = 0
    for t in tables:
        blob[off: off + t.nbytes] = t.view(np.uint8).reshape(-1)
        off += t.nbytes
    mine = torch.from_numpy(blob).to(dev, non_blocking=True)
    everyone = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(everyone, mine)
    if rank != dst:
        return None
    host = everyone.cpu().numpy().reshape(world, cap)
    out = [[] for _ in tables]
    for r in range(world):
        off = 0
        for k, t in enumerate(tables):
            nb = int(all_sizes[r, k])
            out[k].append(host[r, off: off + nb].view(t.dtype).copy())
            off += nb
    return out


def gather_tables(table, dist=None, device=None, dst=0):
    """One table per rank -> list of per-rank tables on `dst`, None elsewhere."""
    res = gather_many([table], dist, device, dst)
    return None if res is None else res[0]


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
