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


class ShmTableRing:
    """The table gather of an N-rank job on ONE host, without any device work: every rank's step
    tables are downloaded (its own D2H, its own PCIe link) straight into a ring of entries in a
    shared-memory segment the rank owns; the destination rank maps all N segments and reads the
    entries in place once their sequence flag is set.

    Why not the NCCL all-gather + one download on the destination (batch.FrameBatchJob keeps that
    path, gather_via = "nccl", for ranks on different hosts): it moves every table over PCIe twice,
    the second time all N ranks' tables through the destination's single link, and its collective
    couples the ranks -- measured on 8 B200s, 2.59 ms per step against 2.48 for the slowest rank
    without any gather (profiles/README.md).  torch.distributed is used once, at setup: the segment
    name, the entry capacity and the go / no-go are agreed with it.

    Segment of rank r:  int64 header [run, final_pos, ack, -, -, -, -, -], int64 flag per entry,
    then n_entries x cap bytes.  Entry of step `pos` (the rank's own count): pos % n_entries; its
    flag becomes pos + 1 when the step's tables are final.  The producer waits before reusing an
    entry the destination has not released (ack)."""
    DATA_OFF = 8192

    def __init__(self, dist, mem, cap, n_entries=16, dst=0, directory=None):
        import mmap
        import os
        self.dist, self.mem, self.cap, self.n, self.dst = dist, mem, int(cap), int(n_entries), int(dst)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        assert 8 * (8 + self.n) <= self.DATA_OFF
        self.size = self.DATA_OFF + self.n * self.cap
        self.dir = directory or ("/dev/shm" if os.path.isdir("/dev/shm") else __import__("tempfile").gettempdir())
        name = [None]
        if self.rank == self.dst:
            name[0] = f"ipb200_tables_{os.getpid()}_{int.from_bytes(os.urandom(4), 'little'):08x}"
        dist.broadcast_object_list(name, src=self.dst)
        self.base = os.path.join(self.dir, name[0])
        self.maps, self.ok = {}, True
        path = f"{self.base}_{self.rank}"
        try:
            fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            try:
                os.posix_fallocate(fd, 0, self.size)           # reserves the pages now: ENOSPC here, not SIGBUS later
                self.maps[self.rank] = mmap.mmap(fd, self.size)
            finally:
                os.close(fd)
            self.mine = np.frombuffer(self.maps[self.rank], dtype=np.uint8)
            self.mine[: self.DATA_OFF] = 0
            self.mine_t = mem.register_host(self.maps[self.rank], self.mine)   # page-locked: the D2H into it is asynchronous
        except (OSError, RuntimeError):
            self.ok = False
        dist.barrier()                            # every rank's segment exists (or its creation failed)
        if self.ok and self.rank == self.dst:
            try:
                for r in range(self.world):
                    if r != self.rank:
                        fd = os.open(f"{self.base}_{r}", os.O_RDWR)
                        try:
                            self.maps[r] = mmap.mmap(fd, self.size)
                        finally:
                            os.close(fd)
            except OSError:                       # a rank on another host, or a segment that could not be made
                self.ok = False
        self.ok = bool(mem.all_reduce_max(0 if self.ok else 1, dist) == 0)     # every rank, or none
        if self.ok and self.rank == self.dst:
            self.segs = [np.frombuffer(self.maps[r], dtype=np.uint8) for r in range(self.world)]
            self.heads = [sg[: self.DATA_OFF].view(np.int64) for sg in self.segs]
            self.next = [0] * self.world          # per producer: the next step expected
            self.released = [0] * self.world      # per producer: steps below this may be overwritten
        dist.barrier()
        try:                                      # every mapping exists: the names can go (no leak if a rank dies)
            os.unlink(path)
        except OSError:
            pass
        if self.ok:
            self.head = self.mine[: self.DATA_OFF].view(np.int64)
            self.run = 0

    # ---- producer side (every rank)
    def entry(self, pos, timeout_s=120.0):
        """(numpy view, page-locked tensor view) of the entry for the rank's step `pos`; waits while
        the destination has not released the step that used the entry before."""
        import time
        t0 = None
        while pos - int(self.head[2]) >= self.n:
            if self.rank == self.dst:
                raise RuntimeError("table ring full on the destination rank: call gathered() at least every "
                                   f"{self.n} steps")
            t0 = t0 or time.perf_counter()
            if time.perf_counter() - t0 > timeout_s:
                raise RuntimeError("table ring full: the destination rank does not collect (gathered())")
            time.sleep(50e-6)
        o = self.DATA_OFF + (pos % self.n) * self.cap
        return self.mine[o: o + self.cap], self.mine_t[o: o + self.cap]

    def publish(self, pos):
        self.head[8 + pos % self.n] = pos + 1

    def end_run(self, n_pos):
        """The rank has published every step below n_pos and will not add to this run."""
        self.run += 1
        self.head[1] = n_pos
        self.head[0] = self.run

    # ---- consumer side (destination rank)
    def release(self, r, upto):
        """Producer r may reuse the entries of its steps below `upto`."""
        self.released[r] = max(self.released[r], int(upto))
        self.heads[r][2] = self.released[r]

    def poll_rank(self, r):
        """Entries of producer r that became final since the last call: [(pos, uint8 view)]."""
        h, got = self.heads[r], []
        while self.next[r] - self.released[r] < self.n and int(h[8 + self.next[r] % self.n]) == self.next[r] + 1:
            o = self.DATA_OFF + (self.next[r] % self.n) * self.cap
            got.append((self.next[r], self.segs[r][o: o + self.cap]))
            self.next[r] += 1
        return got

    def drained(self, run):
        """True when every producer has ended run `run` and all its steps were handed out."""
        return all(int(h[0]) >= run and self.next[r] >= int(h[1]) for r, h in enumerate(self.heads))


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
