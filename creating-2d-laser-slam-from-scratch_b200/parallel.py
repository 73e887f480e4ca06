"""Multi-GPU plumbing for the batched hot path: scan-matches are independent units, so a batch is
partitioned by index over the ranks with NO data-path collective; only results are gathered.  One process per GPU,
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used for the gather and for barriers/timing."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced partition: the first (n_items % world) ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(array, world: int, rank: int):
    lo, hi = shard_bounds(len(array), world, rank)
    return array[lo:hi]


def gather_results(local: np.ndarray, n_total: int, group=None) -> np.ndarray:
    """All-gather per-match result rows ([n_local, k] float64) back into batch order on every rank.  Shards may have
    different lengths (ragged): rows are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        assert len(local) == n_total
        return np.asarray(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    k = local.shape[1] if local.ndim == 2 else 1
    loc = np.asarray(local, dtype=np.float64).reshape(-1, k)
    max_len = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    pad = np.zeros((max_len, k))
    pad[:len(loc)] = loc
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    rows = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        rows.append(outs[r][:hi - lo].cpu().numpy())
    return np.concatenate(rows, axis=0) if rows else np.zeros((0, k))


def max_over_ranks(value: float, group=None) -> float:
    """Timing rule: report the MAX over ranks."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())



def _all_reduce(arr: np.ndarray, op: str, group=None) -> np.ndarray:
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return arr
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def correlate_scan_angle_split(matcher, centers, search, n_angles: int, group=None):
    """One coarse CorrelateScan (Mapper.cpp:315-523) whose ANGLES are split over the ranks: every
    rank holds the same scans + grids, sweeps its contiguous share of the `n_angles` angle steps, and three small
    all-reduces (MAX on best response / per-cell maxima / status, SUM on the tie sums) stand where the reference's
    single-threaded max + tie average run.  Returns the same tuple as ScanMatcher.correlate_scan on every rank."""
    import torch.distributed as dist

    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if dist.is_available() and dist.is_initialized() else (1, 0)
    lo, hi = shard_bounds(n_angles, world, rank)
    best, probs, status = matcher.split_begin(centers, search, lo, hi - lo)
    best = _all_reduce(best, "max", group)
    probs = _all_reduce(probs, "max", group)
    status = _all_reduce(status.astype(np.float64), "max", group).astype(np.int32)
    ties = _all_reduce(matcher.split_ties(best), "sum", group)
    out = matcher.split_finish(best, ties, probs)
    if (status != 0).any():  # a rank saw an out-of-range candidate: the reference throws for the whole sweep
        out[3][status != 0] = status[status != 0]
    return out


def occupancy_grid_sharded(occgrid_mod, laser, ranges_shard, poses_shard, resolution: float, device: int = 0, group=None,
                           nccl_comm=None):
    """karto::OccupancyGrid::CreateFromScans (Karto.h:5659-5673) over a scan list SHARDED across the ranks
    (SURVEY.md §8(e)(iii)): bounding boxes are reduced (MIN/MAX), each rank ray-traces its shard into the globally
    dimensioned grid, the uint32 pass/hit counters are summed with one all-reduce each (in place on the device with
    NCCL; through host arrays with gloo), then every rank thresholds the same cells."""
    import torch
    import torch.distributed as dist

    live = dist.is_available() and dist.is_initialized()
    bbox = occgrid_mod.scans_bbox(laser, ranges_shard, poses_shard, device=device)
    if live:
        lo = _all_reduce(-bbox[:2], "max", group)  # MIN via MAX of the negation
        hi = _all_reduce(bbox[2:], "max", group)
        bbox = np.concatenate([-lo, hi])
    g = occgrid_mod.OccupancyGrid(laser, ranges_shard, poses_shard, resolution, device=device, bbox=bbox)
    if live and g.info.data_size > 0:
        if nccl_comm is not None:  # the collective inside the library: ncclAllReduce(SUM) on the device counters
            g.allreduce_counters(nccl_comm)
        elif dist.get_backend(group) == "nccl":
            for arr in g.device_counters():
                t = torch.as_tensor(arr, device=torch.device("cuda", device))
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            torch.cuda.synchronize(device)
        else:
            a = g.arrays()
            pa = _all_reduce(a["passes"].astype(np.int64), "sum", group)
            hi_ = _all_reduce(a["hits"].astype(np.int64), "sum", group)
            g.set_counters(pa, hi_)
    g.update()
    return g


# ---- a raw NCCL communicator for the in-library collectives (b2s_matcher_correlate_scan_split, b2s_occ_grid_allreduce_counters)
class NcclComm:
    """ncclComm_t created with ncclGetUniqueId / ncclCommInitRank through ctypes; the 128-byte id travels over the
    already-initialised torch.distributed group.  A C++ host does the same with its own bootstrap."""

    def __init__(self, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist

        self.C = C
        self.lib = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_char * 128)]
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        uid = UniqueId()
        if self.rank == 0:
            rc = self.lib.ncclGetUniqueId(C.byref(uid))
            assert rc == 0, rc
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone().to(dev)
        dist.broadcast(t, 0, group=group)
        raw = bytes(t.cpu().numpy().tobytes())
        C.memmove(C.byref(uid), raw, 128)
        self.comm = C.c_void_p()
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        rc = self.lib.ncclCommInitRank(C.byref(self.comm), self.world, uid, self.rank)
        assert rc == 0, rc

    def close(self):
        if getattr(self, "comm", None) and self.comm.value:
            self.lib.ncclCommDestroy.argtypes = [self.C.c_void_p]
            self.lib.ncclCommDestroy(self.comm)
            self.comm = self.C.c_void_p()


def correlate_scan_split_in_library(matcher, centers, search, comm: "NcclComm"):
    """b2s_matcher_correlate_scan_split: the same angle-split sweep with the all-reduces issued by the library itself on
    device buffers (no host staging).  Returns (results tuple, collective timings in ms)."""
    import ctypes as C
    from . import abi
    from .matcher import check, results_to_arrays

    L = matcher.L
    L.b2s_matcher_correlate_scan_split.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double),
                                                   C.POINTER(abi.Search), C.POINTER(abi.MatchResult)]
    L.b2s_matcher_last_split_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    c = np.ascontiguousarray(centers, np.float64).reshape(matcher.batch, 3)
    res = (abi.MatchResult * matcher.batch)()
    check(L.b2s_matcher_correlate_scan_split(matcher.h, comm.comm, comm.rank, comm.world, c.ctypes.data_as(C.POINTER(C.c_double)),
                                             C.byref(search), res))
    t = np.zeros(4)
    check(L.b2s_matcher_last_split_timing(matcher.h, t.ctypes.data_as(C.POINTER(C.c_double))))
    return results_to_arrays(res, matcher.batch), dict(allreduce_best_ms=t[0], allreduce_plane_ms=t[1], allreduce_ties_ms=t[2],
                                                        after_sweep_ms=t[3])
