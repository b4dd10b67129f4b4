"""Host-side sharding of the path over the GPUs of one node (one process per GPU).

Units (lattice points of a scan, poses of a screen, MC chains) are independent (src/lds.ml:2009,
2050: the reference forks one process per ligand / start), so each rank takes a contiguous block and
there is no data-path collective.  The only exchange is the all-gather of the per-rank top-k lists,
merged with the reference's tie rule by libmmo_b200's mmo_topk_merge.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_range(n_units: int, rank: int, world: int):
    """Contiguous block [first, first+count) of rank `rank`; the blocks partition range(n_units)."""
    base, rem = divmod(n_units, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def merge_lists(lib, k, score_lists, frame_lists):
    """mmo_topk_merge over python lists of per-rank (scores, frames)."""
    n = len(score_lists)
    S = np.full((n, k), np.inf)
    F = np.full((n, k), -1, np.int64)
    cnt = np.zeros(n, np.int32)
    for i, (s, f) in enumerate(zip(score_lists, frame_lists)):
        m = min(k, len(s))
        S[i, :m] = s[:m]; F[i, :m] = f[:m]; cnt[i] = m
    out_s, out_f = np.empty(k), np.empty(k, np.int64)
    out_n = C.c_int32()
    rc = lib.mmo_topk_merge(C.c_int32(n), C.c_int32(k), S.ctypes.data_as(C.POINTER(C.c_double)),
                            F.ctypes.data_as(C.POINTER(C.c_int64)), cnt.ctypes.data_as(C.POINTER(C.c_int32)),
                            out_s.ctypes.data_as(C.POINTER(C.c_double)), out_f.ctypes.data_as(C.POINTER(C.c_int64)),
                            C.byref(out_n))
    if rc != 0:
        raise RuntimeError(lib.mmo_last_error().decode())
    return out_s[:out_n.value].copy(), out_f[:out_n.value].copy()


def allgather_topk(dist, lib, k, scores, frames, device=None):
    """torch.distributed flavour of the exchange (gloo on CPU in the tests, NCCL in bench.py):
    all-gather k x (f64, i64) per rank, then the library merge.  Returns (scores, frames)."""
    import torch
    world = dist.get_world_size()
    s = torch.full((k,), float("inf"), dtype=torch.float64, device=device)
    f = torch.full((k,), -1, dtype=torch.int64, device=device)
    n = min(k, len(scores))
    s[:n] = torch.as_tensor(np.asarray(scores[:n], np.float64), device=device)
    f[:n] = torch.as_tensor(np.asarray(frames[:n], np.int64), device=device)
    cnt = torch.tensor([n], dtype=torch.int64, device=device)
    all_s = [torch.empty_like(s) for _ in range(world)]
    all_f = [torch.empty_like(f) for _ in range(world)]
    all_n = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(all_s, s)
    dist.all_gather(all_f, f)
    dist.all_gather(all_n, cnt)
    sl = [a.cpu().numpy()[:int(c.item())] for a, c in zip(all_s, all_n)]
    fl = [a.cpu().numpy()[:int(c.item())] for a, c in zip(all_f, all_n)]
    return merge_lists(lib, k, sl, fl)
