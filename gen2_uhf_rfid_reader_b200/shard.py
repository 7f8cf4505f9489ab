"""Sharding of capture segments over ranks + the one collective of the system.

Capture segments (inventory rounds / slots) are independent units: rank r decodes
segments [r*S/G, (r+1)*S/G) of the global table with no exchange during compute.  The only collective is
one all-gather of fixed-size `rfid_b200_window_result` records (+ per-segment window counts) at the end,
after which any rank can reduce READER_STATS.  Works on NCCL (CUDA tensors) and gloo (CPU tensors).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import abi


def shard_range(n_segments: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) of segment ids owned by `rank` (first n%world ranks get one more)."""
    base, extra = divmod(n_segments, world)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return begin, end


def max_shard(n_segments: int, world: int) -> int:
    return (n_segments + world - 1) // world


def gather_records(results: torch.Tensor, counts: torch.Tensor, n_segments: int, max_windows: int, group=None,
                   out_results: torch.Tensor = None, out_counts: torch.Tensor = None):
    """All-gather the per-rank record blocks.

    results: uint8 [n_local*max_windows, 64] (device or CPU), counts: int32 [n_local]; every rank passes its own
    shard (`shard_range`).  Shards are padded to the largest shard so the collective has equal counts.
    Returns (results uint8 [n_segments*max_windows, 64], counts int32 [n_segments]) in global segment order.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return results, counts
    rank = dist.get_rank(group)
    pad = max_shard(n_segments, world)
    b, e = shard_range(n_segments, rank, world)
    n_local = e - b
    dev = results.device
    send_r = torch.zeros((pad * max_windows, 64), dtype=torch.uint8, device=dev)
    send_c = torch.zeros(pad, dtype=torch.int32, device=dev)
    send_r[: n_local * max_windows] = results[: n_local * max_windows]
    send_c[:n_local] = counts[:n_local]
    all_r = out_results if out_results is not None else torch.empty((world * pad * max_windows, 64), dtype=torch.uint8, device=dev)
    all_c = out_counts if out_counts is not None else torch.empty(world * pad, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_r, send_r, group=group)
    dist.all_gather_into_tensor(all_c, send_c, group=group)
    # strip the padding
    keep_r, keep_c = [], []
    for r in range(world):
        rb, re = shard_range(n_segments, r, world)
        keep_r.append(all_r[r * pad * max_windows:(r * pad + (re - rb)) * max_windows])
        keep_c.append(all_c[r * pad: r * pad + (re - rb)])
    return torch.cat(keep_r), torch.cat(keep_c)


def records_numpy(results: torch.Tensor, counts: torch.Tensor, max_windows: int):
    r = results.cpu().numpy().reshape(-1).view(abi.RESULT_DTYPE)
    c = counts.cpu().numpy()
    return r.reshape(c.size, max_windows), c


def renumber_segments(recs: np.ndarray, first_segment: int):
    """records carry shard-local segment indices; shift them to global ids (in place)."""
    recs["segment"] += first_segment
    return recs
