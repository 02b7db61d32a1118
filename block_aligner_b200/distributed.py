"""Multi-GPU use: one process per GPU, pairs sharded across ranks, no collective on the data path.

Every pair is an independent alignment (the reference's `Block::align` touches only its own scratch), so
the only communication is the host-side gather of the 24-byte AlignResults (plus CIGAR strings when asked), done over
a gloo (CPU) group. A single process that owns several GPUs does not need this module: `ba_align_batch_multi` in the
C ABI shards one batch over the GPUs of the box and delivers the results in order.
Shards are contiguous ranges balanced by total sequence length, so that results concatenate in the original
order.
"""
import numpy as np


def shard_bounds(q_off, r_off, world):
    """Split n pairs into `world` contiguous ranges of roughly equal sum(|q| + |r|). -> int64[world + 1]"""
    n = len(q_off) - 1
    work = (np.asarray(q_off[1:], dtype=np.int64) - np.asarray(q_off[:-1], dtype=np.int64)
            + np.asarray(r_off[1:], dtype=np.int64) - np.asarray(r_off[:-1], dtype=np.int64)) + 64
    csum = np.concatenate([[0], np.cumsum(work)])
    targets = csum[-1] * np.arange(1, world, dtype=np.float64) / world
    cuts = np.searchsorted(csum, targets, side="left")
    return np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)


def slice_batch(q_arena, q_off, r_arena, r_off, lo, hi):
    """Zero-copy views of pairs [lo, hi) (offsets stay absolute; the C ABI rebases them)."""
    return q_arena, q_off[lo:hi + 1], r_arena, r_off[lo:hi + 1]


def align_sharded(aligner, cfg, q_arena, q_off, r_arena, r_off, want_cigars=False):
    """Align this rank's shard and gather all results on every rank.

    Returns (results int64[n, 3], cigars list[str] | None). Works with any initialised torch.distributed
    backend (nccl on GPUs, gloo in the CPU tests); without an initialised group it is a single-rank call."""
    import torch
    import torch.distributed as dist
    from . import api
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    bounds = shard_bounds(q_off, r_off, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    qa, qo, ra, ro = slice_batch(q_arena, q_off, r_arena, r_off, lo, hi)
    b = aligner.upload(cfg, qa, qo, ra, ro)
    try:
        b.run()
        r = b.download()
        mine = np.stack([r["score"].astype(np.int64), r["query_idx"].astype(np.int64), r["reference_idx"].astype(np.int64)],
                        axis=1).reshape(hi - lo, 3)
        cig = [b.cigar_string(k) for k in range(hi - lo)] if want_cigars and (cfg.flags & api.TRACE) else None
    finally:
        b.free()
    if world == 1:
        return mine, cig
    n = len(q_off) - 1
    # host-side gather of the 24-byte results (and CIGAR strings): always over a gloo (CPU) group -- results are already
    # in host memory, and the data path must not depend on NCCL (pairs are independent, there is nothing to exchange)
    group = _host_group()
    sizes = (bounds[1:] - bounds[:-1]).tolist()
    pad = max(sizes)
    buf = torch.zeros((pad, 3), dtype=torch.int64)
    buf[:hi - lo] = torch.from_numpy(mine)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = np.concatenate([p[:s].numpy() for p, s in zip(parts, sizes)], axis=0)
    assert out.shape == (n, 3)
    cigs = None
    if want_cigars and (cfg.flags & api.TRACE):
        gathered = [None] * world
        dist.all_gather_object(gathered, cig, group=group)
        cigs = [c for part in gathered for c in part]
    return out, cigs


_HOST_GROUP = None


def _host_group():
    """A gloo group over all ranks (the default group itself when that already is gloo)."""
    global _HOST_GROUP
    import torch.distributed as dist
    if dist.get_backend() == "gloo":
        return None
    if _HOST_GROUP is None:
        _HOST_GROUP = dist.new_group(backend="gloo")
    return _HOST_GROUP
