"""Multi-GPU plumbing: one process per GPU, recordings sharded in contiguous ranges, no collective
on the data path.  The only exchange is the final all-gather of fixed-width rows (feature table,
per-window Wasserstein values, and — for the mismatched control — the padded audio H1 diagrams of
the reference recordings), over NCCL/NVLink on GPUs and gloo in the CPU tests.

The reference's counterpart is process-level sharding by BATCH_START/BATCH_END + MERGE_PARTIALS
(/root/reference/scripts/tda_eeg_classification_v2.py:55-60, 608-668) and joblib over recordings
(:569-572)."""
from __future__ import annotations


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of n units for `rank`; the first n % world ranks get one extra unit."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allgather_rows(local, n_total, group=None):
    """All-gather row blocks of unequal height (shard_range layout) into (n_total, ...) on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    hmax = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    pad = torch.zeros((hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(out[r * hmax: r * hmax + (hi - lo)])
    return torch.cat(parts, dim=0)


def gather_reference_diagrams(bd_local, counts_local, rec_lo, rec_hi, wanted, group=None):
    """For the mismatched control: every rank needs the audio H1 diagrams of a few recordings that
    may live on other ranks.  bd_local (n_local_rec, items, cap, 2), counts_local (n_local_rec, items);
    `wanted` = sorted global recording ids needed by anyone.  Returns (bd, counts) for `wanted` on
    every rank (rows the rank does not own are filled by the all-reduce of disjoint contributions)."""
    import torch
    import torch.distributed as dist
    idx = torch.as_tensor(wanted, dtype=torch.long, device=bd_local.device)
    bd = torch.zeros((len(wanted),) + tuple(bd_local.shape[1:]), dtype=bd_local.dtype, device=bd_local.device)
    cnt = torch.zeros((len(wanted),) + tuple(counts_local.shape[1:]), dtype=counts_local.dtype,
                      device=counts_local.device)
    mine = (idx >= rec_lo) & (idx < rec_hi)
    if mine.any():
        bd[mine] = bd_local[idx[mine] - rec_lo]
        cnt[mine] = counts_local[idx[mine] - rec_lo]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        # each wanted row has exactly one owner: a sum over ranks is a gather (+inf never appears in
        # H1 births; deaths may be +inf, and inf + 0 = inf is exact)
        dist.all_reduce(bd, group=group)
        dist.all_reduce(cnt, group=group)
    return bd, cnt
