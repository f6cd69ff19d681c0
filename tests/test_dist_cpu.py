"""CPU-only, world_size 2 over gloo: the sharding / gathering logic of the multi-GPU path."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tda_eeg_audio_b200.dist import allgather_rows, gather_reference_diagrams, shard_range
from tda_eeg_audio_b200.pipeline import mismatch_reference_recording


def test_shard_range_partitions():
    for n in (0, 1, 7, 1416):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _table(rec_ids):
    r = torch.as_tensor(rec_ids, dtype=torch.float64)
    return r[:, None] * 1000 + torch.arange(220, dtype=torch.float64)[None, :]


def _worker(rank, world, port, n_rec, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_rec, rank, world)
        full = allgather_rows(_table(range(lo, hi)), n_rec)
        ok1 = torch.equal(full, _table(range(n_rec)))
        # mismatched control: audio H1 diagrams of the reference recordings
        ref = mismatch_reference_recording(n_rec, n_subjects=3)
        wanted = sorted(set(int(x) for x in ref if x >= 0))
        bd_local = torch.zeros((hi - lo, 5, 4, 2), dtype=torch.float32)
        cnt_local = torch.zeros((hi - lo, 5), dtype=torch.int32)
        for k, r in enumerate(range(lo, hi)):
            bd_local[k] = r + 0.5
            bd_local[k, 0, 0, 1] = float("inf")
            cnt_local[k] = r % 4
        bd, cnt = gather_reference_diagrams(bd_local, cnt_local, lo, hi, wanted)
        ok2 = all(float(bd[i, 1, 1, 0]) == w + 0.5 and int(cnt[i, 0]) == w % 4 and np.isinf(float(bd[i, 0, 0, 1]))
                  for i, w in enumerate(wanted))
        q.put((rank, bool(ok1), bool(ok2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rec", [10, 11])
def test_allgather_and_reference_exchange_world2(n_rec):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_rec
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rec, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True, True), (1, True, True)]
