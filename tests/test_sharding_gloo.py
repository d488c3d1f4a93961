"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU path."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cooking_zoo_b200 import _native
from cooking_zoo_b200.sharding import shard_range, shard_rows, global_layout_ids, reduce_stats


def test_shard_range_partitions():
    for total in (1, 7, 8, 131072, 1_048_576 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o1, c1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + c1 == o2


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _native.load_library()
    offset, count = shard_range(total, world, rank)
    lids = global_layout_ids(lib.cz_layout_draw, 99, offset, count, 400)
    recipes = np.arange(total * 2).reshape(total, 2) % 8
    mine = shard_rows(recipes, world, rank)
    assert mine.shape == (count, 2) and (mine == recipes[offset:offset + count]).all()
    # statistics: every rank contributes its shard; the reduced vector must equal the global sum
    stats = torch.tensor([float(count), float(lids.sum()), float(mine.sum())], dtype=torch.float64)
    red = reduce_stats(stats)
    if rank == 0:
        np.savez(out, red=red.numpy(), lids0=lids)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    total, world = 1000, 2
    out = str(tmp_path / "r0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, total, out), nprocs=world, join=True)
    got = np.load(out)
    lib = _native.load_library()
    whole = global_layout_ids(lib.cz_layout_draw, 99, 0, total, 400)
    recipes = np.arange(total * 2).reshape(total, 2) % 8
    want = torch.tensor([float(total), float(whole.sum()), float(recipes.sum())], dtype=torch.float64)
    assert np.array_equal(got["red"], want.numpy())
    # the shard of rank 0 is a prefix of the single-process draw: results do not depend on G
    assert (got["lids0"] == whole[:len(got["lids0"])]).all()
