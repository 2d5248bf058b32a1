"""World-size-2 gloo test (CPU) of the row-sharded search plumbing in ebsd_vae_b200/sharding.py.

The collectives (uneven and even all-gather of query latents, exchange of the packed per-shard candidates so that a
rank ends up with the lists of its own queries) run for real over gloo; the per-shard search and the merge are done by the oracle here because the
kernels need a GPU.  Checked property: sharded search + merge == one search over the whole dictionary.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ebsd_vae_b200 import sharding
        from oracle import topk_ref as T

        rng = np.random.default_rng(123)
        d_all = T.normalize_rows(rng.normal(size=(3001, 16)).astype(np.float32))
        d_all[2000:2010] = d_all[5]  # ties across the shard boundary
        q_all = T.normalize_rows(rng.normal(size=(37, 16)).astype(np.float32))
        q_all[0] = d_all[5]
        cuts_d = [0, 1200, 3001]        # uneven shards
        cuts_q = [0, 20, 37]            # uneven query split
        shard = d_all[cuts_d[rank]:cuts_d[rank + 1]]
        q_loc = torch.from_numpy(q_all[cuts_q[rank]:cuts_q[rank + 1]])

        counts = sharding.all_gather_counts(len(shard), None)
        assert counts == [1200, 1801]
        index_base = sum(counts[:rank])
        q_counts = sharding.all_gather_counts(q_loc.shape[0], None)
        assert q_counts == [20, 17]
        q_glob = sharding.all_gather_rows(q_loc, q_counts)
        np.testing.assert_array_equal(q_glob.numpy(), q_all)

        dot, idx = T.topk(shard, q_glob.numpy(), 10, index_base=index_base)
        packed = torch.from_numpy(T.pack_candidates(dot, idx))
        got = sharding.exchange_packed(packed, q_counts)
        assert tuple(got.shape) == (world, q_counts[rank], 10)
        sd, si = T.unpack_candidates(got.numpy())
        md, mi = T.topk_merge(sd, si)
        wd, wi = T.topk(d_all, q_all[cuts_q[rank]:cuts_q[rank + 1]], 10)
        np.testing.assert_array_equal(mi, wi)
        np.testing.assert_array_equal(md, wd)

        # equal counts (the data-parallel case) gather without padding
        q_even = torch.from_numpy(q_all[rank * 16:(rank + 1) * 16])
        np.testing.assert_array_equal(sharding.all_gather_rows(q_even, [16, 16]).numpy(), q_all[:32])
        # a rank without rows still takes part
        empty = sharding.all_gather_rows(torch.zeros((0 if rank == 0 else 3, 4)), [0, 3])
        assert tuple(empty.shape) == (3, 4)

        # replicated orientation table
        eul = torch.arange(len(shard) * 3, dtype=torch.float64).reshape(-1, 3) + 10000 * rank
        table = sharding.all_gather_rows(eul, counts)
        assert table.shape == (3001, 3)
        assert float(table[1200, 0]) == 10000.0 and float(table[0, 0]) == 0.0
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_search_plumbing_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
