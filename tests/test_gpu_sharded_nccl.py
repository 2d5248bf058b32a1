"""Hardware parity of the NCCL row-sharded dictionary (ebsd_vae_b200/sharding.py): two ranks on two GPUs.

Runs only where at least two CUDA devices are visible (`gpurun --gpus 2`); on a single-GPU box the test is skipped --
bench.py asserts the same equality after its timed region whenever it runs with N > 1.
Checked: search_global / query_similar / find_best_orientations_batch of the sharded database equal a single-GPU
LatentVectorDatabase over the concatenated dictionary bit for bit (the reference contract is one exact list per
query, latice/index/faiss_db.py:241-256), with uneven shards, uneven query counts, a rank without queries, ties
across the shard boundary, and persistence of the shards.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import ebsd_vae_b200 as E
        from ebsd_vae_b200.sharding import ShardedLatentVectorDatabase

        rng = np.random.default_rng(5)
        n_all, k = 150_001, 10
        lat = rng.normal(size=(n_all, 16)).astype(np.float32)
        lat[90_000:90_020] = lat[7]                   # ties across the shard boundary (cut at 90 010)
        eul = rng.uniform(0, 1, size=(n_all, 3)) * np.array([360.0, 180.0, 360.0])
        cut = 90_010
        mine = slice(0, cut) if rank == 0 else slice(cut, n_all)
        cfg = E.LatentVectorDatabaseConfig(persist_directory=os.path.join(out_dir, "store"), collection_name="nccl")
        db = ShardedLatentVectorDatabase(cfg)
        db.add_vectors(lat[mine], eul[mine])
        assert db.get_global_count() == n_all and db.index_base == (0 if rank == 0 else cut)

        single = E.LatentVectorDatabase(E.LatentVectorDatabaseConfig(persist_directory=None))
        single.add_vectors(lat, eul)

        for q_counts in ([3000, 2100], [2500, 0], [1, 1]):     # screen path, a rank without queries, stream path
            nq = q_counts[rank]
            a = sum(q_counts[:rank])
            qs_all = lat[rng.integers(0, n_all, size=sum(q_counts))] + 0.05 * rng.normal(size=(sum(q_counts), 16)).astype(np.float32)
            qs_all[0] = lat[7]
            qs = qs_all[a : a + nq]
            qh = db._prepare_queries(torch.from_numpy(qs.reshape(-1, 16)))
            for counts in (q_counts, None):                      # with and without the caller-supplied counts
                dot_g, idx_g, dist_g = db.search_global(qh, k, counts)
                dot_1, idx_1, dist_1 = single.search_device(single._prepare_queries(torch.from_numpy(qs.reshape(-1, 16))), k)
                assert torch.equal(idx_g, idx_1) and torch.equal(dot_g, dot_1) and torch.equal(dist_g, dist_1)
            if rank == 0 and q_counts[0] >= 2500:
                assert idx_g[0].tolist() == [7] + list(range(90_000, 90_009))   # lowest global rows win the tie
            res_g = db.find_best_orientations_batch(qs.reshape(-1, 16), top_n=k, orientation_threshold=3.0,
                                                    min_required_matches=3, q_counts=q_counts)
            res_1 = single.find_best_orientations_batch(qs.reshape(-1, 16), top_n=k, orientation_threshold=3.0,
                                                        min_required_matches=3)
            np.testing.assert_array_equal(res_g.indices, res_1.indices)
            np.testing.assert_array_equal(res_g.success, res_1.success)
            np.testing.assert_array_equal(res_g.candidate_orientations, res_1.candidate_orientations)
            np.testing.assert_array_equal(res_g.mean_orientations, res_1.mean_orientations)
            np.testing.assert_array_equal(res_g.similar_masks, res_1.similar_masks)

        # replicated rows: every rank holds all normalised rows and searches only its own queries, without any
        # per-batch collective -- same global rows, same lists, bit for bit
        db.replicate_rows()
        assert tuple(db._replica.shape) == (n_all, 16) and torch.equal(db._replica, single._latents[:n_all])
        for q_counts in ([3000, 2100], [2500, 0], [1, 1]):
            nq, a = q_counts[rank], sum(q_counts[:rank])
            qs = (lat[rng.integers(0, n_all, size=sum(q_counts))] + 0.05 * rng.normal(size=(sum(q_counts), 16)).astype(np.float32))[a : a + nq]
            qh = db._prepare_queries(torch.from_numpy(qs.reshape(-1, 16)))
            dot_g, idx_g, dist_g = db.search_global(qh, k)      # no counts needed: nothing is exchanged
            dot_1, idx_1, dist_1 = single.search_device(single._prepare_queries(torch.from_numpy(qs.reshape(-1, 16))), k)
            assert torch.equal(idx_g, idx_1) and torch.equal(dot_g, dot_1) and torch.equal(dist_g, dist_1)
            res_g = db.find_best_orientations_batch(qs.reshape(-1, 16), top_n=k, orientation_threshold=3.0, min_required_matches=3)
            res_1 = single.find_best_orientations_batch(qs.reshape(-1, 16), top_n=k, orientation_threshold=3.0, min_required_matches=3)
            np.testing.assert_array_equal(res_g.indices, res_1.indices)
            np.testing.assert_array_equal(res_g.mean_orientations, res_1.mean_orientations)
        db._replica, db.replicate = None, False             # back to the row-sharded search for the rest of the test

        # query_similar is global too (collective)
        got = db.query_similar(lat[7], n_results=5)
        want = single.query_similar(lat[7], n_results=5)
        assert got["ids"] == want["ids"] and got["metadatas"] == want["metadatas"]

        # every rank persists its shard; load() republishes counts and the replicated orientation tables
        path = db.save()
        assert f"shard{rank}of{world}" in path.name
        db.delete_collection()
        assert db.get_global_count() == 0
        db.load()
        assert db.get_global_count() == n_all and db.index_base == (0 if rank == 0 else cut)
        qh = db._prepare_queries(torch.from_numpy(lat[5:6] if rank == 0 else lat[100_000:100_001]))
        _, idx_g, _ = db.search_global(qh, k, [1, 1])
        assert int(idx_g[0, 0]) == (5 if rank == 0 else 100_000)
        torch.cuda.synchronize()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_nccl_sharded_search_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2); bench.py asserts the same equality at every N > 1")
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
