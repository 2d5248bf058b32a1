"""Row-sharded dictionary over the GPUs of one box (one process per GPU, torch.distributed / NCCL).

The dictionary shards by row: rank g holds the rows it encoded, ``index_base`` = number of
rows on lower ranks, and the small orientation table is replicated.  A query batch is split data-parallel:

    encode Q/G patterns  ->  all-gather latents (Q x 64 B)  ->  local exact top-k of ALL Q queries on the shard
    ->  all-gather the per-shard candidates (dot, global row)  ->  k-way merge of the own Q/G queries
    ->  consensus on the own queries.

The merge is order independent (ties break on the global row index), so results equal a single-GPU search
bit for bit.  The collectives are plain NCCL calls issued on the compute stream; the exchanged volume
(Q*k*12 B per rank) is tiny next to the search itself.

The plumbing below is backend-agnostic (NCCL on GPUs, gloo in the CPU tests); the search / merge / consensus
calls are the native kernels and need a GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _native
from .vector_db import _to_host, LatentVectorDatabase, LatentVectorDatabaseConfig, OrientationResultBatch


def all_gather_counts(n: int, group=None, device="cpu") -> list[int]:
    world = dist.get_world_size(group)
    mine = torch.tensor([n], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [int(t.item()) for t in out]


def all_gather_rows(x: torch.Tensor, counts: list[int], group=None) -> torch.Tensor:
    """All-gather tensors that differ in their first dimension (``counts[r]`` rows on rank r); returns the concatenation."""
    world = dist.get_world_size(group)
    nmax = max(counts) if counts else 0
    if nmax == 0:
        return x.new_zeros((0,) + tuple(x.shape[1:]))
    padded = x.new_zeros((nmax,) + tuple(x.shape[1:]))
    padded[: x.shape[0]] = x
    out = x.new_empty((world, nmax) + tuple(x.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group) if x.is_cuda else dist.all_gather(
        list(out.unbind(0)), padded.contiguous(), group=group)
    return torch.cat([out[r, : counts[r]] for r in range(world)], dim=0)


def exchange_candidates(dot: torch.Tensor, idx: torch.Tensor, q_counts: list[int], group=None):
    """Every rank holds candidates [Q_global,k] from its shard; return, for the OWN query slice, the stack
    [world, Q_own, k] of all shards' candidates (all-gather as in the north star, then slice)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    all_dot = dot.new_empty((world,) + tuple(dot.shape))
    all_idx = idx.new_empty((world,) + tuple(idx.shape))
    if dot.is_cuda:
        dist.all_gather_into_tensor(all_dot, dot.contiguous(), group=group)
        dist.all_gather_into_tensor(all_idx, idx.contiguous(), group=group)
    else:
        dist.all_gather(list(all_dot.unbind(0)), dot.contiguous(), group=group)
        dist.all_gather(list(all_idx.unbind(0)), idx.contiguous(), group=group)
    a = sum(q_counts[:rank])
    b = a + q_counts[rank]
    return all_dot[:, a:b].contiguous(), all_idx[:, a:b].contiguous()


class ShardedLatentVectorDatabase(LatentVectorDatabase):
    """``LatentVectorDatabase`` whose rows are spread over the ranks of a process group.

    ``add_vectors`` takes the rows of THIS rank (e.g. the patterns it encoded in ``build_dictionary``);
    ``find_best_orientations_batch`` takes the queries of THIS rank and returns their results.
    """

    def __init__(self, config: LatentVectorDatabaseConfig | None = None, group=None) -> None:
        super().__init__(config)
        if not dist.is_initialized():
            raise RuntimeError("ShardedLatentVectorDatabase needs torch.distributed to be initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._global_eulers: torch.Tensor | None = None
        self._global_quats: torch.Tensor | None = None
        self._shard_counts = [0] * self.world

    def add_vectors(self, latent_vectors, orientations, batch_size: int = 1000) -> None:
        if self._count:
            raise RuntimeError("ShardedLatentVectorDatabase is built by one add_vectors call per rank")
        super().add_vectors(latent_vectors, orientations, batch_size)
        dev = self._dev()
        self._shard_counts = all_gather_counts(self._count, self.group, dev)
        self.index_base = sum(self._shard_counts[: self.rank])
        # replicate the orientation table (24 + 32 bytes per row)
        self._global_eulers = all_gather_rows(self._eulers[: self._count], self._shard_counts, self.group)
        self._global_quats = all_gather_rows(self._quats[: self._count], self._shard_counts, self.group)

    def _global_count(self) -> int:
        return sum(self._shard_counts)

    def _orientation_tables(self):
        return self._global_eulers, self._global_quats, 0

    def search_global(self, q_hat_local: torch.Tensor, k: int):
        """Exact global top-k for this rank's normalised queries: (dot, idx, dist), each [Q_local,k]."""
        dev = self._dev()
        lib = _native.load()
        q_counts = all_gather_counts(q_hat_local.shape[0], self.group, dev)
        q_all = all_gather_rows(q_hat_local, q_counts, self.group)
        dot, idx, _ = self.search_device(q_all.contiguous(), k)
        sd, si = exchange_candidates(dot, idx, q_counts, self.group)
        nq = q_hat_local.shape[0]
        out_dot = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        out_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
        if nq:
            with torch.cuda.device(dev):
                _native.check(
                    lib.ebsd_topk_merge(sd.data_ptr(), si.data_ptr(), self.world, nq, k, out_dot.data_ptr(),
                                        out_idx.data_ptr(), out_dist.data_ptr(), self._stream(dev)),
                    "ebsd_topk_merge")
        return out_dot, out_idx, out_dist

    def find_best_orientations_batch(self, query_vectors, batch_size: int = 32, top_n: int = 20,
                                     orientation_threshold: float = 1.0, min_required_matches: int = 18,
                                     max_iterations: int = 3) -> OrientationResultBatch:
        k = self._clamp_k(top_n)
        q_in = torch.as_tensor(query_vectors)
        if q_in.dim() == 1:
            q_in = q_in[None]
        q = self._prepare_queries(q_in)
        _, idx, dist_ = self.search_global(q, k)
        if self.config.mode == "chroma" and min(self._global_count(), k) < max_iterations:
            raise IndexError("top_n candidates fewer than max_iterations (chroma_db.py:302-303)")
        _, mean_e, success, mask, _, cand = self.consensus_device(idx, orientation_threshold, min_required_matches,
                                                                  max_iterations)
        qv, idx_h, dist_h, cand_h, succ_h, mean_h, mask_h = _to_host(q_in.detach(), idx, dist_, cand, success, mean_e, mask)
        return OrientationResultBatch(
            query_vectors=qv, indices=idx_h, distances=dist_h, candidate_orientations=cand_h,
            success=succ_h.astype(bool), mean_orientations=mean_h, similar_masks=mask_h.astype(np.uint64),
            faiss_mode=self.config.mode == "faiss")
