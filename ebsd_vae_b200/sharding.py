"""Row-sharded dictionary over the GPUs of one box (one process per GPU, torch.distributed / NCCL).

The dictionary shards by row: rank g holds the rows it encoded, ``index_base`` = number of
rows on lower ranks, and the small orientation table is replicated.  A query batch is split data-parallel:

    encode Q/G patterns  ->  all-gather latents (Q x 64 B)  ->  local exact top-k of ALL Q queries on the shard
    ->  all-to-all of the per-shard candidates, packed (dot, global row) in one word each, so that a rank receives
        only the lists of its own Q/G queries  ->  k-way merge  ->  consensus on the own queries.

The merge is order independent (ties break on the global row index), so results equal a single-GPU search
bit for bit (asserted on hardware by bench.py at N > 1 and tests/test_gpu_sharded_nccl.py).  The collectives are
plain NCCL calls issued on the compute stream, without host synchronisation when the caller passes the per-rank query
counts; the exchanged volume (Q*k*8 B per rank) is tiny next to the search itself.

``replicate=True`` (or ``replicate_rows()``) trades memory for the exchange: every rank fetches ALL normalised rows
once at build time (64 B per row; 10 M rows = 640 MB of a 180 GB GPU) and a query batch then needs no collective at
all -- each rank searches only its own queries against the full dictionary.  The pair count per rank is the same
(Q/G x N instead of Q x N/G) but the search runs in its better shape (long dictionary, fewer queries) and the global
row numbers, hence the results, are the same bit for bit.  The default stays row-sharded.

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
    """Row / query counts of every rank: one collective and ONE host synchronisation."""
    world = dist.get_world_size(group)
    mine = torch.tensor([n], dtype=torch.int64, device=device)
    out = torch.zeros((world,), dtype=torch.int64, device=device)
    if mine.is_cuda:
        dist.all_gather_into_tensor(out, mine, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), mine[0], group=group)
    return [int(v) for v in out.tolist()]


def all_gather_rows(x: torch.Tensor, counts: list[int], group=None) -> torch.Tensor:
    """All-gather tensors that differ in their first dimension (``counts[r]`` rows on rank r); returns the
    concatenation.  Equal counts (the data-parallel case) gather straight into the result: no padding, no copy."""
    world = dist.get_world_size(group)
    nmax = max(counts) if counts else 0
    if nmax == 0:
        return x.new_zeros((0,) + tuple(x.shape[1:]))
    uniform = all(c == nmax for c in counts)
    src = x.contiguous()
    if not uniform:
        src = x.new_zeros((nmax,) + tuple(x.shape[1:]))
        src[: x.shape[0]] = x
    out = x.new_empty((world, nmax) + tuple(x.shape[1:]))
    if x.is_cuda:
        dist.all_gather_into_tensor(out, src, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), src, group=group)
    if uniform:
        return out.view((world * nmax,) + tuple(x.shape[1:]))
    return torch.cat([out[r, : counts[r]] for r in range(world)], dim=0)


def exchange_packed(packed: torch.Tensor, q_counts: list[int], group=None) -> torch.Tensor:
    """``packed`` [Q_global,k]: this shard's candidates for EVERY query.  Returns [world, Q_own, k]: every shard's
    candidates for the OWN query slice -- an all-to-all, so a rank receives 1/world of what an all-gather moves."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    k = packed.shape[1]
    out = packed.new_empty((world, q_counts[rank], k))
    if packed.is_cuda:
        dist.all_to_all_single(out.view(world * q_counts[rank], k), packed.contiguous(),
                               output_split_sizes=[q_counts[rank]] * world, input_split_sizes=list(q_counts), group=group)
    else:  # gloo (CPU tests) has no all_to_all_single: all-gather and slice
        bufs = [packed.new_empty((sum(q_counts), k)) for _ in range(world)]
        dist.all_gather(bufs, packed.contiguous(), group=group)
        a = sum(q_counts[:rank])
        for r in range(world):
            out[r] = bufs[r][a : a + q_counts[rank]]
    return out


class ShardedLatentVectorDatabase(LatentVectorDatabase):
    """``LatentVectorDatabase`` whose rows are spread over the ranks of a process group.

    ``add_vectors`` takes the rows of THIS rank (e.g. the patterns it encoded in ``build_dictionary``);
    ``query_similar`` / ``find_best_orientation(s_batch)`` take the queries of THIS rank and return their GLOBAL
    results; they are collective calls (every rank must make them, possibly with zero queries).
    """

    def __init__(self, config: LatentVectorDatabaseConfig | None = None, group=None, replicate: bool = False) -> None:
        if not dist.is_initialized():
            raise RuntimeError("ShardedLatentVectorDatabase needs torch.distributed to be initialised")
        self.group = group
        self.replicate = bool(replicate)
        self._replica: torch.Tensor | None = None   # [N_global,16] normalised rows of every shard, in global row order
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._global_eulers: torch.Tensor | None = None
        self._global_quats: torch.Tensor | None = None
        self._shard_counts = [0] * self.world
        super().__init__(config)

    def _reopen_on_init(self) -> bool:
        return False   # a shard file is reopened explicitly with load() (collective), never implicitly

    @property
    def npz_path(self):
        """Every rank persists its own shard: ``<persist_directory>/<collection_name>.shard<rank>of<world>.npz``."""
        base = super().npz_path
        return base.with_name(f"{self.collection_name}.shard{self.rank}of{self.world}.npz")

    # ------------------------------------------------------------------ population (collective)
    def _publish(self) -> None:
        """Refresh what every rank must know about all shards: row counts, ``index_base`` and the replicated
        orientation tables (24 + 32 bytes per row).  Collective."""
        dev = self._dev()
        self._shard_counts = all_gather_counts(self._count, self.group, dev)
        self.index_base = sum(self._shard_counts[: self.rank])
        if self._count:
            eul, qua = self._eulers[: self._count], self._quats[: self._count]
        else:   # a rank without rows still takes part in the collectives
            eul = torch.zeros((0, 3), dtype=torch.float64, device=dev)
            qua = torch.zeros((0, 4), dtype=torch.float64, device=dev)
        self._global_eulers = all_gather_rows(eul, self._shard_counts, self.group)
        self._global_quats = all_gather_rows(qua, self._shard_counts, self.group)
        self._replica = None
        if self.replicate:
            self.replicate_rows()

    def replicate_rows(self) -> None:
        """Collective: fetch the normalised rows of every shard (one all-gather of 64 B per row).  Afterwards a query
        batch is searched against the full dictionary on the rank that owns it, with no per-batch collective."""
        dev = self._dev()
        mine = (self._latents[: self._count] if self._count
                else torch.zeros((0, self.dimension), dtype=torch.float32, device=dev))
        self._replica = all_gather_rows(mine, self._shard_counts, self.group).contiguous()
        self.replicate = True

    def add_vectors(self, latent_vectors, orientations, batch_size: int = 1000) -> None:
        if self._count:
            raise RuntimeError("ShardedLatentVectorDatabase is built by one add_vectors call per rank")
        super().add_vectors(latent_vectors, orientations, batch_size)
        self._publish()

    def load(self, path=None) -> None:
        super().load(path)
        self._publish()

    def delete_collection(self) -> None:
        super().delete_collection()
        self._shard_counts = [0] * self.world
        self.index_base = 0
        self._global_eulers = self._global_quats = None
        self._replica = None

    def _global_count(self) -> int:
        return sum(self._shard_counts)

    def get_global_count(self) -> int:
        return self._global_count()

    def _orientation_tables(self):
        if self._global_quats is None or self._global_quats.shape[0] == 0:
            return None, None, 0
        return self._global_eulers, self._global_quats, 0

    # ------------------------------------------------------------------ search (collective)
    def search_global(self, q_hat_local: torch.Tensor, k: int, q_counts: list[int] | None = None):
        """Exact global top-k for this rank's normalised queries: (dot, idx, dist), each [Q_local,k].

        ``q_counts`` = the number of queries on every rank when the caller knows it (data-parallel batches of a
        fixed size); without it the counts are all-gathered first, which costs one host synchronisation.
        Steps: all-gather the latents (Q x 64 B) -> local exact top-k of ALL queries on this shard -> pack (dot, row)
        into one word per candidate -> ONE all-to-all -> k-way merge of the own queries.  No host synchronisation.
        """
        dev = self._dev()
        lib = _native.load()
        nq = q_hat_local.shape[0]
        if self._replica is not None:   # replicated rows: the own queries against every row, no collective
            return self.search_device(q_hat_local, k, rows=self._replica, index_base=0)
        if q_counts is None:
            q_counts = all_gather_counts(nq, self.group, dev)
        elif len(q_counts) != self.world or q_counts[self.rank] != nq:
            raise ValueError(f"q_counts {q_counts} does not describe this rank's {nq} queries")
        q_all = all_gather_rows(q_hat_local, q_counts, self.group)
        dot, idx, _ = self.search_device(q_all, k)
        out_dot = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        out_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
        if self._global_count() >= 2 ** 32 - 1:
            raise ValueError("the packed candidate exchange addresses fewer than 2^32 - 1 dictionary rows")
        packed = torch.empty((q_all.shape[0], k), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            st = self._stream(dev)
            _native.check(lib.ebsd_topk_pack(dot.data_ptr(), idx.data_ptr(), packed.numel(), packed.data_ptr(), st),
                          "ebsd_topk_pack")
            gathered = exchange_packed(packed, q_counts, self.group)
            _native.check(
                lib.ebsd_topk_merge_packed(gathered.data_ptr(), self.world, nq, k, out_dot.data_ptr(),
                                           out_idx.data_ptr(), out_dist.data_ptr(), st),
                "ebsd_topk_merge_packed")
        return out_dot, out_idx, out_dist

    def _search_for_consensus(self, q_hat: torch.Tensor, k: int):
        return self.search_global(q_hat, k, getattr(self, "_next_q_counts", None))

    def query_similar(self, query_vector, n_results: int = 20, include_metadata: bool = True):
        """Global result of one query in Chroma's shape (collective: every rank calls it with its own query)."""
        query_vector = np.asarray(query_vector)
        if query_vector.ndim > 1:
            query_vector = query_vector.squeeze()
        if query_vector.shape[0] != self.dimension:
            raise ValueError(f"Expected query vector of dimension {self.dimension}, got {query_vector.shape[0]}")
        k = self._clamp_k(n_results)
        _, idx, dist_ = self.search_global(self._prepare_queries(query_vector), k)
        idx_h, dist_h = _to_host(idx[0], dist_[0])
        keep = idx_h >= 0
        idx_h, dist_h = idx_h[keep], dist_h[keep]
        out = {"ids": [[f"vec_{int(i)}" for i in idx_h]]}
        if include_metadata:
            orient = (self._global_eulers[torch.as_tensor(idx_h, device=self._dev())].cpu().numpy()
                      if len(idx_h) else np.zeros((0, 3)))
            out["distances"] = [[float(d) for d in dist_h]]
            out["metadatas"] = [[
                {"orientation_str": ",".join(map(str, o.tolist())), "phi1": float(o[0]), "Phi": float(o[1]),
                 "phi2": float(o[2])} for o in orient
            ]]
        return out

    def find_best_orientations_batch(self, query_vectors, batch_size: int = 32, top_n: int = 20,
                                     orientation_threshold: float = 1.0, min_required_matches: int = 18,
                                     max_iterations: int = 3, q_counts: list[int] | None = None) -> OrientationResultBatch:
        """As the base class, on the GLOBAL dictionary.  ``q_counts`` (queries per rank) saves the count all-gather."""
        self._next_q_counts = q_counts
        try:
            return super().find_best_orientations_batch(query_vectors, batch_size, top_n, orientation_threshold,
                                                        min_required_matches, max_iterations)
        finally:
            self._next_q_counts = None
