"""In-memory GPU replacement for the reference's ``LatentVectorDatabase``.

Same public surface as ``ChromaLatentVectorDatabase`` (latice/index/chroma_db.py:87-423): ``add_vectors``,
``create_from_files``, ``query_similar``, ``find_best_orientation``, ``find_best_orientations_batch``,
``get_count``, ``delete_collection``, ``_validate_vectors``; ``LatentVectorDatabaseConfig`` and
``OrientationResult`` keep their fields and defaults (chroma_db.py:25-84).  Instead of Chroma/HNSW (approximate)
the dictionary lives in HBM as normalised fp32 rows and every query is an exact brute-force search
(ebsd_topk), followed by the quaternion consensus kernel (ebsd_consensus).

Consensus semantics follow the Chroma class by default (``mode="chroma"``: the threshold is compared with
radians, ``best_orientation`` stays the nearest candidate); ``mode="faiss"`` gives the FAISS twin's behaviour
(latice/index/faiss_db.py:258-372: threshold in degrees, ``best_orientation`` = mean on success).
"""
from __future__ import annotations

import logging
from collections.abc import Sequence
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np
import torch
from numpy.typing import NDArray

from . import _native

logger = logging.getLogger(__name__)


@dataclass
class LatentVectorDatabaseConfig:
    """Configuration (fields and defaults of latice/index/chroma_db.py:25-38, plus GPU-side options).

    Attributes:
        collection_name: Name of the collection (kept for API compatibility)
        persist_directory: Kept for API compatibility; the dictionary is held in GPU memory
        dimension: Dimension of the latent vectors
        device: CUDA device holding the dictionary
        mode: "chroma" (reference default semantics) or "faiss" (degrees threshold, see module docstring)
    """

    collection_name: str = "latent_vectors"
    persist_directory: str | None = ".chroma_db"
    dimension: int = 16
    device: str = "cuda"
    mode: str = "chroma"


@dataclass
class OrientationResult:
    """Results from orientation matching query (same fields as latice/index/chroma_db.py:41-84)."""

    query_vector: NDArray[np.float64]
    best_orientation: NDArray[np.float64]
    candidate_orientations: NDArray[np.float64]
    distances: NDArray[np.float64]
    mean_orientation: NDArray[np.float64] | None = None
    success: bool = True
    similar_indices: NDArray[np.int64] = None

    def get_top_n_orientations(self, n: int = 5) -> NDArray[np.float64]:
        """Return the top N orientations sorted by similarity (ascending distance)."""
        if self.distances is None or len(self.distances) == 0:
            return self.candidate_orientations[: min(n, len(self.candidate_orientations))]
        order = np.argsort(self.distances)
        return self.candidate_orientations[order[: min(n, len(order))]]


class OrientationResultBatch(Sequence):
    """Struct-of-arrays result of a batched query; behaves like ``list[OrientationResult]``.

    Building tens of thousands of dataclass instances would cost more than the GPU work, so items are
    materialised on access.  The arrays are also exposed directly:
    ``success`` [Q] bool, ``mean_orientations`` [Q,3] (NaN where not successful), ``best_orientations`` [Q,3],
    ``candidate_orientations`` [Q,k,3], ``distances`` [Q,k], ``indices`` [Q,k] (global dictionary rows),
    ``similar_masks`` [Q] uint64, ``query_vectors`` [Q,16].
    """

    def __init__(self, query_vectors, indices, distances, candidate_orientations, success, mean_orientations,
                 similar_masks, faiss_mode: bool):
        self.query_vectors = query_vectors
        self.indices = indices
        self.distances = distances
        self.candidate_orientations = candidate_orientations
        self.success = success
        self.mean_orientations = mean_orientations
        self.similar_masks = similar_masks
        self._faiss = faiss_mode
        valid = indices >= 0
        self._n_valid = valid.sum(axis=1)
        if faiss_mode:
            self.best_orientations = np.where(success[:, None], mean_orientations, candidate_orientations[:, 0, :])
        else:
            self.best_orientations = candidate_orientations[:, 0, :]

    def __len__(self) -> int:
        return len(self.success)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        kv = int(self._n_valid[i])
        ok = bool(self.success[i])
        mask = int(self.similar_masks[i])
        similar = np.array([b for b in range(kv) if (mask >> b) & 1], dtype=np.int64)
        return OrientationResult(
            query_vector=self.query_vectors[i],
            best_orientation=self.best_orientations[i].copy(),
            mean_orientation=self.mean_orientations[i].copy() if ok else None,
            candidate_orientations=self.candidate_orientations[i, :kv].copy(),
            distances=self.distances[i, :kv].astype(np.float64),
            success=ok,
            similar_indices=similar,
        )


def _to_host(*tensors):
    """Device tensors -> numpy arrays with ONE stream synchronisation (pinned staging from torch's caching host
    allocator, non-blocking copies).  Host tensors pass through."""
    staged = []
    dev = None
    for t in tensors:
        if t.device.type == "cuda":
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            dev = t.device
            staged.append(h)
        else:
            staged.append(t)
    if dev is not None:
        torch.cuda.current_stream(dev).synchronize()
    return tuple(h.numpy() for h in staged)  # views: the (pinned) staging lives as long as the arrays do


class LatentVectorDatabase:
    """Exact-search latent dictionary resident in GPU memory."""

    def __init__(self, config: LatentVectorDatabaseConfig | None = None) -> None:
        self.config = config if config is not None else LatentVectorDatabaseConfig()
        self.collection_name = self.config.collection_name
        self.dimension = self.config.dimension
        self.persist_directory = self.config.persist_directory
        if self.config.mode not in ("chroma", "faiss"):
            raise ValueError(f"mode must be 'chroma' or 'faiss', got {self.config.mode!r}")
        self._device: torch.device | None = None
        self._count = 0
        self._capacity = 0
        self._latents: torch.Tensor | None = None    # [cap,16] f32, rows normalised
        self._eulers: torch.Tensor | None = None     # [cap,3] f64 (phi1, Phi, phi2) degrees, as given
        self._quats: torch.Tensor | None = None      # [cap,4] f64 (x,y,z,w)
        self._topk_ws: torch.Tensor | None = None
        self.index_base = 0                          # global row index of local row 0 (row-sharded use)
        logger.info("Created in-memory GPU latent dictionary '%s'", self.collection_name)
        # both reference classes reopen a persisted store on construction (chroma_db.py:113-131 PersistentClient +
        # get_collection; faiss_db.py:129-130 ``if self.npz_path.exists(): self.load()``)
        if self.persist_directory is not None and self._reopen_on_init() and self.npz_path.exists():
            self.load()

    def _reopen_on_init(self) -> bool:
        return True

    # ------------------------------------------------------------------ helpers
    def _dev(self) -> torch.device:
        if self._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("LatentVectorDatabase needs a CUDA device (sm_100a); there is no CPU search path")
            dev = torch.device(self.config.device)
            if dev.type != "cuda":
                raise RuntimeError(f"LatentVectorDatabase needs a CUDA device, got {self.config.device!r}")
            if dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            self._device = dev
        return self._device

    @staticmethod
    def _stream(dev) -> int:
        return torch.cuda.current_stream(dev).cuda_stream

    def _reserve(self, extra: int) -> None:
        need = self._count + extra
        if need <= self._capacity:
            return
        dev = self._dev()
        cap = max(need, int(self._capacity * 1.5), 1024)
        lat = torch.empty((cap, self.dimension), dtype=torch.float32, device=dev)
        eul = torch.empty((cap, 3), dtype=torch.float64, device=dev)
        qua = torch.empty((cap, 4), dtype=torch.float64, device=dev)
        if self._count:
            lat[: self._count] = self._latents[: self._count]
            eul[: self._count] = self._eulers[: self._count]
            qua[: self._count] = self._quats[: self._count]
        self._latents, self._eulers, self._quats, self._capacity = lat, eul, qua, cap

    def _validate_vectors(self, latent_vectors, orientations) -> None:
        if len(latent_vectors) != len(orientations):
            raise ValueError("Number of latent vectors and orientations must match")
        if latent_vectors.shape[1] != self.dimension:
            raise ValueError(
                f"Expected latent vectors of dimension {self.dimension}, got {latent_vectors.shape[1]}"
            )

    # ------------------------------------------------------------------ population
    def add_vectors(self, latent_vectors, orientations, batch_size: int = 1000) -> None:
        """Append latent vectors [n,16] and ZXZ Euler orientations [n,3] (degrees).

        Accepts numpy arrays (as the reference) or torch tensors (CUDA tensors are taken without a host
        round trip).  ``batch_size`` is accepted for API compatibility; the upload is one copy.
        """
        self._validate_vectors(latent_vectors, orientations)
        if self.dimension != 16:
            raise ValueError("the B200 search kernel is specialised for 16-D latents")
        n = len(latent_vectors)
        if n == 0:
            return
        if orientations.shape[1] != 3:
            raise ValueError(f"Expected orientations of shape (n, 3), got {tuple(orientations.shape)}")
        dev = self._dev()
        lib = _native.load()
        self._reserve(n)
        lat = torch.as_tensor(latent_vectors).to(device=dev, dtype=torch.float32)
        eul = torch.as_tensor(orientations).to(device=dev, dtype=torch.float64)
        a, b = self._count, self._count + n
        self._latents[a:b] = lat
        self._eulers[a:b] = eul
        with torch.cuda.device(dev):
            st = self._stream(dev)
            _native.check(lib.ebsd_normalize_rows(self._latents[a:b].data_ptr(), n, self.dimension, st),
                          "ebsd_normalize_rows")
            _native.check(lib.ebsd_euler_to_quat(self._eulers[a:b].data_ptr(), n, self._quats[a:b].data_ptr(), st),
                          "ebsd_euler_to_quat")
        self._count = b
        logger.info("Successfully added %d vectors to the dictionary (total %d)", n, self._count)

    def create_from_files(self, latent_file_path: Path, angles_file_path: Path, batch_size: int = 1000) -> None:
        latent_vectors = np.load(Path(latent_file_path))
        orientations = np.load(Path(angles_file_path))
        self.add_vectors(latent_vectors, orientations, batch_size)

    def get_count(self) -> int:
        return self._count

    def delete_collection(self) -> None:
        self._count = 0
        self._capacity = 0
        self._latents = self._eulers = self._quats = None
        logger.info("Deleted collection '%s'", self.collection_name)

    # ------------------------------------------------------------------ persistence
    @property
    def npz_path(self) -> Path:
        """``<persist_directory>/<collection_name>.npz`` -- the role of Chroma's ``persist_directory``
        (chroma_db.py:113-117) and of the FAISS variant's single ``.npz`` file (faiss_db.py:125, 440-476)."""
        if self.persist_directory is None:
            raise ValueError("persist_directory is None: this dictionary is in-memory only (pass a path to save/load)")
        return Path(self.persist_directory) / f"{self.collection_name}.npz"

    def save(self, path: str | Path | None = None) -> Path:
        """Write the dictionary to one ``.npz``: ``latents`` (fp32 [N,16], rows normalised exactly as searched),
        ``orientations`` (float64 [N,3] degrees -- the key the reference's FAISS file uses, faiss_db.py:448-455),
        ``dimension``.  The FAISS file's ``faiss_index`` blob is a serialised third-party object and is not produced."""
        path = Path(path) if path is not None else self.npz_path
        path = path.with_suffix(".npz")
        path.parent.mkdir(parents=True, exist_ok=True)
        n = self._count
        lat = self._latents[:n].cpu().numpy() if n else np.zeros((0, self.dimension), np.float32)
        eul = self._eulers[:n].cpu().numpy() if n else np.zeros((0, 3), np.float64)
        np.savez_compressed(str(path), latents=lat, orientations=eul, dimension=np.int64(self.dimension))
        logger.info("Saved %d vectors to %s", n, path)
        return path

    def load(self, path: str | Path | None = None) -> None:
        """Replace the in-memory dictionary with the contents of a file written by :meth:`save` (raises
        ``FileNotFoundError("NPZ file missing.")`` like faiss_db.py:463-465).  Rows are stored normalised;
        normalising them again on insertion is the identity up to one fp32 rounding, so searches return the same
        rows (the stored, not the re-normalised, values are used: they are copied verbatim)."""
        path = Path(path) if path is not None else self.npz_path
        path = path.with_suffix(".npz")
        if not path.exists():
            logger.error("Cannot load. NPZ file %s not found.", path)
            raise FileNotFoundError("NPZ file missing.")
        data = np.load(str(path))
        lat, eul = data["latents"], data["orientations"]
        if "dimension" in data.files and int(data["dimension"]) != self.dimension:
            raise ValueError(f"Expected latent vectors of dimension {self.dimension}, got {int(data['dimension'])}")
        self._validate_vectors(lat, eul)
        self.delete_collection()
        n = len(lat)
        if n == 0:
            return
        dev = self._dev()
        self._reserve(n)
        self._latents[:n] = torch.from_numpy(np.ascontiguousarray(lat, dtype=np.float32)).to(dev)   # verbatim
        self._eulers[:n] = torch.from_numpy(np.ascontiguousarray(eul, dtype=np.float64)).to(dev)
        with torch.cuda.device(dev):
            _native.check(_native.load().ebsd_euler_to_quat(self._eulers[:n].data_ptr(), n, self._quats[:n].data_ptr(),
                                                            self._stream(dev)), "ebsd_euler_to_quat")
        self._count = n
        logger.info("Loaded %d vectors from %s", n, path)

    def delete_persistence(self) -> None:
        """Delete the persisted file and reset the in-memory dictionary (faiss_db.py:478-496)."""
        path = self.npz_path.with_suffix(".npz")
        if path.exists():
            path.unlink()
            logger.info("Deleted index file: %s", path)
            self.delete_collection()

    # ------------------------------------------------------------------ device-level search
    def _prepare_queries(self, query_vectors) -> torch.Tensor:
        dev = self._dev()
        q = torch.as_tensor(query_vectors)
        if q.dim() == 1:
            q = q[None]
        if q.shape[1] != self.dimension:
            raise ValueError(f"Expected query vector of dimension {self.dimension}, got {q.shape[1]}")
        q = q.to(device=dev, dtype=torch.float32, copy=True).contiguous()
        if q.shape[0]:
            with torch.cuda.device(dev):
                _native.check(_native.load().ebsd_normalize_rows(q.data_ptr(), q.shape[0], self.dimension,
                                                                 self._stream(dev)), "ebsd_normalize_rows")
        return q

    def search_device(self, q_hat: torch.Tensor, k: int, rows: torch.Tensor | None = None, index_base: int | None = None):
        """Exact top-k of normalised queries [Q,16] (CUDA) against the local rows (or against ``rows``, normalised
        [N,16] on the same device, whose row 0 is global row ``index_base``).

        Returns (dot [Q,k] f32, idx [Q,k] i64 global rows or -1, dist [Q,k] f32 = 1 - dot), all on the device.
        """
        dev = self._dev()
        lib = _native.load()
        nq = q_hat.shape[0]
        if rows is None:
            rows, count, base = self._latents, self._count, self.index_base
        else:
            count, base = int(rows.shape[0]), int(index_base or 0)
        dot = torch.empty((nq, k), dtype=torch.float32, device=dev)
        idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
        if nq == 0:
            return dot, idx, dist
        with torch.cuda.device(dev):   # the plan behind the workspace size depends on the CURRENT device's SM count
            need = int(lib.ebsd_topk_workspace_bytes(count, nq, k))
            if need and (self._topk_ws is None or self._topk_ws.numel() < need):
                self._topk_ws = torch.empty(need, dtype=torch.uint8, device=dev)
            _native.check(
                lib.ebsd_topk(
                    rows.data_ptr() if count else None, count, base,
                    q_hat.data_ptr(), nq, k, dot.data_ptr(), idx.data_ptr(), dist.data_ptr(),
                    self._topk_ws.data_ptr() if need else None, need, self._stream(dev)),
                "ebsd_topk",
            )
        return dot, idx, dist

    def _orientation_tables(self):
        """(euler [N,3] f64, quat [N,4] f64, global row of table row 0) covering every row a search can return."""
        if self._count == 0:
            return None, None, self.index_base
        return self._eulers[: self._count], self._quats[: self._count], self.index_base

    def consensus_device(self, idx: torch.Tensor, orientation_threshold: float, min_required_matches: int,
                         max_iterations: int):
        """Consensus of candidate lists idx [Q,k] (global rows, -1 = empty): one kernel launch, no eager glue.

        Returns (mean_quat [Q,4], mean_euler [Q,3], success [Q] u8, similar_mask [Q] i64, ref_iter [Q] i32,
        candidate Euler triplets [Q,k,3] -- NaN in empty slots), all on the device."""
        dev = self._dev()
        lib = _native.load()
        nq, k = idx.shape
        eulers, quats, base = self._orientation_tables()
        mean_q = torch.empty((nq, 4), dtype=torch.float64, device=dev)
        mean_e = torch.empty((nq, 3), dtype=torch.float64, device=dev)
        success = torch.empty((nq,), dtype=torch.uint8, device=dev)
        mask = torch.empty((nq,), dtype=torch.int64, device=dev)
        ref_it = torch.empty((nq,), dtype=torch.int32, device=dev)
        cand = torch.empty((nq, k, 3), dtype=torch.float64, device=dev)
        faiss = self.config.mode == "faiss"
        if nq:
            n_rows = 0 if quats is None else quats.shape[0]
            with torch.cuda.device(dev):
                _native.check(
                    lib.ebsd_consensus(
                        quats.data_ptr() if n_rows else None, eulers.data_ptr() if n_rows else None, n_rows, int(base),
                        idx.data_ptr(), nq, k, float(orientation_threshold),
                        _native.ANGLE_DEGREES if faiss else _native.ANGLE_RADIANS, int(min_required_matches),
                        int(max_iterations), int(faiss), mean_q.data_ptr(), mean_e.data_ptr(), success.data_ptr(),
                        mask.data_ptr(), ref_it.data_ptr(), cand.data_ptr(), self._stream(dev)),
                    "ebsd_consensus",
                )
        return mean_q, mean_e, success, mask, ref_it, cand

    # ------------------------------------------------------------------ reference API
    def query_similar(self, query_vector, n_results: int = 20, include_metadata: bool = True) -> dict[str, Any]:
        """Chroma-shaped result of one query: ``{"ids": [[..]], "distances": [[..]], "metadatas": [[..]]}``."""
        query_vector = np.asarray(query_vector)
        if query_vector.ndim > 1:
            query_vector = query_vector.squeeze()
        if query_vector.shape[0] != self.dimension:
            raise ValueError(f"Expected query vector of dimension {self.dimension}, got {query_vector.shape[0]}")
        k = self._clamp_k(n_results)
        q = self._prepare_queries(query_vector)
        _, idx, dist = self.search_device(q, k)
        idx_h = idx[0].cpu().numpy()
        dist_h = dist[0].cpu().numpy()
        keep = idx_h >= 0
        idx_h, dist_h = idx_h[keep], dist_h[keep]
        out: dict[str, Any] = {"ids": [[f"vec_{int(i)}" for i in idx_h]]}
        if include_metadata:
            eulers, _, base = self._orientation_tables()
            orient = eulers[torch.as_tensor(idx_h - base, device=eulers.device)].cpu().numpy() if len(idx_h) else \
                np.zeros((0, 3))
            out["distances"] = [[float(d) for d in dist_h]]   # cosine distance 1 - cos, ascending (chroma_db.py:127-130)
            out["metadatas"] = [[
                {"orientation_str": ",".join(map(str, o.tolist())), "phi1": float(o[0]), "Phi": float(o[1]),
                 "phi2": float(o[2])} for o in orient
            ]]
        return out

    def _clamp_k(self, top_n: int) -> int:
        if top_n < 1:
            raise ValueError(f"top_n must be >= 1, got {top_n}")
        if top_n > _native.MAX_TOPK:
            raise ValueError(f"top_n up to {_native.MAX_TOPK} is supported by the B200 search kernel, got {top_n}")
        return int(top_n)

    def find_best_orientations_batch(self, query_vectors, batch_size: int = 32, top_n: int = 20,
                                     orientation_threshold: float = 1.0, min_required_matches: int = 18,
                                     max_iterations: int = 3) -> OrientationResultBatch:
        """Batched ``find_best_orientation``: one exact search + one consensus launch for all queries.

        ``batch_size`` is accepted for API compatibility (chroma_db.py:377-410 loops serially in Python).
        """
        k = self._clamp_k(top_n)
        q_in = torch.as_tensor(query_vectors)
        if q_in.dim() == 1:
            q_in = q_in[None]
        q = self._prepare_queries(q_in)
        dot, idx, dist = self._search_for_consensus(q, k)
        _, mean_e, success, mask, _, cand = self.consensus_device(idx, orientation_threshold, min_required_matches,
                                                                  max_iterations)
        return self._finish_batch(q_in, dot, idx, dist, cand, success, mean_e, mask, max_iterations)

    def _search_for_consensus(self, q_hat: torch.Tensor, k: int):
        return self.search_device(q_hat, k)

    def _finish_batch(self, q_in, dot, idx, dist, cand, success, mean_e, mask, max_iterations) -> OrientationResultBatch:
        faiss = self.config.mode == "faiss"
        # one synchronisation for all result arrays: asynchronous copies into pinned staging, then a single wait.
        # FAISS semantics carry the inner products as ``distances`` (faiss_db.py:216-256, 281-300), Chroma the cosine
        # distance 1 - cos.
        qv, idx_h, dist_h, cand_h, succ_h, mean_h, mask_h = _to_host(q_in.detach(), idx, dot if faiss else dist, cand,
                                                                     success, mean_e, mask)
        succ_h = succ_h.astype(bool)
        if not faiss and len(succ_h):
            # The reference indexes orientations[iteration] unguarded (chroma_db.py:302-303) and leaves the loop on
            # success (:324-326): a query raises IndexError only when every reference orientation it has failed and
            # the next iteration would index past its candidates.  (FAISS clamps the loop instead, faiss_db.py:302.)
            n_valid = (idx_h >= 0).sum(axis=1)
            bad = np.flatnonzero(~succ_h & (n_valid < max_iterations))
            if len(bad):
                n = int(n_valid[bad[0]])
                raise IndexError(f"index {n} is out of bounds for axis 0 with size {n}")
        return OrientationResultBatch(
            query_vectors=qv,
            indices=idx_h,
            distances=dist_h,
            candidate_orientations=cand_h,
            success=succ_h,
            mean_orientations=mean_h,
            similar_masks=mask_h.astype(np.uint64),
            faiss_mode=faiss,
        )

    def _global_count(self) -> int:
        return self._count

    def find_best_orientation(self, query_vector, top_n: int = 20, orientation_threshold: float = 1.0,
                              min_required_matches: int = 18, max_iterations: int = 3) -> OrientationResult:
        """Find the best matching orientation for one query vector (chroma_db.py:261-342).

        A CUDA tensor is taken as it is (``index_pattern`` hands over the encoder's output without a round trip through
        the host); the result then carries its host copy as ``query_vector``."""
        on_device = isinstance(query_vector, torch.Tensor) and query_vector.is_cuda
        qv = query_vector.detach().reshape(-1) if on_device else np.asarray(query_vector)
        if qv.ndim > 1:
            qv = qv.squeeze()
        if qv.shape[0] != self.dimension:
            raise ValueError(f"Expected query vector of dimension {self.dimension}, got {qv.shape[0]}")
        batch = self.find_best_orientations_batch(qv[None], top_n=top_n, orientation_threshold=orientation_threshold,
                                                  min_required_matches=min_required_matches,
                                                  max_iterations=max_iterations)
        res = batch[0]
        if not on_device:
            res.query_vector = query_vector
        if not res.success:
            logger.warning("Failed to find best orientation after %d iterations", max_iterations)
        return res


# Name used by the current reference code (latice/index/chroma_db.py:87); README/notebooks use LatentVectorDatabase.
ChromaLatentVectorDatabase = LatentVectorDatabase


@dataclass
class FaissLatentVectorDatabaseConfig:
    """Fields and defaults of latice/index/faiss_db.py:34-46 (plus the CUDA device holding the dictionary)."""

    npz_path: str = "faiss_index.npz"
    dimension: int = 16
    device: str = "cuda"

    # what the shared implementation reads from a configuration
    mode = "faiss"

    @property
    def collection_name(self) -> str:
        return Path(self.npz_path).with_suffix("").name

    @property
    def persist_directory(self) -> str:
        return str(Path(self.npz_path).parent)


class FaissLatentVectorDatabase(LatentVectorDatabase):
    """Drop-in for the reference's FAISS twin (latice/index/faiss_db.py:92-496): the same exact search and consensus
    kernels with that class's conventions -- ``query_similar`` returns ``(similarities, indices)`` (inner products,
    descending) instead of Chroma's dictionary, ``n_results`` is clamped to the row count, an empty index gives empty
    arrays / a NaN result, thresholds are degrees, ``best_orientation`` is the mean on success, iterations are clamped
    to the candidates, and the dictionary persists to ONE ``.npz`` (``npz_path``), reopened on construction."""

    def __init__(self, config: FaissLatentVectorDatabaseConfig | None = None) -> None:
        super().__init__(config if config is not None else FaissLatentVectorDatabaseConfig())

    def query_similar(self, query_vector, n_results: int = 20):
        """(similarities [k] f32 descending, indices [k] i64) of one query (faiss_db.py:216-256)."""
        if self.get_count() == 0:
            logger.warning("Querying an empty index.")
            return np.array([]), np.array([])
        if self.get_count() < n_results:
            logger.warning("Requested %d results, but index only contains %d vectors. Returning all.", n_results,
                           self.get_count())
            n_results = self.get_count()
        query_vector = np.asarray(query_vector)
        if query_vector.ndim == 1:
            query_vector = query_vector.reshape(1, -1)
        if query_vector.shape[1] != self.dimension:
            raise ValueError(f"Expected query vector of dimension {self.dimension}, got {query_vector.shape[1]}")
        k = self._clamp_k(n_results)
        dot, idx, _ = self.search_device(self._prepare_queries(query_vector[:1]), k)
        dot_h, idx_h = _to_host(dot[0], idx[0])
        return dot_h.copy(), idx_h.copy()

    def find_best_orientation(self, query_vector, top_n: int = 20, orientation_threshold: float = 1.0,
                              min_required_matches: int = 18, max_iterations: int = 3) -> OrientationResult:
        if self.get_count() == 0:   # faiss_db.py:280-291
            logger.warning("No similar vectors found for query.")
            return OrientationResult(query_vector=np.asarray(query_vector).squeeze(),
                                     best_orientation=np.array([np.nan, np.nan, np.nan]),
                                     candidate_orientations=np.array([]), distances=np.array([]),
                                     mean_orientation=None, success=False, similar_indices=None)
        res = super().find_best_orientation(query_vector, top_n=min(top_n, self.get_count()),
                                            orientation_threshold=orientation_threshold,
                                            min_required_matches=min_required_matches, max_iterations=max_iterations)
        if not (isinstance(query_vector, torch.Tensor) and query_vector.is_cuda):
            res.query_vector = np.asarray(query_vector).squeeze().astype(np.float64)   # faiss_db.py:358-360
        return res
