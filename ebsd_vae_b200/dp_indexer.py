"""``DiffractionPatternIndexer`` / ``IndexerConfig`` with the reference's API (latice/index/dp_indexer.py:27-297).

encode -> store -> query -> consensus, with every stage on the GPU:

* patterns are quantised/cropped on the host exactly like the reference transform (ebsd_vae_b200/transform.py),
  shipped as uint8 and encoded by the native encoder (only ``mu`` is computed; the reference runs the whole
  VAE and discards everything else, dp_indexer.py:136,183,284);
* latents never leave the device between encoder, dictionary and search when the caller stays inside
  ``build_dictionary`` / ``index_patterns_batch``.
"""
from __future__ import annotations

import logging
from pathlib import Path
from typing import Literal

import numpy as np
import torch
from numpy.typing import NDArray
from pydantic.dataclasses import dataclass

from .model import EncoderEngine
from .transform import load_patterns, parse_rotation_angles, transform_batch_device, transform_batch_u8
from .vector_db import LatentVectorDatabase, LatentVectorDatabaseConfig, OrientationResult

logger = logging.getLogger(__name__)


@dataclass
class IndexerConfig:
    """Configuration for the diffraction pattern indexer (fields/defaults of dp_indexer.py:26-48).

    ``pattern_path`` / ``angles_path`` are only needed by ``build_dictionary`` and default to None here (the
    reference declares them required but then constructs ``IndexerConfig()`` without them, dp_indexer.py:72).
    ``device`` defaults to "cuda": this implementation has no CPU path.
    """

    pattern_path: Path | None = None
    angles_path: Path | None = None
    batch_size: int = 64
    device: Literal["cuda", "cpu", "mps"] = "cuda"
    latent_dim: int = 16
    random_seed: int = 42
    image_size: tuple[int, int] = (128, 128)
    top_n: int = 20
    orientation_threshold: float = 3.0


class DiffractionPatternIndexer:
    """Indexes diffraction patterns using the VAE encoder and the GPU latent dictionary."""

    #: patterns encoded per native call inside build_dictionary / encode_patterns_batch (host staging granularity)
    ENCODE_CHUNK = 2960
    #: patterns per encoder pass (csrc/encoder.cu kChunkFused): host batches are copied in slices of one pass
    ENCODE_PASS = 1480

    @classmethod
    def _copy_slices(cls, b: int) -> list[tuple[int, int]]:
        """[a, e) slices of a host batch of ``b`` patterns for the pipelined host -> device copy.

        Nothing can be encoded before the first slice has landed, so the first slices are small (the remainder of
        ``b`` over whole encoder passes, halved when it is large) and every later slice is exactly one encoder pass:
        10 000 patterns -> 560 + 560 + 6 x 1480 (0.18 ms of exposed copy instead of the 0.8 ms of a 2500-pattern
        slice; the passes are as long as the device-resident path's).
        """
        p = cls.ENCODE_PASS
        if b <= p // 2:
            return [(0, b)]
        n = (b + p - 1) // p
        first = b - (n - 1) * p
        if n > 1 and first < p // 4:      # a very short remainder: spread it instead of a tiny extra pass
            step = (b + n - 1) // n
            step += step & 1
            sizes = [min(step, b - a) for a in range(0, b, step)]
        else:
            sizes = [first] + [p] * (n - 1)
        if sizes[0] > 640:                # halve the slice the encoder has to wait for
            h = (sizes[0] // 2 + 1) & ~1
            sizes[:1] = [h, sizes[0] - h]
        out, a = [], 0
        for sz in sizes:
            out.append((a, a + sz))
            a += sz
        return out

    def __init__(self, model, db: LatentVectorDatabase | None = None, config: IndexerConfig | None = None) -> None:
        self.config = config if config is not None else IndexerConfig()
        self.db = db if db is not None else LatentVectorDatabase(
            LatentVectorDatabaseConfig(dimension=self.config.latent_dim)
        )
        np.random.seed(self.config.random_seed)
        torch.manual_seed(self.config.random_seed)

        if self.config.device != "cuda":
            raise RuntimeError(
                f"device={self.config.device!r}: ebsd_vae_b200 runs on CUDA (sm_100a) only and has no CPU fallback"
            )
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available; ebsd_vae_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        logger.info(f"Using device: {self.device}")

        self.model = model
        self.model.eval()
        self.model.to(self.device)
        self._engine: EncoderEngine | None = None
        self._copy_stream: torch.cuda.Stream | None = None

    # ------------------------------------------------------------------ encoder plumbing
    @property
    def engine(self) -> EncoderEngine:
        if self._engine is None:
            if hasattr(self.model, "engine") and callable(self.model.engine):
                self._engine = self.model.engine()
            else:  # e.g. the reference's own nn.Module: take its weights
                self._engine = EncoderEngine(self.model.state_dict(), self.device)
        return self._engine

    def _encode_host_tensor(self, t: torch.Tensor, transform: bool = False) -> torch.Tensor:
        """Host tensor [B,H,W] -> mu [B,16] on the device.

        ``transform=False``: ``t`` is [B,128,128] uint8 (k/255 encoded) or float32 (used as is) -- tensors bypass the
        reference's transform (dp_indexer.py:128-131, 165-169).  ``transform=True``: ``t`` is what the reference
        feeds to ``create_default_transform`` (float32 / float64 / uint8 frames of any size); the 8-bit quantise and
        the centre crop run on the device (``ebsd_quantize_crop``), so raw frames cross PCIe once and no per-pattern
        host work remains.

        The host -> device copy is pipelined against the encoder: slices of at most one encoder pass (``_copy_slices``) are
        copied on a side stream into one device buffer while the compute stream works on the slices that have already
        landed (pinned sources copy asynchronously; pageable ones still work, just without the overlap).
        """
        b = t.shape[0]
        if b == 0:
            return torch.empty((0, self.config.latent_dim), dtype=torch.float32, device=self.device)
        t = t.contiguous()
        compute = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        dev = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        mu = torch.empty((b, self.config.latent_dim), dtype=torch.float32, device=self.device)
        self._copy_stream.wait_stream(compute)  # `dev` was allocated on the compute stream
        slices = self._copy_slices(b)
        events = []
        with torch.cuda.stream(self._copy_stream):
            for a, e in slices:
                dev[a:e].copy_(t[a:e], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                events.append(ev)
        for i, (a, e) in enumerate(slices):
            compute.wait_event(events[i])
            part = dev[a:e]
            if transform:
                part = transform_batch_device(part, tuple(self.config.image_size))
            mu[a:e] = self.engine.encode(part)
        return mu

    def _encode_u8_host(self, u8: np.ndarray) -> torch.Tensor:
        """uint8 [B,128,128] on the host -> mu [B,16] on the device."""
        return self._encode_host_tensor(torch.from_numpy(np.ascontiguousarray(u8)))

    def _encode_any(self, patterns) -> torch.Tensor:
        """Reference input rules (dp_indexer.py:124-131, 150-169): ndarrays go through the transform, tensors bypass it."""
        if isinstance(patterns, np.ndarray):
            if patterns.ndim not in (2, 3):
                raise AssertionError(f"Expected 4D tensor, got {patterns.ndim + 1}D")
            if patterns.ndim == 2:
                patterns = patterns[None]
            if patterns.dtype in (np.float32, np.float64, np.uint8):
                return self._encode_host_tensor(torch.from_numpy(np.ascontiguousarray(patterns)), transform=True)
            # anything else (float16, integers the reference rejects ...) goes through the host restatement, which
            # raises the reference's TypeError for unsupported dtypes
            return self._encode_u8_host(transform_batch_u8(patterns, tuple(self.config.image_size)))
        t = patterns
        if t.dim() == 2:
            t = t[None]
        elif t.dim() == 4:
            if t.shape[1] != 1:
                raise ValueError(f"expected a single channel, got {t.shape[1]}")
            t = t[:, 0]
        elif t.dim() != 3:
            raise AssertionError(f"Expected 4D tensor, got {t.dim()}D")
        if t.dtype != torch.uint8:
            t = t.to(torch.float32)
        if t.device.type == "cpu":
            return self._encode_host_tensor(t)
        t = t.to(self.device)
        outs = [self.engine.encode(t[a : a + self.ENCODE_CHUNK]) for a in range(0, t.shape[0], self.ENCODE_CHUNK)]
        return outs[0] if len(outs) == 1 else torch.cat(outs)

    # ------------------------------------------------------------------ reference API
    def build_dictionary(self) -> None:
        """Generate latent vectors from ``config.pattern_path`` / ``config.angles_path`` and add them to the db."""
        if self.config.pattern_path is None or self.config.angles_path is None:
            raise ValueError("IndexerConfig.pattern_path and angles_path are required by build_dictionary")
        logger.info(f"Generating latent vectors from patterns in {self.config.pattern_path}")
        latent_vectors, orientations = self._extract_latent_vectors_with_angles(
            self.config.pattern_path, self.config.angles_path
        )
        logger.info(f"Adding {len(latent_vectors)} vectors to database")
        self.db.add_vectors(latent_vectors, orientations)

    #: bytes of one pinned staging buffer of the streaming dictionary build (two are kept)
    STAGING_BYTES = 128 << 20

    def _extract_latent_vectors_with_angles(self, pattern_path, angles_path):
        """Encode every pattern of the .npy file; returns (latents CUDA [N,16] f32, orientations [N,3] f64).

        Replaces the reference's DataLoader loop (dp_indexer.py:254-297, data_module.py:122-133) by a stream:
        memory-mapped file -> two pinned staging buffers (filled by a few host threads) -> host-to-device copies on a
        side stream -> ``ebsd_quantize_crop`` + encoder on the compute stream, so that reading chunk i+1 overlaps
        encoding chunk i and frames cross PCIe once, in the file's own dtype.  The reference's dataset casts each
        pattern to float64 before the transform (data_module.py:132), so integer-typed files are scaled by 255 and
        wrap modulo 256 -- the kernel does that cast per pixel (``EBSD_SRC_VIA_F64``).
        """
        from concurrent.futures import ThreadPoolExecutor

        data = load_patterns(pattern_path)
        with ThreadPoolExecutor(max_workers=1) as pool:
            # the angle file (text, ~1 us per row in Python) is parsed while the GPU encodes
            angles_job = pool.submit(parse_rotation_angles, angles_path)
            latents = self._encode_frames_streaming(data)
            angles = angles_job.result()
        if len(angles) < len(data):
            raise ValueError(f"angle file has {len(angles)} rows for {len(data)} patterns")
        return latents, angles[: len(data)]

    _STREAM_DTYPES = (np.uint8, np.int16, np.uint16, np.int32, np.int64, np.float32, np.float64)

    def _encode_frames_streaming(self, data: np.ndarray) -> torch.Tensor:
        """[N,H,W] host array (typically a read-only memory map) -> mu [N,16] on the device, dataset semantics."""
        from concurrent.futures import ThreadPoolExecutor
        import os

        n = len(data)
        mu = torch.empty((n, self.config.latent_dim), dtype=torch.float32, device=self.device)
        if n == 0:
            return mu
        cast_on_host = data.dtype not in [np.dtype(t) for t in self._STREAM_DTYPES]   # e.g. float16, bool
        np_dtype = np.dtype(np.float64) if cast_on_host else data.dtype
        frame_bytes = int(np.prod(data.shape[1:])) * np_dtype.itemsize
        chunk = int(max(2, min(self.ENCODE_CHUNK, self.STAGING_BYTES // max(frame_bytes, 1), n)))
        chunk += chunk & 1
        t_dtype = torch.from_numpy(np.empty(0, dtype=np_dtype)).dtype
        shape = (chunk,) + tuple(data.shape[1:])
        key = (shape, t_dtype)
        if getattr(self, "_staging_key", None) != key:   # pinned allocations are slow (tens of ms): keep them
            self._staging = [torch.empty(shape, dtype=t_dtype, pin_memory=True) for _ in range(2)]
            self._staging_key = key
        staging = self._staging
        staging_np = [s.numpy() for s in staging]
        raw = [torch.empty(shape, dtype=t_dtype, device=self.device) for _ in range(2)]
        compute = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        self._copy_stream.wait_stream(compute)   # `raw` was allocated on the compute stream
        copied = [None, None]     # event: H2D of the chunk in staging[i] finished (the staging buffer may be refilled)
        consumed = [None, None]   # event: quantise/crop of raw[i] finished (the device buffer may be overwritten)
        workers = max(1, min(8, (os.cpu_count() or 2) - 1))
        with ThreadPoolExecutor(max_workers=workers) as pool:
            def fill(dst, src):
                rows = len(src)
                cuts = np.linspace(0, rows, min(workers, rows) + 1).astype(int)
                list(pool.map(lambda ab: np.copyto(dst[ab[0]:ab[1]], src[ab[0]:ab[1]], casting="unsafe"),
                              zip(cuts[:-1], cuts[1:])))

            for i, a in enumerate(range(0, n, chunk)):
                b = min(a + chunk, n)
                s = i & 1
                if copied[s] is not None:
                    copied[s].synchronize()
                fill(staging_np[s][: b - a], data[a:b])
                with torch.cuda.stream(self._copy_stream):
                    if consumed[s] is not None:
                        self._copy_stream.wait_event(consumed[s])
                    raw[s][: b - a].copy_(staging[s][: b - a], non_blocking=True)
                    copied[s] = torch.cuda.Event()
                    copied[s].record(self._copy_stream)
                compute.wait_event(copied[s])
                u8 = transform_batch_device(raw[s][: b - a], tuple(self.config.image_size), via_float64=True)
                consumed[s] = torch.cuda.Event()
                consumed[s].record(compute)
                mu[a:b] = self.engine.encode(u8)
        for s in range(2):   # the side stream's last writes of `raw` end before the buffers go back to the allocator
            raw[s].record_stream(self._copy_stream)
            if copied[s] is not None:
                copied[s].synchronize()   # the cached staging buffers may be refilled by the next call
        return mu

    def encode_pattern(self, pattern) -> NDArray[np.float32]:
        """Encode a single diffraction pattern to latent space -> (16,) float32."""
        return self._encode_any(pattern).cpu().numpy().squeeze()

    def encode_patterns_batch(self, patterns) -> NDArray[np.float32]:
        """Encode multiple diffraction patterns [B,H,W] -> (B,16) float32."""
        return self._encode_any(patterns).cpu().numpy()

    def index_pattern(self, pattern, top_n: int | None = None,
                      orientation_threshold: float | None = None) -> OrientationResult:
        """Index a diffraction pattern and return the best orientation (dp_indexer.py:188-214)."""
        top_n = top_n or self.config.top_n
        orientation_threshold = orientation_threshold or self.config.orientation_threshold
        if isinstance(self.db, LatentVectorDatabase):
            latent_vector = self._encode_any(pattern)[0]   # stays on the device: one synchronisation per pattern, not two
        else:                                              # a duck-typed dictionary (e.g. the reference's own classes)
            latent_vector = self.encode_pattern(pattern)
        return self.db.find_best_orientation(
            latent_vector, top_n=top_n, orientation_threshold=orientation_threshold
        )

    def index_patterns_batch(self, patterns, **kwargs):
        """Index multiple patterns; ``kwargs`` go to ``find_best_orientation`` (dp_indexer.py:216-232).

        Latents stay on the device between the encoder and the search.
        """
        latent_vectors = self._encode_any(patterns)
        return self.db.find_best_orientations_batch(latent_vectors, batch_size=self.config.batch_size, **kwargs)
