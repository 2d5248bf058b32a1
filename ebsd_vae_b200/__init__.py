"""ebsd_vae_b200 -- B200-native dictionary-indexing hot path behind the ``latice`` API.

Drop-in names (reference: poyentung/ebsd-vae, package ``latice``):

* ``DiffractionPatternIndexer``, ``IndexerConfig``           (latice/index/dp_indexer.py)
* ``LatentVectorDatabase`` (= ``ChromaLatentVectorDatabase``), ``LatentVectorDatabaseConfig``,
  ``OrientationResult``                                      (latice/index/chroma_db.py)
* ``FaissLatentVectorDatabase``, ``FaissLatentVectorDatabaseConfig``  (latice/index/faiss_db.py)
* ``VariationalAutoEncoderRawData``                          (latice/model.py)
* ``get_color_key``                                          (latice/utils/utils.py:206-240, IPF colours)

All compute runs in libebsd_b200.so (hand-written sm_100a CUDA behind a C ABI, include/ebsd_b200.h).
There is no CPU or eager-PyTorch fallback.
"""
from .colorkey import get_color_key, ipf_colors_device
from .dp_indexer import DiffractionPatternIndexer, IndexerConfig
from .model import EncoderEngine, VariationalAutoEncoderRawData, load_vae_weights
from .vector_db import (
    ChromaLatentVectorDatabase,
    FaissLatentVectorDatabase,
    FaissLatentVectorDatabaseConfig,
    LatentVectorDatabase,
    LatentVectorDatabaseConfig,
    OrientationResult,
    OrientationResultBatch,
)

__all__ = [
    "DiffractionPatternIndexer",
    "IndexerConfig",
    "EncoderEngine",
    "VariationalAutoEncoderRawData",
    "load_vae_weights",
    "ChromaLatentVectorDatabase",
    "FaissLatentVectorDatabase",
    "FaissLatentVectorDatabaseConfig",
    "LatentVectorDatabase",
    "LatentVectorDatabaseConfig",
    "OrientationResult",
    "OrientationResultBatch",
    "get_color_key",
    "ipf_colors_device",
]
