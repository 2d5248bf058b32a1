"""Import the UNMODIFIED reference package from ``/root/reference`` (build container only).

Test infrastructure.  The reference (``latice``) imports chromadb, pytorch_lightning
and a module that is missing from its own tree (``latice/index/latent_vector_db_base.py``,
imported at latice/index/chroma_db.py:18).  None of those are on the hot path's
arithmetic, so they are replaced by empty stand-ins in ``sys.modules`` before the
import; the reference sources themselves are used as they lie.

The GPU box has no ``/root/reference``: nothing that runs there may call this.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EBSD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "latice"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_loaded = False


def load():
    """Return the namespace ``(model, data_module, chroma_db, dp_indexer, constants)`` of reference modules."""
    global _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not _loaded:
        class _InvalidCollection(Exception):
            pass

        if "chromadb" not in sys.modules:
            errs = _stub("chromadb.errors", InvalidCollectionException=_InvalidCollection)
            _stub("chromadb", PersistentClient=lambda *a, **k: None, Client=lambda *a, **k: None, errors=errs)
        if "pytorch_lightning" not in sys.modules:
            _stub(
                "pytorch_lightning",
                LightningDataModule=type("LightningDataModule", (), {"__init__": lambda self, *a, **k: None}),
            )
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import latice.index  # noqa: F401  (package __init__ has no imports)
        import latice.utils.constants as constants

        _stub("latice.index.latent_vector_db_base", LatentVectorDatabaseBase=type("LatentVectorDatabaseBase", (), {}))
        # latice/utils/utils.py drags in altair/matplotlib; only QUAT_SYM is used (chroma_db.py:19) and it is
        # the same table as latice/utils/constants.py:13-39.
        _stub("latice.utils.utils", QUAT_SYM=constants.QUAT_SYM)
        _loaded = True

    import latice.model as model
    import latice.data_module as data_module
    import latice.index.chroma_db as chroma_db
    import latice.index.dp_indexer as dp_indexer
    import latice.utils.constants as constants

    return types.SimpleNamespace(
        model=model, data_module=data_module, chroma_db=chroma_db, dp_indexer=dp_indexer, constants=constants
    )
