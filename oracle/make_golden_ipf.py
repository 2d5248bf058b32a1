"""Generate tests/golden/ipf.npz by running the UNMODIFIED reference colour-key functions (build container only).

    python -m oracle.make_golden_ipf

``get_color_key`` lives in latice/utils/utils.py (206-240), which imports altair / matplotlib / pytorch_lightning for
its plotting helpers; those are replaced by empty stand-ins, the file itself is executed as it lies.  It calls
``ColorKeyGenerator.generate_ipf_color`` (latice/utils/colorkey.py:64-130) for every orientation.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

from . import refload

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def ipf_cases() -> np.ndarray:
    rng = np.random.default_rng(314)
    rand = np.stack([rng.uniform(0, 360, 400), rng.uniform(0, 180, 400), rng.uniform(0, 360, 400)], axis=1)
    sample = np.stack([np.zeros(40), np.arange(1.5, 601.5, 15.0), np.zeros(40)], axis=1)  # anglefile-like (0, i, 0)
    return np.concatenate([rand, sample])


def load_reference_utils():
    if not refload.available():
        raise RuntimeError("reference tree not found")
    class _Anything(types.ModuleType):
        """Stand-in module: any attribute (only used in annotations of the plotting helpers) resolves to ``object``."""

        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return object

    for name in ("altair", "matplotlib", "matplotlib.pyplot", "matplotlib.figure", "pytorch_lightning"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    pl = sys.modules["pytorch_lightning"]
    if not isinstance(getattr(pl, "loggers", None), types.SimpleNamespace):
        pl.loggers = types.SimpleNamespace(TensorBoardLogger=object, WandbLogger=object)
    if refload.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, refload.REFERENCE_ROOT)
    spec = importlib.util.spec_from_file_location(
        "_latice_utils_unmodified", os.path.join(refload.REFERENCE_ROOT, "latice", "utils", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main() -> None:
    utils = load_reference_utils()
    cases = ipf_cases()
    out = {mode: np.asarray(utils.get_color_key(cases, mode=mode)) for mode in ("ipf_x", "ipf_y", "ipf_z")}
    np.savez_compressed(os.path.join(GOLDEN, "ipf.npz"), eulers=cases, **out)
    print("ipf golden:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
