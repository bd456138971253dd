"""Importable alias for the product package.

The product lives in the directory the build contract names,
``blurr-a-boosted-low-resource-inference-for-vision-language-action-model_b200/``,
whose name is not a valid Python identifier.  This shim makes it importable as
``blurr_b200`` by pointing the package search path at that directory, so
``import blurr_b200.pizero`` loads ``<that dir>/pizero.py``.
"""

from __future__ import annotations

import os as _os

PACKAGE_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "blurr-a-boosted-low-resource-inference-for-vision-language-action-model_b200",
)
if not _os.path.isdir(PACKAGE_DIR):  # pragma: no cover - broken checkout
    raise ImportError(f"product package directory missing: {PACKAGE_DIR}")

__path__.insert(0, PACKAGE_DIR)  # type: ignore[name-defined]

with open(_os.path.join(PACKAGE_DIR, "__init__.py"), "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), _os.path.join(PACKAGE_DIR, "__init__.py"), "exec"))
