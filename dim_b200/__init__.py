"""Importable alias of the product package, whose directory name
(``deep-image-matching_b200``) is not a valid Python identifier.

``import dim_b200.extractors.superpoint`` resolves to
``deep-image-matching_b200/extractors/superpoint.py``.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_real = _os.path.join(_os.path.dirname(_here), "deep-image-matching_b200")
__path__.append(_real)  # noqa: F821  (package attribute)
PACKAGE_DIR = _real

exec(open(_os.path.join(_real, "_init.py")).read())
