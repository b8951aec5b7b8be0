"""Importable name of the product package ``deep-image-matching_b200`` (a directory name Python cannot import):
this package's search path IS that directory, so every ``dim_b200.<module>`` is ``deep-image-matching_b200/<module>.py``."""
import os as _os

PACKAGE_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "deep-image-matching_b200")
__path__ = [PACKAGE_DIR]
__version__ = "0.2.0"
