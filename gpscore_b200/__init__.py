"""Importable alias for the package directory
`scoring-rules-for-gaussian-process-regression-a-new-approach-to-inference_b200/`
(its name is not a valid Python identifier).  All code lives there; this file
only points the import system at it: `import gpscore_b200.api` loads
`<that directory>/api.py`.
"""
import os as _os

_PKG_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "scoring-rules-for-gaussian-process-regression-a-new-approach-to-inference_b200",
)
__path__.append(_PKG_DIR)
PKG_DIR = _PKG_DIR
