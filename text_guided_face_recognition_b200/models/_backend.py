"""Resolve the kernel bindings whether this package is imported as
`text_guided_face_recognition_b200.models` or as a top-level `models` (drop-in mode)."""
import os
import sys

try:
    from .. import ops  # type: ignore  # noqa: F401
except (ImportError, ValueError):
    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    from text_guided_face_recognition_b200 import ops  # noqa: F401
