"""TEST INFRASTRUCTURE ONLY (oracle shim): pickle-backed stand-in for klepto's
file_archive (reference utils/util.py:74-83)."""
from . import archives  # noqa: F401
