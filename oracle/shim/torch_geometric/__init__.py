"""TEST INFRASTRUCTURE ONLY (oracle shim) -- never imported by the product path.

Pure-PyTorch restatement of the torch-geometric==1.1.2 symbols the reference
imports (reference Dockerfile:32).  Restated from the published 1.1.x API and
the corroborating subclass code in the reference itself
(model/layers_meta.py:115-172).  The wheel is not under /root/reference and
cannot be fetched offline => the third-party arithmetic is "parity unpinned"
(SURVEY.md 8c, App. A); the reference's OWN code runs unmodified on top of it.
"""
__version__ = '1.1.2-shim'
from . import data, nn, utils, transforms  # noqa: F401
