"""Import shim: the package directory is ``tensor-fusion_b200/`` (the name the
build contract fixes), which is not a valid Python identifier.  This shim makes
``import tensor_fusion_b200.<module>`` resolve into ``tensor-fusion_b200/host``.
"""
import os as _os

REPO_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
PACKAGE_DIR = _os.path.join(REPO_ROOT, "tensor-fusion_b200")
__path__.append(_os.path.join(PACKAGE_DIR, "host"))
