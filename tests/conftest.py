import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import ctypes
        lib = ctypes.CDLL(os.path.join(ROOT, "tensor-fusion_b200", "lib", "libtfw_b200.so"))
        p = ctypes.c_void_p()
        rc = lib.tfw_host_alloc(ctypes.c_size_t(4096), ctypes.byref(p))
        if rc == 0:
            lib.tfw_host_free(p)
            return True
        return False
    except OSError:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a CPU box: they are skipped only when no
    # device exists; on a GPU box a missing extension is a hard failure (the
    # import of tensor_fusion_b200._native raises).
    if HAS_GPU or os.environ.get("TFW_RUN_GPU_MARKED"):  # the latter: provider tests under the mock NVML (test_cpu_provider_nvml.py)
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
