import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_sessionfinish(session, exitstatus):
    """With a -DLP_CHECKED library (LP_B200_LIB=...liblp_b200_checked.so) the kernels count violated index / capacity
    invariants on the device; report them at the end of the GPU run (compute-sanitizer is not available on this pool)."""
    if not os.environ.get("LP_B200_LIB"):
        return
    try:
        import ctypes
        import torch
        if not torch.cuda.is_available():
            return
        from latent_nerf_test_b200 import _lib
        line = ctypes.c_int32(0)
        n = _lib.lib().lp_check_failures(ctypes.byref(line))
        print(f"\n[lp_b200 checked build] device-side invariant violations: {n}" + (f" (first at lp_b200.cu:{line.value})" if n > 0 else ""))
        if n > 0:
            session.exitstatus = 1
    except Exception as exc:      # reporting only
        print(f"\n[lp_b200 checked build] could not read the violation counter: {exc}")
