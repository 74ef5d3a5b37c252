"""Runs last (-m gpu): with a -DLP_CHECKED library (``LP_B200_LIB=.../liblp_b200_checked.so``) every kernel invariant
violated during the whole GPU test session has been counted on the device; there must be none.  With the ordinary library
the checks are compiled out and the counter stays 0."""
import ctypes

import pytest
import torch

from latent_nerf_test_b200 import _lib

pytestmark = pytest.mark.gpu


def test_no_device_side_invariant_violations():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    line = ctypes.c_int32(0)
    n = _lib.lib().lp_check_failures(ctypes.byref(line))
    assert n == 0, f"{n} violated kernel invariants, first at lp_b200.cu:{line.value}"
