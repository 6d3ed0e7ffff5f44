"""div_const (three FMAs) must equal IEEE division for every divisor the kernels use it with."""
import ctypes

import pytest

from fused_util import emu_backend


@pytest.mark.parametrize("d", [9.0, 639.0, 191.0, 1023.0, 319.0, 63.0, 31.0, 329.0, 199.0, 127.0, 39.0])
def test_div_const_is_correctly_rounded(d):
    dll = emu_backend().dll
    dll.emu_div_const_mismatches.restype = ctypes.c_long
    dll.emu_div_const_mismatches.argtypes = [ctypes.c_float]
    assert dll.emu_div_const_mismatches(d) == 0
