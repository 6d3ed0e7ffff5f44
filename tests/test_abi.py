"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/bbd_loss.h declares.
No compute is launched here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest

from baseboostdepth_b200 import _lib
from baseboostdepth_b200 import build as bbd_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return bbd_build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bbd_loss.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bbd_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_symbol(lib_path):
    dll = ctypes.CDLL(lib_path)
    for sym in declared_symbols():
        assert hasattr(dll, sym), sym
    assert dll.bbd_version() == 2
    dll.bbd_reproj_tiles.restype = ctypes.c_int
    assert dll.bbd_reproj_tiles(192, 640) == 23 * 12      # 28x16 tiles


def test_struct_layouts_match_header(lib_path):
    # sizes computed by the C compiler for the same declarations
    import subprocess, tempfile
    src = '#include "bbd_loss.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu", sizeof(bbd_tables),' \
          'sizeof(bbd_ident_args), sizeof(bbd_reproj_args), sizeof(bbd_smooth_args), sizeof(bbd_d2d_args));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    mine = [ctypes.sizeof(t) for t in (_lib.Tables, _lib.IdentArgs, _lib.ReprojArgs, _lib.SmoothArgs, _lib.D2DArgs)]
    assert sizes == mine


def test_no_cpu_fallback_in_product():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _lib.cuda_backend()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "baseboostdepth_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_argument_errors_are_codes_not_crashes(lib_path):
    """Error behaviour of the boundary (include/bbd_loss.h): bad arguments give a negative BBD_E_* code and
    a message in bbd_last_error_string(); nothing throws, exits or touches the device.  (Argument checks
    run before any CUDA call, so this works in the GPU-less build container.)"""
    dll = ctypes.CDLL(lib_path)
    dll.bbd_last_error_string.restype = ctypes.c_char_p
    E_ARG, E_RANGE = -1, -2
    null = ctypes.c_void_p(None)

    assert dll.bbd_reproj_fused(null, null) == E_ARG
    assert b"reproj" in dll.bbd_last_error_string()
    assert dll.bbd_ident_forward(null, null) == E_ARG
    assert dll.bbd_smooth_fused(null, null) == E_ARG
    assert dll.bbd_disp_to_depth_forward(null, null) == E_ARG
    assert dll.bbd_disp_to_depth_backward(null, null) == E_ARG
    assert dll.bbd_u8_to_f32(null, null, ctypes.c_size_t(16), null) == E_ARG
    assert dll.bbd_u8_to_f32(null, null, ctypes.c_size_t(0), null) == E_ARG      # null wins over "nothing to do"

    # a structurally complete argument block with a candidate count outside the supported range
    ra = _lib.ReprojArgs()
    buf = (ctypes.c_float * 4)()
    tab = (ctypes.c_int32 * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p).value
    for name in ("target", "depth", "inv_K", "P", "ident_min", "loss_part"):
        setattr(ra, name, p)
    ra.tab.hdr = ra.tab.rep = ra.tab.ident = ctypes.cast(tab, ctypes.c_void_p).value
    ra.batch, ra.height, ra.width, ra.num_scales, ra.need_grad = 1, 32, 64, 1, 0
    ra.max_rep = 0
    assert dll.bbd_reproj_fused(ctypes.byref(ra), null) == E_RANGE
    ra.max_rep = _lib.MAX_REP + 1
    assert dll.bbd_reproj_fused(ctypes.byref(ra), null) == E_RANGE
    ra.max_rep, ra.height = 2, 1
    assert dll.bbd_reproj_fused(ctypes.byref(ra), null) == E_ARG                 # reflection padding needs >= 2 rows
    ra.height, ra.need_grad = 32, 1
    assert dll.bbd_reproj_fused(ctypes.byref(ra), null) == E_ARG                 # gradient buffers missing
    assert b"gradient" in dll.bbd_last_error_string()

    da = _lib.D2DArgs()
    da.batch, da.levels, da.height, da.width = 1, _lib.MAX_SCALES + 1, 32, 64
    da.depth = p
    assert dll.bbd_disp_to_depth_forward(ctypes.byref(da), null) == E_RANGE
    da.levels = 1
    da.h[0], da.w[0], da.disp[0] = 12, 20, p       # 32/12 is not an integer factor
    da.gdepth = da.gscale = p
    da.gdisp[0] = p
    assert dll.bbd_disp_to_depth_backward(ctypes.byref(da), null) == E_RANGE

    assert dll.bbd_loss_combine_forward(0, null, null, null, ctypes.c_float(4.0), null, null, null) == E_ARG
    assert dll.bbd_loss_combine_forward(9, p, p, p, ctypes.c_float(4.0), p, p, null) == E_RANGE
