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
    assert dll.bbd_version() == 1
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
