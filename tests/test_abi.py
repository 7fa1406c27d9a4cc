"""CPU tests: the C-ABI shared library builds, loads without a GPU and exports every symbol that
include/t2v_b200.h declares; argument validation fails loudly (no compute is attempted)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from t2v_turbo_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    from t2v_turbo_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "t2v_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(t2v_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.t2v_version() >= 100


def test_struct_sizes_match_header(tmp_path):
    """Compile the header with gcc and compare struct sizes with the ctypes mirrors."""
    import subprocess
    from t2v_turbo_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "t2v_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(T2VGemmDesc),sizeof(T2VAttnDesc),sizeof(T2VShortAttnDesc),sizeof(T2VGroupNormDesc),'
                   'sizeof(T2VLayerNormDesc),sizeof(T2VSmallLinearDesc),sizeof(T2VWgradDesc),sizeof(T2VGroupNormBwdDesc),sizeof(T2VAttnBwdDesc),'
                   'sizeof(T2VShortAttnBwdDesc));return 0;}')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    mirrors = [_lib.GemmDesc, _lib.AttnDesc, _lib.ShortAttnDesc, _lib.GroupNormDesc, _lib.LayerNormDesc, _lib.SmallLinearDesc, _lib.WgradDesc,
               _lib.GroupNormBwdDesc, _lib.AttnBwdDesc, _lib.ShortAttnBwdDesc]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_argument_errors_are_loud(lib):
    from t2v_turbo_b200 import _lib
    d = _lib.GemmDesc()
    rc = lib.t2v_gemm(ctypes.byref(d), None)
    assert rc < 0 and b"null" in lib.t2v_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "t2v_gemm")
    assert lib.t2v_lcm_step(None, None, None, None, None, 0, 0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, None) < 0


def test_workspace_size_queries(lib):
    """Host-only entry points: what a non-Python caller must allocate (no device needed)."""
    from t2v_turbo_b200 import _lib
    d = _lib.GemmDesc()
    assert lib.t2v_gemm_workspace_bytes(ctypes.byref(d)) == -1 and lib.t2v_gemm_workspace_bytes(None) == -1      # o_size / b_rows unset
    for j, v in enumerate((160, 4, 1, 1)):
        d.o_size[j] = v
    d.b_rows = 1280
    assert lib.t2v_gemm_workspace_bytes(ctypes.byref(d)) == 640 * 1280 * 4                 # level-3 Linear: fp32 [points][N]
    assert lib.t2v_groupnorm_workspace_bytes(16, 32, 0) == (16 * 32 * 2 + 1) * 4
    assert lib.t2v_groupnorm_workspace_bytes(16, 32, 1) == 16 * 32 * 4 * 4
    assert lib.t2v_groupnorm_workspace_bytes(0, 32, 0) == -1


def test_product_has_no_oracle_or_cpu_fallback():
    """The shipped package must not import the oracle (it is the checker, not the product)."""
    pkg = os.path.join(ROOT, "t2v_turbo_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_plan_box_covers_unet_geometries():
    from t2v_turbo_b200.ops import plan_box
    for sizes in [(64, 40, 16, 1), (32, 20, 16, 1), (16, 10, 16, 1), (8, 5, 16, 1), (2560, 16, 1, 1), (640, 16, 1, 1),
                  (160, 16, 1, 1), (40, 16, 1, 1), (512, 320, 16, 1)]:
        b = plan_box(sizes)
        rows = b[0] * b[1] * b[2] * b[3]
        assert rows <= 128 and rows % 8 == 0
        tiles = 1
        for s, x in zip(sizes, b):
            tiles *= -(-s // x)
        total = sizes[0] * sizes[1] * sizes[2] * sizes[3]
        assert total / (tiles * 128) >= 0.9, (sizes, b)
