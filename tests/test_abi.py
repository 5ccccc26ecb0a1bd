"""CPU suite: the C-ABI library loads and exports every symbol include/gsb.h declares; the pure
host-side entry points (sizes, validation, error text) behave.  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gsb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", src)) - {"gsb_alloc_fn"})


def test_library_exports_every_declared_symbol():
    from gsorb_slam_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libgsb.so not built (make -C gsorb_slam_b200/csrc)"
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(L, n), f"{n} declared in gsb.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of sync with gsb.h"


def test_version_and_sizes():
    from gsorb_slam_b200 import _lib
    L = _lib.lib()
    assert L.gsb_version() == (0 << 16) | 2
    g, i, b = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert L.gsb_workspace_query(1000, 640, 480, 5000, C.byref(g), C.byref(i), C.byref(b)) == 0
    assert g.value == L.gsb_geometry_bytes(1000) and g.value >= 1000 * 48
    assert i.value == L.gsb_image_bytes(640, 480) and i.value >= 640 * 480 * 8
    assert b.value == L.gsb_binning_bytes(5000) and b.value >= 5000 * 12
    assert L.gsb_geometry_bytes(0) > 0
    assert L.gsb_workspace_query(-1, 640, 480, 0, None, None, None) == -1
    assert b"bad sizes" in L.gsb_last_error()


def test_validation_errors_match_reference_messages():
    """Argument checks happen before any CUDA call, so they can be exercised without a GPU."""
    from gsorb_slam_b200 import _lib
    L = _lib.lib()
    a = _lib.RasterArgs()
    a.P, a.width, a.height, a.tan_fovx, a.tan_fovy = 4, 64, 64, 0.5, 0.5
    dummy = C.c_void_p(256)   # never dereferenced: validation fails first
    a.viewmatrix = a.projmatrix = a.background = a.means3D = a.opacities = dummy
    a.scales = a.rotations = dummy
    # neither SHs nor colours
    rc = L.gsb_forward_ws(C.byref(a), dummy, 1 << 30, dummy, 1 << 30, 100, dummy, 1 << 30, dummy, dummy, None, None)
    assert rc == -1 and b"exactly one of either SHs or precomputed colors" in L.gsb_last_error()
    a.colors_precomp = dummy
    a.cov3D_precomp = dummy   # both scale/rotation and cov3D
    rc = L.gsb_forward_ws(C.byref(a), dummy, 1 << 30, dummy, 1 << 30, 100, dummy, 1 << 30, dummy, dummy, None, None)
    assert rc == -1 and b"scale/rotation pair or precomputed 3D covariance" in L.gsb_last_error()
    a.cov3D_precomp = None
    a.width = 0
    rc = L.gsb_forward_ws(C.byref(a), dummy, 1 << 30, dummy, 1 << 30, 100, dummy, 1 << 30, dummy, dummy, None, None)
    assert rc == -1
    a.width = 64
    rc = L.gsb_forward_ws(C.byref(a), dummy, 16, dummy, 1 << 30, 100, dummy, 1 << 30, dummy, dummy, None, None)
    assert rc == -3 and b"geometry workspace too small" in L.gsb_last_error()
    with pytest.raises(ValueError):
        _lib.check(-1)
    # extensions: fused pass needs its extra output / gradient input, tile-row band must lie inside the image, exchange arguments
    rc = L.gsb_forward_fused_ws(C.byref(a), dummy, 1 << 30, dummy, 1 << 30, 100, dummy, 1 << 30, dummy, None, dummy, None, None)
    assert rc == -1 and b"out_depth_sil" in L.gsb_last_error()
    g = _lib.GradOutputs()
    rc = L.gsb_backward_fused(C.byref(a), None, dummy, dummy, dummy, dummy, None, C.byref(g), None, 0, None)
    assert rc == -1 and b"dL_ddepth_sil" in L.gsb_last_error()
    a.tile_row_begin, a.tile_row_end = 2, 9   # 64 px = 4 tile rows
    rc = L.gsb_forward_ws(C.byref(a), dummy, 1 << 30, dummy, 1 << 30, 100, dummy, 1 << 30, dummy, dummy, None, None)
    assert rc == -1 and b"tile row band" in L.gsb_last_error()
    a.tile_row_begin, a.tile_row_end = 0, 0
    # pointer alignment the kernels rely on (128-bit loads of rotations, 64-bit loads of cov3D_precomp)
    a.rotations = C.c_void_p(256 + 4)
    rc = L.gsb_forward_ws(C.byref(a), dummy, 1 << 30, dummy, 1 << 30, 100, dummy, 1 << 30, dummy, dummy, None, None)
    assert rc == -1 and b"rotations must be 16-byte aligned" in L.gsb_last_error()
    a.rotations = a.scales = None
    a.cov3D_precomp = C.c_void_p(256 + 4)
    rc = L.gsb_forward_ws(C.byref(a), dummy, 1 << 30, dummy, 1 << 30, 100, dummy, 1 << 30, dummy, dummy, None, None)
    assert rc == -1 and b"cov3D_precomp must be 8-byte aligned" in L.gsb_last_error()
    a.cov3D_precomp = None
    a.rotations = a.scales = dummy
    ptrs = (C.c_void_p * 2)(256, 512)
    assert L.gsb_exchange_allreduce(None, ptrs, ptrs, 1002, 0, 2, None) == -1          # n % 4 != 0
    assert L.gsb_exchange_allreduce(None, ptrs, ptrs, 1000, 2, 2, None) == -1          # rank out of range
    assert L.gsb_exchange_allreduce(None, None, None, 1000, 0, 2, None) == -1          # missing mappings
    assert L.gsb_exchange_allreduce(None, None, None, 1000, 0, 1, None) == 0           # world 1: no-op
    assert L.gsb_exchange_sync_bytes(8) >= 2 * 148 * 8 * 4


def test_operator_surface_mirrors_reference_names():
    from gsorb_slam_b200 import rasterizer as R
    for n in ("GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "_RasterizeGaussians", "distCUDA2"):
        assert hasattr(R, n)
    for n in ("forward", "Visable", "mark_visible"):
        assert hasattr(R.GaussianRasterizer, n)
    import torch
    rs = R.GaussianRasterizationSettings(64, 64, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False)
    r = R.GaussianRasterizer(rs)
    m = torch.zeros(4, 3)
    with pytest.raises(ValueError, match="exactly one of either SHs or precomputed colors"):
        r.forward(m, m, torch.zeros(4, 1), scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(ValueError, match="scale/rotation pair or precomputed 3D covariance"):
        r.forward(m, m, torch.zeros(4, 1), colors_precomp=m)


def test_libtorch_adapter_builds_and_exposes_the_reference_surface():
    """adapter/gsb_adapter.so = adapter/Rasterizer.{cuh,cc} (the drop-in for include/Rasterizer.cuh + src/Rasterizer.cu +
    src/spatial.cu) compiled against the installed libtorch, plus a pybind harness; built by __graft_entry__.build()."""
    import importlib
    import os
    import sys
    import torch  # noqa: F401
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "adapter", "gsb_adapter.so")
    if not os.path.exists(so):
        import pytest
        pytest.skip("adapter not built (run __graft_entry__.build())")
    sys.path.insert(0, os.path.join(root, "adapter"))
    A = importlib.import_module("gsb_adapter")
    for name in ("forward", "forward_fused", "visable", "mark_visible", "dist_cuda2", "GradientExchange"):
        assert hasattr(A, name)
    src = open(os.path.join(root, "adapter", "Rasterizer.cuh")).read()
    for name in ("GaussianRasterizationSettings", "GaussianRasterizer", "_RasterizeGaussians", "rasterize_gaussians", "filter_radii",
                 "RasterizeGaussiansCUDA", "RasterizeGaussiansBackwardCUDA", "RasterizeGaussiansfilterCUDA", "markVisible",
                 "Visable", "mark_visible", "distCUDA2"):   # names of include/Rasterizer.cuh:28-380 + include/spatial.h
        assert name in src, name


def test_header_is_plain_c99_and_a_c_caller_links(tmp_path):
    """The boundary is a C ABI: include/gsb.h compiles as strict C99 (no C++-isms, no torch types), and a C program linked
    against libgsb.so reaches the host-only entry points (version, blob sizes, argument validation with its error text)."""
    import shutil
    import subprocess
    from gsorb_slam_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "gsb.h"
int main(void)
{
    gsb_raster_args a;
    memset(&a, 0, sizeof a);                       /* P = 0, no pointers, 0 x 0 image: must be refused, not crash */
    if (gsb_version() <= 0) return 1;
    if (gsb_image_bytes(640, 480) == 0 || gsb_geometry_bytes(1000) == 0 || gsb_binning_bytes(4096) == 0) return 2;
    if (gsb_forward_ws(&a, NULL, 0, NULL, 0, 0, NULL, 0, NULL, NULL, NULL, NULL) != GSB_ERR_INVALID_ARGUMENT) return 3;
    if (strstr(gsb_last_error(), "image size") == NULL) return 4;
    printf("%d %lu\n", gsb_version(), (unsigned long)gsb_image_bytes(640, 480));
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L", libdir, "-lgsb", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    ver, nbytes = out.stdout.split()
    L = _lib.lib()
    assert int(ver) == L.gsb_version() and int(nbytes) == L.gsb_image_bytes(640, 480)


def test_no_entry_point_dereferences_null_arguments():
    """Every entry point of include/gsb.h called with NULL for every pointer -- once with zero sizes, once with non-zero ones --
    must come back with a status (GSB_OK for "nothing to do", GSB_ERR_INVALID_ARGUMENT otherwise; a pure size query returns its
    size), never dereference on the host.  Run in a child process so that a crash is reported by name.  (gsb_forward takes
    callbacks and is covered by test_validation_errors_match_reference_messages.)"""
    import subprocess
    import sys
    child = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
from gsorb_slam_b200 import _lib
L = _lib.lib()
for fill in (0, 5):
    for name in sorted(_lib.SIGNATURES):
        if name == "gsb_forward":
            continue
        res, argt = _lib.SIGNATURES[name]
        args = [fill if a in (C.c_int, C.c_longlong, C.c_size_t, C.c_uint) else (0.5 * fill if a in (C.c_float, C.c_double) else None) for a in argt]
        print("calling", name, fill, flush=True)
        r = getattr(L, name)(*args)
        if res is C.c_int and name not in ("gsb_version", "gsb_num_stages"):
            assert r in (0, -1, -2), (name, r)      # -2: a launch on all-NULL outputs found no driver (CPU box); never reached with real work
            if r == -1:
                assert L.gsb_last_error(), name
print("done")
''' % ROOT
    p = subprocess.run([sys.executable, "-c", child], capture_output=True, text=True)
    last = [l for l in p.stdout.splitlines() if l.startswith("calling")][-1:] or ["(none)"]
    assert p.returncode == 0 and p.stdout.strip().endswith("done"), (p.returncode, last, p.stderr[-600:])
