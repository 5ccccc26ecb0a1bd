"""Shared helpers of the parity tests."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL_CASES = ["tiny_default", "tiny_view_bg", "tiny_sh", "tiny_cov", "ragged_100x75"]

INT_KEYS = ["radii", "n_contrib", "ranges", "point_list", "tiles_touched"]
GEOM_KEYS = ["depths", "means2D", "conic_opacity"]
IMAGE_KEYS = ["color", "depth", "final_T"]
GRAD_KEYS = ["dL_dmean2D", "dL_dconic", "dL_dopacity", "dL_dcolor", "dL_dmean3D", "dL_dcov3D", "dL_dscale", "dL_drot", "dL_dsh"]

# Tolerances of BASELINE.json's north_star: RGB / depth / transmittance within 1e-4 relative,
# gradients within 1e-3 -- both relative to the tensor's own scale (max |reference|).
TOL_IMAGE = 1e-4
TOL_GRAD = 1e-3


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    kw = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    dL = kw.pop("dL_dpix")
    for k in ("width", "height", "sh_degree"):
        if k in kw:
            kw[k] = int(kw[k])
    for k in ("tanfovx", "tanfovy"):
        kw[k] = float(kw[k])
    return kw, dL, ref


def rel_to_scale(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    val = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    if os.environ.get("GSB_REL_LOG"):   # margins of the tolerances, per test (profiles/*_rel_margins.txt)
        with open(os.environ["GSB_REL_LOG"], "a") as f:
            f.write(f"{os.environ.get('PYTEST_CURRENT_TEST', '')}\t{val:.3e}\n")
    return val


def to_np(x):
    try:
        import torch
        if isinstance(x, torch.Tensor):
            return x.detach().cpu().numpy()
    except ImportError:
        pass
    return np.asarray(x)
