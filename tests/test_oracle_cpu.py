"""CPU suite: the oracle (oracle/gs_oracle.cpp) against the golden outputs of the reference
kernels (tests/golden/*.npz, generated on a B200 by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from helpers import GEOM_KEYS, GRAD_KEYS, IMAGE_KEYS, SMALL_CASES, TOL_GRAD, TOL_IMAGE, load_case, rel_to_scale


def run_oracle(kw, dL):
    from oracle import gs_oracle
    fr = gs_oracle.OracleFrame(**kw)
    g = fr.backward(dL)
    out = dict(color=fr.color, depth=fr.depth, radii=fr.radii, num_rendered=fr.num_rendered)
    out.update(fr.image_state())
    out.update(fr.binning())
    out.update(fr.geometry())
    out.update({k: v for k, v in g.items() if v is not None})
    return out


@pytest.mark.parametrize("name", SMALL_CASES)
def test_oracle_matches_reference_golden(name):
    kw, dL, ref = load_case(name)
    out = run_oracle(kw, dL)
    assert int(out["num_rendered"]) == int(ref["num_rendered"])
    for k in ("radii", "tiles_touched", "point_list", "ranges", "n_contrib"):
        np.testing.assert_array_equal(np.asarray(out[k]).astype(np.int64).ravel(), ref[k].astype(np.int64).ravel(), err_msg=k)
    for k in GEOM_KEYS:
        np.testing.assert_array_equal(np.asarray(out[k], np.float32).ravel().view(np.uint32), ref[k].ravel().view(np.uint32), err_msg=k)
    for k in IMAGE_KEYS:
        assert rel_to_scale(out[k], ref[k]) <= TOL_IMAGE, k
    for k in GRAD_KEYS:
        if k in out and k in ref and (k != "dL_dsh" or "shs" in kw):
            assert rel_to_scale(out[k], np.asarray(ref[k]).reshape(np.asarray(out[k]).shape)) <= TOL_GRAD, k


def test_oracle_knn_and_visibility_golden():
    from oracle import gs_oracle
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "knn_5000.npz"))
    np.testing.assert_array_equal(gs_oracle.knn_mean_dist2(z["in_points"]), z["ref_mean_dist2"])
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "visible_4000.npz"))
    kw = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    kw["width"], kw["height"] = int(kw["width"]), int(kw["height"])
    kw["tanfovx"], kw["tanfovy"] = float(kw["tanfovx"]), float(kw["tanfovy"])
    np.testing.assert_array_equal(gs_oracle.visible_filter(**kw), z["ref_radii"])
    np.testing.assert_array_equal(gs_oracle.mark_visible(kw["means3D"], kw["viewmatrix"], kw["projmatrix"]), z["ref_present"])


def test_config0_naive_cpu_forward():
    """BASELINE.json configs[0]: 256 Gaussians, 64x64, forward alpha-blend on the host CPU."""
    from gsorb_slam_b200.scene import make_config
    from oracle import gs_oracle
    sc = make_config("cfg0_tiny")
    fr = gs_oracle.frame_from_scene(sc)
    assert fr.color.shape == (3, 64, 64) and np.isfinite(fr.color).all()
    assert fr.num_rendered > 256 and (fr.radii > 0).sum() > 200
    T = fr.image_state()["final_T"]
    assert (T >= 0).all() and (T <= 1).all()
    # colour + T*bg with bg = 0 and colours in [0,1]: sum of weights = 1 - T bounds every channel
    assert (fr.color <= (1 - T)[None] + 1e-5).all()
