"""GaussianModel.ply (src/Utils.cc:182-280, consumed by scripts/replay.py:38-83): byte layout and round trip."""
import os
import struct

import numpy as np
import pytest

from gsorb_slam_b200.ply import PROPERTIES, load_gaussian_model, save_gaussian_model


def _model(P, seed=0):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.normal(0, 1, s).astype(np.float32)
    return dict(means=f(P, 3), rgb=rng.random((P, 3)).astype(np.float32), logit_opacities=f(P, 1), log_scales=f(P, 3), unnorm_quats=f(P, 4))


@pytest.mark.parametrize("P", [0, 1, 1000])
def test_round_trip_and_byte_layout(tmp_path, P):
    m = _model(P)
    path = os.path.join(tmp_path, "GaussianModel.ply")
    save_gaussian_model(path, **m)
    raw = open(path, "rb").read()
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P + "".join("property float %s\n" % n for n in PROPERTIES)
              + "end_header\n").encode()
    assert raw.startswith(header) and len(raw) == len(header) + P * 14 * 4          # what tinyply writes for 14 float properties
    if P:
        first = struct.unpack("<14f", raw[len(header):len(header) + 56])             # one interleaved record per Gaussian
        want = np.concatenate([m["means"][0], m["rgb"][0], m["logit_opacities"][0], m["log_scales"][0], m["unnorm_quats"][0]])
        np.testing.assert_array_equal(np.array(first, np.float32), want)
    back = load_gaussian_model(path)
    for k in m:
        np.testing.assert_array_equal(back[k], m[k])


def test_loader_reads_properties_by_name_in_any_order(tmp_path):
    """plyfile (scripts/replay.py) looks properties up by name; so does the loader -- also with extra properties and doubles."""
    P = 17
    m = _model(P, seed=3)
    names = list(PROPERTIES)[::-1] + ["extra"]
    types = {n: "float" for n in names}
    types["opacity"] = "double"
    flat = dict(zip(PROPERTIES, np.concatenate([m["means"], m["rgb"], m["logit_opacities"], m["log_scales"], m["unnorm_quats"]], 1).T))
    flat["extra"] = np.zeros(P, np.float32)
    rec = np.zeros(P, dtype=[(n, "<f8" if types[n] == "double" else "<f4") for n in names])
    for n in names:
        rec[n] = flat[n]
    header = "ply\nformat binary_little_endian 1.0\ncomment written by a test\nelement vertex %d\n" % P
    header += "".join("property %s %s\n" % (types[n], n) for n in names) + "element face 0\nproperty list uchar int vertex_indices\nend_header\n"
    path = os.path.join(tmp_path, "m.ply")
    open(path, "wb").write(header.encode() + rec.tobytes())
    back = load_gaussian_model(path)
    for k in m:
        np.testing.assert_array_equal(back[k], m[k])
    with pytest.raises(ValueError):
        open(path, "wb").write(b"ply\nformat ascii 1.0\nelement vertex 0\nend_header\n")
        load_gaussian_model(path)
