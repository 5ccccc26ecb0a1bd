"""``GaussianModel.ply`` -- the on-disk format of a GSORB-SLAM map, kept so that a map optimised through libgsb is read by the
reference's tooling (``scripts/replay.py:38-83`` via ``plyfile``) and vice versa.

Written by ``SavePly`` / ``WriteOutputPly`` (src/Utils.cc:182-280) with tinyply: one ``vertex`` element, binary little
endian, 14 ``float`` properties per Gaussian in this order -- ``x y z  rgb_0 rgb_1 rgb_2  opacity  scale_0 scale_1 scale_2
rot_0 rot_1 rot_2 rot_3`` (``ConstructListAttributes``, src/Utils.cc:211-228) -- holding the RAW parameters: world means,
colours, LOGIT opacity, LOG scales, UNNORMALISED quaternion (w, x, y, z).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

PROPERTIES = ("x", "y", "z", "rgb_0", "rgb_1", "rgb_2", "opacity", "scale_0", "scale_1", "scale_2",
              "rot_0", "rot_1", "rot_2", "rot_3")
_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1", "char": "i1",
              "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2", "int": "<i4", "int32": "<i4",
              "uint": "<u4", "uint32": "<u4"}


def save_gaussian_model(path: str, means, rgb, logit_opacities, log_scales, unnorm_quats) -> None:
    """Write the five parameter tensors (anything ``np.asarray`` accepts; torch tensors: pass ``.cpu().numpy()``)."""
    m = np.asarray(means, np.float32).reshape(-1, 3)
    P = m.shape[0]
    cols = [m, np.asarray(rgb, np.float32).reshape(P, 3), np.asarray(logit_opacities, np.float32).reshape(P, 1),
            np.asarray(log_scales, np.float32).reshape(P, 3), np.asarray(unnorm_quats, np.float32).reshape(P, 4)]
    rows = np.ascontiguousarray(np.concatenate(cols, 1), dtype="<f4")          # [P, 14], one record per Gaussian
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P
    header += "".join("property float %s\n" % name for name in PROPERTIES) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rows.tobytes())


def load_gaussian_model(path: str) -> Dict[str, np.ndarray]:
    """Read a ``GaussianModel.ply`` (properties looked up by NAME, in any order and of any scalar type, as ``plyfile`` does).
    Returns ``means [P,3], rgb [P,3], logit_opacities [P,1], log_scales [P,3], unnorm_quats [P,4]`` as float32."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    if not lines or lines[0].strip() != "ply":
        raise ValueError("not a PLY file")
    fmt = [l.split() for l in lines if l.startswith("format")]
    if not fmt or fmt[0][1] != "binary_little_endian":
        raise ValueError("only binary_little_endian PLY is supported (the reference writes nothing else)")
    count, props, in_vertex = 0, [], False
    for l in lines:
        tok = l.split()
        if not tok:
            continue
        if tok[0] == "element":
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                count = int(tok[2])
            elif props:
                break          # the vertex element is complete; later elements are not read
        elif tok[0] == "property" and in_vertex:
            if tok[1] == "list":
                raise ValueError("list properties are not part of GaussianModel.ply")
            props.append((tok[2], _PLY_TYPES[tok[1]]))
    rec = np.dtype(props)
    v = np.frombuffer(data, dtype=rec, count=count, offset=end)
    missing = [n for n in PROPERTIES if n not in rec.names]
    if missing:
        raise ValueError("GaussianModel.ply lacks properties: %s" % ", ".join(missing))
    col = lambda names: np.stack([v[n].astype(np.float32) for n in names], 1)
    return dict(means=col(("x", "y", "z")), rgb=col(("rgb_0", "rgb_1", "rgb_2")), logit_opacities=col(("opacity",)),
                log_scales=col(("scale_0", "scale_1", "scale_2")), unnorm_quats=col(("rot_0", "rot_1", "rot_2", "rot_3")))
