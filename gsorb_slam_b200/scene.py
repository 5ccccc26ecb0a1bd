"""Seeded synthetic "SLAM-like" Gaussian maps and cameras (numpy, host side).

There are no datasets on the GPU box, so every test / bench input is produced here.
The distributions restate what GSORB-SLAM itself produces:

* camera: `Camera::Camera` (src/Camera.cc:7-41) -- symmetric frustum, tanfov = W/(2 fx),
  near 0.01 / far 100; the projection tensor holds P^T, i.e. the 16 floats the kernels
  read column-major (auxiliary.h:58-77).  Default mode of `Render::StartSplatting`
  (src/Render.cc:748-754): identity view matrix, means pre-transformed into the camera
  frame, campos = 0.
* Gaussians: what `Render::InitGaussianPoint` + `Gaussian::AddGaussianPoints` create
  (src/Render.cc:666-707, src/Gaussian.cc:50-74): back-projected pixels, isotropic
  log_scale = log(z / f) (sigma ~ 1 px) -- here with N(0, 0.3) jitter per axis so the
  covariance path is exercised -- random unit quaternions, logit-opacity ~ N(1, 1).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# (W, H, fx, fy) of the reference's YAMLs / BASELINE.json configs
INTRINSICS = {
    "tiny": (64, 64, 64.0, 64.0),                    # config #0
    "tum": (640, 480, 517.306408, 516.469215),       # Examples/RGB-D/tum/TUM1.yaml:13-16
    "replica": (1200, 680, 600.0, 600.0),            # Examples/RGB-D/replica.yaml:12-17
    "scannet_hr": (1296, 968, 1165.0, 1165.0),       # synthetic scale-up (BASELINE.json config #4)
}


@dataclass
class Camera:
    width: int
    height: int
    fx: float
    fy: float
    near: float = 0.01
    far: float = 100.0
    tanfovx: float = field(init=False)
    tanfovy: float = field(init=False)
    viewmatrix: np.ndarray = field(init=False)   # [16] as the kernels read it (column-major)
    projmatrix: np.ndarray = field(init=False)   # [16] full projection, same convention
    campos: np.ndarray = field(init=False)       # [3]

    def __post_init__(self):
        self.tanfovx = np.float32(self.width / (2.0 * self.fx))
        self.tanfovy = np.float32(self.height / (2.0 * self.fy))
        self.set_pose(np.eye(4, dtype=np.float32))

    def proj_rowmajor(self) -> np.ndarray:
        """The OpenGL-style matrix of src/Camera.cc:26-29 in ordinary row-major maths."""
        n, f = np.float32(self.near), np.float32(self.far)
        top = self.tanfovy * n
        right = self.tanfovx * n
        P = np.zeros((4, 4), dtype=np.float32)
        P[0, 0] = 2 * n / (2 * right)
        P[1, 1] = 2 * n / (2 * top)
        P[2, 2] = f / (f - n)
        P[2, 3] = -(f * n) / (f - n)
        P[3, 2] = 1.0
        return P

    def set_pose(self, Tcw: np.ndarray) -> None:
        """Camera::SetPose (src/Camera.cc:44-52): viewmatrix tensor = Tcw^T, full = view @ proj."""
        Tcw = np.asarray(Tcw, dtype=np.float32)
        view_t = Tcw.T.copy()                       # tensor rows = columns of Tcw
        proj_t = self.proj_rowmajor().T.copy()      # Eigen column-major blob read row-major
        full_t = (view_t @ proj_t).astype(np.float32)
        self.viewmatrix = view_t.reshape(16).copy()
        self.projmatrix = full_t.reshape(16).copy()
        self.campos = np.linalg.inv(Tcw.astype(np.float64))[:3, 3].astype(np.float32)


@dataclass
class Scene:
    cam: Camera
    means3D: np.ndarray        # [P,3] camera-frame means (what the rasterizer receives)
    scales: np.ndarray         # [P,3] exp(log_scales)
    rotations: np.ndarray      # [P,4] normalised (w,x,y,z)
    opacities: np.ndarray      # [P]   sigmoid(logit)
    colors: np.ndarray         # [P,3]
    background: np.ndarray     # [3]
    dL_dpix: np.ndarray        # [3,H,W]
    # raw (pre-activation) parameters, for the prologue / Adam pieces
    log_scales: np.ndarray = None
    unnorm_quats: np.ndarray = None
    logit_opacities: np.ndarray = None

    @property
    def P(self) -> int:
        return int(self.means3D.shape[0])


def make_scene(P: int, intr: str | tuple = "tum", seed: int = 0, scale_mul: float = 1.0,
               cull_frac: float = 0.02, background: float = 0.0, z_range=(0.5, 6.0),
               scale_jitter: float = 0.3) -> Scene:
    """SLAM-like map of P Gaussians seen by one camera (SURVEY.md 8d "Synthetic inputs")."""
    W, H, fx, fy = INTRINSICS[intr] if isinstance(intr, str) else intr
    cam = Camera(int(W), int(H), float(fx), float(fy))
    rng = np.random.default_rng(seed)
    f32 = np.float32
    # pixels uniform over the image extended by 5 %
    u = rng.uniform(-0.05 * W, 1.05 * W, P).astype(f32)
    v = rng.uniform(-0.05 * H, 1.05 * H, P).astype(f32)
    z = rng.uniform(z_range[0], z_range[1], P).astype(f32)
    ncull = int(round(cull_frac * P))
    if ncull:
        idx = rng.choice(P, ncull, replace=False)
        z[idx] = rng.uniform(-1.0, 0.2, ncull).astype(f32)   # exercises the z <= 0.2 cull
    # principal point is the image centre by construction (src/Camera.cc:19-31)
    cx, cy = f32((W - 1) / 2.0), f32((H - 1) / 2.0)
    x = (u - cx) / f32(fx) * z
    y = (v - cy) / f32(fy) * z
    means = np.stack([x, y, z], 1).astype(f32)
    fm = f32((fx + fy) / 2.0)
    base = np.log(np.maximum(np.abs(z), f32(0.05)) / fm).astype(f32)
    log_scales = (base[:, None] + rng.normal(0.0, scale_jitter, (P, 3)).astype(f32)
                  + f32(np.log(scale_mul))).astype(f32)
    q = rng.normal(0.0, 1.0, (P, 4)).astype(f32)
    logit = rng.normal(1.0, 1.0, P).astype(f32)
    colors = rng.uniform(0.0, 1.0, (P, 3)).astype(f32)
    dL = (rng.normal(0.0, 1.0, (3, H, W)) / (H * W)).astype(f32)
    qn = (q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), f32(1e-12))).astype(f32)
    return Scene(cam=cam, means3D=means, scales=np.exp(log_scales).astype(f32), rotations=qn,
                 opacities=(1.0 / (1.0 + np.exp(-logit))).astype(f32), colors=colors,
                 background=np.full(3, background, dtype=f32), dL_dpix=dL,
                 log_scales=log_scales, unnorm_quats=q, logit_opacities=logit)


# The five BASELINE.json configs plus the headline metric, as generator arguments.
CONFIGS = {
    "cfg0_tiny": dict(P=256, intr="tiny", scale_mul=4.0),
    "cfg1_100k": dict(P=100_000, intr="tum"),
    "cfg2_500k": dict(P=500_000, intr="tum"),
    "headline_1m": dict(P=1_000_000, intr="tum"),
    "cfg3_2m": dict(P=2_000_000, intr="replica"),
    "cfg4_5m": dict(P=5_000_000, intr="scannet_hr"),
    # stress variants of the headline map: 3x larger splats (long tile lists: the 128-KB sort class, saturated pixels)
    # and a close-up where a third of the Gaussians are outside the frustum
    "dense_1m": dict(P=1_000_000, intr="tum", scale_mul=3.0),
    "culled_1m": dict(P=1_000_000, intr="tum", cull_frac=0.35),
}


def raster_ordered(sc: Scene) -> Scene:
    """Same map, Gaussians re-ordered tile by tile in raster order -- the memory order a real GSORB-SLAM map has:
    `Render::InitGaussianPoint` / `AddGaussian` create Gaussians pixel by pixel (src/Render.cc:666-707, 617-655), so
    neighbours in memory are neighbours on screen.  Exercises same-address contention of the binning atomics."""
    cam = sc.cam
    z = np.maximum(sc.means3D[:, 2], 1e-3)
    u = sc.means3D[:, 0] / z * cam.fx + (cam.width - 1) / 2.0
    v = sc.means3D[:, 1] / z * cam.fy + (cam.height - 1) / 2.0
    key = (np.clip(v, -64, cam.height + 64).astype(np.int64) + 64) * (cam.width + 256) + np.clip(u, -64, cam.width + 64).astype(np.int64) + 64
    order = np.argsort(key, kind="stable")
    for name in ("means3D", "scales", "rotations", "opacities", "colors", "log_scales", "unnorm_quats", "logit_opacities"):
        setattr(sc, name, np.ascontiguousarray(getattr(sc, name)[order]))
    return sc


def make_config(name: str, seed: int = 0) -> Scene:
    if name.endswith("_raster"):
        return raster_ordered(make_scene(seed=seed, **CONFIGS[name[:-len("_raster")]]))
    return make_scene(seed=seed, **CONFIGS[name])


# ---- BASELINE.json config #2 ("camera-pose backward enabled"): a camera that is NOT at the origin ----------------------
def cfg2_pose() -> np.ndarray:
    ay, ax = 0.07, -0.03
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    T = np.eye(4)
    T[:3, :3] = Ry @ Rx
    T[:3, 3] = [0.12, -0.04, 0.25]
    return T.astype(np.float32)


def pose_transform_f32(Tcw: np.ndarray, means_world: np.ndarray) -> np.ndarray:
    """src/Render.cc:750-752 (``Tcw.repeat(N,1,1).bmm([mean;1])``) restated with a FIXED fp32 operation order (products
    rounded, summed left to right) so that the golden generator and the tests hand bit-identical camera-frame means to the
    reference kernels and to libgsb."""
    T = Tcw.astype(np.float32)
    m = means_world.astype(np.float32)
    out = np.empty_like(m)
    for r in range(3):
        out[:, r] = ((T[r, 0] * m[:, 0] + T[r, 1] * m[:, 1]) + T[r, 2] * m[:, 2]) + T[r, 3]
    return out


def quantised_raster_scene(P: int, intr: str = "tum", seed: int = 0, depth_factor: float = 5000.0, planes: int = 24) -> Scene:
    """What ``Render::InitWorld`` / ``InitGaussianPoint`` (src/Render.cc:496-553, 666-707) really produce: one Gaussian per
    valid depth pixel, back-projected at an identity pose from a depth image whose values are QUANTISED (TUM
    ``DepthMapFactor`` 5000, Replica 16-bit PNG), created in raster order.  Piecewise-planar depth (``planes`` fronto-
    parallel patches) so that a tile holds hundreds of EXACTLY equal view-space depths -- the worst case of the tile
    sort's tie handling, which continuous synthetic depths never exercise."""
    sc = make_scene(P, intr, seed=seed)
    cam = sc.cam
    rng = np.random.default_rng(seed + 1000)
    f32 = np.float32
    z = sc.means3D[:, 2].copy()
    vis = z > 0.2
    u = sc.means3D[:, 0] / np.where(vis, z, 1) * f32(cam.fx) + f32((cam.width - 1) / 2.0)
    v = sc.means3D[:, 1] / np.where(vis, z, 1) * f32(cam.fy) + f32((cam.height - 1) / 2.0)
    # depth = one of `planes` levels chosen by the screen cell (6 x 4 patches), then quantised to 1/depth_factor metres
    cell = (np.clip(u / cam.width * 6, 0, 5.999).astype(np.int64) + 6 * np.clip(v / cam.height * 4, 0, 3.999).astype(np.int64)) % planes
    levels = rng.uniform(0.8, 5.0, planes)
    zq = (np.round(levels[cell] * depth_factor) / depth_factor).astype(f32)
    znew = np.where(vis, zq, z).astype(f32)
    sc.means3D[:, 0] = np.where(vis, (u - f32((cam.width - 1) / 2.0)) / f32(cam.fx) * znew, sc.means3D[:, 0])
    sc.means3D[:, 1] = np.where(vis, (v - f32((cam.height - 1) / 2.0)) / f32(cam.fy) * znew, sc.means3D[:, 1])
    sc.means3D[:, 2] = znew
    fm = f32((cam.fx + cam.fy) / 2.0)
    sc.log_scales = np.repeat(np.log(np.maximum(np.abs(znew), f32(0.05)) / fm)[:, None], 3, 1).astype(f32)   # isotropic, Gaussian.cc:60-66
    sc.scales = np.exp(sc.log_scales).astype(f32)
    return raster_ordered(sc)


def make_large_case(name: str):
    """name -> (scene whose ``means3D`` are the camera-frame means handed to the rasterizer, extras).  The cases behind
    tests/golden/large_digests.json; ``tum_<P>`` are the round-1 names of cfg1_100k / headline_1m."""
    if name.startswith("tum_"):
        return make_scene(int(name[4:]), "tum", seed=0), {}
    if name == "cfg2_500k_pose":
        sc = make_config("cfg2_500k")
        Tcw = cfg2_pose()
        Twc = np.linalg.inv(Tcw.astype(np.float64))
        means_world = (sc.means3D.astype(np.float64) @ Twc[:3, :3].T + Twc[:3, 3]).astype(np.float32)
        sc.means3D = pose_transform_f32(Tcw, means_world)
        return sc, dict(Tcw=Tcw, means_world=means_world)
    if name == "quantised_1m":
        return quantised_raster_scene(1_000_000), {}
    return make_config(name), {}
