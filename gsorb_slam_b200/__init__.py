"""gsorb_slam_b200 -- B200-native (sm_100a) differentiable 3D-Gaussian rasterizer behind
GSORB-SLAM's operator surface.  The product is libgsb.so (include/gsb.h, csrc/); this package
is its host-side mirror of the reference interface (rasterizer.py), a raw C-ABI driver
(lowlevel.py), the multi-GPU gradient exchange (distributed.py) and the synthetic scene
generator used by tests and bench (scene.py)."""
__version__ = "0.1.0"
