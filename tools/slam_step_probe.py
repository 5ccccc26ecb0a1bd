import sys, os, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from gsorb_slam_b200.scene import make_config
from gsorb_slam_b200.mapping import MapOptimizer
from gsorb_slam_b200.lowlevel import frame_from_scene
dev = torch.device("cuda:0")
sc = make_config("headline_1m")
W, H = sc.cam.width, sc.cam.height
mo = MapOptimizer(sc.means3D, sc.colors, sc.logit_opacities, sc.log_scales, sc.unnorm_quats, width=W, height=H, tanfovx=sc.cam.tanfovx,
                  tanfovy=sc.cam.tanfovy, projmatrix=sc.cam.projmatrix, device=dev, max_rendered=4 * sc.P + 4096)
T = torch.eye(4, device=dev)
c, ds, med, _ = mo.render_fused(T)
gt_c = c.clone().clamp(0, 1); gt_d = torch.rand(H, W, device=dev) * 5 + 0.5
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    mo.step_slam(T, gt_c, gt_d)
torch.cuda.synchronize()
# tracking iterations on the same (now frozen) map
from gsorb_slam_b200.tracking import PoseOptimizer
po = PoseOptimizer(mo, [1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0])
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 0):
    po.step(gt_c, gt_d, 0.7, 1.0, True)
torch.cuda.synchronize()
