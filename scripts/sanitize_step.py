"""Two eager nerfacto steps at a tiny size, for compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfstudio_b200.nerfacto import NerfactoModel, NerfactoModelConfig
from nerfstudio_b200.engine import NerfactoStep
from nerfstudio_b200.scene import synthetic_rays
torch.manual_seed(0)
cfg = NerfactoModelConfig(implementation="torch", average_init_density=0.01, num_levels=4, max_res=128, log2_hashmap_size=12,
                          proposal_net_args_list=[
                              {"hidden_dim": 16, "log2_hashmap_size": 10, "num_levels": 5, "max_res": 64, "use_linear": False},
                              {"hidden_dim": 16, "log2_hashmap_size": 10, "num_levels": 5, "max_res": 128, "use_linear": False}])
model = NerfactoModel(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), 8).cuda().train()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 300
eng = NerfactoStep(model, R, use_graph=False, always_update_proposals=True)
rays, gt = synthetic_rays(R, 8, 0)
eng.set_batch(rays["origins"].cuda(), rays["directions"].cuda(), rays["camera_indices"].cuda(), gt.cuda())
for _ in range(2):
    l = eng.step()
torch.cuda.synchronize()
print("losses", [float(v) for v in l])
