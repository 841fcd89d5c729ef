"""PSNR of nerfacto trained by the captured B200 step vs the autograd path over the drop-in modules, on a synthetic
scene rendered by a fixed random 'teacher' field (no dataset is available offline).  Prints PSNR on held-out rays."""
import copy, math, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfstudio_b200.nerfacto import NerfactoModel, NerfactoModelConfig, Trainer
from nerfstudio_b200.engine import NerfactoStep
from nerfstudio_b200.scene import bundle_from, synthetic_rays

def cfg():
    return NerfactoModelConfig(implementation="torch", average_init_density=0.01, num_levels=8, max_res=256,
                               log2_hashmap_size=15, background_color="black", use_appearance_embedding=False)
aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
torch.manual_seed(123)
teacher = NerfactoModel(cfg(), aabb, 8).cuda().eval()
with torch.no_grad():
    teacher.field.mlp_base.model[0].hash_table.mul_(3000.0)
    for p in teacher.proposal_networks:
        p.encoding.hash_table.mul_(3000.0)
R, NB = 4096, 64
def batch(seed):
    rays, _ = synthetic_rays(R, 8, seed)
    rays = {k: v.cuda() for k, v in rays.items()}
    with torch.no_grad():
        gt = teacher(bundle_from(rays))["rgb"]
    return rays, gt
train = [batch(s) for s in range(NB)]
held = [batch(10_000 + s) for s in range(4)]
def psnr(model):
    model.eval(); mse = 0.0
    with torch.no_grad():
        for rays, gt in held:
            mse += float(((model(bundle_from(rays))["rgb"] - gt) ** 2).mean())
    model.train(); return -10 * math.log10(mse / len(held))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
seeds = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [7]
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["graph", "autograd"]
out = {}
for name, seed in [(m, sd) for sd in seeds for m in modes]:
    torch.manual_seed(seed)
    student = NerfactoModel(cfg(), aabb, 8).cuda().train()
    if name != "autograd":
        eng = NerfactoStep(student, R, use_graph=(name == "graph"))
    else:
        eng = Trainer(student)
    torch.manual_seed(99)
    traj = [(0, psnr(student))]
    for it in range(steps):
        rays, gt = train[it % NB]
        if name != "autograd":
            eng.set_batch(rays["origins"], rays["directions"], rays["camera_indices"], gt); eng.step()
        else:
            eng.train_iteration(bundle_from(rays), {"image": gt})
        if (it + 1) % (steps // 4) == 0:
            traj.append((it + 1, psnr(student)))
    out[f"{name}/{seed}"] = traj
    print(name, seed, " ".join(f"{s}:{p:.2f}dB" for s, p in traj))
print(json.dumps(out))
