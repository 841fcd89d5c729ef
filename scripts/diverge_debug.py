import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfstudio_b200.nerfacto import NerfactoModel, NerfactoModelConfig
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
    for p in teacher.proposal_networks: p.encoding.hash_table.mul_(3000.0)
R, NB = 4096, 64
def batch(seed):
    rays, _ = synthetic_rays(R, 8, seed); rays = {k: v.cuda() for k, v in rays.items()}
    with torch.no_grad(): gt = teacher(bundle_from(rays))["rgb"]
    return rays, gt
train = [batch(s) for s in range(NB)]
mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
kw = dict(use_graph=(mode != "eager"))
if mode == "simt": kw["mlp_backend"] = "simt"
if mode == "unfused": kw["fused_proposals"] = False
if mode == "always": kw["always_update_proposals"] = True
torch.manual_seed(7)
student = NerfactoModel(cfg(), aabb, 8).cuda().train()
eng = NerfactoStep(student, R, **kw)
torch.manual_seed(99)
for it in range(600):
    rays, gt = train[it % NB]
    eng.set_batch(rays["origins"], rays["directions"], rays["camera_indices"], gt)
    l = eng.step().clone(); torch.cuda.synchronize()
    if os.environ.get("EVAL_AT") and it == int(os.environ["EVAL_AT"]):
        student.eval()
        with torch.no_grad():
            r2, g2 = batch(10_000)
            o2 = student(bundle_from(r2))
            print("   mid-training eval mse", float(((o2["rgb"] - g2) ** 2).mean()))
        student.train()
    bad = not bool(torch.isfinite(l).all()) or float(l[3]) > 1.0
    if it % 50 == 0 or bad or (os.environ.get('EVAL_AT') and 0 <= it - int(os.environ['EVAL_AT']) < 6):
        print(mode, it, [round(float(v), 6) for v in l], "upd" if eng._steps_since_update == 1 else "")
    if bad:
        for nm in ("rgb_out", "acc", "rgb", "hin", "d_rgb", "d_hin"):
            t = getattr(eng, nm); print("  ", nm, "finite" if torch.isfinite(t).all() else "NONFINITE", float(t.abs().max()))
        for i in range(3):
            for nm in ("dens", "w", "enc", "h", "d_w", "d_dens", "eb", "sb"):
                t = getattr(eng, nm)[i]; print("  ", nm, i, "finite" if torch.isfinite(t).all() else "NONFINITE", float(t.abs().max()) if t.numel() else 0)
        print("   params finite", bool(torch.isfinite(eng.optim.flat).all()), "grad finite", bool(torch.isfinite(eng.optim.flat_grad).all()), float(eng.optim.flat.abs().max()))
        break
# ---- evaluation through the module path after engine training
student.eval()
rays, gt = batch(10_000)
with torch.no_grad():
    for nm, near in (("eval", None),):
        out = student(bundle_from(rays))
        print("EVAL rgb mean", float(out["rgb"].mean()), "acc mean", float(out["accumulation"].mean()), "gt mean", float(gt.mean()),
              "mse", float(((out["rgb"] - gt) ** 2).mean()), "finite", bool(torch.isfinite(out["rgb"]).all()))
    student.train()
    out = student(bundle_from(rays))
    print("TRAIN-mode fwd rgb mean", float(out["rgb"].mean()), "acc", float(out["accumulation"].mean()), "mse", float(((out["rgb"] - gt) ** 2).mean()))
    for i, p in enumerate(student.proposal_networks):
        print("prop", i, "table absmax", float(p.encoding.hash_table.abs().max()), "finite", bool(torch.isfinite(p.encoding.hash_table).all()))
    f = student.field
    print("field table absmax", float(f.mlp_base.model[0].hash_table.abs().max()), [float(l.weight.abs().max()) for l in f.mlp_head.layers])
