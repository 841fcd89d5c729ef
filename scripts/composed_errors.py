"""Actual end-to-end errors of the UNSTAGED composed pipeline (engine step and autograd modules) against the golden
nerfacto_pipeline fixture — the numbers behind the tolerances of tests/test_gpu_engine.py / test_gpu_modules.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import rel_err, load_golden
import test_gpu_engine as TE
from test_gpu_modules import _named_params

g = load_golden("nerfacto_pipeline")
model, eng = TE._mk(g, use_graph=False)
model.proposal_sampler.set_anneal(0.7)
eng._anneal = lambda step: 0.7
eng.optim.lr = 0.0
losses = eng.step().cpu()
for i, k in enumerate(["loss_rgb", "loss_interlevel", "loss_distortion", "loss"]):
    print("engine %-18s rel %.3e" % (k, rel_err(losses[i], g[k])))
for i in range(3):
    print("engine sbins%d rel %.3e   w%d rel %.3e" % (i, rel_err(eng.sb[i].cpu(), g[f"train_sbins{i}"].reshape(eng.sb[i].shape)), i,
          rel_err(eng.w[i].cpu().reshape(-1), g[f"train_w{i}"].reshape(-1))))
worst = {}
for k, p in _named_params(model).items():
    a, b = p.grad.detach().cpu().double().flatten(), g["g_" + k].double().flatten()
    cos = float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))
    print("engine g_%-28s entry rel %.3e  cos %.7f  norm ratio %.6f" % (k, rel_err(p.grad.cpu(), g["g_" + k]), cos, float(a.norm() / b.norm())))
dm = (eng.depth_med.cpu() - g["train_depth"].reshape(-1)).abs() <= 1e-6 * g["train_depth"].reshape(-1).abs()
print("engine median depth identical on %.4f of rays" % dm.float().mean().item())
