"""Multi-GPU consistency of the captured step (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multi_gpu_check.py

Trains the BASELINE nerfacto configuration for a few steps in both gradient-exchange modes — (a) reduce-scatter ->
sharded Adam -> all-gather, (b) all-reduce + replicated Adam — from the same weights on the same rays and checks:
replicas hold identical parameters after flush() in both modes, the two modes' loss curves agree (float atomics only),
and reports the step time of each.  Prints one JSON line on rank 0.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from nerfstudio_b200 import distributed as D
from nerfstudio_b200.cameras.camera_optimizers import CameraOptimizerConfig
from nerfstudio_b200.engine import NerfactoStep
from nerfstudio_b200.nerfacto import NerfactoModel, NerfactoModelConfig
from nerfstudio_b200.scene import synthetic_rays

R, STEPS = 4096, int(os.environ.get("STEPS", "40"))
rank, local, world = D.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
out = {"world": world, "steps": STEPS}
curves = {}
for mode in ("sharded", "allreduce"):
    torch.manual_seed(0)
    cfg = NerfactoModelConfig(implementation="torch", average_init_density=0.01, camera_optimizer=CameraOptimizerConfig(mode="SO3xR3"))
    model = NerfactoModel(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_train_data=16).to(dev)
    model.proposal_sampler.update_sched = lambda step: -1
    eng = NerfactoStep(model, R, allreduce=D.FlatGradAllReduce(), use_graph=True, always_update_proposals=True,
                       sharded_update=(mode == "sharded"))
    D.broadcast_parameters(eng.optim.flat)
    torch.manual_seed(42 + rank)
    batches = []
    for b in range(4):
        rays, gt = synthetic_rays(R, 16, seed=1000 * rank + b)
        batches.append(({k: v.to(dev) for k, v in rays.items()}, gt.to(dev)))
    losses = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(STEPS):
        if it == STEPS // 2:
            torch.cuda.synchronize(), dist.barrier(), torch.cuda.synchronize()
            t0.record()
        rays, gt = batches[it % 4]
        eng.set_batch(rays["origins"], rays["directions"], rays["camera_indices"], gt)
        l = eng.step()
        losses.append(l[3].clone())
    t1.record()
    eng.flush()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / (STEPS - STEPS // 2)
    flat = eng.optim.flat
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    diff = (flat - ref).abs().max()
    dist.all_reduce(diff, op=dist.ReduceOp.MAX)
    curve = torch.stack(losses)
    dist.all_reduce(curve, op=dist.ReduceOp.SUM)
    curves[mode] = (curve / world).cpu()
    tms = torch.tensor([ms], device=dev)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    out[mode] = {"ms_per_step": float(tms), "max_param_diff_across_ranks": float(diff), "first_loss": float(curves[mode][0]),
                 "last_loss": float(curves[mode][-1]), "finite": bool(torch.isfinite(flat).all())}
    del eng, model
rel = ((curves["sharded"] - curves["allreduce"]).abs() / curves["allreduce"].abs().clamp_min(1e-12))
out["loss_curve_max_rel_diff_first_10_steps"] = float(rel[:10].max())
out["loss_curve_rel_diff_last_step"] = float(rel[-1])
ok = (out["sharded"]["max_param_diff_across_ranks"] == 0.0 and out["allreduce"]["max_param_diff_across_ranks"] == 0.0
      and out["loss_curve_max_rel_diff_first_10_steps"] < 1e-3 and out["sharded"]["finite"]
      and out["sharded"]["last_loss"] < out["sharded"]["first_loss"])
out["ok"] = bool(ok)
if rank == 0:
    print(json.dumps(out))
dist.barrier()
sys.exit(0 if ok else 1)
