#!/usr/bin/env python
"""bench.py — train rays/sec of the nerfacto hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 30 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

One "step" = one full optimisation step of nerfacto on one batch of 4096 synthetic rays per GPU
(BASELINE configs[2]: proposal sampler 256 -> 96 -> 48 samples, L=16/T=2^19/F=2 main grid, two L=5/T=2^17
proposal grids, scene contraction, appearance embedding): ray batch -> proposal sampling (2 density networks)
-> main field -> weights -> compositing -> rgb + interlevel + distortion losses -> backward -> (allreduce) ->
fused Adam.  Nothing is skipped or cached inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

RAYS_PER_GPU = 4096
NUM_IMAGES = 200
METRIC = "train_rays_per_sec"
WORKLOAD = "nerfacto 4096 rays/GPU x (256,96)->48 samples, L16/T2^19/F2 grid + 2x(L5/T2^17) proposal grids"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def oracle_step_factory(n_rays: int, seed: int = 0):
    """The reference's CPU path for the same step, restated in oracle/nerf_oracle.py (torch CPU, fp32, all host
    threads): forward + losses + backward + Adam on `n_rays` rays of the same workload."""
    from oracle import nerf_oracle as O
    from nerfstudio_b200.scene import synthetic_rays

    torch.manual_seed(seed)

    def table(L, log2T):
        return ((torch.rand(L << log2T, 2) * 2 - 1) * 1e-3).requires_grad_(True)

    def lin(o, i):
        l = torch.nn.Linear(i, o)
        return l.weight.detach().clone().requires_grad_(True), l.bias.detach().clone().requires_grad_(True)

    props = []
    for max_res in (128, 256):
        w0, b0 = lin(16, 10)
        w1, b1 = lin(1, 16)
        props.append(dict(table=table(5, 17), scalings=O.hash_level_scalings(5, 16, max_res), log2_T=17, w=[w0, w1], b=[b0, b1]))
    wb0, bb0 = lin(64, 32)
    wb1, bb1 = lin(16, 64)
    wh0, bh0 = lin(64, 63)
    wh1, bh1 = lin(64, 64)
    wh2, bh2 = lin(3, 64)
    field = dict(table=table(16, 19), scalings=O.hash_level_scalings(16, 16, 2048), log2_T=19,
                 embedding=torch.randn(NUM_IMAGES, 32).requires_grad_(True), w_base=[wb0, wb1], b_base=[bb0, bb1],
                 w_head=[wh0, wh1, wh2], b_head=[bh0, bh1, bh2])
    P = dict(props=props, field=field)
    leaves = []
    for p in props:
        leaves += [p["table"]] + p["w"] + p["b"]
    leaves += [field["table"], field["embedding"]] + field["w_base"] + field["b_base"] + field["w_head"] + field["b_head"]
    state = [(torch.zeros_like(l), torch.zeros_like(l)) for l in leaves]
    rays, gt = synthetic_rays(n_rays, NUM_IMAGES, seed)
    orays = dict(origins=rays["origins"], directions=rays["directions"], nears=torch.full((n_rays, 1), 0.05),
                 fars=torch.full((n_rays, 1), 1000.0), camera_indices=rays["camera_indices"][:, 0], rgb=gt)
    cfg = dict(num_prop_samples=(256, 96), num_nerf_samples=48, aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]]),
               contraction=True, avg_init=0.01, anneal=1.0)
    counter = {"step": 0}

    def step():
        counter["step"] += 1
        rng = dict(jitter0=torch.rand(n_rays, 1), jitter_pdf=[torch.rand(n_rays, 1), torch.rand(n_rays, 1)])
        out = O.nerfacto_forward(P, orays, cfg, rng, training=True)
        grads = torch.autograd.grad(out["loss"], leaves)
        with torch.no_grad():
            for l, g, (m, v) in zip(leaves, grads, state):
                O.adam_step(l, g, m, v, counter["step"], 1e-2)
        return float(out["loss"].detach())

    return step


def pick_cpu_threads():
    """The thread count at which the CPU port runs fastest on this host (torch's intra-op parallelism stops scaling —
    and on a 128-thread box reverses — well before all hardware threads are used): one 256-ray step per candidate."""
    cores = os.cpu_count() or 1
    best = (0.0, 1)
    for c in sorted({c for c in (8, 16, 32, 64, cores) if c <= cores}):
        torch.set_num_threads(c)
        step = oracle_step_factory(256)
        step()
        t0 = time.perf_counter()
        step()
        rps = 256 / (time.perf_counter() - t0)
        if rps > best[0]:
            best = (rps, c)
    return best[1], best[0]


def time_cpu(n_rays: int, steps: int, warmup: int, threads: int):
    torch.set_num_threads(threads)
    step = oracle_step_factory(n_rays)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_rays * steps / dt, dt / steps, threads


def run_reference(args) -> None:
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference is pure
    Python/PyTorch so there is nothing to compile into oracle/_ref) on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads, rps_est = pick_cpu_threads()
    # rays per step: the full 4096-ray batch if steps + warmup of it fit in ~150 s at the calibrated rate, else a
    # bounded sample of the same workload (multiple of 64 rays, at least 64)
    budget_rays = rps_est * 150.0 / max(args.steps + args.warmup, 1)
    n_rays = int(min(RAYS_PER_GPU, max(64, (int(budget_rays) // 64) * 64)))
    rps, sec, cores = time_cpu(n_rays, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": rps, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{n_rays} rays/step of the 4096-ray batch (CPU bounded sample)"},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {n_rays} rays, fwd+bwd+Adam, torch CPU fp32, {cores} threads "
                                   f"(fastest of 8/16/32/64/{os.cpu_count()} on this host)"},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    from nerfstudio_b200 import distributed as D
    from nerfstudio_b200 import functional as F
    from nerfstudio_b200 import lib
    from nerfstudio_b200.nerfacto import NerfactoModel, NerfactoModelConfig, Trainer
    from nerfstudio_b200.scene import bundle_from, synthetic_rays

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 core has no CPU fallback")
    rank, local, world = D.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib.load()
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        assert lib.tune(k, int(v)), f"unknown tuning key {k}"
    torch.manual_seed(0)  # identical initial weights on every rank
    cfg = NerfactoModelConfig(implementation="torch", average_init_density=0.01)
    model = NerfactoModel(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_train_data=NUM_IMAGES).to(dev)
    if args.force_proposal_update:
        model.proposal_sampler.update_sched = lambda step: -1  # proposal networks trained on every step
    allreduce = D.FlatGradAllReduce() if world > 1 else None
    if args.engine == "autograd":
        trainer = Trainer(model, allreduce=allreduce)
        engine = None
    else:
        from nerfstudio_b200.engine import NerfactoStep

        engine = NerfactoStep(model, RAYS_PER_GPU, allreduce=allreduce, use_graph=(args.engine == "graph"),
                              always_update_proposals=args.force_proposal_update, mlp_backend=args.mlp,
                              fused_proposals=not args.unfused_proposals)
        trainer = engine
    D.broadcast_parameters(trainer.optim.flat)
    n_params = trainer.optim.flat.numel()

    # independent ray draws per rank (scripts/train.py:98: seed + rank)
    n_batches = 4
    host = []
    for b in range(n_batches):
        rays, gt = synthetic_rays(RAYS_PER_GPU, NUM_IMAGES, seed=1000 * rank + b)
        host.append(({k: v.pin_memory() for k, v in rays.items()}, gt.pin_memory()))
    resident = [({k: v.to(dev) for k, v in r.items()}, g.to(dev)) for r, g in host]
    torch.manual_seed(42 + rank)

    packed = ([engine.pack_batch(r["origins"], r["directions"], r["camera_indices"], g) for r, g in host]
              if engine is not None else None)  # pinned, laid out by the engine's loader-side helper
    packed_dev = [b.to(dev) for b in packed] if packed is not None else None

    def step_resident(i):
        rays, gt = resident[i % n_batches]
        if engine is not None:
            engine.set_batch_packed(packed_dev[i % n_batches])  # device -> device, batch already in HBM
            return engine.step()
        return trainer.train_iteration(bundle_from(rays), {"image": gt})

    def step_e2e(i):
        rays, gt = host[i % n_batches]  # pinned host memory -> device inside the timed region
        if engine is not None:
            engine.set_batch_packed(packed[i % n_batches])  # one async H2D copy of the whole batch
            return float(engine.step()[3].item())  # device -> host read of the step's loss
        d_rays = {k: v.to(dev, non_blocking=True) for k, v in rays.items()}
        stats = trainer.train_iteration(bundle_from(d_rays), {"image": gt.to(dev, non_blocking=True)})
        return float(stats["loss"].item())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item()), clocks

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    lib.LAUNCHES = 0
    F.KERNEL_TIMES.clear()
    F.PROFILE_KERNELS = engine is None
    ms, clocks = timed(step_resident, args.steps, ClockSampler(local))
    F.PROFILE_KERNELS = False
    launches = lib.LAUNCHES
    torch.cuda.synchronize()
    value = world * RAYS_PER_GPU * args.steps / (ms * 1e-3)
    kt = F.kernel_time_summary()
    if engine is not None:
        # per-kernel durations: the same launch sequence run eagerly with CUDA events around every C-ABI call
        # (a CUDA graph has no per-node events); `launches` = our kernel launches replayed per graph x steps
        was_graph, engine.use_graph = engine.use_graph, False
        lib.LAUNCHES, lib.PROFILE = 0, None
        step_resident(0)
        launches = lib.LAUNCHES * args.steps
        lib.PROFILE = {}
        n_prof = 5
        for i in range(n_prof):
            step_resident(i)
        torch.cuda.synchronize()
        prof, lib.PROFILE = lib.profile_summary(), None
        engine.use_graph = was_graph
        S = engine.S
        # algorithmic bytes of every kernel that gathers from / scatters into a hash table: 8 corners x F(2) x 4 B per
        # (point, level); scatter counted as read + write.  The fused density-field kernels carry the proposal grids.
        hb = {}
        for i, L in ((0, 5), (1, 5), (2, 16)):
            n_pts, b = RAYS_PER_GPU * S[i], RAYS_PER_GPU * S[i] * L * 8 * 2 * 4
            hb[f"b2n_hashgrid_fwd[n={n_pts}]"], hb[f"b2n_hashgrid_bwd[n={n_pts}]"] = b, 2 * b
            hb[f"b2n_density_field_fwd[n={n_pts}]"] = b
            hb[f"b2n_density_field_bwd[n={n_pts}]"] = 3 * b  # re-gather + scatter
        kt = {k: {"launches": c, "ms_total": t, "ms_avg": t / c, "bytes_per_launch": hb[k]} for k, (c, t) in prof.items() if k in hb}
        kernel_table = {k: round(t / n_prof, 4) for k, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
        ms_prof = sum(t for _, t in prof.values()) / n_prof
    else:
        kernel_table, ms_prof, n_prof = None, ms / args.steps, args.steps

    for i in range(3):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, args.steps)
    e2e_value = world * RAYS_PER_GPU * args.steps / (ms_e2e * 1e-3)
    if packed is not None:
        h2d = packed[0].numel() * packed[0].element_size()
    else:
        h2d = sum(v.numel() * v.element_size() for v in host[0][0].values()) + host[0][1].numel() * 4

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (hash-grid gather / scatter), algorithmic bytes: 8 corners x F x 4 B
    peak, peak_src = peaks()
    dom = max(kt.items(), key=lambda kv: kv[1]["ms_total"]) if kt else None
    roofline = None
    if dom is not None:
        name, st = dom
        ach = st["bytes_per_launch"] / (st["ms_avg"] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_dram_traffic.json")  # dram read+write per launch, ncu --set full
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(name.split("[")[0], {}).get(name)
        roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src, "launches_per_step": st["launches"] / n_prof,
                    "ms_avg": st["ms_avg"], "share_of_step": st["ms_avg"] * st["launches"] / n_prof / ms_prof,
                    "all_hash_kernels": {k: {"ms_avg": v["ms_avg"], "GBps": v["bytes_per_launch"] / (v["ms_avg"] * 1e-3) / 1e9,
                                             "launches_per_step": v["launches"] / n_prof} for k, v in kt.items()},
                    "kernel_ms_per_step": kernel_table}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads, _ = pick_cpu_threads()
        rps, sec, cores = time_cpu(512, 3, 1, threads)
        cpu = {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
               "sample": f"3 steps x 512 rays of the same workload (fwd+bwd+Adam), torch CPU fp32, {cores} threads"}
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch_rays": world * RAYS_PER_GPU,
                   "parallelism": f"ray-batch data parallel x{world}, one flat-gradient allreduce/step" if world > 1 else "single GPU",
                   "precision": ("fp32 tables; MLPs on tcgen05 tensor cores, 3xTF32 split, fp32 accumulate in TMEM (1e-4 parity mode)"
                                 if (engine is not None and engine.tc) else "fp32 tables, fp32 SIMT MLPs (1e-4 parity mode)"), "optimizer": "fused Adam over one flat buffer", "engine": args.engine,
                   "proposal_update": "every step" if args.force_proposal_update else "reference schedule",
                   "l2": f"per-step working set {4 * 4 * n_params / 1e6:.0f} MB (params+grads+Adam moments) > 126 MB L2",
                   "params": n_params},
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mlp", default="auto", choices=["auto", "tc", "simt"],
                    help="tiny-MLP kernels of the graph/eager engine: tcgen05 3xTF32 (tc) or fp32 SIMT")
    ap.add_argument("--unfused-proposals", action="store_true", help="proposal networks as separate grid/MLP launches")
    ap.add_argument("--tune", default="", help="comma separated key=value launch-geometry knobs (lib.tune)")
    ap.add_argument("--engine", default="graph", choices=["graph", "eager", "autograd"],
                    help="graph: CUDA-graph replay of the hand-written step (default); eager: same launches without a "
                         "graph; autograd: the drop-in modules under torch.autograd")
    ap.add_argument("--reference-schedule", dest="force_proposal_update", action="store_false",
                    help="use nerfacto's proposal-update schedule instead of training the proposal nets every step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
