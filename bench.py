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
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

RAYS_PER_GPU = 4096
NUM_IMAGES = 200
METRIC = "train_rays_per_sec"
CAMERA_OPTIMIZER = "SO3xR3"  # set from --camera-optimizer; nerfacto's default (models/nerfacto.py:131)
WORKLOAD = "nerfacto 4096 rays/GPU x (256,96)->48 samples, L16/T2^19/F2 grid + 2x(L5/T2^17) proposal grids"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of this rank's GPU, sampled through NVML from a thread of this process.

    The thread is started long before the timed region (nothing is forked or initialised between the barrier and the
    first timed launch — round 1 spawned nvidia-smi there and, on an 8-GPU node, lost ~85 ms of the window to NVML
    start-up).  Samples carry a host timestamp; `window(t0, t1)` summarises those that fall inside a timed region."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, device_index: int, period_s: float = 0.004):
        self.rows, self.period, self.err = [], period_s, None
        self._stop = threading.Event()
        self.h = self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:  # noqa: BLE001
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[device_index]) if vis and vis.split(",")[device_index].isdigit() else device_index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._sample()  # first NVML queries (slow path) happen here, not in the timed region
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as e:  # noqa: BLE001
            self.err = f"NVML unavailable: {e}"

    def _sample(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:  # noqa: BLE001
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.rows.append((time.perf_counter(), mhz, mask))

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                return
            time.sleep(self.period)

    def window(self, spans) -> dict:
        """spans: list of (t0, t1) host times of the timed regions."""
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [self.err or "NVML unavailable"]}
        rows = [r for r in self.rows if any(t0 <= r[0] <= t1 for t0, t1 in spans)]
        sm = sorted(r[1] for r in rows)
        reasons = sorted({name for _, _, m in rows for name, bit in self.REASONS if m & bit})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "reasons": reasons, "source": "NVML, in-process thread started before warm-up"}

    def stop(self):
        self._stop.set()


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


def reference_step_factory(n_rays: int, seed: int = 0):
    """The REFERENCE ITSELF on the host cores: the unmodified `nerfstudio.models.nerfacto.NerfactoModel` (torch
    implementation, its own samplers / fields / renderers / losses, torch.optim.Adam as configs/method_configs.py:106-113
    configures it) imported from /root/reference or oracle/_ref (oracle/make_ref.py).  None when it is not importable."""
    from oracle import ref_loader

    if ref_loader.load() is None:
        return None
    import nerfstudio.models.nerfacto as NM
    from nerfstudio.cameras.rays import RayBundle
    from nerfstudio.data.scene_box import SceneBox
    from nerfstudio_b200.scene import synthetic_rays

    torch.manual_seed(seed)
    cfg = NM.NerfactoModelConfig(implementation="torch", average_init_density=0.01)
    cfg.camera_optimizer.mode = CAMERA_OPTIMIZER  # nerfacto's default is SO3xR3 (models/nerfacto.py:131)
    model = NM.NerfactoModel(cfg, scene_box=SceneBox(aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]])),
                             num_train_data=NUM_IMAGES).train()
    model.proposal_sampler.update_sched = lambda step: -1  # proposal networks trained every step, as the GPU arm does
    groups = model.get_param_groups()
    opts = [torch.optim.Adam(groups[k], lr=1e-3 if k == "camera_opt" else 1e-2, eps=1e-15) for k in groups]
    rays, gt = synthetic_rays(n_rays, NUM_IMAGES, seed)
    counter = {"step": 0}

    def step():
        rb = RayBundle(origins=rays["origins"], directions=rays["directions"], pixel_area=rays["pixel_area"],
                       camera_indices=rays["camera_indices"])
        for o in opts:
            o.zero_grad(set_to_none=True)
        out = model(rb)
        batch = {"image": gt}
        md = model.get_metrics_dict(out, batch)
        loss = sum(model.get_loss_dict(out, batch, md).values())
        loss.backward()
        for o in opts:
            o.step()
        model.proposal_sampler.step_cb(counter["step"])
        counter["step"] += 1
        return float(loss.detach())

    return step


def cpu_step_factory(n_rays: int):
    """(step fn, kind): the reference itself when importable, else the oracle's restatement of it."""
    try:
        step = reference_step_factory(n_rays)
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"[bench] reference not runnable ({type(e).__name__}: {e}); using the oracle port\n")
        step = None
    if step is not None:
        return step, "reference"
    return oracle_step_factory(n_rays), "port"


def pick_cpu_threads():
    """The thread count at which the CPU port runs fastest on this host (torch's intra-op parallelism stops scaling —
    and on a 128-thread box reverses — well before all hardware threads are used): one 256-ray step per candidate."""
    cores = os.cpu_count() or 1
    best = (0.0, 1)
    for c in sorted({c for c in (8, 16, 32, 64, cores) if c <= cores}):
        torch.set_num_threads(c)
        step, _ = cpu_step_factory(256)
        step()
        t0 = time.perf_counter()
        step()
        rps = 256 / (time.perf_counter() - t0)
        if rps > best[0]:
            best = (rps, c)
    return best[1], best[0]


def time_cpu(n_rays: int, steps: int, warmup: int, threads: int):
    torch.set_num_threads(threads)
    step, kind = cpu_step_factory(n_rays)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_rays * steps / dt, dt / steps, threads, kind


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
    rps, sec, cores, kind = time_cpu(n_rays, args.steps, args.warmup, threads)
    what = ("the unmodified reference NerfactoModel (torch path, torch.optim.Adam)" if kind == "reference"
            else "oracle port of the reference's torch path")
    line = {
        "impl": "reference", "metric": METRIC, "value": rps, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{n_rays} rays/step of the 4096-ray batch (CPU bounded sample)"},
        "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} steps x {n_rays} rays, fwd+bwd+Adam, {what}, torch CPU fp32, {cores} threads "
                                   f"(fastest of 8/16/32/64/{os.cpu_count()} on this host)"},
        "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
STATE_STEP = 100  # the timed windows start at this optimisation step (fixed training state: gradient sparsity of the
#                   proposal levels, hence density_field_bwd's time, depends on how far training has progressed)

# algorithmic flops of the tiny MLPs: 2 * sum(in_i * out_i) per row forward; backward = dA + dW = 2x
_MLP_FLOPS = {(32, 16): 2 * (32 * 64 + 64 * 16), (64, 3): 2 * (63 * 64 + 64 * 64 + 64 * 3), (63, 3): 2 * (63 * 64 + 64 * 64 + 64 * 3)}


def sustained_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured sustained bf16 (MEASURED_PEAKS.json)"
    return 1500.0, "fallback (B200_PROFILING.md)"


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
    sampler = ClockSampler(local)  # NVML thread: started now, long before any timed region
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        assert lib.tune(k, int(v)), f"unknown tuning key {k}"
    torch.manual_seed(0)  # identical initial weights on every rank
    from nerfstudio_b200.cameras.camera_optimizers import CameraOptimizerConfig

    cfg = NerfactoModelConfig(implementation="torch", average_init_density=0.01,
                              camera_optimizer=CameraOptimizerConfig(mode=CAMERA_OPTIMIZER))
    model = NerfactoModel(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_train_data=NUM_IMAGES).to(dev)
    if args.force_proposal_update:
        model.proposal_sampler.update_sched = lambda step: -1  # proposal networks trained on every step
    allreduce = D.FlatGradAllReduce() if world > 1 else None

    # independent ray draws per rank (scripts/train.py:98: seed + rank)
    n_batches = 4
    host = []
    for b in range(n_batches):
        rays, gt = synthetic_rays(RAYS_PER_GPU, NUM_IMAGES, seed=1000 * rank + b)
        host.append(({k: v.pin_memory() for k, v in rays.items()}, gt.pin_memory()))
    resident = [({k: v.to(dev) for k, v in r.items()}, g.to(dev)) for r, g in host]

    fused = None
    if args.engine == "autograd":
        trainer = Trainer(model, allreduce=allreduce)
        engine = None
    else:
        # the captured step behind the reference's training surface (Trainer.train_iteration /
        # Pipeline.get_train_loss_dict): nerfstudio_b200/pipeline.py
        from nerfstudio_b200.pipeline import FusedTrainStep, HostRayQueue

        fused = FusedTrainStep(model, None, RAYS_PER_GPU, allreduce=allreduce, use_graph=(args.engine == "graph"),
                               always_update_proposals=args.force_proposal_update, mlp_backend=args.mlp,
                               fused_proposals=not args.unfused_proposals, sharded_update=not args.no_sharded_update)
        engine = trainer = fused.engine
        fused.datamanager = HostRayQueue(engine, host)
    D.broadcast_parameters(trainer.optim.flat)
    n_params = trainer.optim.flat.numel()
    torch.manual_seed(42 + rank)
    packed_dev = [it[2].to(dev) for it in fused.datamanager.items] if fused is not None else None

    def step_resident(i):
        rays, gt = resident[i % n_batches]
        if engine is not None:
            engine.set_batch_packed(packed_dev[i % n_batches])  # device -> device, batch already in HBM
            return engine.step()
        return trainer.train_iteration(bundle_from(rays), {"image": gt})

    def step_e2e(i):
        if fused is not None:
            # THE plugin call: Trainer.train_iteration(step) -> pinned host RayBundle from the data manager, H2D copy,
            # one graph replay (fwd + losses + bwd + allreduce + Adam), then a device -> host read of the loss
            loss, _loss_dict, _metrics = fused.train_iteration(i)
            return float(loss)
        rays, gt = host[i % n_batches]  # pinned host memory -> device inside the timed region
        d_rays = {k: v.to(dev, non_blocking=True) for k, v in rays.items()}
        stats = trainer.train_iteration(bundle_from(d_rays), {"image": gt.to(dev, non_blocking=True)})
        return float(stats["loss"].item())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    spans = []

    def timed(fn, steps, windows, first_step):
        """`windows` back-to-back timed regions of exactly `steps` steps each; every region is bracketed by
        barrier + synchronize on both sides, timed with CUDA events, max over ranks.  Returns the per-window ms."""
        out = []
        for w in range(windows):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            t0 = time.perf_counter()
            e0.record()
            for i in range(steps):
                fn(first_step + w * steps + i)
            e1.record()
            barrier()
            spans.append((t0, time.perf_counter()))
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
            out.append(float(ms.item()))
        return out

    # warm-up: at least the requested steps, and always up to the fixed training state STATE_STEP (also warms NCCL's
    # channels at N > 1: the driver asks for only 5 warm-up steps)
    n_warm = max(args.warmup, 3, args.state_step)
    for i in range(n_warm):
        step_resident(i)
    lib.LAUNCHES = 0
    F.KERNEL_TIMES.clear()
    F.PROFILE_KERNELS = engine is None
    win_ms = timed(step_resident, args.steps, args.windows, n_warm)
    F.PROFILE_KERNELS = False
    launches = lib.LAUNCHES
    ms = sorted(win_ms)[len(win_ms) // 2]  # median window
    value = world * RAYS_PER_GPU * args.steps / (ms * 1e-3)
    clocks = sampler.window(list(spans))

    for i in range(3):
        step_e2e(i)
    e2e_ms = timed(step_e2e, args.steps, args.windows, 3)
    ms_e2e = sorted(e2e_ms)[len(e2e_ms) // 2]
    e2e_value = world * RAYS_PER_GPU * args.steps / (ms_e2e * 1e-3)
    if fused is not None:
        h2d = fused.datamanager.bytes_per_batch
    else:
        h2d = sum(v.numel() * v.element_size() for v in host[0][0].values()) + host[0][1].numel() * 4

    # ---- per-kernel durations (after the timed regions): the same launch sequence run eagerly with CUDA events around
    # every C-ABI call on the launching stream (a CUDA graph has no per-node events)
    kernel_table, ms_prof, n_prof, prof, scatter_frac, dense_ms = None, ms / args.steps, args.steps, {}, {}, {}
    if engine is not None:
        # (branches serialised for this pass: a kernel timed while another runs beside it would be charged the overlap)
        was_graph, engine.use_graph = engine.use_graph, False
        was_conc, engine.concurrent = engine.concurrent, False
        lib.LAUNCHES, lib.PROFILE = 0, None
        step_resident(0)
        launches_per_step = lib.LAUNCHES
        launches = launches_per_step * args.steps * args.windows
        lib.PROFILE = {}
        n_prof = 10
        for i in range(n_prof):
            step_resident(i)
        torch.cuda.synchronize()
        prof, lib.PROFILE = lib.profile_summary(), None
        engine.use_graph, engine.concurrent = was_graph, was_conc
        kernel_table = {k: round(t / n_prof, 4) for k, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
        ms_prof = sum(t for _, t in prof.values()) / n_prof
        # measured fraction of samples whose density gradient is non-zero: density_field_bwd skips the scatter for the
        # others, so its algorithmic bytes are gather + frac * (read + write)
        for lvl in (0, 1):
            scatter_frac[RAYS_PER_GPU * engine.S[lvl]] = float((engine.d_dens[lvl] != 0).float().mean().item())
        # the same kernel with a gradient on EVERY sample (the regime at the start of training / on scenes where the
        # proposal histograms under-estimate everywhere): its time is data dependent, so both regimes are reported
        if engine.fused_props:
            for lvl in (0, 1):
                engine.d_dens[lvl].fill_(1e-6)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                engine.density_field_bwd_launch(lvl)
                e0.record()
                for _ in range(5):
                    engine.density_field_bwd_launch(lvl)
                e1.record()
                torch.cuda.synchronize()
                dense_ms[f"b2n_density_field_bwd[n={RAYS_PER_GPU * engine.S[lvl]}]"] = e0.elapsed_time(e1) / 5
    # ---- full-image eval path: 1920x1080 frame through the graph-captured chunk loop, rays generated on the
    # device, chunks sharded over the ranks; ms per frame with CUDA events (median of 3), max over ranks
    eval_line = None
    if engine is not None and not args.no_eval:
        from nerfstudio_b200.cameras.cameras import Cameras
        from nerfstudio_b200.render_engine import NerfactoRender

        Hh, Ww = 1080, 1920
        c2w = torch.eye(4, device=dev)[:3][None].clone()
        c2w[0, :, 3] = torch.tensor([0.0, 0.0, 1.2], device=dev)
        cams = Cameras(c2w, torch.tensor([1200.0], device=dev), torch.tensor([1200.0], device=dev),
                       torch.tensor([Ww / 2], device=dev), torch.tensor([Hh / 2], device=dev),
                       width=torch.tensor([Ww], device=dev), height=torch.tensor([Hh], device=dev))
        engine.flush()  # multi-GPU sharded update: the last step's parameter all-gather must have landed before the model is read
        model.eval()
        ren = NerfactoRender(model, chunk_rays=1 << 15, mlp_backend=args.mlp)
        ren.render_camera(cams, 0, shard=world > 1)  # warm-up + graph capture
        frame_ms = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            img = ren.render_camera(cams, 0, shard=world > 1)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            frame_ms.append(float(t.item()))
        model.train()
        fm = sorted(frame_ms)[1]
        eval_line = {"frame": "1920x1080, one perspective camera, 32768-ray chunks (eval_num_rays_per_chunk)", "ms_per_frame": fm,
                     "fps": 1e3 / fm, "rays_per_sec": Hh * Ww / (fm * 1e-3), "frames_ms": frame_ms,
                     "finite": bool(torch.isfinite(img["rgb"]).all().item()),
                     "through": "render_engine.NerfactoRender.render_camera (Model.get_outputs_for_camera surface)"}
    sampler.stop()
    if rank != 0:
        return

    # ---- rooflines.  (1) the step's DOMINANT kernel (largest time per step); (2) the hash gather on the 16-level grid —
    # the kernel north_star's ">= 60 % of the HBM roofline" names; (3) every hash kernel, for the record.
    hbm_peak, peak_src = peaks()
    # per-launch DRAM (read+write) and L2 (lts__t_bytes) bytes of the same kernels from the committed `ncu --set full`
    # capture of this build (scripts/ncu_traffic.py -> profiles/ncu_traffic.json: {call key: {dram_bytes, lts_bytes, ...}})
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    ncu = json.load(open(tpath)) if os.path.exists(tpath) else {}

    def ncu_of(key, field):
        return (ncu.get(key) or {}).get(field)

    roofline = gather_roof = None
    hash_rows = {}
    if prof:
        S = engine.S
        hb = {}
        for i, L in ((0, 5), (1, 5), (2, 16)):
            n_pts = RAYS_PER_GPU * S[i]
            b = n_pts * L * 8 * 2 * 4  # 8 corners x F(2) x 4 B per (point, level)
            hb[f"b2n_hashgrid_fwd[n={n_pts}]"], hb[f"b2n_hashgrid_bwd[n={n_pts}]"] = b, 2 * b
            hb[f"b2n_density_field_fwd[n={n_pts}]"] = b
            hb[f"b2n_density_field_bwd[n={n_pts}]"] = int(b * (1 + 2 * scatter_frac.get(n_pts, 1.0)))
        for k, (c, t) in prof.items():
            if k in hb:
                hash_rows[k] = {"ms_avg": t / c, "GBps": hb[k] / (t / c * 1e-3) / 1e9, "frac_of_hbm_peak": hb[k] / (t / c * 1e-3) / 1e9 / hbm_peak,
                                "bytes_per_launch": hb[k], "launches_per_step": c / n_prof}
        name, (c, t) = max(prof.items(), key=lambda kv: kv[1][1])
        ms_avg = t / c
        share = t / n_prof / ms_prof
        if name.startswith("b2n_mlp_tc") and "n=" in name and "in=" in name:  # (b2n_mlp_tc_pack carries no shape)
            tpeak, tsrc = sustained_tensor_peak()
            n_rows = int(name.split("n=")[1].split(",")[0])
            key = (int(name.split("in=")[1].split(",")[0]), int(name.split("out=")[1].split("]")[0]))
            fl = _MLP_FLOPS.get(key, 0) * n_rows * (2 if "bwd" in name else 1)
            ach = fl / (ms_avg * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": name, "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                        "traffic": ncu_of(name, "dram_bytes"), "peak_source": tsrc, "flops_per_launch": fl,
                        "note": "algorithmic fp32 flops of the tiny MLP (3xTF32 issues 3 tensor-core passes per product); "
                                "K, N <= 64 GEMMs are latency-chained per 128-row tile, never near the tensor peak"}
        else:
            by = hb.get(name)
            ach = by / (ms_avg * 1e-3) / 1e9 if by else None
            roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak if ach else None, "traffic": ncu_of(name, "dram_bytes"), "peak_source": peak_src,
                        "bytes_per_launch": by}
        roofline.update({"launches_per_step": c / n_prof, "ms_avg": ms_avg, "share_of_step": share,
                         "scatter_fraction_measured": scatter_frac or None})
        gname = f"b2n_hashgrid_fwd[n={RAYS_PER_GPU * S[2]}]"
        if gname in hash_rows:
            g = hash_rows[gname]
            gather_roof = {"bound": "hbm", "kernel": gname, "achieved": g["GBps"], "peak": hbm_peak, "unit": "GB/s",
                           "frac": g["frac_of_hbm_peak"], "traffic": ncu_of(gname, "dram_bytes"), "l2_bytes": ncu_of(gname, "lts_bytes"),
                           "bytes_per_launch": g["bytes_per_launch"], "ms_avg": g["ms_avg"], "peak_source": peak_src,
                           "note": "64 B per (sample, level); the 64 MiB table is L2-resident on B200, so DRAM traffic < algorithmic bytes"}
        # whole step against the hash roofline (1.99 GB algorithmic per 4096-ray step)
        step_bytes = sum(v["bytes_per_launch"] * v["launches_per_step"] for v in hash_rows.values())
        step_roof = {"algorithmic_hash_bytes_per_step": step_bytes, "GBps": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                     "frac_of_hbm_peak": step_bytes / (ms / args.steps * 1e-3) / 1e9 / hbm_peak}
    else:
        step_roof = None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample()
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch_rays": world * RAYS_PER_GPU,
                   "parallelism": (f"ray-batch data parallel x{world}: " + ("field gradients reduce-scattered, Adam on 1/N of the table, "
                                   "parameters all-gathered under the next step's proposal sampling; camera + proposal "
                                   "gradients all-reduced" if not args.no_sharded_update and args.engine != "autograd" else
                                   "flat-gradient all-reduce in two overlapped segments")) if world > 1 else "single GPU",
                   "precision": ("fp32 tables; MLPs on tcgen05 tensor cores, 3xTF32 split, fp32 accumulate in TMEM (1e-4 parity mode)"
                                 if (engine is not None and engine.tc) else "fp32 tables, fp32 SIMT MLPs (1e-4 parity mode)"), "optimizer": "fused Adam over one flat buffer", "engine": args.engine,
                   "proposal_update": "every step" if args.force_proposal_update else "reference schedule",
                   "camera_optimizer": f"{CAMERA_OPTIMIZER} (nerfacto default), trained inside the captured step" if CAMERA_OPTIMIZER != "off" else "off",
                   "l2": f"per-step working set {4 * 4 * n_params / 1e6:.0f} MB (params+grads+Adam moments) > 126 MB L2",
                   "timing": f"median of {args.windows} windows of exactly {args.steps} steps, each bracketed by barrier+synchronize; "
                             f"windows start at optimisation step {n_warm}",
                   "graph": ("one CUDA graph per step with parallel branches (proposal backward, main-grid position gradient, "
                             "gradient memset + weight packing beside the main chain)" if world == 1 else
                             "four graph pieces per step with the collectives between them"),
                   "params": n_params},
        "windows_ms": win_ms, "e2e_windows_ms": e2e_ms,
        "roofline": roofline, "roofline_hash_gather": gather_roof, "roofline_step": step_roof,
        "hash_kernels": hash_rows or None, "kernel_ms_per_step": kernel_table,
        "data_dependent_kernels": {"fraction_of_samples_with_gradient": scatter_frac or None, "ms_in_dense_gradient_regime": dense_ms or None,
                                   "note": "density_field_bwd skips warps whose samples all have zero d_density (exact zeros from the "
                                           "interlevel loss); kernel_ms_per_step is the regime of the timed windows, each "
                                           "kernel timed alone (branches serialised): in the captured step the proposal "
                                           "backward, hashgrid_dx, the memset and the weight packing run beside the main chain, "
                                           "so the table sums to more than ms_per_step"},
        "cpu_baseline": cpu, "eval_render": eval_line,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps,
                "through": ("FusedTrainStep.train_iteration(step): the reference's Trainer.train_iteration / "
                            "Pipeline.get_train_loss_dict surface (nerfstudio_b200/pipeline.py)") if fused is not None else "Trainer.train_iteration (autograd path)"},
        "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def run_ngp(args) -> None:
    """--workload ngp: BASELINE configs[1] (instant-ngp: 16-level hash grid T=2^19 F=2, 64-wide MLPs, 128^3 x 4 occupancy
    grid, cone_angle 0.004, alpha_thre 0.01) on the analytic-sphere scene of SURVEY 8d — the packed path end to end:
    occupancy update every 16 steps, march, sigma_fn + pruning, field on M packed samples, packed compositing, loss,
    backward, Adam.  M (samples per step) is data dependent and reported."""
    from nerfstudio_b200 import distributed as D
    from nerfstudio_b200 import lib
    from nerfstudio_b200.instant_ngp import InstantNGPModelConfig, NGPModel, NGPTrainer
    from nerfstudio_b200.scene import bundle_from, sphere_scene_rays

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 core has no CPU fallback")
    rank, local, world = D.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib.load()
    sampler = ClockSampler(local)
    torch.manual_seed(0)
    cfg = InstantNGPModelConfig(implementation="torch", background_color="black")
    model = NGPModel(cfg, torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]]), num_train_data=100).to(dev)
    allreduce = D.FlatGradAllReduce() if world > 1 else None
    trainer = NGPTrainer(model, allreduce=allreduce)
    D.broadcast_parameters(trainer.optim.flat)
    n_batches = 8
    host = []
    for b in range(n_batches):
        rays, gt = sphere_scene_rays(RAYS_PER_GPU, seed=1000 * rank + b)
        host.append(({k: v.pin_memory() for k, v in rays.items()}, gt.pin_memory()))
    resident = [({k: v.to(dev) for k, v in r.items()}, g.to(dev)) for r, g in host]
    torch.manual_seed(42 + rank)
    samples = []

    def step_resident(i):
        rays, gt = resident[i % n_batches]
        out = trainer.train_iteration(bundle_from(rays), {"image": gt})
        samples.append(trainer.num_samples)
        return out

    def step_e2e(i):
        rays, gt = host[i % n_batches]
        d_rays = {k: v.to(dev, non_blocking=True) for k, v in rays.items()}
        return float(trainer.train_iteration(bundle_from(d_rays), {"image": gt.to(dev, non_blocking=True)})["loss"].item())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    spans = []

    def timed(fn, steps, windows, first):
        out = []
        for w in range(windows):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            t0 = time.perf_counter()
            e0.record()
            for i in range(steps):
                fn(first + w * steps + i)
            e1.record()
            barrier()
            spans.append((t0, time.perf_counter()))
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
            out.append(float(ms.item()))
        return out

    n_warm = max(args.warmup, 320)  # past the occupancy grid's 256-step warm-up: the grid has converged to the sphere
    for i in range(n_warm):
        step_resident(i)
    samples.clear()
    lib.LAUNCHES = 0
    win_ms = timed(step_resident, args.steps, args.windows, n_warm)
    launches = lib.LAUNCHES
    m_mean = float(torch.stack([s.sum() for s in samples]).float().mean().item())
    ms = sorted(win_ms)[len(win_ms) // 2]
    value = world * RAYS_PER_GPU * args.steps / (ms * 1e-3)
    clocks = sampler.window(list(spans))
    for i in range(3):
        step_e2e(i)
    e2e_ms = timed(step_e2e, args.steps, args.windows, n_warm + 3)
    ms_e2e = sorted(e2e_ms)[len(e2e_ms) // 2]
    # per-kernel profile (CUDA events around every C-ABI launch), 32 steps = two occupancy updates
    lib.PROFILE_BY_SIZE, lib.PROFILE = False, {}
    n_prof = 32
    first = trainer.step
    for i in range(n_prof):
        step_resident(first + i)
    torch.cuda.synchronize()
    prof, sizes = lib.profile_summary(), lib.profile_sizes()
    lib.PROFILE, lib.PROFILE_BY_SIZE = None, True
    sampler.stop()
    if rank != 0:
        return
    hbm_peak, peak_src = peaks()
    table = {k: {"ms_per_step": round(t / n_prof, 4), "launches_per_step": round(c / n_prof, 2)} for k, (c, t) in
             sorted(prof.items(), key=lambda kv: -kv[1][1])}
    ms_prof = sum(t for _, t in prof.values()) / n_prof
    name, (c, t) = max(prof.items(), key=lambda kv: kv[1][1])
    roofline = {"kernel": name, "ms_per_step": t / n_prof, "launches_per_step": c / n_prof, "share_of_kernel_time": t / n_prof / ms_prof}
    hash_rows = {}
    for k, mult in (("b2n_hashgrid_fwd", 1), ("b2n_hashgrid_bwd", 2)):
        if k in prof:
            by = sizes[k] * 16 * 8 * 2 * 4 * mult  # 64 B per (sample, level), 16 levels; scatter = read + write
            gbps = by / (prof[k][1] * 1e-3) / 1e9
            hash_rows[k] = {"points_per_step": sizes[k] / n_prof, "GBps": gbps, "frac_of_hbm_peak": gbps / hbm_peak}
    if name in hash_rows:
        roofline.update({"bound": "hbm", "achieved": hash_rows[name]["GBps"], "peak": hbm_peak, "unit": "GB/s",
                         "frac": hash_rows[name]["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src})
    else:
        roofline.update({"bound": "latency", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                         "note": "the step is launch/latency bound at this sample count; see hash_kernels for the gather"})
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "instant-ngp 4096 rays/GPU, L16/T2^19/F2 grid + 64-wide MLPs, 128^3 x 4 occupancy grid, "
                               "cone_angle 0.004, alpha_thre 0.01, aabb +-1.5; analytic sphere r=0.5 (SURVEY 8d config 2)",
                   "global_batch_rays": world * RAYS_PER_GPU, "samples_per_step_M": m_mean,
                   "samples_per_ray": m_mean / RAYS_PER_GPU, "engine": "autograd over the drop-in modules (packed path)",
                   "background": "black (opaque synthetic targets)", "occupancy_update": "every 16 steps, inside the timed region",
                   "timing": f"median of {args.windows} windows of exactly {args.steps} steps; windows start at step {n_warm}"},
        "windows_ms": win_ms, "e2e_windows_ms": e2e_ms, "roofline": roofline, "hash_kernels": hash_rows,
        "kernel_ms_per_step": table, "cpu_baseline": None,
        "e2e": {"value": world * RAYS_PER_GPU * args.steps / (ms_e2e * 1e-3), "unit": "rays/s",
                "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host[0][0].values()) + host[0][1].numel() * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "through": "NGPTrainer.train_iteration"},
        "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def run_splat(args) -> None:
    """--workload splat: BASELINE configs[4] — splatfacto's rasteriser behind the gsplat API: 1 000 000 Gaussians (means
    U(-1,1)^3, scales exp(N(-4,0.5)), random unit quaternions, opacities sigmoid(N(0,1)), SH degree 3), one 1920x1080 camera
    at z = -3, fx = fy = 1200.  A step = one forward render (projection + SH + binning + sort + tile
    rasterisation); `--splat-train` adds the backward.  Reported as Gaussians/s and frames/s."""
    from nerfstudio_b200 import lib
    from nerfstudio_b200.shims import gsplat as G

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the B200 core has no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    lib.load()
    sampler = ClockSampler(dev.index)
    N, W, H = 1_000_000, 1920, 1080
    g = torch.Generator().manual_seed(0)
    means = (torch.rand(N, 3, generator=g) * 2 - 1).to(dev)
    quats = torch.randn(N, 4, generator=g).to(dev)
    scales = torch.exp(torch.randn(N, 3, generator=g) * 0.5 - 4.0).to(dev)
    opac = torch.sigmoid(torch.randn(N, generator=g)).to(dev)
    sh = (torch.randn(N, 16, 3, generator=g) * 0.2).to(dev)
    view = torch.eye(4)
    view[2, 3] = 3.0
    K = torch.tensor([[1200.0, 0, W / 2], [0, 1200.0, H / 2], [0, 0, 1]])
    vm, kk = view[None].to(dev), K[None].to(dev)
    leaves = [means, quats, scales, opac, sh]
    if args.splat_train:
        leaves = [t.requires_grad_(True) for t in leaves]

    def step(_i):
        out, alpha, info = G.rasterization(*leaves, vm, kk, W, H, sh_degree=3)
        if args.splat_train:
            for t in leaves:
                t.grad = None
            (out.mean() + alpha.mean()).backward()
        return out

    for i in range(max(args.warmup, 3)):
        step(i)
    spans, win = [], []
    lib.LAUNCHES = 0
    for _w in range(args.windows):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            out = step(i)
        e1.record()
        torch.cuda.synchronize()
        spans.append((t0, time.perf_counter()))
        win.append(e0.elapsed_time(e1))
    launches = lib.LAUNCHES
    ms = sorted(win)[len(win) // 2] / args.steps
    lib.PROFILE_BY_SIZE, lib.PROFILE = False, {}
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    prof = lib.profile_summary()
    lib.PROFILE, lib.PROFILE_BY_SIZE = None, True
    clocks = sampler.window(spans)
    sampler.stop()
    hbm_peak, peak_src = peaks()
    # algorithmic bytes of the projection: read 10 floats + 48 SH floats, write 2+1+3+1+1+3 per Gaussian
    proj_bytes = N * 4 * (3 + 4 + 3 + 48 + 11)
    kt = {k: round(t / 3, 4) for k, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    pj = prof.get("b2n_gs_project_fwd")
    line = {"metric": "gaussians_per_sec", "value": N / (ms * 1e-3), "unit": "gaussians/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "splatfacto rasteriser (gsplat API): 1M Gaussians, SH degree 3, 1920x1080, one camera "
                                   "(BASELINE configs[4])", "mode": "forward + backward" if args.splat_train else "forward render",
                       "sort": "64-bit key radix sort = torch.sort (library primitive, as gsplat uses cub)"},
            "fps": 1e3 / ms, "windows_ms": win, "kernel_ms_per_step": kt, "alpha_mean": float(out.abs().mean().item()),
            "roofline": {"bound": "hbm", "kernel": "b2n_gs_project_fwd", "achieved": proj_bytes / (pj[1] / pj[0] * 1e-3) / 1e9 if pj else None,
                         "peak": hbm_peak, "unit": "GB/s", "frac": (proj_bytes / (pj[1] / pj[0] * 1e-3) / 1e9 / hbm_peak) if pj else None,
                         "traffic": None, "peak_source": peak_src,
                         "note": "projection + SH: 268 B per Gaussian; the tile rasteriser is latency/occupancy bound (see kernel_ms_per_step)"},
            "cpu_baseline": None, "e2e": {"value": N / (ms * 1e-3), "unit": "gaussians/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": 0,
                                          "through": "shims.gsplat.rasterization (gsplat.rendering.rasterization surface)"},
            "gpu_launches": launches, "clocks": clocks}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample():
    threads, _ = pick_cpu_threads()
    rps, sec, cores, kind = time_cpu(512, 3, 1, threads)
    return {"value": rps, "unit": "rays/s", "cores": cores, "kind": kind,
            "sample": f"3 steps x 512 rays of the same workload (fwd+bwd+Adam; {'the unmodified reference model' if kind == 'reference' else 'oracle port'}), "
                      f"torch CPU fp32, {cores} threads"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="nerfacto", choices=["nerfacto", "ngp", "splat"],
                    help="nerfacto = BASELINE configs[2], the metric's workload (default); ngp = configs[1] (instant-ngp); "
                         "splat = configs[4] (3DGS rasteriser behind the gsplat API)")
    ap.add_argument("--splat-train", action="store_true", help="splat workload: include the backward pass")
    ap.add_argument("--windows", type=int, default=3, help="timed windows of exactly --steps steps each (median reported)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--state-step", type=int, default=STATE_STEP,
                    help="optimisation step at which the timed windows start (profiling runs under ncu use a small value)")
    ap.add_argument("--no-eval", action="store_true", help="skip the full-image eval-render measurement")
    ap.add_argument("--no-sharded-update", action="store_true",
                    help="N > 1: all-reduce the whole gradient buffer and run Adam replicated (instead of reduce-scatter -> "
                         "Adam on 1/N of the hash table -> all-gather overlapped with the next step's proposal sampling)")
    ap.add_argument("--mlp", default="auto", choices=["auto", "tc", "simt"],
                    help="tiny-MLP kernels of the graph/eager engine: tcgen05 3xTF32 (tc) or fp32 SIMT")
    ap.add_argument("--unfused-proposals", action="store_true", help="proposal networks as separate grid/MLP launches")
    ap.add_argument("--tune", default="", help="comma separated key=value launch-geometry knobs (lib.tune)")
    ap.add_argument("--engine", default="graph", choices=["graph", "eager", "autograd"],
                    help="graph: CUDA-graph replay of the hand-written step (default); eager: same launches without a "
                         "graph; autograd: the drop-in modules under torch.autograd")
    ap.add_argument("--reference-schedule", dest="force_proposal_update", action="store_false",
                    help="use nerfacto's proposal-update schedule instead of training the proposal nets every step")
    ap.add_argument("--camera-optimizer", default="SO3xR3", choices=["SO3xR3", "off"],
                    help="nerfacto's per-camera pose optimiser (default: on, as in the reference's nerfacto config)")
    args = ap.parse_args()
    global CAMERA_OPTIMIZER
    CAMERA_OPTIMIZER = args.camera_optimizer
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "ngp":
        run_ngp(args)
    elif args.workload == "splat":
        run_splat(args)
    else:
        run_b200(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
