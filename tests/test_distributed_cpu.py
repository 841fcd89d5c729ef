"""CPU, world_size 2 over gloo: the host logic of the ray-batch data-parallel path (flat gradient all-reduce,
parameter broadcast, rank-0-only reference arm).  No CUDA calls."""
import os
import socket
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from nerfstudio_b200 import distributed as D

    r, l, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    flat = torch.full((1000,), float(rank + 1))
    scale = D.FlatGradAllReduce()(flat)
    params = torch.full((10,), float(rank))
    D.broadcast_parameters(params, src=0)
    out[rank] = (float(flat[0]), scale, float(params[0]))
    dist.destroy_process_group()


def test_flat_grad_allreduce_and_broadcast_world2():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in (0, 1):
        total, scale, p0 = out[rank]
        assert total == 3.0      # 1 + 2: summed over ranks
        assert scale == 0.5      # the mean is taken inside the fused Adam kernel (grad_scale)
        assert p0 == 0.0         # every replica starts from rank 0's weights


def _worker_segments(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from nerfstudio_b200 import distributed as D

    D.init_from_env("gloo")
    ar = D.FlatGradAllReduce()
    flat = torch.arange(12.0) * (rank + 1)  # [field segment | proposal segment], split at 8 like engine.grad_split
    h_field = ar.start(flat[:8])            # summed while the "proposal backward" would run ...
    flat[8:] += 100.0 * (rank + 1)          # ... which still writes the other segment
    h_prop = ar.start(flat[8:])
    ar.finish(h_field, h_prop)
    out[rank] = flat.tolist()
    dist.destroy_process_group()


def test_segmented_async_allreduce_world2():
    """engine.NerfactoStep at N > 1: two async collectives over views of ONE flat buffer, summed in place."""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_segments, args=(2, port, out), nprocs=2, join=True)
    expect = [3.0 * i for i in range(8)] + [3.0 * i + 300.0 for i in range(8, 12)]
    assert out[0] == expect and out[1] == expect


def _worker_sharded(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from nerfstudio_b200 import distributed as D

    D.init_from_env("gloo")
    ar = D.FlatGradAllReduce()
    n_field, n_all = 21, 30                  # field segment [0, 21), late segments [21, 30) as engine.grad_split = 21
    chunk = ar.shard_chunk(n_field)          # (21 // 2) // 4 * 4 = 8 -> slices [0,8) [8,16), tail [16, 21)
    params = torch.zeros(n_all)
    grad = torch.arange(float(n_all)) * (rank + 1)
    h_rs = ar.start_reduce_scatter(grad[: world * chunk])
    h_rest = ar.start(grad[world * chunk:])  # field tail + camera + proposal gradients
    ar.finish(h_rs, h_rest)
    # "Adam" (p -= g) on what this rank owns: its slice, the field tail, the late segments
    lo, hi = rank * chunk, (rank + 1) * chunk
    params[lo:hi] -= grad[lo:hi]
    params[world * chunk:] -= grad[world * chunk:]
    h_ag = ar.start_all_gather(params[: world * chunk])
    ar.finish(h_ag)
    out[rank] = (chunk, params.tolist())
    dist.destroy_process_group()


def test_sharded_update_world2():
    """engine.NerfactoStep._step_sharded's exchange: reduce-scatter of the field gradients, all-reduce of the rest, every
    rank updates its slice, all-gather of the parameters — replicas end up with the all-reduce result."""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_sharded, args=(2, port, out), nprocs=2, join=True)
    expect = [-3.0 * i for i in range(30)]
    for rank in (0, 1):
        chunk, params = out[rank]
        assert chunk == 8
        assert params == expect, (rank, params)


def test_single_process_allreduce_is_identity():
    sys.path.insert(0, ROOT)
    from nerfstudio_b200.distributed import FlatGradAllReduce

    g = torch.arange(8.0)
    assert FlatGradAllReduce()(g) == 1.0
    assert torch.equal(g, torch.arange(8.0))


def test_reference_arm_runs_on_rank0_only():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 prints the JSON line, rank 1 exits 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_flat_adam_skips_frozen_groups(monkeypatch):
    """optim.FlatAdam.step(active=...): a group that received no gradients this step is not touched (parameters, moments,
    bias-correction count) — the reference's Optimizers.optimizer_scaler_step_all skips groups whose grads are None
    (engine/optimizers.py:142-159).  Host logic only: the kernel call is replaced by the oracle's Adam."""
    from oracle import nerf_oracle as O
    from nerfstudio_b200 import functional as F
    from nerfstudio_b200.optim import FlatAdam

    calls = []

    def fake_adam(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-15, grad_scale=1.0):
        calls.append((p.data_ptr(), p.numel(), step))
        O.adam_step(p, g * grad_scale, m, v, step, lr, betas[0], betas[1], eps)

    monkeypatch.setattr(F, "adam_step", fake_adam)

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.field = torch.nn.Linear(5, 3)
            self.proposal_networks = torch.nn.Linear(4, 2)

        def get_param_groups(self):
            return {"proposal_networks": list(self.proposal_networks.parameters()), "fields": list(self.field.parameters())}

    torch.manual_seed(0)
    m = M()
    opt = FlatAdam(m, lr=1e-2)
    assert [s[0] for s in opt.segments] == ["fields", "proposal_networks"] and opt.segment_of("proposal_networks")[1] == opt.flat.numel()
    ref = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    ref_opt = {"fields": torch.optim.Adam(ref[:2], lr=1e-2, eps=1e-15), "proposal_networks": torch.optim.Adam(ref[2:], lr=1e-2, eps=1e-15)}
    for step, active in enumerate([None, ["fields"], ["fields"], None]):
        opt.zero_grad()
        gs = [torch.randn_like(p) for p in ref]
        for p, r, gr in zip(m.parameters(), ref, gs):
            is_prop = any(p is q for q in m.proposal_networks.parameters())
            live = active is None or not is_prop
            p.grad.copy_(gr if live else torch.zeros_like(gr))
            r.grad = gr.clone() if live else None
        calls.clear()
        opt.step(active=active)
        for name, o in ref_opt.items():
            if active is None or name in active:
                o.step()
        assert len(calls) == (1 if step == 0 else (1 if active else 2)), calls  # equal counters share one launch
        for p, r in zip(m.parameters(), ref):
            assert torch.allclose(p, r, atol=1e-7), step
    assert opt.group_steps == {"fields": 4, "proposal_networks": 2}


def _worker_render_shards(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from nerfstudio_b200 import distributed as D
    from nerfstudio_b200.render_engine import chunk_owner

    D.init_from_env("gloo")
    N, chunk = 1000, 96
    img = torch.zeros(N, 3)
    for ci, i in enumerate(range(0, N, chunk)):
        if chunk_owner(ci, world) == rank:
            img[i: i + chunk] = torch.arange(i, min(N, i + chunk)).float()[:, None]
    dist.all_reduce(img, op=dist.ReduceOp.SUM)  # every pixel written by exactly one rank: the sum is a gather
    out[rank] = img[:, 0].tolist()
    dist.destroy_process_group()


def test_render_chunk_sharding_gathers_the_image():
    """Multi-GPU eval: chunks dealt round-robin, image planes summed — host logic at world size 2 (gloo)."""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_render_shards, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == [float(i) for i in range(1000)] and out[1] == out[0]
