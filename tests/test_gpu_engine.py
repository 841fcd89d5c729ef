"""GPU: the graph-captured step (engine.NerfactoStep) does the same arithmetic as the autograd path over the
drop-in modules (which test_gpu_modules.py pins to the reference), eagerly and when replayed from a CUDA graph."""
import copy

import pytest
import torch

from conftest import assert_close, assert_grad_close
from test_gpu_modules import _bundle, _pipeline_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _mk(g, use_graph, always=True, **kw):
    from nerfstudio_b200.engine import NerfactoStep

    model = _pipeline_model(g).train()
    eng = NerfactoStep(model, n_rays=g["origins"].shape[0], use_graph=use_graph, always_update_proposals=always, **kw)
    eng.set_batch(g["origins"].cuda(), g["directions"].cuda(), g["train_cams"].cuda(), g["gt"].cuda())
    eng.fixed_jitter = [g["train_rand0"].cuda(), g["train_rand1"].cuda(), g["train_rand2"].cuda()]
    return model, eng


def test_engine_step_equals_reference_losses_and_grads(cuda, golden):
    """Losses and every parameter gradient of the hand-written backward vs the reference's autograd (golden)."""
    g = golden("nerfacto_pipeline")
    model, eng = _mk(g, use_graph=False)
    model.proposal_sampler.set_anneal(0.7)
    # golden was recorded with anneal 0.7: drive the engine's device-side exponent to the same value
    eng._anneal = lambda step: 0.7
    eng.optim.lr = 0.0  # keep the weights so that gradients can be read back unchanged
    losses = eng.step().cpu()
    assert_close(losses[0], g["loss_rgb"], 1e-4)
    # a sum of a few clipped squares: the ~1e-5 sample shifts of two stacked inverse-CDF steps move it by 4.8e-4
    # (measured, scripts/composed_errors.py); with the reference's samples fed in (staged test below) it meets 1e-4
    assert_close(losses[1], g["loss_interlevel"], 1e-3)
    assert_close(losses[2], g["loss_distortion"], 1e-4)
    assert_close(losses[3], g["loss"], 1e-4)
    from test_gpu_modules import _named_params

    for k, p in _named_params(model).items():
        # switching gradients (see conftest.assert_grad_close): everything that only the interlevel loss reaches
        # (proposal networks) and the hash tables; the MLP weights are compared entry-wise at 2.5e-3 (measured: up to
        # 1.2e-3 on the base MLP's first layer, 1e-5 and below on the head) — end to end they sit behind two
        # ill-conditioned resampling steps: the density networks agree with torch's to 1.6e-7, the inverse-CDF steps turn
        # that into 7e-6 / 1.7e-5 sample shifts, and the x1000 test tables amplify those.  The 1e-4 bar is enforced
        # entry-wise on identical inputs per operator (test_gpu_ops.py / test_gpu_tc.py) and composed in the staged test
        assert_grad_close(p.grad, g["g_" + k], "g_" + k, 2.5e-3, sparse_switching=k.startswith("p") or "table" in k)
    assert_close(eng.rgb_out, g["train_rgb"], 1e-4)
    assert_close(eng.acc[:, None], g["train_acc"], 1e-4)


def test_engine_staged_levels_entrywise(cuda, golden):
    """The hand-written backward held to the 1e-4 bar entry by entry: with every level fed the reference's recorded
    samples (`train_sbins{i}` / `train_ebins{i}`) the captured step computes exactly the reference's graph, so weights,
    all losses and every parameter gradient — hash tables and proposal networks included — must agree at 1e-4 of the
    max-norm, no direction/norm comparison."""
    from test_gpu_modules import _named_params

    g = golden("nerfacto_pipeline")
    model, eng = _mk(g, use_graph=False)
    eng._anneal = lambda step: 0.7
    eng.optim.lr = 0.0
    eng.fixed_bins = [(g[f"train_sbins{i}"].cuda(), g[f"train_ebins{i}"].cuda()) for i in range(3)]
    losses = eng.step().cpu()
    for i in range(3):
        assert_close(eng.w[i], g[f"train_w{i}"][..., 0], 1e-4, f"w{i}")
    assert_close(losses[0], g["loss_rgb"], 1e-4), assert_close(losses[1], g["loss_interlevel"], 1e-4)
    assert_close(losses[2], g["loss_distortion"], 1e-4), assert_close(losses[3], g["loss"], 1e-4)
    assert_close(eng.rgb_out, g["train_rgb"], 1e-4), assert_close(eng.depth_exp[:, None], g["train_exp_depth"], 1e-4)
    assert torch.equal(eng.depth_med.cpu()[:, None], g["train_depth"]), "median depth sample"
    for k, p in _named_params(model).items():
        # the proposal MLPs' gradients are sums of ~3000 signed terms that cancel to ~1e-6 (|sum| / sum|terms| ~ 1e-2):
        # fp32 summation order alone moves them by 1e-4 of their max-norm in either implementation -> 3e-4 there
        assert_close(p.grad, g["g_" + k], 3e-4 if (k.startswith("p") and "table" not in k) else 1e-4, "staged g_" + k)


def test_engine_camera_optimizer_pose_gradients(cuda, golden):
    """nerfacto's default has the camera optimiser ON: the captured step applies the SO3xR3 pose corrections, carries the
    photometric / interlevel / distortion gradients back through the sample positions of all three levels (hashgrid_dx,
    fused proposal backward with ray gradients, positions_bwd) and adds the regulariser.  Staged on the reference's
    recorded samples; pose gradient and losses against the reference's autograd (pipeline_camopt golden)."""
    from nerfstudio_b200.engine import NerfactoStep

    g = golden("pipeline_camopt")
    model = _pipeline_model(g, camera_optimizer="SO3xR3").train()
    with torch.no_grad():
        model.camera_optimizer.pose_adjustment.copy_(g["pose"].cuda())
    # (the proposal backward and the main grid's position gradient run as forked branches; also checked serialised)
    for use_graph, concurrent in ((False, True), (True, True), (True, False)):
        eng = NerfactoStep(model, n_rays=g["origins"].shape[0], use_graph=use_graph, always_update_proposals=True,
                           concurrent_backward=concurrent)
        assert eng.camopt is not None and eng.optim.segment_of("camera_opt") is not None
        assert eng.grad_split == eng.optim.segment_of("camera_opt")[0]  # [field | camera_opt | proposals]
        eng.set_batch(g["origins"].cuda(), g["directions"].cuda(), g["cams"].cuda(), g["gt"].cuda())
        eng.fixed_jitter = [g["rand0"].cuda(), g["rand1"].cuda(), g["rand2"].cuda()]
        eng.fixed_bins = [(g[f"sbins{i}"].cuda(), g[f"ebins{i}"].cuda()) for i in range(3)]
        eng._anneal = lambda step: 0.7
        eng.optim.lr, eng.camera_lr = 0.0, 0.0
        losses = eng.step().cpu()
        torch.cuda.synchronize()
        assert_close(losses[0], g["loss_rgb"], 1e-4), assert_close(losses[1], g["loss_interlevel"], 1e-4)
        assert_close(losses[2], g["loss_distortion"], 1e-4), assert_close(losses[4], g["regularizer"], 1e-4)
        assert_close(losses[3], g["loss"], 1e-4)
        assert_close(eng.rgb_out, g["rgb"], 1e-4)
        assert_close(model.camera_optimizer.pose_adjustment.grad, g["g_pose_total"], 1e-3, "d loss / d pose")
    # and it trains: poses move, loss goes down, the autograd path over the modules agrees on the first losses
    from nerfstudio_b200.nerfacto import Trainer
    from test_gpu_modules import FakeRand

    m2 = _pipeline_model(g, camera_optimizer="SO3xR3").train()
    with torch.no_grad():
        m2.camera_optimizer.pose_adjustment.copy_(g["pose"].cuda())
    eng = NerfactoStep(model, n_rays=g["origins"].shape[0], use_graph=True, always_update_proposals=True)
    eng.set_batch(g["origins"].cuda(), g["directions"].cuda(), g["cams"].cuda(), g["gt"].cuda())
    eng.fixed_jitter = [g["rand0"].cuda(), g["rand1"].cuda(), g["rand2"].cuda()]
    tr = Trainer(m2)
    p0 = model.camera_optimizer.pose_adjustment.detach().clone()
    for it in range(3):
        with FakeRand([g["rand0"], g["rand1"], g["rand2"]]):
            stats = tr.train_iteration(_bundle(g["origins"], g["directions"], g["cams"]), {"image": g["gt"].cuda()})
        l = eng.step()
        assert_close(l[3], stats["loss"], 2e-3 if it else 1e-4, f"loss step {it}")
    assert float((model.camera_optimizer.pose_adjustment - p0).abs().max()) > 1e-4
    rel = float((model.camera_optimizer.pose_adjustment - m2.camera_optimizer.pose_adjustment).norm()
                / m2.camera_optimizer.pose_adjustment.norm())
    assert rel < 5e-2, rel


def test_frozen_proposal_networks_are_not_stepped(cuda, golden):
    """Reference schedule: after step 10 the proposal networks are trained only every few steps; on the other steps
    their gradients are None in the reference and Optimizers.optimizer_scaler_step_all skips the group
    (engine/optimizers.py:142-159) — parameters and Adam moments stay bit-identical, and the group's bias-correction count
    does not advance.  Checked for the captured step (eager and graph) and for the autograd Trainer."""
    from nerfstudio_b200.nerfacto import Trainer

    g = golden("nerfacto_pipeline")
    for use_graph in (False, True):
        model, eng = _mk(g, use_graph=use_graph, always=False)
        model.proposal_sampler.update_sched = lambda step: 3  # update when 4+ steps have passed (or step < 10)
        eng.fixed_jitter = None
        a, b = eng.optim.segment_of("proposal_networks")
        frozen_seen = 0
        for it in range(16):
            due = eng._update_due(it)
            before = [t[a:b].clone() for t in (eng.optim.flat, eng.optim.exp_avg, eng.optim.exp_avg_sq)]
            f_before = eng.optim.flat[:a].clone()
            eng.step()
            torch.cuda.synchronize()
            after = [t[a:b] for t in (eng.optim.flat, eng.optim.exp_avg, eng.optim.exp_avg_sq)]
            if due:
                assert not torch.equal(before[0], after[0]), it
            else:
                frozen_seen += 1
                assert all(torch.equal(x, y) for x, y in zip(before, after)), f"step {it}: frozen proposals moved"
            assert not torch.equal(f_before, eng.optim.flat[:a]), it  # the field trains every step
        assert frozen_seen >= 3 and eng.optim.group_steps["proposal_networks"] == 16 - frozen_seen
        assert eng.optim.group_steps["fields"] == 16
    model = _pipeline_model(g).train()
    model.proposal_sampler.update_sched = lambda step: 3
    tr = Trainer(model)
    a, b = tr.optim.segment_of("proposal_networks")
    frozen_seen = 0
    for it in range(14):
        before = tr.optim.flat[a:b].clone()
        tr.train_iteration(_bundle(g["origins"], g["directions"], g["train_cams"]), {"image": g["gt"].cuda()})
        if not model.proposal_sampler.last_updated:
            frozen_seen += 1
            assert torch.equal(before, tr.optim.flat[a:b]), it
    assert frozen_seen >= 2


def test_engine_matches_autograd_trainer_over_steps(cuda, golden):
    from nerfstudio_b200.nerfacto import Trainer
    from test_gpu_modules import FakeRand

    g = golden("nerfacto_pipeline")
    model, eng = _mk(g, use_graph=False)
    ref = copy.deepcopy(model)
    tr = Trainer(ref)
    draws = [g["train_rand0"], g["train_rand1"], g["train_rand2"]]
    batch = {"image": g["gt"].cuda()}
    for it in range(3):
        with FakeRand(list(draws)):
            stats = tr.train_iteration(_bundle(g["origins"], g["directions"], g["train_cams"]), batch)
        losses = eng.step()
        assert_close(losses[3], stats["loss"], 1e-4, f"loss step {it}")
    # Adam with eps=1e-15 is scale invariant: an entry whose gradient is rounding noise still moves by ~lr, so
    # parameters are compared where they matter (the big tables, by relative L2) rather than entry by entry
    for (k, a), (_, b) in zip(model.state_dict().items(), ref.state_dict().items()):
        if a.dtype.is_floating_point and a.numel() > 1:
            rel = float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
            assert rel < 2e-2, (k, rel)


def test_graph_replay_equals_eager(cuda, golden):
    g = golden("nerfacto_pipeline")
    m_e, eager = _mk(g, use_graph=False)
    m_g, graph = _mk(g, use_graph=True)
    for it in range(4):
        le, lg = eager.step().clone(), graph.step().clone()
        torch.cuda.synchronize()
        assert_close(lg, le, 1e-6, f"losses step {it}")
    for (k, a), (_, b) in zip(m_g.state_dict().items(), m_e.state_dict().items()):
        if a.dtype.is_floating_point and a.numel() > 1:
            # float atomics reorder between runs and Adam (eps 1e-15) turns a sign flip of a rounding-noise gradient into
            # a +-lr move of that entry: compare in relative L2, where those isolated entries do not dominate
            rel = float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
            assert rel < 1e-4, (k, rel)


def test_fused_ray_tail_equals_operator_chain(cuda, golden):
    """b2n_nerfacto_ray_tail (weights, renderers, losses and their backward for all three levels in one launch) runs the
    per-ray bodies of the separate operators: every deterministic product is bit-identical to the 15-launch chain, the
    loss scalars (different reduction order) agree to 1e-6, the parameter gradients (float atomics downstream) to 1e-5."""
    g = golden("nerfacto_pipeline")
    outs = []
    for fused in (False, True):
        model, eng = _mk(g, use_graph=False, fused_tail=fused)
        assert eng.fused_tail == fused
        eng.optim.lr = 0.0
        losses = eng.step().clone()
        torch.cuda.synchronize()
        outs.append((eng, losses))
    (a, la), (b, lb) = outs
    for name in ("rgb_out", "acc", "depth_exp", "depth_med", "d_rgb", "d_hpre", "d_w_dist"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    for lvl in range(3):
        assert torch.equal(a.w[lvl], b.w[lvl]) and torch.equal(a.dens[lvl], b.dens[lvl]), lvl
        assert torch.equal(a.d_w[lvl], b.d_w[lvl]), lvl
        assert torch.equal(a.rows[lvl], b.rows[lvl]), lvl
    for lvl in range(2):
        assert torch.equal(a.d_dens[lvl], b.d_dens[lvl]), lvl
    assert_close(lb, la, 1e-6, "losses")
    assert_close(b.optim.flat_grad, a.optim.flat_grad, 1e-5, "flat gradient")


def test_branches_do_not_change_the_update(cuda, golden):
    """The captured step with its parallel branches (proposal backward + its Adam, position gradient, static head-input
    columns, embedding-row gradients, ...) against the same step issued as one chain: same losses and, after three
    optimisation steps, the same parameters (up to float-atomics order)."""
    g = golden("nerfacto_pipeline")
    m_c, chain = _mk(g, use_graph=True, concurrent_backward=False)
    m_b, branched = _mk(g, use_graph=True, concurrent_backward=True)
    assert branched.concurrent and not chain.concurrent
    for it in range(3):
        lc, lb = chain.step().clone(), branched.step().clone()
        torch.cuda.synchronize()
        assert_close(lb, lc, 1e-5, f"losses step {it}")
    for (k, a), (_, b) in zip(m_b.state_dict().items(), m_c.state_dict().items()):
        if a.dtype.is_floating_point and a.numel() > 1:
            rel = float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
            assert rel < 1e-4, (k, rel)
    assert float(branched.optim.exp_avg.abs().sum()) > 0


def test_engine_trains(cuda, golden):
    """Sanity of the optimisation loop under graph replay with fresh random jitter: the loss goes down."""
    g = golden("nerfacto_pipeline")
    model, eng = _mk(g, use_graph=True, always=False)
    eng.fixed_jitter = None
    torch.manual_seed(0)
    first = None
    for it in range(60):
        l = eng.step()
        if it == 0:
            first = float(l[0])
    last = float(eng.losses[0])
    assert last < 0.6 * first, (first, last)


def test_graph_engine_converges_like_autograd_on_teacher_scene(cuda):
    """Held-out PSNR over 300 un-synchronised graph steps (the CPU runs ahead of the stream the whole time) on a scene
    rendered by a fixed random teacher field: the best of three late checkpoints is well above the start (a collapsed
    run sits at ~6 dB) and within 5 dB of the autograd path's.  The band is wide on purpose: single checkpoints of
    either path move by 1-4 dB from one evaluation to the next (atomics-ordered trajectories at lr 1e-2); exactness
    of the step itself is pinned by the tests above."""
    import math

    from nerfstudio_b200.engine import NerfactoStep
    from nerfstudio_b200.nerfacto import NerfactoModel, NerfactoModelConfig, Trainer
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
    R, NB = 4096, 32

    def batch(seed):
        rays, _ = synthetic_rays(R, 8, seed)
        rays = {k: v.cuda() for k, v in rays.items()}
        with torch.no_grad():
            return rays, teacher(bundle_from(rays))["rgb"]

    train, held = [batch(s) for s in range(NB)], [batch(10_000 + s) for s in range(2)]

    def psnr(model):
        model.eval()
        with torch.no_grad():
            mse = sum(float(((model(bundle_from(r))["rgb"] - gt) ** 2).mean()) for r, gt in held) / len(held)
        model.train()
        return -10 * math.log10(mse)

    res = {}
    for name in ("graph", "autograd"):
        torch.manual_seed(7)
        student = NerfactoModel(cfg(), aabb, 8).cuda().train()
        eng = NerfactoStep(student, R, use_graph=True) if name == "graph" else Trainer(student)
        start, late = psnr(student), []
        for it in range(300):
            rays, gt = train[it % NB]
            if name == "graph":
                eng.set_batch(rays["origins"], rays["directions"], rays["camera_indices"], gt)
                eng.step()
            else:
                eng.train_iteration(bundle_from(rays), {"image": gt})
            if it + 1 in (200, 250, 300):
                late.append(psnr(student))
        res[name] = (start, max(late))
    assert res["graph"][1] > res["graph"][0] + 8.0, res
    assert abs(res["graph"][1] - res["autograd"][1]) < 5.0, res


def test_packed_batch_equals_set_batch(cuda, golden):
    """pack_batch + set_batch_packed (one copy) fill the static inputs exactly like set_batch (four copies)."""
    g = golden("nerfacto_pipeline")
    model, eng = _mk(g, use_graph=False)
    ref = [t.clone() for t in (eng.origins, eng.directions, eng.cams, eng.gt)]
    blob = eng.pack_batch(g["origins"], g["directions"], g["train_cams"], g["gt"])
    assert blob.is_pinned()
    eng.inputs.zero_()
    eng.set_batch_packed(blob)
    torch.cuda.synchronize()
    for a, b in zip(ref, (eng.origins, eng.directions, eng.cams, eng.gt)):
        assert torch.equal(a, b)


def test_render_engine_matches_golden_eval_and_module_path(cuda, golden):
    """the graph-captured chunk loop (render_engine.NerfactoRender) vs the reference's recorded eval outputs
    and vs the module path's chunked `get_outputs_for_camera_ray_bundle`, including a ragged last chunk."""
    from nerfstudio_b200.render_engine import NerfactoRender
    from test_gpu_modules import _bundle

    g = golden("nerfacto_pipeline")
    model = _pipeline_model(g).eval()
    model.proposal_sampler.set_anneal(1.0)
    R = g["origins"].shape[0]
    for chunk, graph in ((R, False), (40, True), (64, True)):  # 96 rays: one chunk / 40+40+16 / 64+32
        ren = NerfactoRender(model, chunk_rays=chunk, use_graph=graph)
        ren.step.nears.fill_(0.05)  # the golden eval pass was recorded with explicit nears = 0.05 (no collider)
        out = ren.render_rays(g["origins"].cuda(), g["directions"].cuda(), g["eval_cams"].cuda())
        assert_close(out["rgb"], g["eval_rgb"], 1e-4, f"rgb chunk={chunk}")
        assert_close(out["accumulation"], g["eval_acc"], 1e-4)
        assert_close(out["expected_depth"], g["eval_exp_depth"], 1e-4)
        ref_d = g["eval_depth"]
        same = ((out["depth"].cpu() - ref_d).abs() <= 1e-3 * ref_d.abs()).float().mean().item()
        assert same >= 0.97, f"median depth agrees on {same:.3f} of rays"
    # through the collider (eval: near plane reset to 0, scene_colliders.py:169-191) against the module path
    out = NerfactoRender(model, chunk_rays=64, use_graph=True).render_rays(g["origins"].cuda(), g["directions"].cuda(),
                                                                            g["eval_cams"].cuda())
    rb = _bundle(g["origins"], g["directions"], g["eval_cams"])
    rb.nears = rb.fars = None
    with torch.no_grad():
        mod = model.get_outputs_for_camera_ray_bundle(rb)
    for k in ("rgb", "accumulation", "expected_depth"):
        assert_close(out[k], mod[k], 1e-4, k)  # tensor-core (3xTF32) vs SIMT MLPs in the main field
    for k in ("depth", "prop_depth_0", "prop_depth_1"):
        same = ((out[k] - mod[k]).abs() <= 1e-5 * mod[k].abs()).float().mean().item()
        assert same >= 0.97, (k, same)


def test_render_engine_whole_camera(cuda, golden):
    """`get_outputs_for_camera`: keep_shape rays from the ray-generation kernel, [H, W, C] outputs, chunked == unchunked."""
    from nerfstudio_b200.cameras.cameras import Cameras
    from nerfstudio_b200.render_engine import NerfactoRender

    g = golden("nerfacto_pipeline")
    model = _pipeline_model(g).eval()
    H, W = 24, 40
    c2w = torch.eye(4)[:3][None].repeat(2, 1, 1)
    c2w[:, :, 3] = torch.tensor([[0.0, 0.0, 1.5], [0.3, -0.2, 1.2]])
    cams = Cameras(c2w.cuda(), torch.full((2,), 30.0).cuda(), torch.full((2,), 30.0).cuda(), torch.full((2,), W / 2).cuda(),
                   torch.full((2,), H / 2).cuda(), width=torch.full((2,), W).cuda(), height=torch.full((2,), H).cuda())
    a = NerfactoRender(model, chunk_rays=H * W, use_graph=False).render_camera(cams, 1)
    b = NerfactoRender(model, chunk_rays=256, use_graph=True).render_camera(cams, 1)
    assert a["rgb"].shape == (H, W, 3) and a["depth"].shape == (H, W, 1)
    for k in a:
        assert_close(b[k], a[k], 1e-6, k)
    assert torch.isfinite(a["rgb"]).all() and float(a["accumulation"].max()) <= 1.0 + 1e-5
