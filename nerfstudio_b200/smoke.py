"""smoke(): one small nerfacto training step on cuda:0 through the C-ABI, checked against the CPU oracle."""
from __future__ import annotations

import copy

import torch


def run() -> None:
    from oracle import nerf_oracle as O  # checker only

    from .nerfacto import NerfactoModel, NerfactoModelConfig
    from .scene import bundle_from, synthetic_rays

    torch.manual_seed(0)
    cfg = NerfactoModelConfig(
        num_levels=8, max_res=512, log2_hashmap_size=13, num_proposal_samples_per_ray=(32, 20),
        num_nerf_samples_per_ray=12, average_init_density=0.01, implementation="torch",
        proposal_net_args_list=[
            {"hidden_dim": 16, "log2_hashmap_size": 12, "num_levels": 5, "max_res": 128, "use_linear": False},
            {"hidden_dim": 16, "log2_hashmap_size": 12, "num_levels": 5, "max_res": 256, "use_linear": False}])
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    model = NerfactoModel(cfg, aabb, num_train_data=8)
    with torch.no_grad():  # non-trivial densities
        model.field.mlp_base.model[0].hash_table.mul_(1000.0)
        for p in model.proposal_networks:
            p.encoding.hash_table.mul_(2000.0)
    cpu_state = copy.deepcopy(model.state_dict())
    model = model.cuda().eval()
    R = 256
    rays, gt = synthetic_rays(R, num_images=8, seed=1)
    out = model(bundle_from({k: v.cuda() for k, v in rays.items()}))

    # oracle on the same weights / rays (eval mode: deterministic sampling)
    def field_params(prefix, with_head):
        sd = cpu_state
        if not with_head:
            return dict(table=sd[f"{prefix}.encoding.hash_table"], log2_T=12,
                        scalings=O.hash_level_scalings(5, 16, 128 if prefix.endswith("0") else 256),
                        w=[sd[f"{prefix}.mlp_base.1.layers.{i}.weight"] for i in range(2)],
                        b=[sd[f"{prefix}.mlp_base.1.layers.{i}.bias"] for i in range(2)])
        return dict(table=sd["field.mlp_base.model.0.hash_table"], log2_T=13, scalings=O.hash_level_scalings(8, 16, 512),
                    embedding=sd["field.embedding_appearance.embedding.weight"],
                    w_base=[sd[f"field.mlp_base.model.1.layers.{i}.weight"] for i in range(2)],
                    b_base=[sd[f"field.mlp_base.model.1.layers.{i}.bias"] for i in range(2)],
                    w_head=[sd[f"field.mlp_head.layers.{i}.weight"] for i in range(3)],
                    b_head=[sd[f"field.mlp_head.layers.{i}.bias"] for i in range(3)])

    P = dict(props=[field_params("proposal_networks.0", False), field_params("proposal_networks.1", False)],
             field=field_params("field", True))
    orays = dict(origins=rays["origins"], directions=rays["directions"], nears=torch.zeros(R, 1),
                 fars=torch.full((R, 1), 1000.0), camera_indices=rays["camera_indices"][:, 0], rgb=gt)
    ocfg = dict(num_prop_samples=(32, 20), num_nerf_samples=12, aabb=aabb, contraction=True, avg_init=0.01)
    ref = O.nerfacto_forward(P, orays, ocfg, {}, training=False)
    # eval mode uses the mean appearance embedding (use_average_appearance_embedding=True)
    err = (out["rgb"].cpu() - _eval_rgb(O, P, orays, ocfg)).abs().max().item()
    assert err < 1e-4, f"smoke: rgb differs from the oracle by {err:.2e}"
    acc_err = (out["accumulation"].cpu() - ref["accumulation"]).abs().max().item()
    assert acc_err < 1e-4, f"smoke: accumulation differs from the oracle by {acc_err:.2e}"

    # one optimisation step through backward + fused Adam
    from .nerfacto import Trainer

    model.train()
    tr = Trainer(model)
    stats = tr.train_iteration(bundle_from({k: v.cuda() for k, v in rays.items()}), {"image": gt.cuda()})
    torch.cuda.synchronize()
    assert torch.isfinite(stats["loss"]).item()

    # the captured step (hand-written backward, tcgen05 MLPs, TMA-staged weights, camera optimiser) on the same rays
    from .engine import NerfactoStep

    eng = NerfactoStep(model, R, use_graph=True)
    eng.set_batch(rays["origins"].cuda(), rays["directions"].cuda(), rays["camera_indices"].cuda(), gt.cuda())
    first = float(eng.step()[3])
    for _ in range(5):
        eng.step()
    torch.cuda.synchronize()
    last = float(eng.losses[3])
    assert first == first and last == last and last < first * 1.5, f"smoke: captured step diverged ({first} -> {last})"


def _eval_rgb(O, P, rays, cfg):
    import torch as _t

    out = O.nerfacto_forward(P, rays, cfg, {}, training=False)
    # nerfacto_forward(training=False) uses zeros for the appearance embedding; redo the field with the mean
    sb, eb = out["sdist_list"][-1], out["euclid_list"][-1]
    starts, ends = eb[:, :-1], eb[:, 1:]
    pos = O.frustum_positions(rays["origins"], rays["directions"], starts, ends)
    S = starts.shape[1]
    dens, rgb = O.nerfacto_field(pos, rays["directions"][:, None].expand(-1, S, -1), None, P["field"], cfg["aabb"], True,
                                 cfg["avg_init"], training=False, use_average_appearance=True)
    w = O.get_weights((ends - starts)[..., None], dens)
    return O.composite_rgb(rgb, w, "last_sample", training=False)
