"""Generate golden vectors from the ACTUAL reference (nerfstudio at /root/reference, pure-PyTorch path).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Each file holds the inputs, the weights and the outputs (and, where the
row is differentiable, the gradients for a recorded upstream gradient) of one SURVEY §8(a) row, produced
by the reference's own modules with `implementation="torch"`.  `viser` and `nerfacc` are stubbed at import
(they are import-time dependencies of the reference that are not installed; nothing on the torch path
calls into them).  Random draws inside the reference (stratified jitter) are made reproducible by
swapping `torch.rand` for a recording wrapper during the call; `torch.searchsorted` is wrapped the same
way to capture the int64 indices the reference computes internally.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("NERFSTUDIO_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
for _name in ("viser", "viser.transforms", "nerfacc"):
    _m = types.ModuleType(_name)
    _m.OccGridEstimator = object
    sys.modules[_name] = _m

import warnings  # noqa: E402

warnings.filterwarnings("ignore")

from nerfstudio.cameras.cameras import Cameras, CameraType  # noqa: E402
from nerfstudio.cameras.rays import Frustums, RayBundle, RaySamples  # noqa: E402
from nerfstudio.data.scene_box import SceneBox  # noqa: E402
from nerfstudio.field_components.encodings import HashEncoding, NeRFEncoding, SHEncoding  # noqa: E402
from nerfstudio.field_components.field_heads import FieldHeadNames  # noqa: E402
from nerfstudio.field_components.mlp import MLP  # noqa: E402
from nerfstudio.field_components.spatial_distortions import SceneContraction  # noqa: E402
from nerfstudio.fields.density_fields import HashMLPDensityField  # noqa: E402
from nerfstudio.fields.nerfacto_field import NerfactoField  # noqa: E402
from nerfstudio.fields.vanilla_nerf_field import NeRFField  # noqa: E402
from nerfstudio.model_components.losses import distortion_loss, interlevel_loss  # noqa: E402
from nerfstudio.model_components.ray_generators import RayGenerator  # noqa: E402
from nerfstudio.model_components.ray_samplers import (  # noqa: E402
    PDFSampler,
    ProposalNetworkSampler,
    UniformLinDispPiecewiseSampler,
    UniformSampler,
)
from nerfstudio.model_components.renderers import AccumulationRenderer, DepthRenderer, RGBRenderer  # noqa: E402
from nerfstudio.model_components.scene_colliders import AABBBoxCollider, NearFarCollider  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: npy(v) for k, v in arrs.items()})
    print(f"  wrote {name}.npz  ({', '.join(arrs)})")


class Recorder:
    """Swap torch.rand / torch.searchsorted for recording wrappers."""

    def __init__(self):
        self.rands, self.inds = [], []

    def __enter__(self):
        self._rand, self._ss = torch.rand, torch.searchsorted

        def rand(*a, **k):
            r = self._rand(*a, **k)
            self.rands.append(r.clone())
            return r

        def ss(*a, **k):
            r = self._ss(*a, **k)
            self.inds.append(r.clone())
            return r

        torch.rand, torch.searchsorted = rand, ss
        return self

    def __exit__(self, *exc):
        torch.rand, torch.searchsorted = self._rand, self._ss


def make_rays(R, seed, ring=True):
    g = torch.Generator().manual_seed(seed)
    if ring:
        ang = torch.rand(R, generator=g) * 2 * np.pi
        o = torch.stack([torch.cos(ang), torch.sin(ang), torch.zeros(R)], -1) + 0.3 * torch.randn(R, 3, generator=g)
        target = 0.3 * torch.randn(R, 3, generator=g)
        d = torch.nn.functional.normalize(target - o, dim=-1)
    else:
        o = torch.randn(R, 3, generator=g)
        d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    return o, d


def bundle(o, d, cam_hi=8, seed=0, near=0.05, far=1000.0):
    g = torch.Generator().manual_seed(seed + 77)
    R = o.shape[0]
    return RayBundle(
        origins=o.clone(), directions=d.clone(), pixel_area=torch.full((R, 1), 1e-6),
        camera_indices=torch.randint(0, cam_hi, (R, 1), generator=g),
        nears=torch.full((R, 1), near), fars=torch.full((R, 1), far),
    )


# ----------------------------------------------------------------------------------------
def g_hash():
    torch.manual_seed(1)
    cfgs = dict(
        small=dict(num_levels=4, min_res=16, max_res=128, log2_hashmap_size=10, features_per_level=2),
        f4=dict(num_levels=3, min_res=4, max_res=64, log2_hashmap_size=8, features_per_level=4),
        mid=dict(num_levels=8, min_res=16, max_res=1024, log2_hashmap_size=12, features_per_level=2),
    )
    out = {}
    for name, c in cfgs.items():
        enc = HashEncoding(**c, implementation="torch")
        N = 257
        x = torch.rand(N, 3)
        x[:8] = torch.tensor(  # exact grid hits / boundaries: ceil==floor cases
            [[0, 0, 0], [0.5, 0.5, 0.5], [0.25, 0.75, 0.125], [1 - 2 ** -24, 0.5, 0.0625], [0.0625] * 3, [0.999] * 3,
             [1e-7, 0.3, 0.6], [0.5, 0.0, 1 - 2 ** -24]], dtype=torch.float32)
        y = enc(x)
        dy = torch.randn_like(y)
        (gt,) = torch.autograd.grad(y, enc.hash_table, dy)
        # indices straight from the reference's hash_fn on the same corner selections
        scaled = x[..., None, :] * enc.scalings.view(-1, 1)
        c_, f_ = torch.ceil(scaled).int(), torch.floor(scaled).int()
        sel = [(c_, c_, c_), (c_, f_, c_), (f_, f_, c_), (f_, c_, c_), (c_, c_, f_), (c_, f_, f_), (f_, f_, f_), (f_, c_, f_)]
        idx = torch.stack(
            [enc.hash_fn(torch.cat([a[..., 0:1], b[..., 1:2], cc[..., 2:3]], -1)) for a, b, cc in sel], -1)
        out.update({f"{name}_x": x, f"{name}_table": enc.hash_table, f"{name}_scalings": enc.scalings,
                    f"{name}_y": y, f"{name}_dy": dy, f"{name}_dtable": gt, f"{name}_idx": idx,
                    f"{name}_cfg": np.array([c["num_levels"], c["min_res"], c["max_res"], c["log2_hashmap_size"],
                                             c["features_per_level"]])})
    # scalings of the BASELINE configs
    for nm, (L, lo, hi) in dict(main=(16, 16, 2048), prop0=(5, 16, 128), prop1=(5, 16, 256), default=(16, 16, 1024),
                                ngp=(16, 16, 2048)).items():
        out[f"scalings_{nm}"] = HashEncoding(num_levels=L, min_res=lo, max_res=hi, log2_hashmap_size=4,
                                            implementation="torch").scalings
    save("hash_encoding", **out)


def g_encodings():
    torch.manual_seed(2)
    d = torch.nn.functional.normalize(torch.randn(300, 3), dim=-1)
    out = {"dirs": d}
    for lv in range(1, 6):
        out[f"sh{lv}"] = SHEncoding(levels=lv, implementation="torch")((d + 1) / 2)
        out[f"sh{lv}_raw"] = SHEncoding(levels=lv, implementation="torch")(d)
    x = torch.rand(200, 3) * 4 - 2
    out["x"] = x
    pe = NeRFEncoding(in_dim=3, num_frequencies=10, min_freq_exp=0.0, max_freq_exp=8.0, include_input=True)
    de = NeRFEncoding(in_dim=3, num_frequencies=4, min_freq_exp=0.0, max_freq_exp=4.0, include_input=True)
    p2 = NeRFEncoding(in_dim=3, num_frequencies=2, min_freq_exp=0, max_freq_exp=1, include_input=False)
    xr = x.clone().requires_grad_(True)
    y = pe(xr)
    dy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, xr, dy)
    out.update(pe_y=y, pe_dy=dy, pe_dx=gx, de_y=de(x), p2_y=p2(x))
    pos = torch.randn(400, 3) * 2
    pos[:4] = torch.tensor([[0.5, 0.2, -0.1], [1.0, 0.0, 0.0], [3.0, -4.0, 0.5], [0.0, 0.0, 0.0]])
    out.update(contract_x=pos, contract_y=SceneContraction(order=float("inf"))(pos))
    save("encodings", **out)


def g_mlp():
    torch.manual_seed(3)
    out = {}
    cfgs = dict(
        base=dict(in_dim=32, num_layers=2, layer_width=64, out_dim=16, act=None),
        head=dict(in_dim=63, num_layers=3, layer_width=64, out_dim=3, act=torch.nn.Sigmoid()),
        prop=dict(in_dim=10, num_layers=2, layer_width=16, out_dim=1, act=None),
        skip=dict(in_dim=12, num_layers=6, layer_width=32, out_dim=None, act=torch.nn.ReLU(), skip=(3,)),
        one=dict(in_dim=7, num_layers=1, layer_width=8, out_dim=5, act=None),
    )
    for name, c in cfgs.items():
        m = MLP(in_dim=c["in_dim"], num_layers=c["num_layers"], layer_width=c["layer_width"], out_dim=c["out_dim"],
                skip_connections=c.get("skip"), out_activation=c["act"], implementation="torch")
        x = torch.randn(150, c["in_dim"], requires_grad=True)
        y = m(x)
        dy = torch.randn_like(y)
        grads = torch.autograd.grad(y, [x] + list(m.parameters()), dy)
        out.update({f"{name}_x": x, f"{name}_y": y, f"{name}_dy": dy, f"{name}_dx": grads[0]})
        for i, layer in enumerate(m.layers):
            out[f"{name}_w{i}"], out[f"{name}_b{i}"] = layer.weight, layer.bias
            out[f"{name}_dw{i}"], out[f"{name}_db{i}"] = grads[1 + 2 * i], grads[2 + 2 * i]
    save("mlp", **out)


def g_density_field():
    torch.manual_seed(4)
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    out = {"aabb": aabb}
    for name, contraction in (("contract", True), ("aabb", False)):
        f = HashMLPDensityField(aabb, num_layers=2, hidden_dim=16, num_levels=5, max_res=128, base_res=16,
                                log2_hashmap_size=12, features_per_level=2, average_init_density=0.01,
                                spatial_distortion=SceneContraction(order=float("inf")) if contraction else None,
                                implementation="torch")
        with torch.no_grad():  # make the network non-trivial
            f.encoding.hash_table.mul_(300.0)
        pos = torch.randn(40, 24, 3) * (2.0 if contraction else 0.7)
        dens = f.density_fn(pos)
        dy = torch.randn_like(dens)
        params = [f.encoding.hash_table] + [p for l in f.mlp_base[1].layers for p in (l.weight, l.bias)]
        grads = torch.autograd.grad(dens, params, dy)
        out.update({f"{name}_pos": pos, f"{name}_density": dens, f"{name}_dy": dy, f"{name}_table": params[0],
                    f"{name}_dtable": grads[0], f"{name}_scalings": f.encoding.scalings})
        for i in range(2):
            out[f"{name}_w{i}"], out[f"{name}_b{i}"] = params[1 + 2 * i], params[2 + 2 * i]
            out[f"{name}_dw{i}"], out[f"{name}_db{i}"] = grads[1 + 2 * i], grads[2 + 2 * i]
    save("density_field", **out)


def _samples_from_bins(rb, ebins, sbins=None):
    return rb.get_ray_samples(bin_starts=ebins[..., :-1, None], bin_ends=ebins[..., 1:, None],
                              spacing_starts=None if sbins is None else sbins[..., :-1, None],
                              spacing_ends=None if sbins is None else sbins[..., 1:, None])


def g_nerfacto_field():
    torch.manual_seed(5)
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    out = {"aabb": aabb}
    for name, contraction, training in (("train", True, True), ("eval_avg", True, False), ("aabb", False, True)):
        f = NerfactoField(aabb, num_images=8, num_levels=6, base_res=16, max_res=256, log2_hashmap_size=12,
                          spatial_distortion=SceneContraction(order=float("inf")) if contraction else None,
                          average_init_density=0.01, use_average_appearance_embedding=(name == "eval_avg"),
                          implementation="torch")
        f.train(training)
        with torch.no_grad():
            f.mlp_base.model[0].hash_table.mul_(300.0)
        o, d = make_rays(48, 11)
        rb = bundle(o, d)
        eb = torch.sort(torch.rand(48, 17) * 4.0 + 0.05, dim=-1).values
        rs = _samples_from_bins(rb, eb)
        fo = f(rs)
        dens, rgb = fo[FieldHeadNames.DENSITY], fo[FieldHeadNames.RGB]
        d_d, d_rgb = torch.randn_like(dens), torch.randn_like(rgb)
        named = dict(table=f.mlp_base.model[0].hash_table, emb=f.embedding_appearance.embedding.weight)
        for i, l in enumerate(f.mlp_base.model[1].layers):
            named[f"wb{i}"], named[f"bb{i}"] = l.weight, l.bias
        for i, l in enumerate(f.mlp_head.layers):
            named[f"wh{i}"], named[f"bh{i}"] = l.weight, l.bias
        keys = [k for k in named if not (k == "emb" and not training)]
        grads = torch.autograd.grad([dens, rgb], [named[k] for k in keys], [d_d, d_rgb], allow_unused=True)
        out.update({f"{name}_origins": o, f"{name}_directions": d, f"{name}_cams": rb.camera_indices,
                    f"{name}_ebins": eb, f"{name}_density": dens, f"{name}_rgb": rgb, f"{name}_d_density": d_d,
                    f"{name}_d_rgb": d_rgb, f"{name}_scalings": f.mlp_base.model[0].scalings})
        for k, v in named.items():
            out[f"{name}_{k}"] = v
        for k, gr in zip(keys, grads):
            out[f"{name}_g_{k}"] = gr if gr is not None else torch.zeros_like(named[k])
    save("nerfacto_field", **out)


def g_normals():
    """Field.get_normals (fields/base_field.py:80-102): -normalize(d density_before_activation / d sample_locations) through
    `forward(ray_samples, compute_normals=True)` — the position gradient of the hash grid + base MLP."""
    torch.manual_seed(17)
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    out = {"aabb": aabb}
    for name, contraction in (("contract", True), ("aabb", False)):
        f = NerfactoField(aabb, num_images=8, num_levels=6, base_res=16, max_res=256, log2_hashmap_size=12,
                          spatial_distortion=SceneContraction(order=float("inf")) if contraction else None,
                          average_init_density=0.01, implementation="torch")
        f.eval()
        with torch.no_grad():
            f.mlp_base.model[0].hash_table.mul_(300.0)
        o, d = make_rays(40, 23)
        rb = bundle(o, d)
        eb = torch.sort(torch.rand(40, 13) * 3.0 + 0.05, dim=-1).values
        fo = f(_samples_from_bins(rb, eb), compute_normals=True)
        raw = torch.autograd.grad(f._density_before_activation, f._sample_locations,
                                  grad_outputs=torch.ones_like(f._density_before_activation), retain_graph=True)[0]
        out.update({f"{name}_origins": o, f"{name}_directions": d, f"{name}_cams": rb.camera_indices, f"{name}_ebins": eb,
                    f"{name}_normals": fo[FieldHeadNames.NORMALS], f"{name}_density": fo[FieldHeadNames.DENSITY],
                    f"{name}_grad_raw": raw, f"{name}_table": f.mlp_base.model[0].hash_table,
                    f"{name}_emb": f.embedding_appearance.embedding.weight})
        for i, l in enumerate(f.mlp_base.model[1].layers):
            out[f"{name}_wb{i}"], out[f"{name}_bb{i}"] = l.weight, l.bias
        for i, l in enumerate(f.mlp_head.layers):
            out[f"{name}_wh{i}"], out[f"{name}_bh{i}"] = l.weight, l.bias
    save("normals", **out)


def g_samplers():
    torch.manual_seed(6)
    out = {}
    o, d = make_rays(64, 21)
    near = torch.full((64, 1), 0.05)
    far = torch.full((64, 1), 1000.0)
    far[::3] = 6.0
    near[::5] = 2.0
    rb = bundle(o, d)
    rb.nears, rb.fars = near, far
    out.update(nears=near, fars=far)
    for kind, cls in (("piecewise", UniformLinDispPiecewiseSampler), ("uniform", UniformSampler)):
        for mode in ("eval", "single", "multi"):
            s = cls(num_samples=32, single_jitter=(mode == "single"))
            s.train(mode != "eval")
            with Recorder() as rec:
                rs = s(rb)
            out[f"{kind}_{mode}_sbins"] = torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[:, -1:, 0]], -1)
            out[f"{kind}_{mode}_ebins"] = torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[:, -1:, 0]], -1)
            if rec.rands:
                out[f"{kind}_{mode}_jitter"] = rec.rands[0]
    # PDF sampler on top of the piecewise eval samples
    init = UniformLinDispPiecewiseSampler(num_samples=32)
    init.eval()
    rs0 = init(rb)
    w = torch.rand(64, 32, 1) ** 4
    w[3] = 0.0  # zero-weight ray
    w[4, 5:] = 0.0
    w[5] = 1e-9
    out["pdf_weights"] = w
    for mode, inc in (("eval", False), ("single", False), ("multi", False), ("eval_inc", True)):
        s = PDFSampler(num_samples=16, single_jitter=(mode == "single"), include_original=inc,
                       train_stratified=True)
        s.train(mode in ("single", "multi"))
        with Recorder() as rec:
            rs = s(rb, rs0, w)
        out[f"pdf_{mode}_sbins"] = torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[:, -1:, 0]], -1)
        out[f"pdf_{mode}_ebins"] = torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[:, -1:, 0]], -1)
        out[f"pdf_{mode}_inds"] = rec.inds[0]
        if rec.rands:
            out[f"pdf_{mode}_jitter"] = rec.rands[0]
    # colliders
    box = SceneBox(aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]]))
    col = AABBBoxCollider(box, near_plane=0.1)
    col.train()
    rb2 = RayBundle(origins=o.clone(), directions=d.clone(), pixel_area=torch.ones(64, 1))
    rb2 = col(rb2)
    out.update(origins=o, directions=d, aabb_nears=rb2.nears, aabb_fars=rb2.fars)
    nf = NearFarCollider(0.05, 1000.0)
    nf.eval()
    rb3 = nf(RayBundle(origins=o.clone(), directions=d.clone(), pixel_area=torch.ones(64, 1)))
    out.update(nf_eval_nears=rb3.nears, nf_eval_fars=rb3.fars)
    save("samplers", **out)


def g_render():
    torch.manual_seed(7)
    out = {}
    R, S = 50, 24
    eb = torch.sort(torch.rand(R, S + 1) * 5 + 0.1, dim=-1).values
    sb = torch.sort(torch.rand(R, S + 1), dim=-1).values
    o, d = make_rays(R, 31)
    rb = bundle(o, d)
    rs = _samples_from_bins(rb, eb, sb)
    dens = (torch.rand(R, S, 1) ** 3 * 30).requires_grad_(True)
    with torch.no_grad():
        dens[2] = 0.0
        dens[3, 4] = 1e9
    w = rs.get_weights(dens)
    dw = torch.randn_like(w)
    (gd,) = torch.autograd.grad(w, dens, dw, retain_graph=True)
    out.update(ebins=eb, sbins=sb, density=dens, weights=w, d_weights=dw, d_density=gd)
    rgb = torch.rand(R, S, 3, requires_grad=True)
    for bg in ("last_sample", "white", "black", "random"):
        ren = RGBRenderer(background_color=bg)
        ren.train()
        wd = w.detach().clone().requires_grad_(True)
        comp = ren(rgb=rgb, weights=wd)
        dc = torch.randn_like(comp)
        g_rgb, g_w = torch.autograd.grad(comp, [rgb, wd], dc)
        out.update({f"rgb_{bg}": comp, f"rgb_{bg}_dout": dc, f"rgb_{bg}_drgb": g_rgb, f"rgb_{bg}_dw": g_w})
    ren = RGBRenderer(background_color="last_sample")
    ren.eval()
    rgb_nan = rgb.detach().clone()
    rgb_nan[1, 2, 0] = float("nan")
    out.update(rgb_samples=rgb, rgb_samples_nan=rgb_nan, rgb_eval=ren(rgb=rgb_nan, weights=w.detach()))
    out["accumulation"] = AccumulationRenderer()(weights=w)
    out["depth_median"] = DepthRenderer("median")(weights=w.detach(), ray_samples=rs)
    wd = w.detach().clone().requires_grad_(True)
    de = DepthRenderer("expected")(weights=wd, ray_samples=rs)
    dde = torch.randn_like(de)
    (g_w,) = torch.autograd.grad(de, wd, dde)
    out.update(depth_expected=de, depth_expected_dout=dde, depth_expected_dw=g_w)
    save("render", **out)


def g_losses():
    torch.manual_seed(8)
    R = 40
    o, d = make_rays(R, 41)
    rb = bundle(o, d)

    def lvl(S):
        sb = torch.sort(torch.rand(R, S + 1), dim=-1).values
        sb[:, 0], sb[:, -1] = 0.0, 1.0
        w = torch.rand(R, S, 1) ** 2
        w = (w / w.sum(1, keepdim=True) * torch.rand(R, 1, 1)).requires_grad_(True)
        return _samples_from_bins(rb, sb * 5 + 0.05, sb), w, sb

    rs0, w0, sb0 = lvl(32)
    rs1, w1, sb1 = lvl(20)
    rs2, w2, sb2 = lvl(12)
    li = interlevel_loss([w0, w1, w2], [rs0, rs1, rs2])
    g0, g1 = torch.autograd.grad(li, [w0, w1])
    ld = distortion_loss([w0, w1, w2], [rs0, rs1, rs2])
    (g2,) = torch.autograd.grad(ld, [w2])
    save("losses", sb0=sb0, sb1=sb1, sb2=sb2, w0=w0, w1=w1, w2=w2, interlevel=li, interlevel_dw0=g0,
         interlevel_dw1=g1, distortion=ld, distortion_dw2=g2)


def g_raygen():
    torch.manual_seed(9)
    C = 5
    c2w = torch.eye(4)[None, :3, :].repeat(C, 1, 1)
    rot = torch.linalg.qr(torch.randn(C, 3, 3)).Q
    c2w[:, :3, :3] = rot
    c2w[:, :3, 3] = torch.randn(C, 3)
    fx = torch.tensor([1111.1, 900.0, 1200.0, 500.0, 800.0])
    fy = torch.tensor([1111.1, 905.0, 1190.0, 500.0, 790.0])
    cx = torch.tensor([400.0, 320.0, 960.0, 25.0, 400.5])
    cy = torch.tensor([400.0, 240.0, 540.0, 25.0, 300.25])
    hw = torch.tensor([[800, 800], [480, 640], [1080, 1920], [50, 50], [600, 800]])
    dist = torch.zeros(C, 6)
    dist[1] = torch.tensor([0.05, -0.01, 0.001, 0.0, 0.002, -0.001])
    dist[2] = torch.tensor([-0.1, 0.02, 0.0, 0.0, 0.0, 0.0])
    out = dict(c2w=c2w, fx=fx, fy=fy, cx=cx, cy=cy, hw=hw, dist=dist)
    for name, dd in (("nodist", None), ("dist", dist)):
        cams = Cameras(camera_to_worlds=c2w, fx=fx, fy=fy, cx=cx, cy=cy, height=hw[:, 0:1], width=hw[:, 1:2],
                       distortion_params=dd, camera_type=CameraType.PERSPECTIVE)
        gen = RayGenerator(cams)
        cam = torch.randint(0, C, (200,))
        row = (torch.rand(200) * hw[cam, 0]).long()
        col = (torch.rand(200) * hw[cam, 1]).long()
        ri = torch.stack([cam, row, col], -1)
        rb = gen(ri)
        out.update({f"{name}_ray_indices": ri, f"{name}_origins": rb.origins, f"{name}_directions": rb.directions,
                    f"{name}_pixel_area": rb.pixel_area, f"{name}_camera_indices": rb.camera_indices,
                    f"{name}_directions_norm": rb.metadata["directions_norm"]})
    save("raygen", **out)


def g_vanilla():
    torch.manual_seed(10)
    pe = NeRFEncoding(in_dim=3, num_frequencies=10, min_freq_exp=0.0, max_freq_exp=8.0, include_input=True)
    de = NeRFEncoding(in_dim=3, num_frequencies=4, min_freq_exp=0.0, max_freq_exp=4.0, include_input=True)
    f = NeRFField(position_encoding=pe, direction_encoding=de, base_mlp_num_layers=8, base_mlp_layer_width=64,
                  head_mlp_num_layers=2, head_mlp_layer_width=32)
    o, d = make_rays(20, 51)
    rb = bundle(o, d, near=2.0, far=6.0)
    s = UniformSampler(num_samples=12)
    s.eval()
    rs = s(rb)
    fo = f(rs)
    out = dict(origins=o, directions=d, ebins=torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[:, -1:, 0]], -1),
               density=fo[FieldHeadNames.DENSITY], rgb=fo[FieldHeadNames.RGB])
    for i, l in enumerate(f.mlp_base.layers):
        out[f"wb{i}"], out[f"bb{i}"] = l.weight, l.bias
    for i, l in enumerate(f.mlp_head.layers):
        out[f"wh{i}"], out[f"bh{i}"] = l.weight, l.bias
    out["w_sigma"], out["b_sigma"] = f.field_output_density.net.weight, f.field_output_density.net.bias
    out["w_rgb"], out["b_rgb"] = f.field_heads[0].net.weight, f.field_heads[0].net.bias
    save("vanilla_field", **out)


def g_pipeline():
    """Full nerfacto composition: ProposalNetworkSampler + NerfactoField + renderers + losses (+ grads)."""
    torch.manual_seed(12)
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    con = SceneContraction(order=float("inf"))
    props = torch.nn.ModuleList([
        HashMLPDensityField(aabb, hidden_dim=16, num_levels=5, max_res=mr, base_res=16, log2_hashmap_size=12,
                            spatial_distortion=con, average_init_density=0.01, implementation="torch")
        for mr in (128, 256)])
    field = NerfactoField(aabb, num_images=8, num_levels=8, base_res=16, max_res=512, log2_hashmap_size=13,
                          spatial_distortion=con, average_init_density=0.01, implementation="torch")
    with torch.no_grad():
        for p in props:
            p.encoding.hash_table.mul_(2000.0)
        field.mlp_base.model[0].hash_table.mul_(1000.0)
    R = 96
    o, d = make_rays(R, 61)
    gt = torch.rand(R, 3)
    out = dict(aabb=aabb, origins=o, directions=d, gt=gt)
    for mode in ("train", "eval"):
        training = mode == "train"
        props.train(training), field.train(training)
        sampler = ProposalNetworkSampler(num_nerf_samples_per_ray=12, num_proposal_samples_per_ray=(32, 20),
                                         num_proposal_network_iterations=2, single_jitter=True)
        sampler.train(training)
        sampler.set_anneal(0.7 if training else 1.0)
        rb = bundle(o, d, seed=3)
        with Recorder() as rec:
            rs, wl, rsl = sampler(rb, density_fns=[p.density_fn for p in props])
        fo = field(rs)
        w = rs.get_weights(fo[FieldHeadNames.DENSITY])
        wl.append(w), rsl.append(rs)
        ren = RGBRenderer(background_color="last_sample")
        ren.train(training)
        rgb = ren(rgb=fo[FieldHeadNames.RGB], weights=w)
        acc = AccumulationRenderer()(weights=w)
        dep = DepthRenderer("median")(weights=w.detach(), ray_samples=rs)
        edep = DepthRenderer("expected")(weights=w, ray_samples=rs)
        out.update({f"{mode}_rgb": rgb, f"{mode}_acc": acc, f"{mode}_depth": dep, f"{mode}_exp_depth": edep,
                    f"{mode}_cams": rb.camera_indices})
        for i, (ww, rr) in enumerate(zip(wl, rsl)):
            out[f"{mode}_w{i}"] = ww
            out[f"{mode}_sbins{i}"] = torch.cat([rr.spacing_starts[..., 0], rr.spacing_ends[:, -1:, 0]], -1)
            out[f"{mode}_ebins{i}"] = torch.cat([rr.frustums.starts[..., 0], rr.frustums.ends[:, -1:, 0]], -1)
        for i, r_ in enumerate(rec.rands):
            out[f"{mode}_rand{i}"] = r_
        for i, ind in enumerate(rec.inds):
            out[f"{mode}_inds{i}"] = ind
        if training:
            l_rgb = torch.nn.functional.mse_loss(gt, rgb)
            l_il = interlevel_loss(wl, rsl)
            l_di = 0.002 * distortion_loss(wl, rsl)
            loss = l_rgb + l_il + l_di
            named = {}
            for j, p in enumerate(props):
                named[f"p{j}_table"] = p.encoding.hash_table
                for i, l in enumerate(p.mlp_base[1].layers):
                    named[f"p{j}_w{i}"], named[f"p{j}_b{i}"] = l.weight, l.bias
            named["f_table"] = field.mlp_base.model[0].hash_table
            named["f_emb"] = field.embedding_appearance.embedding.weight
            for i, l in enumerate(field.mlp_base.model[1].layers):
                named[f"f_wb{i}"], named[f"f_bb{i}"] = l.weight, l.bias
            for i, l in enumerate(field.mlp_head.layers):
                named[f"f_wh{i}"], named[f"f_bh{i}"] = l.weight, l.bias
            grads = torch.autograd.grad(loss, list(named.values()))
            out.update(loss=loss, loss_rgb=l_rgb, loss_interlevel=l_il, loss_distortion=l_di)
            for (k, v), gr in zip(named.items(), grads):
                out[k], out["g_" + k] = v, gr
            out["p0_scalings"], out["p1_scalings"] = props[0].encoding.scalings, props[1].encoding.scalings
            out["f_scalings"] = field.mlp_base.model[0].scalings
    save("nerfacto_pipeline", **out)


def g_pipeline_camopt():
    """nerfacto composition WITH the camera optimiser on (its default: models/nerfacto.py:131, get_outputs :300-301): the
    SO3xR3 pose corrections move origins / directions, and the photometric + interlevel + distortion losses reach
    `pose_adjustment` through the sample positions of all three levels (Frustums.get_positions, rays.py:50-59)."""
    from nerfstudio.cameras.camera_optimizers import CameraOptimizer, CameraOptimizerConfig

    torch.manual_seed(29)
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    con = SceneContraction(order=float("inf"))
    props = torch.nn.ModuleList([
        HashMLPDensityField(aabb, hidden_dim=16, num_levels=5, max_res=mr, base_res=16, log2_hashmap_size=12,
                            spatial_distortion=con, average_init_density=0.01, implementation="torch")
        for mr in (128, 256)])
    field = NerfactoField(aabb, num_images=8, num_levels=8, base_res=16, max_res=512, log2_hashmap_size=13,
                          spatial_distortion=con, average_init_density=0.01, implementation="torch")
    with torch.no_grad():
        for p in props:
            p.encoding.hash_table.mul_(2000.0)
        field.mlp_base.model[0].hash_table.mul_(1000.0)
    C, R = 8, 96
    opt = CameraOptimizer(CameraOptimizerConfig(mode="SO3xR3"), num_cameras=C, device="cpu")
    with torch.no_grad():
        opt.pose_adjustment.copy_(torch.randn(C, 6) * 0.02)
    o, d = make_rays(R, 67)
    gt = torch.rand(R, 3)
    props.train(), field.train()
    sampler = ProposalNetworkSampler(num_nerf_samples_per_ray=12, num_proposal_samples_per_ray=(32, 20),
                                     num_proposal_network_iterations=2, single_jitter=True)
    sampler.train()
    sampler.set_anneal(0.7)
    rb = bundle(o, d, cam_hi=C, seed=5)
    opt.apply_to_raybundle(rb)
    with Recorder() as rec:
        rs, wl, rsl = sampler(rb, density_fns=[p.density_fn for p in props])
    fo = field(rs)
    w = rs.get_weights(fo[FieldHeadNames.DENSITY])
    wl.append(w), rsl.append(rs)
    ren = RGBRenderer(background_color="last_sample")
    ren.train()
    rgb = ren(rgb=fo[FieldHeadNames.RGB], weights=w)
    l_rgb = torch.nn.functional.mse_loss(gt, rgb)
    l_il = interlevel_loss(wl, rsl)
    l_di = 0.002 * distortion_loss(wl, rsl)
    reg = {}
    opt.get_loss_dict(reg)
    loss = l_rgb + l_il + l_di + reg["camera_opt_regularizer"]
    out = dict(aabb=aabb, origins=o, directions=d, gt=gt, cams=rb.camera_indices, pose=opt.pose_adjustment, rgb=rgb,
               loss=loss, loss_rgb=l_rgb, loss_interlevel=l_il, loss_distortion=l_di, regularizer=reg["camera_opt_regularizer"])
    parts = dict(rgb=l_rgb, interlevel=l_il, distortion=l_di, total=loss)
    for k, v in parts.items():
        (gp,) = torch.autograd.grad(v, [opt.pose_adjustment], retain_graph=True)
        out["g_pose_" + k] = gp
    for i, rr in enumerate(rsl):
        out[f"sbins{i}"] = torch.cat([rr.spacing_starts[..., 0], rr.spacing_ends[:, -1:, 0]], -1)
        out[f"ebins{i}"] = torch.cat([rr.frustums.starts[..., 0], rr.frustums.ends[:, -1:, 0]], -1)
    for i, r_ in enumerate(rec.rands):
        out[f"rand{i}"] = r_
    for j, p in enumerate(props):
        out[f"p{j}_table"] = p.encoding.hash_table
        for i, l in enumerate(p.mlp_base[1].layers):
            out[f"p{j}_w{i}"], out[f"p{j}_b{i}"] = l.weight, l.bias
    out["f_table"], out["f_emb"] = field.mlp_base.model[0].hash_table, field.embedding_appearance.embedding.weight
    for i, l in enumerate(field.mlp_base.model[1].layers):
        out[f"f_wb{i}"], out[f"f_bb{i}"] = l.weight, l.bias
    for i, l in enumerate(field.mlp_head.layers):
        out[f"f_wh{i}"], out[f"f_bh{i}"] = l.weight, l.bias
    save("pipeline_camopt", **out)


def g_samplers_extra():
    """The other SpacedSampler subclasses the reference tests (tests/model_components/test_ray_sampler.py:37-86):
    LinearDisparitySampler, SqrtSampler, LogSampler — that test's set-up (10 rays, near 2 / far 4, 15 samples) plus
    64 rays with varied near/far, eval and single-jitter training mode."""
    from nerfstudio.model_components.ray_samplers import LinearDisparitySampler, LogSampler, SqrtSampler

    torch.manual_seed(8)
    out = {}
    o, d = make_rays(64, 22)
    near, far = torch.full((64, 1), 0.05), torch.full((64, 1), 1000.0)
    far[::3], near[::5] = 6.0, 2.0
    rb = bundle(o, d)
    rb.nears, rb.fars = near, far
    rb_t = RayBundle(origins=torch.zeros(10, 3), directions=torch.ones(10, 3), pixel_area=torch.ones(10, 1))
    col = NearFarCollider(near_plane=2, far_plane=4)
    col.train()
    rb_t = col(rb_t)
    out.update(nears=near, fars=far, t_nears=rb_t.nears, t_fars=rb_t.fars)
    for kind, cls in (("lindisp", LinearDisparitySampler), ("sqrt", SqrtSampler), ("log", LogSampler)):
        s = cls(num_samples=15)
        s.eval()
        rs = s(rb_t)
        assert rs.frustums.get_positions().shape[-2] == 15
        out[f"{kind}_t_ebins"] = torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[:, -1:, 0]], -1)
        for mode in ("eval", "single"):
            s = cls(num_samples=24, single_jitter=True)
            s.train(mode != "eval")
            with Recorder() as rec:
                rs = s(rb)
            out[f"{kind}_{mode}_sbins"] = torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[:, -1:, 0]], -1)
            out[f"{kind}_{mode}_ebins"] = torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[:, -1:, 0]], -1)
            if rec.rands:
                out[f"{kind}_{mode}_jitter"] = rec.rands[0]
    save("samplers_extra", **out)


def g_aabb_intersect():
    """nerfstudio.utils.math.intersect_aabb — the function the reference itself compares with
    nerfacc.ray_aabb_intersect (tests/utils/test_aabb_intersection.py:117-183, rtol 1e-3): the only pinned statement
    about the nerfacc call used by the packed path."""
    from nerfstudio.utils.math import intersect_aabb

    g = torch.Generator().manual_seed(496)
    lo = (torch.rand(3, generator=g) - 0.5) * 100
    aabb = torch.cat([lo, lo + torch.rand(3, generator=g) * 100 + 1.0])
    centre, half = (aabb[:3] + aabb[3:]) / 2, (aabb[3:] - aabb[:3]) / 2
    R = 500
    o = centre + (torch.rand(R, 3, generator=g) - 0.5) * 1000
    target = centre + (torch.rand(R, 3, generator=g) - 0.5) * 2 * half * 1.5   # some rays miss the box
    d = torch.nn.functional.normalize(target - o, dim=-1)
    o[:50] = centre + (torch.rand(50, 3, generator=g) - 0.5) * half               # origins inside the box
    t_min, t_max = intersect_aabb(o, d, aabb)
    save("aabb_intersect", aabb=aabb, origins=o, directions=d, t_min=t_min, t_max=t_max)


def g_camera_opt():
    """CameraOptimizer.apply_to_raybundle (SO3xR3), with the pose gradients and the regulariser."""
    from nerfstudio.cameras.camera_optimizers import CameraOptimizer, CameraOptimizerConfig

    torch.manual_seed(11)
    C, R = 8, 512
    opt = CameraOptimizer(CameraOptimizerConfig(mode="SO3xR3"), num_cameras=C, device="cpu")
    with torch.no_grad():
        opt.pose_adjustment.copy_(torch.randn(C, 6) * 0.05)
        opt.pose_adjustment[0].zero_()               # the initial state: |w|^2 clamped at 1e-4
        opt.pose_adjustment[1, 3:] *= 0.01            # below the clamp
        opt.pose_adjustment[2, 3:] *= 20.0            # a large rotation (~1 rad)
    o, d = make_rays(R, 5)
    rb = bundle(o, d, cam_hi=C, seed=3)
    opt.apply_to_raybundle(rb)
    go, gd = torch.randn(R, 3), torch.randn(R, 3)
    (g_pose,) = torch.autograd.grad((rb.origins * go).sum() + (rb.directions * gd).sum(), [opt.pose_adjustment])
    loss = {}
    opt.get_loss_dict(loss)
    mats = opt(torch.arange(C))
    save("camera_opt", pose=opt.pose_adjustment, cams=rb.camera_indices, origins=o, directions=d, out_origins=rb.origins,
         out_directions=rb.directions, go=go, gd=gd, g_pose=g_pose, regularizer=loss["camera_opt_regularizer"], matrices=mats)


if __name__ == "__main__":
    torch.set_num_threads(4)
    fns = (g_hash, g_encodings, g_mlp, g_density_field, g_nerfacto_field, g_samplers, g_render, g_losses,
           g_raygen, g_vanilla, g_pipeline, g_camera_opt, g_samplers_extra, g_aabb_intersect, g_normals, g_pipeline_camopt)
    only = set(sys.argv[1:])  # e.g. `python tests/golden/make_golden.py g_camera_opt` regenerates one fixture
    for fn in fns:
        if only and fn.__name__ not in only:
            continue
        print(fn.__name__)
        fn()
