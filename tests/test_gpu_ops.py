"""GPU parity: every C-ABI kernel vs the golden vectors recorded from the reference and vs the oracle.

Bars (north_star): bit-exact for integer / index outputs, <= 1e-4 relative (max-norm) for fp32 — most ops are
held to a tighter 2e-5.  All calls go through nerfstudio_b200.functional -> ctypes -> libb200nerf.so.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-4
TIGHT = 2e-5


@pytest.fixture(scope="module")
def F():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nerfstudio_b200 import functional

    return functional


def cu(t):
    return t.cuda() if isinstance(t, torch.Tensor) else t


# ------------------------------------------------------------------------------------------------
def test_hashgrid_golden(F, golden):
    g = golden("hash_encoding")
    for nm in ("small", "f4", "mid"):
        L, lo, hi, log2T, Fd = (int(v) for v in g[f"{nm}_cfg"])
        grid = F.GridSpec(g[f"{nm}_scalings"].tolist(), log2T, Fd)
        x, table = cu(g[f"{nm}_x"]), cu(g[f"{nm}_table"])
        y, idx = F.hashgrid_forward(x, table, grid, want_indices=True)
        assert torch.equal(idx.cpu(), g[f"{nm}_idx"]), f"{nm}: corner indices must be bit-exact"
        assert_close(y, g[f"{nm}_y"], TIGHT, nm)
        dtable, dx = F.hashgrid_backward(x, table, cu(g[f"{nm}_dy"]), grid, want_dx=True)
        assert_close(dtable, g[f"{nm}_dtable"], TIGHT, nm + " dtable")
        # dx against autograd of the oracle
        xr = g[f"{nm}_x"].clone().requires_grad_(True)
        yo = O.hash_encode(xr, g[f"{nm}_table"], g[f"{nm}_scalings"], log2T)
        (gx,) = torch.autograd.grad(yo, xr, g[f"{nm}_dy"])
        assert_close(dx, gx, REL, nm + " dx")


def test_hashgrid_full_size_properties(F):
    """BASELINE size (L=16, T=2^19, N=196608): indices vs oracle on a slice, linearity in the table, and
    scatter/gather adjointness  <dy, enc(table)> == <dtable, table>."""
    torch.manual_seed(0)
    scal = O.hash_level_scalings(16, 16, 2048)
    grid = F.GridSpec(scal.tolist(), 19, 2)
    N = 196608
    x = torch.rand(N, 3, device="cuda")
    t1 = (torch.rand(16 << 19, 2, device="cuda") * 2 - 1) * 1e-3
    t2 = torch.randn_like(t1) * 1e-3
    y1, idx = F.hashgrid_forward(x, t1, grid, want_indices=True)
    idx_o, _ = O.hash_corner_indices(x[:4096].cpu(), scal, 19)
    assert torch.equal(idx[:4096].cpu(), idx_o)
    y2 = F.hashgrid_forward(x, t2, grid)
    y12 = F.hashgrid_forward(x, t1 + 2 * t2, grid)
    assert_close(y12, y1 + 2 * y2, 1e-5, "linearity")
    dy = torch.randn_like(y1)
    dt, _ = F.hashgrid_backward(x, t1, dy, grid)
    lhs = (dy.double() * y1.double()).sum()
    rhs = (dt.double() * t1.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-4 * abs(float(lhs)) + 1e-9
    yo = O.hash_encode(x[:2048].cpu(), t1.cpu(), scal, 19)
    assert_close(y1[:2048], yo, TIGHT)


def test_hashgrid_pair_kernel_bitwise_equals_thread_per_point(F):
    """The lane-pair gather (default for F = 2) forms every output from the same 8 corners in the same association as the
    one-thread-per-point kernel: bit-identical, at the BASELINE size, odd point counts, 5- and 16-level grids, both modes."""
    from nerfstudio_b200 import lib

    torch.manual_seed(2)
    cases = [(F.GridSpec(O.hash_level_scalings(16, 16, 2048).tolist(), 19, 2), 16 << 19, 196608),
             (F.GridSpec(O.hash_level_scalings(5, 16, 128).tolist(), 17, 2), 5 << 17, 100001),
             (F.GridSpec(O.hash_level_scalings(7, 16, 512).tolist(), 12, 2), 7 << 12, 1)]
    g = F.GridSpec.tcnn(8, 16, 1.3819, 12, 2)
    cases.append((g, g.n_rows, 3333))
    for grid, rows, N in cases:
        x = torch.rand(N, 3, device="cuda")
        x[: min(N, 7)] = torch.tensor([0.0, 0.5, 1.0])[:3]  # exact grid points: ceil == floor
        table = torch.randn(rows, 2, device="cuda")
        try:
            assert lib.tune("hash_fwd_pair", 0)
            ref = F.hashgrid_forward(x, table, grid)
        finally:
            lib.tune("hash_fwd_pair", 1)
        assert torch.equal(F.hashgrid_forward(x, table, grid), ref), (grid.n_levels, N)


def test_hashgrid_dx_kernels(F):
    """b2n_hashgrid_dx (gather-only position gradient; lane-pair and thread-per-point variants) == the dx of hashgrid_bwd."""
    from nerfstudio_b200 import lib
    from nerfstudio_b200.lib import call, ptr, stream
    import ctypes as C

    torch.manual_seed(4)
    for grid, rows, N in ((F.GridSpec(O.hash_level_scalings(16, 16, 2048).tolist(), 19, 2), 16 << 19, 50001),
                          (F.GridSpec.tcnn(8, 16, 1.3819, 12, 2), None, 777)):
        rows = rows or grid.n_rows
        x = torch.rand(N, 3, device="cuda")
        table = torch.randn(rows, 2, device="cuda")
        dy = torch.randn(N, grid.out_dim, device="cuda")
        _, ref = F.hashgrid_backward(x, table, dy, grid, want_dx=True)
        for pair in (1, 0):
            try:
                assert lib.tune("hash_fwd_pair", pair)
                dx = torch.empty(N, 3, device="cuda")
                call("b2n_hashgrid_dx", C.byref(grid.c), ptr(x), ptr(table), ptr(dy), N, ptr(dx), stream())
            finally:
                lib.tune("hash_fwd_pair", 1)
            assert_close(dx, ref, 1e-5, f"dx pair={pair}")


def test_hashgrid_tcnn_mode(F):
    """tcnn-mode addressing vs the oracle's restatement of the published semantics (parity unpinned upstream)."""
    torch.manual_seed(1)
    g = 1.3819
    meta, rows = O.tcnn_grid_meta(8, 16, g, 12)
    grid = F.GridSpec.tcnn(8, 16, g, 12, 2)
    assert grid.n_rows == rows
    x = torch.rand(500, 3)
    table = torch.randn(rows, 2) * 0.1
    y = F.hashgrid_forward(x.cuda(), table.cuda(), grid)
    assert_close(y, O.tcnn_hash_encode(x, table, meta), TIGHT)


def test_hashgrid_empty_and_errors(F):
    grid = F.GridSpec([16.0, 32.0], 8, 2)
    y = F.hashgrid_forward(torch.empty(0, 3, device="cuda"), torch.zeros(512, 2, device="cuda"), grid)
    assert y.shape == (0, 4)
    with pytest.raises(ValueError):
        F.GridSpec([16.0], 8, 3)
    with pytest.raises(RuntimeError):
        F.hashgrid_forward(torch.zeros(4, 3), torch.zeros(512, 2), grid)  # CPU tensors: no fallback


# ------------------------------------------------------------------------------------------------
def _mlp_cfg(nm):
    return dict(base=(2, (), "none"), head=(3, (), "sigmoid"), prop=(2, (), "none"), skip=(6, (3,), "relu"),
                one=(1, (), "none"))[nm]


def test_mlp_golden(F, golden):
    g = golden("mlp")
    for nm in ("base", "head", "prop", "skip", "one"):
        n, skip, oact = _mlp_cfg(nm)
        ws = [cu(g[f"{nm}_w{i}"]).requires_grad_(True) for i in range(n)]
        bs = [cu(g[f"{nm}_b{i}"]).requires_grad_(True) for i in range(n)]
        x = cu(g[f"{nm}_x"]).requires_grad_(True)
        spec = F.MlpSpec(x.shape[1], [w.shape[0] for w in ws], skip=skip, out_act=oact)
        y = F.mlp(spec, x, ws, bs)
        assert_close(y, g[f"{nm}_y"], TIGHT, nm)
        grads = torch.autograd.grad(y, [x] + ws + bs, cu(g[f"{nm}_dy"]))
        assert_close(grads[0], g[f"{nm}_dx"], TIGHT, nm + " dx")
        for i in range(n):
            assert_close(grads[1 + i], g[f"{nm}_dw{i}"], TIGHT, f"{nm} dw{i}")
            assert_close(grads[1 + n + i], g[f"{nm}_db{i}"], TIGHT, f"{nm} db{i}")


def test_mlp_large_batch_vs_oracle(F):
    """Ragged batch (not a multiple of the 128-row tile), many tiles per CTA; no-bias (tcnn-style) variant too."""
    torch.manual_seed(3)
    for bias in (True, False):
        n = 128 * 700 + 37
        x = torch.randn(n, 63)
        ws = [torch.randn(64, 63) * 0.2, torch.randn(64, 64) * 0.2, torch.randn(3, 64) * 0.2]
        bs = [torch.randn(64) * 0.1, torch.randn(64) * 0.1, torch.randn(3) * 0.1] if bias else [None] * 3
        dy = torch.randn(n, 3)
        wl = [w.clone().requires_grad_(True) for w in ws]
        bl = [b.clone().requires_grad_(True) if b is not None else None for b in bs]
        yo = O.mlp_forward(x, wl, bl, out_act="sigmoid")
        go = torch.autograd.grad(yo, wl + [b for b in bl if b is not None], dy)
        wc = [w.cuda().requires_grad_(True) for w in ws]
        bc = [b.cuda().requires_grad_(True) if b is not None else None for b in bs]
        spec = F.MlpSpec(63, [64, 64, 3], out_act="sigmoid", bias=bias)
        y = F.mlp(spec, x.cuda(), wc, bc)
        assert_close(y, yo, TIGHT)
        gc = torch.autograd.grad(y, wc + [b for b in bc if b is not None], dy.cuda())
        for a, b in zip(gc, go):
            assert_close(a, b, REL)


# ------------------------------------------------------------------------------------------------
def test_encodings_golden(F, golden):
    g = golden("encodings")
    d = cu(g["dirs"])
    for lv in range(1, 6):
        assert_close(F.sh_encode(d, lv, remap01=True), g[f"sh{lv}"], TIGHT, f"sh{lv}")
        assert_close(F.sh_encode(d, lv, remap01=False), g[f"sh{lv}_raw"], TIGHT)
    with pytest.raises(ValueError):
        F.sh_encode(d, 6)
    x = cu(g["x"]).requires_grad_(True)
    freqs = (2 ** torch.linspace(0.0, 8.0, 10)).tolist()
    y = F.freq_encode(x, freqs, True)
    # arguments up to ~3.2e3: (2*pi*x)*freq (+pi/2) is formed in the reference's rounding order (encodings.py:162-169)
    assert_close(y, g["pe_y"], TIGHT, "nerf enc")
    (gx,) = torch.autograd.grad(y, x, cu(g["pe_dy"]))
    assert_close(gx, g["pe_dx"], REL)
    assert_close(F.freq_encode(cu(g["x"]), (2 ** torch.linspace(0.0, 4.0, 4)).tolist(), True), g["de_y"], TIGHT)
    assert_close(F.freq_encode(cu(g["x"]), (2 ** torch.linspace(0.0, 1.0, 2)).tolist(), False), g["p2_y"], TIGHT)
    # contraction through the positions kernel (point form): (contract(x)+2)/4 * selector
    px, sel = F.positions_to_unit_cube(cu(g["contract_x"]), None, None, True, None)
    ref, sel_o = O.normalize_and_select(g["contract_x"], None, True)
    assert torch.equal(px.cpu(), ref), "contracted positions must be bit-identical (they feed floor/ceil)"
    assert torch.equal(sel.cpu().bool(), sel_o)


def test_positions_ray_form_bit_exact(F, golden):
    g = golden("nerfacto_field")
    for nm, con in (("train", True), ("aabb", False)):
        o, d, eb = g[f"{nm}_origins"], g[f"{nm}_directions"], g[f"{nm}_ebins"]
        pos = O.frustum_positions(o, d, eb[:, :-1], eb[:, 1:])
        ref, sel_o = O.normalize_and_select(pos, g["aabb"], con)
        x, sel = F.positions_to_unit_cube(cu(o), cu(d), cu(eb), con, g["aabb"].flatten().tolist())
        assert torch.equal(x.cpu().view_as(ref), ref), nm
        assert torch.equal(sel.cpu().bool().view_as(sel_o), sel_o)


# ------------------------------------------------------------------------------------------------
def test_samplers_golden(F, golden):
    g = golden("samplers")
    for kind in ("piecewise", "uniform"):
        for mode in ("eval", "single", "multi"):
            jit = g.get(f"{kind}_{mode}_jitter")
            sb, eb = F.spaced_sample(cu(g["nears"]), cu(g["fars"]), 32, kind, cu(jit) if jit is not None else None)
            assert torch.equal(sb.cpu(), g[f"{kind}_{mode}_sbins"].expand(64, -1)), (kind, mode)
            assert torch.equal(eb.cpu(), g[f"{kind}_{mode}_ebins"]), (kind, mode)
    sb0 = cu(g["piecewise_eval_sbins"])
    w = cu(g["pdf_weights"][..., 0])
    for mode in ("eval", "single", "multi"):
        jit = g.get(f"pdf_{mode}_jitter")
        nsb, neb, cdf, inds = F.pdf_sample(sb0, w, 16, cu(jit) if jit is not None else None, cu(g["nears"]),
                                           cu(g["fars"]), "piecewise", want_aux=True)
        ref = O.pdf_sample(g["piecewise_eval_sbins"], g["pdf_weights"][..., 0], 16, jit)
        assert_close(cdf, ref["cdf"], 1e-6, "cdf")
        # index work: bit-exact given identical floating-point inputs (the kernel's own cdf) ...
        u = ref["u"]
        assert torch.equal(inds.cpu(), torch.searchsorted(cdf.cpu(), u, side="right")), mode
        # ... and against the reference's recorded indices: the cdf is bit-identical (normaliser summed in torch's
        # order, exact fp64 running sums), so every index and every resampled bin edge is too
        assert torch.equal(cdf.cpu(), ref["cdf"]), f"{mode}: cdf differs from the reference's"
        assert torch.equal(inds.cpu(), g[f"pdf_{mode}_inds"]), f"{mode}: searchsorted indices differ from the reference"
        assert torch.equal(nsb.cpu(), g[f"pdf_{mode}_sbins"]), mode
        assert torch.equal(neb.cpu(), g[f"pdf_{mode}_ebins"]), mode
    n, f = F.aabb_collide(cu(g["origins"]), cu(g["directions"]), [-1, -1, -1, 1, 1, 1], 0.1)
    assert torch.equal(n.cpu(), g["aabb_nears"]) and torch.equal(f.cpu(), g["aabb_fars"])


def test_torch_row_sum_bit_exact(F):
    """The PDF sampler's normaliser: same bits as torch's CPU sum for every row length the path uses (and odd ones)."""
    torch.manual_seed(11)
    for S in (1, 3, 5, 8, 16, 31, 32, 48, 96, 97, 256, 257, 1024, 2051, 4096):
        x = torch.rand(300, S) ** 3 + 0.01
        assert torch.equal(F.torch_row_sum(x.cuda()).cpu(), x.sum(-1)), S


def test_pdf_sample_full_size(F):
    """4096 rays x 256 -> 96: inds vs searchsorted on the kernel's cdf (bit-exact), bins sorted, within [0,1]."""
    torch.manual_seed(5)
    R, S = 4096, 256
    sb, eb = F.spaced_sample(torch.full((R, 1), 0.05).cuda(), torch.full((R, 1), 1000.0).cuda(), S, "piecewise",
                             torch.rand(R, 1).cuda())
    w = (torch.rand(R, S) ** 8).cuda()
    jit = torch.rand(R, 1).cuda()
    near, far = torch.full((R, 1), 0.05).cuda(), torch.full((R, 1), 1000.0).cuda()
    nsb, neb, cdf, inds = F.pdf_sample(sb, w, 96, jit, near, far, "piecewise", want_aux=True)
    ref = O.pdf_sample(sb.cpu(), w.cpu(), 96, jit.cpu())
    assert torch.equal(cdf.cpu(), ref["cdf"])
    assert torch.equal(inds.cpu(), ref["inds"]), "searchsorted indices must be bit-exact at the BASELINE size"
    assert torch.equal(nsb.cpu(), ref["bins"])
    assert bool((nsb[:, 1:] >= nsb[:, :-1]).all()) and float(nsb.min()) >= 0 and float(nsb.max()) <= 1


# ------------------------------------------------------------------------------------------------
def test_render_golden(F, golden):
    g = golden("render")
    eb = cu(g["ebins"])
    dens = cu(g["density"][..., 0]).requires_grad_(True)
    w = F.render_weights(eb, dens)
    assert_close(w, g["weights"][..., 0], TIGHT, "weights")
    (gd,) = torch.autograd.grad(w, dens, cu(g["d_weights"][..., 0]))
    assert_close(gd, g["d_density"][..., 0], REL, "d_density")
    for bg in ("last_sample", "white", "black", "random"):
        rgb = cu(g["rgb_samples"]).requires_grad_(True)
        wd = cu(g["weights"][..., 0]).requires_grad_(True)
        comp, acc, dep = F.composite(rgb, wd, eb, bg)
        assert_close(comp, g[f"rgb_{bg}"], TIGHT, bg)
        g_rgb, g_w = torch.autograd.grad(comp, [rgb, wd], cu(g[f"rgb_{bg}_dout"]))
        assert_close(g_rgb, g[f"rgb_{bg}_drgb"], TIGHT)
        assert_close(g_w, g[f"rgb_{bg}_dw"][..., 0], TIGHT)
    comp, acc, dep = F.composite(cu(g["rgb_samples_nan"]), cu(g["weights"][..., 0]), eb, "last_sample", eval_mode=True)
    assert_close(comp, g["rgb_eval"], TIGHT)
    assert_close(acc, g["accumulation"][..., 0], TIGHT)
    md, idx = F.median_depth(cu(g["weights"][..., 0]), eb)
    starts, ends = g["ebins"][:, :-1, None], g["ebins"][:, 1:, None]
    ref_d, ref_i = O.depth_median(g["weights"], starts, ends)
    assert torch.equal(idx.cpu(), ref_i[:, 0]), "median index must be bit-exact"
    assert torch.equal(md.cpu(), g["depth_median"][:, 0])
    wd = cu(g["weights"][..., 0]).requires_grad_(True)
    _, _, dep = F.composite(cu(g["rgb_samples"]), wd, eb, "random")
    steps = (g["ebins"][:, :-1] + g["ebins"][:, 1:]) / 2
    depc = torch.clip(dep, float(steps.min()), float(steps.max()))
    assert_close(depc, g["depth_expected"][:, 0], TIGHT)
    (gw,) = torch.autograd.grad(depc, wd, cu(g["depth_expected_dout"][:, 0]))
    assert_close(gw, g["depth_expected_dw"][..., 0], REL)


def test_losses_golden(F, golden):
    g = golden("losses")
    w = [cu(g[f"w{i}"][..., 0]).requires_grad_(True) for i in range(3)]
    sd = [cu(g[f"sb{i}"]) for i in range(3)]
    li = F.interlevel_loss(w, sd)
    assert_close(li, g["interlevel"], TIGHT)
    g0, g1 = torch.autograd.grad(li, w[:2])
    assert_close(g0, g["interlevel_dw0"][..., 0], REL)
    assert_close(g1, g["interlevel_dw1"][..., 0], REL)
    ld = F.distortion_loss(w[2], sd[2])
    assert_close(ld, g["distortion"], TIGHT)
    (g2,) = torch.autograd.grad(ld, [w[2]])
    assert_close(g2, g["distortion_dw2"][..., 0], REL)


def test_raygen_golden(F, golden):
    g = golden("raygen")
    intr = torch.stack([g["fx"], g["fy"], g["cx"], g["cy"]], -1)
    for nm, dist in (("nodist", None), ("dist", g["dist"])):
        r = F.generate_rays(cu(g["c2w"]), cu(intr), cu(dist) if dist is not None else None, cu(g[f"{nm}_ray_indices"]))
        assert torch.equal(r["origins"].cpu(), g[f"{nm}_origins"])
        assert_close(r["directions"], g[f"{nm}_directions"], TIGHT, nm)
        assert_close(r["pixel_area"], g[f"{nm}_pixel_area"], REL, nm)
        assert_close(r["directions_norm"], g[f"{nm}_directions_norm"], TIGHT)
        assert torch.equal(r["camera_indices"].cpu(), g[f"{nm}_camera_indices"])


def test_cameras_generate_rays_reference_signature(F, golden):
    """Cameras.generate_rays(camera_indices, coords, ..., keep_shape, aabb_box) on the kernel: training-batch form vs the
    golden rays; whole-image form ([H, W] and [H, W, n]) consistent with the per-pixel form bit for bit."""
    from nerfstudio_b200.cameras.cameras import Cameras

    g = golden("raygen")
    for nm, dist in (("nodist", None), ("dist", g["dist"])):
        cams = Cameras(cu(g["c2w"]), cu(g["fx"]), cu(g["fy"]), cu(g["cx"]), cu(g["cy"]), width=cu(g["hw"][:, 1]),
                       height=cu(g["hw"][:, 0]), distortion_params=None if dist is None else cu(dist))
        ri = g[f"{nm}_ray_indices"]
        coords = ri[:, 1:].float() + 0.5  # pixel centres, (y, x) — what get_image_coords / the pixel sampler produce
        rb = cams.generate_rays(camera_indices=cu(ri[:, :1]), coords=cu(coords))
        assert torch.equal(rb.origins.cpu(), g[f"{nm}_origins"])
        assert_close(rb.directions, g[f"{nm}_directions"], TIGHT, nm)
        assert_close(rb.pixel_area, g[f"{nm}_pixel_area"], REL, nm)
        assert_close(rb.metadata["directions_norm"], g[f"{nm}_directions_norm"], TIGHT)
        assert torch.equal(rb.camera_indices.cpu(), g[f"{nm}_camera_indices"])
        ref = F.generate_rays(cu(g["c2w"]), cams.intrinsics(), None if dist is None else cu(dist), cu(ri))
        assert torch.equal(rb.directions, ref["directions"]) and torch.equal(rb.pixel_area, ref["pixel_area"])
    # whole images (coords=None): [H, W] for an int index, [H, W, n] for n listed cameras
    H, W = 12, 20
    cams = Cameras(cu(g["c2w"]), cu(g["fx"]), cu(g["fy"]), cu(g["cx"]), cu(g["cy"]), width=torch.full((5,), W).cuda(),
                   height=torch.full((5,), H).cuda(), distortion_params=cu(g["dist"]))
    img = cams.generate_rays(camera_indices=3, keep_shape=True)
    assert tuple(img.shape) == (H, W) and img.origins.shape == (H, W, 3)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    ri = torch.stack([torch.full_like(yy, 3), yy, xx], -1).reshape(-1, 3)
    per_px = F.generate_rays(cu(g["c2w"]), cams.intrinsics(), cu(g["dist"]), cu(ri))
    assert torch.equal(img.directions.reshape(-1, 3), per_px["directions"])
    assert torch.equal(img.pixel_area.reshape(-1, 1), per_px["pixel_area"])
    multi = cams.generate_rays(camera_indices=torch.tensor([[3], [1]]).cuda())
    assert tuple(multi.shape) == (H, W, 2) and torch.equal(multi.directions[:, :, 0], img.directions)
    assert bool((multi.camera_indices[:, :, 1] == 1).all())
    flat = cams.generate_rays(camera_indices=3, keep_shape=False, aabb_box=[-1.0, -1, -1, 1, 1, 1])
    assert tuple(flat.shape) == (H * W,) and flat.nears.shape == (H * W, 1)
    to, tx, ho = O.ray_aabb_intersect(flat.origins.cpu(), flat.directions.cpu(), torch.tensor([-1.0, -1, -1, 1, 1, 1]), 0.0, 1e10)
    assert torch.equal(flat.nears.cpu()[ho][:, 0], to[ho]) and bool((flat.nears.cpu()[~ho] == 1e10).all())


def test_device_ray_pipeline(F, golden):
    """pixel sampling + colour lookup + ray generation in one launch over a uint8 image cache in HBM, against
    the reference's arithmetic: indices = (rand * [C,H,W]).long() (data/pixel_samplers.py:168-171), image[c,y,x] / 255
    (get_image_float32), rays = RayGenerator on those indices (bit-identical to the ray-generation kernel)."""
    from nerfstudio_b200.cameras.cameras import Cameras
    from nerfstudio_b200.data.device_pipeline import DeviceRayPipeline

    g = golden("raygen")
    torch.manual_seed(31)
    C_, H, W = 5, 37, 53
    images = torch.randint(0, 256, (C_, H, W, 4), dtype=torch.uint8)  # RGBA cache: alpha ignored
    cams = Cameras(cu(g["c2w"]), cu(g["fx"]), cu(g["fy"]), cu(g["cx"]), cu(g["cy"]), distortion_params=cu(g["dist"]))
    pipe = DeviceRayPipeline(cams, images.cuda(), num_rays_per_batch=4096)
    u = torch.rand(4096, 3)
    u[0] = torch.tensor([0.0, 0.0, 0.0])
    u[1] = torch.tensor([1.0 - 2 ** -24] * 3)
    bundle, batch = pipe.sample(u.cuda())
    idx = (u * torch.tensor([C_, H, W])).long()
    assert torch.equal(batch["indices"].cpu(), idx)
    assert torch.equal(batch["image"].cpu(), images[idx[:, 0], idx[:, 1], idx[:, 2], :3].float() / 255.0)
    ref = F.generate_rays(cu(g["c2w"]), cams.intrinsics(), cu(g["dist"]), idx.cuda())
    assert torch.equal(bundle.origins, ref["origins"]) and torch.equal(bundle.directions, ref["directions"])
    assert torch.equal(bundle.pixel_area, ref["pixel_area"]) and torch.equal(bundle.camera_indices, ref["camera_indices"])
    # a cached SUBSET of the cameras: indices are corrected to absolute camera ids (pixel_samplers.py:312)
    sub = DeviceRayPipeline(cams, images[:2].cuda(), 64, image_idx=torch.tensor([4, 2]))
    b2, batch2 = sub.sample(u[:64].cuda())
    local = (u[:64, 0] * 2).long()
    assert torch.equal(batch2["indices"][:, 0].cpu(), torch.tensor([4, 2])[local])
    rb, bt = pipe.next_train(0)
    assert rb.origins.shape == (4096, 3) and bt["image"].shape == (4096, 3) and float(bt["image"].max()) <= 1.0


# ------------------------------------------------------------------------------------------------
def test_packed_path_vs_oracle(F):
    torch.manual_seed(7)
    R = 300
    counts = torch.randint(0, 40, (R,))
    counts[5] = 0
    counts[R - 1] = 0
    ri = torch.repeat_interleave(torch.arange(R), counts)
    M = ri.numel()
    info = F.pack_info(ri.cuda(), R)
    assert torch.equal(info.cpu(), O.pack_info(ri, R))
    ts = torch.rand(M).cumsum(0) * 0.01
    te = ts + 0.005
    sig = (torch.rand(M) * 50).requires_grad_(True)
    wo, to, ao = O.packed_weights(ts, te, sig, ri, R)
    dw = torch.randn(M)
    (gs,) = torch.autograd.grad(wo, sig, dw)
    sc = sig.detach().cuda().requires_grad_(True)
    w, tr, al = F.packed_weights(ts.cuda(), te.cuda(), sc, info)
    assert_close(w, wo, TIGHT)
    assert_close(tr, to, TIGHT)
    assert_close(al, ao, TIGHT)
    (gc,) = torch.autograd.grad(w, sc, dw.cuda())
    assert_close(gc, gs, REL)
    vals = torch.rand(M, 3, requires_grad=True)
    wl = wo.detach().clone().requires_grad_(True)
    acc_o = O.accumulate_along_rays(wl, vals, ri, R)
    dout = torch.randn(R, 3)
    gwo, gvo = torch.autograd.grad(acc_o, [wl, vals], dout)
    wc, vc = wo.detach().cuda().requires_grad_(True), vals.detach().cuda().requires_grad_(True)
    acc = F.packed_accumulate(wc, vc, info)
    assert_close(acc, acc_o, TIGHT)
    gw, gv = torch.autograd.grad(acc, [wc, vc], dout.cuda())
    assert_close(gw, gwo, TIGHT)
    assert_close(gv, gvo, TIGHT)
    assert_close(F.packed_accumulate(wc, None, info), O.accumulate_along_rays(wo.detach(), None, ri, R), TIGHT)


def test_occgrid_march_bit_exact_vs_oracle(F):
    torch.manual_seed(8)
    levels, res = 2, 16
    binaries = torch.rand(levels, res, res, res) > 0.6
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1])
    R = 48
    o = torch.randn(R, 3) * 1.5
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(R, 3), dim=-1)
    jit = torch.rand(R)
    for cone, jitter in ((0.0, None), (0.01, jit)):
        ri_o, ts_o, te_o = O.occgrid_march(o, d, binaries, aabb, 0.05, 0.0, 1e10, cone, jitter)
        ri, ts, te = F.occgrid_march(o.cuda(), d.cuda(), binaries.cuda(), aabb.tolist(), 0.05, 0.0, 1e10, cone,
                                     None if jitter is None else jitter.cuda())
        assert torch.equal(ri.cpu(), ri_o), "ray indices / counts must be bit-exact"
        assert torch.equal(ts.cpu(), ts_o) and torch.equal(te.cpu(), te_o)
    # empty grid -> no samples
    ri, ts, te = F.occgrid_march(o.cuda(), d.cuda(), torch.zeros_like(binaries).cuda(), aabb.tolist(), 0.05)
    assert ri.numel() == 0


def test_occgrid_march_full_size_warp_equals_serial(F):
    """BASELINE config 2 geometry (128^3 x 4 levels, aabb +-1.5, step = diag/1000, cone_angle 0.004, 4096 rays, sphere
    occupancy): the warp-per-ray march emits exactly the serial march's samples (which is bit-exact vs the oracle)."""
    from nerfstudio_b200 import lib
    from nerfstudio_b200.scene import sphere_scene_rays

    torch.manual_seed(9)
    res, levels = 128, 4
    idx = torch.stack(torch.meshgrid([torch.arange(res)] * 3, indexing="ij"), -1).float()
    binaries = torch.zeros(levels, res, res, res, dtype=torch.bool)
    for lvl in range(levels):
        c = ((idx + 0.5) / res * 2 - 1) * 1.5 * 2 ** lvl
        binaries[lvl] = c.norm(dim=-1) < 0.55
    rays, _ = sphere_scene_rays(4096, seed=3)
    o, d = rays["origins"].cuda(), rays["directions"].cuda()
    step = (3 * 3.0 ** 2) ** 0.5 / 1000
    jit = torch.rand(4096).cuda()
    outs = []
    for mode in (0, 1):
        try:
            assert lib.tune("march_warp", mode)
            outs.append(F.occgrid_march(o, d, binaries.cuda(), [-1.5] * 3 + [1.5] * 3, step, 0.05, 1e3, 0.004, jit))
        finally:
            lib.tune("march_warp", 1)
    assert outs[0][0].numel() > 10000
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_scan_counts(F):
    torch.manual_seed(12)
    for n in (1, 31, 1024, 1025, 4096, 5000, 8193, 1_000_003):
        c = torch.randint(0, 300, (n,), dtype=torch.int32)
        off, total = F.scan_counts(c.cuda())
        cs = torch.cumsum(c.long(), 0)
        assert total == int(cs[-1]) and torch.equal(off.cpu(), cs - c.long()), n


def test_packed_prune_bit_exact_vs_oracle(F):
    """K8 (visibility pruning inside OccGridEstimator.sampling): same kept samples, same order, as the oracle."""
    torch.manual_seed(13)
    R = 500
    counts = torch.randint(0, 90, (R,))
    counts[0] = counts[77] = 0
    ri = torch.repeat_interleave(torch.arange(R), counts)
    M = ri.numel()
    ts = torch.rand(M) * 0.01 + torch.arange(M) * 0.01
    te = ts + 0.01
    sig = torch.rand(M) ** 4 * 400
    _, trans_o, alphas_o = O.packed_weights(ts, te, sig, ri, R)
    info = F.pack_info(ri.cuda(), R)
    for eps, thre, mean in ((1e-4, 0.01, None), (1e-4, 0.01, 0.003), (1e-2, 0.0, None)):
        ri_o, ts_o, te_o, keep = O.packed_visibility_prune(ri, ts, te, sig, R, eps, thre, mean)
        cap = None if mean is None else torch.tensor([mean]).cuda()
        # identical floating inputs (the oracle's T and alpha) -> identical index work
        r2, s2, e2 = F.packed_prune(ri.cuda(), ts.cuda(), te.cuda(), trans_o.cuda(), alphas_o.cuda(), info, eps, thre, cap)
        assert torch.equal(r2.cpu(), ri_o) and torch.equal(s2.cpu(), ts_o) and torch.equal(e2.cpu(), te_o), (eps, thre, mean)
        assert 0 < r2.numel() < M
        # and through the kernels' own T / alpha: counts agree up to samples sitting on a threshold
        _, tr, al = F.packed_weights(ts.cuda(), te.cuda(), sig.cuda(), info)
        r3, _, _ = F.packed_prune(ri.cuda(), ts.cuda(), te.cuda(), tr, al, info, eps, thre, cap)
        assert abs(r3.numel() - ri_o.numel()) <= 2
    x = F.packed_positions(torch.rand(R, 3).cuda(), torch.rand(R, 3).cuda(), ri.cuda(), ts.cuda(), te.cuda())
    assert x.shape == (M, 3)


def test_packed_positions_bit_exact(F):
    torch.manual_seed(14)
    R, M = 64, 3000
    o, d = torch.randn(R, 3), torch.randn(R, 3)
    ri = torch.sort(torch.randint(0, R, (M,))).values
    ts = torch.rand(M) * 5
    te = ts + torch.rand(M) * 0.1
    ref = o[ri] + d[ri] * (ts + te)[:, None] / 2.0  # model_components/ray_samplers.py:420
    assert torch.equal(F.packed_positions(o.cuda(), d.cuda(), ri.cuda(), ts.cuda(), te.cuda()).cpu(), ref)


def test_occgrid_update_bit_exact_vs_oracle(F):
    """K9 (a27): jittered cell centres, EMA-max with duplicate cells, fp64 mean threshold, binaries — against the oracle
    on a RECORDED cell-sample / jitter stream, warm-up (all cells) and sampled (uniform + occupied) updates."""
    from nerfstudio_b200.shims import nerfacc

    torch.manual_seed(15)
    res, levels = 16, 3
    per = res ** 3
    grid = nerfacc.OccGridEstimator(roi_aabb=[-1.0, -1, -1, 1, 1, 1], resolution=res, levels=levels).cuda().train()
    aabbs = grid.aabbs.cpu()

    def occ_fn(x):  # exact in fp32 on both devices: comparisons and separately rounded mul/add only
        r2 = x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1] + x[:, 2] * x[:, 2]
        return (r2 < 0.49).to(x.dtype) * 0.03 + x[:, 0].abs() * 0.001

    occs_o = torch.zeros(levels * per)
    # step 0: warm-up -> every cell
    cells = [None] * levels
    jit = [torch.rand(per, 3) for _ in range(levels)]
    occs_o, bin_o, thre_o = O.occgrid_update(occs_o, levels, res, aabbs, cells, jit, occ_fn)
    grid.update_every_n_steps(0, occ_fn, cells=cells, jitters=[j.cuda() for j in jit])
    x = F.occgrid_points(None, jit[1].cuda(), res, aabbs[1].tolist())
    assert torch.equal(x.cpu(), O.occgrid_cell_points(torch.arange(per), jit[1], res, aabbs[1])), "cell sample points"
    assert torch.equal(grid.occs.cpu(), occs_o), "occupancy values after the warm-up update"
    assert torch.equal(grid.binaries.flatten().cpu(), bin_o) and float(grid._stats[0]) == float(thre_o)
    assert 0 < int(bin_o.sum()) < bin_o.numel()
    # step 16 (as if past warm-up): sampled cells, with duplicates (uniform draws + occupied cells overlap)
    for rep in range(3):
        cells, jit = [], []
        for lvl in range(levels):
            uni = torch.randint(per, (per // 4,))
            occ_ids = torch.nonzero(bin_o.view(levels, per)[lvl])[:, 0]
            if per // 4 < occ_ids.numel():
                occ_ids = occ_ids[torch.randint(occ_ids.numel(), (per // 4,))]
            ids = torch.cat([uni, occ_ids, uni[:50]])
            cells.append(ids), jit.append(torch.rand(ids.numel(), 3))
        occs_o, bin_o, thre_o = O.occgrid_update(occs_o, levels, res, aabbs, cells, jit, occ_fn)
        grid.update_every_n_steps(16 * (rep + 1), occ_fn, warmup_steps=0, cells=[c.cuda() for c in cells],
                                  jitters=[j.cuda() for j in jit])
        assert torch.equal(grid.occs.cpu(), occs_o), rep
        assert torch.equal(grid.binaries.flatten().cpu(), bin_o), rep
    # the fresh-draw path (no recorded stream) runs and keeps the invariants
    grid.update_every_n_steps(64, occ_fn, warmup_steps=0)
    assert float(grid.occs.min()) >= 0 and bool(grid.binaries.any())
    st = grid._occ_stats()
    assert abs(float(st[1]) - float(grid.occs.double().mean())) < 1e-7


def test_ray_aabb_intersect_golden(F, golden):
    """nerfacc.ray_aabb_intersect through the reference's own proxy (utils/math.py:138-175 recorded in the golden file)."""
    g = golden("aabb_intersect")
    tmin, tmax, hit = F.ray_aabb_intersect(cu(g["origins"]), cu(g["directions"]), cu(g["aabb"]).reshape(1, 6), 0.0, 1e10, 1e10)
    to, tx, ho = O.ray_aabb_intersect(g["origins"], g["directions"], g["aabb"], near=0.0, far=1e10)
    assert torch.equal(hit[:, 0].cpu(), ho)
    assert torch.equal(tmin[:, 0].cpu()[ho], to[ho]) and torch.equal(tmax[:, 0].cpu()[ho], tx[ho])
    assert torch.equal(ho, g["t_min"] < 1e10)  # the reference's recorded hits (tests/utils/test_aabb_intersection.py)
    assert torch.allclose(tmin[:, 0].cpu()[ho], g["t_min"][ho], rtol=1e-3) and torch.allclose(tmax[:, 0].cpu()[ho], g["t_max"][ho], rtol=1e-3)
    assert bool((tmin[:, 0].cpu()[~ho] == 1e10).all())


def test_adam_vs_oracle(F):
    torch.manual_seed(9)
    n = 100003
    p, g = torch.randn(n), torch.randn(n) * 1e-3
    g[::7] = 0.0
    m, v = torch.zeros(n), torch.zeros(n)
    pc, mc, vc = p.cuda(), m.cuda(), v.cuda()
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=1e-2, eps=1e-15)
    for step in range(1, 4):
        O.adam_step(p, g * step, m, v, step, 1e-2)
        F.adam_step(pc, (g * step).cuda(), mc, vc, step, 1e-2)
        pt.grad = (g * step).clone()
        opt.step()
    assert_close(p, pt.detach(), 1e-6, "oracle adam vs torch.optim.Adam")
    assert_close(pc, p, 1e-5)
    assert_close(mc, m, 1e-5)
    assert_close(vc, v, 1e-5)


def test_camera_optimizer_pose_apply_vs_reference(F, golden):
    """a4: fused SO3xR3 ray correction and its pose gradients vs the reference's CameraOptimizer (golden)."""
    from nerfstudio_b200.cameras.camera_optimizers import CameraOptimizer, CameraOptimizerConfig
    from nerfstudio_b200.cameras.rays import RayBundle

    g = golden("camera_opt")
    C = g["pose"].shape[0]
    opt = CameraOptimizer(CameraOptimizerConfig(mode="SO3xR3"), num_cameras=C, device="cuda")
    with torch.no_grad():
        opt.pose_adjustment.copy_(g["pose"].cuda())
    rb = RayBundle(origins=g["origins"].cuda(), directions=g["directions"].cuda(),
                   pixel_area=torch.ones(g["origins"].shape[0], 1, device="cuda"), camera_indices=g["cams"].cuda())
    opt.apply_to_raybundle(rb)
    assert_close(rb.origins, g["out_origins"], 1e-6, "origins")
    assert_close(rb.directions, g["out_directions"], 1e-5, "directions")
    (gp,) = torch.autograd.grad((rb.origins * g["go"].cuda()).sum() + (rb.directions * g["gd"].cuda()).sum(),
                                [opt.pose_adjustment])
    assert_close(gp, g["g_pose"], 1e-4, "g_pose")
    assert_close(opt(torch.arange(C, device="cuda")), g["matrices"], 1e-6, "matrices")
    losses = {}
    opt.get_loss_dict(losses)
    assert_close(losses["camera_opt_regularizer"], g["regularizer"], 1e-6, "regularizer")

    # non-trainable cameras: identity transform and no gradient (camera_optimizers.py:127-131)
    frozen_idx = torch.tensor([2, 5])
    opt2 = CameraOptimizer(CameraOptimizerConfig(mode="SO3xR3"), num_cameras=C, device="cuda",
                           non_trainable_camera_indices=frozen_idx)
    with torch.no_grad():
        opt2.pose_adjustment.copy_(g["pose"].cuda())
    rb2 = RayBundle(origins=g["origins"].cuda(), directions=g["directions"].cuda(),
                    pixel_area=torch.ones(g["origins"].shape[0], 1, device="cuda"), camera_indices=g["cams"].cuda())
    opt2.apply_to_raybundle(rb2)
    is_frozen = torch.isin(g["cams"].reshape(-1), frozen_idx)
    assert torch.equal(rb2.origins.cpu()[is_frozen], g["origins"][is_frozen])
    assert torch.equal(rb2.directions.cpu()[is_frozen], g["directions"][is_frozen])
    assert_close(rb2.directions.cpu()[~is_frozen], g["out_directions"][~is_frozen], 1e-5, "unfrozen directions")
    (gp2,) = torch.autograd.grad(rb2.origins.sum() + (rb2.directions * g["gd"].cuda()).sum(), [opt2.pose_adjustment])
    assert float(gp2[frozen_idx].abs().max()) == 0.0
    # mode "off" leaves the bundle alone and owns no parameters
    off = CameraOptimizer(CameraOptimizerConfig(mode="off"), num_cameras=C, device="cuda")
    off.apply_to_raybundle(rb2)
    assert len(list(off.parameters())) == 0


def test_other_spaced_samplers_golden(F, golden):
    """LinearDisparity / Sqrt / Log spacing (reference tests/model_components/test_ray_sampler.py:37-86): spacing bins
    bit-exact; euclidean bins bit-exact for 1/x and sqrt (IEEE-rounded), ≤ 1e-6 for log/exp (libm vs CUDA ulps)."""
    g = golden("samplers_extra")
    for kind in ("lindisp", "sqrt", "log"):
        tol = 1e-6 if kind == "log" else 0.0
        _, eb = F.spaced_sample(cu(g["t_nears"]), cu(g["t_fars"]), 15, kind, None)
        assert eb.shape == (10, 16)
        assert_close(eb, g[f"{kind}_t_ebins"], tol, f"{kind} test set-up")
        for mode in ("eval", "single"):
            jit = g.get(f"{kind}_{mode}_jitter")
            sb, eb = F.spaced_sample(cu(g["nears"]), cu(g["fars"]), 24, kind, cu(jit) if jit is not None else None)
            assert torch.equal(sb.cpu(), g[f"{kind}_{mode}_sbins"].expand(64, -1)), (kind, mode)
            assert_close(eb, g[f"{kind}_{mode}_ebins"], tol, f"{kind} {mode}")


def test_step_begin_philox_and_zeroing():
    """b2n_step_begin: draws bit-equal to the Philox oracle for (seed, draw), the draw counter advances per launch (also
    across replays of a captured graph), accumulators are zeroed."""
    from nerfstudio_b200 import lib
    from nerfstudio_b200.lib import call, ptr, stream

    n = 3 * 4096 + 5
    seed = 0x1234ABCD5678
    state = torch.tensor([seed, 7, 0], dtype=torch.int64, device="cuda")
    u = torch.empty(n, device="cuda")
    z0, z1 = torch.ones(5, device="cuda"), torch.ones(2 * 4096 * 3, device="cuda")
    call("b2n_step_begin", ptr(u), n, ptr(state, torch.int64), ptr(z0), z0.numel(), ptr(z1), z1.numel(), stream())
    assert np.array_equal(u.cpu().numpy(), O.philox_uniform(n, seed, 7))
    assert state.tolist() == [seed, 8, 0]
    assert float(z0.abs().max()) == 0.0 and float(z1.abs().max()) == 0.0
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            call("b2n_step_begin", ptr(u), n, ptr(state, torch.int64), ptr(None), 0, ptr(None), 0, stream())
    for k in range(3):
        g.replay()
        torch.cuda.synchronize()
        assert np.array_equal(u.cpu().numpy(), O.philox_uniform(n, seed, 8 + k)), k
    assert state.tolist() == [seed, 11, 0]
    a, b = torch.randn(1000, device="cuda"), torch.randn(1000, device="cuda")
    want = a + b
    call("b2n_add_inplace", ptr(a), ptr(b), 1000, stream())
    assert torch.equal(a, want)
    t = torch.tensor([0.1, 0.2, 0.3, 0.0, 0.7], device="cuda")
    call("b2n_loss_total", ptr(t), 3, ptr(t[4:]), ptr(t[3:]), stream())
    assert float(t[3]) == float(((t[0] + t[1]) + t[2]) + t[4])


def test_fused_weights_pdf_sample_equals_the_two_operators():
    """b2n_weights_pdf_sample = b2n_weights_fwd followed by b2n_pdf_sample, bit for bit (weights and new edges), with the
    anneal exponent on and off, ragged sample counts."""
    from nerfstudio_b200.functional import _linspace
    from nerfstudio_b200.lib import SPACING, call, ptr, stream

    torch.manual_seed(3)
    for R, S, nb, anneal in ((513, 256, 97, 0.7), (257, 96, 49, 1.0), (64, 37, 20, 0.35)):
        sb = torch.sort(torch.rand(R, S + 1, device="cuda"), dim=-1).values
        near, far = torch.full((R,), 0.05, device="cuda"), torch.full((R,), 1000.0, device="cuda")
        eb = 1.0 / (1.0 / near[:, None] * (1 - sb) + 1.0 / far[:, None] * sb)
        eb = torch.sort(eb, dim=-1).values.contiguous()
        dens = torch.rand(R, S, device="cuda") * 5 * (torch.rand(R, S, device="cuda") > 0.4)
        jit = torch.rand(R, 1, device="cuda")
        u_base = _linspace(0.0, 1.0 - (1.0 / nb), nb, sb.device)
        w0 = torch.empty(R, S, device="cuda")
        call("b2n_weights_fwd", ptr(eb), ptr(eb.view(-1)[1:]), S + 1, ptr(dens), R, S, ptr(w0), stream())  # ends = starts + 1 element
        ns0, ne0 = torch.empty(R, nb, device="cuda"), torch.empty(R, nb, device="cuda")
        call("b2n_pdf_sample", ptr(sb), ptr(w0), ptr(u_base), ptr(jit), 0, ptr(near), ptr(far), R, S, nb, float(anneal), ptr(None),
             0.01, 1e-5, SPACING["piecewise"], ptr(ns0), ptr(ne0), ptr(None), ptr(None), stream())
        w1, ns1, ne1 = torch.empty_like(w0), torch.empty_like(ns0), torch.empty_like(ne0)
        call("b2n_weights_pdf_sample", ptr(sb), ptr(eb), ptr(dens), ptr(u_base), ptr(jit), 0, ptr(near), ptr(far), R, S, nb,
             float(anneal), ptr(None), 0.01, 1e-5, SPACING["piecewise"], ptr(w1), ptr(ns1), ptr(ne1), stream())
        assert torch.equal(w0, w1) and torch.equal(ns0, ns1) and torch.equal(ne0, ne1), (R, S)


def test_head_input_parts_compose():
    """b2n_head_input_fwd_part: the static (SH + embedding) and the geo column groups together write exactly what the whole
    operator writes; backward parts 1 + 2 produce what part 0 produces."""
    from nerfstudio_b200.lib import call, ptr, stream

    torch.manual_seed(5)
    R, S, n_sh, geo, n_emb, bw, stride = 129, 48, 16, 15, 32, 16, 64
    sh, base = torch.randn(R, n_sh, device="cuda"), torch.randn(R * S, bw, device="cuda")
    emb, cam = torch.randn(7, n_emb, device="cuda"), torch.randint(0, 7, (R,), device="cuda")
    outs = []
    for parts in ((0,), (1, 2), (2, 1)):
        out = torch.full((R * S, stride), float("nan"), device="cuda")
        for part in parts:
            call("b2n_head_input_fwd_part", ptr(sh), n_sh, ptr(base), bw, geo, ptr(emb), ptr(cam, torch.int64), n_emb, 1, R, S,
                 ptr(out), stride, part, stream())
        outs.append(out)
    assert not torch.isnan(outs[0]).any()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = torch.cat([sh[:, None, :].expand(R, S, n_sh).reshape(R * S, n_sh), base[:, 1:1 + geo], emb[cam][:, None, :].expand(R, S, n_emb)
                     .reshape(R * S, n_emb), torch.zeros(R * S, 1, device="cuda")], dim=1)
    assert torch.equal(outs[0], ref)
    d_in, d_pre = torch.randn(R * S, stride, device="cuda"), torch.randn(R * S, device="cuda")
    res = []
    for parts in ((0,), (1, 2)):
        d_base, d_emb = torch.full((R * S, bw), float("nan"), device="cuda"), torch.zeros_like(emb)
        for part in parts:
            call("b2n_head_input_bwd_part", ptr(d_in), stride, n_sh, geo, n_emb, ptr(d_pre), ptr(cam, torch.int64), R, S, ptr(d_base), bw,
                 ptr(d_emb), part, stream())
        res.append((d_base, d_emb))
    assert torch.equal(res[0][0], res[1][0])
    assert_close(res[1][1], res[0][1], 1e-5, "embedding-row gradients")  # float atomics across rays of one camera
    want = torch.zeros_like(emb).index_add_(0, cam, d_in[:, n_sh + geo: n_sh + geo + n_emb].reshape(R, S, n_emb).sum(1))
    assert_close(res[0][1], want, 1e-4, "embedding-row gradients vs torch")
