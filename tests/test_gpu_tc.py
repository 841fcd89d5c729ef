"""GPU: tcgen05 building blocks and the tensor-core MLP (3xTF32) against fp64 / the oracle."""
import zlib

import pytest
import torch

from conftest import assert_close
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nerfstudio_b200 import functional

    return functional


def _selftest(mode, three, A, B, M, N, K):
    from nerfstudio_b200.lib import call, ptr, stream

    out = torch.full((128, N), float("nan"), device="cuda")
    call("b2n_tc_selftest", mode, three, ptr(A), A.shape[0], A.shape[1], ptr(B), B.shape[0], B.shape[1], M, N, K,
         ptr(out), stream())
    return out


def test_tcgen05_tf32_gemm_and_3xtf32_split(F):
    """K-major x K-major UMMA from hand-built descriptors; one pass ~ tf32 accuracy, three passes ~ fp32."""
    torch.manual_seed(0)
    for N, K in ((16, 8), (64, 64), (16, 64), (64, 32), (32, 16)):
        A, B = torch.randn(128, K, device="cuda"), torch.randn(N, K, device="cuda")
        ref = A.double() @ B.double().T
        one = _selftest(0, 0, A, B, 128, N, K)
        three = _selftest(0, 1, A, B, 128, N, K)
        e1 = float((one.double() - ref).abs().max() / ref.abs().max())
        e3 = float((three.double() - ref).abs().max() / ref.abs().max())
        assert 1e-5 < e1 < 5e-3, (N, K, e1)  # a single tf32 pass is NOT good enough for the 1e-4 parity bar
        assert e3 < 5e-6, (N, K, e3)


TC_BWD_ISSUER_DEFAULT = 1  # csrc/mlp_tc.cu g_tc_bwd_issuer

CFGS = {
    "base": (32, [64, 16], "none"),
    "head": (63, [64, 64, 3], "sigmoid"),
    "prop": (10, [16, 1], "none"),
    "deep": (16, [32, 32, 32, 8], "relu"),
}


@pytest.mark.parametrize("name", list(CFGS))
def test_mlp_tc_forward_backward_vs_oracle(F, name):
    in_dim, dims, out_act = CFGS[name]
    torch.manual_seed(zlib.crc32(name.encode()) % 1000)  # str hashes are salted per process
    n = 128 * 300 + 77  # ragged, > 2 tiles per CTA
    spec = F.MlpSpec(in_dim, dims, out_act=out_act)
    assert F.mlp_tc_supported(spec)
    x = torch.randn(n, in_dim)
    ws, prev = [], in_dim
    for d in dims:
        ws.append(torch.randn(d, prev) * (1.5 / prev ** 0.5))
        prev = d
    bs = [torch.randn(d) * 0.1 for d in dims]
    dy = torch.randn(n, dims[-1])
    # Rows with a hidden pre-activation within 1e-5 of the ReLU kink are ill-conditioned: two correct fp32
    # evaluations may land on different sides and their gradients then differ by O(1).  Such rows (about one per
    # million hidden units) get a zero upstream gradient so the comparison below stays strict for all others.
    with torch.no_grad():
        h, near_kink = x, torch.zeros(n, dtype=torch.bool)
        relu_layers = len(ws) if out_act == "relu" else len(ws) - 1  # a ReLU output has a kink of its own
        for w_, b_ in list(zip(ws, bs))[:relu_layers]:
            z = torch.nn.functional.linear(h, w_, b_)
            near_kink |= (z.abs() < 1e-5).any(dim=1)
            h = torch.relu(z)
    dy[near_kink] = 0.0
    xo = x.clone().requires_grad_(True)
    wl = [w.clone().requires_grad_(True) for w in ws]
    bl = [b.clone().requires_grad_(True) for b in bs]
    yo = O.mlp_forward(xo, wl, bl, out_act=out_act)
    go = torch.autograd.grad(yo, [xo] + wl + bl, dy)
    xc, wc, bc = x.cuda(), [w.cuda() for w in ws], [b.cuda() for b in bs]
    y, hidden = F.mlp_tc_forward(spec, xc, wc, bc, save_hidden=True)
    assert_close(y, yo, 2e-5, name + " y")
    dws, dbs = [torch.zeros_like(w) for w in wc], [torch.zeros_like(b) for b in bc]
    dx = F.mlp_tc_backward(spec, xc, y, hidden, dy.cuda(), wc, bc, dws, dbs, want_dx=True)
    assert_close(dx, go[0], 1e-4, name + " dx")
    L = len(dims)
    for i in range(L):
        assert_close(dws[i], go[1 + i], 1e-4, f"{name} dw{i}")
        assert_close(dbs[i], go[1 + L + i], 1e-4, f"{name} db{i}")


def test_mlp_tc_padded_rows_and_small_batch(F):
    """x with a padded row stride (the engine's 64-wide head input) and a batch smaller than one tile."""
    torch.manual_seed(5)
    spec = F.MlpSpec(63, [64, 64, 3], out_act="sigmoid")
    n = 50
    x = torch.zeros(n, 64)
    x[:, :63] = torch.randn(n, 63)
    ws = [torch.randn(64, 63) * 0.2, torch.randn(64, 64) * 0.2, torch.randn(3, 64) * 0.2]
    bs = [torch.randn(64) * 0.1, torch.randn(64) * 0.1, torch.randn(3) * 0.1]
    yo = O.mlp_forward(x[:, :63], ws, bs, out_act="sigmoid")
    y, _ = F.mlp_tc_forward(spec, x.cuda(), [w.cuda() for w in ws], [b.cuda() for b in bs], save_hidden=False)
    assert_close(y, yo, 2e-5)


def test_mlp_tc_rejects_unsupported(F):
    spec = F.MlpSpec(63, [256, 256, 3])
    assert not F.mlp_tc_supported(spec)
    with pytest.raises(ValueError):
        F.mlp_tc_forward(spec, torch.randn(8, 63).cuda(), [torch.randn(256, 63).cuda(), torch.randn(256, 256).cuda(),
                                                           torch.randn(3, 256).cuda()], [None] * 3, False)


@pytest.mark.parametrize("name", ["base", "head", "deep"])
def test_forward_variants_agree_bitwise(F, name):
    """The serial kernel and the two-slot warp-specialised kernel issue the same MMA sequence per tile: identical bits."""
    from nerfstudio_b200 import lib

    in_dim, dims, out_act = CFGS[name]
    torch.manual_seed(zlib.crc32(name.encode()) % 1000 + 1)
    n = 128 * 149 + 5  # one CTA gets two tiles (both slots busy), the others one
    spec = F.MlpSpec(in_dim, dims, out_act=out_act)
    x = torch.randn(n, in_dim, device="cuda")
    ws, prev = [], in_dim
    for d in dims:
        ws.append(torch.randn(d, prev, device="cuda") / prev ** 0.5)
        prev = d
    bs = [torch.randn(d, device="cuda") * 0.1 for d in dims]
    outs = []
    try:
        for slots in (1, 2):
            assert lib.tune("tc_fwd_slots", slots)
            y, hidden = F.mlp_tc_forward(spec, x, ws, bs, save_hidden=True)
            torch.cuda.synchronize()
            outs.append((y.clone(), hidden.clone()))
    finally:
        lib.tune("tc_fwd_slots", 2)
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("name", ["base", "head", "deep"])
def test_mlp_tc_tma_weight_image_bitwise(F, name):
    """Weights staged by TMA from the packed image (b2n_mlp_tc_pack + *_ws) give bit-identical outputs, hidden
    activations and input gradients to the per-CTA scalar staging; weight gradients agree to atomics-order round-off."""
    in_dim, dims, out_act = CFGS[name]
    torch.manual_seed(5)
    n = 128 * 200 + 13
    spec = F.MlpSpec(in_dim, dims, out_act=out_act)
    x = torch.randn(n, in_dim, device="cuda")
    ws_, bs_ = [], []
    prev = in_dim
    for o in dims:
        ws_.append((torch.randn(o, prev, device="cuda") / prev ** 0.5).contiguous()), bs_.append(torch.randn(o, device="cuda") * 0.1)
        prev = o
    image = F.mlp_tc_pack(spec, ws_, bs_)
    y0, h0 = F.mlp_tc_forward(spec, x, ws_, bs_, True)
    y1, h1 = F.mlp_tc_forward(spec, x, ws_, bs_, True, workspace=image)
    assert torch.equal(y0, y1) and torch.equal(h0, h1)
    dy = torch.randn_like(y0)
    outs = []
    for ws in (None, image):
        dws, dbs = [torch.zeros_like(w) for w in ws_], [torch.zeros_like(b) for b in bs_]
        dx = F.mlp_tc_backward(spec, x, y0, h0, dy, ws_, bs_, dws, dbs, True, workspace=ws)
        outs.append((dx, dws, dbs))
    assert torch.equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1] + outs[0][2], outs[1][1] + outs[1][2]):
        assert_close(a, b, 1e-5)


@pytest.mark.parametrize("name", list(CFGS))
@pytest.mark.parametrize("want_dx", [True, False])
def test_backward_variants_agree(F, name, want_dx):
    """The backward with the asynchronous MMA issue (tune tc_bwd_issuer = 1: mbarrier hand-offs instead of CTA barriers)
    runs the same MMA sequence on the same operands as the serial kernel: dx bit-identical, dW/db to atomics-order
    round-off.  Ragged batch with several tiles per CTA, with and without the TMA weight image."""
    from nerfstudio_b200 import lib

    in_dim, dims, out_act = CFGS[name]
    torch.manual_seed(zlib.crc32(name.encode()) % 1000 + 2)
    n = 128 * 148 * 3 + 41
    spec = F.MlpSpec(in_dim, dims, out_act=out_act)
    x = torch.randn(n, in_dim, device="cuda")
    ws_, bs_, prev = [], [], in_dim
    for o in dims:
        ws_.append((torch.randn(o, prev, device="cuda") / prev ** 0.5).contiguous()), bs_.append(torch.randn(o, device="cuda") * 0.1)
        prev = o
    image = F.mlp_tc_pack(spec, ws_, bs_)
    y, hidden = F.mlp_tc_forward(spec, x, ws_, bs_, True)
    dy = torch.randn_like(y)
    outs = {}
    try:
        for issuer in (0, 2):
            assert lib.tune("tc_bwd_issuer", issuer)
            for tag, wsp in (("scalar", None), ("tma", image)):
                dws, dbs = [torch.zeros_like(w) for w in ws_], [torch.zeros_like(b) for b in bs_]
                dx = F.mlp_tc_backward(spec, x, y, hidden, dy, ws_, bs_, dws, dbs, want_dx, workspace=wsp)
                torch.cuda.synchronize()
                outs[(issuer, tag)] = (dx, dws + dbs)
    finally:
        lib.tune("tc_bwd_issuer", TC_BWD_ISSUER_DEFAULT)
    for tag in ("scalar", "tma"):
        a, b = outs[(0, tag)], outs[(2, tag)]
        if want_dx:
            assert torch.equal(a[0], b[0]), tag
        for u, v in zip(a[1], b[1]):
            assert_close(u, v, 1e-5, tag)
