"""Micro-benchmark of the MLP kernels (SIMT vs tensor-core) at the nerfacto shapes (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfstudio_b200 import functional as F

torch.manual_seed(0)
CFG = {"head": (63, [64, 64, 3], "sigmoid", 196608), "base": (32, [64, 16], "none", 196608), "prop0": (10, [16, 1], "none", 1048576)}
which = sys.argv[1:] or list(CFG)
for name in which:
    in_dim, dims, oact, n = CFG[name]
    spec = F.MlpSpec(in_dim, dims, out_act=oact)
    stride = (in_dim + 3) // 4 * 4
    x = torch.zeros(n, stride, device="cuda"); x[:, :in_dim] = torch.randn(n, in_dim, device="cuda")
    ws, prev = [], in_dim
    for d in dims:
        ws.append(torch.randn(d, prev, device="cuda") / prev ** 0.5); prev = d
    bs = [torch.zeros(d, device="cuda") for d in dims]
    dy = torch.randn(n, dims[-1], device="cuda")
    dws, dbs = [torch.zeros_like(w) for w in ws], [torch.zeros_like(b) for b in bs]
    xs = x[:, :in_dim].contiguous()
    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    y, hid = F.mlp_tc_forward(spec, x, ws, bs, True, x_stride=stride)
    ys, hids = F.mlp_forward(spec, xs, ws, bs, True)
    from nerfstudio_b200 import lib
    image = F.mlp_tc_pack(spec, ws, bs)
    for issuer in (0, 2):
        lib.tune("tc_bwd_issuer", issuer)
        print(name, "tc bwd (tma image) issuer=%d %.3f ms" % (issuer, t(lambda: F.mlp_tc_backward(spec, x, y, hid, dy, ws, bs, dws, dbs, True, workspace=image), reps=20)))
    lib.tune("tc_bwd_issuer", 1)
    print(name, "tc fwd %.3f ms" % t(lambda: F.mlp_tc_forward(spec, x, ws, bs, True, x_stride=stride)),
          "simt fwd %.3f ms" % t(lambda: F.mlp_forward(spec, xs, ws, bs, True)),
          "tc bwd %.3f ms" % t(lambda: F.mlp_tc_backward(spec, x, y, hid, dy, ws, bs, dws, dbs, True)),
          "simt bwd %.3f ms" % t(lambda: F.mlp_backward(spec, xs, ys, hids, dy, ws, bs, dws, dbs, True)))
