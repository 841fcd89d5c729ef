"""Per-phase cycle counts of the tensor-core MLP forward kernel (library built with B2N_NVCC_EXTRA=-DB2N_TC_PROF)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfstudio_b200 import functional as F, lib
L = lib.load()
CFG = {"head": (63, [64, 64, 3], "sigmoid", 196608), "base": (32, [64, 16], "none", 196608)}
names_b = ["load_dz_a", "store_Z", "stage_T", "sync_pre_mma", "mma_issue", "mma_wait", "ld_dA", "sync_post", ] + [f"p{i}" for i in range(8, 16)]
names = ["issuer_wait_full", "issuer_issue", "w_stage_x", "w_wait_done", "w_epilogue", "p5", "p6"] + [f"p{i}" for i in range(7, 16)]
for name in sys.argv[1:] or list(CFG):
    in_dim, dims, oact, n = CFG[name]
    spec = F.MlpSpec(in_dim, dims, out_act=oact)
    stride = (in_dim + 3) // 4 * 4
    x = torch.zeros(n, stride, device="cuda"); x[:, :in_dim] = torch.randn(n, in_dim, device="cuda")
    ws, prev = [], in_dim
    for d in dims:
        ws.append(torch.randn(d, prev, device="cuda") / prev ** 0.5); prev = d
    bs = [torch.zeros(d, device="cuda") for d in dims]
    y, hid = F.mlp_tc_forward(spec, x, ws, bs, True, x_stride=stride)
    dy = torch.randn(n, dims[-1], device="cuda")
    dws, dbs = [torch.zeros_like(w) for w in ws], [torch.zeros_like(b) for b in bs]
    out = (C.c_ulonglong * 16)()
    for kind in ("fwd", "bwd"):
        L.b2n_tc_prof_read(out, 1)
        reps = 5
        for _ in range(reps):
            if kind == "fwd": F.mlp_tc_forward(spec, x, ws, bs, True, x_stride=stride)
            else: F.mlp_tc_backward(spec, x, y, hid, dy, ws, bs, dws, dbs, True)
        L.b2n_tc_prof_read(out, 1)
        tot = sum(out)
        tiles = (n // 128 + 147) // 148
        print(name, kind, "cycles/launch %d (%.1f us @1.965GHz), tiles/CTA %d" % (tot / reps, tot / reps / 1965, tiles))
        for nm, v in zip(names if kind == 'fwd' else names_b, out):
            if v: print("   %-14s %8d cyc/launch  %6d cyc/tile  %5.1f%%" % (nm, v / reps, v / reps / tiles, 100 * v / tot))
