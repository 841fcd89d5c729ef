import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfstudio_b200 import functional as F
from oracle import nerf_oracle as O
torch.manual_seed(1)
in_dim, dims, out_act = 63, [64, 64, 3], "sigmoid"
for n in (128, 128*3, 128*300+77):
    spec = F.MlpSpec(in_dim, dims, out_act=out_act)
    x = torch.randn(n, in_dim)
    ws, prev = [], in_dim
    for d in dims:
        ws.append(torch.randn(d, prev) * (1.5 / prev ** 0.5)); prev = d
    bs = [torch.randn(d) * 0.1 for d in dims]
    dy = torch.randn(n, dims[-1])
    xo = x.clone().requires_grad_(True)
    wl = [w.clone().requires_grad_(True) for w in ws]; bl = [b.clone().requires_grad_(True) for b in bs]
    yo = O.mlp_forward(xo, wl, bl, out_act=out_act)
    go = torch.autograd.grad(yo, [xo] + wl + bl, dy)
    xc, wc, bc = x.cuda(), [w.cuda() for w in ws], [b.cuda() for b in bs]
    for rep in range(2):
        y, hidden = F.mlp_tc_forward(spec, xc, wc, bc, save_hidden=True)
        dws, dbs = [torch.zeros_like(w) for w in wc], [torch.zeros_like(b) for b in bc]
        dx = F.mlp_tc_backward(spec, xc, y, hidden, dy.cuda(), wc, bc, dws, dbs, want_dx=True)
        torch.cuda.synchronize()
        e = (dx.cpu() - go[0]).abs()
        bad_rows = (e.max(1).values > 1e-3).nonzero()[:, 0]
        bad_cols = (e.max(0).values > 1e-3).nonzero()[:, 0]
        print(f"n={n} rep={rep} y err {(y.cpu()-yo).abs().max():.2e} dx err {e.max():.2e} bad rows {len(bad_rows)} (first {bad_rows[:8].tolist()}) bad cols {bad_cols.tolist()[:70]}")
        for i in range(3):
            print(f"   dw{i} err {(dws[i].cpu()-go[1+i]).abs().max()/go[1+i].abs().max():.2e} db{i} err {(dbs[i].cpu()-go[4+i]).abs().max()/go[4+i].abs().max():.2e}")
