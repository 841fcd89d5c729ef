import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfstudio_b200.lib import call, ptr, stream
def run(mode, three, A, B, M, N, K):
    out = torch.full((128, N), float('nan'), device='cuda')
    call("b2n_tc_selftest", mode, three, ptr(A), A.shape[0], A.shape[1], ptr(B), B.shape[0], B.shape[1], M, N, K, ptr(out), stream())
    torch.cuda.synchronize()
    return out.cpu()
torch.set_printoptions(linewidth=250, precision=0, sci_mode=False)
K, N = 16, 16
W = (torch.arange(K)[:, None] * 100 + torch.arange(N)[None, :]).float().cuda()   # W[k][n] = 100k+n
for k0 in (0, 1, 5, 9):
    A = torch.zeros(128, K, device='cuda'); A[:, k0] = 1
    D = run(1, 0, A, W, 128, N, K)
    print(f"mode1 k0={k0}: D[0,:]=", D[0].tolist())
    print(f"           D[3,:]=", D[3].tolist())
# mode 2: D[m][n] = sum_r A[r][m] B[r][n]; A = one-hot row r0 -> D[m][n] = A[r0][m] * B[r0][n]
Ma, N = 64, 16
B = (torch.arange(128)[:, None] * 100 + torch.arange(N)[None, :]).float().cuda()
for r0, m0 in ((0, 0), (5, 3), (17, 9)):
    A = torch.zeros(128, Ma, device='cuda'); A[r0, m0] = 1
    D = run(2, 0, A, B, 64, N, 128)
    nz = torch.nonzero(D.nan_to_num(0))
    print(f"mode2 r0={r0} m0={m0}: nonzero lanes {sorted(set(nz[:,0].tolist()))[:10]} values row:", D[nz[0,0]].tolist() if len(nz) else None, "expected", B[r0].tolist())
