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
K, N = 16, 32
W = torch.zeros(K, N, device='cuda')
for swap in (0, 4):
    for k0 in (0, 1, 2, 8, 9):
        A = torch.zeros(128, K, device='cuda'); A[:, k0] = 1
        lo = run(1, (1 | swap) << 1, A, W, 128, N, K)[0]
        hi = run(1, (2 | swap) << 1, A, W, 128, N, K)[0]
        words = (hi * 1024 + lo).long().tolist()
        print(f"mode1 (B MN-major) swap={swap} k0={k0}: byte offsets read for n=0..{N-1}:", [w * 4 for w in words])
# K-major reference decode (mode 0): D[r][n] = sum_k A[r][k] B[n][k]
for k0 in (0, 1, 4, 8):
    A = torch.zeros(128, K, device='cuda'); A[:, k0] = 1
    Bz = torch.zeros(N, K, device='cuda')
    lo = run(0, 1 << 1, A, Bz, 128, N, K)[0]; hi = run(0, 2 << 1, A, Bz, 128, N, K)[0]
    print(f"mode0 (B K-major) k0={k0}: byte offsets for n=0..:", [int(w) * 4 for w in (hi * 1024 + lo).tolist()])
