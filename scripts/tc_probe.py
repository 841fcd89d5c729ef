"""Probe of the tcgen05 self-test on a B200 (development aid): prints max errors per mode / shape."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfstudio_b200.lib import call, ptr, stream

torch.manual_seed(0)
def run(mode, three, A, B, M, N, K):
    out = torch.full((128, N), float('nan'), device='cuda')
    call("b2n_tc_selftest", mode, three, ptr(A), A.shape[0], A.shape[1], ptr(B), B.shape[0], B.shape[1], M, N, K, ptr(out), stream())
    torch.cuda.synchronize()
    return out

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())

for three in (0, 1):
    for (N, K) in ((16, 8), (64, 64), (16, 64), (64, 32), (32, 16)):
        A = torch.randn(128, K, device='cuda'); B = torch.randn(N, K, device='cuda')
        D = run(0, three, A, B, 128, N, K)
        print(f"mode0 three={three} N={N} K={K}: rel err {rel(D, A.double() @ B.double().T):.3e}")
    for (N, K) in ((64, 64), (32, 16), (64, 16), (16, 64)):
        A = torch.randn(128, K, device='cuda'); W = torch.randn(K, N, device='cuda')
        D = run(1, three, A, W, 128, N, K)
        print(f"mode1 three={three} N={N} K={K}: rel err {rel(D, A.double() @ W.double()):.3e}")
    for (Ma, N) in ((64, 64), (64, 32), (64, 16)):
        A = torch.randn(128, Ma, device='cuda'); B = torch.randn(128, N, device='cuda')
        ref = A.double().T @ B.double()
        D = run(2, three, A, B, 64, N, 128)
        # M=64: which TMEM lanes hold the 64 rows?
        best = None
        for name, rows in (("lanes0-63", list(range(64))), ("16-per-warp", [32 * (i // 16) + i % 16 for i in range(64)])):
            e = rel(D[rows], ref)
            print(f"mode2 M=64 three={three} Ma={Ma} N={N} layout {name}: rel err {e:.3e}")
    # stacked M=128 (two 64-column halves of A as one MN operand)
    A = torch.randn(128, 64, device='cuda'); B = torch.randn(128, 64, device='cuda')
    D = run(2, three, A, B, 128, 64, 128) if False else None
A = torch.randn(128, 128 if False else 64, device='cuda')
