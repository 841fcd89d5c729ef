import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfstudio_b200.lib import call, ptr, stream
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for N in (16, 64):
    for issuers in (1, 2, 3, 4):
        for n in (24, 48):
            if n % issuers: continue
            call("b2n_tc_timing", n | (issuers << 24), N, 20, ptr(out, torch.int64), stream()); torch.cuda.synchronize()
            print(f"N={N} issuers={issuers} n_mma={n}: issue {int(out[0])} cyc, issue->done {int(out[1])} cyc")
