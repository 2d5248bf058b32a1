"""Encode a few chunks of synthetic patterns once (profiling target for ncu): python tools/encode_once.py [B] [warm-up reps] [timed reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ebsd_vae_b200 as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(42)
eng = E.EncoderEngine(E.VariationalAutoEncoderRawData().state_dict(), "cuda")
pats = torch.randint(0, 256, (B, 128, 128), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    mu = eng.encode(pats)
torch.cuda.synchronize()
timed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ms = []
for _ in range(timed):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    mu = eng.encode(pats)
    ev1.record()
    torch.cuda.synchronize()
    ms.append(ev0.elapsed_time(ev1))
ms.sort()
print("B", B, "ms min %.3f median %.3f" % (ms[0], ms[len(ms) // 2]), "img/s", B / ms[len(ms) // 2] * 1e3, float(mu.abs().sum()))
