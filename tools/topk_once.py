"""Run the exact top-k once (ncu target / timing): python tools/topk_once.py [N] [Q] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ebsd_vae_b200 as E

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(2024)
lat = torch.randn((N, 16), generator=g, device="cuda")
eul = torch.rand((N, 3), generator=g, device="cuda", dtype=torch.float64) * 360
db = E.LatentVectorDatabase()
db.add_vectors(lat, eul)
q = db._prepare_queries(lat[:Q] + 0.05 * torch.randn((Q, 16), generator=g, device="cuda"))
for _ in range(2):
    db.search_device(q, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    dot, idx, dist = db.search_device(q, 10)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"N {N} Q {Q}: {ms:.3f} ms per search, {Q / ms * 1e3:.0f} queries/s, {32.0 * Q * N / ms / 1e9:.2f} TFLOP/s fp32, {(64.0 * N + 64.0 * Q + 120.0 * Q) / ms / 1e6:.0f} GB/s algorithmic, self-hit {(idx[:, 0] == torch.arange(Q, device='cuda')).float().mean().item():.3f}")
