"""Time the consensus stage alone: python tools/time_consensus.py [Q] [N]  (10 candidates per query, headline settings)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ebsd_vae_b200 as E

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
g = torch.Generator(device="cuda").manual_seed(7)
db = E.LatentVectorDatabase()
db.add_vectors(torch.randn((N, 16), generator=g, device="cuda"), torch.rand((N, 3), generator=g, device="cuda", dtype=torch.float64) * 360)
q = db._prepare_queries(torch.randn((Q, 16), generator=g, device="cuda"))
_, idx, _ = db.search_device(q, 10)
for _ in range(3):
    db.consensus_device(idx, 3.0, 5, 3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = db.consensus_device(idx, 3.0, 5, 3)
e1.record()
torch.cuda.synchronize()
print(f"consensus Q {Q}: {e0.elapsed_time(e1) / 20:.4f} ms per call, success rate {out[2].float().mean().item():.3f}")
