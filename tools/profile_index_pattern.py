"""Where the wall time of index_pattern goes (host side): python tools/profile_index_pattern.py [ROWS] [CALLS]"""
import cProfile, os, pstats, sys, time, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ebsd_vae_b200 as E

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 300
logging.getLogger("ebsd_vae_b200.vector_db").setLevel(logging.ERROR)
torch.manual_seed(42)
model = E.VariationalAutoEncoderRawData()
indexer = E.DiffractionPatternIndexer(model, db=E.LatentVectorDatabase(E.LatentVectorDatabaseConfig(persist_directory=None)),
                                      config=E.IndexerConfig(device="cuda", top_n=10))
g = torch.Generator(device="cuda").manual_seed(1)
indexer.db.add_vectors(torch.randn((rows, 16), generator=g, device="cuda"), torch.rand((rows, 3), device="cuda", dtype=torch.float64) * 180)
pat = np.random.default_rng(0).random((128, 128)).astype(np.float32)
for _ in range(20):
    indexer.index_pattern(pat)
t = []
for _ in range(calls):
    t0 = time.perf_counter(); indexer.index_pattern(pat); t.append(time.perf_counter() - t0)
t.sort()
print("index_pattern wall ms: median %.3f min %.3f" % (1e3 * t[len(t) // 2], 1e3 * t[0]))
pr = cProfile.Profile()
pr.enable()
for _ in range(calls):
    indexer.index_pattern(pat)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(32)
