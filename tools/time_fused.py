"""Time one fused block with roles switched off: python tools/time_fused.py LAYER NIMG

Needs the role-profiling build of the library (the product build has no such switches):
    make -C ebsd_vae_b200/csrc OUT=$PWD/ab_libs/libebsd_profile.so BUILD=build_profile EXTRA=-DEBSD_ROLE_PROFILE
    EBSD_B200_LIB=ab_libs/libebsd_profile.so python tools/time_fused.py 1 1184
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ebsd_vae_b200 as E
from ebsd_vae_b200 import _native

PLAN = {1: (32, 32, 128, 1), 2: (32, 64, 64, 0), 3: (64, 64, 64, 1), 4: (64, 128, 32, 0), 5: (128, 128, 32, 1),
        6: (128, 128, 16, 0), 7: (128, 128, 16, 1), 8: (128, 128, 8, 0), 9: (128, 128, 8, 1)}
layer = int(sys.argv[1]); n = int(sys.argv[2])
cin, cout, hw, pool = PLAN[layer]
torch.manual_seed(0)
eng = E.EncoderEngine(E.VariationalAutoEncoderRawData().state_dict(), "cuda")
lib = _native.load()
if layer == 1:
    src = torch.randint(0, 256, (n, 128, 128), dtype=torch.uint8, device="cuda")
    src_sums = torch.zeros((n, 32, 2), dtype=torch.float64, device="cuda")
else:
    src = torch.randn((n, hw, hw, cin), device="cuda")
    x = src.double()
    src_sums = torch.stack([x.sum(dim=(1, 2)), (x * x).sum(dim=(1, 2))], dim=2).contiguous()
ho = hw // 2 if pool else hw
raw = torch.empty((n, ho, ho, cout), device="cuda")
sums = torch.zeros((n, cout, 2), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    _native.check(lib.ebsd_encoder_block(eng._handle, layer, 0, src.data_ptr(), src_sums.data_ptr(), hw * hw, n,
                                             raw.data_ptr(), sums.data_ptr(), st), "dbg")
REPS = 10
for flags in (0, 1, 2, 4, 8, 16, 32, 1 | 2, 1 | 4, 2 | 4, 1 | 2 | 4):
    lib.ebsd_profile_set_flags(flags)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS): run()
    e1.record(); torch.cuda.synchronize()
    print(f"layer {layer} n {n} flags {flags:2d}: {e0.elapsed_time(e1) / REPS * 1e3:8.1f} us per call" + (" (incl. conv0 stats)" if layer == 1 else ""))
lib.ebsd_profile_set_flags(0)
