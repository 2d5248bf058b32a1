"""Time the front-end block (conv0 + conv1, encoder_front.cuh) with roles switched off: python tools/time_front.py NIMG

Needs the role-profiling build (the product build has no such switches):
    make -C ebsd_vae_b200/csrc OUT=$PWD/ab_libs/libebsd_profile.so BUILD=build_profile EXTRA=-DEBSD_ROLE_PROFILE
    EBSD_B200_LIB=ab_libs/libebsd_profile.so python tools/time_front.py 1184
Flags: 1 no conv0 MMAs, 2 no fp16 conv1 MMAs, 4 no fp8 conv1 MMAs, 32 producers idle, 64 epilogue idle, 128 im2col
builders idle (barriers still cycle, so the pipeline structure is kept).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ebsd_vae_b200 as E
from ebsd_vae_b200 import _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
torch.manual_seed(0)
eng = E.EncoderEngine(E.VariationalAutoEncoderRawData().state_dict(), "cuda")
lib = _native.load()
src = torch.randint(0, 256, (n, 128, 128), dtype=torch.uint8, device="cuda")
src_sums = torch.zeros((n, 32, 2), dtype=torch.float64, device="cuda")
raw = torch.empty((n, 64, 64, 32), device="cuda")
sums = torch.zeros((n, 32, 2), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    _native.check(lib.ebsd_encoder_block(eng._handle, 1, 0, src.data_ptr(), src_sums.data_ptr(), 128 * 128, n,
                                         raw.data_ptr(), sums.data_ptr(), st), "block")
REPS = 10
names = {1: "-conv0", 2: "-f16", 4: "-fp8", 32: "-producers", 64: "-epilogue", 128: "-builders"}
for flags in (0, 1, 2, 4, 6, 7, 32, 64, 128, 32 | 64, 32 | 64 | 128, 7 | 64, 7 | 32, 7 | 32 | 64, 7 | 32 | 64 | 128):
    lib.ebsd_profile_set_flags(flags)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS): run()
    e1.record(); torch.cuda.synchronize()
    label = "".join(v for k, v in names.items() if flags & k) or "all roles"
    print(f"front n {n} flags {flags:3d} {label:40s}: {e0.elapsed_time(e1) / REPS * 1e3:8.1f} us per call (incl. conv0 stats)")
lib.ebsd_profile_set_flags(0)
