"""Time one conv layer of the tensor path under the profiling switches: python tools/time_layer.py LAYER NIMG"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ebsd_vae_b200 as E
from ebsd_vae_b200 import _native

PLAN = {1: (32, 32, 128), 2: (32, 64, 64), 3: (64, 64, 64), 4: (64, 128, 32), 5: (128, 128, 32)}
layer = int(sys.argv[1]); n = int(sys.argv[2])
cin, cout, hw = PLAN[layer]
torch.manual_seed(0)
eng = E.EncoderEngine(E.VariationalAutoEncoderRawData().state_dict(), "cuda")
lib = _native.load()
act = torch.randn((n, hw, hw, cin), device="cuda")
raw = torch.empty((n, hw, hw, cout), device="cuda")
sums = torch.zeros((n, cout, 2), dtype=torch.float64, device="cuda")
ws = torch.empty(n * (hw + 2) * (hw + 2) * cin * 4 + 4096, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    _native.check(lib.ebsd_debug_conv_layer(eng._handle, layer, 2, act.data_ptr(), n, raw.data_ptr(), sums.data_ptr(),
                                            ws.data_ptr(), ws.numel(), st), "dbg")
REPS = 1 if os.environ.get('NCU') else 10
for flags in (0, 1, 2, 4, 8, 16, 4 | 2, 8 | 4 | 2, 16 | 8 | 4 | 2):
    lib.ebsd_debug_set_flags(flags)
    for _ in range(1 if os.environ.get('NCU') else 3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS): run()
    e1.record(); torch.cuda.synchronize()
    print(f"layer {layer} n {n} flags {flags:2d}: {e0.elapsed_time(e1) / REPS * 1e3:8.1f} us per call (incl. split/pad + memset)")
lib.ebsd_debug_set_flags(0)
