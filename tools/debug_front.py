"""Debug build only: run block 1 on a few images and report which bounded wait timed out.
   make -C ebsd_vae_b200/csrc OUT=$PWD/ab_libs/libebsd_debug.so BUILD=build_debug EXTRA=-DEBSD_DEBUG_NOTRAP
   EBSD_B200_LIB=ab_libs/libebsd_debug.so python tools/debug_front.py [NIMG]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ebsd_vae_b200 as E
from ebsd_vae_b200 import _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
eng = E.EncoderEngine(E.VariationalAutoEncoderRawData().state_dict(), "cuda")
lib = _native.load()
src = torch.randint(0, 256, (n, 128, 128), dtype=torch.uint8, device="cuda")
src_sums = torch.zeros((n, 32, 2), dtype=torch.float64, device="cuda")
raw = torch.zeros((n, 64, 64, 32), device="cuda")
sums = torch.zeros((n, 32, 2), dtype=torch.float64, device="cuda")
rc = lib.ebsd_encoder_block(eng._handle, 1, 0, src.data_ptr(), src_sums.data_ptr(), 128 * 128, n, raw.data_ptr(), sums.data_ptr(),
                            torch.cuda.current_stream().cuda_stream)
print("rc", rc)
try:
    torch.cuda.synchronize()
    print("sync ok")
except Exception as e:
    print("sync error:", str(e).splitlines()[0])
info = (ctypes.c_ulonglong * 4)()
print("info rc", lib.ebsd_debug_timeout_info(info))
print("timed out:", info[0], "block", info[1] >> 32, "thread", info[1] & 0xffffffff, "warp", (info[1] & 0xffffffff) // 32,
      "barrier smem addr", hex(info[2]), "parity", info[3])
print("raw nonzero frac", float((raw != 0).float().mean()), "sums", float(sums.abs().sum()))
