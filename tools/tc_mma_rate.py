"""Developer aid (GPU box): per-instruction cost of tcgen05.mma kind::tf32 M=64 in the no-swizzle K-major layout."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "two-for-one-diffusion_b200")]
import torch
from dff_b200 import _native as nat
lib = nat.lib()
for n in (64, 128, 192, 256):
    for k in (64, 256):
        if (64 + n) * k * 8 > 200 * 1024:
            continue
        a = torch.randn(64, k); b = torch.randn(n, k); d = torch.zeros(64, n); ms = C.c_float()
        reps = 4000
        nat.check(lib.dff_debug_tc_gemm(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(d.data_ptr()), n, k, reps, C.byref(ms)))
        mmas = reps * (k // 8) * 3
        err = float((d - a @ b.T).abs().max() / (a @ b.T).abs().max())
        print(f"N={n:3d} K={k:3d}: {ms.value*1e6/mmas:7.1f} ns/MMA = {ms.value*1e6/mmas*1.965:6.0f} cyc  (floor {n/2:.0f} cyc)  err {err:.1e}")
