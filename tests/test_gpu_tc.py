"""GPU: the tcgen05 (UMMA) building block -- split-precision TF32 GEMM with the accumulator in TMEM -- against fp64."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(n, k, reps=1, seed=0):
    from dff_b200 import _native as nat
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(64, k, generator=g)
    b = torch.randn(n, k, generator=g)
    d = torch.empty(64, n)
    ms = C.c_float()
    nat.check(nat.lib().dff_debug_tc_gemm(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(d.data_ptr()), n, k, reps,
                                          C.c_void_p(C.addressof(ms))))
    ref = a.double() @ b.double().t()
    err = float((d.double() - ref).abs().max() / ref.abs().max())
    return err, ms.value


@pytest.mark.parametrize("n,k", [(64, 64), (128, 64), (192, 64), (64, 192), (128, 128), (256, 32), (8, 8), (72, 40)])
def test_tcgen05_3xtf32_gemm_matches_fp64(n, k):
    err, _ = _run(n, k)
    assert err < 2e-6, (n, k, err)          # fp32-grade: single-pass TF32 would be ~5e-4


def test_tcgen05_throughput_report():
    n, k, reps = 192, 64, 2000
    err, ms = _run(n, k, reps)
    macs = 64.0 * n * k * reps
    print(f"\ntcgen05 3xTF32 [64x{n}x{k}] x{reps}: {ms:.3f} ms, {macs / (ms * 1e-3) / 1e9:.1f} useful GMAC/s on ONE SM")
    assert err < 2e-6
