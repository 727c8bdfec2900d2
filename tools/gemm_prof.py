"""Cycle attribution of the tcgen05 GEMM roles (producer / MMA / epilogue) via nrw_debug_gemm_profile."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neuralrecon-w_b200"))
import torch
from nrw import _lib
L = _lib.lib(); dev = "cuda"
def run(M, N, K, planes, act, mn=0, ks=1):
    if mn: A = torch.randn(K, M, device=dev); B = torch.randn(K, N, device=dev) / 30
    else:  A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev) / 30
    bi = torch.zeros(N, device=dev); D = torch.zeros(M, N, device=dev)
    scratch = torch.empty(L.nrw_gemm_test_scratch_bytes(M, N, K) + 1024, dtype=torch.uint8, device=dev)
    sp = (scratch.data_ptr() + 1023) // 1024 * 1024
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    for it in range(3):
        buf.zero_(); torch.cuda.synchronize()
        L.nrw_debug_gemm_profile(C.c_void_p(buf.data_ptr()))
        _lib.check(L.nrw_gemm_test(0, planes, mn, ks, M, N, K, _lib.ptr(A), _lib.ptr(B), _lib.ptr(bi), act, _lib.ptr(D), C.c_void_p(sp), _lib.stream_ptr()), "gemm")
        torch.cuda.synchronize()
        L.nrw_debug_gemm_profile(None)
    b = buf.view(148, 16).double().cpu()
    tiles = b[:, 6].clamp(min=1)
    tot = b[:, 5].mean()
    pt = lambda s: (b[:, s] / tiles).mean() / 1.9e3
    print(f"M={M} N={N} K={K} P={planes} act={act} mn={mn} ks={ks}: kernel {tot/1.9e3:6.1f}us | %kernel: prod_wait {100*b[:,0].mean()/tot:4.1f} mma_wait_full {100*b[:,1].mean()/tot:4.1f} "
          f"mma_wait_acc {100*b[:,2].mean()/tot:4.1f} epi_wait {100*b[:,3].mean()/tot:4.1f} epi_busy {100*b[:,4].mean()/tot:4.1f} | per tile us: busy {pt(4):.2f} ldtm {pt(8):.2f} bias+pre {pt(9):.2f} aux {pt(10):.2f} math {pt(11):.2f} stores {pt(12):.2f}")
run(32768, 512, 512, 2, 1)
run(32768, 512, 512, 2, 0)
run(32768, 512, 64, 2, 1)
run(512, 512, 32768, 2, 0, mn=1, ks=18)
