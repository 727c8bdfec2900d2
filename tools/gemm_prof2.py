"""Cycle attribution of the 2-CTA tcgen05 GEMM roles for a forward-layer shaped launch (NRW_GEMM_TEST_LAYER=1 for the layer stores)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neuralrecon-w_b200"))
import torch
from nrw import _lib
L = _lib.lib(); dev = "cuda"
M, N, K, planes, act = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "262144,512,512,2,1").split(",")]
A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev) / 30
bi = torch.zeros(N, device=dev); D = torch.zeros(M, N, device=dev)
scratch = torch.empty(L.nrw_gemm_test_scratch_bytes(M, N, K) + 1024, dtype=torch.uint8, device=dev)
sp = (scratch.data_ptr() + 1023) // 1024 * 1024
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
for it in range(3):
    buf.zero_(); torch.cuda.synchronize()
    L.nrw_debug_gemm_profile(C.c_void_p(buf.data_ptr()))
    _lib.check(L.nrw_gemm_test(0, planes, 0, 1, M, N, K, _lib.ptr(A), _lib.ptr(B), _lib.ptr(bi), act, _lib.ptr(D), C.c_void_p(sp), _lib.stream_ptr()), "gemm")
    torch.cuda.synchronize()
    L.nrw_debug_gemm_profile(None)
b = buf.view(148, 16).double().cpu()
lead, peer = b[0::2], b[1::2]
tot = b[:, 5].mean()
print(f"kernel {tot/1.9e3:.1f} us  tiles/pair {lead[:,6].mean():.1f}")
for name, x in (("leader", lead), ("peer", peer)):
    t = x[:, 5].mean()
    print(f" {name}: prod_wait_stage {100*x[:,0].mean()/t:5.1f}%  mma_wait_tma {100*x[:,1].mean()/t:5.1f}%  mma_wait_acc {100*x[:,2].mean()/t:5.1f}%  "
          f"epi_wait_acc {100*x[:,3].mean()/t:5.1f}%  epi_busy {100*x[:,4].mean()/t:5.1f}%  (epi busy/tile {x[:,4].sum()/x[:,6].sum()/1.9e3:.2f} us)")
