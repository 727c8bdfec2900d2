// GEMM front-end: D[M,N] = sum over plane products of A_p * B_q^T, fused epilogue (epilogue.cuh).
#pragma once
#include "epilogue.cuh"

namespace nrw {

struct GemmDesc {
  Planes A;           // mn_major=0: [M,K] rows, K contiguous.  mn_major=1: [K,M] rows, M contiguous
  Planes B;           // mn_major=0: [N,K] rows, K contiguous.  mn_major=1: [K,N] rows, N contiguous
  int n_planes = 1;   // split-precision planes used from A and B (1, 2 or 3)
  int M = 0, N = 0, K = 0;
  int mn_major = 0;
  int k_slices = 1;   // split-K (epilogue must be atomic)
  Epi epi;
};

enum GemmBackend { GEMM_TCGEN05 = 0, GEMM_SIMT = 1 };

// tcgen05 / TMEM / TMA implementation (gemm_tc.cu)
int gemm_tc(const GemmDesc& g, cudaStream_t stream);
// fp32 CUDA-core implementation over the same operands (gemm_simt.cu); verification backend
int gemm_simt(const GemmDesc& g, cudaStream_t stream);

inline int gemm(int backend, const GemmDesc& g, cudaStream_t stream) {
  return backend == GEMM_SIMT ? gemm_simt(g, stream) : gemm_tc(g, stream);
}

// number of (a_plane, b_plane) products issued for a given plane count: 1, 3, 6
inline int n_products(int n_planes) { return n_planes == 1 ? 1 : (n_planes == 2 ? 3 : 6); }

// Fused forward-only SDF chain (gemm_tc.cu::sdf_fused_kernel): points -> sdf, activations resident in shared memory.
// W[l] = packed K-major weights of SDF layer l ([512 x Kp] bf16, two planes), bias[l] fp32 [512]; head_w / head_b = lin8 row 0.
struct SdfFusedDesc {
  const float* pts = nullptr;   // [M, 3]
  float* sdf = nullptr;         // [M]
  int M = 0;
  Planes W[8];
  const float* bias[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const float* head_w = nullptr;
  const float* head_b = nullptr;
};
int sdf_fused_forward(const SdfFusedDesc& d, cudaStream_t stream);

long long gemm_tc_launch_count();
// true when gemm_tc runs this split-K weight-gradient shape (M x N output) on 256 x 512 pair tiles: the caller sizes k_slices for
// one item per CTA pair
bool gemm_tc_wide_dw(int M, int N, int n_planes);
// debug: when non-null, every tcgen05 GEMM launch accumulates per-CTA cycle attribution into buf[148*8]
void gemm_tc_set_profile_buffer(unsigned long long* buf);
// measurement: CUDA events around every tcgen05 GEMM launch (on the launching stream)
void gemm_tc_timing_enable(bool on);
int gemm_tc_timing_read(double* ms, double* flops, double* mma_flops, long long* launches, double* bytes);

}  // namespace nrw
