// Generic runtime-parameterised GEMM epilogue shared by the tcgen05 and SIMT GEMM kernels.
//
// For an accumulator element acc(m,n) of D = A * B^T the epilogue computes, in this order,
//   v  = acc [+ bias[n]] [+ rowvec[m]*colvec[n]]
//   out_pre[m,n] = v                                   (optional fp32 store, full N)
//   w  = aux_sig ? v * softplus100'(aux_sig[m,n]) : act(v)
//   out2[m,n] = scale * v * aux_q[m,n] * softplus100''(aux_sig[m,n])   (optional)
//   w  = aux_relu ? (aux_relu[m,n] > 0 ? w : 0) : w
//   w  = w * scale [+ aux_add[m,n]]
//   out_f32[m,n] (=|+=) w ; planes(out_pl)[m,n] = split_bf16(w)       (columns < n_store)
// which covers every fused layer of the SDF / colour / background MLPs and their hand-derived
// backward passes (DESIGN.md "GEMM call sites").
#pragma once
#include "common.cuh"

namespace nrw {

enum { ACT_NONE = 0, ACT_SOFTPLUS100 = 1, ACT_RELU = 2, ACT_SIGMOID = 3 };

struct Epi {
  const float* bias = nullptr;
  const float* rowvec = nullptr;
  const float* colvec = nullptr;
  const float* aux_sig = nullptr;
  const float* aux_q = nullptr;
  const float* aux_add = nullptr;
  int aux_q_bcast = 0;  // aux_q is a [N] row vector broadcast over rows
  int ld_aux = 0;  // shared by aux_sig / aux_q / aux_add
  const bf16* aux_relu = nullptr;
  int ld_relu = 0;
  int act = ACT_NONE;
  float scale = 1.0f;
  float* out_pre = nullptr;
  int ld_pre = 0;
  float* out_f32 = nullptr;
  int ld_f32 = 0;
  int atomic = 0;
  float* out2 = nullptr;
  int ld_out2 = 0;
  Planes out_pl = {nullptr, 0, 0};
  int n_planes = 0;
  int n_store = 1 << 30;  // column bound for out_f32 / out_pl / out2
};

template <int NC>
__device__ __forceinline__ void epi_apply(const Epi& e, int m, int n0, float (&acc)[NC], int N) {
  const int n_hi = min(N, n0 + NC);         // bound for out_pre
  const int ns_hi = min(e.n_store, n_hi);   // bound for main outputs
  const float rv = e.rowvec ? e.rowvec[m] : 0.0f;
  float w[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int n = n0 + j;
    float v = acc[j];
    if (n < n_hi) {
      if (e.bias) v += e.bias[n];
      if (e.rowvec) v += rv * e.colvec[n];
    }
    acc[j] = v;
  }
  if (e.out_pre) {
    float* dst = e.out_pre + (long long)m * e.ld_pre + n0;
    if (n0 + NC <= n_hi && (NC % 4) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < NC; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (n0 + j < n_hi) dst[j] = acc[j];
    }
  }
  if (n0 >= ns_hi) return;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int n = n0 + j;
    const bool ok = n < ns_hi;
    float v = acc[j];
    float r;
    if (e.aux_sig) {
      const float a = ok ? e.aux_sig[(long long)m * e.ld_aux + n] : 0.0f;
      r = v * softplus100_d1(a);
      if (e.out2 && ok) {
        const float q = e.aux_q_bcast ? e.aux_q[n] : e.aux_q[(long long)m * e.ld_aux + n];
        e.out2[(long long)m * e.ld_out2 + n] = e.scale * v * q * softplus100_d2(a);
      }
    } else {
      switch (e.act) {
        case ACT_SOFTPLUS100: r = softplus100(v); break;
        case ACT_RELU: r = fmaxf(v, 0.0f); break;
        case ACT_SIGMOID: r = sigmoidf_(v); break;
        default: r = v;
      }
    }
    if (e.aux_relu && ok) {
      if (!(__bfloat162float(e.aux_relu[(long long)m * e.ld_relu + n]) > 0.0f)) r = 0.0f;
    }
    r *= e.scale;
    if (e.aux_add && ok) r += e.aux_add[(long long)m * e.ld_aux + n];
    w[j] = r;
  }
  if (e.out_f32) {
    float* dst = e.out_f32 + (long long)m * e.ld_f32 + n0;
    if (e.atomic) {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (n0 + j < ns_hi) atomicAdd(dst + j, w[j]);
    } else if (n0 + NC <= ns_hi && (NC % 4) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < NC; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (n0 + j < ns_hi) dst[j] = w[j];
    }
  }
  if (e.n_planes > 0) {
    const long long base = (long long)m * e.out_pl.ld + n0;
    const bool vec = (n0 + NC <= ns_hi) && (NC % 8) == 0 &&
                     ((reinterpret_cast<uintptr_t>(e.out_pl.p + base) & 15) == 0) &&
                     ((e.out_pl.pstride & 7) == 0);
    float res[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) res[j] = w[j];
    for (int pl = 0; pl < e.n_planes; ++pl) {
      bf16* dst = e.out_pl.plane(pl) + base;
      if (vec) {
        if constexpr ((NC % 8) == 0) {
#pragma unroll
          for (int j = 0; j < NC; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              bf16 lo = __float2bfloat16_rn(res[j + 2 * t]);
              bf16 hi = __float2bfloat16_rn(res[j + 2 * t + 1]);
              res[j + 2 * t] -= __bfloat162float(lo);
              res[j + 2 * t + 1] -= __bfloat162float(hi);
              pk[t] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
            }
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          bf16 b = __float2bfloat16_rn(res[j]);
          res[j] -= __bfloat162float(b);
          if (n0 + j < ns_hi) dst[j] = b;
        }
      }
    }
  }
}

}  // namespace nrw
