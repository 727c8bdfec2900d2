/* nrw_math.h - written-down fp32 primitives with a FIXED operation order.
 *
 * The voxel-guided sampler produces integer artefacts (searchsorted indices, merge order) from
 * fp32 arithmetic; "bit-exact" is only meaningful against a restatement that uses the same
 * operation order and the same exp().  Every operation below is a single IEEE-754 binary32
 * operation (no contraction): the CUDA build uses the _rn intrinsics, the host build must be
 * compiled with -ffp-contract=off.  fmaf is used only where written explicitly.
 */
#ifndef NRW_MATH_H_
#define NRW_MATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define NRW_HD __host__ __device__ __forceinline__
#else
#define NRW_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define NRW_MUL(a, b) __fmul_rn((a), (b))
#define NRW_ADD(a, b) __fadd_rn((a), (b))
#define NRW_SUB(a, b) __fsub_rn((a), (b))
#define NRW_DIV(a, b) __fdiv_rn((a), (b))
#define NRW_SQRT(a) __fsqrt_rn((a))
#define NRW_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define NRW_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define NRW_ADD(a, b) ((float)((float)(a) + (float)(b)))
#define NRW_SUB(a, b) ((float)((float)(a) - (float)(b)))
#define NRW_DIV(a, b) ((float)((float)(a) / (float)(b)))
#define NRW_SQRT(a) sqrtf((a))
#define NRW_FMA(a, b, c) fmaf((a), (b), (c))
#endif

/* exp(x): k = rint(x*log2e); r = x - k*ln2 (2-term Cody-Waite); degree-6 Horner; scale by 2^k. */
NRW_HD float nrw_exp_f32(float x) {
  if (x > 88.0f) x = 88.0f;
  if (x < -86.0f) x = -86.0f;
  const float k = rintf(NRW_MUL(x, 1.44269504088896341f));
  float r = NRW_FMA(k, -0.693145751953125f, x);
  r = NRW_FMA(k, -1.42860682030941723212e-6f, r);
  float p = 1.0f / 720.0f;
  p = NRW_FMA(p, r, 1.0f / 120.0f);
  p = NRW_FMA(p, r, 1.0f / 24.0f);
  p = NRW_FMA(p, r, 1.0f / 6.0f);
  p = NRW_FMA(p, r, 0.5f);
  p = NRW_FMA(p, r, 1.0f);
  p = NRW_FMA(p, r, 1.0f);
  int32_t bits;
  memcpy(&bits, &p, 4);
  bits += ((int32_t)k) << 23;
  float out;
  memcpy(&out, &bits, 4);
  return out;
}

NRW_HD float nrw_sigmoid_f32(float x) {
  return NRW_DIV(1.0f, NRW_ADD(1.0f, nrw_exp_f32(-x)));
}

/* torch.linspace(start, end, steps)[i] for float32 (ATen RangeFactories: symmetric evaluation) */
NRW_HD float nrw_linspace_f32(float start, float end, int steps, int i) {
  if (steps == 1) return start;
  const float step = NRW_DIV(NRW_SUB(end, start), (float)(steps - 1));
  const int halfway = steps / 2;
  if (i < halfway) return NRW_ADD(start, NRW_MUL(step, (float)i));
  return NRW_SUB(end, NRW_MUL(step, (float)(steps - i - 1)));
}

#endif /* NRW_MATH_H_ */
