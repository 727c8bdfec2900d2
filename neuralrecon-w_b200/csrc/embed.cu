// Input encodings of the three MLPs and column sums, one thread per OUTPUT element so every store is
// coalesced (the sample-per-thread versions were 10x slower: strided 2-byte stores).
#include "pointwise.h"

namespace nrw {

static constexpr float INV_SQRT2 = 0.70710678118654752440f;

// value j of the 3-D positional encoding [x, sin(2^k x), cos(2^k x)]_k (models/neuconw.py:7-55)
__device__ __forceinline__ float pe3(const float* x, int j) {
  if (j < 3) return x[j];
  const int k = (j - 3) / 6, r = (j - 3) % 6, c = r % 3;
  const float t = x[c] * (float)(1 << k);
  return r >= 3 ? cosf(t) : sinf(t);
}
// 4-D variant (models/nerf.py:8-39)
__device__ __forceinline__ float pe4(const float* x, int j) {
  if (j < 4) return x[j];
  const int k = (j - 4) / 8, r = (j - 4) % 8, c = r % 4;
  const float t = x[c] * (float)(1 << k);
  return r >= 4 ? cosf(t) : sinf(t);
}

// U0[m, 0:64] = [PE6(x) 39 | 0];  U4[m, 473:512] = PE6(x) / sqrt(2).  128 threads per sample.
__global__ void __launch_bounds__(256) sdf_embed_kernel(const float* __restrict__ pts, int M, int n_planes, Planes U0,
                                                        Planes U4) {
  const int m = blockIdx.x * 2 + (threadIdx.x >> 7), j = threadIdx.x & 127;
  if (m >= M) return;
  const float x[3] = {pts[m * 3], pts[m * 3 + 1], pts[m * 3 + 2]};
  if (j < 64) {
    planes_store(U0, n_planes, (long long)m * U0.ld + j, j < 39 ? pe3(x, j) : 0.0f);
  } else if (j < 64 + 39 && U4.p) {
    planes_store(U4, n_planes, (long long)m * U4.ld + 473 + (j - 64), pe3(x, j - 64) * INV_SQRT2);
  }
}
int launch_sdf_embed(const float* pts, int M, int n_planes, Planes U0, Planes U4, cudaStream_t s) {
  sdf_embed_kernel<<<cdiv(M, 2), 256, 0, s>>>(pts, M, n_planes, U0, U4);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// IN1[m, 512:640] = [viewPE4(d) 27 | a n_a | 0];  IN2[m, 128:192] = [pts 3 | normal 3 | 0].  192 threads/sample.
__global__ void __launch_bounds__(192) color_embed_kernel(const float* __restrict__ dirs, const float* __restrict__ a,
                                                          int n_a, int rows_per_src, const float* __restrict__ pts,
                                                          const float* __restrict__ nrm, int M, int n_planes, Planes IN1,
                                                          Planes IN2) {
  const int m = blockIdx.x, j = threadIdx.x;
  if (m >= M) return;
  const int r = m / rows_per_src;
  if (j < 128) {
    float v = 0.0f;
    if (j < 27) {
      const float x[3] = {dirs[r * 3], dirs[r * 3 + 1], dirs[r * 3 + 2]};
      v = pe3(x, j);
    } else if (j < 27 + n_a) {
      v = a[(long long)r * n_a + (j - 27)];
    }
    planes_store(IN1, n_planes, (long long)m * IN1.ld + 512 + j, v);
  } else {
    const int q = j - 128;
    float v = 0.0f;
    if (q < 3) v = pts[m * 3 + q];
    else if (q < 6) v = nrm[m * 3 + (q - 3)];
    planes_store(IN2, n_planes, (long long)m * IN2.ld + 128 + q, v);
  }
}
int launch_color_embed(const float* dirs, const float* a, int n_a, int rows_per_src, const float* pts,
                       const float* nrm, int M, int n_planes, Planes IN1, Planes IN2, cudaStream_t s) {
  color_embed_kernel<<<M, 192, 0, s>>>(dirs, a, n_a, rows_per_src, pts, nrm, M, n_planes, IN1, IN2);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// Background NeRF inputs (renderer.py:157-203; models/nerf.py:156-160).  256 threads per sample:
// j < 128: IN0[m, j] and IN5[m, 256 + j] = [PE10(pts4) 84 | 0];  j >= 128: FEATN[m, 256 + (j-128)] = [viewPE 27 | a | 0]
__global__ void __launch_bounds__(256) nerf_embed_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                         const float* __restrict__ z,
                                                         const float* __restrict__ sample_dist,
                                                         const float* __restrict__ pts4_in, const float* __restrict__ a,
                                                         int n_a, int T, int rows_per_src, int M, int n_planes, Planes IN0,
                                                         Planes IN5, Planes FEATN, float* __restrict__ dists_out) {
  const int m = blockIdx.x, j = threadIdx.x;
  if (m >= M) return;
  const int r = m / rows_per_src;
  if (j < 128) {
    float v = 0.0f;
    if (j < 84) {
      float p4[4];
      if (pts4_in) {
#pragma unroll
        for (int c = 0; c < 4; ++c) p4[c] = pts4_in[(long long)m * 4 + c];
      } else {
        const int i = m % T;
        const float t0 = z[m];
        const float dist = (i + 1 < T) ? __fsub_rn(z[m + 1], t0) : sample_dist[r];
        const float mid = __fadd_rn(t0, __fmul_rn(dist, 0.5f));
        float p[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) p[c] = __fadd_rn(o[r * 3 + c], __fmul_rn(d[r * 3 + c], mid));
        float nr = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        nr = fminf(fmaxf(nr, 1.0f), 1e10f);
        p4[0] = p[0] / nr; p4[1] = p[1] / nr; p4[2] = p[2] / nr; p4[3] = 1.0f / nr;
        if (j == 0 && dists_out) dists_out[m] = dist;
      }
      v = pe4(p4, j);
    }
    planes_store(IN0, n_planes, (long long)m * IN0.ld + j, v);
    planes_store(IN5, n_planes, (long long)m * IN5.ld + 256 + j, v);
  } else {
    const int q = j - 128;
    float v = 0.0f;
    if (q < 27) {
      const float x[3] = {d[r * 3], d[r * 3 + 1], d[r * 3 + 2]};
      v = pe3(x, q);
    } else if (q < 27 + n_a) {
      v = a[(long long)r * n_a + (q - 27)];
    }
    planes_store(FEATN, n_planes, (long long)m * FEATN.ld + 256 + q, v);
  }
}
int launch_nerf_embed(const float* o, const float* d, const float* z, const float* sample_dist,
                      const float* pts4_in, const float* a, int n_a, int T, int rows_per_src, int M,
                      int n_planes, Planes IN0, Planes IN5, Planes FEATN, float* dists_out, cudaStream_t s) {
  nerf_embed_kernel<<<M, 256, 0, s>>>(o, d, z, sample_dist, pts4_in, a, n_a, T, rows_per_src, M, n_planes, IN0, IN5,
                                      FEATN, dists_out);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// ---- column sums: out[n] += sum_m rowscale[m] * X[m,n]  (bias gradients, sdf-head weight gradient) ----------
// X is either bf16 planes or fp32.  A block owns ROWS rows; a thread owns 8 consecutive columns (16-byte loads)
// of every (256 / (N/8))-th row; partial sums are combined in shared memory, then one atomicAdd per column.
static constexpr int CS_ROWS = 128;
__global__ void __launch_bounds__(256) colsum_kernel(Planes X, int n_planes, const float* __restrict__ Xf, int ld, int M,
                                                     int N, const float* __restrict__ rowscale, float* __restrict__ out,
                                                     float* __restrict__ out_rs) {
  __shared__ float red[256 * 8];
  const int tpr = N >> 3;                 // threads per row (N % 8 == 0, N <= 640 -> tpr <= 80)
  const int rpp = 256 / tpr;              // rows per pass
  const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
  const int row0 = blockIdx.x * CS_ROWS, row1 = min(M, row0 + CS_ROWS);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float rs_acc = 0.0f;
  if (tr < rpp) {
    for (int m = row0 + tr; m < row1; m += rpp) {
      const float rsc = rowscale ? rowscale[m] : 1.0f;
      if (tc == 0) rs_acc += rsc;
      float v[8];
      if (Xf) {
        const float4 a0 = *reinterpret_cast<const float4*>(Xf + (long long)m * ld + tc * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(Xf + (long long)m * ld + tc * 8 + 4);
        v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.0f;
        for (int pl = 0; pl < n_planes; ++pl) {
          const uint4 t = *reinterpret_cast<const uint4*>(X.plane(pl) + (long long)m * X.ld + tc * 8);
          const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            v[2 * k] += __uint_as_float(u[k] << 16);
            v[2 * k + 1] += __uint_as_float(u[k] & 0xFFFF0000u);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(rsc, v[k], acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  // thread t < N sums column t over the rpp row groups
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int c = n >> 3, k = n & 7;
    float s = 0.0f;
    for (int g = 0; g < rpp; ++g) s += red[(g * tpr + c) * 8 + k];
    atomicAdd(&out[n], s);
  }
  if (out_rs && tc == 0 && tr < rpp) atomicAdd(out_rs, rs_acc);
}
int launch_colsum(Planes X, int n_planes, const float* Xf, int ld, int M, int N, const float* rowscale,
                  float* out, float* out_rowscale_sum, cudaStream_t s) {
  NRW_CHECK(N % 8 == 0 && N <= 640, NRW_ERR_ARG, "colsum: N=%d must be a multiple of 8 and <= 640", N);
  colsum_kernel<<<cdiv(M, CS_ROWS), 256, 0, s>>>(X, n_planes, Xf, ld, M, N, rowscale, out, out_rowscale_sum);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

}  // namespace nrw
