// Memory-bound helper kernels around the GEMM chain: positional encodings (and their Jacobians),
// the narrow output heads (N <= 3) and their backward, column sums for bias gradients.
#include "pointwise.h"

namespace nrw {

static constexpr float INV_SQRT2 = 0.70710678118654752440f;

// pts[m] = o[r] + d[r] * t,  t = z (use_mid = 0) or z + 0.5 * dist (use_mid = 1; renderer.py:586-593)
// explicit _rn ops: torch evaluates mul and add separately (no FMA contraction).
__global__ void points_kernel(const float* __restrict__ o, const float* __restrict__ d,
                              const float* __restrict__ z, const float* __restrict__ sample_dist, int R,
                              int S, int use_mid, float* __restrict__ pts) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= (long long)R * S) return;
  const int r = (int)(m / S), i = (int)(m % S);
  float t = z[m];
  if (use_mid) {
    const float dist = (i + 1 < S) ? __fsub_rn(z[m + 1], t) : sample_dist[r];
    t = __fadd_rn(t, __fmul_rn(dist, 0.5f));
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) pts[m * 3 + c] = __fadd_rn(o[r * 3 + c], __fmul_rn(d[r * 3 + c], t));
}
int launch_points(const float* o, const float* d, const float* z, const float* sample_dist, int R, int S,
                  int use_mid, float* pts, cudaStream_t s) {
  const long long M = (long long)R * S;
  if (M == 0) return NRW_OK;
  points_kernel<<<cdiv(M, 256), 256, 0, s>>>(o, d, z, sample_dist, R, S, use_mid, pts);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// ---- SDF positional encoding (L=6, 39 wide; models/neuconw.py:7-55) -----------------------------
// U0[m, 0:64] = [x, sin(2^k x), cos(2^k x) ..., 0 pad];  U4[m, 473:512] = PE / sqrt(2)
__global__ void sdf_embed_kernel(const float* __restrict__ pts, int M, int n_planes, Planes U0, Planes U4) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float pe[39];
  const float x0 = pts[m * 3], x1 = pts[m * 3 + 1], x2 = pts[m * 3 + 2];
  pe[0] = x0; pe[1] = x1; pe[2] = x2;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float f = (float)(1 << k);
    float s, c;
    sincosf(x0 * f, &s, &c); pe[3 + 6 * k] = s; pe[6 + 6 * k] = c;
    sincosf(x1 * f, &s, &c); pe[4 + 6 * k] = s; pe[7 + 6 * k] = c;
    sincosf(x2 * f, &s, &c); pe[5 + 6 * k] = s; pe[8 + 6 * k] = c;
  }
  for (int j = 0; j < 64; ++j) planes_store(U0, n_planes, (long long)m * U0.ld + j, j < 39 ? pe[j] : 0.0f);
  if (U4.p)
    for (int j = 0; j < 39; ++j) planes_store(U4, n_planes, (long long)m * U4.ld + 473 + j, pe[j] * INV_SQRT2);
}
int launch_sdf_embed(const float* pts, int M, int n_planes, Planes U0, Planes U4, cudaStream_t s) {
  sdf_embed_kernel<<<cdiv(M, 128), 128, 0, s>>>(pts, M, n_planes, U0, U4);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// ---- SDF head: sdf = softplus(a7) . w0 + b0 ; optional gbar7 = softplus'(a7) * w0 ------------------
__global__ void __launch_bounds__(256) sdf_head_kernel(const float* __restrict__ A7, int M,
                                                       const float* __restrict__ w0,
                                                       const float* __restrict__ b0, float* __restrict__ sdf,
                                                       int n_planes, Planes G7) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* a = A7 + (long long)warp * 512;
  float acc = 0.0f;
  for (int j = lane; j < 512; j += 32) {
    const float v = a[j], w = w0[j];
    acc = fmaf(softplus100(v), w, acc);
    if (G7.p) planes_store(G7, n_planes, (long long)warp * G7.ld + j, softplus100_d1(v) * w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) sdf[warp] = acc + b0[0];
}
int launch_sdf_head(const float* A7, int M, const float* w0, const float* b0, float* sdf, int n_planes,
                    Planes G7, cudaStream_t s) {
  sdf_head_kernel<<<cdiv((long long)M * 32, 256), 256, 0, s>>>(A7, M, w0, b0, sdf, n_planes, G7);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// ---- normal = J_PE(x)^T (q0[:39] + q4[473:512] / sqrt2)   -----------------------------
__global__ void sdf_normal_kernel(const float* __restrict__ pts, const float* __restrict__ Q0,
                                  const float* __restrict__ Q4, int M, float* __restrict__ nrm) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float v[39];
  for (int j = 0; j < 39; ++j) v[j] = Q0[(long long)m * 64 + j] + Q4[(long long)m * 512 + 473 + j] * INV_SQRT2;
  float out[3] = {v[0], v[1], v[2]};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = pts[m * 3 + c];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float f = (float)(1 << k);
      float sn, cs;
      sincosf(x * f, &sn, &cs);
      out[c] += v[3 + 6 * k + c] * f * cs - v[6 + 6 * k + c] * f * sn;
    }
  }
  nrm[m * 3] = out[0]; nrm[m * 3 + 1] = out[1]; nrm[m * 3 + 2] = out[2];
}
int launch_sdf_normal(const float* pts, const float* Q0, const float* Q4, int M, float* nrm, cudaStream_t s) {
  sdf_normal_kernel<<<cdiv(M, 128), 128, 0, s>>>(pts, Q0, Q4, M, nrm);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// backward of the above w.r.t. (q0, q4 tail): t = J_PE(x) dn
__global__ void sdf_normal_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ dn, int M,
                                      int n_planes, Planes DQ0, Planes DQ4) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float t[39];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = pts[m * 3 + c], g = dn[m * 3 + c];
    t[c] = g;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float f = (float)(1 << k);
      float sn, cs;
      sincosf(x * f, &sn, &cs);
      t[3 + 6 * k + c] = g * f * cs;
      t[6 + 6 * k + c] = -g * f * sn;
    }
  }
  for (int j = 0; j < 64; ++j) planes_store(DQ0, n_planes, (long long)m * DQ0.ld + j, j < 39 ? t[j] : 0.0f);
  for (int j = 0; j < 39; ++j) planes_store(DQ4, n_planes, (long long)m * DQ4.ld + 473 + j, t[j] * INV_SQRT2);
}
int launch_sdf_normal_bwd(const float* pts, const float* dn, int M, int n_planes, Planes DQ0, Planes DQ4,
                          cudaStream_t s) {
  sdf_normal_bwd_kernel<<<cdiv(M, 128), 128, 0, s>>>(pts, dn, M, n_planes, DQ0, DQ4);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// ---- view-direction PE (L=4, 27 wide) + appearance code + geometry inputs for the colour net --------
// IN1[m, 512:640] = [viewPE(d) 27 | a n_a | 0];  IN2[m, 128:192] = [pts 3 | normal 3 | 0]
__global__ void color_embed_kernel(const float* __restrict__ dirs, const float* __restrict__ a, int n_a,
                                   int rows_per_src, const float* __restrict__ pts,
                                   const float* __restrict__ nrm, int M, int n_planes, Planes IN1, Planes IN2) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int r = m / rows_per_src;
  float pe[27];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = dirs[r * 3 + c];
    pe[c] = x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float sn, cs;
      sincosf(x * (float)(1 << k), &sn, &cs);
      pe[3 + 6 * k + c] = sn;
      pe[6 + 6 * k + c] = cs;
    }
  }
  const long long b1 = (long long)m * IN1.ld + 512;
  for (int j = 0; j < 128; ++j) {
    float v = 0.0f;
    if (j < 27) v = pe[j];
    else if (j < 27 + n_a) v = a[(long long)r * n_a + (j - 27)];
    planes_store(IN1, n_planes, b1 + j, v);
  }
  const long long b2 = (long long)m * IN2.ld + 128;
  for (int j = 0; j < 64; ++j) {
    float v = 0.0f;
    if (j < 3) v = pts[m * 3 + j];
    else if (j < 6) v = nrm[m * 3 + (j - 3)];
    planes_store(IN2, n_planes, b2 + j, v);
  }
}
int launch_color_embed(const float* dirs, const float* a, int n_a, int rows_per_src, const float* pts,
                       const float* nrm, int M, int n_planes, Planes IN1, Planes IN2, cudaStream_t s) {
  color_embed_kernel<<<cdiv(M, 128), 128, 0, s>>>(dirs, a, n_a, rows_per_src, pts, nrm, M, n_planes, IN1, IN2);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// ---- background NeRF inputs (renderer.py:157-203; models/nerf.py:156-160) -----------------------------
// mode 0: from rays: z_feed [R,T] -> dists, mid, p = o + d*mid, r = clip(|p|,1,1e10), pts4 = [p/r, 1/r]
// mode 1: explicit pts4 [M,4] (dirs/a per point when rows_per_src == 1)
// IN0[m,0:128] = [PE10(pts4) 84 | 0]; IN5[m,256:384] = same; FEATN[m,256:384] = [viewPE 27 | a | 0]
__global__ void nerf_embed_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                  const float* __restrict__ z, const float* __restrict__ sample_dist,
                                  const float* __restrict__ pts4_in, const float* __restrict__ a, int n_a,
                                  int T, int rows_per_src, int M, int n_planes, Planes IN0, Planes IN5,
                                  Planes FEATN, float* __restrict__ dists_out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int r = m / rows_per_src;
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
    if (dists_out) dists_out[m] = dist;
  }
  const long long b0 = (long long)m * IN0.ld, b5 = (long long)m * IN5.ld + 256;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    planes_store(IN0, n_planes, b0 + c, p4[c]);
    planes_store(IN5, n_planes, b5 + c, p4[c]);
  }
  for (int k = 0; k < 10; ++k) {
    const float f = (float)(1 << k);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float sn, cs;
      sincosf(p4[c] * f, &sn, &cs);
      planes_store(IN0, n_planes, b0 + 4 + 8 * k + c, sn);
      planes_store(IN0, n_planes, b0 + 8 + 8 * k + c, cs);
      planes_store(IN5, n_planes, b5 + 4 + 8 * k + c, sn);
      planes_store(IN5, n_planes, b5 + 8 + 8 * k + c, cs);
    }
  }
  for (int j = 84; j < 128; ++j) {
    planes_store(IN0, n_planes, b0 + j, 0.0f);
    planes_store(IN5, n_planes, b5 + j, 0.0f);
  }
  float pe[27];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = d[r * 3 + c];
    pe[c] = x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float sn, cs;
      sincosf(x * (float)(1 << k), &sn, &cs);
      pe[3 + 6 * k + c] = sn;
      pe[6 + 6 * k + c] = cs;
    }
  }
  const long long bf = (long long)m * FEATN.ld + 256;
  for (int j = 0; j < 128; ++j) {
    float v = 0.0f;
    if (j < 27) v = pe[j];
    else if (j < 27 + n_a) v = a[(long long)r * n_a + (j - 27)];
    planes_store(FEATN, n_planes, bf + j, v);
  }
}
int launch_nerf_embed(const float* o, const float* d, const float* z, const float* sample_dist,
                      const float* pts4_in, const float* a, int n_a, int T, int rows_per_src, int M,
                      int n_planes, Planes IN0, Planes IN5, Planes FEATN, float* dists_out, cudaStream_t s) {
  nerf_embed_kernel<<<cdiv(M, 128), 128, 0, s>>>(o, d, z, sample_dist, pts4_in, a, n_a, T, rows_per_src, M,
                                                 n_planes, IN0, IN5, FEATN, dists_out);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// ---- narrow heads: out[m,c] = act(X[m,:] . W[c,:] + b[c]),  NOUT in {1,3} --------------------------
// act: 0 none, 3 sigmoid.  For the NeRF alpha head (NOUT=1) `dists` turns density into
// alpha = 1 - exp(-softplus(density) * dist) (renderer.py:205-207); density is kept in out2.
template <int NOUT>
__global__ void __launch_bounds__(256) head_kernel(Planes X, int n_planes, int K, int M,
                                                   const float* __restrict__ W, const float* __restrict__ b,
                                                   int act, const float* __restrict__ dists,
                                                   float* __restrict__ out, float* __restrict__ out2) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  float acc[NOUT];
#pragma unroll
  for (int c = 0; c < NOUT; ++c) acc[c] = 0.0f;
  for (int j = lane; j < K; j += 32) {
    const float x = planes_load(X, n_planes, (long long)warp * X.ld + j);
#pragma unroll
    for (int c = 0; c < NOUT; ++c) acc[c] = fmaf(x, W[c * K + j], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < NOUT; ++c)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < NOUT; ++c) {
      float v = acc[c] + b[c];
      if (act == ACT_SIGMOID) v = sigmoidf_(v);
      if (NOUT == 1 && dists) {
        if (out2) out2[warp] = v;
        const float sp = v > 20.0f ? v : log1pf(expf(v));
        v = 1.0f - expf(-sp * dists[warp]);
      }
      out[(long long)warp * NOUT + c] = v;
    }
  }
}
int launch_head(int nout, Planes X, int n_planes, int K, int M, const float* W, const float* b, int act,
                const float* dists, float* out, float* out2, cudaStream_t s) {
  const int grid = cdiv((long long)M * 32, 256);
  if (nout == 1) head_kernel<1><<<grid, 256, 0, s>>>(X, n_planes, K, M, W, b, act, dists, out, out2);
  else head_kernel<3><<<grid, 256, 0, s>>>(X, n_planes, K, M, W, b, act, dists, out, out2);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// backward of a narrow head.  dpre[m,c] is formed on the fly:
//   mode 0: dpre = g_out                      (linear head)
//   mode 1: dpre = g_out * y * (1-y)          (sigmoid head; y = forward output)
//   mode 2: NeRF alpha head: dpre = g_alpha * exp(-softplus(density)*dist) * dist * sigmoid(density)
// writes dX[m,k] = relu'(X[m,k]) * sum_c dpre[m,c] W[c,k] as planes (or, when dX.p == null, only dpre to
// dpre_out), and accumulates dW[c,k] += sum_m dpre X, db[c] += sum_m dpre with block-level partials.
template <int NOUT>
__global__ void __launch_bounds__(256) head_bwd_kernel(Planes X, int n_planes, int K, int M,
                                                       const float* __restrict__ W,
                                                       const float* __restrict__ g_out,
                                                       const float* __restrict__ y_or_density,
                                                       const float* __restrict__ dists, int mode, Planes dX,
                                                       float* __restrict__ dpre_out, float* __restrict__ dW,
                                                       float* __restrict__ db) {
  extern __shared__ float sm[];  // [NOUT*K] partial dW, [NOUT] partial db
  float* sW = sm;
  float* sb = sm + NOUT * K;
  for (int i = threadIdx.x; i < NOUT * K + NOUT; i += blockDim.x) sm[i] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int rows_per_block = 64;
  const int row0 = blockIdx.x * rows_per_block;
  for (int rr = wib; rr < rows_per_block; rr += wpb) {
    const int m = row0 + rr;
    if (m >= M) break;
    float dp[NOUT];
#pragma unroll
    for (int c = 0; c < NOUT; ++c) {
      float g = g_out[(long long)m * NOUT + c];
      if (mode == 1) {
        const float y = y_or_density[(long long)m * NOUT + c];
        g = g * y * (1.0f - y);
      } else if (mode == 2) {
        const float dens = y_or_density[m];
        const float sp = dens > 20.0f ? dens : log1pf(expf(dens));
        const float dsp = dens > 20.0f ? 1.0f : sigmoidf_(dens);
        g = g * expf(-sp * dists[m]) * dists[m] * dsp;
      }
      dp[c] = g;
    }
    if (dpre_out && lane < NOUT) dpre_out[(long long)m * NOUT + lane] = dp[lane < NOUT ? lane : 0];
    for (int j = lane; j < K; j += 32) {
      const float x = planes_load(X, n_planes, (long long)m * X.ld + j);
      float acc = 0.0f;
#pragma unroll
      for (int c = 0; c < NOUT; ++c) {
        acc = fmaf(dp[c], W[c * K + j], acc);
        atomicAdd(&sW[c * K + j], dp[c] * x);
      }
      if (dX.p) planes_store(dX, n_planes, (long long)m * dX.ld + j, x > 0.0f ? acc : 0.0f);
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < NOUT; ++c) atomicAdd(&sb[c], dp[c]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NOUT * K; i += blockDim.x) atomicAdd(&dW[i], sW[i]);
  if (threadIdx.x < NOUT) atomicAdd(&db[threadIdx.x], sb[threadIdx.x]);
}
int launch_head_bwd(int nout, Planes X, int n_planes, int K, int M, const float* W, const float* g_out,
                    const float* y_or_density, const float* dists, int mode, Planes dX, float* dpre_out,
                    float* dW, float* db, cudaStream_t s) {
  const int grid = cdiv(M, 64);
  const size_t smem = (size_t)(nout * K + nout) * sizeof(float);
  if (nout == 1)
    head_bwd_kernel<1><<<grid, 256, smem, s>>>(X, n_planes, K, M, W, g_out, y_or_density, dists, mode, dX, dpre_out, dW, db);
  else
    head_bwd_kernel<3><<<grid, 256, smem, s>>>(X, n_planes, K, M, W, g_out, y_or_density, dists, mode, dX, dpre_out, dW, db);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// ---- column sums: out[n] += sum_m X[m,n]  (bias gradients) ----------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(Planes X, int n_planes, const float* __restrict__ Xf, int ld,
                                                     int M, int N, const float* __restrict__ rowscale,
                                                     float* __restrict__ out, float* __restrict__ out_rs) {
  // block handles 256 rows x all columns; thread t owns columns t, t+256, ...
  const int row0 = blockIdx.x * 256, row1 = min(M, row0 + 256);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float acc = 0.0f;
    for (int m = row0; m < row1; ++m) {
      float v = Xf ? Xf[(long long)m * ld + n] : planes_load(X, n_planes, (long long)m * X.ld + n);
      if (rowscale) v *= rowscale[m];
      acc += v;
    }
    atomicAdd(&out[n], acc);
  }
  if (out_rs && threadIdx.x == 0) {
    float acc = 0.0f;
    for (int m = row0; m < row1; ++m) acc += rowscale[m];
    atomicAdd(out_rs, acc);
  }
}
int launch_colsum(Planes X, int n_planes, const float* Xf, int ld, int M, int N, const float* rowscale,
                  float* out, float* out_rowscale_sum, cudaStream_t s) {
  colsum_kernel<<<cdiv(M, 256), 256, 0, s>>>(X, n_planes, Xf, ld, M, N, rowscale, out, out_rowscale_sum);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// generic small utilities
__global__ void fill_kernel(float* p, long long n, float v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
int launch_fill(float* p, long long n, float v, cudaStream_t s) {
  if (n == 0) return NRW_OK;
  fill_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, n, v);
  NRW_LAUNCH_OK();
  return NRW_OK;
}
__global__ void add3_kernel(float* dst, const float* a, const float* b, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = a[i] + (b ? b[i] : 0.0f);
}
int launch_add(float* dst, const float* a, const float* b, long long n, cudaStream_t s) {
  if (n == 0) return NRW_OK;
  add3_kernel<<<cdiv(n, 256), 256, 0, s>>>(dst, a, b, n);
  NRW_LAUNCH_OK();
  return NRW_OK;
}
// split an fp32 matrix [rows, cols] (ld_src) into planes (ld = P.ld), zero padding cols..ld
__global__ void split_planes_kernel(const float* __restrict__ src, long long rows, int cols, int ld_src,
                                    int n_planes, Planes P) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * P.ld) return;
  const long long r = i / P.ld;
  const int c = (int)(i % P.ld);
  planes_store(P, n_planes, i, c < cols ? src[r * ld_src + c] : 0.0f);
}
int launch_split_planes(const float* src, long long rows, int cols, int ld_src, int n_planes, Planes P,
                        cudaStream_t s) {
  split_planes_kernel<<<cdiv(rows * P.ld, 256), 256, 0, s>>>(src, rows, cols, ld_src, n_planes, P);
  NRW_LAUNCH_OK();
  return NRW_OK;
}
// per-ray segment sum: out[r, c] = sum_{i<S} X[(r*S+i), c]   (appearance-code gradient)
__global__ void segsum_kernel(const float* __restrict__ X, int ld, int col0, int ncols, int R, int S,
                              float* __restrict__ out, int accumulate) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    float acc = 0.0f;
    for (int i = 0; i < S; ++i) acc += X[((long long)r * S + i) * ld + col0 + c];
    if (accumulate) out[(long long)r * ncols + c] += acc; else out[(long long)r * ncols + c] = acc;
  }
}
int launch_segsum(const float* X, int ld, int col0, int ncols, int R, int S, float* out, int accumulate,
                  cudaStream_t s) {
  if (R == 0) return NRW_OK;
  segsum_kernel<<<R, 64, 0, s>>>(X, ld, col0, ncols, R, S, out, accumulate);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

}  // namespace nrw
