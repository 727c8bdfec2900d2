// Voxel-guided hierarchical sampler (NeuconWRenderer.sparse_sampler, rendering/renderer.py:458-568):
// coarse / outside strata, SDF-driven importance resampling (up_sample + sample_pdf, renderer.py:15-48,
// 257-341) and the sorted merge that replaces torch.sort in cat_z_vals (renderer.py:343-363).
//
// One thread owns one ray and walks its samples front to back with the operation order written
// down in include/nrw_math.h, so the integer artefacts (searchsorted indices, merge permutation) are
// bit-reproducible against the C restatement in oracle/sampler_ref.c.  The work is O(R * S) scalar
// operations on a few MB - it is latency-, not bandwidth-bound, and < 1 % of a training step.
#include "../../include/nrw_math.h"
#include "pointwise.h"

namespace nrw {

__global__ void coarse_z_kernel(nrw_sampler_cfg c, int R, const float* __restrict__ near,
                                const float* __restrict__ far, const float* __restrict__ s_near,
                                const float* __restrict__ s_far, const float* __restrict__ u_ray,
                                const float* __restrict__ u_out, float* __restrict__ z,
                                float* __restrict__ z_out, float* __restrict__ sample_dist) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float sn = s_near ? s_near[r] : near[r];
  const float sf = s_far ? s_far[r] : far[r];
  const float range = NRW_SUB(sf, sn);
  const float ns = (float)c.n_samples;
  sample_dist[r] = NRW_DIV(range, ns);
  float shift = 0.0f;
  if (c.perturb) shift = NRW_DIV(NRW_MUL(NRW_MUL(range, NRW_SUB(u_ray[r], 0.5f)), 2.0f), ns);
  for (int j = 0; j < c.n_samples; ++j) {
    float v = NRW_ADD(sn, NRW_MUL(range, nrw_linspace_f32(0.0f, 1.0f, c.n_samples, j)));
    if (c.perturb) v = NRW_ADD(v, shift);
    z[(long long)r * c.n_samples + j] = v;
  }
  const int no = c.n_outside;
  if (no > 0) {
    const float hi = (float)(1.0 - 1.0 / ((double)no + 1.0));
    const float add = (float)(1.0 / (double)c.n_samples);
    for (int j = 0; j < no; ++j) {
      // element j of the result uses the flipped stratum jj = no-1-j
      const int jj = no - 1 - j;
      float b = nrw_linspace_f32(1e-3f, hi, no, jj);
      if (c.perturb) {
        const float bl = jj > 0 ? nrw_linspace_f32(1e-3f, hi, no, jj - 1) : b;
        const float bu = jj + 1 < no ? nrw_linspace_f32(1e-3f, hi, no, jj + 1) : b;
        const float lower = jj > 0 ? NRW_MUL(0.5f, NRW_ADD(b, bl)) : b;
        const float upper = jj + 1 < no ? NRW_MUL(0.5f, NRW_ADD(bu, b)) : b;
        b = NRW_ADD(lower, NRW_MUL(NRW_SUB(upper, lower), u_out[(long long)r * no + jj]));
      }
      z_out[(long long)r * no + j] = NRW_ADD(NRW_DIV(far[r], b), add);
    }
  }
}
int launch_coarse_z(const nrw_sampler_cfg& c, int R, const float* near, const float* far, const float* s_near,
                    const float* s_far, const float* u_ray, const float* u_out, float* z, float* z_out,
                    float* sample_dist, cudaStream_t s) {
  coarse_z_kernel<<<cdiv(R, 128), 128, 0, s>>>(c, R, near, far, s_near, s_far, u_ray, u_out, z, z_out, sample_dist);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// One importance-resampling round for one ray.  cdf is a caller-provided scratch row of m floats.
__device__ void upsample_ray(int m, int n_new, float inv_s, const float* o, const float* d, const float* z,
                             const float* sdf, float* cdf, float* z_new, float* z_merged, int32_t* inds,
                             int32_t* order) {
  // pass 1: per-interval weights, unnormalised, stored in cdf[1..m-1]; running sum
  float prev_cos_raw = 0.0f, T = 1.0f, wsum = 0.0f;
  float px = NRW_ADD(o[0], NRW_MUL(d[0], z[0])), py = NRW_ADD(o[1], NRW_MUL(d[1], z[0])),
        pz = NRW_ADD(o[2], NRW_MUL(d[2], z[0]));
  float rad_prev = NRW_SQRT(NRW_ADD(NRW_ADD(NRW_MUL(px, px), NRW_MUL(py, py)), NRW_MUL(pz, pz)));
  for (int j = 0; j + 1 < m; ++j) {
    const float z0 = z[j], z1 = z[j + 1], s0 = sdf[j], s1 = sdf[j + 1];
    px = NRW_ADD(o[0], NRW_MUL(d[0], z1)); py = NRW_ADD(o[1], NRW_MUL(d[1], z1)); pz = NRW_ADD(o[2], NRW_MUL(d[2], z1));
    const float rad = NRW_SQRT(NRW_ADD(NRW_ADD(NRW_MUL(px, px), NRW_MUL(py, py)), NRW_MUL(pz, pz)));
    const float inside = (rad_prev < 1.0f || rad < 1.0f) ? 1.0f : 0.0f;
    rad_prev = rad;
    const float dz = NRW_SUB(z1, z0);
    const float cos_raw = NRW_DIV(NRW_SUB(s1, s0), NRW_ADD(dz, 1e-5f));
    float cv = fminf(prev_cos_raw, cos_raw);
    prev_cos_raw = cos_raw;
    cv = NRW_MUL(fminf(fmaxf(cv, -1e3f), 0.0f), inside);
    const float mid = NRW_MUL(NRW_ADD(s0, s1), 0.5f);
    const float h = NRW_MUL(NRW_MUL(cv, dz), 0.5f);
    const float pc = nrw_sigmoid_f32(NRW_MUL(NRW_SUB(mid, h), inv_s));
    const float nc = nrw_sigmoid_f32(NRW_MUL(NRW_ADD(mid, h), inv_s));
    const float alpha = NRW_DIV(NRW_ADD(NRW_SUB(pc, nc), 1e-5f), NRW_ADD(pc, 1e-5f));
    const float w = NRW_ADD(NRW_MUL(alpha, T), 1e-5f);  // weights + 1e-5 (renderer.py:19)
    T = NRW_MUL(T, NRW_ADD(NRW_SUB(1.0f, alpha), 1e-7f));
    cdf[j + 1] = w;
    wsum = NRW_ADD(wsum, w);
  }
  // pass 2: normalise and accumulate
  cdf[0] = 0.0f;
  float run = 0.0f;
  for (int j = 1; j < m; ++j) {
    run = NRW_ADD(run, NRW_DIV(cdf[j], wsum));
    cdf[j] = run;
  }
  // pass 3: invert at the n_new stratified midpoints (searchsorted right=True), two-pointer walk
  const float u0 = (float)(0.0 + 0.5 / (double)n_new), u1 = (float)(1.0 - 0.5 / (double)n_new);
  int ind = 0;
  for (int t = 0; t < n_new; ++t) {
    const float u = nrw_linspace_f32(u0, u1, n_new, t);
    while (ind < m && cdf[ind] <= u) ++ind;
    const int below = ind - 1 > 0 ? ind - 1 : 0;
    const int above = ind < m - 1 ? ind : m - 1;
    float den = NRW_SUB(cdf[above], cdf[below]);
    if (den < 1e-5f) den = 1.0f;
    const float tt = NRW_DIV(NRW_SUB(u, cdf[below]), den);
    z_new[t] = NRW_ADD(z[below], NRW_MUL(tt, NRW_SUB(z[above], z[below])));
    if (inds) inds[t] = ind;
  }
  // pass 4: stable merge of the two ascending runs (ties: existing sample first)
  int a = 0, b = 0;
  for (int k = 0; k < m + n_new; ++k) {
    const bool take_a = (b >= n_new) || (a < m && z[a] <= z_new[b]);
    if (take_a) { z_merged[k] = z[a]; if (order) order[k] = a; ++a; }
    else { z_merged[k] = z_new[b]; if (order) order[k] = m + b; ++b; }
  }
}

__global__ void upsample_round_kernel(int R, int m, int n_new, float inv_s, const float* __restrict__ o,
                                      const float* __restrict__ d, const float* __restrict__ z,
                                      const float* __restrict__ sdf, float* __restrict__ cdf,
                                      float* __restrict__ z_new, float* __restrict__ z_merged,
                                      int32_t* __restrict__ inds, int32_t* __restrict__ order) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  upsample_ray(m, n_new, inv_s, o + r * 3, d + r * 3, z + (long long)r * m, sdf + (long long)r * m,
               cdf + (long long)r * m, z_new + (long long)r * n_new, z_merged + (long long)r * (m + n_new),
               inds ? inds + (long long)r * n_new : nullptr, order ? order + (long long)r * (m + n_new) : nullptr);
}
int launch_upsample_round(int R, int m, int n_new, float inv_s, const float* o, const float* d, const float* z,
                          const float* sdf, float* cdf_scratch, float* z_new, float* z_merged, int32_t* inds,
                          int32_t* order, cudaStream_t s) {
  upsample_round_kernel<<<cdiv(R, 64), 64, 0, s>>>(R, m, n_new, inv_s, o, d, z, sdf, cdf_scratch, z_new,
                                                   z_merged, inds, order);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

__global__ void merge_sdf_kernel(int R, int m, int n_new, const float* __restrict__ sdf_old,
                                 const float* __restrict__ sdf_new, const int32_t* __restrict__ order,
                                 float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int tot = m + n_new;
  if (i >= (long long)R * tot) return;
  const int r = (int)(i / tot);
  const int k = order[i];
  out[i] = k < m ? sdf_old[(long long)r * m + k] : sdf_new[(long long)r * n_new + (k - m)];
}
int launch_merge_sdf(int R, int m, int n_new, const float* sdf_old, const float* sdf_new, const int32_t* order,
                     float* sdf_merged, cudaStream_t s) {
  merge_sdf_kernel<<<cdiv((long long)R * (m + n_new), 256), 256, 0, s>>>(R, m, n_new, sdf_old, sdf_new, order, sdf_merged);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// boundary samples (renderer.py:549-566): n_near on [near, z_0), n_far on (z_last, far], then merge
__global__ void boundary_kernel(int R, int S0, int nb, const float* __restrict__ near,
                                const float* __restrict__ far, const float* __restrict__ z,
                                float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int n_near = nb / 2, n_far = nb - n_near;
  const float* zr = z + (long long)r * S0;
  float* o = out + (long long)r * (S0 + nb);
  // The reference sorts the concatenation (torch.sort, renderer.py:565).  Each run is monotone, so a 3-way merge gives
  // the same values: the near run ascends when near <= z0 and DESCENDS when the fine sampling window starts before the
  // (octree-overridden) near, the far run likewise; descending runs are walked from their last element.
  int a = 0, b = 0, c = 0;
  const float z0 = zr[0], zl = zr[S0 - 1], nr = near[r], fr = far[r];
  const bool a_desc = z0 < nr, b_desc = fr < zl;
  for (int k = 0; k < S0 + nb; ++k) {
    const int ia = a_desc ? n_near - 1 - a : a, ib = b_desc ? n_far - 1 - b : b;
    const float va = a < n_near ? NRW_ADD(nr, NRW_MUL(NRW_SUB(z0, nr), nrw_linspace_f32(0.0f, 1.0f, n_near + 1, ia))) : INFINITY;
    const float vb = b < n_far ? NRW_ADD(zl, NRW_MUL(NRW_SUB(fr, zl), nrw_linspace_f32(0.0f, 1.0f, n_far + 1, ib + 1))) : INFINITY;
    const float vc = c < S0 ? zr[c] : INFINITY;
    // smallest first; ties resolved in concatenation order [near run, far run, z]
    if (a < n_near && va <= vb && va <= vc) { o[k] = va; ++a; }
    else if (b < n_far && vb <= vc) { o[k] = vb; ++b; }
    else { o[k] = vc; ++c; }
  }
}
int launch_boundary(int R, int S0, int nb, const float* near, const float* far, const float* z, float* out,
                    cudaStream_t s) {
  boundary_kernel<<<cdiv(R, 128), 128, 0, s>>>(R, S0, nb, near, far, z, out);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

// out[r] = stable merge of the ascending rows a[r, :na], b[r, :nb]  (z_vals_feed, renderer.py:835-836)
__global__ void merge_sorted_kernel(int R, int na, int nb, const float* __restrict__ a,
                                    const float* __restrict__ b, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* ar = a + (long long)r * na;
  const float* br = b + (long long)r * nb;
  float* o = out + (long long)r * (na + nb);
  int i = 0, j = 0;
  for (int k = 0; k < na + nb; ++k) {
    if (j >= nb || (i < na && ar[i] <= br[j])) o[k] = ar[i++];
    else o[k] = br[j++];
  }
}
int launch_merge_sorted(int R, int na, int nb, const float* a, const float* b, float* out, cudaStream_t s) {
  merge_sorted_kernel<<<cdiv(R, 128), 128, 0, s>>>(R, na, nb, a, b, out);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

}  // namespace nrw
