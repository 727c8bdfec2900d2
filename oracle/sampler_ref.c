/* oracle/sampler_ref.c - TEST INFRASTRUCTURE (never linked into libnrw.so).
 *
 * Plain-C restatement of one importance-resampling round of NeuralRecon-W's voxel-guided sampler:
 *   NeuconWRenderer.up_sample  (rendering/renderer.py:257-341)
 *   sample_pdf(det=True)       (rendering/renderer.py:15-48)
 *   NeuconWRenderer.cat_z_vals (rendering/renderer.py:343-363; torch.sort replaced by a stable merge of the
 *                               two ascending runs, ties: existing sample first)
 * and of the coarse / outside strata of sparse_sampler (renderer.py:488-514).
 *
 * Every arithmetic step is a single IEEE-754 binary32 operation in the order written down in
 * include/nrw_math.h (compile with -ffp-contract=off), with sequential fp32 accumulation of the weights
 * and of the cdf.  The CUDA kernel (csrc/sampler.cu) implements the same sequence, so searchsorted indices,
 * merge permutation and z values must agree BIT FOR BIT (tests/test_sampler_bitexact.py).  Agreement with
 * the reference's torch ops is statistical (torch's CPU cumsum accumulates in double, its exp/sigmoid come
 * from a vector math library): pinned by tests/golden to ~1e-6 relative on z with identical indices.
 *
 * Build: make -C oracle   ->  oracle/_build/libsampler_ref.so
 */
#include "../include/nrw_math.h"

void nrw_ref_upsample_round(int R, int m, int n_new, float inv_s, const float* o, const float* d, const float* z_all,
                            const float* sdf_all, float* cdf_all, float* z_new_all, float* z_merged_all,
                            int32_t* inds_all, int32_t* order_all) {
  for (int r = 0; r < R; ++r) {
    const float* oo = o + 3 * r;
    const float* dd = d + 3 * r;
    const float* z = z_all + (long)r * m;
    const float* sdf = sdf_all + (long)r * m;
    float* cdf = cdf_all + (long)r * m;
    float* z_new = z_new_all + (long)r * n_new;
    float* zm = z_merged_all + (long)r * (m + n_new);
    int32_t* inds = inds_all ? inds_all + (long)r * n_new : 0;
    int32_t* order = order_all ? order_all + (long)r * (m + n_new) : 0;
    /* per-interval weights (renderer.py:263-312), unnormalised, + 1e-5 (renderer.py:19) */
    float prev_cos_raw = 0.0f, T = 1.0f, wsum = 0.0f;
    float px = NRW_ADD(oo[0], NRW_MUL(dd[0], z[0])), py = NRW_ADD(oo[1], NRW_MUL(dd[1], z[0])),
          pz = NRW_ADD(oo[2], NRW_MUL(dd[2], z[0]));
    float rad_prev = NRW_SQRT(NRW_ADD(NRW_ADD(NRW_MUL(px, px), NRW_MUL(py, py)), NRW_MUL(pz, pz)));
    for (int j = 0; j + 1 < m; ++j) {
      const float z0 = z[j], z1 = z[j + 1], s0 = sdf[j], s1 = sdf[j + 1];
      px = NRW_ADD(oo[0], NRW_MUL(dd[0], z1));
      py = NRW_ADD(oo[1], NRW_MUL(dd[1], z1));
      pz = NRW_ADD(oo[2], NRW_MUL(dd[2], z1));
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
      const float w = NRW_ADD(NRW_MUL(alpha, T), 1e-5f);
      T = NRW_MUL(T, NRW_ADD(NRW_SUB(1.0f, alpha), 1e-7f));
      cdf[j + 1] = w;
      wsum = NRW_ADD(wsum, w);
    }
    cdf[0] = 0.0f;
    float run = 0.0f;
    for (int j = 1; j < m; ++j) {
      run = NRW_ADD(run, NRW_DIV(cdf[j], wsum));
      cdf[j] = run;
    }
    /* inverse cdf at the stratified midpoints, searchsorted(right=True) (renderer.py:24-46) */
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
    /* stable merge */
    int a = 0, b = 0;
    for (int k = 0; k < m + n_new; ++k) {
      const int take_a = (b >= n_new) || (a < m && z[a] <= z_new[b]);
      if (take_a) { zm[k] = z[a]; if (order) order[k] = a; ++a; }
      else { zm[k] = z_new[b]; if (order) order[k] = m + b; ++b; }
    }
  }
}

/* coarse + outside strata (renderer.py:488-514); u_ray/u_out may be NULL when perturb == 0 */
void nrw_ref_coarse(int R, int n_samples, int n_outside, int perturb, const float* near, const float* far,
                    const float* s_near, const float* s_far, const float* u_ray, const float* u_out, float* z,
                    float* z_out, float* sample_dist) {
  for (int r = 0; r < R; ++r) {
    const float sn = s_near ? s_near[r] : near[r], sf = s_far ? s_far[r] : far[r];
    const float range = NRW_SUB(sf, sn), ns = (float)n_samples;
    sample_dist[r] = NRW_DIV(range, ns);
    float shift = 0.0f;
    if (perturb) shift = NRW_DIV(NRW_MUL(NRW_MUL(range, NRW_SUB(u_ray[r], 0.5f)), 2.0f), ns);
    for (int j = 0; j < n_samples; ++j) {
      float v = NRW_ADD(sn, NRW_MUL(range, nrw_linspace_f32(0.0f, 1.0f, n_samples, j)));
      if (perturb) v = NRW_ADD(v, shift);
      z[(long)r * n_samples + j] = v;
    }
    const int no = n_outside;
    if (no > 0) {
      const float hi = (float)(1.0 - 1.0 / ((double)no + 1.0));
      const float add = (float)(1.0 / (double)n_samples);
      for (int j = 0; j < no; ++j) {
        const int jj = no - 1 - j;
        float b = nrw_linspace_f32(1e-3f, hi, no, jj);
        if (perturb) {
          const float bl = jj > 0 ? nrw_linspace_f32(1e-3f, hi, no, jj - 1) : b;
          const float bu = jj + 1 < no ? nrw_linspace_f32(1e-3f, hi, no, jj + 1) : b;
          const float lower = jj > 0 ? NRW_MUL(0.5f, NRW_ADD(b, bl)) : b;
          const float upper = jj + 1 < no ? NRW_MUL(0.5f, NRW_ADD(bu, b)) : b;
          b = NRW_ADD(lower, NRW_MUL(NRW_SUB(upper, lower), u_out[(long)r * no + jj]));
        }
        z_out[(long)r * no + j] = NRW_ADD(NRW_DIV(far[r], b), add);
      }
    }
  }
}

/* Boundary samples of the fine-sampling branch (rendering/renderer.py:546-566): bound_near_num = nb/2 values
 * near + (z_0 - near) * linspace(0,1,n_near+1)[:-1], bound_far_num = nb - nb/2 values
 * z_last + (far - z_last) * linspace(0,1,n_far+1)[1:], concatenated with z and sorted (torch.sort).  Each run is
 * monotone (DESCENDING when the sampling window starts before near / ends after far), so the sort is restated as a
 * 3-way merge that walks descending runs from their end; the sorted VALUES are what the reference produces. */
void nrw_ref_boundary(int R, int S0, int nb, const float* near, const float* far, const float* z_all, float* out_all) {
  const int n_near = nb / 2, n_far = nb - n_near;
  for (int r = 0; r < R; ++r) {
    const float* zr = z_all + (long)r * S0;
    float* o = out_all + (long)r * (S0 + nb);
    const float z0 = zr[0], zl = zr[S0 - 1], nr = near[r], fr = far[r];
    const int a_desc = z0 < nr, b_desc = fr < zl;
    int a = 0, b = 0, c = 0;
    for (int k = 0; k < S0 + nb; ++k) {
      const int ia = a_desc ? n_near - 1 - a : a, ib = b_desc ? n_far - 1 - b : b;
      const float va = a < n_near ? NRW_ADD(nr, NRW_MUL(NRW_SUB(z0, nr), nrw_linspace_f32(0.0f, 1.0f, n_near + 1, ia))) : INFINITY;
      const float vb = b < n_far ? NRW_ADD(zl, NRW_MUL(NRW_SUB(fr, zl), nrw_linspace_f32(0.0f, 1.0f, n_far + 1, ib + 1))) : INFINITY;
      const float vc = c < S0 ? zr[c] : INFINITY;
      if (a < n_near && va <= vb && va <= vc) { o[k] = va; ++a; }
      else if (b < n_far && vb <= vc) { o[k] = vb; ++b; }
      else { o[k] = vc; ++c; }
    }
  }
}
