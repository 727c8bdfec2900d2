// Orchestration of the per-ray hot path over L2-sized chunks of samples.
#pragma once
#include <vector>

#include "gemm.h"
#include "params.h"
#include "pointwise.h"

// Forward activations of one chunk.  With enough HBM (180 GB on B200) every chunk of a training batch
// keeps its own slot, so the backward pass consumes them directly instead of recomputing the forward.
struct FwdSdfSlot {
  float* PTS;
  nrw::Planes U0, U[9], G[8], FEAT;
  float* Q[8];
  nrw::bf16* Qh[8];     // bf16 twin of Q[l] (l != 0, 4) when the context keeps backward-only side streams in bf16
  float *c_sdf, *c_nrm;
  float* HP;            // [Mc, 8] row partials of the fused SDF head (forward-only queries)
  nrw::Planes IN1, H1, IN2, X[5];
  float* c_rgb;
};
struct FwdNerfSlot {
  nrw::Planes IN0, NH[9], IN5, FEATN, AP[5];
  float *c_density, *c_alpha, *c_rgbbg, *c_dists;
};

struct nrw_ctx {
  int n_planes = 2, backend = 0, n_vocab = 0, n_a = 48;
  int bwd_planes = 0;   // 0: same as n_planes; 1: 'mixed' precision (backward GEMMs use the hi plane only)
  int cur_planes = 2;   // planes used by the GEMM helpers of the pass in flight
  int bwd_gate_planes = 0;   // planes of u read for the softplus gates of the BACKWARD sweeps (0 = all forward planes)
  int gate_planes() const { return bwd_gate_planes > 0 ? bwd_gate_planes : n_planes; }
  std::vector<FwdSdfSlot> sdf_slots;
  std::vector<FwdNerfSlot> nerf_slots;
  int n_slots_sdf = 1, n_slots_nerf = 1;
  bool fwd_cached = false;          // slots hold the forward of the last render_forward call
  int cached_R = 0, cached_S = 0, cached_T = 0, cached_gen = 0;   // gen: nrw_render_cfg::reserved0 of that call
  void use_sdf_slot(int i) {
    const FwdSdfSlot& s = sdf_slots[i];
    PTS = s.PTS; U0 = s.U0; FEAT = s.FEAT; c_sdf = s.c_sdf; c_nrm = s.c_nrm; HP = s.HP;
    for (int l = 0; l < 9; ++l) U[l] = s.U[l];
    for (int l = 0; l < 8; ++l) { G[l] = s.G[l]; Q[l] = s.Q[l]; Qh[l] = s.Qh[l]; }
    IN1 = s.IN1; H1 = s.H1; IN2 = s.IN2; c_rgb = s.c_rgb;
    for (int l = 0; l < 5; ++l) X[l] = s.X[l];
  }
  void use_nerf_slot(int i) {
    const FwdNerfSlot& s = nerf_slots[i];
    IN0 = s.IN0; IN5 = s.IN5; FEATN = s.FEATN;
    for (int l = 0; l < 9; ++l) NH[l] = s.NH[l];
    for (int l = 0; l < 5; ++l) AP[l] = s.AP[l];
    c_density = s.c_density; c_alpha = s.c_alpha; c_rgbbg = s.c_rgbbg; c_dists = s.c_dists;
  }
  std::vector<nrw::ParamInfo> tab;
  nrw::PackedModel pm;
  char* packed = nullptr;
  nrw::bf16* bf_area = nullptr;
  float* f_area = nullptr;
  bool bound = false, packed_valid = false;
  const float* params = nullptr;
  int Mc = 0, with_bwd = 0, max_rays = 0, max_T = 0;

  // ---- chunk workspace (rows = Mc) ----
  float* PTS = nullptr;
  nrw::Planes U0, U[9], G[8], FEAT;
  float* Q[8] = {nullptr};   // Q[0] is [Mc,64]
  nrw::bf16* Qh[8] = {nullptr};
  nrw::bf16* DA2h[8] = {nullptr};
  bool aux_bf16 = false;     // 'mixed': Q_l (l != 0, 4) and the second-order terms DA2_l are stored as one bf16 plane
  float* c_sdf = nullptr;
  float* HP = nullptr;
  float* c_nrm = nullptr;
  nrw::Planes IN1, H1, IN2, X[5];
  float* c_rgb = nullptr;
  nrw::Planes IN0, NH[9], IN5, FEATN, AP[5];
  float *c_density = nullptr, *c_alpha = nullptr, *c_rgbbg = nullptr, *c_dists = nullptr;
  // backward
  nrw::Planes DQ0, DQodd, DQeven, DQ4, DA[2], DFEAT;
  float* DQ8f = nullptr;
  float* DA2[8] = {nullptr};
  nrw::Planes dX[2], dH2, dH1, dXF, dNA[2], dNF, dNH[2];
  float *tail = nullptr, *c_dn = nullptr, *c_ddens = nullptr, *c_dpre3 = nullptr;
  float* gs = nullptr;       // gradient scratch (packed layout)
  float* ge_acc = nullptr;   // [2]
  // ---- per-call global arrays (rows = max_rays * max_T) ----
  float *gz[2] = {nullptr, nullptr}, *gsdf[2] = {nullptr, nullptr}, *gznew = nullptr, *gsdfnew = nullptr,
        *gcdf = nullptr;
  int32_t* gorder = nullptr;
  float *g_dsdf = nullptr, *g_dnrm = nullptr, *g_drgb = nullptr, *g_dbga = nullptr, *g_dbgc = nullptr;
  float* g_pts = nullptr;    // [max_rays*max_T, 3]

  nrw::Planes W(int layer) const {
    return nrw::Planes{bf_area + pm.layers[layer].W_off, pm.plane_stride[layer], pm.layers[layer].Kp};
  }
  nrw::Planes WT(int layer) const {
    return nrw::Planes{bf_area + pm.layers[layer].WT_off, pm.plane_stride[layer], pm.layers[layer].Np};
  }
  const float* bias(int layer) const { return f_area + pm.layers[layer].bias_off; }
  float* dW(int layer) const { return gs + pm.layers[layer].dW_off; }
  float* db(int layer) const { return gs + pm.layers[layer].db_off; }
};

namespace nrw {

long long workspace_bytes(const nrw_ctx& c, int chunk_rows, int with_bwd, int max_rays, int max_T, int n_slots_sdf,
                          int n_slots_nerf);
int carve_workspace(nrw_ctx& c, void* base, long long bytes, int chunk_rows, int with_bwd, int max_rays,
                    int max_T, int n_slots_sdf, int n_slots_nerf, cudaStream_t s);

// SDF value (+ normals, + feature planes) for M rows at positions pts [M,3]; results in c.c_sdf / c.c_nrm / c.FEAT
int sdf_chunk_forward(nrw_ctx& c, int M, const float* pts, bool need_normal, bool need_feat, cudaStream_t s);
int color_chunk_forward(nrw_ctx& c, int M, const float* pts, const float* dirs, const float* a, int rows_per_src,
                        cudaStream_t s);
int nerf_chunk_forward(nrw_ctx& c, int M, const float* o, const float* d, const float* z, const float* sdist,
                       const float* pts4, const float* a, int T, int rows_per_src, cudaStream_t s);
// backward of the three networks for the chunk currently resident in the workspace
int color_chunk_backward(nrw_ctx& c, int M, const float* d_rgb, const float* d_nrm_comp, int rows_per_src,
                         float* d_a_rays, int R_chunk, cudaStream_t s);
int sdf_chunk_backward(nrw_ctx& c, int M, const float* pts, const float* d_sdf, cudaStream_t s);
int nerf_chunk_backward(nrw_ctx& c, int M, const float* d_bga, const float* d_bgc, float* d_a_rays, int R_chunk,
                        int T, cudaStream_t s);

int sdf_query(nrw_ctx& c, const float* pts, long long n, float* sdf, cudaStream_t s);
int sample(nrw_ctx& c, const nrw_sampler_cfg& cfg, int R, const float* o, const float* d, const float* near,
           const float* far, const float* s_near, const float* s_far, const float* u_ray, const float* u_out,
           float* z_vals, float* z_out, float* sample_dist, int32_t* trace_inds, int32_t* trace_order,
           cudaStream_t s);
int render_forward(nrw_ctx& c, const nrw_render_cfg& cfg, const nrw_render_io& io, cudaStream_t s);
int render_backward(nrw_ctx& c, const nrw_render_cfg& cfg, const nrw_render_io& io, const nrw_render_grads& g,
                    cudaStream_t s);

}  // namespace nrw
