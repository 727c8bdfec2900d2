/* nrw.h - C ABI of the B200-native NeuralRecon-W per-ray training core (libnrw.so).
 *
 * The reference (zju3dv/NeuralRecon-W) is pure Python: its seam for this path is the
 * duck-typed Python object NeuconWRenderer (rendering/renderer.py:51-961) plus the
 * nn.Modules NeuconW (models/neuconw.py:299-376) and NeRF (models/nerf.py:86-184).  This
 * header declares what a reference-side binding (ctypes) would bind for
 * each of those entry points.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked host;
 *   - the caller (PyTorch) owns all memory; the library never allocates device memory;
 *     scratch is passed in through nrw_ctx_bind and sized by nrw_*_bytes();
 *   - every call is asynchronous on the cudaStream_t passed as `stream` (a void*);
 *   - return 0 (NRW_OK) or a negative nrw_status; message via nrw_last_error() (thread-local);
 *   - one process per GPU; a context is bound to the device current at creation.
 */
#ifndef NRW_H_
#define NRW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NRW_API __attribute__((visibility("default")))
#else
#define NRW_API
#endif

typedef enum {
  NRW_OK = 0,
  NRW_ERR_ARG = -1,       /* bad argument / unsupported configuration */
  NRW_ERR_CUDA = -2,      /* CUDA runtime or driver error */
  NRW_ERR_WORKSPACE = -3, /* bound workspace / packed-weight buffer too small */
  NRW_ERR_STATE = -4      /* call order violated (e.g. render before pack) */
} nrw_status;

typedef struct nrw_ctx nrw_ctx;

/* GEMM backends: 0 = tcgen05 tensor cores (product path), 1 = fp32 CUDA cores (verification). */
#define NRW_GEMM_TCGEN05 0
#define NRW_GEMM_SIMT 1

NRW_API const char* nrw_last_error(void);
NRW_API int nrw_version(void);

/* ---- parameter layout ------------------------------------------------------------------
 * All trainable tensors of NeuconWSystem (embedding_a, neuconw, nerf; SURVEY.md 9.4 /
 * lightning_modules/neuconw_system.py:74-103) live in ONE flat fp32 buffer; gradients in a
 * second buffer with the same layout (single NCCL all-reduce).  The table below is the
 * single source of truth for names / shapes / offsets (in floats). */
typedef struct {
  const char* name; /* reference state_dict key, e.g. "neuconw.sdf_net.lin0.weight_v" */
  int rows, cols;   /* cols == 0 for 1-D tensors, rows == cols == 0 for scalars */
  long long offset; /* float offset into the flat buffer */
  long long numel;
} nrw_param_info;
NRW_API int nrw_param_count(void);
NRW_API int nrw_param_table(int n_vocab, int n_a, nrw_param_info* out /* host, nrw_param_count() */);
NRW_API long long nrw_param_total(int n_vocab, int n_a);

/* ---- context ---------------------------------------------------------------------------
 * n_planes: bf16 operand planes per fp32 tensor (1 = bf16, 2 = bf16x3 products, 3 = bf16x6
 * products ~ fp32).  chunk_rows: samples per MLP chunk (multiple of 128). */
NRW_API int nrw_ctx_create(nrw_ctx** out, int n_planes, int gemm_backend, int n_vocab, int n_a);
NRW_API int nrw_ctx_destroy(nrw_ctx* ctx);
/* mixed precision: the backward GEMMs use only the first n operand planes (0 = same as forward).  n = 1 with
 * n_planes = 2 keeps every rendered output at split-bf16 accuracy and computes gradients in plain bf16. */
NRW_API int nrw_ctx_set_backward_planes(nrw_ctx* ctx, int n);
/* planes of the stored softplus outputs read by the BACKWARD sweeps to rebuild the gates softplus'(a), softplus''(a)
 * (0 = all forward planes; 1 halves that stream at ~1e-3 relative gate error).  The forward gradient chain always reads all. */
NRW_API int nrw_ctx_set_backward_gate_planes(nrw_ctx* ctx, int n);
NRW_API long long nrw_packed_bytes(const nrw_ctx* ctx);
/* n_slots_sdf / n_slots_nerf: how many chunks keep their forward activations resident for the backward pass
 * (>= number of chunks of a batch: no forward recompute in backward; 1: recompute, minimum memory). */
NRW_API long long nrw_workspace_bytes(const nrw_ctx* ctx, int chunk_rows, int with_backward, int max_rays,
                                      int max_samples_per_ray, int n_slots_sdf, int n_slots_nerf);
NRW_API int nrw_ctx_bind(nrw_ctx* ctx, void* packed, long long packed_bytes, void* workspace,
                         long long workspace_bytes, int chunk_rows, int with_backward, int max_rays,
                         int max_samples_per_ray, int n_slots_sdf, int n_slots_nerf, void* stream);
/* weight-norm materialisation + bf16 plane split + transposes of every layer (replaces what
 * torch.nn.utils.weight_norm recomputes on every call, models/neuconw.py:104-105,256-257). */
NRW_API int nrw_pack_weights(nrw_ctx* ctx, const float* params, void* stream);

/* ---- NeuconWRenderer.sdf / NeuconW.sdf  (rendering/renderer.py:947-949) -----------------
 * With two-plane operands on the tcgen05 backend the whole query is ONE launch of the fused on-chip chain (encoding, 8 layers,
 * head; 12 B in / 4 B out of HBM per point) and touches no workspace; otherwise it runs chunk by chunk through the bound
 * workspace.  Results do not depend on how the caller batches the points. */
NRW_API int nrw_sdf_query(nrw_ctx* ctx, const float* pts /*[n,3]*/, long long n, float* sdf /*[n]*/,
                          void* stream);
/* NeuconW.forward pieces (models/neuconw.py:339-376): sdf, features' consumer rgb, normals. */
NRW_API int nrw_neuconw_forward(nrw_ctx* ctx, const float* pts /*[n,3]*/, const float* dirs /*[n,3]*/,
                                const float* a /*[n,n_a]*/, long long n, float* rgb /*[n,3]*/,
                                float* sdf /*[n]*/, float* normals /*[n,3]*/, void* stream);
/* NeRF.forward (models/nerf.py:156-182): pts4 [n,4], dirs [n,3], a [n,n_a] -> density[n], rgb[n,3] */
NRW_API int nrw_nerf_forward(nrw_ctx* ctx, const float* pts4, const float* dirs, const float* a,
                             long long n, float* density, float* rgb, void* stream);

/* ---- NeuconWRenderer.sparse_sampler (rendering/renderer.py:458-568) ---------------------- */
typedef struct {
  int n_samples, n_importance, up_sample_steps, n_outside, s_val_base;
  int boundary_samples; /* only used when sample_near/sample_far are given */
  int perturb;          /* 0/1: u_ray / u_out must be given when 1 */
} nrw_sampler_cfg;
/* o,d [R,3] (unit-sphere frame), near,far [R]; sample_near/sample_far [R] or NULL (no fine octree);
 * u_ray [R], u_out [R,n_outside] uniform draws or NULL.  Outputs: z_vals [R,S], z_out [R,n_outside],
 * sample_dist [R].  Optional trace (may be NULL): inds int32 [steps,R,n_imp/steps] (searchsorted
 * indices), order int32 [steps,R,S_round_max] (merge permutation). */
NRW_API int nrw_sample(nrw_ctx* ctx, const nrw_sampler_cfg* cfg, int R, const float* o, const float* d,
                       const float* near, const float* far, const float* sample_near,
                       const float* sample_far, const float* u_ray, const float* u_out, float* z_vals,
                       float* z_out, float* sample_dist, int32_t* trace_inds, int32_t* trace_order,
                       void* stream);
NRW_API int nrw_samples_per_ray(const nrw_sampler_cfg* cfg, int with_fine_octree);
/* one up-sampling round with injected sdf (stage-wise bit-exactness test; renderer.py:257-363) */
NRW_API int nrw_upsample_round(int R, int m, int n_new, float inv_s, const float* o, const float* d,
                               const float* z /*[R,m]*/, const float* sdf /*[R,m]*/,
                               float* cdf_scratch /*[R,m]*/, float* z_new /*[R,n_new]*/,
                               float* z_merged /*[R,m+n_new]*/, int32_t* inds /*[R,n_new]*/,
                               int32_t* order /*[R,m+n_new]*/, void* stream);

/* boundary samples of the fine-sampling branch (stage-wise bit-exactness test; renderer.py:546-566):
 * z [R,S0] ascending -> out [R,S0+nb] = sort(cat(nb/2 samples on [near,z_0), nb-nb/2 on (z_last,far], z)) */
NRW_API int nrw_boundary_samples(int R, int S0, int nb, const float* near, const float* far, const float* z,
                                 float* out, void* stream);

/* ---- NeuconWRenderer.render core (rendering/renderer.py:157-228,570-783) ----------------- */
typedef struct {
  int R, S, n_outside;        /* T = S + n_outside */
  float cos_anneal_ratio;
  const float* background_rgb; /* device [3] or NULL (renderer.py:753-754) */
  int reserved0;               /* generation stamp: render_backward reuses the cached forward activations only when
                                  it equals the stamp of the render_forward call that produced them (else recompute) */
  int trim_sphere;
} nrw_render_cfg;

/* Device pointers of one render call.  Inputs first, then outputs, then the small per-sample
 * tensors the backward pass needs ("saved", caller-allocated like everything else). */
typedef struct {
  /* inputs */
  const float* o;           /* [R,3] */
  const float* d;           /* [R,3] */
  const float* z_vals;      /* [R,S] */
  const float* z_out;       /* [R,n_outside] */
  const float* sample_dist; /* [R] */
  const float* a_emb;       /* [R,n_a] */
  const float* inv_s;       /* [1] */
  /* outputs (16-key dict of renderer.py:899-916, minus the loss-side glue kept in torch) */
  float* color;             /* [R,3] */
  float* color_sphere;      /* [R,3] */
  float* color_bg;          /* [R,3] */
  float* cdf;               /* [R,S] */
  float* gradients;         /* [R,S,3] */
  float* weights;           /* [R,T] */
  float* weights_sum;       /* [R] */
  float* inside_sphere;     /* [R,S] */
  float* depth;             /* [R] */
  float* normals;           /* [R,3] */
  float* gradient_error;    /* [1] */
  /* saved for backward */
  float* sv_sdf;            /* [R,S] */
  float* sv_rgb;            /* [R,S,3] */
  float* sv_bg_alpha;       /* [R,T] */
  float* sv_bg_rgb;         /* [R,T,3] */
  float* sv_z_feed;         /* [R,T] */
  float* sv_relax_sum;      /* [1] */
} nrw_render_io;

NRW_API int nrw_render_forward(nrw_ctx* ctx, const nrw_render_cfg* cfg, const nrw_render_io* io, void* stream);

/* upstream gradients (NULL = zero) and gradient outputs */
typedef struct {
  const float* g_color;        /* [R,3] */
  const float* g_color_sphere; /* [R,3] */
  const float* g_color_bg;     /* [R,3] */
  const float* g_cdf;          /* [R,S] */
  const float* g_gradients;    /* [R,S,3] */
  const float* g_weights;      /* [R,T] */
  const float* g_weights_sum;  /* [R] */
  const float* g_depth;        /* [R] */
  const float* g_normals;      /* [R,3] */
  const float* g_gradient_error; /* [1] */
  float* grad_params;          /* flat, same layout as params; ACCUMULATED into */
  float* grad_a_emb;           /* [R,n_a] (written) */
  float* grad_inv_s;           /* [1] (written) */
} nrw_render_grads;
NRW_API int nrw_render_backward(nrw_ctx* ctx, const nrw_render_cfg* cfg, const nrw_render_io* io,
                                const nrw_render_grads* g, void* stream);

/* stage-wise compositing (K4) with injected per-sample inputs, for parity tests */
NRW_API int nrw_composite_forward(const nrw_render_cfg* cfg, const nrw_render_io* io, const float* sdf,
                                  const float* normals_ps /*[R,S,3]*/, const float* rgb /*[R,S,3]*/,
                                  const float* bg_alpha, const float* bg_rgb, float* scratch2 /*[2]*/,
                                  void* stream);
NRW_API int nrw_composite_backward(const nrw_render_cfg* cfg, const nrw_render_io* io,
                                   const nrw_render_grads* g, const float* normals_ps, float* d_sdf,
                                   float* d_normals_ps, float* d_rgb, float* d_bg_alpha, float* d_bg_rgb,
                                   void* stream);

/* ---- octree near/far (tools/prepare_data/generate_voxel.py:311-439) ---------------------- */
/* octree bytes (breadth-first, one per non-leaf), prefix = exclusive popcount sum (#nodes entries),
 * pyramid int32 [2, level+2].  rays in the SfM frame.  Outputs near,far [R] (already * scale),
 * pid int32 [R] (-1 = miss), count int32 [R] (# leaf voxels hit). */
NRW_API int nrw_octree_near_far(const uint8_t* octree, const int32_t* prefix, const int32_t* pyramid_host,
                                int level, const float* rays_o, const float* rays_d, int R,
                                const float scene_origin[3], float scale, float* near, float* far,
                                int32_t* pid, int32_t* count, void* stream);
/* compacted hit list (ray_index, point_index, depth) front-to-back per ray; offsets = exclusive scan of count */
NRW_API int nrw_octree_hits(const uint8_t* octree, const int32_t* prefix, const int32_t* pyramid_host,
                            int level, const float* rays_o, const float* rays_d, int R,
                            const float scene_origin[3], float scale, const int64_t* offsets,
                            int32_t* ray_index, int32_t* point_index, float* depth, void* stream);

/* ---- octree build (K0; tools/prepare_data/generate_voxel.py:149-150 quantize_points + unbatched_points_to_octree,
 *      :173-178 scan_octrees + generate_points; called by get_octree renderer.py:137-155 and octree_update
 *      neuconw_system.py:268-312) ------------------------------------------------------------------------------- */
/* points [n,3] (float32, or float64 when points_are_f64) already normalised to the open cube (-1,1)
 * (generate_voxel.py:113-127).  Device outputs: octree uint8 [cap_nonleaf] (breadth-first child masks, zero padded),
 * prefix int32 [cap_nonleaf] (exclusive popcount sum), pyramid int32 [2, level+2], points_out int16 [cap_total,3]
 * (node coordinates of every level, breadth-first), counts_out int32 [2] = {#non-leaf nodes, #nodes}.  Nothing is
 * written out of bounds when a capacity is too small: compare counts_out with the capacities after synchronising
 * (n_points * level / n_points * (level+1) always suffice).  scratch: nrw_octree_build_scratch_bytes, 256-byte aligned. */
NRW_API long long nrw_octree_build_scratch_bytes(int n_points, int level, int cap_nonleaf);
NRW_API int nrw_octree_build(const void* points, int points_are_f64, int n_points, int level, uint8_t* octree,
                             int32_t* prefix, int32_t* pyramid, int16_t* points_out, int cap_nonleaf, int cap_total,
                             int32_t* counts_out, void* scratch, void* stream);

/* ---- fused clip + Adam on flat fp32 buffers (train.py:61 gradient_clip_val -> clip_grad_norm_;
 *      utils/__init__.py:30 torch.optim.Adam(eps=1e-7)) ---------------------------------------------------------- */
/* acc[0] (device double, zeroed by the caller) += sum g^2; call once per gradient buffer of the clipped group */
NRW_API int nrw_grad_sumsq(const float* grad, long long n, double* acc, void* stream);
/* one Adam step (t = step >= 1) on p with moments m, v; the gradient is scaled by min(1, max_norm/(sqrt(*sumsq)+1e-6))
 * when sumsq != NULL and max_norm > 0 (torch.nn.utils.clip_grad_norm_).  Nothing is read back to the host. */
NRW_API int nrw_adam_clip_step(float* p, const float* grad, float* m, float* v, long long n, const double* sumsq,
                               double max_norm, double lr, double beta1, double beta2, double eps, int step,
                               void* stream);

/* ---- ray-cache batch gather: PhototourismDataset.__getitem__ with semantics over an index vector
 *      (datasets/phototourism.py:709-724) fused with training_step's RAY_MASK_LIST filter
 *      (lightning_modules/neuconw_system.py:345-355).  cache_rays [n,12] = o3,d3,near,far,ts,label,depth,weight and
 *      cache_rgbs [n,3] are the reference's cache arrays (tools/prepare_data/prepare_data_cache.py:128-151) resident in HBM.
 *      Rows whose label equals one of mask_labels_host[0..n_mask) (<= 8 ids, host array) are dropped; kept rows are
 *      written in index order: rays [m,10] = row[0:8] ++ row[10:12], rgbs [m,3], ts [m] int64, label [m]; n_valid[0] = m
 *      (device int64).  scratch: nrw_compact_scratch_bytes(batch) bytes, 256-byte aligned. ---------------------------- */
NRW_API long long nrw_compact_scratch_bytes(long long n);
NRW_API int nrw_raycache_gather(const float* cache_rays, const float* cache_rgbs, long long n_cache, const int64_t* index,
                                int batch, const int32_t* mask_labels_host, int n_mask, float* rays, float* rgbs, int64_t* ts,
                                float* label, int64_t* n_valid, void* scratch, void* stream);
/* ---- mesh-extraction / octree-refresh query pipeline ------------------------------------------------
 * dense lattice of utils/visualization.py:42-52: out[t] = (lin_x[i], lin_y[j], lin_z[k]) for linear index i0+t =
 * (i*dim + j)*dim + k, lin_c = torch.linspace(lo[c], hi[c], dim) (float32). */
NRW_API int nrw_grid_points_dense(int dim, const float lo[3], const float hi[3], long long i0, long long n, float* out /*[n,3]*/,
                                  void* stream);
/* up-sampled sparse lattice of tools/extract_mesh.py:73-95 / neuconw_system.py:213-234: candidate i0+t -> leaf (i0+t)/up^3
 * (leaves int16 [n_leaves,3], lexicographic = torch.nonzero order), sub-voxel unravel((i0+t)%up^3); xyz_sfm (optional) =
 * float32(index)*voxel_size + vol_origin, xyz_train = (xyz_sfm - scene_origin)/scene_radius. */
NRW_API int nrw_grid_points_sparse(const int16_t* leaves, long long n_leaves, int up_times, float voxel_size,
                                   const float vol_origin[3], const float scene_origin[3], float scene_radius, long long i0,
                                   long long n, float* xyz_sfm, float* xyz_train, void* stream);
/* stable compaction xyz[sdf <= threshold] (neuconw_system.py:259), appended at out[count[0]...]; count (device int64) is
 * increased by the number of rows kept.  scratch: nrw_compact_scratch_bytes(n). */
NRW_API int nrw_threshold_compact(const float* sdf, const float* xyz /*[n,3]*/, long long n, float threshold, float* out,
                                  int64_t* count, void* scratch, void* stream);

/* ---- unit-test hooks ---------------------------------------------------------------------- */
/* D[M,N] = (sum planes of A)[M,K] * (sum planes of B)[N,K]^T from fp32 inputs: splits into planes in
 * scratch (caller-provided, nrw_gemm_test_scratch_bytes) and runs the selected backend. */
NRW_API long long nrw_gemm_test_scratch_bytes(int M, int N, int K);
NRW_API int nrw_gemm_test(int backend, int n_planes, int mn_major, int k_slices, int M, int N, int K,
                          const float* A, const float* B, const float* bias, int act, float* D,
                          void* scratch, void* stream);
NRW_API long long nrw_launch_count(void);
/* measurement: while enabled, every tcgen05 GEMM launch is bracketed by CUDA events on its stream; a call with
 * out5 != NULL synchronises those events and returns {sum of kernel ms, algorithmic FLOP (2MNK), MMA FLOP
 * (x plane products), launches, algorithmic HBM bytes (operands + epilogue streams)} since the last read
 * (bench.py roofline). */
NRW_API int nrw_gemm_timing(int enable, double* out5_host);
/* debug: per-CTA cycle attribution of the tcgen05 GEMM (u64 [SMs,16], zeroed by the caller; NULL = off) */
NRW_API int nrw_debug_gemm_profile(void* device_buf_u64);

#ifdef __cplusplus
}
#endif
#endif /* NRW_H_ */
