// K1a: ray / sparse-octree intersection (octree.cu)
#pragma once
#include "common.cuh"

namespace nrw {
int octree_near_far(const uint8_t* octree, const int32_t* prefix, const int32_t* pyramid_host, int level,
                    const float* rays_o, const float* rays_d, int R, const float scene_origin[3], float scale,
                    float* near, float* far, int32_t* pid, int32_t* count, cudaStream_t s);
int octree_hits(const uint8_t* octree, const int32_t* prefix, const int32_t* pyramid_host, int level,
                const float* rays_o, const float* rays_d, int R, const float scene_origin[3], float scale,
                const int64_t* offsets, int32_t* ray_index, int32_t* point_index, float* depth, cudaStream_t s);
}  // namespace nrw

// K0: octree builder (octree_build.cu)
namespace nrw {
long long octree_build_scratch_bytes(int n_points, int level, int cap_nonleaf);
int octree_build(const void* points, int is_f64, int n, int level, uint8_t* octree, int32_t* prefix, int32_t* pyramid,
                 int16_t* points_out, int cap_nonleaf, int cap_total, int32_t* counts_out, void* scratch, cudaStream_t s);
}  // namespace nrw

// fused clip + Adam (optim.cu)
namespace nrw {
int grad_sumsq(const float* g, long long n, double* acc, cudaStream_t s);
int adam_clip_step(float* p, const float* g, float* m, float* v, long long n, const double* sumsq, double max_norm, double lr,
                   double b1, double b2, double eps, int step, cudaStream_t s);
}  // namespace nrw

// data movers either side of the hot path (dataio.cu): ray-cache batch gather + label filter, query-point generators of
// the mesh-extraction / octree-refresh pipelines, stable threshold compaction
namespace nrw {
long long compact_scratch_bytes(long long n);
int raycache_gather(const float* cache_rays, const float* cache_rgbs, long long n_cache, const int64_t* index, int batch,
                    const int32_t* mask_labels, int n_mask, float* rays, float* rgbs, int64_t* ts, float* label,
                    int64_t* n_valid, void* scratch, cudaStream_t s);
int threshold_compact(const float* sdf, const float* xyz, long long n, float thr, float* out, int64_t* count, void* scratch,
                      cudaStream_t s);
int grid_points_dense(int dim, const float lo[3], const float hi[3], long long i0, long long n, float* out, cudaStream_t s);
int grid_points_sparse(const int16_t* leaves, long long n_leaves, int up, float voxel, const float vol_origin[3],
                       const float scene_origin[3], float scene_radius, long long i0, long long n, float* xyz_sfm,
                       float* xyz_train, cudaStream_t s);
}  // namespace nrw
