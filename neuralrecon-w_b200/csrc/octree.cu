// K1a: ray <-> sparse-octree intersection (replaces kaolin.render.spc.unbatched_raytrace as used by
// get_near_far, tools/prepare_data/generate_voxel.py:311-439).
//
// Input is Kaolin's SPC encoding: `octree` = one occupancy byte per non-leaf node in breadth-first
// order (bit j set <=> child with Morton digit j = x<<2|y<<1|z exists), `prefix` = exclusive popcount
// sum (children of node i start at hierarchy index 1 + prefix[i]), `pyramid[1][l]` = first hierarchy
// index of level l.  One thread walks one ray depth-first; a voxel is reported iff the slab test below
// passes for it AND for all of its ancestors - a pure function of (ray, voxel), so the hit SET does not
// depend on traversal order and is bit-reproducible against oracle/octree_port.py.
#include "../../include/nrw_math.h"
#include "octree.h"

namespace nrw {

struct RayN { float o[3], d[3]; };

__device__ __forceinline__ RayN normalise_ray(const float* ro, const float* rd, int r, float ox, float oy, float oz,
                                              float scale) {
  RayN q;
  const float so[3] = {ox, oy, oz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    q.d[a] = NRW_ADD(rd[r * 3 + a], 1e-7f);                                   // generate_voxel.py:332
    q.o[a] = NRW_DIV(NRW_SUB(NRW_ADD(ro[r * 3 + a], 1e-7f), so[a]), scale);   // :333,345
  }
  return q;
}

// slab test against the voxel (x,y,z) of `level`; returns entry depth (>= 0) or -1 when missed
__device__ __forceinline__ float slab(const RayN& q, int x, int y, int z, int level) {
  const float r = 1.0f / (float)(1 << level);
  const int p[3] = {x, y, z};
  float tmin = -INFINITY, tmax = INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float c = NRW_SUB(NRW_MUL(r, (float)(2 * p[a] + 1)), 1.0f);
    const float t1 = NRW_DIV(NRW_SUB(NRW_SUB(c, r), q.o[a]), q.d[a]);
    const float t2 = NRW_DIV(NRW_SUB(NRW_ADD(c, r), q.o[a]), q.d[a]);
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
  }
  if (!(tmax >= tmin) || !(tmax >= 0.0f)) return -1.0f;
  return fmaxf(tmin, 0.0f);
}

static constexpr int MAX_LEVEL = 16;

// mode 0: near/far/pid/count.  mode 1: write the hit list at offsets[r] and sort it front-to-back.
template <int MODE>
__global__ void octree_trace_kernel(const uint8_t* __restrict__ octree, const int32_t* __restrict__ prefix, int level,
                                    int leaf_base, const float* __restrict__ ro, const float* __restrict__ rd, int R,
                                    float ox, float oy, float oz, float scale, float* __restrict__ near,
                                    float* __restrict__ far, int32_t* __restrict__ pid, int32_t* __restrict__ count,
                                    const int64_t* __restrict__ offsets, int32_t* __restrict__ ray_index,
                                    int32_t* __restrict__ point_index, float* __restrict__ depth) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const RayN q = normalise_ray(ro, rd, r, ox, oy, oz, scale);
  int node[MAX_LEVEL + 1], child[MAX_LEVEL + 1], cx[MAX_LEVEL + 1], cy[MAX_LEVEL + 1], cz[MAX_LEVEL + 1];
  float tn = INFINITY, tf = -INFINITY;
  int best = -1, n_hit = 0;
  const long long base = MODE == 1 ? offsets[r] : 0;
  int l = 0;
  node[0] = 0; child[0] = 0; cx[0] = cy[0] = cz[0] = 0;
  if (slab(q, 0, 0, 0, 0) < 0.0f) l = -1;
  while (l >= 0) {
    if (child[l] >= 8) { --l; continue; }
    const int j = child[l]++;
    const uint8_t byte = octree[node[l]];
    if (!((byte >> j) & 1)) continue;
    const int nx = cx[l] * 2 + ((j >> 2) & 1), ny = cy[l] * 2 + ((j >> 1) & 1), nz = cz[l] * 2 + (j & 1);
    const float t = slab(q, nx, ny, nz, l + 1);
    if (t < 0.0f) continue;
    const int idx = 1 + prefix[node[l]] + __popc((unsigned)byte & ((1u << j) - 1u));
    if (l + 1 == level) {
      if (MODE == 0) {
        if (t < tn || (t == tn && idx < best)) { tn = t; best = idx; }
        if (t > tf) tf = t;
      } else {
        ray_index[base + n_hit] = r;
        point_index[base + n_hit] = idx;
        depth[base + n_hit] = t;
      }
      ++n_hit;
    } else {
      ++l;
      node[l] = idx; child[l] = 0; cx[l] = nx; cy[l] = ny; cz[l] = nz;
    }
  }
  (void)leaf_base;
  if (MODE == 0) {
    // post-processing of get_near_far (generate_voxel.py:393-400,437-439)
    float nr = n_hit ? tn : 0.0f, fr = n_hit ? tf : 0.0f;
    int pd = n_hit ? best : -1;
    if (!(nr > 1e-4f)) { nr = 0.0f; fr = 0.0f; pd = -1; }
    near[r] = NRW_MUL(nr, scale);
    far[r] = NRW_MUL(fr, scale);
    pid[r] = pd;
    count[r] = n_hit;
  } else {
    // insertion sort by (depth, point index): front-to-back, Morton order on ties
    for (int i = 1; i < n_hit; ++i) {
      const float dk = depth[base + i];
      const int pk = point_index[base + i];
      int j = i - 1;
      while (j >= 0 && (depth[base + j] > dk || (depth[base + j] == dk && point_index[base + j] > pk))) {
        depth[base + j + 1] = depth[base + j];
        point_index[base + j + 1] = point_index[base + j];
        --j;
      }
      depth[base + j + 1] = dk;
      point_index[base + j + 1] = pk;
    }
  }
}

int octree_near_far(const uint8_t* octree, const int32_t* prefix, const int32_t* pyramid_host, int level,
                    const float* rays_o, const float* rays_d, int R, const float so[3], float scale, float* near,
                    float* far, int32_t* pid, int32_t* count, cudaStream_t s) {
  NRW_CHECK(level >= 1 && level <= MAX_LEVEL, NRW_ERR_ARG, "octree: level %d out of range", level);
  const int leaf_base = pyramid_host ? pyramid_host[(level + 2) + level] : 0;
  octree_trace_kernel<0><<<cdiv(R, 128), 128, 0, s>>>(octree, prefix, level, leaf_base, rays_o, rays_d, R, so[0], so[1],
                                                      so[2], scale, near, far, pid, count, nullptr, nullptr, nullptr,
                                                      nullptr);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

int octree_hits(const uint8_t* octree, const int32_t* prefix, const int32_t* pyramid_host, int level,
                const float* rays_o, const float* rays_d, int R, const float so[3], float scale, const int64_t* offsets,
                int32_t* ray_index, int32_t* point_index, float* depth, cudaStream_t s) {
  NRW_CHECK(level >= 1 && level <= MAX_LEVEL, NRW_ERR_ARG, "octree: level %d out of range", level);
  const int leaf_base = pyramid_host ? pyramid_host[(level + 2) + level] : 0;
  octree_trace_kernel<1><<<cdiv(R, 128), 128, 0, s>>>(octree, prefix, level, leaf_base, rays_o, rays_d, R, so[0], so[1],
                                                      so[2], scale, nullptr, nullptr, nullptr, nullptr, offsets,
                                                      ray_index, point_index, depth);
  NRW_LAUNCH_OK();
  return NRW_OK;
}

}  // namespace nrw
