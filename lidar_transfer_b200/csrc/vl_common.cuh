// vl_common.cuh -- shared device structures and helpers of libvlidar (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/vlidar.h"

// ---------------------------------------------------------------------------
// error plumbing (thread-local text behind vl_last_error())
// ---------------------------------------------------------------------------
void vl_set_error(const char* fmt, ...);
void vl_count_launch();
int vl_sm_count();   // multiProcessorCount of the current device (cached per device; 148 on B200): persistent grids are sized from it

// ---------------------------------------------------------------------------
// optional per-stage device timing (CUDA events on the launching stream), off by default.
// bench.py switches it on to obtain the per-kernel durations behind the roofline figures.
// ---------------------------------------------------------------------------
enum VlStage {
  VL_ST_BOUNDS = 0, VL_ST_MORTON, VL_ST_SORT_PASS, VL_ST_EMIT_CLIMB, VL_ST_TOP_CLIMB,
  VL_ST_TRACE, VL_ST_PROJECT_SCATTER, VL_ST_PROJECT_GATHER, VL_ST_TSDF_INIT, VL_ST_TSDF_INTEGRATE,
  VL_ST_MESH_COUNT, VL_ST_MESH_SCAN, VL_ST_MESH_COMPACT, VL_ST_MESH_EMIT,
  VL_ST_BEAMS, VL_ST_CAST_INIT, VL_ST_CAST_SETUP, VL_ST_CAST_ITEMS, VL_ST_CAST_RESOLVE, VL_ST_COMPARE, VL_ST_COUNT
};
void vl_prof_begin(int stage, cudaStream_t stream);   // also opens an NVTX range named after the stage
void vl_prof_end(int stage, cudaStream_t stream);
struct VlProfScope {
  int stage; cudaStream_t stream;
  VlProfScope(int st, cudaStream_t s) : stage(st), stream(s) { vl_prof_begin(st, s); }
  ~VlProfScope() { vl_prof_end(stage, stream); }
};

#define VL_CUDA_CHECK(expr)                                                           \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      vl_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                   __LINE__);                                                         \
      return VL_ECUDA;                                                                \
    }                                                                                 \
  } while (0)

#define VL_LAUNCH_CHECK(name)                                                         \
  do {                                                                                \
    vl_count_launch();                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      vl_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));          \
      return VL_ECUDA;                                                                \
    }                                                                                 \
  } while (0)

// ---------------------------------------------------------------------------
// BVH blob layout (one caller-provided device allocation, all sections 256 B aligned)
//
//   [header 256 B][nodes 64 B x max(n-1,1)][tris 48 B x n][c0 16 B x n]
//   [sort keys 2 x 4 B x n][sort vals 2 x 4 B x n][flags 4 B x n][radix-sort scratch]
//
// HBM layout rationale: a node carries BOTH children's boxes so one 64 B (half-line)
// fetch decides both descents; a triangle record is three float4 (v0, e1, e2) so a leaf
// of <= 4 Morton-adjacent triangles is one contiguous <= 192 B run.
// ---------------------------------------------------------------------------
struct __align__(16) VlNode {
  // q[0], q[1]: left child   (min x y z, max x | max y z, child ref, first key of its range)
  // q[2], q[3]: right child  (min x y z, max x | max y z, child ref, last key of its range)
  // each half is one aligned 32 B chunk, written by the child that owns it during the build
  float4 q[4];
};

struct __align__(16) VlTri {
  float4 v0;  // xyz = vertex 0, w = original face index (bits)
  float4 e1;  // xyz = v1 - v0,  w = mean remission
  float4 e2;  // xyz = v2 - v0,  w = unused
};

struct VlHeader {
  int n_tris;
  int root_ref;
  int n_bad_faces;
  int max_climb;
  unsigned int bounds_min[3];  // order-preserving uint encoding of float
  unsigned int bounds_max[3];
  int n_pending;               // nodes queued by k_emit_climb for k_top_climb
  float box_pad;               // conservative leaf-box padding (k_morton)
  int pad[52];
};
static_assert(sizeof(VlHeader) == 256, "header is 256 B");

// A child reference: >= 0 -> inner node index; < 0 -> leaf, ~ref = (first_tri << 3) | count.
#define VL_LEAF_MAX 4
__host__ __device__ inline int vl_make_leaf(int first, int count) { return ~((first << 3) | count); }
__host__ __device__ inline int vl_leaf_first(int ref) { return (~ref) >> 3; }
__host__ __device__ inline int vl_leaf_count(int ref) { return (~ref) & 7; }

#ifndef VL_SORT_THREADS
#define VL_SORT_THREADS 256
#endif
#ifndef VL_SORT_ITEMS
#define VL_SORT_ITEMS 8
#endif
#define VL_SORT_TILE (VL_SORT_THREADS * VL_SORT_ITEMS)  // keys per radix-sort tile
#define VL_SORT_PASSES 4                                // 8-bit digits over the 32-bit Morton key
#define VL_TOP_LEVELS 11                                // levels of the tree kept in heap order for shared-memory staging
#define VL_TOP_NODES ((1 << VL_TOP_LEVELS) - 1)         // 2047 nodes x 64 B = 128 KB

// sort scratch: [digit histograms 4 x 256 u32][4 tile tickets, padded to 256 B][tile + group states 4 x n_state_tiles x 256 u32]
struct VlBlobLayout {
  size_t off_nodes, off_tris, off_c0, off_keys0, off_keys1, off_vals0, off_vals1, off_flags;
  size_t off_ghist, off_tickets, off_tile_state;
  size_t off_top;     // the top VL_TOP_LEVELS levels of the tree in heap order (k_top_pack), staged by TMA in k_trace_persistent
  size_t total;
  int n_sort_tiles;
  int n_state_tiles;  // state rows per pass: one per tile + one per group of 32 tiles
};

__host__ inline size_t vl_align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ inline VlBlobLayout vl_blob_layout(int n) {
  VlBlobLayout L;
  size_t nn = n > 0 ? (size_t)n : 1;
  size_t off = 256;
  L.off_nodes = off; off = vl_align256(off + 64 * nn);
  L.off_tris = off;  off = vl_align256(off + 48 * nn);
  L.off_c0 = off;    off = vl_align256(off + 16 * nn);
  L.off_keys0 = off; off = vl_align256(off + 4 * nn);
  L.off_keys1 = off; off = vl_align256(off + 4 * nn);
  L.off_vals0 = off; off = vl_align256(off + 4 * nn);
  L.off_vals1 = off; off = vl_align256(off + 4 * nn);
  L.off_flags = off; off = vl_align256(off + 4 * nn);
  L.n_sort_tiles = (int)((nn + VL_SORT_TILE - 1) / VL_SORT_TILE);
  L.off_ghist = off;      off = vl_align256(off + 4 * 256 * VL_SORT_PASSES);
  L.off_tickets = off;    off = vl_align256(off + 4 * VL_SORT_PASSES);
  L.n_state_tiles = L.n_sort_tiles + (L.n_sort_tiles + 31) / 32;
  L.off_tile_state = off; off = vl_align256(off + 4 * 256 * (size_t)VL_SORT_PASSES * (size_t)L.n_state_tiles);
  L.off_top = off;        off = vl_align256(off + 64 * (size_t)VL_TOP_NODES);
  L.total = off;
  return L;
}

// ---------------------------------------------------------------------------
// exactly-rounded float helpers: the parity-critical arithmetic (normalise, Moller-
// Trumbore, hit point) must not be FMA-contracted -- the canonical reference build is
// -ffp-contract=off and the oracle restatement states each rounding explicitly.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float vl_dot(float ax, float ay, float az, float bx, float by, float bz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}
__device__ __forceinline__ float3 vl_cross(float ax, float ay, float az, float bx, float by, float bz) {
  return make_float3(__fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by)),
                     __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz)),
                     __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx)));
}

// order-preserving float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned int vl_float_to_ordered(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float vl_ordered_to_float(unsigned int u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// Moller-Trumbore with the reference's operation order (Triangle.h:27-50).
// d must be the normalised direction.  Returns true and t on a hit.
__device__ __forceinline__ bool vl_tri_hit(const float4 v0, const float4 e1, const float4 e2,
                                           const float3 o, const float3 d, float* t_out) {
  const float eps = 0.000001f;
  float3 h = vl_cross(d.x, d.y, d.z, e2.x, e2.y, e2.z);
  float a = vl_dot(e1.x, e1.y, e1.z, h.x, h.y, h.z);
  if (a < eps && a > -eps) return false;
  float inv_a = __fdiv_rn(1.0f, a);
  float sx = __fsub_rn(o.x, v0.x), sy = __fsub_rn(o.y, v0.y), sz = __fsub_rn(o.z, v0.z);
  float u = __fmul_rn(vl_dot(sx, sy, sz, h.x, h.y, h.z), inv_a);
  if (u < 0.0f || u > 1.0f) return false;
  float3 q = vl_cross(sx, sy, sz, e1.x, e1.y, e1.z);
  float v = __fmul_rn(vl_dot(d.x, d.y, d.z, q.x, q.y, q.z), inv_a);
  if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;
  float t = __fmul_rn(vl_dot(e2.x, e2.y, e2.z, q.x, q.y, q.z), inv_a);
  if (t < eps) return false;
  *t_out = t;
  return true;
}

// normalize(): IEEE 1/sqrt with the reference's summation order (Vector3.h:73-89 hadd tree).
__device__ __forceinline__ float3 vl_normalize(float x, float y, float z) {
  float D = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fadd_rn(__fmul_rn(z, z), 0.0f));
  float r = __fdiv_rn(1.0f, __fsqrt_rn(D));
  return make_float3(__fmul_rn(x, r), __fmul_rn(y, r), __fmul_rn(z, r));
}

// direction of ray r: normalised here, or -- VL_RAYS_NORMALIZED -- taken as given (the host normalised it with the
// reference's own rsqrtps + Newton step, vl_normalize_rays, so vl_tri_hit sees the reference's bits)
__device__ __forceinline__ float3 vl_ray_dir(const float* __restrict__ rays, size_t r, bool prenorm) {
  const float x = __ldg(rays + 3 * r), y = __ldg(rays + 3 * r + 1), z = __ldg(rays + 3 * r + 2);
  return prenorm ? make_float3(x, y, z) : vl_normalize(x, y, z);
}

// internal launchers (implemented in the .cu files, called by vl_api.cu)
int vl_bvh_build_launch(const float* d_verts, const int* d_faces, const int* d_colors, const float* d_rem,
                        int n_verts, int n_faces, void* d_blob, cudaStream_t stream);
int vl_trace_launch(const void* d_blob, int n_faces, const float* d_rays, const float* d_origin, int n_rays,
                    int height, float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem,
                    int* d_tri_id, int flags, cudaStream_t stream);
int vl_trace_bruteforce_launch(const float* d_verts, const int* d_faces, const int* d_colors, const float* d_rem,
                               int n_verts, int n_faces, const float* d_rays, const float* d_origin, int n_rays,
                               int height, float* d_endpoints, int* d_endcolors, float* d_range,
                               float* d_endrem, int* d_tri_id, int flags, cudaStream_t stream);
// vl_cast.cu
size_t vl_beams_bytes_impl(int n_rays, int height);
size_t vl_cast_workspace_bytes_impl(int n_rays, int n_faces);
int vl_beams_build_launch(const float* d_rays, int n_rays, int height, void* d_beams, int flags, cudaStream_t stream);
int vl_cast_launch(const void* d_beams, const float* d_verts, const int* d_faces, const int* d_colors,
                   const float* d_rem, int n_verts, int n_faces, const float* d_origin, int n_rays, int height,
                   float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem, int* d_tri_id, int flags,
                   void* d_ws, cudaStream_t stream, int phases = 3);
// phases of one cast: FRONT = reset + setup + units (needs verts and faces only), RESOLVE = the write-back (needs the
// colours and remissions too) -- the host-pointer ctrace launches the front while the colours are still on their way
#define VL_CAST_PHASE_FRONT 1
#define VL_CAST_PHASE_RESOLVE 2
int vl_cast_status_read(const void* d_ws, cudaStream_t stream, int* info);
int vl_cast_graph_create_impl(const void* d_beams, const float* d_origin, int n_rays, int height, float* d_endpoints,
                              int* d_endcolors, float* d_range, float* d_endrem, int* d_tri_id, int flags, void* d_ws,
                              int max_faces, const void* h_desc, int* h_status, cudaStream_t stream, void** out_exec);
int vl_cast_graph_launch_impl(void* exec, cudaStream_t stream);
int vl_cast_graph_destroy_impl(void* exec);
