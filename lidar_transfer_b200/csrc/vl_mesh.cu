// vl_mesh.cu -- iso-surface extraction from the TSDF volume + per-vertex label / remission lookup, sm_100a.
//
// Replaces TSDFVolume.get_volume + get_mesh (auxiliary/fusion_lidar.py:395-424): the reference copies three
// volumes to the host (3.4 GB at the default 284 M voxels), runs scikit-image's marching_cubes_lewiner on the
// CPU (:407) and looks vertex colours / remissions up with numpy (:409-423).  Here the volumes stay in HBM:
//
//   k_mesh_count  one thread per cube (cube index == voxel index of its low corner), 8 corner reads,
//                 triangle count from the 256-case table, per-chunk totals
//   k_mesh_scan   exclusive scan of the chunk totals (single CTA) -> chunk offsets + grand total
//   k_mesh_emit   same sweep; a CTA-wide exclusive scan per 256-cube slab gives every cube its output slot,
//                 so triangles come out in cube order (deterministic, no atomics)
//
// Output is an indexed triangle SOUP: 3 vertices per triangle, faces = (3t, 3t+1, 3t+2).  Vertex positions,
// world transform (verts * voxel_size + origin, float32, :412), nearest-voxel lookup (np.round = half-to-even,
// :409), colour split and the uint8 wrap (:417-423) follow the reference; the case table is this project's own
// derivation (tools/gen_mc_table.py), and scikit-image itself is absent from the reference tree and its
// version unpinned, so parity of the TOPOLOGY against skimage is unpinned.  The oracle restates
// this file's algorithm (oracle/vl_oracle.c: vlo_mesh_extract) and must match bit for bit.
#include "vl_common.cuh"
#include "vl_mc_table.inc"

namespace {

constexpr int kThreads = 256;
constexpr int kSlabs = 64;                       // 256-cube slabs per chunk
constexpr int kChunk = kThreads * kSlabs;        // cubes per CTA

__constant__ signed char c_tri_table[256][15] = VL_MC_TRI_TABLE;
__constant__ unsigned char c_tri_count[256] = VL_MC_TRI_COUNT;
__constant__ unsigned char c_edge_corners[12][2] = VL_MC_EDGE_CORNERS;

struct MeshParams {
  int dx, dy, dz;
  float level, voxel_size, ox, oy, oz;
};

// case index of the cube whose low corner is voxel (x,y,z); -1 when the cube leaves the volume
__device__ __forceinline__ int cube_case(const float* __restrict__ tsdf, const MeshParams& P, long long vi, float* v) {
  const int yz = P.dy * P.dz;
  const int x = (int)(vi / yz);
  const int rem = (int)(vi - (long long)x * yz);
  const int y = rem / P.dz, z = rem - y * P.dz;
  if (x >= P.dx - 1 || y >= P.dy - 1 || z >= P.dz - 1) return -1;
  int mask = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    v[c] = __ldg(tsdf + vi + (long long)(c & 1) * yz + ((c >> 1) & 1) * P.dz + ((c >> 2) & 1));
    mask |= (v[c] < P.level) ? (1 << c) : 0;
  }
  return mask;
}

__global__ void __launch_bounds__(kThreads)
k_mesh_count(const float* __restrict__ tsdf, const MeshParams P, long long n_vox, int* __restrict__ chunk_count) {
  __shared__ int s_sum[kThreads / 32];
  const long long base = (long long)blockIdx.x * kChunk;
  int mine = 0;
  float v[8];
  for (int s = 0; s < kSlabs; ++s) {
    const long long vi = base + (long long)s * kThreads + threadIdx.x;
    if (vi < n_vox) {
      const int m = cube_case(tsdf, P, vi, v);
      if (m > 0) mine += c_tri_count[m];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += s_sum[w];
    chunk_count[blockIdx.x] = t;
  }
}

// single CTA: chunk_offset[c] = sum of chunk_count[0..c), total[0] = grand total (long long)
__global__ void __launch_bounds__(1024)
k_mesh_scan(const int* __restrict__ chunk_count, long long* __restrict__ chunk_offset, int n_chunks,
            long long* __restrict__ total) {
  __shared__ long long warp_sums[32];
  __shared__ long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_chunks; base += 1024) {
    const int idx = base + tid;
    const long long v = idx < n_chunks ? (long long)chunk_count[idx] : 0ll;
    long long incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const long long ws = warp_sums[lane];
      long long wi = ws;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, wi, off);
        if (lane >= off) wi += t;
      }
      warp_sums[lane] = wi - ws;
    }
    __syncthreads();
    const long long excl = carry_s + warp_sums[wid] + (incl - v);
    if (idx < n_chunks) chunk_offset[idx] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) total[0] = carry_s;
}

__global__ void __launch_bounds__(kThreads)
k_mesh_emit(const float* __restrict__ tsdf, const float* __restrict__ color_vol, const float* __restrict__ rem_vol,
            const MeshParams P, long long n_vox, const long long* __restrict__ chunk_offset, long long capacity,
            float* __restrict__ verts, int* __restrict__ faces, float* __restrict__ norms,
            unsigned char* __restrict__ colors, float* __restrict__ rem_out) {
  __shared__ int s_warp[kThreads / 32];
  __shared__ long long s_run;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long base = (long long)blockIdx.x * kChunk;
  if (tid == 0) s_run = chunk_offset[blockIdx.x];
  __syncthreads();
  const int yz = P.dy * P.dz;
  for (int s = 0; s < kSlabs; ++s) {
    const long long vi = base + (long long)s * kThreads + tid;
    float v[8];
    int m = 0;
    if (vi < n_vox) m = cube_case(tsdf, P, vi, v);
    const int cnt = m > 0 ? c_tri_count[m] : 0;
    // CTA-wide exclusive scan of cnt
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int before = 0, slab_total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      const int t = s_warp[w];
      if (w < wid) before += t;
      slab_total += t;
    }
    const long long out0 = s_run + before + (incl - cnt);
    __syncthreads();
    if (tid == 0) s_run += slab_total;
    if (cnt > 0) {
      const int x = (int)(vi / yz);
      const int r2 = (int)(vi - (long long)x * yz);
      const int y = r2 / P.dz, z = r2 - y * P.dz;
      for (int t = 0; t < cnt; ++t) {
        const long long tri = out0 + t;
        if (tri >= capacity) break;
        float pw[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int e = c_tri_table[m][3 * t + k];
          const int ca = c_edge_corners[e][0], cb = c_edge_corners[e][1];
          // vertex on the edge ca -> cb (cb = ca + one axis step), float32 like skimage's output
          const float va = v[ca], vb = v[cb];
          const float tt = __fdiv_rn(__fsub_rn(P.level, va), __fsub_rn(vb, va));
          float pv[3] = {(float)(x + (ca & 1)), (float)(y + ((ca >> 1) & 1)), (float)(z + ((ca >> 2) & 1))};
          const int axis = (ca ^ cb) == 1 ? 0 : ((ca ^ cb) == 2 ? 1 : 2);
          pv[axis] = __fadd_rn(pv[axis], tt);
          // nearest voxel (np.round: half to even), clamped to the volume
          const int ix = min(max(__float2int_rn(pv[0]), 0), P.dx - 1);
          const int iy = min(max(__float2int_rn(pv[1]), 0), P.dy - 1);
          const int iz = min(max(__float2int_rn(pv[2]), 0), P.dz - 1);
          const long long ni = ((long long)ix * P.dy + iy) * P.dz + iz;
          const float rgb = __ldg(color_vol + ni);
          // fusion_lidar.py:417-423 (float32 arithmetic, then astype(uint8) wraps modulo 256)
          const float cb_ = floorf(__fdiv_rn(rgb, 65536.0f));
          const float cg_ = floorf(__fdiv_rn(__fsub_rn(rgb, __fmul_rn(cb_, 65536.0f)), 256.0f));
          const float cr_ = __fsub_rn(__fsub_rn(rgb, __fmul_rn(cb_, 65536.0f)), __fmul_rn(cg_, 256.0f));
          const long long vtx = 3 * tri + k;
          colors[3 * vtx + 0] = (unsigned char)((long long)floorf(cr_) & 255);
          colors[3 * vtx + 1] = (unsigned char)((long long)floorf(cg_) & 255);
          colors[3 * vtx + 2] = (unsigned char)((long long)floorf(cb_) & 255);
          rem_out[vtx] = __ldg(rem_vol + ni);
          // :412 verts * voxel_size + origin
          pw[k][0] = __fadd_rn(__fmul_rn(pv[0], P.voxel_size), P.ox);
          pw[k][1] = __fadd_rn(__fmul_rn(pv[1], P.voxel_size), P.oy);
          pw[k][2] = __fadd_rn(__fmul_rn(pv[2], P.voxel_size), P.oz);
          verts[3 * vtx + 0] = pw[k][0];
          verts[3 * vtx + 1] = pw[k][1];
          verts[3 * vtx + 2] = pw[k][2];
          faces[vtx] = (int)vtx;
        }
        // flat normal of the triangle for all three vertices (only consumed by meshwrite)
        const float ax = pw[1][0] - pw[0][0], ay = pw[1][1] - pw[0][1], az = pw[1][2] - pw[0][2];
        const float bx = pw[2][0] - pw[0][0], by = pw[2][1] - pw[0][1], bz = pw[2][2] - pw[0][2];
        float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
        const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
        if (nn > 0.f) { nx /= nn; ny /= nn; nz /= nn; }
        if (norms) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            norms[3 * (3 * tri + k) + 0] = nx;
            norms[3 * (3 * tri + k) + 1] = ny;
            norms[3 * (3 * tri + k) + 2] = nz;
          }
        }
      }
    }
  }
}

}  // namespace

extern "C" size_t vl_mesh_workspace_bytes(long long n_voxels) {
  const long long n_chunks = (n_voxels + kChunk - 1) / kChunk;
  return vl_align256((size_t)(n_chunks > 0 ? n_chunks : 1) * 4) + vl_align256((size_t)(n_chunks > 0 ? n_chunks : 1) * 8) + 256;
}

static int mesh_args(const char* who, const float* d_tsdf, int dx, int dy, int dz, void* d_ws, size_t ws_bytes,
                     MeshParams* P, float level, float voxel_size, const float* origin) {
  if (!d_tsdf || dx <= 0 || dy <= 0 || dz <= 0 || !d_ws || (((uintptr_t)d_ws) & 255)) {
    vl_set_error("%s: invalid argument (dim %d x %d x %d, workspace %p)", who, dx, dy, dz, d_ws);
    return VL_EINVAL;
  }
  const long long n = (long long)dx * dy * dz;
  if (ws_bytes < vl_mesh_workspace_bytes(n)) {
    vl_set_error("%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, vl_mesh_workspace_bytes(n));
    return VL_ENOSPACE;
  }
  P->dx = dx; P->dy = dy; P->dz = dz; P->level = level; P->voxel_size = voxel_size;
  P->ox = origin ? origin[0] : 0.f; P->oy = origin ? origin[1] : 0.f; P->oz = origin ? origin[2] : 0.f;
  return VL_OK;
}

extern "C" int vl_mesh_count(const float* d_tsdf, int dx, int dy, int dz, float level, void* d_workspace,
                             size_t workspace_bytes, long long* d_total, vl_stream stream_) {
  MeshParams P;
  int rc = mesh_args("vl_mesh_count", d_tsdf, dx, dy, dz, d_workspace, workspace_bytes, &P, level, 1.f, nullptr);
  if (rc) return rc;
  if (!d_total) { vl_set_error("vl_mesh_count: null d_total"); return VL_EINVAL; }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long n = (long long)dx * dy * dz;
  const int n_chunks = (int)((n + kChunk - 1) / kChunk);
  int* chunk_count = static_cast<int*>(d_workspace);
  long long* chunk_offset = reinterpret_cast<long long*>(static_cast<char*>(d_workspace) + vl_align256((size_t)n_chunks * 4));
  { VlProfScope ps(VL_ST_MESH_COUNT, stream);
  k_mesh_count<<<n_chunks, kThreads, 0, stream>>>(d_tsdf, P, n, chunk_count); }
  VL_LAUNCH_CHECK("k_mesh_count");
  { VlProfScope ps(VL_ST_MESH_SCAN, stream);
  k_mesh_scan<<<1, 1024, 0, stream>>>(chunk_count, chunk_offset, n_chunks, d_total); }
  VL_LAUNCH_CHECK("k_mesh_scan");
  return VL_OK;
}

extern "C" int vl_mesh_emit(const float* d_tsdf, const float* d_color, const float* d_rem, int dx, int dy, int dz,
                            float level, float voxel_size, const float vol_origin[3], const void* d_workspace,
                            size_t workspace_bytes, long long capacity_tris, float* d_verts, int* d_faces,
                            float* d_norms, unsigned char* d_colors, float* d_rem_out, vl_stream stream_) {
  MeshParams P;
  int rc = mesh_args("vl_mesh_emit", d_tsdf, dx, dy, dz, const_cast<void*>(d_workspace), workspace_bytes, &P, level,
                     voxel_size, vol_origin);
  if (rc) return rc;
  if (!d_color || !d_rem || !vol_origin || capacity_tris < 0 ||
      (capacity_tris > 0 && (!d_verts || !d_faces || !d_colors || !d_rem_out))) {
    vl_set_error("vl_mesh_emit: invalid argument");
    return VL_EINVAL;
  }
  if (capacity_tris == 0) return VL_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long n = (long long)dx * dy * dz;
  const int n_chunks = (int)((n + kChunk - 1) / kChunk);
  const long long* chunk_offset =
      reinterpret_cast<const long long*>(static_cast<const char*>(d_workspace) + vl_align256((size_t)n_chunks * 4));
  VlProfScope ps(VL_ST_MESH_EMIT, stream);
  k_mesh_emit<<<n_chunks, kThreads, 0, stream>>>(d_tsdf, d_color, d_rem, P, n, chunk_offset, capacity_tris, d_verts,
                                                d_faces, d_norms, d_colors, d_rem_out);
  VL_LAUNCH_CHECK("k_mesh_emit");
  return VL_OK;
}
