// vl_mesh.cu -- iso-surface extraction from the TSDF volume + per-vertex label / remission lookup, sm_100a.
//
// Replaces TSDFVolume.get_volume + get_mesh (auxiliary/fusion_lidar.py:395-424): the reference copies three
// volumes to the host (3.4 GB at the default 284 M voxels), runs scikit-image's marching_cubes_lewiner on the
// CPU (:407) and looks vertex colours / remissions up with numpy (:409-423).  Here the volumes stay in HBM:
//
//   k_mesh_count  one thread per cube (cube index == voxel index of its low corner): 4 corner reads + 4 from the
//                 neighbour lane, the 256-case index stored as one byte per cube, per-chunk triangle totals
//   k_mesh_scan   exclusive scan of the chunk totals (single CTA) -> chunk offsets + grand total
//   k_mesh_emit   sweep over the case bytes; a CTA-wide exclusive scan per 256-cube slab gives every cube its
//                 output slot (cube order: deterministic, no atomics), then the slab's triangles are expanded one
//                 per thread so that every lane works and consecutive lanes write consecutive triangles
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

// Thread mapping of both sweeps: blockIdx.y = x (one yz-plane per grid row), blockIdx.x = chunk of kChunk
// consecutive voxels j = y * dz + z of that plane.  Chunk c of plane x has the linear chunk id x * chunks_per_plane + c;
// chunk ids increase with the voxel index, so triangles still come out in cube order.

// Case index of the cube whose low corner is voxel (x, y, z): bit c set <=> value at corner c < level; 0 when the
// cube leaves the volume.  Every lane of the warp must call this (shuffles): lane L holds voxel j, lane L+1 voxel
// j+1 = the same column one step up in z, so the four z+1 corners come from the neighbour lane instead of memory.
__device__ __forceinline__ int cube_case(const float* __restrict__ plane0, const MeshParams& P, int x, int j, bool in_plane) {
  const int yz = P.dy * P.dz;
  const int y = j / P.dz, z = j - y * P.dz;
  const bool x1 = x + 1 < P.dx, y1 = in_plane && (y + 1 < P.dy);
  // corner (cx, cy, 0)
  float v00 = 1.f, v10 = 1.f, v01 = 1.f, v11 = 1.f;
  if (in_plane) {
    v00 = __ldg(plane0 + j);
    if (x1) v10 = __ldg(plane0 + yz + j);
    if (y1) v01 = __ldg(plane0 + j + P.dz);
    if (x1 && y1) v11 = __ldg(plane0 + yz + j + P.dz);
  }
  // corner (cx, cy, 1) = the neighbour lane's (cx, cy, 0)
  float w00 = __shfl_down_sync(0xffffffffu, v00, 1), w10 = __shfl_down_sync(0xffffffffu, v10, 1);
  float w01 = __shfl_down_sync(0xffffffffu, v01, 1), w11 = __shfl_down_sync(0xffffffffu, v11, 1);
  const bool valid = in_plane && x1 && y1 && (z + 1 < P.dz);
  if (valid && (threadIdx.x & 31) == 31) {  // the last lane has no neighbour
    w00 = __ldg(plane0 + j + 1); w10 = __ldg(plane0 + yz + j + 1);
    w01 = __ldg(plane0 + j + P.dz + 1); w11 = __ldg(plane0 + yz + j + P.dz + 1);
  }
  if (!valid) return 0;
  const float L = P.level;
  return (v00 < L ? 1 : 0) | (v10 < L ? 2 : 0) | (v01 < L ? 4 : 0) | (v11 < L ? 8 : 0) |
         (w00 < L ? 16 : 0) | (w10 < L ? 32 : 0) | (w01 < L ? 64 : 0) | (w11 < L ? 128 : 0);
}

__global__ void __launch_bounds__(kThreads)
k_mesh_count(const float* __restrict__ tsdf, const MeshParams P, int chunks_per_plane, int* __restrict__ chunk_count,
             unsigned char* __restrict__ cases) {
  __shared__ int s_sum[kThreads / 32];
  const int x = blockIdx.y, yz = P.dy * P.dz;
  const float* plane0 = tsdf + (size_t)x * yz;
  unsigned char* cases0 = cases + (size_t)x * yz;
  int mine = 0;
  for (int s = 0; s < kSlabs; ++s) {
    const int j = blockIdx.x * kChunk + s * kThreads + threadIdx.x;
    const bool in_plane = j < yz;
    const int m = cube_case(plane0, P, x, in_plane ? j : 0, in_plane);
    if (in_plane) {
      cases0[j] = (unsigned char)m;
      mine += c_tri_count[m];
    }
    if ((blockIdx.x * kChunk + (s + 1) * kThreads) >= yz) break;  // uniform: the plane ends inside this chunk
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += s_sum[w];
    chunk_count[x * chunks_per_plane + blockIdx.x] = t;
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

// Second sweep: reads the case byte of every cube (1 B instead of 8 floats), gives every cube its output slot with
// a CTA-wide scan per 256-cube slab, then EXPANDS: triangle t of the slab is produced by thread t % 256 (binary
// search of t in the slab's offsets), so all lanes work and consecutive lanes write consecutive triangles.
__global__ void __launch_bounds__(kThreads)
k_mesh_emit(const float* __restrict__ tsdf, const float* __restrict__ color_vol, const float* __restrict__ rem_vol,
            const MeshParams P, int chunks_per_plane, const unsigned char* __restrict__ cases,
            const long long* __restrict__ chunk_offset, long long capacity, float* __restrict__ verts,
            int* __restrict__ faces, float* __restrict__ norms, unsigned char* __restrict__ colors,
            float* __restrict__ rem_out) {
  __shared__ int s_warp[kThreads / 32];
  __shared__ int s_off[kThreads + 1];       // exclusive triangle offsets of the slab's cubes
  __shared__ unsigned char s_case[kThreads];
  __shared__ long long s_run;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int x = blockIdx.y, yz = P.dy * P.dz;
  const float* plane0 = tsdf + (size_t)x * yz;
  const unsigned char* cases0 = cases + (size_t)x * yz;
  if (tid == 0) s_run = chunk_offset[x * chunks_per_plane + blockIdx.x];
  for (int s = 0; s < kSlabs; ++s) {
    const int j0 = blockIdx.x * kChunk + s * kThreads;
    if (j0 >= yz) break;
    const int j = j0 + tid;
    const int m = j < yz ? cases0[j] : 0;
    const int cnt = c_tri_count[m];
    if (__syncthreads_or(cnt) == 0) continue;  // nothing to emit in this slab (also orders s_run / s_off reuse)
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    s_case[tid] = (unsigned char)m;
    __syncthreads();
    int before = 0, slab_total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      const int t = s_warp[w];
      if (w < wid) before += t;
      slab_total += t;
    }
    s_off[tid] = before + incl - cnt;
    if (tid == 0) s_off[kThreads] = slab_total;
    const long long run = s_run;
    __syncthreads();
    if (tid == 0) s_run = run + slab_total;
    for (int t = tid; t < slab_total; t += kThreads) {
      const long long tri = run + t;
      if (tri >= capacity) break;
      // source cube: the last slot whose offset is <= t
      int lo = 0, hi = kThreads - 1;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_off[mid] <= t) lo = mid; else hi = mid - 1;
      }
      const int src = lo, local = t - s_off[src], mc = s_case[src];
      const int jj = j0 + src, y = jj / P.dz, z = jj - y * P.dz;
      float pw[3][3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = c_tri_table[mc][3 * local + k];
        const int ca = c_edge_corners[e][0], cb = c_edge_corners[e][1];
        // vertex on the edge ca -> cb (cb = ca + one axis step), float32 like skimage's output
        const float va = __ldg(plane0 + (size_t)(ca & 1) * yz + jj + ((ca >> 1) & 1) * P.dz + ((ca >> 2) & 1));
        const float vb = __ldg(plane0 + (size_t)(cb & 1) * yz + jj + ((cb >> 1) & 1) * P.dz + ((cb >> 2) & 1));
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
      if (norms) {  // flat normal of the triangle for all three vertices (only consumed by meshwrite)
        const float ax = pw[1][0] - pw[0][0], ay = pw[1][1] - pw[0][1], az = pw[1][2] - pw[0][2];
        const float bx = pw[2][0] - pw[0][0], by = pw[2][1] - pw[0][1], bz = pw[2][2] - pw[0][2];
        float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
        const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
        if (nn > 0.f) { nx /= nn; ny /= nn; nz /= nn; }
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

}  // namespace

static int mesh_chunks_per_plane(int dy, int dz) { return (int)(((long long)dy * dz + kChunk - 1) / kChunk); }

// workspace: [chunk counts i32][chunk offsets i64][case byte per voxel]
static size_t mesh_ws_layout(int dx, int dy, int dz, size_t* off_offsets, size_t* off_cases) {
  const size_t n_chunks = (size_t)dx * mesh_chunks_per_plane(dy, dz);
  size_t off = vl_align256(n_chunks * 4);
  if (off_offsets) *off_offsets = off;
  off = vl_align256(off + n_chunks * 8);
  if (off_cases) *off_cases = off;
  off = vl_align256(off + (size_t)dx * dy * dz);
  return off;
}

extern "C" size_t vl_mesh_workspace_bytes(int dx, int dy, int dz) {
  if (dx <= 0 || dy <= 0 || dz <= 0) return 256;
  return mesh_ws_layout(dx, dy, dz, nullptr, nullptr);
}

static int mesh_args(const char* who, const float* d_tsdf, int dx, int dy, int dz, const void* d_ws, size_t ws_bytes,
                     MeshParams* P, float level, float voxel_size, const float* origin) {
  if (!d_tsdf || dx <= 0 || dy <= 0 || dz <= 0 || !d_ws || (((uintptr_t)d_ws) & 255) || dx > 65535 ||
      (long long)dx * dy * dz >= (1ll << 31)) {
    vl_set_error("%s: invalid argument (dim %d x %d x %d, workspace %p)", who, dx, dy, dz, d_ws);
    return VL_EINVAL;
  }
  if (ws_bytes < vl_mesh_workspace_bytes(dx, dy, dz)) {
    vl_set_error("%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, vl_mesh_workspace_bytes(dx, dy, dz));
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
  const int cpp = mesh_chunks_per_plane(dy, dz);
  size_t off_offsets, off_cases;
  mesh_ws_layout(dx, dy, dz, &off_offsets, &off_cases);
  char* ws = static_cast<char*>(d_workspace);
  { VlProfScope ps(VL_ST_MESH_COUNT, stream);
  k_mesh_count<<<dim3(cpp, dx), kThreads, 0, stream>>>(d_tsdf, P, cpp, reinterpret_cast<int*>(ws),
                                                      reinterpret_cast<unsigned char*>(ws + off_cases)); }
  VL_LAUNCH_CHECK("k_mesh_count");
  { VlProfScope ps(VL_ST_MESH_SCAN, stream);
  k_mesh_scan<<<1, 1024, 0, stream>>>(reinterpret_cast<const int*>(ws), reinterpret_cast<long long*>(ws + off_offsets),
                                      cpp * dx, d_total); }
  VL_LAUNCH_CHECK("k_mesh_scan");
  return VL_OK;
}

extern "C" int vl_mesh_emit(const float* d_tsdf, const float* d_color, const float* d_rem, int dx, int dy, int dz,
                            float level, float voxel_size, const float vol_origin[3], const void* d_workspace,
                            size_t workspace_bytes, long long capacity_tris, float* d_verts, int* d_faces,
                            float* d_norms, unsigned char* d_colors, float* d_rem_out, vl_stream stream_) {
  MeshParams P;
  int rc = mesh_args("vl_mesh_emit", d_tsdf, dx, dy, dz, d_workspace, workspace_bytes, &P, level, voxel_size, vol_origin);
  if (rc) return rc;
  if (!d_color || !d_rem || !vol_origin || capacity_tris < 0 ||
      (capacity_tris > 0 && (!d_verts || !d_faces || !d_colors || !d_rem_out))) {
    vl_set_error("vl_mesh_emit: invalid argument");
    return VL_EINVAL;
  }
  if (capacity_tris == 0) return VL_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int cpp = mesh_chunks_per_plane(dy, dz);
  size_t off_offsets, off_cases;
  mesh_ws_layout(dx, dy, dz, &off_offsets, &off_cases);
  const char* ws = static_cast<const char*>(d_workspace);
  VlProfScope ps(VL_ST_MESH_EMIT, stream);
  k_mesh_emit<<<dim3(cpp, dx), kThreads, 0, stream>>>(d_tsdf, d_color, d_rem, P, cpp,
                                                     reinterpret_cast<const unsigned char*>(ws + off_cases),
                                                     reinterpret_cast<const long long*>(ws + off_offsets), capacity_tris,
                                                     d_verts, d_faces, d_norms, d_colors, d_rem_out);
  VL_LAUNCH_CHECK("k_mesh_emit");
  return VL_OK;
}
