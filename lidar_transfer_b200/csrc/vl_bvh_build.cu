// vl_bvh_build.cu -- (i) LBVH build over the per-scan triangle mesh, sm_100a.
//
// Replaces the reference's serial top-down build: Triangle construction
// (auxiliary/raytracer/RayTracer.cpp:32-51) + BVH::build (auxiliary/raytracer/BVH.cpp:143-243).
// The tree is a different (Morton-order) tree; closest-hit results do not depend on the tree
// (DESIGN.md "parity under a different tree").
//
// Pipeline (all on the caller's stream, no allocation, no host sync):
//   k_init_header  -> k_bounds (vertex AABB, warp-reduce + ordered-uint atomics)
//   -> k_morton (validate faces, centroid -> 32-bit cubic-cell Morton key, zero climb flags)
//   -> 4 x { k_sort_hist, k_sort_scan, k_sort_scatter }   stable 8-bit LSD radix sort (key, face id)
//   -> k_emit_climb (write 48 B triangle records in sorted order, then Apetrei-style bottom-up
//      hierarchy + box refit in the same kernel; sub-trees of <= 4 triangles collapse to leaves)
#include "vl_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void k_init_header(VlHeader* hdr, int n_tris) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    hdr->n_tris = n_tris;
    hdr->root_ref = vl_make_leaf(0, 0);
    hdr->n_bad_faces = 0;
    hdr->max_climb = 0;
    for (int k = 0; k < 3; ++k) {
      hdr->bounds_min[k] = 0xffffffffu;
      hdr->bounds_max[k] = 0u;
    }
  }
}

// Vertex AABB: grid-stride float loads, warp shuffle reduction, one atomic pair per warp.
__global__ void __launch_bounds__(kThreads) k_bounds(const float* __restrict__ verts, int n_verts, VlHeader* hdr) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_verts; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = __ldg(verts + 3 * (size_t)i + k);
      mn[k] = fminf(mn[k], v);
      mx[k] = fmaxf(mx[k], v);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], off));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], off));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (mn[k] <= mx[k]) {
        atomicMin(&hdr->bounds_min[k], vl_float_to_ordered(mn[k]));
        atomicMax(&hdr->bounds_max[k], vl_float_to_ordered(mx[k]));
      }
    }
  }
}

__device__ __forceinline__ unsigned int expand10(unsigned int v) {  // 10 bits -> every third bit
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

// 32-bit Morton key over CUBIC cells: x,y get 11 bits, z 10 bits, one common scale
// s = 2048 / max(ext_x, ext_y, 2 ext_z) -- LiDAR scenes are flat, equal-size cells keep the
// radix tree's implicit splits isotropic.
__global__ void __launch_bounds__(kThreads)
k_morton(const float* __restrict__ verts, const int* __restrict__ faces, int n_verts, int n_faces,
         VlHeader* hdr, unsigned int* __restrict__ keys, int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_faces) return;
  flags[i] = 0;
  int i0 = __ldg(faces + 3 * (size_t)i), i1 = __ldg(faces + 3 * (size_t)i + 1), i2 = __ldg(faces + 3 * (size_t)i + 2);
  if ((unsigned)i0 >= (unsigned)n_verts || (unsigned)i1 >= (unsigned)n_verts || (unsigned)i2 >= (unsigned)n_verts) {
    atomicAdd(&hdr->n_bad_faces, 1);
    keys[i] = 0xffffffffu;
    return;
  }
  float bmin[3], ext[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    bmin[k] = vl_ordered_to_float(hdr->bounds_min[k]);
    ext[k] = vl_ordered_to_float(hdr->bounds_max[k]) - bmin[k];
  }
  float m = fmaxf(fmaxf(ext[0], ext[1]), 2.0f * ext[2]);
  float s = m > 0.0f ? 2048.0f / m : 0.0f;
  float c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    c[k] = (__ldg(verts + 3 * (size_t)i0 + k) + __ldg(verts + 3 * (size_t)i1 + k) + __ldg(verts + 3 * (size_t)i2 + k)) *
           (1.0f / 3.0f);
  unsigned int qx = (unsigned int)fminf(fmaxf((c[0] - bmin[0]) * s, 0.0f), 2047.0f);
  unsigned int qy = (unsigned int)fminf(fmaxf((c[1] - bmin[1]) * s, 0.0f), 2047.0f);
  unsigned int qz = (unsigned int)fminf(fmaxf((c[2] - bmin[2]) * s, 0.0f), 1023.0f);
  unsigned int key = ((qx >> 10) << 31) | ((qy >> 10) << 30) | (expand10(qz) << 2) | (expand10(qx & 1023u) << 1) |
                     expand10(qy & 1023u);
  keys[i] = key;
}

// ---------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits, tiles of VL_SORT_TILE keys
// hist layout: hist[digit * n_tiles + tile]  (an exclusive scan of the flat array yields
// the global output offset of every (digit, tile) bucket)
// ---------------------------------------------------------------------------
constexpr int kItems = VL_SORT_TILE / kThreads;  // 16

__global__ void __launch_bounds__(kThreads)
k_sort_hist(const unsigned int* __restrict__ keys, int n, int shift, unsigned int* __restrict__ hist, int n_tiles) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int base = blockIdx.x * VL_SORT_TILE;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    int idx = base + j * kThreads + threadIdx.x;
    if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// single-CTA exclusive scan over `total` counters (<= a few hundred thousand)
__global__ void __launch_bounds__(1024) k_sort_scan(unsigned int* __restrict__ hist, int total) {
  __shared__ unsigned int warp_sums[32];
  __shared__ unsigned int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  constexpr int kPer = 4;
  for (int base = 0; base < total; base += 1024 * kPer) {
    unsigned int v[kPer], sum = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      int idx = base + tid * kPer + k;
      v[k] = idx < total ? hist[idx] : 0u;
      sum += v[k];
    }
    unsigned int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      unsigned int ws = warp_sums[lane], wi = ws;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, wi, off);
        if (lane >= off) wi += t;
      }
      warp_sums[lane] = wi - ws;  // exclusive
    }
    __syncthreads();
    unsigned int run = carry_s + warp_sums[wid] + (incl - sum);
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      int idx = base + tid * kPer + k;
      if (idx < total) hist[idx] = run;
      run += v[k];
    }
    __syncthreads();
    if (tid == 1023) carry_s = run;
    __syncthreads();
  }
}

// Stable scatter.  Warp w owns the contiguous 512-key run [tile*4096 + 512 w, +512) and walks it in
// 16 rounds of 32; __match_any_sync ranks equal digits inside a round, a per-warp digit counter in
// shared memory carries the rank across rounds, a 256-thread column scan orders the warps.
__global__ void __launch_bounds__(kThreads)
k_sort_scatter(const unsigned int* __restrict__ keys_in, const unsigned int* __restrict__ vals_in,
               unsigned int* __restrict__ keys_out, unsigned int* __restrict__ vals_out, int n, int shift,
               const unsigned int* __restrict__ hist, int n_tiles) {
  __shared__ unsigned int cnt[kThreads / 32][256];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int k = tid; k < (kThreads / 32) * 256; k += kThreads) (&cnt[0][0])[k] = 0;
  __syncthreads();
  const int base = blockIdx.x * VL_SORT_TILE + w * (32 * kItems);
  const unsigned int lt_mask = (1u << lane) - 1u;
  unsigned int key[kItems], val[kItems];
  unsigned short rank[kItems];
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    int idx = base + j * 32 + lane;
    bool valid = idx < n;
    key[j] = valid ? keys_in[idx] : 0xffffffffu;
    val[j] = valid ? (vals_in ? vals_in[idx] : (unsigned int)idx) : 0u;
    unsigned int d = (key[j] >> shift) & 255u;
    unsigned int peers = __match_any_sync(0xffffffffu, valid ? d : 0x100u);
    int leader = __ffs(peers) - 1;
    unsigned int basecnt = 0;
    if (lane == leader && valid) {
      basecnt = cnt[w][d];
      cnt[w][d] = basecnt + __popc(peers);
    }
    basecnt = __shfl_sync(0xffffffffu, basecnt, leader);
    rank[j] = (unsigned short)(basecnt + __popc(peers & lt_mask));
    __syncwarp();
  }
  __syncthreads();
  {  // thread d: exclusive scan of digit d over the 8 warps, seeded with the global bucket offset
    unsigned int run = hist[tid * n_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kThreads / 32; ++ww) {
      unsigned int c = cnt[ww][tid];
      cnt[ww][tid] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    int idx = base + j * 32 + lane;
    if (idx < n) {
      unsigned int d = (key[j] >> shift) & 255u;
      unsigned int out = cnt[w][d] + rank[j];
      keys_out[out] = key[j];
      vals_out[out] = val[j];
    }
  }
}

// ---------------------------------------------------------------------------
// emit sorted triangle records + bottom-up hierarchy (Apetrei 2014 formulation of the
// Karras radix tree: inner node i sits between sorted keys i and i+1)
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long key_delta(const unsigned int* __restrict__ keys, int i) {
  return ((unsigned long long)(keys[i] ^ keys[i + 1]) << 32) | (unsigned int)(i ^ (i + 1));
}

__global__ void __launch_bounds__(kThreads)
k_emit_climb(const float* __restrict__ verts, const int* __restrict__ faces, const int* __restrict__ colors,
             const float* __restrict__ rem, int n_verts, int n, const unsigned int* __restrict__ keys,
             const unsigned int* __restrict__ vals, VlHeader* hdr, VlNode* nodes, VlTri* __restrict__ tris,
             int4* __restrict__ c0, int* flags) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int f = (int)vals[p];
  int i0 = __ldg(faces + 3 * (size_t)f), i1 = __ldg(faces + 3 * (size_t)f + 1), i2 = __ldg(faces + 3 * (size_t)f + 2);
  float bmin[3], bmax[3];
  if ((unsigned)i0 >= (unsigned)n_verts || (unsigned)i1 >= (unsigned)n_verts || (unsigned)i2 >= (unsigned)n_verts) {
    // invalid face: a record no ray can hit (a = 0) and an empty box
    VlTri t;
    t.v0 = make_float4(0.f, 0.f, 0.f, __int_as_float(f));
    t.e1 = make_float4(0.f, 0.f, 0.f, 0.f);
    t.e2 = make_float4(0.f, 0.f, 0.f, 0.f);
    tris[p] = t;
    c0[p] = make_int4(0, 0, 0, f);
#pragma unroll
    for (int k = 0; k < 3; ++k) { bmin[k] = INFINITY; bmax[k] = -INFINITY; }
  } else {
    float v0[3], v1[3], v2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      v0[k] = __ldg(verts + 3 * (size_t)i0 + k);
      v1[k] = __ldg(verts + 3 * (size_t)i1 + k);
      v2[k] = __ldg(verts + 3 * (size_t)i2 + k);
    }
    // Triangle.h:63-70 mean remission, RayTracer.cpp:36-48 colours pass through float
    float r = __fdiv_rn(__fadd_rn(__fadd_rn(__ldg(rem + i0), __ldg(rem + i1)), __ldg(rem + i2)), 3.0f);
    VlTri t;
    t.v0 = make_float4(v0[0], v0[1], v0[2], __int_as_float(f));
    t.e1 = make_float4(__fsub_rn(v1[0], v0[0]), __fsub_rn(v1[1], v0[1]), __fsub_rn(v1[2], v0[2]), r);
    t.e2 = make_float4(__fsub_rn(v2[0], v0[0]), __fsub_rn(v2[1], v0[1]), __fsub_rn(v2[2], v0[2]), 0.f);
    tris[p] = t;
    c0[p] = make_int4((int)(float)__ldg(colors + 3 * (size_t)i0), (int)(float)__ldg(colors + 3 * (size_t)i0 + 1),
                      (int)(float)__ldg(colors + 3 * (size_t)i0 + 2), f);
    // conservative leaf box: pad by 2^-21 of the scene's largest |coordinate| so that rounding in
    // the slab test can never cull a triangle the Moller-Trumbore arithmetic would accept
    float amax = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      amax = fmaxf(amax, fmaxf(fabsf(vl_ordered_to_float(hdr->bounds_min[k])), fabsf(vl_ordered_to_float(hdr->bounds_max[k]))));
    const float pad = amax * 4.76837158203125e-07f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      bmin[k] = fminf(v0[k], fminf(v1[k], v2[k])) - pad;
      bmax[k] = fmaxf(v0[k], fmaxf(v1[k], v2[k])) + pad;
    }
  }

  int l = p, r = p;
  int ref = vl_make_leaf(p, 1);
  if (n == 1) { hdr->root_ref = ref; return; }
  int climb = 0;
  while (true) {
    bool is_left;
    if (l == 0) is_left = true;
    else if (r == n - 1) is_left = false;
    else is_left = key_delta(keys, r) < key_delta(keys, l - 1);
    const int parent = is_left ? r : l - 1;
    float* nf = reinterpret_cast<float*>(&nodes[parent]);
    int* ni = reinterpret_cast<int*>(&nodes[parent]);
    const int boff = is_left ? 0 : 6;
#pragma unroll
    for (int k = 0; k < 3; ++k) { nf[boff + k] = bmin[k]; nf[boff + 3 + k] = bmax[k]; }
    ni[is_left ? 12 : 13] = ref;
    ni[is_left ? 14 : 15] = is_left ? l : r;
    __threadfence();
    if (atomicExch(&flags[parent], 1) == 0) break;  // first child to arrive stops here
    __threadfence();
    const int soff = is_left ? 6 : 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      bmin[k] = fminf(bmin[k], __ldcg(nf + soff + k));
      bmax[k] = fmaxf(bmax[k], __ldcg(nf + soff + 3 + k));
    }
    if (is_left) r = __ldcg(ni + 15); else l = __ldcg(ni + 14);
    const int size = r - l + 1;
    ref = size <= VL_LEAF_MAX ? vl_make_leaf(l, size) : parent;
    ++climb;
    if (l == 0 && r == n - 1) {
      hdr->root_ref = ref;
      atomicMax(&hdr->max_climb, climb);
      break;
    }
  }
}

}  // namespace

int vl_bvh_build_launch(const float* d_verts, const int* d_faces, const int* d_colors, const float* d_rem,
                        int n_verts, int n_faces, void* d_blob, cudaStream_t stream) {
  char* blob = static_cast<char*>(d_blob);
  VlBlobLayout L = vl_blob_layout(n_faces);
  VlHeader* hdr = reinterpret_cast<VlHeader*>(blob);
  k_init_header<<<1, 32, 0, stream>>>(hdr, n_faces);
  VL_LAUNCH_CHECK("k_init_header");
  if (n_faces <= 0) return VL_OK;
  unsigned int* keys0 = reinterpret_cast<unsigned int*>(blob + L.off_keys0);
  unsigned int* keys1 = reinterpret_cast<unsigned int*>(blob + L.off_keys1);
  unsigned int* vals0 = reinterpret_cast<unsigned int*>(blob + L.off_vals0);
  unsigned int* vals1 = reinterpret_cast<unsigned int*>(blob + L.off_vals1);
  unsigned int* hist = reinterpret_cast<unsigned int*>(blob + L.off_hist);
  int* flags = reinterpret_cast<int*>(blob + L.off_flags);

  int nb_verts = (n_verts + kThreads - 1) / kThreads;
  if (nb_verts > 148 * 8) nb_verts = 148 * 8;
  if (nb_verts < 1) nb_verts = 1;
  { VlProfScope ps(VL_ST_BOUNDS, stream);
  k_bounds<<<nb_verts, kThreads, 0, stream>>>(d_verts, n_verts, hdr); }
  VL_LAUNCH_CHECK("k_bounds");
  const int nb_faces = (n_faces + kThreads - 1) / kThreads;
  { VlProfScope ps(VL_ST_MORTON, stream);
  k_morton<<<nb_faces, kThreads, 0, stream>>>(d_verts, d_faces, n_verts, n_faces, hdr, keys0, flags); }
  VL_LAUNCH_CHECK("k_morton");

  const int nt = L.n_sort_tiles;
  const unsigned int* kin = keys0;
  const unsigned int* vin = nullptr;  // pass 0 generates the identity permutation on the fly
  unsigned int* kout = keys1;
  unsigned int* vout = vals1;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 8 * pass;
    { VlProfScope ps(VL_ST_SORT_HIST, stream);
    k_sort_hist<<<nt, kThreads, 0, stream>>>(kin, n_faces, shift, hist, nt); }
    VL_LAUNCH_CHECK("k_sort_hist");
    { VlProfScope ps(VL_ST_SORT_SCAN, stream);
    k_sort_scan<<<1, 1024, 0, stream>>>(hist, 256 * nt); }
    VL_LAUNCH_CHECK("k_sort_scan");
    { VlProfScope ps(VL_ST_SORT_SCATTER, stream);
    k_sort_scatter<<<nt, kThreads, 0, stream>>>(kin, vin, kout, vout, n_faces, shift, hist, nt); }
    VL_LAUNCH_CHECK("k_sort_scatter");
    kin = kout;
    vin = vout;
    kout = (kout == keys1) ? keys0 : keys1;
    vout = (vout == vals1) ? vals0 : vals1;
  }
  // after 4 passes the sorted (key, face id) pairs are back in keys0 / vals0
  VlProfScope ps_emit(VL_ST_EMIT_CLIMB, stream);
  k_emit_climb<<<nb_faces, kThreads, 0, stream>>>(d_verts, d_faces, d_colors, d_rem, n_verts, n_faces, keys0, vals0, hdr,
                                                 reinterpret_cast<VlNode*>(blob + L.off_nodes),
                                                 reinterpret_cast<VlTri*>(blob + L.off_tris),
                                                 reinterpret_cast<int4*>(blob + L.off_c0), flags);
  VL_LAUNCH_CHECK("k_emit_climb");
  return VL_OK;
}
