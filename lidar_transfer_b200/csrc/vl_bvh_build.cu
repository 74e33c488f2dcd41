// vl_bvh_build.cu -- (i) LBVH build over the per-scan triangle mesh, sm_100a.
//
// Replaces the reference's serial top-down build: Triangle construction
// (auxiliary/raytracer/RayTracer.cpp:32-51) + BVH::build (auxiliary/raytracer/BVH.cpp:143-243).
// The tree is a different (Morton-order) tree; closest-hit results do not depend on the tree
// (DESIGN.md "parity under a different tree").
//
// Pipeline: 8 launches on the caller's stream, no allocation, no host sync:
//   k_init_header  header + zeroed digit histograms / tile tickets
//   k_bounds       vertex AABB (block reduction, 6 ordered-uint atomics per CTA)
//   k_morton       validate faces, centroid -> 32-bit cubic-cell Morton key, the four 8-bit digit
//                  histograms of the whole key array (shared-memory privatised), zeroed climb flags
//                  and tile states
//   4 x k_sort_pass  stable LSD radix sort of (key, face id), ONE kernel per 8-bit digit: tiles take
//                  tickets, rank their keys with __match_any_sync and obtain their global offsets by
//                  decoupled look-back over per-tile digit counts (single-pass "onesweep" scheme) --
//                  each key/value is read once and written once per pass
//   k_emit_climb   sorted 48 B triangle records + bottom-up hierarchy (Apetrei's formulation of the
//                  Karras radix tree) with box refit in the same kernel.  A CTA owns a run of sorted
//                  triangles and resolves every inner node whose two children lie inside that run
//                  through SHARED memory (flag / sibling box / sibling ref); only the few nodes that
//                  straddle CTA boundaries climb on through global memory + atomics.  Sub-trees of
//                  <= 4 triangles collapse to leaves and never touch the node array.
#include "vl_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(1024) k_init_header(VlHeader* hdr, int n_tris, unsigned int* ghist,
                                                      unsigned int* tickets) {
  const int t = threadIdx.x;
  if (t < 256 * VL_SORT_PASSES) ghist[t] = 0u;
  if (t < VL_SORT_PASSES) tickets[t] = 0u;
  if (t == 0) {
    hdr->n_tris = n_tris;
    hdr->root_ref = vl_make_leaf(0, 0);
    hdr->n_bad_faces = 0;
    hdr->max_climb = 0;
    hdr->n_pending = 0;
    for (int k = 0; k < 3; ++k) {
      hdr->bounds_min[k] = 0xffffffffu;
      hdr->bounds_max[k] = 0u;
    }
  }
}

// Vertex AABB: grid-stride float4-free loads (12 B stride), warp shuffle + shared-memory block
// reduction, one atomic set per CTA.
__global__ void __launch_bounds__(kThreads) k_bounds(const float* __restrict__ verts, int n_verts, VlHeader* hdr) {
  __shared__ float s_mn[kThreads / 32][3], s_mx[kThreads / 32][3];
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  // the flat float array is read fully coalesced: element e belongs to axis e % 3
  const size_t total = 3 * (size_t)n_verts;
  const size_t stride = (size_t)gridDim.x * kThreads * 3;
  for (size_t base = ((size_t)blockIdx.x * kThreads + threadIdx.x) * 3; base < total; base += stride) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = __ldg(verts + base + k);
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
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { s_mn[w][k] = mn[k]; s_mx[w][k] = mx[k]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int k = threadIdx.x;
    float a = s_mn[0][k], b = s_mx[0][k];
#pragma unroll
    for (int ww = 1; ww < kThreads / 32; ++ww) { a = fminf(a, s_mn[ww][k]); b = fmaxf(b, s_mx[ww][k]); }
    if (a <= b) {
      atomicMin(&hdr->bounds_min[k], vl_float_to_ordered(a));
      atomicMax(&hdr->bounds_max[k], vl_float_to_ordered(b));
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
// radix tree's implicit splits isotropic.  Persistent grid: every CTA also accumulates the four
// digit histograms of its keys in shared memory and flushes the non-zero bins once.
__global__ void __launch_bounds__(kThreads)
k_morton(const float* __restrict__ verts, const int* __restrict__ faces, int n_verts, int n_faces,
         VlHeader* hdr, unsigned int* __restrict__ keys, int* __restrict__ flags, unsigned int* __restrict__ ghist,
         unsigned int* __restrict__ tile_state, int n_state_words) {
  __shared__ unsigned int h[VL_SORT_PASSES][256];
#pragma unroll
  for (int p = 0; p < VL_SORT_PASSES; ++p) h[p][threadIdx.x] = 0u;
  float bmin[3], ext[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    bmin[k] = vl_ordered_to_float(hdr->bounds_min[k]);
    ext[k] = vl_ordered_to_float(hdr->bounds_max[k]) - bmin[k];
  }
  const float m = fmaxf(fmaxf(ext[0], ext[1]), 2.0f * ext[2]);
  const float s = m > 0.0f ? 2048.0f / m : 0.0f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // conservative leaf-box padding: 2^-21 of the scene's largest |coordinate|, so that rounding in the slab
    // test can never cull a triangle the Moller-Trumbore arithmetic would accept
    float amax = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) amax = fmaxf(amax, fmaxf(fabsf(bmin[k]), fabsf(bmin[k] + ext[k])));
    hdr->box_pad = amax * 4.76837158203125e-07f;
  }
  __syncthreads();
  const int stride = gridDim.x * kThreads;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n_state_words; i += stride) tile_state[i] = 0u;
  int n_bad = 0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n_faces; i += stride) {
    flags[i] = 0;
    const int i0 = __ldg(faces + 3 * (size_t)i), i1 = __ldg(faces + 3 * (size_t)i + 1), i2 = __ldg(faces + 3 * (size_t)i + 2);
    unsigned int key;
    if ((unsigned)i0 >= (unsigned)n_verts || (unsigned)i1 >= (unsigned)n_verts || (unsigned)i2 >= (unsigned)n_verts) {
      ++n_bad;
      key = 0xffffffffu;
    } else {
      float c[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        c[k] = (__ldg(verts + 3 * (size_t)i0 + k) + __ldg(verts + 3 * (size_t)i1 + k) + __ldg(verts + 3 * (size_t)i2 + k)) *
               (1.0f / 3.0f);
      const unsigned int qx = (unsigned int)fminf(fmaxf((c[0] - bmin[0]) * s, 0.0f), 2047.0f);
      const unsigned int qy = (unsigned int)fminf(fmaxf((c[1] - bmin[1]) * s, 0.0f), 2047.0f);
      const unsigned int qz = (unsigned int)fminf(fmaxf((c[2] - bmin[2]) * s, 0.0f), 1023.0f);
      key = ((qx >> 10) << 31) | ((qy >> 10) << 30) | (expand10(qz) << 2) | (expand10(qx & 1023u) << 1) |
            expand10(qy & 1023u);
    }
    keys[i] = key;
#pragma unroll
    for (int p = 0; p < VL_SORT_PASSES; ++p) atomicAdd(&h[p][(key >> (8 * p)) & 255u], 1u);
  }
  if (n_bad) atomicAdd(&hdr->n_bad_faces, n_bad);
  __syncthreads();
#pragma unroll
  for (int p = 0; p < VL_SORT_PASSES; ++p) {
    const unsigned int c = h[p][threadIdx.x];
    if (c) atomicAdd(&ghist[p * 256 + threadIdx.x], c);
  }
}

// ---------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits, one kernel per digit (decoupled look-back)
//
// tile state word (per tile, per digit): [31:30] 0 = not ready, 1 = this tile's count,
// 2 = inclusive count of tiles 0..this;  [29:0] the count (n < 2^28).
// Tiles are numbered by a ticket counter, so a tile only ever waits for tiles that started earlier.
// ---------------------------------------------------------------------------
constexpr int kItems = VL_SORT_ITEMS;
constexpr int kSortWarps = VL_SORT_THREADS / 32;
constexpr unsigned int kStAggregate = 1u << 30, kStPrefix = 2u << 30, kStMask = (1u << 30) - 1u;

__global__ void __launch_bounds__(VL_SORT_THREADS)
k_sort_pass(const unsigned int* __restrict__ keys_in, const unsigned int* __restrict__ vals_in,
            unsigned int* __restrict__ keys_out, unsigned int* __restrict__ vals_out, int n, int shift,
            const unsigned int* __restrict__ ghist, volatile unsigned int* tile_state, unsigned int* ticket, int n_tiles) {
  __shared__ __align__(16) unsigned int cnt[kSortWarps][256];
  __shared__ unsigned int warp_sums[8];
  __shared__ unsigned int s_texcl[256], s_gbase[256];
  __shared__ unsigned int s_keys[VL_SORT_TILE], s_vals[VL_SORT_TILE];
  __shared__ int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_tile = (int)atomicAdd(ticket, 1u);
  for (int k = tid; k < kSortWarps * 64; k += VL_SORT_THREADS) reinterpret_cast<uint4*>(&cnt[0][0])[k] = make_uint4(0u, 0u, 0u, 0u);
  // exclusive scan of the global digit histogram: thread d (< 256) -> first output slot of digit d
  unsigned int gcount = 0, incl = 0;
  if (tid < 256) {
    gcount = ghist[tid];
    incl = gcount;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
  }
  __syncthreads();
  unsigned int gbase = incl - gcount;
  if (tid < 256) {
#pragma unroll
    for (int ww = 0; ww < 8; ++ww)
      if (ww < w) gbase += warp_sums[ww];
  }
  const int tile = s_tile;

  // rank: warp w owns the contiguous run [tile*TILE + w*32*kItems, +32*kItems), walked in kItems rounds of 32
  const int base = tile * VL_SORT_TILE + w * (32 * kItems);
  const unsigned int lt_mask = (1u << lane) - 1u;
  unsigned int key[kItems], val[kItems];
  unsigned short rank[kItems];
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int idx = base + j * 32 + lane;
    const bool valid = idx < n;
    key[j] = valid ? keys_in[idx] : 0xffffffffu;
    val[j] = valid ? (vals_in ? vals_in[idx] : (unsigned int)idx) : 0u;
  }
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const bool valid = base + j * 32 + lane < n;
    const unsigned int d = (key[j] >> shift) & 255u;
    const unsigned int peers = __match_any_sync(0xffffffffu, valid ? d : 0x100u);
    const int leader = __ffs(peers) - 1;
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
  if (tid < 256) {  // thread d: digit d's count in this tile, published; look back for the tiles before
    unsigned int run = 0;
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ++ww) {
      const unsigned int c = cnt[ww][tid];
      cnt[ww][tid] = run;
      run += c;
    }
    // Two-level decoupled look-back.  Every tile publishes its count; the count of all tiles before tile t is
    //   (inclusive count of the previous GROUP of kGroup tiles) + (counts of the earlier tiles of t's own group).
    // The last tile of a group publishes the group's aggregate, looks back over GROUPS (a few steps, decoupled:
    // aggregate or inclusive, whichever is there) and publishes the group's inclusive count.  With every tile
    // resident at once a flat look-back walks back ~t/2 tiles; this one reads <= kGroup - 1 tile words + 1 group word.
    constexpr int kGroup = 32, kBatch = 8;
    volatile unsigned int* group_state = tile_state + (size_t)n_tiles * 256;
    tile_state[(size_t)tile * 256 + tid] = kStAggregate | run;
    const int g = tile / kGroup, first = g * kGroup;
    unsigned int before = 0;
    for (int t = tile - 1; t >= first;) {  // the earlier tiles of this group, kBatch independent loads per round trip
      unsigned int st[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) st[k] = (t - k >= first) ? tile_state[(size_t)(t - k) * 256 + tid] : (1u << 30);
      bool ready = true;
#pragma unroll
      for (int k = 0; k < kBatch; ++k) ready = ready && (st[k] >> 30) != 0u;
      if (!ready) { __nanosleep(32); continue; }
#pragma unroll
      for (int k = 0; k < kBatch; ++k) before += st[k] & kStMask;
      t -= kBatch;
    }
    if ((tile % kGroup) == kGroup - 1) {  // this tile completes its group
      const unsigned int group_total = before + run;
      unsigned int groups_before = 0;
      if (g == 0) {
        group_state[tid] = kStPrefix | group_total;
      } else {
        group_state[(size_t)g * 256 + tid] = kStAggregate | group_total;
        int t = g - 1;
        while (true) {  // kBatch earlier groups per round trip; a virtual group -1 holds the inclusive count 0
          unsigned int st[kBatch];
#pragma unroll
          for (int k = 0; k < kBatch; ++k) st[k] = (t - k >= 0) ? group_state[(size_t)(t - k) * 256 + tid] : (2u << 30);
          bool done = false;
          int used = 0;
#pragma unroll
          for (int k = 0; k < kBatch; ++k) {
            if (!done && used == k && (st[k] >> 30) != 0u) {
              groups_before += st[k] & kStMask;
              ++used;
              done = (st[k] >> 30) == 2u;
            }
          }
          if (done) break;
          t -= used;
          if (used == 0) __nanosleep(32);
        }
        group_state[(size_t)g * 256 + tid] = kStPrefix | (groups_before + group_total);
      }
      before += groups_before;
    } else if (g > 0) {  // inclusive count of the previous group
      unsigned int st;
      while (((st = group_state[(size_t)(g - 1) * 256 + tid]) >> 30) != 2u) __nanosleep(32);
      before += st & kStMask;
    }
    // position of digit d inside the tile's locally sorted order (exclusive scan of `run` over the 256 digits)
    unsigned int tincl = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, tincl, off);
      if (lane >= off) tincl += t;
    }
    if (lane == 31) warp_sums[w] = tincl;
    s_texcl[tid] = tincl - run;       // + the sums of the digit warps before, added after the barrier
    s_gbase[tid] = gbase + before;    // first global slot of this tile's digit-d keys
  }
  __syncthreads();
  if (tid < 256) {
    unsigned int add = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww)
      if (ww < w) add += warp_sums[ww];
    const unsigned int texcl = s_texcl[tid] + add;
    s_gbase[tid] -= texcl;            // global slot = s_gbase[d] + position in the tile's sorted order
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ++ww) cnt[ww][tid] += texcl;
  }
  __syncthreads();
  // stage the tile in digit order in shared memory, then write runs of equal digits with consecutive lanes:
  // neighbouring lanes hit neighbouring addresses instead of 32 different sectors
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int idx = base + j * 32 + lane;
    if (idx < n) {
      const unsigned int d = (key[j] >> shift) & 255u;
      const unsigned int lpos = cnt[w][d] + rank[j];
      s_keys[lpos] = key[j];
      s_vals[lpos] = val[j];
    }
  }
  __syncthreads();
  const int tile_n = min(VL_SORT_TILE, n - tile * VL_SORT_TILE);
  for (int i = tid; i < tile_n; i += VL_SORT_THREADS) {
    const unsigned int k = s_keys[i];
    const unsigned int out = s_gbase[(k >> shift) & 255u] + (unsigned int)i;
    keys_out[out] = k;
    vals_out[out] = s_vals[i];
  }
}

// ---------------------------------------------------------------------------
// emit sorted triangle records + bottom-up hierarchy (Apetrei 2014 formulation of the
// Karras radix tree: inner node i sits between sorted keys i and i+1)
// ---------------------------------------------------------------------------
constexpr int kClimbThreads = 512;
constexpr int kClimbWarps = kClimbThreads / 32;
constexpr unsigned long long kDeltaInf = ~0ull;

// delta(i): the split level between sorted keys i and i+1 (larger = the two keys part ways higher up the
// radix tree); equal keys are told apart by the index bits, so all deltas of one array are distinct.
__device__ __forceinline__ unsigned long long key_delta(const unsigned int* __restrict__ keys, int i) {
  return ((unsigned long long)(__ldg(keys + i) ^ __ldg(keys + i + 1)) << 32) | (unsigned int)(i ^ (i + 1));
}

__device__ __forceinline__ unsigned long long shfl_up64(unsigned long long v, int off) {
  return ((unsigned long long)__shfl_up_sync(0xffffffffu, (unsigned int)(v >> 32), off) << 32) |
         __shfl_up_sync(0xffffffffu, (unsigned int)v, off);
}
__device__ __forceinline__ unsigned long long shfl_down64(unsigned long long v, int off) {
  return ((unsigned long long)__shfl_down_sync(0xffffffffu, (unsigned int)(v >> 32), off) << 32) |
         __shfl_down_sync(0xffffffffu, (unsigned int)v, off);
}

struct __align__(16) ClimbSlot {
  float4 a;  // box min xyz, box max x
  float4 b;  // box max y z, child ref (bits), outer range bound (bits)
};

__global__ void __launch_bounds__(kClimbThreads)
k_emit_climb(const float* __restrict__ verts, const int* __restrict__ faces, const int* __restrict__ colors,
             const float* __restrict__ rem, int n_verts, int n, const unsigned int* __restrict__ keys,
             const unsigned int* __restrict__ vals, VlHeader* hdr, VlNode* nodes, VlTri* __restrict__ tris,
             int4* __restrict__ c0, int* flags, unsigned int* __restrict__ pending) {
  // Slot k of the CTA = inner node B0 + k (it sits between sorted keys B0+k and B0+k+1).  Node i is LOCAL
  // when its whole key range lies inside the CTA's run [B0, B1): its range grows left / right until it
  // meets a larger delta, so   local(i)  <=>  max delta[B0-1 .. i-1] > delta(i)  and  max delta[i+1 .. B1-1] > delta(i)
  // (delta(-1) = delta(n-1) = +inf).  Both children of a node evaluate the same predicate, and every leaf
  // under a local node belongs to this CTA, so the two children always meet in the same place.
  //
  // Phases: (0) deltas + locality flags; (1) one thread per triangle emits its record and parks its box;
  // (2) every thread finds, from the deltas alone, the maximal local sub-tree of <= 4 triangles it belongs
  // to -- those become the leaves of the BVH; (3) the first thread of each leaf is compacted to the front of
  // the CTA (dense warps) and climbs from there.
  __shared__ unsigned long long s_delta[kClimbThreads + 1];  // [0] = delta(B0-1), [1+k] = delta(B0+k)
  __shared__ unsigned long long s_wmax[2][kClimbWarps];
  __shared__ unsigned char s_local[kClimbThreads];
  __shared__ int s_flag[kClimbThreads];
  __shared__ ClimbSlot s_slot[kClimbThreads][2];  // [left child | right child]; its head doubles as s_tbox until phase 3
  __shared__ int s_lead[kClimbThreads];           // (first - B0) | count << 16 of the compacted leaves
  __shared__ int s_wcount[kClimbWarps];
  __shared__ unsigned char s_info[kClimbThreads];
  float(*s_tbox)[6] = reinterpret_cast<float(*)[6]>(&s_slot[0][0]);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int B0 = blockIdx.x * kClimbThreads;
  const int B1 = min(B0 + kClimbThreads, n);
  const int p = B0 + tid;
  const bool active = p < n;
  s_flag[tid] = 0;
  // ---- phase 0 ----
  unsigned long long dl = 0ull;  // delta(p); 0 for slots past the run never wins a max
  if (active) dl = (p == n - 1) ? kDeltaInf : key_delta(keys, p);
  if (tid == 0) s_delta[0] = (B0 == 0) ? kDeltaInf : key_delta(keys, B0 - 1);
  s_delta[1 + tid] = dl;
  unsigned long long pm = dl, sm = dl;  // inclusive prefix / suffix max inside the warp
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long a = shfl_up64(pm, off), b = shfl_down64(sm, off);
    if (lane >= off) pm = max(pm, a);
    if (lane + off < 32) sm = max(sm, b);
  }
  if (lane == 31) s_wmax[0][wid] = pm;
  if (lane == 0) s_wmax[1][wid] = sm;
  __syncthreads();
  if (wid < 2) {  // warp 0: exclusive prefix max of the warp totals, warp 1: exclusive suffix max (0 is neutral)
    unsigned long long acc = lane < kClimbWarps ? s_wmax[wid][lane] : 0ull;
#pragma unroll
    for (int off = 1; off < kClimbWarps; off <<= 1) {
      const unsigned long long up = shfl_up64(acc, off), down = shfl_down64(acc, off);
      if (wid == 0) { if (lane >= off) acc = max(acc, up); }
      else if (lane + off < 32) acc = max(acc, down);
    }
    const unsigned long long up1 = shfl_up64(acc, 1), down1 = shfl_down64(acc, 1);
    unsigned long long ex = wid == 0 ? (lane == 0 ? 0ull : up1) : (lane == 31 ? 0ull : down1);
    __syncwarp();
    if (lane < kClimbWarps) s_wmax[wid][lane] = ex;
  }
  // ---- phase 1: triangle record + box ----
  float bmin[3], bmax[3];
  if (active) {
    const int f = (int)vals[p];
    const int i0 = __ldg(faces + 3 * (size_t)f), i1 = __ldg(faces + 3 * (size_t)f + 1), i2 = __ldg(faces + 3 * (size_t)f + 2);
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
      const float r = __fdiv_rn(__fadd_rn(__fadd_rn(__ldg(rem + i0), __ldg(rem + i1)), __ldg(rem + i2)), 3.0f);
      VlTri t;
      t.v0 = make_float4(v0[0], v0[1], v0[2], __int_as_float(f));
      t.e1 = make_float4(__fsub_rn(v1[0], v0[0]), __fsub_rn(v1[1], v0[1]), __fsub_rn(v1[2], v0[2]), r);
      t.e2 = make_float4(__fsub_rn(v2[0], v0[0]), __fsub_rn(v2[1], v0[1]), __fsub_rn(v2[2], v0[2]), 0.f);
      tris[p] = t;
      c0[p] = make_int4((int)(float)__ldg(colors + 3 * (size_t)i0), (int)(float)__ldg(colors + 3 * (size_t)i0 + 1),
                        (int)(float)__ldg(colors + 3 * (size_t)i0 + 2), f);
      const float pad = hdr->box_pad;  // set by k_morton
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        bmin[k] = fminf(v0[k], fminf(v1[k], v2[k])) - pad;
        bmax[k] = fmaxf(v0[k], fmaxf(v1[k], v2[k])) + pad;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { s_tbox[tid][k] = bmin[k]; s_tbox[tid][3 + k] = bmax[k]; }
  }
  __syncthreads();
  {  // locality flag of node p (exclusive maxima: [B0-1 .. p-1] and [p+1 .. B1-1])
    unsigned long long pre = shfl_up64(pm, 1), suf = shfl_down64(sm, 1);
    if (lane == 0) pre = 0ull;
    if (lane == 31) suf = 0ull;
    pre = max(max(pre, s_delta[0]), s_wmax[0][wid]);
    suf = max(suf, s_wmax[1][wid]);
    s_local[tid] = (active && p <= B1 - 2 && pre > dl && suf > dl) ? 1 : 0;
  }
  __syncthreads();
  // ---- phase 2: the leaf (maximal local sub-tree of <= VL_LEAF_MAX triangles) this triangle belongs to ----
  // Node i covers keys [i - Lx, i + 1 + Rx], Lx / Rx = how many deltas directly left / right of delta(i) are
  // smaller than it.  Small nodes (<= 4 keys) have Lx + Rx <= 2, so three comparisons per side decide it --
  // no loops, no divergence.  s_info[k]: 0 = not a small local node, else 0x10 | Lx | Rx << 2.
  {
    unsigned int info = 0;
    if (s_local[tid]) {  // node p is local: its range ends inside the CTA, the windows never leave s_delta
      int lx = 0, rx = 0;
      bool go = true;
#pragma unroll
      for (int j = 1; j <= 3; ++j) {
        go = go && (tid - j >= -1) && s_delta[1 + tid - j] < dl;
        lx += go ? 1 : 0;
      }
      go = true;
#pragma unroll
      for (int j = 1; j <= 3; ++j) {
        go = go && (tid + j < kClimbThreads) && s_delta[1 + tid + j] < dl;
        rx += go ? 1 : 0;
      }
      if (lx + rx <= VL_LEAF_MAX - 2) info = 0x10u | (unsigned)lx | ((unsigned)rx << 2);
    }
    s_info[tid] = (unsigned char)info;
  }
  __syncthreads();
  int l = p, r = p;
  if (active) {
    // the small local nodes that contain key p are nested (they are ancestors of leaf p): take the largest
    int best = 1;
#pragma unroll
    for (int j = -(VL_LEAF_MAX - 1); j <= VL_LEAF_MAX - 2; ++j) {
      const int k = tid + j;  // candidate node
      if (k >= 0 && k < kClimbThreads) {
        const unsigned int info = s_info[k];
        const int nl = k - (int)(info & 3u), nr = k + 1 + (int)((info >> 2) & 3u);
        const int size = nr - nl + 1;
        if (info && nl <= tid && tid <= nr && size > best) { best = size; l = B0 + nl; r = B0 + nr; }
      }
    }
  }
  // ---- phase 3: compact the first thread of every leaf to the front of the CTA ----
  const bool leader = active && p == l;
  const unsigned int bal = __ballot_sync(0xffffffffu, leader);
  if (lane == 0) s_wcount[wid] = __popc(bal);
  __syncthreads();
  int before = 0, n_leaders = 0;
#pragma unroll
  for (int ww = 0; ww < kClimbWarps; ++ww) {
    const int c = s_wcount[ww];
    if (ww < wid) before += c;
    n_leaders += c;
  }
  if (leader) s_lead[before + __popc(bal & ((1u << lane) - 1u))] = (l - B0) | ((r - l + 1) << 16);
  __syncthreads();
  const bool run = tid < n_leaders;
  int ref = 0;
  if (run) {
    const int packed = s_lead[tid], first = packed & 0xffff, count = packed >> 16;
    l = B0 + first; r = l + count - 1;
    ref = vl_make_leaf(l, count);
#pragma unroll
    for (int k = 0; k < 3; ++k) { bmin[k] = s_tbox[first][k]; bmax[k] = s_tbox[first][3 + k]; }
    for (int j = 1; j < count; ++j) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        bmin[k] = fminf(bmin[k], s_tbox[first + j][k]);
        bmax[k] = fmaxf(bmax[k], s_tbox[first + j][3 + k]);
      }
    }
  }
  __syncthreads();  // every box has been read: the slots may be written from here on
  if (!run) return;
  if (l == 0 && r == n - 1) { hdr->root_ref = ref; return; }
  while (true) {  // [l, r] stays inside [B0, B1): every delta comes from shared memory
    const bool is_left = (l == 0) ? true : ((r == n - 1) ? false : (s_delta[1 + r - B0] < s_delta[l - B0]));
    const int parent = is_left ? r : l - 1;
    if (parent >= B0 && parent < B1 && s_local[parent - B0]) {
      // ---- the whole sub-tree of `parent` lives in this CTA: the children meet in shared memory ----
      const int k = parent - B0, me = is_left ? 0 : 1;
      // (the slot words are written and read with shared-memory atomics: the hand-over is ordered by the flag, which
      // racecheck cannot see -- atomic accesses are the form of this release / acquire pair that it accepts)
      {
        int* sw = reinterpret_cast<int*>(&s_slot[k][me]);
        const float wv[8] = {bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2], __int_as_float(ref), __int_as_float(is_left ? l : r)};
#pragma unroll
        for (int q = 0; q < 8; ++q) atomicExch(sw + q, __float_as_int(wv[q]));
      }
      // Release / acquire hand-over between the two children of node k (the classic bottom-up LBVH refit): each child
      // publishes its slot, fences, then exchanges the flag -- exactly one of them reads 1, and it is the one that ran
      // its exchange SECOND, i.e. after the sibling's fence made the sibling's slot visible; the fence below orders its
      // own reads after the exchange.  compute-sanitizer --tool racecheck orders shared-memory accesses by barriers only
      // and does not model ordering through an atomic flag: with plain loads / stores on the slots it reported them as
      // hazards (profiles/r01_sanitizer.md, r02_sanitizer.md); the build is bit-reproducible
      // (tests/test_trace_edge_gpu.py::test_full_size_properties).
      __threadfence_block();
      if (atomicExch(&s_flag[k], 1) == 0) break;  // first child to arrive stops here
      __threadfence_block();
      float4 sa, sb;
      {
        int* sr = reinterpret_cast<int*>(&s_slot[k][me ^ 1]);
        sa = make_float4(__int_as_float(atomicOr(sr + 0, 0)), __int_as_float(atomicOr(sr + 1, 0)), __int_as_float(atomicOr(sr + 2, 0)), __int_as_float(atomicOr(sr + 3, 0)));
        sb = make_float4(__int_as_float(atomicOr(sr + 4, 0)), __int_as_float(atomicOr(sr + 5, 0)), __int_as_float(atomicOr(sr + 6, 0)), __int_as_float(atomicOr(sr + 7, 0)));
      }
      const int sref = __float_as_int(sb.z);
      if (is_left) r = __float_as_int(sb.w); else l = __float_as_int(sb.w);
      const int size = r - l + 1;
      if (size > VL_LEAF_MAX) {  // a real inner node: the whole 64 B record in one go
        const float4 ma = make_float4(bmin[0], bmin[1], bmin[2], bmax[0]);
        const float4 mb = make_float4(bmax[1], bmax[2], __int_as_float(ref), __int_as_float(is_left ? l : r));
        float4* q = nodes[parent].q;
        q[is_left ? 0 : 2] = ma; q[is_left ? 1 : 3] = mb;
        q[is_left ? 2 : 0] = sa; q[is_left ? 3 : 1] = sb;
      }
      bmin[0] = fminf(bmin[0], sa.x); bmin[1] = fminf(bmin[1], sa.y); bmin[2] = fminf(bmin[2], sa.z);
      bmax[0] = fmaxf(bmax[0], sa.w); bmax[1] = fmaxf(bmax[1], sb.x); bmax[2] = fmaxf(bmax[2], sb.y);
      ref = size <= VL_LEAF_MAX ? vl_make_leaf(l, size) : parent;
    } else {
      // ---- the sub-tree of `parent` spans CTAs: park this half in the node record and stop; whichever child
      // completes the pair queues the node for k_top_climb (no fence: the kernel boundary orders the data) ----
      float4* q = nodes[parent].q + (is_left ? 0 : 2);
      q[0] = make_float4(bmin[0], bmin[1], bmin[2], bmax[0]);
      q[1] = make_float4(bmax[1], bmax[2], __int_as_float(ref), __int_as_float(is_left ? l : r));
      if (atomicAdd(&flags[parent], 1) == 1) pending[atomicAdd(&hdr->n_pending, 1)] = parent;
      break;
    }
    if (l == 0 && r == n - 1) { hdr->root_ref = ref; break; }  // the whole mesh fits one CTA
  }
}

// The few nodes whose key range spans CTAs of k_emit_climb: one thread per queued node (both halves parked by
// k_emit_climb) merges it and climbs on through global memory -- write own half, fence, count the arrival; the
// first child to arrive at a node stops, the second carries the merged box upwards.  A few thousand threads
// with a dependent chain of ~10-25 levels: latency-bound, but it holds next to no SM resources.
constexpr int kTopThreads = 128;

__global__ void __launch_bounds__(kTopThreads)
k_top_climb(const unsigned int* __restrict__ keys, int n, VlHeader* hdr, VlNode* nodes, int* flags,
            const unsigned int* __restrict__ pending) {
  const int n_pending = hdr->n_pending;
  for (int i = blockIdx.x * kTopThreads + threadIdx.x; i < n_pending; i += gridDim.x * kTopThreads) {
    int parent = (int)pending[i];
    float bmin[3], bmax[3];
    int l, r, ref, climb = 0;
    {
      const float4* q = nodes[parent].q;
      const float4 q0 = __ldcg(q), q1 = __ldcg(q + 1), q2 = __ldcg(q + 2), q3 = __ldcg(q + 3);
      bmin[0] = fminf(q0.x, q2.x); bmin[1] = fminf(q0.y, q2.y); bmin[2] = fminf(q0.z, q2.z);
      bmax[0] = fmaxf(q0.w, q2.w); bmax[1] = fmaxf(q1.x, q3.x); bmax[2] = fmaxf(q1.y, q3.y);
      l = __float_as_int(q1.w); r = __float_as_int(q3.w);
    }
    while (true) {
      const int size = r - l + 1;
      ref = size <= VL_LEAF_MAX ? vl_make_leaf(l, size) : parent;
      ++climb;
      if (l == 0 && r == n - 1) {
        hdr->root_ref = ref;
        atomicMax(&hdr->max_climb, climb);
        break;
      }
      const bool is_left = (l == 0) ? true : ((r == n - 1) ? false : (key_delta(keys, r) < key_delta(keys, l - 1)));
      parent = is_left ? r : l - 1;
      float4* q = nodes[parent].q;
      q[is_left ? 0 : 2] = make_float4(bmin[0], bmin[1], bmin[2], bmax[0]);
      q[is_left ? 1 : 3] = make_float4(bmax[1], bmax[2], __int_as_float(ref), __int_as_float(is_left ? l : r));
      __threadfence();
      if (atomicAdd(&flags[parent], 1) == 0) break;  // first child to arrive stops here
      __threadfence();
      const float4 sa = __ldcg(q + (is_left ? 2 : 0)), sb = __ldcg(q + (is_left ? 3 : 1));
      bmin[0] = fminf(bmin[0], sa.x); bmin[1] = fminf(bmin[1], sa.y); bmin[2] = fminf(bmin[2], sa.z);
      bmax[0] = fmaxf(bmax[0], sa.w); bmax[1] = fmaxf(bmax[1], sb.x); bmax[2] = fmaxf(bmax[2], sb.y);
      if (is_left) r = __float_as_int(sb.w); else l = __float_as_int(sb.w);
    }
  }
}

// The top VL_TOP_LEVELS levels of the finished tree, copied into heap order (node i -> children 2 i + 1, 2 i + 2) so that
// k_trace_persistent can stage them in shared memory with ONE bulk copy (cp.async.bulk) and address them without
// following references.  A slot whose node does not exist (its parent's child is a leaf, or absent) is never read.
__global__ void __launch_bounds__(1024)
k_top_pack(const VlHeader* __restrict__ hdr, const VlNode* __restrict__ nodes, VlNode* __restrict__ top) {
  __shared__ int s_ref[VL_TOP_NODES + 1];
  if (threadIdx.x == 0) s_ref[0] = hdr->n_tris > 0 ? hdr->root_ref : -1;
  __syncthreads();
  for (int L = 0; L < VL_TOP_LEVELS; ++L) {
    const int n = 1 << L, base = n - 1;
    for (int t = threadIdx.x; t < n; t += 1024) {
      const int i = base + t, g = s_ref[i];
      int c0 = -1, c1 = -1;
      if (g >= 0) {
        const float4 a = nodes[g].q[0], b = nodes[g].q[1], c = nodes[g].q[2], e = nodes[g].q[3];
        top[i].q[0] = a; top[i].q[1] = b; top[i].q[2] = c; top[i].q[3] = e;
        c0 = __float_as_int(b.z); c1 = __float_as_int(e.z);
      }
      if (L + 1 < VL_TOP_LEVELS) { s_ref[2 * i + 1] = c0; s_ref[2 * i + 2] = c1; }
    }
    __syncthreads();
  }
}

}  // namespace

int g_debug_build_stop = 0;  // vl_debug_build_stop(): 0 full build; 1 / 2 / 3 = return after bounds / morton / sort (timing only)
extern "C" void vl_debug_build_stop(int stage) { g_debug_build_stop = stage; }

int vl_bvh_build_launch(const float* d_verts, const int* d_faces, const int* d_colors, const float* d_rem,
                        int n_verts, int n_faces, void* d_blob, cudaStream_t stream) {
  char* blob = static_cast<char*>(d_blob);
  VlBlobLayout L = vl_blob_layout(n_faces);
  VlHeader* hdr = reinterpret_cast<VlHeader*>(blob);
  unsigned int* ghist = reinterpret_cast<unsigned int*>(blob + L.off_ghist);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(blob + L.off_tickets);
  unsigned int* tile_state = reinterpret_cast<unsigned int*>(blob + L.off_tile_state);
  k_init_header<<<1, 1024, 0, stream>>>(hdr, n_faces, ghist, tickets);
  VL_LAUNCH_CHECK("k_init_header");
  if (n_faces <= 0) return VL_OK;
  unsigned int* keys0 = reinterpret_cast<unsigned int*>(blob + L.off_keys0);
  unsigned int* keys1 = reinterpret_cast<unsigned int*>(blob + L.off_keys1);
  unsigned int* vals0 = reinterpret_cast<unsigned int*>(blob + L.off_vals0);
  unsigned int* vals1 = reinterpret_cast<unsigned int*>(blob + L.off_vals1);
  int* flags = reinterpret_cast<int*>(blob + L.off_flags);

  // persistent grids: SM count (148 on B200) x 4 resident CTAs
  int nb_verts = (n_verts + kThreads - 1) / kThreads;
  if (nb_verts > vl_sm_count() * 4) nb_verts = vl_sm_count() * 4;
  if (nb_verts < 1) nb_verts = 1;
  { VlProfScope ps(VL_ST_BOUNDS, stream);
  k_bounds<<<nb_verts, kThreads, 0, stream>>>(d_verts, n_verts, hdr); }
  VL_LAUNCH_CHECK("k_bounds");
  if (g_debug_build_stop == 1) return VL_OK;
  const int nt = L.n_sort_tiles;
  const int n_state_words = 256 * VL_SORT_PASSES * L.n_state_tiles;
  int nb_faces = (n_faces + kThreads - 1) / kThreads;
  if (nb_faces > vl_sm_count() * 4) nb_faces = vl_sm_count() * 4;
  { VlProfScope ps(VL_ST_MORTON, stream);
  k_morton<<<nb_faces, kThreads, 0, stream>>>(d_verts, d_faces, n_verts, n_faces, hdr, keys0, flags, ghist, tile_state,
                                             n_state_words); }
  VL_LAUNCH_CHECK("k_morton");
  if (g_debug_build_stop == 2) return VL_OK;

  const unsigned int* kin = keys0;
  const unsigned int* vin = nullptr;  // pass 0 generates the identity permutation on the fly
  unsigned int* kout = keys1;
  unsigned int* vout = vals1;
  for (int pass = 0; pass < VL_SORT_PASSES; ++pass) {
    { VlProfScope ps(VL_ST_SORT_PASS, stream);
    k_sort_pass<<<nt, VL_SORT_THREADS, 0, stream>>>(kin, vin, kout, vout, n_faces, 8 * pass, ghist + 256 * pass,
                                                   tile_state + (size_t)256 * L.n_state_tiles * pass, tickets + pass, nt); }
    VL_LAUNCH_CHECK("k_sort_pass");
    kin = kout;
    vin = vout;
    kout = (kout == keys1) ? keys0 : keys1;
    vout = (vout == vals1) ? vals0 : vals1;
  }
  // after 4 passes the sorted (key, face id) pairs are back in keys0 / vals0
  if (g_debug_build_stop == 3) return VL_OK;
  const int nb_climb = (n_faces + kClimbThreads - 1) / kClimbThreads;
  VlNode* nodes = reinterpret_cast<VlNode*>(blob + L.off_nodes);
  { VlProfScope ps(VL_ST_EMIT_CLIMB, stream);
  // vals1 is free after the sort: it becomes the queue of nodes for k_top_climb (at most n / 2 entries)
  k_emit_climb<<<nb_climb, kClimbThreads, 0, stream>>>(d_verts, d_faces, d_colors, d_rem, n_verts, n_faces, keys0, vals0,
                                                      hdr, nodes, reinterpret_cast<VlTri*>(blob + L.off_tris),
                                                      reinterpret_cast<int4*>(blob + L.off_c0), flags, vals1); }
  VL_LAUNCH_CHECK("k_emit_climb");
  if (nb_climb > 1) {
    VlProfScope ps(VL_ST_TOP_CLIMB, stream);
    int nb_top = (nb_climb * 16 + kTopThreads - 1) / kTopThreads;  // ~ a dozen queued nodes per CTA is typical
    if (nb_top > vl_sm_count() * 8) nb_top = vl_sm_count() * 8;
    k_top_climb<<<nb_top, kTopThreads, 0, stream>>>(keys0, n_faces, hdr, nodes, flags, vals1);
    VL_LAUNCH_CHECK("k_top_climb");
  }
  k_top_pack<<<1, 1024, 0, stream>>>(hdr, nodes, reinterpret_cast<VlNode*>(blob + L.off_top));
  VL_LAUNCH_CHECK("k_top_pack");
  return VL_OK;
}
