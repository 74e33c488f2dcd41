// vl_trace.cu -- (ii) per-ray closest-hit BVH traversal + Moller-Trumbore, sm_100a.
//
// Replaces the reference's hot loop: RayTracer.cpp:62-92 (per-pixel loop and write-back),
// BVH::getIntersection BVH.cpp:19-110, BBox::intersect BBox.cpp:52-100,
// Triangle::getIntersection Triangle.h:27-50, normalize Vector3.h:73-89.
//
// Result contract: the closest hit under the reference's Moller-Trumbore
// arithmetic over ALL triangles -- the tree only prunes.  The slab test is conservative
// (padded leaf boxes + relaxed far bound) so it can never cull a triangle the triangle test
// would accept; exact-t ties go to the smaller original face index.
#include "vl_common.cuh"

namespace {

constexpr int kTraceThreads = 128;
constexpr int kSmemStack = 24;   // per-thread stack entries held in shared memory
constexpr int kLocalStack = 48;  // overflow (LBVH height is bounded by 64 key bits)

struct Hit {
  float t;
  int pos;     // sorted triangle position
  int orig;    // original face index
  float rem;
};

// slab test of one box; lanes with NaN (0 * inf) impose no constraint, like the reference's
// min/max ordering (BBox.cpp:70-80).  `exact_nan` is only needed when a direction component is 0.
template <bool kNanFilter>
__device__ __forceinline__ void slab(const float bminx, const float bminy, const float bminz, const float bmaxx,
                                     const float bmaxy, const float bmaxz, const float3 o, const float3 inv_d,
                                     float* tnear, float* tfar) {
  float l1x = (bminx - o.x) * inv_d.x, l2x = (bmaxx - o.x) * inv_d.x;
  float l1y = (bminy - o.y) * inv_d.y, l2y = (bmaxy - o.y) * inv_d.y;
  float l1z = (bminz - o.z) * inv_d.z, l2z = (bmaxz - o.z) * inv_d.z;
  float lox, loy, loz, hix, hiy, hiz;
  if (kNanFilter) {
    hix = fmaxf(fminf(l1x, INFINITY), fminf(l2x, INFINITY)); lox = fminf(fmaxf(l1x, -INFINITY), fmaxf(l2x, -INFINITY));
    hiy = fmaxf(fminf(l1y, INFINITY), fminf(l2y, INFINITY)); loy = fminf(fmaxf(l1y, -INFINITY), fmaxf(l2y, -INFINITY));
    hiz = fmaxf(fminf(l1z, INFINITY), fminf(l2z, INFINITY)); loz = fminf(fmaxf(l1z, -INFINITY), fmaxf(l2z, -INFINITY));
    if (l1x != l1x || l2x != l2x) { lox = -INFINITY; hix = INFINITY; }
    if (l1y != l1y || l2y != l2y) { loy = -INFINITY; hiy = INFINITY; }
    if (l1z != l1z || l2z != l2z) { loz = -INFINITY; hiz = INFINITY; }
  } else {
    lox = fminf(l1x, l2x); hix = fmaxf(l1x, l2x);
    loy = fminf(l1y, l2y); hiy = fmaxf(l1y, l2y);
    loz = fminf(l1z, l2z); hiz = fmaxf(l1z, l2z);
  }
  *tnear = fmaxf(fmaxf(lox, loy), loz);
  *tfar = fminf(fminf(hix, hiy), hiz);
}

// "while-while" traversal: every lane first walks inner nodes until it holds a leaf (lanes that got there early
// idle instead of dragging the warp through both code paths), then the leaves are tested together.
constexpr int kDoneRef = (int)0x80000000;  // not a valid leaf reference (first would be >= 2^28)

template <bool kNanFilter>
__device__ __forceinline__ void traverse(const VlNode* __restrict__ nodes, const VlTri* __restrict__ tris,
                                         int root_ref, const float3 o, const float3 d, const float3 inv_d,
                                         uint2 (*sstack)[kTraceThreads], Hit* best, int* n_nodes, int* n_tris) {
  uint2 lstack[kLocalStack];
  int sp = 0;
  int ref = root_ref;
  const int tid = threadIdx.x;
  // pop, skipping entries that can no longer beat the best hit (BVH.cpp:41)
  auto pop = [&]() -> int {
    while (sp > 0) {
      --sp;
      const uint2 ent = sp < kSmemStack ? sstack[sp][tid] : lstack[sp - kSmemStack];
      if (__uint_as_float(ent.y) <= best->t) return (int)ent.x;
    }
    return kDoneRef;
  };
  while (ref != kDoneRef) {
    while (ref >= 0) {
      ++*n_nodes;
      const float4* q = nodes[ref].q;
      const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), e = __ldg(q + 3);
      float tn0, tf0, tn1, tf1;
      slab<kNanFilter>(a.x, a.y, a.z, a.w, b.x, b.y, o, inv_d, &tn0, &tf0);
      slab<kNanFilter>(c.x, c.y, c.z, c.w, e.x, e.y, o, inv_d, &tn1, &tf1);
      // BBox.cpp:97 `tfar >= 0 && tfar >= tnear`, with a relaxed far bound; BVH.cpp:41 prune
      tf0 = tf0 * 1.0000004f; tf1 = tf1 * 1.0000004f;
      const bool h0 = (tf0 >= 0.f) & (tf0 >= tn0) & (tn0 <= best->t);
      const bool h1 = (tf1 >= 0.f) & (tf1 >= tn1) & (tn1 <= best->t);
      const int r0 = __float_as_int(b.z), r1 = __float_as_int(e.z);
      if (h0 & h1) {
        const bool swap = tn1 < tn0;  // BVH.cpp:77: nearer child first
        const int far_ref = swap ? r0 : r1;
        const float far_t = swap ? tn0 : tn1;
        ref = swap ? r1 : r0;
        const uint2 ent = make_uint2((unsigned)far_ref, __float_as_uint(far_t));
        if (sp < kSmemStack) sstack[sp][tid] = ent; else lstack[sp - kSmemStack] = ent;
        ++sp;
      } else if (h0) {
        ref = r0;
      } else if (h1) {
        ref = r1;
      } else {
        ref = pop();
      }
    }
    if (ref == kDoneRef) break;
    const int first = vl_leaf_first(ref), count = vl_leaf_count(ref);
    *n_tris += count;
    for (int k = 0; k < count; ++k) {
      const float4* tq = reinterpret_cast<const float4*>(tris + first + k);
      const float4 v0 = __ldg(tq), e1 = __ldg(tq + 1), e2 = __ldg(tq + 2);
      float t;
      if (vl_tri_hit(v0, e1, e2, o, d, &t)) {
        const int orig = __float_as_int(v0.w);
        if (t < best->t || (t == best->t && orig < best->orig)) {
          best->t = t; best->pos = first + k; best->orig = orig; best->rem = e1.w;
        }
      }
    }
    ref = pop();
  }
}

// kTiled: a warp is an 8 x 4 tile of the H x W beam grid and a CTA a 16 x 8 tile (neighbouring beams walk the
// same sub-trees, so their node fetches hit in L1 and their control flow stays together); otherwise rays are
// taken in storage order (ray sets that are not a beam grid).
template <bool kTiled>
__global__ void __launch_bounds__(kTraceThreads)
k_trace(const VlHeader* __restrict__ hdr, const VlNode* __restrict__ nodes, const VlTri* __restrict__ tris,
        const int4* __restrict__ c0, const float* __restrict__ rays, const float* __restrict__ origin, int n_traced,
        int width, int height, float* __restrict__ endpoints, int* __restrict__ endcolors, float* __restrict__ range,
        float* __restrict__ endrem, int* __restrict__ tri_id, bool zero_misses, bool prenorm, int* __restrict__ stats) {
  __shared__ uint2 sstack[kSmemStack][kTraceThreads];
  int r;
  if (kTiled) {
    const int tiles_x = (width + 15) >> 4;
    const int bx = blockIdx.x % tiles_x, by = blockIdx.x / tiles_x;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = (bx << 4) + ((w & 1) << 3) + (lane & 7), row = (by << 3) + ((w >> 1) << 2) + (lane >> 3);
    if (col >= width || row >= height) return;
    r = row * width + col;
  } else {
    r = blockIdx.x * kTraceThreads + threadIdx.x;
    if (r >= n_traced) return;
  }
  const float3 o = make_float3(__ldg(origin), __ldg(origin + 1), __ldg(origin + 2));
  const float3 d = vl_ray_dir(rays, (size_t)r, prenorm);
  const float3 inv_d = make_float3(__fdiv_rn(1.0f, d.x), __fdiv_rn(1.0f, d.y), __fdiv_rn(1.0f, d.z));  // Ray.h:11-12
  Hit best;
  best.t = 999999999.f;  // BVH.cpp:20
  best.pos = -1; best.orig = 0x7fffffff; best.rem = 0.f;
  const int root = hdr->root_ref;
  int n_nodes = 0, n_tris = 0;
  if (hdr->n_tris > 0) {
    if (d.x == 0.f || d.y == 0.f || d.z == 0.f || !(d.x == d.x))
      traverse<true>(nodes, tris, root, o, d, inv_d, sstack, &best, &n_nodes, &n_tris);
    else
      traverse<false>(nodes, tris, root, o, d, inv_d, sstack, &best, &n_nodes, &n_tris);
  }
  if (stats) { stats[2 * (size_t)r] = n_nodes; stats[2 * (size_t)r + 1] = n_tris; }
  if (best.pos >= 0) {
    // RayTracer.cpp:73-90 write-back, BVH.cpp:106-107 hit = o + d * t
    const int4 col = __ldg(c0 + best.pos);
    endpoints[3 * (size_t)r + 0] = __fadd_rn(o.x, __fmul_rn(d.x, best.t));
    endpoints[3 * (size_t)r + 1] = __fadd_rn(o.y, __fmul_rn(d.y, best.t));
    endpoints[3 * (size_t)r + 2] = __fadd_rn(o.z, __fmul_rn(d.z, best.t));
    endcolors[3 * (size_t)r + 0] = col.x;
    endcolors[3 * (size_t)r + 1] = col.y;
    endcolors[3 * (size_t)r + 2] = col.z;
    endrem[r] = best.rem;
    range[r] = best.t;
  } else if (zero_misses) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { endpoints[3 * (size_t)r + k] = 0.f; endcolors[3 * (size_t)r + k] = 0; }
    endrem[r] = 0.f;
    range[r] = 0.f;
  }
  if (tri_id) tri_id[r] = best.pos >= 0 ? best.orig : -1;
}

// ---------------------------------------------------------------------------
// test aid: every ray against every triangle, triangles staged through shared memory
// ---------------------------------------------------------------------------
constexpr int kBfTile = 128;

__global__ void __launch_bounds__(kTraceThreads)
k_trace_bruteforce(const float* __restrict__ verts, const int* __restrict__ faces, const int* __restrict__ colors,
                   const float* __restrict__ rem, int n_verts, int n_faces, const float* __restrict__ rays,
                   const float* __restrict__ origin, int n_traced, float* __restrict__ endpoints,
                   int* __restrict__ endcolors, float* __restrict__ range, float* __restrict__ endrem,
                   int* __restrict__ tri_id, bool prenorm) {
  __shared__ float4 sv0[kBfTile], se1[kBfTile], se2[kBfTile];
  const int r = blockIdx.x * kTraceThreads + threadIdx.x;
  const bool active = r < n_traced;
  const float3 o = make_float3(__ldg(origin), __ldg(origin + 1), __ldg(origin + 2));
  float3 d = make_float3(0.f, 0.f, 0.f);
  if (active) d = vl_ray_dir(rays, (size_t)r, prenorm);
  float best_t = 999999999.f;
  int best = -1;
  for (int base = 0; base < n_faces; base += kBfTile) {
    __syncthreads();
    const int f = base + threadIdx.x;
    if (threadIdx.x < kBfTile) {
      float4 v0 = make_float4(0, 0, 0, 0), e1 = v0, e2 = v0;
      if (f < n_faces) {
        int i0 = faces[3 * (size_t)f], i1 = faces[3 * (size_t)f + 1], i2 = faces[3 * (size_t)f + 2];
        if ((unsigned)i0 < (unsigned)n_verts && (unsigned)i1 < (unsigned)n_verts && (unsigned)i2 < (unsigned)n_verts) {
          v0 = make_float4(verts[3 * (size_t)i0], verts[3 * (size_t)i0 + 1], verts[3 * (size_t)i0 + 2], 0.f);
          e1 = make_float4(__fsub_rn(verts[3 * (size_t)i1], v0.x), __fsub_rn(verts[3 * (size_t)i1 + 1], v0.y),
                           __fsub_rn(verts[3 * (size_t)i1 + 2], v0.z), 0.f);
          e2 = make_float4(__fsub_rn(verts[3 * (size_t)i2], v0.x), __fsub_rn(verts[3 * (size_t)i2 + 1], v0.y),
                           __fsub_rn(verts[3 * (size_t)i2 + 2], v0.z), 0.f);
        }
      }
      sv0[threadIdx.x] = v0; se1[threadIdx.x] = e1; se2[threadIdx.x] = e2;
    }
    __syncthreads();
    if (active) {
      const int lim = min(kBfTile, n_faces - base);
      for (int k = 0; k < lim; ++k) {
        float t;
        if (vl_tri_hit(sv0[k], se1[k], se2[k], o, d, &t) && t < best_t) { best_t = t; best = base + k; }
      }
    }
  }
  if (!active) return;
  if (best >= 0) {
    int i0 = faces[3 * (size_t)best], i1 = faces[3 * (size_t)best + 1], i2 = faces[3 * (size_t)best + 2];
    endpoints[3 * (size_t)r + 0] = __fadd_rn(o.x, __fmul_rn(d.x, best_t));
    endpoints[3 * (size_t)r + 1] = __fadd_rn(o.y, __fmul_rn(d.y, best_t));
    endpoints[3 * (size_t)r + 2] = __fadd_rn(o.z, __fmul_rn(d.z, best_t));
    endcolors[3 * (size_t)r + 0] = (int)(float)colors[3 * (size_t)i0];
    endcolors[3 * (size_t)r + 1] = (int)(float)colors[3 * (size_t)i0 + 1];
    endcolors[3 * (size_t)r + 2] = (int)(float)colors[3 * (size_t)i0 + 2];
    endrem[r] = __fdiv_rn(__fadd_rn(__fadd_rn(rem[i0], rem[i1]), rem[i2]), 3.0f);
    range[r] = best_t;
  }
  if (tri_id) tri_id[r] = best;
}

// ---------------------------------------------------------------------------
// warp-packet traversal: the 32 rays of a warp are a TW x TH tile of the H x W beam grid
// (neighbouring beams), they walk the tree TOGETHER with one shared stack:
//   * node / triangle records are fetched with warp-uniform addresses (1 L1 wavefront per
//     128-bit load instead of up to 32 for per-lane pointers -- the per-thread kernel above is
//     L1TEX-wavefront bound, profiles/r01_*),
//   * a child is entered when ANY lane's slab test passes and is not pruned by that lane's best t
//     (ballot), the child preferred as "near" by more lanes first (BVH.cpp:77 generalised),
//   * the far child is pushed with PER-LANE entry distances (+inf = lane does not enter), so
//     popping prunes per lane exactly like BVH.cpp:41.
// Every lane still sees every triangle the per-thread traversal would test (a superset), so the
// result is the same closest hit; only the amount of work differs.
// ---------------------------------------------------------------------------
constexpr int kPktWarps = 4;
constexpr int kPktStack = 64;  // LBVH height <= 32 key bits + 32 tie-break bits

template <int TW>
__global__ void __launch_bounds__(kPktWarps * 32)
k_trace_packet(const VlHeader* __restrict__ hdr, const VlNode* __restrict__ nodes, const VlTri* __restrict__ tris,
               const int4* __restrict__ c0, const float* __restrict__ rays, const float* __restrict__ origin,
               int width, int height, float* __restrict__ endpoints, int* __restrict__ endcolors,
               float* __restrict__ range, float* __restrict__ endrem, int* __restrict__ tri_id, bool zero_misses,
               bool prenorm, int* __restrict__ stats) {
  constexpr int TH = 32 / TW;
  __shared__ float s_tn[kPktWarps][kPktStack][32];
  __shared__ int s_ref[kPktWarps][kPktStack];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int tiles_x = (width + TW - 1) / TW, tiles_y = (height + TH - 1) / TH;
  const int tile = blockIdx.x * kPktWarps + wib;
  if (tile >= tiles_x * tiles_y) return;  // warp-uniform
  const int col = (tile % tiles_x) * TW + (lane % TW), row = (tile / tiles_x) * TH + (lane / TW);
  const bool valid = col < width && row < height;
  const size_t r = (size_t)row * width + col;
  const float3 o = make_float3(__ldg(origin), __ldg(origin + 1), __ldg(origin + 2));
  float3 d = make_float3(0.f, 0.f, 0.f);
  if (valid) d = vl_ray_dir(rays, r, prenorm);
  const float3 inv_d = make_float3(__fdiv_rn(1.0f, d.x), __fdiv_rn(1.0f, d.y), __fdiv_rn(1.0f, d.z));
  // a lane with a zero / NaN direction component needs the NaN-filtering slab (BBox.cpp:70-80)
  const bool odd_lane = valid && (d.x == 0.f || d.y == 0.f || d.z == 0.f || !(d.x == d.x));
  const bool any_odd = __any_sync(0xffffffffu, odd_lane);
  float best_t = valid ? 999999999.f : -INFINITY;  // BVH.cpp:20; invalid lanes never enter anything
  int best_pos = -1, best_orig = 0x7fffffff;
  float best_rem = 0.f;
  int n_nodes = 0, n_tris = 0;
  int ref = hdr->root_ref;
  int sp = 0;
  if (hdr->n_tris > 0) {
    while (true) {
      if (ref >= 0) {
        ++n_nodes;
        const float4* q = nodes[ref].q;  // warp-uniform address
        const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), e = __ldg(q + 3);
        float tn0, tf0, tn1, tf1;
        if (any_odd) {
          slab<true>(a.x, a.y, a.z, a.w, b.x, b.y, o, inv_d, &tn0, &tf0);
          slab<true>(c.x, c.y, c.z, c.w, e.x, e.y, o, inv_d, &tn1, &tf1);
        } else {
          slab<false>(a.x, a.y, a.z, a.w, b.x, b.y, o, inv_d, &tn0, &tf0);
          slab<false>(c.x, c.y, c.z, c.w, e.x, e.y, o, inv_d, &tn1, &tf1);
        }
        tf0 = tf0 * 1.0000004f; tf1 = tf1 * 1.0000004f;
        const bool h0 = (tf0 >= 0.f) & (tf0 >= tn0) & (tn0 <= best_t);
        const bool h1 = (tf1 >= 0.f) & (tf1 >= tn1) & (tn1 <= best_t);
        const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
        const int r0 = __float_as_int(b.z), r1 = __float_as_int(e.z);
        if (m0 && m1) {
          const unsigned want1 = __ballot_sync(0xffffffffu, h1 && (!h0 || tn1 < tn0));
          const bool first1 = 2 * __popc(want1) > __popc(m0 | m1);
          s_ref[wib][sp] = first1 ? r0 : r1;
          s_tn[wib][sp][lane] = first1 ? (h0 ? tn0 : INFINITY) : (h1 ? tn1 : INFINITY);
          ++sp;
          ref = first1 ? r1 : r0;
          continue;
        }
        if (m0) { ref = r0; continue; }
        if (m1) { ref = r1; continue; }
      } else {
        const int first = vl_leaf_first(ref), count = vl_leaf_count(ref);
        n_tris += count;
        for (int k = 0; k < count; ++k) {
          const float4* tq = reinterpret_cast<const float4*>(tris + first + k);  // warp-uniform
          const float4 v0 = __ldg(tq), e1 = __ldg(tq + 1), e2 = __ldg(tq + 2);
          float t;
          if (valid && vl_tri_hit(v0, e1, e2, o, d, &t)) {
            const int orig = __float_as_int(v0.w);
            if (t < best_t || (t == best_t && orig < best_orig)) {
              best_t = t; best_pos = first + k; best_orig = orig; best_rem = e1.w;
            }
          }
        }
      }
      bool found = false;
      while (sp > 0) {
        --sp;
        if (__any_sync(0xffffffffu, s_tn[wib][sp][lane] <= best_t)) { ref = s_ref[wib][sp]; found = true; break; }
      }
      if (!found) break;
    }
  }
  if (!valid) return;
  if (best_pos >= 0) {
    const int4 col4 = __ldg(c0 + best_pos);
    endpoints[3 * r + 0] = __fadd_rn(o.x, __fmul_rn(d.x, best_t));
    endpoints[3 * r + 1] = __fadd_rn(o.y, __fmul_rn(d.y, best_t));
    endpoints[3 * r + 2] = __fadd_rn(o.z, __fmul_rn(d.z, best_t));
    endcolors[3 * r + 0] = col4.x;
    endcolors[3 * r + 1] = col4.y;
    endcolors[3 * r + 2] = col4.z;
    endrem[r] = best_rem;
    range[r] = best_t;
  } else if (zero_misses) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { endpoints[3 * r + k] = 0.f; endcolors[3 * r + k] = 0; }
    endrem[r] = 0.f;
    range[r] = 0.f;
  }
  if (tri_id) tri_id[r] = best_pos >= 0 ? best_orig : -1;
  if (stats) { stats[2 * r] = n_nodes; stats[2 * r + 1] = n_tris; }
}

// ---------------------------------------------------------------------------
// North-star (ii) as BASELINE.json spells it: persistent warps that pull rays from a global counter, warp-wide
// compaction (lanes whose ray has finished take the next rays as soon as fewer than 8 lanes of the warp are live), the
// per-lane stack in shared memory, and the top VL_TOP_LEVELS levels of the tree (2047 nodes x 64 B = 128 KB, heap
// order, k_top_pack) staged ONCE per CTA by the TMA engine: cp.async.bulk global -> shared, completion on an mbarrier
// (UBLKCP in the SASS).  One 512-thread CTA per SM: 128 KB of nodes + 48 KB of stacks.  Same result contract and the
// same arithmetic as k_trace (slab<>, vl_tri_hit): bit-identical outputs, tests/test_trace_gpu.py.
// ---------------------------------------------------------------------------
constexpr int kPtThreads = 512;
constexpr int kPtStack = 12;        // stack entries per lane in shared memory (deeper: local memory)
constexpr int kPtLocal = 52;
constexpr int kTopFlag = 0x40000000;   // reference = heap index into the staged top of the tree
constexpr int kRefillIdle = 24;     // refill as soon as 24 lanes are idle (fewer than 8 live)
constexpr size_t kPtTopBytes = 64 * (size_t)VL_TOP_NODES;                     // 131 008, a multiple of 16
constexpr size_t kPtSmemBytes = 64 * (size_t)(VL_TOP_NODES + 1) + sizeof(uint2) * kPtStack * kPtThreads;

__device__ unsigned int g_pt_counters[64];

__device__ __forceinline__ unsigned int smem_addr(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

template <bool kNanFilter>
__device__ __forceinline__ int pt_node_step(int ref, const VlNode* __restrict__ nodes, const VlNode* s_top, const float3 o,
                                            const float3 inv_d, float best_t, uint2* s_stack, uint2* lstack, int& sp, int tid,
                                            bool* popped) {
  float4 a, b, c, e;
  const bool cached = (ref & kTopFlag) != 0;
  const int hi = ref & ~kTopFlag;
  if (cached) { const float4* q = s_top[hi].q; a = q[0]; b = q[1]; c = q[2]; e = q[3]; }
  else { const float4* q = nodes[ref].q; a = __ldg(q); b = __ldg(q + 1); c = __ldg(q + 2); e = __ldg(q + 3); }
  float tn0, tf0, tn1, tf1;
  slab<kNanFilter>(a.x, a.y, a.z, a.w, b.x, b.y, o, inv_d, &tn0, &tf0);
  slab<kNanFilter>(c.x, c.y, c.z, c.w, e.x, e.y, o, inv_d, &tn1, &tf1);
  tf0 = tf0 * 1.0000004f; tf1 = tf1 * 1.0000004f;
  const bool h0 = (tf0 >= 0.f) & (tf0 >= tn0) & (tn0 <= best_t);
  const bool h1 = (tf1 >= 0.f) & (tf1 >= tn1) & (tn1 <= best_t);
  int r0 = __float_as_int(b.z), r1 = __float_as_int(e.z);
  if (cached && hi < (1 << (VL_TOP_LEVELS - 1)) - 1) {   // the children of a staged node above the last level are staged too
    if (r0 >= 0) r0 = (2 * hi + 1) | kTopFlag;
    if (r1 >= 0) r1 = (2 * hi + 2) | kTopFlag;
  }
  *popped = false;
  if (h0 & h1) {
    const bool swap = tn1 < tn0;
    const uint2 ent = make_uint2((unsigned)(swap ? r0 : r1), __float_as_uint(swap ? tn0 : tn1));
    if (sp < kPtStack) s_stack[sp * kPtThreads + tid] = ent; else lstack[sp - kPtStack] = ent;
    ++sp;
    return swap ? r1 : r0;
  }
  if (h0) return r0;
  if (h1) return r1;
  *popped = true;
  return kDoneRef;   // the caller pops
}

extern __shared__ __align__(128) unsigned char pt_smem[];

__global__ void __launch_bounds__(kPtThreads, 1)
k_trace_persistent(const VlHeader* __restrict__ hdr, const VlNode* __restrict__ nodes, const VlNode* __restrict__ top_g,
                   const VlTri* __restrict__ tris, const int4* __restrict__ c0, const float* __restrict__ rays,
                   const float* __restrict__ origin, int n_traced, int width, int height, float* __restrict__ endpoints,
                   int* __restrict__ endcolors, float* __restrict__ range, float* __restrict__ endrem,
                   int* __restrict__ tri_id, bool zero_misses, bool prenorm, unsigned int* __restrict__ counter) {
  VlNode* s_top = reinterpret_cast<VlNode*>(pt_smem);
  uint2* s_stack = reinterpret_cast<uint2*>(pt_smem + 64 * (size_t)(VL_TOP_NODES + 1));
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x, lane = tid & 31;
  const unsigned int lt = (1u << lane) - 1u;
  // ---- stage the top of the tree: one thread arms the mbarrier with the byte count and issues the bulk copies
  const unsigned int bar = smem_addr(&s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned int)kPtTopBytes) : "memory");
    const char* src = reinterpret_cast<const char*>(top_g);
    for (unsigned int off = 0; off < (unsigned int)kPtTopBytes; off += 32768u) {
      const unsigned int n = min(32768u, (unsigned int)kPtTopBytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_addr(pt_smem + off)), "l"(src + off), "r"(n), "r"(bar) : "memory");
    }
  }
  {
    unsigned int done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
  }
  // ---- persistent warps
  const bool grid_rays = height >= 4 && width >= 8;
  const int tiles_x = (width + 7) >> 3, tiles_y = (height + 3) >> 2;
  const int n_work = grid_rays ? tiles_x * tiles_y * 32 : n_traced;
  const float3 o = make_float3(__ldg(origin), __ldg(origin + 1), __ldg(origin + 2));
  const int root = hdr->n_tris > 0 ? (hdr->root_ref >= 0 ? (0 | kTopFlag) : hdr->root_ref) : kDoneRef;
  uint2 lstack[kPtLocal];
  int r = -1, ref = kDoneRef, sp = 0;
  float3 d = make_float3(0.f, 0.f, 0.f), inv_d = d;
  bool odd = false;
  Hit best;
  best.t = 0.f; best.pos = -1; best.orig = 0; best.rem = 0.f;
  bool more = true;
  for (;;) {
    const unsigned int idle = __ballot_sync(0xffffffffu, r < 0);
    if (idle == 0xffffffffu && !more) break;
    if (more && __popc(idle) >= kRefillIdle) {               // warp-uniform
      const int cnt = __popc(idle);
      int base = 0;
      if (lane == 0) base = (int)atomicAdd(counter, (unsigned int)cnt);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (r < 0) {
        const int idx = base + __popc(idle & lt);
        if (idx < n_work) {
          int rr = idx;
          if (grid_rays) {                                     // 8 x 4 beam tiles, tile-major: neighbours stay together
            const int tile = idx >> 5, in = idx & 31;
            const int col = (tile % tiles_x) * 8 + (in & 7), row = (tile / tiles_x) * 4 + (in >> 3);
            rr = (col < width && row < height) ? row * width + col : -1;
          }
          if (rr >= 0) {
            r = rr;
            d = vl_ray_dir(rays, (size_t)r, prenorm);
            inv_d = make_float3(__fdiv_rn(1.0f, d.x), __fdiv_rn(1.0f, d.y), __fdiv_rn(1.0f, d.z));  // Ray.h:11-12
            odd = d.x == 0.f || d.y == 0.f || d.z == 0.f || !(d.x == d.x);
            best.t = 999999999.f;  // BVH.cpp:20
            best.pos = -1; best.orig = 0x7fffffff; best.rem = 0.f;
            sp = 0;
            ref = root;
          }
        }
      }
      if (base + cnt >= n_work) more = false;
    }
    if (r >= 0) {
      auto pop = [&]() -> int {
        while (sp > 0) {
          --sp;
          const uint2 ent = sp < kPtStack ? s_stack[sp * kPtThreads + tid] : lstack[sp - kPtStack];
          if (__uint_as_float(ent.y) <= best.t) return (int)ent.x;
        }
        return kDoneRef;
      };
      while (ref >= 0) {                                       // inner nodes until this lane holds a leaf (or is done)
        bool popped;
        int nxt = odd ? pt_node_step<true>(ref, nodes, s_top, o, inv_d, best.t, s_stack, lstack, sp, tid, &popped)
                      : pt_node_step<false>(ref, nodes, s_top, o, inv_d, best.t, s_stack, lstack, sp, tid, &popped);
        ref = popped ? pop() : nxt;
      }
      if (ref != kDoneRef) {
        const int first = vl_leaf_first(ref), count = vl_leaf_count(ref);
        for (int k = 0; k < count; ++k) {
          const float4* tq = reinterpret_cast<const float4*>(tris + first + k);
          const float4 v0 = __ldg(tq), e1 = __ldg(tq + 1), e2 = __ldg(tq + 2);
          float t;
          if (vl_tri_hit(v0, e1, e2, o, d, &t)) {
            const int orig = __float_as_int(v0.w);
            if (t < best.t || (t == best.t && orig < best.orig)) { best.t = t; best.pos = first + k; best.orig = orig; best.rem = e1.w; }
          }
        }
        ref = pop();
      }
      if (ref == kDoneRef) {                                   // this ray is finished: write it, free the lane
        if (best.pos >= 0) {
          const int4 col = __ldg(c0 + best.pos);
          endpoints[3 * (size_t)r + 0] = __fadd_rn(o.x, __fmul_rn(d.x, best.t));
          endpoints[3 * (size_t)r + 1] = __fadd_rn(o.y, __fmul_rn(d.y, best.t));
          endpoints[3 * (size_t)r + 2] = __fadd_rn(o.z, __fmul_rn(d.z, best.t));
          endcolors[3 * (size_t)r + 0] = col.x;
          endcolors[3 * (size_t)r + 1] = col.y;
          endcolors[3 * (size_t)r + 2] = col.z;
          endrem[r] = best.rem;
          range[r] = best.t;
        } else if (zero_misses) {
#pragma unroll
          for (int k = 0; k < 3; ++k) { endpoints[3 * (size_t)r + k] = 0.f; endcolors[3 * (size_t)r + k] = 0; }
          endrem[r] = 0.f;
          range[r] = 0.f;
        }
        if (tri_id) tri_id[r] = best.pos >= 0 ? best.orig : -1;
        r = -1;
      }
    }
  }
}

int* g_debug_stats = nullptr;  // vl_debug_trace_stats(): per-ray {inner nodes visited, triangles tested}
int g_debug_mode = 0;          // vl_debug_trace_mode(): 0 auto, 1 per-ray storage order, 2 per-ray 16x8 tiles, 4/8/16/32 packet tile width

}  // namespace

extern "C" void vl_debug_trace_mode(int mode) { g_debug_mode = mode; }

extern "C" void vl_debug_trace_stats(int* d_stats) { g_debug_stats = d_stats; }

int vl_trace_launch(const void* d_blob, int n_faces, const float* d_rays, const float* d_origin, int n_rays,
                    int height, float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem,
                    int* d_tri_id, int flags, cudaStream_t stream) {
  const int width = n_rays / height;           // RayTracer.cpp:56
  const long long n_traced = (long long)width * height;
  if (d_tri_id && n_rays > n_traced) {         // rays beyond width*height are never cast
    VL_CUDA_CHECK(cudaMemsetAsync(d_tri_id + n_traced, 0xff, sizeof(int) * (size_t)(n_rays - n_traced), stream));
  }
  if (n_traced <= 0) return VL_OK;
  const char* blob = static_cast<const char*>(d_blob);
  VlBlobLayout L = vl_blob_layout(n_faces);
  const VlHeader* hdr = reinterpret_cast<const VlHeader*>(blob);
  const VlNode* nodes = reinterpret_cast<const VlNode*>(blob + L.off_nodes);
  const VlTri* tris = reinterpret_cast<const VlTri*>(blob + L.off_tris);
  const int4* c0 = reinterpret_cast<const int4*>(blob + L.off_c0);
  const bool zm = (flags & VL_TRACE_ZERO_MISSES) != 0, pn = (flags & VL_RAYS_NORMALIZED) != 0;
  int mode = g_debug_mode;
  // measured on B200 (gpurun_out/trace_stats_710.txt): per-thread stacks 0.180 ms, 8x4 packets 0.185 ms per
  // 131 072 rays over 1.05 M triangles -- packets test 1.65x the triangles, so per-thread is the default
  if (mode == 0) mode = (flags & VL_TRACE_PERSISTENT) ? 3 : ((flags & VL_TRACE_PACKET) ? ((height >= 4 && width >= 8) ? 8 : 32) : 2);
  VlProfScope ps(VL_ST_TRACE, stream);
  if (mode == 2 && !(height >= 4 && width >= 8)) mode = 1;
  if (mode == 3) {   // persistent warps + ray compaction + TMA-staged top of the tree (VL_TRACE_PERSISTENT)
    static unsigned int next_slot = 0;
    const int sm_count = vl_sm_count();
    VL_CUDA_CHECK(cudaFuncSetAttribute(k_trace_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPtSmemBytes));
    unsigned int* counters = nullptr;
    VL_CUDA_CHECK(cudaGetSymbolAddress(reinterpret_cast<void**>(&counters), g_pt_counters));
    unsigned int* counter = counters + (next_slot++ & 63u);
    VL_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    const VlNode* top = reinterpret_cast<const VlNode*>(blob + L.off_top);
    k_trace_persistent<<<sm_count, kPtThreads, kPtSmemBytes, stream>>>(hdr, nodes, top, tris, c0, d_rays, d_origin, (int)n_traced,
                                                                      width, height, d_endpoints, d_endcolors, d_range, d_endrem,
                                                                      d_tri_id, zm, pn, counter);
  } else if (mode == 1) {
    const int nb = (int)((n_traced + kTraceThreads - 1) / kTraceThreads);
    k_trace<false><<<nb, kTraceThreads, 0, stream>>>(hdr, nodes, tris, c0, d_rays, d_origin, (int)n_traced, width, height,
                                                    d_endpoints, d_endcolors, d_range, d_endrem, d_tri_id, zm, pn, g_debug_stats);
  } else if (mode == 2) {
    const int nb = ((width + 15) / 16) * ((height + 7) / 8);
    k_trace<true><<<nb, kTraceThreads, 0, stream>>>(hdr, nodes, tris, c0, d_rays, d_origin, (int)n_traced, width, height,
                                                   d_endpoints, d_endcolors, d_range, d_endrem, d_tri_id, zm, pn, g_debug_stats);
  } else {
    const int tw = mode, th = 32 / mode;
    const long long n_tiles = (long long)((width + tw - 1) / tw) * ((height + th - 1) / th);
    const int nb = (int)((n_tiles + kPktWarps - 1) / kPktWarps);
#define VL_PKT(TW_)                                                                                              \
  k_trace_packet<TW_><<<nb, kPktWarps * 32, 0, stream>>>(hdr, nodes, tris, c0, d_rays, d_origin, width, height,   \
                                                         d_endpoints, d_endcolors, d_range, d_endrem, d_tri_id,  \
                                                         zm, pn, g_debug_stats)
    if (tw == 8) VL_PKT(8); else if (tw == 16) VL_PKT(16); else if (tw == 4) VL_PKT(4); else VL_PKT(32);
#undef VL_PKT
  }
  VL_LAUNCH_CHECK("k_trace");
  return VL_OK;
}

int vl_trace_bruteforce_launch(const float* d_verts, const int* d_faces, const int* d_colors, const float* d_rem,
                               int n_verts, int n_faces, const float* d_rays, const float* d_origin, int n_rays,
                               int height, float* d_endpoints, int* d_endcolors, float* d_range,
                               float* d_endrem, int* d_tri_id, int flags, cudaStream_t stream) {
  const int width = n_rays / height;
  const long long n_traced = (long long)width * height;
  if (d_tri_id && n_rays > n_traced)
    VL_CUDA_CHECK(cudaMemsetAsync(d_tri_id + n_traced, 0xff, sizeof(int) * (size_t)(n_rays - n_traced), stream));
  if (n_traced <= 0) return VL_OK;
  const int nb = (int)((n_traced + kTraceThreads - 1) / kTraceThreads);
  k_trace_bruteforce<<<nb, kTraceThreads, 0, stream>>>(d_verts, d_faces, d_colors, d_rem, n_verts, n_faces, d_rays,
                                                      d_origin, (int)n_traced, d_endpoints, d_endcolors, d_range,
                                                      d_endrem, d_tri_id, (flags & VL_RAYS_NORMALIZED) != 0);
  VL_LAUNCH_CHECK("k_trace_bruteforce");
  return VL_OK;
}
