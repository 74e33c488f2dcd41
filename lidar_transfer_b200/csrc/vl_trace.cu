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
  if (mode == 0) mode = (flags & VL_TRACE_PACKET) ? ((height >= 4 && width >= 8) ? 8 : 32) : 2;
  VlProfScope ps(VL_ST_TRACE, stream);
  if (mode == 2 && !(height >= 4 && width >= 8)) mode = 1;
  if (mode == 1) {
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
