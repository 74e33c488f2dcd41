// vl_cast.cu -- (ii-b) scene-streaming closest-hit cast for single-origin ray sets, sm_100a.
//
// Replaces the same reference code as vl_bvh_build.cu + vl_trace.cu together (Triangle construction
// RayTracer.cpp:32-51, BVH::build BVH.cpp:143-243, the ray loop RayTracer.cpp:62-92, BVH::getIntersection
// BVH.cpp:19-110, Triangle::getIntersection Triangle.h:27-50) with the roles of the two sets swapped.
//
// Why: in this pipeline every scan has its own mesh (~1 M triangles) that is cast ONCE with ~131 k rays,
// and the ctrace ABI gives all rays ONE origin (RayTracer.cpp:116-124: `float* origin` is a single point).
// Building a hierarchy over the large, single-use set (the triangles) to query it with the small, constant
// set (the sensor's beams) is backwards on a bandwidth machine.  Here the BEAMS are indexed once per sensor
// (a direction-space cell grid: azimuth x sin(elevation)), and each scan's triangles are streamed through it
// exactly once: a triangle's central projection is a spherical triangle, its padded (azimuth, sine) bounding
// rectangle selects a few cells, the beams in those cells are tested with the reference's Moller-Trumbore
// arithmetic (vl_tri_hit, bit-identical to vl_trace.cu) and the closest hit per beam is kept with a 64-bit
// atomicMin on (t bits << 32 | face index).  No per-scan sort, no per-scan tree, no divergent traversal.
//
// Per scan: k_cast_init resets the per-beam keys; k_cast_setup streams the faces in batches of
// 512 per CTA -- a cheap cull (sine interval vs the beam rows) that ~70 % of a LiDAR scene's triangles do not
// survive, then on dense warps the full rectangle, a 64-byte record and one work unit per run of <= 8 cells of a
// cell row; k_cast_units pools the beams of 32 units per warp and tests them 32 at a time; k_cast_resolve writes the
// outputs of RayTracer.cpp:73-90 for the winning triangle of each beam.  Records and units live in L2 between the
// two kernels.  The four launches of a stream slot can be captured once and replayed per scan (vl_cast_graph_*).
//
// Result contract (same as vl_trace.cu): the closest hit over ALL triangles under the
// reference arithmetic; exact-t ties go to the smaller face index (that is what the packed key orders by).
// The cell rectangle may only over-select: every bound below is conservative (padded by kPad0 plus the
// rounding of v - o), so a beam the triangle test would accept is always among the candidates.
#include <stdlib.h>
#include "vl_common.cuh"

#ifndef VL_SETUP_MINB
#define VL_SETUP_MINB 4
#endif
#define VL_SETUP_MINB_DEFAULT VL_SETUP_MINB

namespace {

constexpr int kCastThreads = 256;
constexpr int kCastWarps = kCastThreads / 32;
constexpr int kFineBins = 4096;     // fine sin(elevation) bins of the next-beam-sine table (the arithmetic early-out)
#ifndef VL_SEG_SHIFT
#define VL_SEG_SHIFT 3
#endif
constexpr int kSegShift = VL_SEG_SHIFT;   // an item is one run of <= 8 cells of one cell row ...
constexpr int kSegShiftWide = 6;    // ... or of <= 64 cells for a triangle wider than kWideCols (bounds the unit count)
constexpr int kWideCols = 128;
constexpr int kUnitItems = 1;       // items per work unit (one lane of k_cast_units)
#ifndef VL_SETUP_FPT
#define VL_SETUP_FPT 2   // measured (8 scans in flight): 1 / 2 / 4 / 8 faces per thread and batch -> 41.0 / 36.0 / 36.9 / 37.6 us per scan
#endif
constexpr int kBatch = VL_SETUP_FPT * kCastThreads;   // faces per culling batch of k_cast_setup
constexpr int kUnitBits = 36;       // packed reservation counter: [records : 28][units : 36]
constexpr float kPad0 = 2e-5f;      // angular slack (rad): >> fp32 rounding of azimuth / sine / Moller-Trumbore edges

struct VlBeamHeader {
  unsigned int sine_min_ord, sine_max_ord;  // order-preserving uint encodings, reduced by k_beam_prep
  int n_binned;                             // rays with a finite direction (the others can never hit)
  int pad[61];
};
static_assert(sizeof(VlBeamHeader) == 256, "beam header is 256 B");

struct VlCastHeader {
  int n_bad_faces;
  int overflow;                  // more work units than the workspace holds: results invalid (VL_ENOSPACE)
  unsigned long long reserved;   // [records : 28][units : 36], bumped once per pass by k_cast_setup
  // the three counters above as k_cast_resolve found them: what the host reads (vl_cast_status, the 16-byte status copy)
  int snap_bad_faces;
  int snap_overflow;
  unsigned long long snap_reserved;
  int pad[56];
};
constexpr size_t kSnapOffset = 16;
static_assert(sizeof(VlCastHeader) == 256, "cast header is 256 B");

// The mesh of one scan.  Passed to the kernels by value (vl_cast), or read from device memory (a replayed CUDA
// graph, vl_cast_graph_*: the graph's first node copies it from a pinned host copy the caller rewrites per scan).
struct VlMeshDesc {
  const float* verts;
  const int* faces;
  const int* colors;
  const float* rem;
  int n_verts, n_faces;
  long long pad[3];
};
static_assert(sizeof(VlMeshDesc) == 64, "mesh descriptor is 64 B");
constexpr size_t kDescOffset = 128;   // device copy of the descriptor inside the 256-byte workspace header (the counters use the first 24)

struct BeamLayout {
  int n;        // rays cast = width * height (RayTracer.cpp:56)
  int cw, ch;   // direction cells: yaw x sine
  size_t off_dir, off_sorted, off_slot_of, off_cell_start, off_cursor, off_fine, off_mask, off_blk, off_rowlim, total;
};

int g_cells_per_row = 1;   // cell rows per beam row (vl_debug_cast_cells)
// vl_debug_cast_rearm: 1 = a slot's graph has no reset kernel (k_cast_resolve re-arms the keys and counters it has just
// read), 0 (default) = k_cast_init per scan.  Measured (profiles/r02_experiments.md): WITHOUT the 3.6 us reset kernel the
// step is 15 % slower (2720 vs 3136 Mrays/s at 8 streams, A/B on one box) -- kept as a switch, off.
int g_graph_rearm = 0;
int g_items_ctas_per_sm = 4;
// vl_debug_cast_split / VLIDAR_CAST_SPLIT: 1 = cull and setup as two kernels (k_cast_cull + k_cast_setup2)
int g_cast_split = getenv("VLIDAR_CAST_SPLIT") ? atoi(getenv("VLIDAR_CAST_SPLIT")) : 0;
int g_row_trim = 1;          // vl_debug_cast_row_trim: 0 = rectangles keep every cell row their sine interval touches (A/B aid)
int g_setup_ctas_per_sm = VL_SETUP_MINB_DEFAULT;

BeamLayout beam_layout(int n_rays, int height) {
  BeamLayout L;
  const int width = height > 0 ? n_rays / height : 0;
  const long long n = (long long)width * height;
  L.n = (int)n;
  int cw = width < 1 ? 1 : (width > 4096 ? 4096 : width);
  long long ch = (long long)g_cells_per_row * height;
  if ((long long)cw * ch < n) ch = (n + cw - 1) / cw;
  if (ch < 1) ch = 1;
  if (ch > 4096) ch = 4096;
  L.cw = cw; L.ch = (int)ch;
  const size_t nn = n > 0 ? (size_t)n : 1, ncell = (size_t)L.cw * L.ch;
  size_t off = 256;
  L.off_dir = off;        off = vl_align256(off + 16 * nn);
  L.off_sorted = off;     off = vl_align256(off + 16 * nn);
  L.off_slot_of = off;    off = vl_align256(off + 4 * nn);
  L.off_cell_start = off; off = vl_align256(off + 4 * (ncell + 1));
  L.off_cursor = off;     off = vl_align256(off + 4 * ncell);
  L.off_fine = off;       off = vl_align256(off + 4 * kFineBins);
  L.off_mask = off;       off = vl_align256(off + 4 * kFineBins);   // next-beam-sine table (floats), see beam_in
  L.off_blk = off;        off = vl_align256(off + 4 * (ncell / 4096 + 1));
  L.off_rowlim = off;     off = vl_align256(off + 8 * (size_t)L.ch);   // per cell row: smallest / largest sine of its beams (ordered uints)
  L.total = off;
  return L;
}

struct BeamParams {
  int cw, ch;
  float cw_inv;           // cw / 4 (cells per unit of pseudo-angle)
  float lo, hi;           // sine range of the binned rays
  float ch_inv, nf_inv;   // cells / fine bins per unit sine (0 when the range is empty)
};

// the same expressions on the ray side (k_beam_count / k_beam_scatter) and the triangle side (tri_setup)
__device__ __forceinline__ BeamParams beam_params(const VlBeamHeader* hdr, int cw, int ch) {
  BeamParams P;
  P.cw = cw; P.ch = ch;
  P.cw_inv = (float)cw * 0.25f;
  P.lo = vl_ordered_to_float(hdr->sine_min_ord);
  P.hi = vl_ordered_to_float(hdr->sine_max_ord);
  const float span = P.hi - P.lo;
  const bool ok = span > 1e-12f;   // false also for the empty set (lo, hi = NaN)
  P.ch_inv = ok ? (float)ch / span : 0.f;
  P.nf_inv = ok ? (float)kFineBins / span : 0.f;
  return P;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int row_of(float s, const BeamParams& P) { return clampi((int)floorf((s - P.lo) * P.ch_inv), 0, P.ch - 1); }
__device__ __forceinline__ int fine_of(float s, const BeamParams& P) { return clampi((int)floorf((s - P.lo) * P.nf_inv), 0, kFineBins - 1); }
// Azimuth as a pseudo-angle ("diamond angle") in [-2, 2]: monotonic in atan2(y, x), same wrap point (x < 0, y = 0),
// quarter turns at -1, 0, 1, one division instead of an arctangent.  d(pseudo)/d(yaw) lies in [0.5, 1] per radian,
// so an angular pad in radians is a valid pad in pseudo-angle units, and "does not fit in a half circle" is
// "extent >= 2" exactly (pseudo(yaw + pi) = pseudo(yaw) + 2).
__device__ __forceinline__ float pseudo_yaw(float y, float x) {
  const float m = fabsf(x) + fabsf(y);
  const float t = m > 0.f ? __fdividef(y, m) : 0.f;   // a direction along the z axis has no azimuth: any value will do
  return x >= 0.f ? t : (y >= 0.f ? 2.f - t : -2.f - t);
}
__device__ __forceinline__ float wrap_2(float x) { return x - 4.f * rintf(x * 0.25f); }

// Is there a beam whose sine lies in [slo, shi]?  nxt[b] = the smallest beam sine among the fine bins >= b (+inf when
// there is none).  A beam with z >= slo sits in a bin >= fine_of(slo) (fine_of is monotonic), so z >= nxt[fine_of(slo)]:
// if any beam lies in [slo, shi] then nxt[fine_of(slo)] <= shi.  The converse fails only when the smallest sine of bin
// fine_of(slo) itself lies below slo -- "maybe" is always a valid answer.  One bin index, one load, one comparison.
__device__ __forceinline__ bool beam_in(const float* __restrict__ nxt, float slo, float shi, const BeamParams& P) {
  return __ldg(nxt + fine_of(slo, P)) <= shi;
}

// ---------------------------------------------------------------------------
// beam index (once per sensor / ray set)
// ---------------------------------------------------------------------------
__global__ void k_beam_init(VlBeamHeader* hdr, int* cell_cnt, int ncell_p1, int* fine, uint2* row_lim, int ch) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncell_p1; i += stride) cell_cnt[i] = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ch; i += stride) row_lim[i] = make_uint2(0xffffffffu, 0u);   // empty row
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kFineBins; i += stride) fine[i] = (int)0xffffffffu;   // smallest sine of the bin (ordered), none yet
  if (blockIdx.x == 0 && threadIdx.x == 0) { hdr->sine_min_ord = 0xffffffffu; hdr->sine_max_ord = 0u; hdr->n_binned = 0; }
}

// normalised direction (the one vl_tri_hit is given: vl_normalize = Vector3.h:73-89 with IEEE 1/sqrt, or the caller's
// own unit vectors -- vl_normalize_rays, the reference's rsqrtps + Newton step evaluated on the host) + azimuth
__global__ void __launch_bounds__(kCastThreads)
k_beam_prep(const float* __restrict__ rays, int n, bool prenorm, float4* __restrict__ dir, VlBeamHeader* hdr) {
  __shared__ float s_min[kCastWarps], s_max[kCastWarps];
  __shared__ int s_cnt[kCastWarps];
  const int r = blockIdx.x * kCastThreads + threadIdx.x;
  float smin = INFINITY, smax = -INFINITY;
  int cnt = 0;
  if (r < n) {
    const float3 d = vl_ray_dir(rays, (size_t)r, prenorm);
    float yaw = pseudo_yaw(d.y, d.x);
    const bool ok = isfinite(d.x) && isfinite(d.y) && isfinite(d.z) && isfinite(yaw);
    if (!ok) yaw = __int_as_float(0x7fc00000);
    dir[r] = make_float4(d.x, d.y, d.z, yaw);
    if (ok) { smin = smax = d.z; cnt = 1; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
    smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_min[w] = smin; s_max[w] = smax; s_cnt[w] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 1; k < kCastWarps; ++k) { smin = fminf(smin, s_min[k]); smax = fmaxf(smax, s_max[k]); cnt += s_cnt[k]; }
    if (cnt > 0) {
      atomicMin(&hdr->sine_min_ord, vl_float_to_ordered(smin));
      atomicMax(&hdr->sine_max_ord, vl_float_to_ordered(smax));
      atomicAdd(&hdr->n_binned, cnt);
    }
  }
}

__device__ __forceinline__ int ray_cell(const float4 d, const BeamParams& P) {
  int col = (int)floorf((d.w + 2.f) * P.cw_inv);
  if (col >= P.cw) col -= P.cw;   // azimuth +2 is the same direction as -2
  col = clampi(col, 0, P.cw - 1);
  return row_of(d.z, P) * P.cw + col;
}

__global__ void __launch_bounds__(kCastThreads)
k_beam_count(const float4* __restrict__ dir, int n, const VlBeamHeader* __restrict__ hdr, int cw, int ch,
             int* __restrict__ cell_cnt, int* __restrict__ fine, uint2* __restrict__ row_lim) {
  const int r = blockIdx.x * kCastThreads + threadIdx.x;
  if (r >= n) return;
  const float4 d = dir[r];
  if (!(d.w == d.w)) return;
  const BeamParams P = beam_params(hdr, cw, ch);
  atomicAdd(&cell_cnt[ray_cell(d, P)], 1);
  {   // sine range of the beams of this cell row (tri_setup trims the rows at both ends of a rectangle with it)
    const int row = row_of(d.z, P);
    const unsigned int ord = vl_float_to_ordered(d.z);
    atomicMin(&row_lim[row].x, ord);
    atomicMax(&row_lim[row].y, ord);
  }
  atomicMin(reinterpret_cast<unsigned int*>(fine) + fine_of(d.z, P), vl_float_to_ordered(d.z));
}

// exclusive prefix sums of the cell counts in three coalesced steps (local 4096-element scans, scan of the block
// totals, apply) -- in place, plus a copy as the scatter cursor; the middle step also folds the fine bins into
// their occupancy bitmap
constexpr int kScanBlock = 4096;

__device__ __forceinline__ int block_excl_1024(int v, int* s_warp, int* s_total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
  __syncthreads();
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    const int x = s_warp[lane];
    int xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= d) xi += u; }
    s_warp[lane] = xi - x;
    if (lane == 31) *s_total = xi;
  }
  __syncthreads();
  return s_warp[w] + incl - v;
}

__global__ void __launch_bounds__(1024)
k_beam_scan_local(int* __restrict__ cell, int ncell, int* __restrict__ blk_sum) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int base = blockIdx.x * kScanBlock + threadIdx.x * 4;
  int v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = base + k < ncell ? cell[base + k] : 0;
  int run = block_excl_1024(v[0] + v[1] + v[2] + v[3], s_warp, &s_total);
#pragma unroll
  for (int k = 0; k < 4; ++k) { if (base + k < ncell) cell[base + k] = run; run += v[k]; }
  if (threadIdx.x == 0) blk_sum[blockIdx.x] = s_total;
}

__global__ void __launch_bounds__(1024)
k_beam_scan_top(int* __restrict__ blk_sum, int nblk, int* __restrict__ cell, int ncell, const int* __restrict__ fine,
                float* __restrict__ nxt, uint2* __restrict__ row_lim, int ch) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  __shared__ unsigned int s_min[kFineBins];
  const int per = (nblk + 1023) / 1024;
  const int b = min((int)threadIdx.x * per, nblk), e = min(b + per, nblk);
  int sum = 0;
  for (int i = b; i < e; ++i) sum += blk_sum[i];
  int run = block_excl_1024(sum, s_warp, &s_total);
  for (int i = b; i < e; ++i) { const int c = blk_sum[i]; blk_sum[i] = run; run += c; }
  if (threadIdx.x == 0) cell[ncell] = s_total;
  // next-beam-sine table: suffix minimum of the per-bin smallest sines (ordered uints order like the floats)
  for (int bin = threadIdx.x; bin < kFineBins; bin += 1024) s_min[bin] = (unsigned int)fine[bin];
  __syncthreads();
  for (int d = 1; d < kFineBins; d <<= 1) {
    unsigned int v[kFineBins / 1024];
#pragma unroll
    for (int k = 0; k < kFineBins / 1024; ++k) {
      const int bin = threadIdx.x + 1024 * k;
      v[k] = bin + d < kFineBins ? min(s_min[bin], s_min[bin + d]) : s_min[bin];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kFineBins / 1024; ++k) s_min[threadIdx.x + 1024 * k] = v[k];
    __syncthreads();
  }
  for (int bin = threadIdx.x; bin < kFineBins; bin += 1024)
    nxt[bin] = s_min[bin] == 0xffffffffu ? INFINITY : vl_ordered_to_float(s_min[bin]);
  // per cell row: the sine limits as floats, NaN for a row without beams (every comparison with them fails)
  for (int r = threadIdx.x; r < ch; r += 1024) {
    const uint2 l = row_lim[r];
    const bool any = l.x <= l.y;
    const float lo = any ? vl_ordered_to_float(l.x) : __int_as_float(0x7fc00000), hi = any ? vl_ordered_to_float(l.y) : __int_as_float(0x7fc00000);
    row_lim[r] = make_uint2(__float_as_uint(lo), __float_as_uint(hi));
  }
}

__global__ void __launch_bounds__(1024)
k_beam_scan_apply(int* __restrict__ cell, int ncell, const int* __restrict__ blk_sum, int* __restrict__ cursor) {
  const int i = blockIdx.x * 1024 + threadIdx.x;
  if (i >= ncell) return;
  const int v = cell[i] + blk_sum[i / kScanBlock];
  cell[i] = v;
  cursor[i] = v;
}

// sorted[slot] = direction + azimuth of the ray that landed in `slot` (cell order); slot_of[r] = its slot, -1 for a
// ray without a finite direction.  The per-scan hit keys are kept per SLOT, so the cast never needs a ray index.
__global__ void __launch_bounds__(kCastThreads)
k_beam_scatter(const float4* __restrict__ dir, int n, const VlBeamHeader* __restrict__ hdr, int cw, int ch,
               int* __restrict__ cursor, float4* __restrict__ sorted, int* __restrict__ slot_of) {
  const int r = blockIdx.x * kCastThreads + threadIdx.x;
  if (r >= n) return;
  const float4 d = dir[r];
  if (!(d.w == d.w)) { slot_of[r] = -1; return; }
  const BeamParams P = beam_params(hdr, cw, ch);
  const int slot = atomicAdd(&cursor[ray_cell(d, P)], 1);
  sorted[slot] = d;
  slot_of[r] = slot;
}

// ---------------------------------------------------------------------------
// per-scan cast
// ---------------------------------------------------------------------------
struct TriRec {
  float v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z;
  int orig;
  float ymid, yhalf;   // yaw interval centre / half width incl. padding; yhalf < 0 = every yaw
  float slo, shi;      // padded sine interval
  int ca, ncx, ra, ncy;
};

// Conservative (azimuth, sine) rectangle of a triangle as seen from o, and the cells it overlaps.  The central
// projection of a planar triangle is a spherical triangle with great-circle edges:
//   * azimuth is monotonic along an edge, so the azimuth interval is spanned by the three vertex azimuths unless
//     the z axis pierces the triangle -- exactly when they do not fit in a half circle;
//   * sin(elevation) has no interior extremum on the face except at the poles, and along an edge of arc
//     length L it exceeds its end values by at most L^2 / 8 (|d2/dphi2 sin e| <= 1 on a great circle).
// kFull = false: the cheap part only (sine interval against the vertical field of view and the beam rows), returns
// 1 when the triangle survives; kFull = true: the whole rectangle, returns the number of items (runs of
// cells of one cell row; 0 = no beam can hit it).
template <bool kFull>
__device__ __forceinline__ int tri_setup(int f, int i0, int i1, int i2, const float* __restrict__ verts, const float3 o,
                                         const BeamParams& P, const float* __restrict__ fine_mask,
                                         const float2* __restrict__ row_lim, TriRec& T) {
  const float ax = __ldg(verts + 3 * (size_t)i0), ay = __ldg(verts + 3 * (size_t)i0 + 1), az = __ldg(verts + 3 * (size_t)i0 + 2);
  const float bx = __ldg(verts + 3 * (size_t)i1), by = __ldg(verts + 3 * (size_t)i1 + 1), bz = __ldg(verts + 3 * (size_t)i1 + 2);
  const float cx = __ldg(verts + 3 * (size_t)i2), cy = __ldg(verts + 3 * (size_t)i2 + 1), cz = __ldg(verts + 3 * (size_t)i2 + 2);
  const float p0x = ax - o.x, p0y = ay - o.y, p0z = az - o.z;
  const float p1x = bx - o.x, p1y = by - o.y, p1z = bz - o.z;
  const float p2x = cx - o.x, p2y = cy - o.y, p2z = cz - o.z;
  // a non-finite vertex makes Moller-Trumbore's determinant non-finite: t is then 0, inf or NaN, never accepted
  // (x * 0 == 0 exactly for finite x only)
  if (!(((p0x * 0.f + p0y * 0.f + p0z * 0.f) + (p1x * 0.f + p1y * 0.f + p1z * 0.f) + (p2x * 0.f + p2y * 0.f + p2z * 0.f)) == 0.f)) return 0;
  const float q0 = p0x * p0x + p0y * p0y + p0z * p0z, q1 = p1x * p1x + p1y * p1y + p1z * p1z, q2 = p2x * p2x + p2y * p2y + p2z * p2z;
  float slo, shi;
  bool all_yaw;
  float y0 = 0.f, lo_d = 0.f, hi_d = 0.f, pad_y = 0.f;
  const float qmin = fminf(q0, fminf(q1, q2)), qmax = fmaxf(q0, fmaxf(q1, q2));
  if (!(qmin > 1e-30f && qmax < 1e30f)) {
    if (!kFull) return 1;
    slo = -2.f; shi = 2.f; all_yaw = true;   // a vertex at the origin / out of float range: every beam is a candidate
  } else {
    const float r0 = rsqrtf(q0), r1 = rsqrtf(q1), r2 = rsqrtf(q2);
    // absolute rounding of p = v - o (2 ulp of the largest coordinate involved), as an angle at each vertex
    const float amax = fmaxf(fmaxf(fmaxf(fabsf(ax), fabsf(ay)), fmaxf(fabsf(az), fabsf(bx))),
                             fmaxf(fmaxf(fmaxf(fabsf(by), fabsf(bz)), fmaxf(fabsf(cx), fabsf(cy))),
                                   fmaxf(fmaxf(fabsf(cz), fabsf(o.x)), fmaxf(fabsf(o.y), fabsf(o.z)))));
    const float E = amax * 2.4e-7f;
    const float u0x = p0x * r0, u0y = p0y * r0, u0z = p0z * r0;
    const float u1x = p1x * r1, u1y = p1y * r1, u1z = p1z * r1;
    const float u2x = p2x * r2, u2y = p2y * r2, u2z = p2z * r2;
    const float c01 = (u0x - u1x) * (u0x - u1x) + (u0y - u1y) * (u0y - u1y) + (u0z - u1z) * (u0z - u1z);
    const float c12 = (u1x - u2x) * (u1x - u2x) + (u1y - u2y) * (u1y - u2y) + (u1z - u2z) * (u1z - u2z);
    const float c20 = (u2x - u0x) * (u2x - u0x) + (u2y - u0y) * (u2y - u0y) + (u2z - u0z) * (u2z - u0z);
    // arc L <= (pi/2) chord  =>  L^2 / 8 <= 0.3085 chord^2
    const float bulge = 0.31f * fmaxf(c01, fmaxf(c12, c20));
    const float pad_s = kPad0 + 2.f * E * fmaxf(r0, fmaxf(r1, r2));
    slo = fminf(u0z, fminf(u1z, u2z)) - bulge - pad_s;
    shi = fmaxf(u0z, fmaxf(u1z, u2z)) + bulge + pad_s;
    // reject on the vertical field of view and on the beam rows before any azimuth work.  The interval so far is
    // valid unless the z axis pierces the triangle, which needs the origin inside its xy bounding box.
    if (shi < P.lo || slo > P.hi) return 0;
    const float e2 = 2.f * E;
    const bool near_axis = fminf(p0x, fminf(p1x, p2x)) <= e2 && fmaxf(p0x, fmaxf(p1x, p2x)) >= -e2 &&
                           fminf(p0y, fminf(p1y, p2y)) <= e2 && fmaxf(p0y, fmaxf(p1y, p2y)) >= -e2;
    if (!near_axis && !beam_in(fine_mask, slo, shi, P)) return 0;
    if (!kFull) return 1;
    const float h0 = p0x * p0x + p0y * p0y, h1 = p1x * p1x + p1y * p1y, h2 = p2x * p2x + p2y * p2y;
    const float hmin = fminf(h0, fminf(h1, h2));
    all_yaw = !(hmin > 1e-30f);
    if (!all_yaw) {
      pad_y = kPad0 + 2.f * E * rsqrtf(hmin);
      y0 = pseudo_yaw(p0y, p0x);
      const float d1 = wrap_2(pseudo_yaw(p1y, p1x) - y0), d2 = wrap_2(pseudo_yaw(p2y, p2x) - y0);
      lo_d = fminf(0.f, fminf(d1, d2));
      hi_d = fmaxf(0.f, fmaxf(d1, d2));
      all_yaw = !((hi_d - lo_d) + 2.f * pad_y < 2.f - 1e-3f);
    }
    if (all_yaw && near_axis) {   // the z axis may pierce the triangle: the elevation reaches the pole on that side
      if (fmaxf(p0z, fmaxf(p1z, p2z)) > 0.f) shi = 2.f;
      if (fminf(p0z, fminf(p1z, p2z)) < 0.f) slo = -2.f;
    }
  }
  if (shi < P.lo || slo > P.hi) return 0;
  if (!beam_in(fine_mask, slo, shi, P)) return 0;   // no beam row inside the sine interval
  T.v0x = ax; T.v0y = ay; T.v0z = az;
  T.e1x = __fsub_rn(bx, ax); T.e1y = __fsub_rn(by, ay); T.e1z = __fsub_rn(bz, az);
  T.e2x = __fsub_rn(cx, ax); T.e2y = __fsub_rn(cy, ay); T.e2z = __fsub_rn(cz, az);
  T.orig = f;
  T.slo = slo; T.shi = shi;
  // cell rows of the sine interval, trimmed at both ends by rows none of whose beams lies inside it (the same
  // comparison on the same values as the per-beam filter of k_cast_units: an exact saving, half the candidates of a
  // typical LiDAR triangle, whose interval is narrower than a cell row and straddles a row boundary)
  int ra = row_of(slo, P), rb = row_of(shi, P);
  if (row_lim) {
    // float comparisons, as the filter's (an empty row holds NaN limits and fails both)
    for (; ra <= rb; ++ra) { const float2 l = __ldg(row_lim + ra); if (l.y >= slo && l.x <= shi) break; }
    for (; rb > ra; --rb) { const float2 l = __ldg(row_lim + rb); if (l.y >= slo && l.x <= shi) break; }
    if (ra > rb) return 0;
  }
  T.ra = ra;
  T.ncy = rb - ra + 1;
  if (all_yaw) {
    T.ca = 0; T.ncx = P.cw; T.ymid = 0.f; T.yhalf = -1.f;
  } else {
    const int ca = (int)floorf((y0 + lo_d - pad_y + 2.f) * P.cw_inv), cb = (int)floorf((y0 + hi_d + pad_y + 2.f) * P.cw_inv);
    const int ncx = cb - ca + 1;
    if (ncx >= P.cw) {
      T.ca = 0; T.ncx = P.cw; T.ymid = 0.f; T.yhalf = -1.f;
    } else {
      // y0 + lo_d - pad_y > -2 - 2 - pad: ca >= -cw - 1, one or two conditional wraps instead of a modulo
      int c = ca;
      if (c < 0) c += P.cw;
      if (c < 0) c += P.cw;
      if (c >= P.cw) c -= P.cw;
      T.ca = c; T.ncx = ncx;
      T.ymid = y0 + 0.5f * (lo_d + hi_d);
      T.yhalf = 0.5f * (hi_d - lo_d) + pad_y;
    }
  }
  const int sh = T.ncx > kWideCols ? kSegShiftWide : kSegShift;
  return T.ncy * ((T.ncx + (1 << sh) - 1) >> sh);
}

// The cull of k_cast_setup: the sine interval of tri_setup with cheaper, looser bounds (every replacement only widens
// the interval): chord between unit vectors <= |pa - pb| / min(|pa|, |pb|) instead of normalising the vertices,
// |v| <= |o| + |p| for the rounding term, one MUFU per inverse length.  Returns false only for a triangle that no
// beam can hit; anything unusual (non-finite, at the origin, out of float range, around the z axis) survives and
// is decided by tri_setup<true>.
__device__ __forceinline__ float rsqrt_mufu(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ bool tri_cull(int i0, int i1, int i2, const float* __restrict__ verts, const float3 o,
                                         float o_max, const BeamParams& P, const float* __restrict__ fine_mask) {
  const float p0x = __ldg(verts + 3 * (size_t)i0) - o.x, p0y = __ldg(verts + 3 * (size_t)i0 + 1) - o.y, p0z = __ldg(verts + 3 * (size_t)i0 + 2) - o.z;
  const float p1x = __ldg(verts + 3 * (size_t)i1) - o.x, p1y = __ldg(verts + 3 * (size_t)i1 + 1) - o.y, p1z = __ldg(verts + 3 * (size_t)i1 + 2) - o.z;
  const float p2x = __ldg(verts + 3 * (size_t)i2) - o.x, p2y = __ldg(verts + 3 * (size_t)i2 + 1) - o.y, p2z = __ldg(verts + 3 * (size_t)i2 + 2) - o.z;
  const float q0 = p0x * p0x + p0y * p0y + p0z * p0z, q1 = p1x * p1x + p1y * p1y + p1z * p1z, q2 = p2x * p2x + p2y * p2y + p2z * p2z;
  const float qmin = fminf(q0, fminf(q1, q2)), qmax = fmaxf(q0, fmaxf(q1, q2));
  if (!((q0 + q1) + q2 < 1e30f) || !(qmin > 1e-30f)) return true;   // NaN / inf / huge / at the origin: not decided here
  const float r0 = rsqrt_mufu(q0), r1 = rsqrt_mufu(q1), r2 = rsqrt_mufu(q2);
  const float rmax = fmaxf(r0, fmaxf(r1, r2)), rmin = fminf(r0, fminf(r1, r2));
  const float s0 = p0z * r0, s1 = p1z * r1, s2 = p2z * r2;
  const float d01 = (p0x - p1x) * (p0x - p1x) + (p0y - p1y) * (p0y - p1y) + (p0z - p1z) * (p0z - p1z);
  const float d12 = (p1x - p2x) * (p1x - p2x) + (p1y - p2y) * (p1y - p2y) + (p1z - p2z) * (p1z - p2z);
  const float d20 = (p2x - p0x) * (p2x - p0x) + (p2y - p0y) * (p2y - p0y) + (p2z - p0z) * (p2z - p0z);
  const float bulge = 0.31f * fmaxf(d01, fmaxf(d12, d20)) * rmax * rmax;
  const float E = 2.4e-7f * (o_max + qmax * rmin);   // every |coordinate| <= |o|_inf + |p|_2
  const float pad_s = kPad0 + 2.f * E * rmax;
  const float slo = fminf(s0, fminf(s1, s2)) - bulge - pad_s, shi = fmaxf(s0, fmaxf(s1, s2)) + bulge + pad_s;
  if (shi < P.lo || slo > P.hi) return false;
  const float e2 = 2.f * E;
  const bool near_axis = fminf(p0x, fminf(p1x, p2x)) <= e2 && fmaxf(p0x, fmaxf(p1x, p2x)) >= -e2 &&
                         fminf(p0y, fminf(p1y, p2y)) <= e2 && fmaxf(p0y, fmaxf(p1y, p2y)) >= -e2;
  return near_axis || beam_in(fine_mask, slo, shi, P);
}

__device__ __forceinline__ unsigned long long init_key() {
  return ((unsigned long long)__float_as_uint(999999999.f) << 32) | 0x7fffffffull;   // BVH.cpp:20
}

__global__ void k_cast_init(unsigned long long* __restrict__ best, int n, VlCastHeader* hdr) {
  const int stride = gridDim.x * blockDim.x;
  const unsigned long long k = init_key();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) best[i] = k;
  if (blockIdx.x == 0 && threadIdx.x == 0) { hdr->n_bad_faces = 0; hdr->overflow = 0; hdr->reserved = 0ull; }
}

// Step 1: stream the faces once, in batches of 512 per CTA.
//   cull:  every face gets the cheap test (sine interval vs the beam rows); ~70 % of a LiDAR scene's triangles lie
//          between two beam rows or outside the vertical field of view and stop here.  Survivors are queued in
//          shared memory, so that
//   setup: runs on dense warps: the full rectangle, a 64-byte record (edges for the triangle test + rectangle)
//          and ceil(items / 4) work units (record, first item) appended to global lists -- one packed atomicAdd
//          per pass reserves both.  List order is irrelevant: a unit is self-contained.
template <bool kDescPtr>
__global__ void __launch_bounds__(kCastThreads, VL_SETUP_MINB)
k_cast_setup(const VlBeamHeader* __restrict__ bhdr, int cw, int ch, const float* __restrict__ s_mask,
             const VlMeshDesc mesh_val, const VlMeshDesc* __restrict__ mesh_ptr,
             const float* __restrict__ origin, VlCastHeader* chdr, float4* __restrict__ recs, int rec_cap,
             int2* __restrict__ units, unsigned long long unit_cap, const float2* __restrict__ row_lim) {
  const float* __restrict__ verts = kDescPtr ? mesh_ptr->verts : mesh_val.verts;
  const int* __restrict__ faces = kDescPtr ? mesh_ptr->faces : mesh_val.faces;
  const int n_verts = kDescPtr ? mesh_ptr->n_verts : mesh_val.n_verts;
  const int n_faces = kDescPtr ? mesh_ptr->n_faces : mesh_val.n_faces;
  // ONE barrier per batch: the survivor queue is double-buffered (batch b + 2 overwrites the queue of batch b only after
  // the barrier of batch b + 1, which every thread reaches after its share of batch b's setup), the two counters are
  // triple-buffered (the set of batch b + 2 is cleared by thread 0 right after the barrier of batch b, i.e. before it
  // arrives at the barrier of batch b + 1 that precedes every use)
  __shared__ int s_queue[2][kBatch];
  __shared__ int s_nq[3], s_next[3];
  if (threadIdx.x < 3) { s_nq[threadIdx.x] = 0; s_next[threadIdx.x] = 0; }
  __syncthreads();
  const BeamParams P = beam_params(bhdr, cw, ch);
  const float3 o = make_float3(__ldg(origin), __ldg(origin + 1), __ldg(origin + 2));
  const float o_max = fmaxf(fabsf(o.x), fmaxf(fabsf(o.y), fabsf(o.z)));
  const int lane = threadIdx.x & 31;
  const int n_batches = (n_faces + kBatch - 1) / kBatch;
  const unsigned long long units_mask = (1ull << kUnitBits) - 1ull;
  int n_bad = 0;
  int it = 0;
  for (int batch = blockIdx.x; batch < n_batches; batch += gridDim.x, ++it) {
    const int cb = it % 3, qb = it & 1;
#ifndef VL_NO_SETUP_PREFETCH
    {   // the faces (a soup: the vertices) of this CTA's NEXT batch start their way from DRAM to L2 now: one line per thread
      const int nb = batch + gridDim.x;
      if (nb < n_batches) {
        const size_t per = faces ? 12 : 36;
        const char* p = faces ? reinterpret_cast<const char*>(faces) + (size_t)nb * kBatch * 12
                              : reinterpret_cast<const char*>(verts) + (size_t)nb * kBatch * 36;
        const size_t bytes = (size_t)min(kBatch, n_faces - nb * kBatch) * per;
        for (size_t o = (size_t)threadIdx.x * 128; o < bytes; o += (size_t)kCastThreads * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
      }
    }
#endif
    // ---- cull
    int idx[kBatch / kCastThreads][3];
#pragma unroll
    for (int k = 0; k < kBatch / kCastThreads; ++k) {   // all index loads first: faces -> verts is a dependent gather
      const int f = batch * kBatch + k * kCastThreads + threadIdx.x;
      if (f < n_faces) {
        if (faces) { idx[k][0] = __ldg(faces + 3 * (size_t)f); idx[k][1] = __ldg(faces + 3 * (size_t)f + 1); idx[k][2] = __ldg(faces + 3 * (size_t)f + 2); }
        else { idx[k][0] = 3 * f; idx[k][1] = 3 * f + 1; idx[k][2] = 3 * f + 2; }   // triangle soup: faces = (3f, 3f+1, 3f+2), no index array
      }
    }
#pragma unroll
    for (int k = 0; k < kBatch / kCastThreads; ++k) {
      const int f = batch * kBatch + k * kCastThreads + threadIdx.x;
      bool keep = false;
      if (f < n_faces) {
        if ((unsigned)idx[k][0] >= (unsigned)n_verts || (unsigned)idx[k][1] >= (unsigned)n_verts || (unsigned)idx[k][2] >= (unsigned)n_verts) {
          ++n_bad;
        } else {
          keep = tri_cull(idx[k][0], idx[k][1], idx[k][2], verts, o, o_max, P, s_mask);
        }
      }
      const unsigned int m = __ballot_sync(0xffffffffu, keep);
      if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_nq[cb], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) s_queue[qb][base + __popc(m & ((1u << lane) - 1u))] = f;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { const int rb = (it + 2) % 3; s_nq[rb] = 0; s_next[rb] = 0; }
    const int nq = s_nq[cb];
    // ---- setup of the survivors: every WARP takes 32 of them at a time from the queue and reserves its records and
    // units with its own packed atomicAdd (warp scan by shuffles) -- no block scan and no barrier inside this phase
    for (;;) {
      int start = 0;
      if (lane == 0) start = atomicAdd(&s_next[cb], 32);
      start = __shfl_sync(0xffffffffu, start, 0);
      if (start >= nq) break;                                   // warp-uniform
      const int j = start + lane;
      TriRec T;
      int n_i = 0;
      if (j < nq) {
        const int f = s_queue[qb][j];
        const int i0 = faces ? __ldg(faces + 3 * (size_t)f) : 3 * f, i1 = faces ? __ldg(faces + 3 * (size_t)f + 1) : 3 * f + 1,
                  i2 = faces ? __ldg(faces + 3 * (size_t)f + 2) : 3 * f + 2;
        n_i = tri_setup<true>(f, i0, i1, i2, verts, o, P, s_mask, row_lim, T);
      }
      const int n_u = (n_i + kUnitItems - 1) / kUnitItems;
      const unsigned long long mine = n_i > 0 ? ((1ull << kUnitBits) | (unsigned long long)n_u) : 0ull;
      unsigned long long incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += u;
      }
      const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned long long base = 0ull;
      if (lane == 0 && total) {
        base = atomicAdd(&chdr->reserved, total);
        if ((base & units_mask) + (total & units_mask) > unit_cap) chdr->overflow = 1;
      }
      base = __shfl_sync(0xffffffffu, base, 0);
      if (n_i > 0) {
        const unsigned long long at = base + incl - mine;
        const int pos = (int)(at >> kUnitBits);
        const unsigned long long u0 = at & units_mask;
        if (pos < rec_cap && u0 + (unsigned long long)n_u <= unit_cap) {   // always, unless the unit list overflowed
          float4* r = recs + 4 * (size_t)pos;
          r[0] = make_float4(T.v0x, T.v0y, T.v0z, T.e1x);
          r[1] = make_float4(T.e1y, T.e1z, T.e2x, T.e2y);
          r[2] = make_float4(T.e2z, __int_as_float(T.orig), T.ymid, T.yhalf);
          r[3] = make_float4(T.slo, T.shi, __int_as_float(T.ca | (T.ncx << 16) | (T.ncx > kWideCols ? (1 << 30) : 0)),
                             __int_as_float(T.ra | (T.ncy << 16)));
          for (int k = 0; k < n_u; ++k) units[u0 + k] = make_int2(pos, k * kUnitItems);
        }
      }
    }
  }
  if (n_bad) atomicAdd(&chdr->n_bad_faces, n_bad);
}

// The same step as TWO kernels (vl_debug_cast_split(1)): the cull alone needs ~40 registers and runs at 6 CTAs per SM instead of
// 4 -- half as many again warps to hide the face -> vertex gathers behind --, the survivors' setup (64 registers) then runs
// on a dense list.  No shared-memory queue and no barrier in the cull: every CTA appends its survivors to its OWN segment
// of a global list (a warp-aggregated shared-memory counter; capacity = the faces the CTA's batches hold), k_cast_setup2 is
// launched with the same grid and CTA c sets up segment c, a warp per 32 entries.
#ifndef VL_CULL_MINB
#define VL_CULL_MINB 6
#endif
constexpr int kSegCounts = 4096;   // ints at the head of the list: survivors per CTA (grids are <= 148 x 8 CTAs)

__device__ __forceinline__ int cast_seg_cap(int n_faces, int grid) {
  const int n_batches = (n_faces + kBatch - 1) / kBatch;
  return ((n_batches + grid - 1) / grid) * kBatch;
}

template <bool kDescPtr>
__global__ void __launch_bounds__(kCastThreads, VL_CULL_MINB)
k_cast_cull(const VlBeamHeader* __restrict__ bhdr, int cw, int ch, const float* __restrict__ nxt, const VlMeshDesc mesh_val,
            const VlMeshDesc* __restrict__ mesh_ptr, const float* __restrict__ origin, VlCastHeader* chdr,
            int* __restrict__ list) {
  const float* __restrict__ verts = kDescPtr ? mesh_ptr->verts : mesh_val.verts;
  const int* __restrict__ faces = kDescPtr ? mesh_ptr->faces : mesh_val.faces;
  const int n_verts = kDescPtr ? mesh_ptr->n_verts : mesh_val.n_verts;
  const int n_faces = kDescPtr ? mesh_ptr->n_faces : mesh_val.n_faces;
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const BeamParams P = beam_params(bhdr, cw, ch);
  const float3 o = make_float3(__ldg(origin), __ldg(origin + 1), __ldg(origin + 2));
  const float o_max = fmaxf(fabsf(o.x), fmaxf(fabsf(o.y), fabsf(o.z)));
  const int lane = threadIdx.x & 31;
  const int n_batches = (n_faces + kBatch - 1) / kBatch;
  int* __restrict__ seg = list + kSegCounts + (size_t)blockIdx.x * cast_seg_cap(n_faces, gridDim.x);
  int n_bad = 0;
  for (int batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    {   // the faces (a soup: the vertices) of this CTA's NEXT batch start their way from DRAM to L2 now: one line per thread
      const int nb = batch + gridDim.x;
      if (nb < n_batches) {
        const size_t per = faces ? 12 : 36;
        const char* p = faces ? reinterpret_cast<const char*>(faces) + (size_t)nb * kBatch * 12
                              : reinterpret_cast<const char*>(verts) + (size_t)nb * kBatch * 36;
        const size_t bytes = (size_t)min(kBatch, n_faces - nb * kBatch) * per;
        for (size_t o2 = (size_t)threadIdx.x * 128; o2 < bytes; o2 += (size_t)kCastThreads * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o2));
      }
    }
    int idx[kBatch / kCastThreads][3];
#pragma unroll
    for (int k = 0; k < kBatch / kCastThreads; ++k) {   // all index loads first: faces -> verts is a dependent gather
      const int f = batch * kBatch + k * kCastThreads + threadIdx.x;
      if (f < n_faces) {
        if (faces) { idx[k][0] = __ldg(faces + 3 * (size_t)f); idx[k][1] = __ldg(faces + 3 * (size_t)f + 1); idx[k][2] = __ldg(faces + 3 * (size_t)f + 2); }
        else { idx[k][0] = 3 * f; idx[k][1] = 3 * f + 1; idx[k][2] = 3 * f + 2; }
      }
    }
#pragma unroll
    for (int k = 0; k < kBatch / kCastThreads; ++k) {
      const int f = batch * kBatch + k * kCastThreads + threadIdx.x;
      bool keep = false;
      if (f < n_faces) {
        if ((unsigned)idx[k][0] >= (unsigned)n_verts || (unsigned)idx[k][1] >= (unsigned)n_verts || (unsigned)idx[k][2] >= (unsigned)n_verts) {
          ++n_bad;
        } else {
          keep = tri_cull(idx[k][0], idx[k][1], idx[k][2], verts, o, o_max, P, nxt);
        }
      }
      const unsigned int m = __ballot_sync(0xffffffffu, keep);
      if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) seg[base + __popc(m & ((1u << lane) - 1u))] = f;
      }
    }
  }
  if (n_bad) atomicAdd(&chdr->n_bad_faces, n_bad);
  __syncthreads();
  if (threadIdx.x == 0) list[blockIdx.x] = s_count;
}

template <bool kDescPtr>
__global__ void __launch_bounds__(kCastThreads, VL_SETUP_MINB)
k_cast_setup2(const VlBeamHeader* __restrict__ bhdr, int cw, int ch, const float* __restrict__ nxt, const VlMeshDesc mesh_val,
              const VlMeshDesc* __restrict__ mesh_ptr, const float* __restrict__ origin, VlCastHeader* chdr,
              float4* __restrict__ recs, int rec_cap, int2* __restrict__ units, unsigned long long unit_cap,
              const float2* __restrict__ row_lim, const int* __restrict__ list) {
  const float* __restrict__ verts = kDescPtr ? mesh_ptr->verts : mesh_val.verts;
  const int* __restrict__ faces = kDescPtr ? mesh_ptr->faces : mesh_val.faces;
  const int n_faces = kDescPtr ? mesh_ptr->n_faces : mesh_val.n_faces;
  const BeamParams P = beam_params(bhdr, cw, ch);
  const float3 o = make_float3(__ldg(origin), __ldg(origin + 1), __ldg(origin + 2));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned long long units_mask = (1ull << kUnitBits) - 1ull;
  const int nq = __ldg(list + blockIdx.x);
  const int* __restrict__ seg = list + kSegCounts + (size_t)blockIdx.x * cast_seg_cap(n_faces, gridDim.x);
  for (int start = w * 32; start < nq; start += kCastWarps * 32) {
    const int j = start + lane;
    TriRec T;
    int n_i = 0;
    if (j < nq) {
      const int f = __ldg(seg + j);
      const int i0 = faces ? __ldg(faces + 3 * (size_t)f) : 3 * f, i1 = faces ? __ldg(faces + 3 * (size_t)f + 1) : 3 * f + 1,
                i2 = faces ? __ldg(faces + 3 * (size_t)f + 2) : 3 * f + 2;
      n_i = tri_setup<true>(f, i0, i1, i2, verts, o, P, nxt, row_lim, T);
    }
    const int n_u = (n_i + kUnitItems - 1) / kUnitItems;
    const unsigned long long mine = n_i > 0 ? ((1ull << kUnitBits) | (unsigned long long)n_u) : 0ull;
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += u;
    }
    const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0ull;
    if (lane == 0 && total) {
      base = atomicAdd(&chdr->reserved, total);
      if ((base & units_mask) + (total & units_mask) > unit_cap) chdr->overflow = 1;
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (n_i > 0) {
      const unsigned long long at = base + incl - mine;
      const int pos = (int)(at >> kUnitBits);
      const unsigned long long u0 = at & units_mask;
      if (pos < rec_cap && u0 + (unsigned long long)n_u <= unit_cap) {   // always, unless the unit list overflowed
        float4* r = recs + 4 * (size_t)pos;
        r[0] = make_float4(T.v0x, T.v0y, T.v0z, T.e1x);
        r[1] = make_float4(T.e1y, T.e1z, T.e2x, T.e2y);
        r[2] = make_float4(T.e2z, __int_as_float(T.orig), T.ymid, T.yhalf);
        r[3] = make_float4(T.slo, T.shi, __int_as_float(T.ca | (T.ncx << 16) | (T.ncx > kWideCols ? (1 << 30) : 0)),
                           __int_as_float(T.ra | (T.ncy << 16)));
        for (int k = 0; k < n_u; ++k) units[u0 + k] = make_int2(pos, k * kUnitItems);
      }
    }
  }
}

// Step 2: a warp per 32 work units.  A unit is one item: a run of <= 8 (64 for very wide triangles) cells of one cell row of a triangle's
// rectangle; the beams of consecutive cells are consecutive in the sorted beam list, so a run is one contiguous
// range of it (two when it wraps at the azimuth seam).  Each lane decodes its unit and parks the triangle in shared
// memory; then the warp pools the beams of its 32 units and tests them 32 at a time -- full lanes whatever the
// mix of units -- keeping the closest hit per beam slot with a fire-and-forget 64-bit atomicMin.  No CTA barrier; a
// triangle in front of the sensor (thousands of runs) is spread over as many lanes as it has units.
struct UnitRec {
  float v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z;
  unsigned int orig;
  float ymid, yhalf, slo, shi;
  int ka0, ka1, kb0, kb1;
};

__global__ void __launch_bounds__(kCastThreads)
k_cast_units(const VlBeamHeader* __restrict__ bhdr, int cw, int ch, const int* __restrict__ cell_start,
             const float4* __restrict__ sorted, const float* __restrict__ origin,
             unsigned long long* __restrict__ best, const VlCastHeader* __restrict__ chdr,
             const float4* __restrict__ recs, const int2* __restrict__ units) {
  __shared__ UnitRec s_rec[kCastWarps][32];
  if (chdr->overflow) return;
  const unsigned long long n_units = chdr->reserved & ((1ull << kUnitBits) - 1ull);
  const float3 o = make_float3(__ldg(origin), __ldg(origin + 1), __ldg(origin + 2));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned long long n_groups = (n_units + 31) / 32;
  const unsigned long long g_stride = (unsigned long long)gridDim.x * kCastWarps;
  for (unsigned long long g = (unsigned long long)blockIdx.x * kCastWarps + w; g < n_groups; g += g_stride) {
    const unsigned long long u = g * 32 + lane;
    int n_r = 0;
    const float4* rec = recs;
    float slo = 0.f, shi = 0.f;
    int ka0 = 0, ka1 = 0, kb0 = 0, kb1 = 0;
    if (u < n_units) {
      const int2 unit = __ldg(units + u);
      rec = recs + 4 * (size_t)unit.x;
      const float4 q3 = __ldg(rec + 3);
      slo = q3.x; shi = q3.y;
      const int pk0 = __float_as_int(q3.z), pk1 = __float_as_int(q3.w);
      const int ca = pk0 & 0xffff, ncx = (pk0 >> 16) & 0x1fff, ra = pk1 & 0xffff;
      const int sh = (pk0 >> 30) & 1 ? kSegShiftWide : kSegShift;
      const int nseg = (ncx + (1 << sh) - 1) >> sh;
      int yy = unit.y, sx = 0;
      if (nseg > 1) { yy = unit.y / nseg; sx = unit.y - yy * nseg; }
      const int len = min(1 << sh, ncx - (sx << sh));
      int c_begin = ca + (sx << sh);
      if (c_begin >= cw) c_begin -= cw;
      const int len_a = min(len, cw - c_begin), len_b = len - len_a;
      const int* row = cell_start + (size_t)(ra + yy) * cw;
      ka0 = __ldg(row + c_begin); ka1 = __ldg(row + c_begin + len_a);
      if (len_b > 0) { kb0 = __ldg(row); kb1 = __ldg(row + len_b); }
      n_r = ka1 - ka0 + kb1 - kb0;
    }
    int incl = n_r;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;   // warp-uniform
    const int off = incl - n_r;                                     // first pooled candidate of this lane's unit
    const unsigned int nz = __ballot_sync(0xffffffffu, n_r > 0);
    if (n_r > 0) {   // the units that have beams at all, parked densely: slot = rank among them (their offsets are strictly increasing)
      const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2);
      UnitRec& R = s_rec[w][__popc(nz & ((1u << lane) - 1u))];
      R.v0x = q0.x; R.v0y = q0.y; R.v0z = q0.z; R.e1x = q0.w; R.e1y = q1.x; R.e1z = q1.y; R.e2x = q1.z; R.e2y = q1.w; R.e2z = q2.x;
      R.orig = (unsigned int)__float_as_int(q2.y); R.ymid = q2.z; R.yhalf = q2.w; R.slo = slo; R.shi = shi;
      R.ka0 = ka0 - off; R.ka1 = ka1; R.kb0 = kb0; R.kb1 = kb1;      // candidate number -> sorted-beam index: ka0 - off + it
    }
    __syncwarp();
    int jb = 0;   // parked units that start before `base`
    for (int base = 0; base < total; base += 32) {
      // candidate -> unit without a search: the units that START inside this stretch of 32 candidates mark their first
      // candidate in one word (one warp-wide OR); a candidate's unit is the number of marks at or below it
      const int rel = off - base;
      const unsigned int M = __reduce_or_sync(0xffffffffu, (n_r > 0 && rel >= 0 && rel < 32) ? (1u << rel) : 0u);
      const int it = base + lane;
      if (it < total) {
        const UnitRec& R = s_rec[w][jb + __popc(M & (0xffffffffu >> (31 - lane))) - 1];
        int k = R.ka0 + it;
        if (k >= R.ka1) k = R.kb0 + (k - R.ka1);
        const float4 rd = __ldg(sorted + k);
        if (!(rd.z < R.slo || rd.z > R.shi)) {
          float t;
          if (vl_tri_hit(make_float4(R.v0x, R.v0y, R.v0z, 0.f), make_float4(R.e1x, R.e1y, R.e1z, 0.f),
                         make_float4(R.e2x, R.e2y, R.e2z, 0.f), o, make_float3(rd.x, rd.y, rd.z), &t)) {
            // a NaN t orders above the initial key and never wins; no value is read back (RED, not ATOM)
            atomicMin(best + k, ((unsigned long long)__float_as_uint(t) << 32) | R.orig);
          }
        }
      }
      jb += __popc(M);
    }
    __syncwarp();
  }
}

// RayTracer.cpp:73-90 write-back for the winning triangle of each beam; BVH.cpp:106-107 hit = o + d * t
template <bool kDescPtr>
__global__ void __launch_bounds__(kCastThreads)
k_cast_resolve(unsigned long long* best, int n, const float4* __restrict__ dir,
               const int* __restrict__ slot_of, const float* __restrict__ origin, const VlMeshDesc mesh_val,
               const VlMeshDesc* __restrict__ mesh_ptr, float* __restrict__ endpoints, int* __restrict__ endcolors,
               float* __restrict__ range, float* __restrict__ endrem, int* __restrict__ tri_id, bool zero_misses,
               bool colors_u8, VlCastHeader* __restrict__ chdr, bool rearm) {
  const int r = blockIdx.x * kCastThreads + threadIdx.x;
  if (r == 0) {   // the counters of this scan for the host; rearm: ... and cleared for the next scan (no k_cast_init then)
    chdr->snap_bad_faces = chdr->n_bad_faces; chdr->snap_overflow = chdr->overflow; chdr->snap_reserved = chdr->reserved;
    if (rearm) { chdr->n_bad_faces = 0; chdr->overflow = 0; chdr->reserved = 0ull; }
  }
  if (r >= n) return;
  const int* __restrict__ faces = kDescPtr ? mesh_ptr->faces : mesh_val.faces;
  const int* __restrict__ colors = kDescPtr ? mesh_ptr->colors : mesh_val.colors;
  const float* __restrict__ rem = kDescPtr ? mesh_ptr->rem : mesh_val.rem;
  const int slot = __ldg(slot_of + r);
  const unsigned long long key = slot >= 0 ? best[slot] : init_key();
  if (rearm && slot >= 0) best[slot] = init_key();   // every slot belongs to exactly one beam: re-armed for the next scan
  if (key < init_key()) {
    const int f = (int)(unsigned int)(key & 0xffffffffull);
    const float t = __uint_as_float((unsigned int)(key >> 32));
    const float4 d = dir[r];
    const float3 o = make_float3(__ldg(origin), __ldg(origin + 1), __ldg(origin + 2));
    const int i0 = faces ? __ldg(faces + 3 * (size_t)f) : 3 * f, i1 = faces ? __ldg(faces + 3 * (size_t)f + 1) : 3 * f + 1,
              i2 = faces ? __ldg(faces + 3 * (size_t)f + 2) : 3 * f + 2;
    endpoints[3 * (size_t)r + 0] = __fadd_rn(o.x, __fmul_rn(d.x, t));
    endpoints[3 * (size_t)r + 1] = __fadd_rn(o.y, __fmul_rn(d.y, t));
    endpoints[3 * (size_t)r + 2] = __fadd_rn(o.z, __fmul_rn(d.z, t));
    // RayTracer.cpp:36-48 colours pass through float; Triangle.h:53-56 colour of vertex 0
    if (colors_u8) {   // VL_COLORS_U8: the mesh extraction's uint8 colours as they are (colors.astype(np.int32), fusion_lidar.py:437)
      const unsigned char* c8 = reinterpret_cast<const unsigned char*>(colors) + 3 * (size_t)i0;
      endcolors[3 * (size_t)r + 0] = (int)__ldg(c8);
      endcolors[3 * (size_t)r + 1] = (int)__ldg(c8 + 1);
      endcolors[3 * (size_t)r + 2] = (int)__ldg(c8 + 2);
    } else {
      endcolors[3 * (size_t)r + 0] = (int)(float)__ldg(colors + 3 * (size_t)i0);
      endcolors[3 * (size_t)r + 1] = (int)(float)__ldg(colors + 3 * (size_t)i0 + 1);
      endcolors[3 * (size_t)r + 2] = (int)(float)__ldg(colors + 3 * (size_t)i0 + 2);
    }
    endrem[r] = __fdiv_rn(__fadd_rn(__fadd_rn(__ldg(rem + i0), __ldg(rem + i1)), __ldg(rem + i2)), 3.0f);   // Triangle.h:63-70
    range[r] = t;
    if (tri_id) tri_id[r] = f;
  } else {
    if (zero_misses) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { endpoints[3 * (size_t)r + k] = 0.f; endcolors[3 * (size_t)r + k] = 0; }
      endrem[r] = 0.f;
      range[r] = 0.f;
    }
    if (tri_id) tri_id[r] = -1;
  }
}

}  // namespace

extern "C" void vl_debug_cast_rearm(int on) { g_graph_rearm = on ? 1 : 0; }
extern "C" void vl_debug_cast_cells(int cells_per_beam_row) { g_cells_per_row = cells_per_beam_row < 1 ? 1 : cells_per_beam_row; }
extern "C" void vl_debug_cast_split(int on) { g_cast_split = on ? 1 : 0; }
extern "C" void vl_debug_cast_row_trim(int on) { g_row_trim = on ? 1 : 0; }
extern "C" void vl_debug_cast_ctas(int ctas_per_sm) { g_items_ctas_per_sm = ctas_per_sm < 1 ? 1 : ctas_per_sm; }
extern "C" void vl_debug_cast_setup_ctas(int ctas_per_sm) { g_setup_ctas_per_sm = ctas_per_sm < 1 ? 1 : ctas_per_sm; }

size_t vl_beams_bytes_impl(int n_rays, int height) { return beam_layout(n_rays, height).total; }

struct CastLayout { size_t off_best, off_units, off_recs, off_list, total; unsigned long long unit_cap; };

CastLayout cast_layout(int n_rays, int n_faces) {
  CastLayout C;
  const size_t nr = n_rays > 0 ? (size_t)n_rays : 1, nf = n_faces > 0 ? (size_t)n_faces : 1;
  size_t off = 256;
  C.off_best = off;       off = vl_align256(off + 8 * nr);
  C.unit_cap = 4 * nf + (1ull << 20);   // a unit is a run of <= 8 (64) cells of one cell row; LiDAR meshes need ~0.6 nf
  C.off_units = off;      off = vl_align256(off + 8 * (size_t)C.unit_cap);
  C.off_recs = off;       off = vl_align256(off + 64 * nf);
  // survivors of the cull per CTA (vl_debug_cast_split(1)): counts + one segment per CTA, each a whole number of batches
  C.off_list = off;       off = vl_align256(off + 4 * (kSegCounts + nf + (size_t)kSegCounts * kBatch));
  C.total = off;
  return C;
}

size_t vl_cast_workspace_bytes_impl(int n_rays, int n_faces) { return cast_layout(n_rays, n_faces).total; }

int vl_beams_build_launch(const float* d_rays, int n_rays, int height, void* d_beams, int flags, cudaStream_t stream) {
  const BeamLayout L = beam_layout(n_rays, height);
  char* B = static_cast<char*>(d_beams);
  VlBeamHeader* hdr = reinterpret_cast<VlBeamHeader*>(B);
  float4* dir = reinterpret_cast<float4*>(B + L.off_dir);
  float4* sorted = reinterpret_cast<float4*>(B + L.off_sorted);
  int* slot_of = reinterpret_cast<int*>(B + L.off_slot_of);
  int* cell_start = reinterpret_cast<int*>(B + L.off_cell_start);
  int* cursor = reinterpret_cast<int*>(B + L.off_cursor);
  int* fine = reinterpret_cast<int*>(B + L.off_fine);
  float* fine_mask = reinterpret_cast<float*>(B + L.off_mask);
  int* blk_sum = reinterpret_cast<int*>(B + L.off_blk);
  const int ncell = L.cw * L.ch;
  VlProfScope ps(VL_ST_BEAMS, stream);
  uint2* row_lim = reinterpret_cast<uint2*>(B + L.off_rowlim);
  k_beam_init<<<vl_sm_count(), 256, 0, stream>>>(hdr, cell_start, ncell + 1, fine, row_lim, L.ch);
  VL_LAUNCH_CHECK("k_beam_init");
  const int nb = (L.n + kCastThreads - 1) / kCastThreads;
  if (L.n > 0) {
    k_beam_prep<<<nb, kCastThreads, 0, stream>>>(d_rays, L.n, (flags & VL_RAYS_NORMALIZED) != 0, dir, hdr);
    VL_LAUNCH_CHECK("k_beam_prep");
    k_beam_count<<<nb, kCastThreads, 0, stream>>>(dir, L.n, hdr, L.cw, L.ch, cell_start, fine, row_lim);
    VL_LAUNCH_CHECK("k_beam_count");
  }
  const int nblk = (ncell + kScanBlock - 1) / kScanBlock;
  k_beam_scan_local<<<nblk, 1024, 0, stream>>>(cell_start, ncell, blk_sum);
  VL_LAUNCH_CHECK("k_beam_scan_local");
  k_beam_scan_top<<<1, 1024, 0, stream>>>(blk_sum, nblk, cell_start, ncell, fine, fine_mask, row_lim, L.ch);
  VL_LAUNCH_CHECK("k_beam_scan_top");
  k_beam_scan_apply<<<(ncell + 1023) / 1024, 1024, 0, stream>>>(cell_start, ncell, blk_sum, cursor);
  VL_LAUNCH_CHECK("k_beam_scan_apply");
  if (L.n > 0) {
    k_beam_scatter<<<nb, kCastThreads, 0, stream>>>(dir, L.n, hdr, L.cw, L.ch, cursor, sorted, slot_of);
    VL_LAUNCH_CHECK("k_beam_scatter");
  }
  return VL_OK;
}

// the four kernels of one cast.  cap_faces sizes the workspace sections (the actual mesh for vl_cast, the largest
// mesh of a stream slot for a graph); mesh is read from `d_desc` when by_ptr (graph replay) and passed by value otherwise.
static int cast_enqueue(const void* d_beams, const VlMeshDesc& mesh, bool by_ptr, int cap_faces, const float* d_origin,
                        int n_rays, int height, float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem,
                        int* d_tri_id, int flags, void* d_ws, cudaStream_t stream, bool init, bool rearm,
                        int phases = VL_CAST_PHASE_FRONT | VL_CAST_PHASE_RESOLVE) {
  const BeamLayout L = beam_layout(n_rays, height);
  const bool front = (phases & VL_CAST_PHASE_FRONT) != 0, back = (phases & VL_CAST_PHASE_RESOLVE) != 0;
  if (back && d_tri_id && n_rays > L.n)   // rays beyond width * height are never cast (RayTracer.cpp:56)
    VL_CUDA_CHECK(cudaMemsetAsync(d_tri_id + L.n, 0xff, sizeof(int) * (size_t)(n_rays - L.n), stream));
  if (L.n <= 0) {   // no ray is cast (fewer rays than rows): no kernel runs, but vl_cast_status must not read a stale header
    if (d_ws) VL_CUDA_CHECK(cudaMemsetAsync(d_ws, 0, sizeof(VlCastHeader), stream));
    return VL_OK;
  }
  const char* B = static_cast<const char*>(d_beams);
  const VlBeamHeader* bhdr = reinterpret_cast<const VlBeamHeader*>(B);
  const float4* dir = reinterpret_cast<const float4*>(B + L.off_dir);
  const float4* sorted = reinterpret_cast<const float4*>(B + L.off_sorted);
  const int* slot_of = reinterpret_cast<const int*>(B + L.off_slot_of);
  const int* cell_start = reinterpret_cast<const int*>(B + L.off_cell_start);
  const float* fine_mask = reinterpret_cast<const float*>(B + L.off_mask);
  const float2* row_lim = g_row_trim ? reinterpret_cast<const float2*>(B + L.off_rowlim) : nullptr;
  char* Wk = static_cast<char*>(d_ws);
  const CastLayout C = cast_layout(n_rays, cap_faces);
  VlCastHeader* chdr = reinterpret_cast<VlCastHeader*>(Wk);
  const VlMeshDesc* d_desc = reinterpret_cast<const VlMeshDesc*>(Wk + kDescOffset);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(Wk + C.off_best);
  int2* units = reinterpret_cast<int2*>(Wk + C.off_units);
  float4* recs = reinterpret_cast<float4*>(Wk + C.off_recs);
  const int rec_cap = cap_faces > 0 ? cap_faces : 1;
  if (init && front) {
    VlProfScope ps(VL_ST_CAST_INIT, stream);
    k_cast_init<<<vl_sm_count(), 256, 0, stream>>>(best, L.n, chdr);
    VL_LAUNCH_CHECK("k_cast_init");
  }
  if (front && (by_ptr || mesh.n_faces > 0)) {
    {
      VlProfScope ps(VL_ST_CAST_SETUP, stream);
      const int n_batches = ((by_ptr ? cap_faces : mesh.n_faces) + kBatch - 1) / kBatch;
      const int cap = vl_sm_count() * g_setup_ctas_per_sm;
      const int nb = n_batches < cap ? (n_batches > 0 ? n_batches : 1) : cap;
      int* list = reinterpret_cast<int*>(Wk + C.off_list);
      if (g_cast_split) {
        int nc = vl_sm_count() * VL_CULL_MINB;
        if (nc > kSegCounts) nc = kSegCounts;
        if (n_batches < nc) nc = n_batches > 0 ? n_batches : 1;
        if (by_ptr) {
          k_cast_cull<true><<<nc, kCastThreads, 0, stream>>>(bhdr, L.cw, L.ch, fine_mask, mesh, d_desc, d_origin, chdr, list);
          VL_LAUNCH_CHECK("k_cast_cull");
          k_cast_setup2<true><<<nc, kCastThreads, 0, stream>>>(bhdr, L.cw, L.ch, fine_mask, mesh, d_desc, d_origin, chdr, recs, rec_cap,
                                                              units, C.unit_cap, row_lim, list);
        } else {
          k_cast_cull<false><<<nc, kCastThreads, 0, stream>>>(bhdr, L.cw, L.ch, fine_mask, mesh, d_desc, d_origin, chdr, list);
          VL_LAUNCH_CHECK("k_cast_cull");
          k_cast_setup2<false><<<nc, kCastThreads, 0, stream>>>(bhdr, L.cw, L.ch, fine_mask, mesh, d_desc, d_origin, chdr, recs, rec_cap,
                                                               units, C.unit_cap, row_lim, list);
        }
        VL_LAUNCH_CHECK("k_cast_setup2");
      } else
      if (by_ptr)
        k_cast_setup<true><<<nb, kCastThreads, 0, stream>>>(bhdr, L.cw, L.ch, fine_mask, mesh, d_desc, d_origin, chdr, recs,
                                                           rec_cap, units, C.unit_cap, row_lim);
      else
        k_cast_setup<false><<<nb, kCastThreads, 0, stream>>>(bhdr, L.cw, L.ch, fine_mask, mesh, d_desc, d_origin, chdr, recs,
                                                            rec_cap, units, C.unit_cap, row_lim);
      VL_LAUNCH_CHECK("k_cast_setup");
    }
    {
      VlProfScope ps(VL_ST_CAST_ITEMS, stream);
      k_cast_units<<<vl_sm_count() * g_items_ctas_per_sm, kCastThreads, 0, stream>>>(bhdr, L.cw, L.ch, cell_start, sorted, d_origin,
                                                                          best, chdr, recs, units);
      VL_LAUNCH_CHECK("k_cast_units");
    }
  }
  if (back) {
    VlProfScope ps(VL_ST_CAST_RESOLVE, stream);
    const int nb = (L.n + kCastThreads - 1) / kCastThreads;
    const bool zm = (flags & VL_TRACE_ZERO_MISSES) != 0, c8 = (flags & VL_COLORS_U8) != 0;
    if (by_ptr)
      k_cast_resolve<true><<<nb, kCastThreads, 0, stream>>>(best, L.n, dir, slot_of, d_origin, mesh, d_desc, d_endpoints,
                                                           d_endcolors, d_range, d_endrem, d_tri_id, zm, c8, chdr, rearm);
    else
      k_cast_resolve<false><<<nb, kCastThreads, 0, stream>>>(best, L.n, dir, slot_of, d_origin, mesh, d_desc, d_endpoints,
                                                            d_endcolors, d_range, d_endrem, d_tri_id, zm, c8, chdr, rearm);
    VL_LAUNCH_CHECK("k_cast_resolve");
  }
  return VL_OK;
}

int vl_cast_launch(const void* d_beams, const float* d_verts, const int* d_faces, const int* d_colors,
                   const float* d_rem, int n_verts, int n_faces, const float* d_origin, int n_rays, int height,
                   float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem, int* d_tri_id, int flags,
                   void* d_ws, cudaStream_t stream, int phases) {
  VlMeshDesc mesh = {};
  mesh.verts = d_verts; mesh.faces = d_faces; mesh.colors = d_colors; mesh.rem = d_rem;
  mesh.n_verts = n_verts; mesh.n_faces = n_faces;
  return cast_enqueue(d_beams, mesh, false, n_faces, d_origin, n_rays, height, d_endpoints, d_endcolors, d_range, d_endrem,
                      d_tri_id, flags, d_ws, stream, true, false, phases);
}

// ---------------------------------------------------------------------------
// one scan = one graph launch: the cast of a stream slot (fixed beams, outputs, workspace) captured once, the mesh
// changes per scan through a 64-byte descriptor in pinned host memory that the graph's first node copies to the device
// ---------------------------------------------------------------------------
int vl_cast_graph_create_impl(const void* d_beams, const float* d_origin, int n_rays, int height, float* d_endpoints,
                              int* d_endcolors, float* d_range, float* d_endrem, int* d_tri_id, int flags, void* d_ws,
                              int max_faces, const void* h_desc, int* h_status, cudaStream_t stream, void** out_exec) {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  {   // the workspace of a slot is armed ONCE, here; every replay's k_cast_resolve re-arms it for the next scan
    const BeamLayout L = beam_layout(n_rays, height);
    const CastLayout C = cast_layout(n_rays, max_faces);
    char* Wk = static_cast<char*>(d_ws);
    k_cast_init<<<vl_sm_count(), 256, 0, stream>>>(reinterpret_cast<unsigned long long*>(Wk + C.off_best), L.n > 0 ? L.n : 0,
                                                  reinterpret_cast<VlCastHeader*>(Wk));
    VL_LAUNCH_CHECK("k_cast_init");
  }
  VL_CUDA_CHECK(cudaStreamSynchronize(stream));
  VL_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
  VlMeshDesc none = {};
  int rc = VL_OK;
  cudaError_t e = cudaMemcpyAsync(static_cast<char*>(d_ws) + kDescOffset, h_desc, sizeof(VlMeshDesc), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess)
    rc = cast_enqueue(d_beams, none, true, max_faces, d_origin, n_rays, height, d_endpoints, d_endcolors, d_range, d_endrem,
                      d_tri_id, flags, d_ws, stream, !g_graph_rearm, g_graph_rearm != 0);
  if (e == cudaSuccess && rc == VL_OK && h_status)
    e = cudaMemcpyAsync(h_status, static_cast<char*>(d_ws) + kSnapOffset, 16, cudaMemcpyDeviceToHost, stream);
  const cudaError_t e2 = cudaStreamEndCapture(stream, &graph);   // always leave capture mode
  if (rc != VL_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
  VL_CUDA_CHECK(e);
  VL_CUDA_CHECK(e2);
  VL_CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
  VL_CUDA_CHECK(cudaGraphDestroy(graph));
  *out_exec = exec;
  return VL_OK;
}

int vl_cast_graph_launch_impl(void* exec, cudaStream_t stream) {
  VL_CUDA_CHECK(cudaGraphLaunch(static_cast<cudaGraphExec_t>(exec), stream));
  for (int k = 0; k < (g_graph_rearm ? 3 : 4); ++k) vl_count_launch();   // setup, units, resolve (the slot was armed once, at creation)
  return VL_OK;
}

int vl_cast_graph_destroy_impl(void* exec) {
  if (exec) VL_CUDA_CHECK(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(exec)));
  return VL_OK;
}

int vl_cast_status_read(const void* d_ws, cudaStream_t stream, int* info) {
  VlCastHeader h;
  VL_CUDA_CHECK(cudaMemcpyAsync(&h, d_ws, sizeof(h), cudaMemcpyDeviceToHost, stream));
  VL_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (info) {
    info[0] = h.snap_bad_faces;
    info[1] = (int)(h.snap_reserved >> kUnitBits);                                       // triangles that can be hit at all
    const unsigned long long items = h.snap_reserved & ((1ull << kUnitBits) - 1ull);   // work units (<= 4 cell runs each)
    info[2] = (int)(items & 0x7fffffffull);
    info[3] = (int)(items >> 31);
  }
  if (h.snap_overflow) {
    vl_set_error("vl_cast: the mesh needs more work units (%llu) than the workspace holds -- results are invalid, use vl_bvh_build + vl_trace",
                 (unsigned long long)(h.snap_reserved & ((1ull << kUnitBits) - 1ull)));
    return VL_ENOSPACE;
  }
  if (h.snap_bad_faces > 0) {
    vl_set_error("mesh has %d face(s) with a vertex index outside [0, n_verts)", h.snap_bad_faces);
    return VL_EBADMESH;
  }
  return VL_OK;
}
