// vl_project.cu -- (iii) spherical range-image projection as an atomicMin-on-depth scatter, sm_100a.
//
// Replaces LaserScan.do_range_projection_new(method="depth") (auxiliary/laserscan.py:294-391,
// a per-point PYTHON loop) and SemLaserScan.do_label_projection_new (:672-676).
//
// The reference loop is sequential and compares the float64 depth of point i against the
// float32 value already stored in the image (:376-378):   depth[i] < range_image[py,px].
// Let R(i) = float32(depth[i]).  Per pixel the loop's final winner is (proof):
//   R* = min_i R(i);  among points with R(i) == R*:
//     the LAST  index with depth[i] <  R*  (such a point always displaces the incumbent), else
//     the FIRST index (all have depth[i] >= R*, none can displace the first entrant).
// One 64-bit atomicMin reproduces that exactly:
//   key = R bits << 32 | (depth < R ? 0x7fffffff - i : 0x80000000 | i).
// Compile with -fmad=false: the float64 expressions must round like numpy's.
#include "vl_common.cuh"

namespace {

constexpr int kThreads = 256;

struct ProjParams {
  double fov_down_abs, fov, pi;
  int H, W, remove;
  int method;   // VL_PROJECT_DEPTH (laserscan.py:369-391), VL_PROJECT_PDIST (:392-416), VL_PROJECT_DEPTHFAST (:418-437)
};

// One point through laserscan.py:304-360: kept or not, its pixel, its depth and the quantity the method orders by --
// the depth, or ('pdist', :399-400) the distance of its image position from the pixel centre,
// np.linalg.norm([proj_y, proj_x] - [py + 0.5, px + 0.5]) = sqrt(a * a + b * b) in float64.
struct PointPix { bool keep; int pix; double depth, q; };

__device__ __forceinline__ PointPix point_pixel(const double* __restrict__ pts, long i, const ProjParams& P,
                                                const double* __restrict__ beam_angles, int n_beam_angles) {
  PointPix r;
  r.keep = false; r.pix = 0; r.q = 0.0;
  const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  const double depth = sqrt((x * x + y * y) + z * z);  // np.linalg.norm(points, 2, axis=1), :304
  r.depth = depth;
  if (depth != 0.0) {                                   // :307-309
    const double yaw = -atan2(y, x);
    double pitch = asin(z / depth);
    if (n_beam_angles > 0) {  // :321-327  pitch <- the list entry nearest to it (argmin: first minimum)
      int best = 0;
      double best_d = fabs(pitch - __ldg(beam_angles));
      for (int k = 1; k < n_beam_angles; ++k) {
        const double d = fabs(pitch - __ldg(beam_angles + k));
        if (d < best_d) { best_d = d; best = k; }
      }
      pitch = __ldg(beam_angles + best);
    }
    double proj_x = 0.5 * (yaw / P.pi + 1.0);           // :329-330
    double proj_y = 1.0 - (pitch + P.fov_down_abs) / P.fov;
    if (!P.remove || (proj_y >= 0.0 && proj_y <= 1.0)) {  // :337-345
      proj_x *= P.W;
      proj_y *= P.H;
      const double fx = fmax(0.0, fmin((double)(P.W - 1), floor(proj_x)));  // :355-360
      const double fy = fmax(0.0, fmin((double)(P.H - 1), floor(proj_y)));
      r.pix = (int)fy * P.W + (int)fx;
      r.keep = true;
      r.q = depth;
      if (P.method == VL_PROJECT_PDIST) {
        const double a = proj_y - ((double)(int)fy + 0.5), b = proj_x - ((double)(int)fx + 0.5);
        r.q = sqrt(a * a + b * b);
      }
    }
  }
  return r;
}

__global__ void __launch_bounds__(kThreads)
k_project_scatter(const double* __restrict__ pts, long n, ProjParams P, const double* __restrict__ beam_angles,
                  int n_beam_angles, unsigned long long* __restrict__ keys, unsigned int* __restrict__ masks,
                  unsigned int* __restrict__ warp_counts, uint8_t* __restrict__ keep_out) {
  const long i = (long)blockIdx.x * kThreads + threadIdx.x;
  bool keep = false;
  if (i < n) {
    const PointPix pp = point_pixel(pts, i, P, beam_angles, n_beam_angles);
    keep = pp.keep;
    if (keep) {
      if (P.method == VL_PROJECT_DEPTHFAST) {
        // the smallest float64 depth per pixel (bits of a positive double order like its value); the index follows in
        // k_project_select
        atomicMin(keys + pp.pix, (unsigned long long)__double_as_longlong(pp.depth));
      } else {
        // 'depth' and 'pdist' both compare a float64 quantity against the float32 image it was last stored in
        const float R = (float)pp.q;
        const unsigned int lo = (pp.q < (double)R) ? (0x7fffffffu - (unsigned int)i) : (0x80000000u | (unsigned int)i);
        atomicMin(keys + pp.pix, ((unsigned long long)__float_as_uint(R) << 32) | lo);
      }
    }
    if (keep_out) keep_out[i] = keep ? 1 : 0;
  }
  const unsigned int m = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0) {
    const long w = i >> 5;
    if (w * 32 < n) { masks[w] = m; warp_counts[w] = __popc(m); }
  }
}

// 'depthfast', second pass: among the points whose depth IS the pixel's minimum the smallest index (the reference's
// winner among equal depths is whatever numpy's unstable argsort leaves last: unspecified)
__global__ void __launch_bounds__(kThreads)
k_project_select(const double* __restrict__ pts, long n, ProjParams P, const double* __restrict__ beam_angles,
                 int n_beam_angles, const unsigned long long* __restrict__ keys, unsigned int* __restrict__ winner) {
  const long i = (long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const PointPix pp = point_pixel(pts, i, P, beam_angles, n_beam_angles);
  if (pp.keep && (unsigned long long)__double_as_longlong(pp.depth) == keys[pp.pix]) atomicMin(winner + pp.pix, (unsigned int)i);
}

// single-CTA exclusive scan of the per-warp kept counts (n/32 entries); total -> n_kept
__global__ void __launch_bounds__(1024) k_scan_counts(unsigned int* __restrict__ counts, long total, int* n_kept) {
  __shared__ unsigned int warp_sums[32];
  __shared__ unsigned int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  constexpr int kPer = 4;
  for (long base = 0; base < total; base += 1024 * kPer) {
    unsigned int v[kPer], sum = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      long idx = base + (long)tid * kPer + k;
      v[k] = idx < total ? counts[idx] : 0u;
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
      warp_sums[lane] = wi - ws;
    }
    __syncthreads();
    unsigned int run = carry_s + warp_sums[wid] + (incl - sum);
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      long idx = base + (long)tid * kPer + k;
      if (idx < total) counts[idx] = run;
      run += v[k];
    }
    __syncthreads();
    if (tid == 1023) carry_s = run;
    __syncthreads();
  }
  if (tid == 0 && n_kept) *n_kept = (int)carry_s;
}

// per pixel: decode the winner, translate its index into the kept-point numbering, gather
__global__ void __launch_bounds__(kThreads)
k_project_gather(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ masks,
                 const unsigned int* __restrict__ warp_prefix, const float* __restrict__ remissions,
                 const uint32_t* __restrict__ labels, int n_pix, float* __restrict__ range, int32_t* __restrict__ index,
                 int32_t* __restrict__ label, float* __restrict__ rem, int method, const double* __restrict__ pts,
                 const unsigned int* __restrict__ winner) {
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= n_pix) return;
  const unsigned long long key = keys[p];
  if (key == ~0ull) {  // :362-367 initial values ('depthfast' writes into proj_range, which starts at -1)
    range[p] = method == VL_PROJECT_DEPTHFAST ? -1.0f : 0.0f; index[p] = -1; label[p] = 0; rem[p] = -1.0f;
    return;
  }
  unsigned int i;
  if (method == VL_PROJECT_DEPTHFAST) {
    i = winner[p];
    range[p] = (float)__longlong_as_double((long long)key);   // :428 float64 depth stored into the float32 image
  } else {
    const unsigned int lo = (unsigned int)key;
    i = (lo & 0x80000000u) ? (lo & 0x7fffffffu) : (0x7fffffffu - lo);
    if (method == VL_PROJECT_PDIST) {   // the key holds the distance to the pixel centre; the image gets the point's depth (:405)
      const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
      range[p] = (float)sqrt((x * x + y * y) + z * z);
    } else {
      range[p] = __uint_as_float((unsigned int)(key >> 32));
    }
  }
  index[p] = (int)(warp_prefix[i >> 5] + __popc(masks[i >> 5] & ((1u << (i & 31)) - 1u)));
  label[p] = (int32_t)labels[i];   // :672-676
  rem[p] = remissions[i];
}

// Reverse projection of the `cp` adaption (laserscan.py:475-501): pixel coordinates + depth -> xyz, float64 like the
// reference's numpy (proj_x / proj_y are the per-pixel image coordinates of the winning point, float or clamped).
__global__ void __launch_bounds__(kThreads)
k_reverse_project(const float* __restrict__ depth_im, const double* __restrict__ proj_x, const double* __restrict__ proj_y,
                  int n, int H, int W, double fov_down_abs, double fov, double pi, double* __restrict__ back_points) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const double depth = (double)depth_im[i];
  const double px = proj_x[i] / (double)W, py = proj_y[i] / (double)H;          // :484-489
  const double yaw = (px * 2 - 1.0) * pi;                                          // :492
  const double pitch = pi / 2 - ((1.0 * fov - py * fov) - fov_down_abs);          // :494
  back_points[3 * (size_t)i] = depth * sin(pitch) * cos(-yaw);                     // :495-497
  back_points[3 * (size_t)i + 1] = depth * sin(pitch) * sin(-yaw);
  back_points[3 * (size_t)i + 2] = depth * cos(pitch);
}

// SemLaserScan.get_bnds (laserscan.py: np.amin / np.amax of the points over axis 0) restricted to the points the
// projection kept: one CTA, no atomics -- minima and maxima are exact whatever the order.  out[0..2] = min xyz,
// out[3..5] = max xyz (+inf / -inf when no point is kept).
__global__ void __launch_bounds__(1024)
k_points_bounds(const double* __restrict__ pts, const uint8_t* __restrict__ keep, long n, double* __restrict__ out) {
  __shared__ double s_lo[3][32], s_hi[3][32];
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (long i = threadIdx.x; i < n; i += 1024) {
    if (keep && !keep[i]) continue;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double v = pts[3 * i + k];
      lo[k] = v < lo[k] ? v : lo[k];     // np.amin / np.amax propagate NaN; points with NaN never survive the filters
      hi[k] = v > hi[k] ? v : hi[k];
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    if (lane == 0) { s_lo[k][w] = lo[k]; s_hi[k][w] = hi[k]; }
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double a = s_lo[k][lane], b = s_hi[k][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
      }
      if (lane == 0) { out[k] = a; out[3 + k] = b; }
    }
  }
}

}  // namespace

extern "C" int vl_points_bounds(const double* d_points, const uint8_t* d_keep, long n_points, double* d_bounds6,
                                vl_stream stream_) {
  if (n_points < 0 || !d_bounds6 || (n_points > 0 && !d_points)) {
    vl_set_error("vl_points_bounds: invalid argument");
    return VL_EINVAL;
  }
  k_points_bounds<<<1, 1024, 0, static_cast<cudaStream_t>(stream_)>>>(d_points, d_keep, n_points, d_bounds6);
  VL_LAUNCH_CHECK("k_points_bounds");
  return VL_OK;
}

extern "C" int vl_reverse_project(const float* d_depth_im, const double* d_proj_x, const double* d_proj_y, int H, int W,
                                  double fov_up_deg, double fov_down_deg, double* d_back_points, vl_stream stream_) {
  if (H <= 0 || W <= 0 || !d_depth_im || !d_proj_x || !d_proj_y || !d_back_points) {
    vl_set_error("vl_reverse_project: invalid argument (H %d, W %d)", H, W);
    return VL_EINVAL;
  }
  const double pi = 3.141592653589793;
  const double fu = fov_up_deg / 180.0 * pi, fd = fov_down_deg / 180.0 * pi;       // :477-479
  const double fov = fabs(fd) + fabs(fu);
  const int n = H * W;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  k_reverse_project<<<(n + kThreads - 1) / kThreads, kThreads, 0, stream>>>(d_depth_im, d_proj_x, d_proj_y, n, H, W, fabs(fd),
                                                                           fov, pi, d_back_points);
  VL_LAUNCH_CHECK("k_reverse_project");
  return VL_OK;
}

extern "C" size_t vl_project_workspace_bytes(long n_points, int H, int W) {
  size_t nw = (size_t)((n_points + 31) / 32) + 1;
  return vl_align256(8 * (size_t)H * W) + 2 * vl_align256(4 * nw) + vl_align256(4 * (size_t)H * W);   // keys, masks, counts, 'depthfast' winners
}

extern "C" int vl_project_select(const double* d_points, const float* d_remissions, const uint32_t* d_labels,
                                 long n_points, double fov_up_deg, double fov_down_deg, int H, int W, int remove,
                                 const double* d_beam_angles, int n_beam_angles, int method, float* d_range, int32_t* d_index,
                                 int32_t* d_label, float* d_rem, uint8_t* d_keep, int* d_n_kept, void* d_workspace,
                                 size_t workspace_bytes, vl_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (method != VL_PROJECT_DEPTH && method != VL_PROJECT_PDIST && method != VL_PROJECT_DEPTHFAST) {
    vl_set_error("vl_project: unknown method %d", method);
    return VL_EINVAL;
  }
  if (H <= 0 || W <= 0 || n_points < 0 || n_points >= 0x7fffffffL || !d_range || !d_index || !d_label || !d_rem ||
      !d_workspace || (n_points > 0 && (!d_points || !d_remissions || !d_labels)) || n_beam_angles < 0 ||
      (n_beam_angles > 0 && !d_beam_angles)) {
    vl_set_error("vl_project: invalid argument");
    return VL_EINVAL;
  }
  if (workspace_bytes < vl_project_workspace_bytes(n_points, H, W)) {
    vl_set_error("vl_project: workspace too small (%zu < %zu)", workspace_bytes, vl_project_workspace_bytes(n_points, H, W));
    return VL_ENOSPACE;
  }
  const size_t nw = (size_t)((n_points + 31) / 32);
  char* ws = static_cast<char*>(d_workspace);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws);
  unsigned int* masks = reinterpret_cast<unsigned int*>(ws + vl_align256(8 * (size_t)H * W));
  unsigned int* counts = reinterpret_cast<unsigned int*>(ws + vl_align256(8 * (size_t)H * W) + vl_align256(4 * (nw + 1)));
  unsigned int* winner = reinterpret_cast<unsigned int*>(ws + vl_align256(8 * (size_t)H * W) + 2 * vl_align256(4 * (nw + 1)));
  ProjParams P;
  P.pi = 3.141592653589793;  // np.pi
  const double fov_up = fov_up_deg / 180.0 * P.pi, fov_down = fov_down_deg / 180.0 * P.pi;  // :299-301
  P.fov_down_abs = fabs(fov_down);
  P.fov = fabs(fov_down) + fabs(fov_up);
  P.H = H; P.W = W; P.remove = remove; P.method = method;
  VL_CUDA_CHECK(cudaMemsetAsync(keys, 0xff, 8 * (size_t)H * W, stream));
  if (n_points > 0) {
    const int nb = (int)((n_points + kThreads - 1) / kThreads);
    VlProfScope ps(VL_ST_PROJECT_SCATTER, stream);
    k_project_scatter<<<nb, kThreads, 0, stream>>>(d_points, n_points, P, d_beam_angles, n_beam_angles, keys, masks, counts,
                                                   d_keep);
    VL_LAUNCH_CHECK("k_project_scatter");
    if (method == VL_PROJECT_DEPTHFAST) {
      VL_CUDA_CHECK(cudaMemsetAsync(winner, 0xff, 4 * (size_t)H * W, stream));
      k_project_select<<<nb, kThreads, 0, stream>>>(d_points, n_points, P, d_beam_angles, n_beam_angles, keys, winner);
      VL_LAUNCH_CHECK("k_project_select");
    }
  }
  k_scan_counts<<<1, 1024, 0, stream>>>(counts, (long)nw, d_n_kept);
  VL_LAUNCH_CHECK("k_scan_counts");
  const int n_pix = H * W;
  VlProfScope ps(VL_ST_PROJECT_GATHER, stream);
  k_project_gather<<<(n_pix + kThreads - 1) / kThreads, kThreads, 0, stream>>>(keys, masks, counts, d_remissions, d_labels,
                                                                             n_pix, d_range, d_index, d_label, d_rem, method, d_points,
                                                                             winner);
  VL_LAUNCH_CHECK("k_project_gather");
  return VL_OK;
}

extern "C" int vl_project_snap(const double* d_points, const float* d_remissions, const uint32_t* d_labels,
                               long n_points, double fov_up_deg, double fov_down_deg, int H, int W, int remove,
                               const double* d_beam_angles, int n_beam_angles, float* d_range, int32_t* d_index,
                               int32_t* d_label, float* d_rem, uint8_t* d_keep, int* d_n_kept, void* d_workspace,
                               size_t workspace_bytes, vl_stream stream) {
  return vl_project_select(d_points, d_remissions, d_labels, n_points, fov_up_deg, fov_down_deg, H, W, remove, d_beam_angles,
                           n_beam_angles, VL_PROJECT_DEPTH, d_range, d_index, d_label, d_rem, d_keep, d_n_kept, d_workspace,
                           workspace_bytes, stream);
}

extern "C" int vl_project(const double* d_points, const float* d_remissions, const uint32_t* d_labels, long n_points,
                          double fov_up_deg, double fov_down_deg, int H, int W, int remove, float* d_range,
                          int32_t* d_index, int32_t* d_label, float* d_rem, uint8_t* d_keep, int* d_n_kept,
                          void* d_workspace, size_t workspace_bytes, vl_stream stream) {
  return vl_project_snap(d_points, d_remissions, d_labels, n_points, fov_up_deg, fov_down_deg, H, W, remove, nullptr, 0,
                         d_range, d_index, d_label, d_rem, d_keep, d_n_kept, d_workspace, workspace_bytes, stream);
}
