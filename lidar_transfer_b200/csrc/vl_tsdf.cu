// vl_tsdf.cu -- (iv) class-aware TSDF voxel integration, sm_100a.
//
// Replaces the reference's pycuda `integrate` kernel (auxiliary/fusion_lidar.py:66-229, the
// class-aware branch `merge == true`, :190-227) and its launch loop (:252-287); vl_tsdf_init
// replaces the host-side volume construction + 4 uploads (:48-63).  The volumes never leave HBM.
//
// Arithmetic follows the kernel string operation by operation, including its quirks
// (SURVEY.md A.4): float-precision voxel index decode, `dist_old` read from the WEIGHT volume,
// no weight update on a class switch, mixed float/double constants.  The reference is JIT-compiled
// by nvcc with default -fmad=true; this file is compiled with -fmad=false and spells out the three
// places where that contraction happens (__fmaf_rn), so the rounding is explicit.
// The one-element out-of-bounds access of the reference's `voxel_idx > N` guard (:92) is NOT
// reproduced: voxel_idx == N is never touched.
#include "vl_common.cuh"

namespace {

constexpr int kThreads = 256;
#define VL_PI 3.14159265358979323846

struct TsdfParams {
  int dx, dy, dz;
  float ox, oy, oz;
  float voxel_size, trunc_margin, obs_weight;
  float fov_up, fov_down;  // radians: float(deg * PI / 180.0) evaluated in double on the host, like :119-120 per thread
  int im_h, im_w;
};

__global__ void __launch_bounds__(kThreads)
k_tsdf_init(float4* __restrict__ tsdf, float4* __restrict__ weight, float4* __restrict__ color, float4* __restrict__ rem,
            long long n4, float* tsdf_s, float* weight_s, float* color_s, float* rem_s, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    tsdf[i] = one; weight[i] = zero; color[i] = zero; rem[i] = zero;
  }
  // tail (n not a multiple of 4)
  for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    tsdf_s[i] = 1.f; weight_s[i] = 0.f; color_s[i] = 0.f; rem_s[i] = 0.f;
  }
}

// image column of a voxel position (x, y): :125 yaw, :134-141 proj_x -- depends on x and y only
__device__ __forceinline__ int tsdf_pixel_x(float cam_pt_x, float cam_pt_y, int im_w) {
  float yaw = -atan2f(cam_pt_y, cam_pt_x);
  float proj_x = 0.5 * (yaw / VL_PI + 1.0);        // :134
  proj_x *= im_w;
  int proj_x_cl = (int)floorf(proj_x);             // :139-141
  proj_x_cl = min(im_w - 1, proj_x_cl);
  proj_x_cl = max(0, proj_x_cl);
  return proj_x_cl;
}

// The dz voxels of a z column share x and y, hence the arctangent, the double-precision image coordinate and
// the pixel column: computed once per column here (same expressions, same bits), looked up per voxel below.
__global__ void __launch_bounds__(kThreads)
k_tsdf_columns(int* __restrict__ col_px, const TsdfParams P) {
  const int c = blockIdx.x * kThreads + threadIdx.x;
  if (c >= P.dx * P.dy) return;
  const int vx = c / P.dy, vy = c - vx * P.dy;
  col_px[c] = tsdf_pixel_x(__fmaf_rn((float)vx, P.voxel_size, P.ox), __fmaf_rn((float)vy, P.voxel_size, P.oy), P.im_w);
}

// kTable: pixel column from the per-column table.  kFresh: the volumes are known to hold their initial state
// (tsdf 1, weight / colour / remission 0, fusion_lidar.py:48-63) -- nothing is read, every voxel is written (its
// updated value, or the initial one): vl_tsdf_init + vl_tsdf_integrate in one pass over the volume.
template <bool kTable, bool kFresh>
__device__ __forceinline__ void tsdf_voxel(const int voxel_idx, float* __restrict__ tsdf_vol, float* __restrict__ weight_vol,
                                           float* __restrict__ color_vol, float* __restrict__ rem_vol, const TsdfParams& P,
                                           const float* __restrict__ color_im, const float* __restrict__ depth_im,
                                           const float* __restrict__ rem_im, const int* __restrict__ col_px) {
  float out_tsdf = 1.f, out_weight = 0.f, out_color = 0.f, out_rem = 0.f;   // kFresh: what is stored at the end
  do {
    const int vol_dim_y = P.dy, vol_dim_z = P.dz;
    // :96-98 (float division: can decode (x+1, -1, z) once voxel_idx exceeds 2^24)
    float voxel_x = floorf(((float)voxel_idx) / ((float)(vol_dim_y * vol_dim_z)));
    float voxel_y = floorf(((float)(voxel_idx - ((int)voxel_x) * vol_dim_y * vol_dim_z)) / ((float)vol_dim_z));
    float voxel_z = (float)(voxel_idx - ((int)voxel_x) * vol_dim_y * vol_dim_z - ((int)voxel_y) * vol_dim_z);
    // :101-104
    float voxel_size = P.voxel_size;
    float pt_x = __fmaf_rn(voxel_x, voxel_size, P.ox);
    float pt_y = __fmaf_rn(voxel_y, voxel_size, P.oy);
    float pt_z = __fmaf_rn(voxel_z, voxel_size, P.oz);
    float cam_pt_x = pt_x, cam_pt_z = pt_z, cam_pt_y = pt_y;
    // :119-128
    int im_h = P.im_h, im_w = P.im_w;
    const float fov_up = P.fov_up, fov_down = P.fov_down;  // :119-120, hoisted (same IEEE double expression on the host)
    float fov = fabsf(fov_up) + fabsf(fov_down);
    float depth = norm3df(cam_pt_x, cam_pt_y, cam_pt_z);
    float pitch = asinf(cam_pt_z / depth);
    if (pitch > fov_up || pitch < fov_down) break;  // :131-132
    // :125, :134-141 -- from the column table when the (float-decoded, :96-98) voxel lies inside it
    const int vxi = (int)voxel_x, vyi = (int)voxel_y;
    int proj_x_cl;
    if (kTable && (unsigned)vxi < (unsigned)P.dx && (unsigned)vyi < (unsigned)P.dy) proj_x_cl = __ldg(col_px + vxi * P.dy + vyi);
    else proj_x_cl = tsdf_pixel_x(cam_pt_x, cam_pt_y, im_w);
    float proj_y = 1.0 - (pitch + fabsf(fov_down)) / fov;   // :135
    proj_y *= im_h;
    int proj_y_cl = (int)floorf(proj_y);             // :142-144
    proj_y_cl = min(im_h - 1, proj_y_cl);
    proj_y_cl = max(0, proj_y_cl);
    int pixel_x = proj_x_cl, pixel_y = proj_y_cl;
    float depth_value = __ldg(depth_im + pixel_y * im_w + pixel_x);  // :154-156
    if (depth_value == 0) break;
    // :190-227 class-aware integration
    float trunc_margin = P.trunc_margin;
    float depth_diff = depth_value - depth;
    if (depth_diff < -trunc_margin) break;
    float dist = fminf(1.0f, depth_diff / trunc_margin);
    float dist_old = kFresh ? 0.f : weight_vol[voxel_idx];  // sic (:197)
    float old_color = kFresh ? 0.f : color_vol[voxel_idx];
    float new_color = __ldg(color_im + pixel_y * im_w + pixel_x);
    if (old_color == new_color) {
      float w_old = dist_old;
      float w_new = w_old + P.obs_weight;
      out_weight = w_new;
      out_tsdf = __fmaf_rn(kFresh ? 1.f : tsdf_vol[voxel_idx], w_old, dist) / w_new;
      float old_rem = kFresh ? 0.f : rem_vol[voxel_idx];
      float new_rem = __ldg(rem_im + pixel_y * im_w + pixel_x);
      out_rem = __fmaf_rn(old_rem, w_old, new_rem) / w_new;
      if (!kFresh) { weight_vol[voxel_idx] = out_weight; tsdf_vol[voxel_idx] = out_tsdf; rem_vol[voxel_idx] = out_rem; }
      else out_color = old_color;
    } else if (dist < dist_old) {
      out_tsdf = dist;
      float new_b = floorf(new_color / (256 * 256));
      float new_g = floorf((new_color - new_b * 256 * 256) / 256);
      float new_r = new_color - new_b * 256 * 256 - new_g * 256;
      out_color = new_b * 256 * 256 + new_g * 256 + new_r;
      out_rem = __ldg(rem_im + pixel_y * im_w + pixel_x);
      if (!kFresh) { tsdf_vol[voxel_idx] = out_tsdf; color_vol[voxel_idx] = out_color; rem_vol[voxel_idx] = out_rem; }
    }
  } while (false);
  if (kFresh) {
    tsdf_vol[voxel_idx] = out_tsdf; weight_vol[voxel_idx] = out_weight; color_vol[voxel_idx] = out_color; rem_vol[voxel_idx] = out_rem;
  }
}

template <bool kTable, bool kFresh>
__global__ void __launch_bounds__(kThreads)
k_tsdf_integrate(float* __restrict__ tsdf_vol, float* __restrict__ weight_vol, float* __restrict__ color_vol,
                 float* __restrict__ rem_vol, const TsdfParams P, const float* __restrict__ color_im,
                 const float* __restrict__ depth_im, const float* __restrict__ rem_im, long long n_vox,
                 const int* __restrict__ col_px) {
  const long long vi = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (vi >= n_vox) return;
  tsdf_voxel<kTable, kFresh>((int)vi, tsdf_vol, weight_vol, color_vol, rem_vol, P, color_im, depth_im, rem_im, col_px);
}

// ---- first integration into a NEW volume: only a thin shell of voxels can change ---------------------------------
//
// With tsdf 1 / weight 0 / colour 0 / remission 0 in the volume (fusion_lidar.py:48-52), the kernel string's update
// (:190-227) leaves a voxel at its initial value unless its pixel is non-empty (:156), depth_diff >= -trunc (:193) and
//   * the pixel's colour is 0 (== the volume's: running-average branch, weight becomes obs_weight), or
//   * dist < dist_old == 0 (class-switch branch), i.e. depth_diff < 0: the voxel lies BEHIND the surface.
// So per pixel the voxel depths that matter are (lo, hi] with lo = the pixel's depth (-inf for colour 0) and
// hi = depth + trunc: k_tsdf_shell stores that pair per pixel.  The sweep then needs, per voxel, only a CONSERVATIVE
// bracket of what the reference computes: the exact pixel column of its z column (k_tsdf_columns), its depth to
// 1e-5 (rsqrt.approx instead of norm3df) and its fractional image row to +-eps_row (arcsine by its odd series up to
// s^15: below |asin|, by < 6e-6 rad for |s| <= 0.62, monotonic beyond) -- one candidate pixel, two when the row
// estimate is within eps_row of a row boundary.  A voxel whose bracket misses the shell of its candidate pixels, or
// that is outside the vertical field of view by more than the series' error, is written with the initial values
// straight away; every other voxel is queued in shared memory and evaluated by tsdf_voxel<true, true> -- the
// reference arithmetic, dense over the queue.  Voxels within kDecodeWindow of a slab boundary (where the float
// index decode :96-98 can land in the neighbouring slab) always take that path.
// kVec 4: a thread owns 4 consecutive voxels of one z column (dz % 4 == 0, 16-byte aligned volumes): one decode,
// one column lookup, float4 stores.
constexpr int kFastChunk = 1024;      // voxels per CTA
constexpr int kDecodeWindow = 256;    // >= 64 (float(idx) for idx < 2^31) + n_vox * 2^-24 (rounding of the quotient)
constexpr float kAsinErr = 1e-5f;     // series truncation (< 6e-6 for |s| <= 0.62) + float rounding
constexpr float kDepthRel = 1e-5f;

struct ShellParams {
  int slab;                // dy * dz  (<= 2^24: the decode of y and z is then exact in float)
  float inv_dz;
  float fov_abs_down, h_over_fov, pitch_hi, pitch_lo;   // pitch_hi = fov_up + kAsinErr, pitch_lo = fov_down - kAsinErr
  float eps_row;           // bound on |estimated - exact| fractional image row (< 0.5)
};

// shell[c * H + r]: COLUMN-major, so that the rows a z column's voxels fall into lie next to each other
__global__ void __launch_bounds__(kThreads)
k_tsdf_shell(const float* __restrict__ depth_im, const float* __restrict__ color_im, int H, int W, float trunc,
             float2* __restrict__ shell) {
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= H * W) return;
  const float d = __ldg(depth_im + p);
  float lo = INFINITY, hi = -INFINITY;                       // :156 empty pixel: nothing is integrated
  if (d != 0.f) {
    if (!(fabsf(d) < INFINITY)) { lo = -INFINITY; hi = INFINITY; }   // NaN / inf depth: leave it to the exact path
    else { lo = __ldg(color_im + p) == 0.f ? -INFINITY : d; hi = d + trunc; }
  }
  const int r = p / W, c = p - r * W;
  shell[c * H + r] = make_float2(lo, hi);
}

// Bracket of one voxel (x, y fixed per column: xy2 = x^2 + y^2) at height z: its depth to +-kDepthRel and the one or
// two image rows it can fall into; r0 < 0: outside the vertical field of view for certain (:131).
struct VoxelBracket { float d_lo, d_hi; int r0, r1; };

__device__ __forceinline__ VoxelBracket shell_bracket(float xy2, float z, const TsdfParams& P, const ShellParams& S) {
  // bracket only: contracted and approximate arithmetic is fine here (the file is compiled with -fmad=false)
  const float q = __fmaf_rn(z, z, xy2);
  float rq;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rq) : "f"(q));   // 2 ulp; q == 0 (the origin voxel) gives inf
  const float d = q * rq;                                    // ... and d = NaN: no comparison rules that voxel out
  const float sn = z * rq, s2 = sn * sn;
  // asin(s) = s + s^3/6 + 3 s^5/40 + 15 s^7/336 + 105 s^9/3456 + 945 s^11/42240 + 10395 s^13/599040
  //           + 135135 s^15/9676800 + ...   (remainder < 6e-6 for |s| <= 0.62)
  float poly = __fmaf_rn(s2, 135135.f / 9676800, 10395.f / 599040);
  poly = __fmaf_rn(s2, poly, 945.f / 42240);
  poly = __fmaf_rn(s2, poly, 105.f / 3456);
  poly = __fmaf_rn(s2, poly, 15.f / 336);
  poly = __fmaf_rn(s2, poly, 3.f / 40);
  poly = __fmaf_rn(s2, poly, 1.f / 6);
  const float pitch = __fmaf_rn(sn * s2, poly, sn);
  const float rowf = __fmaf_rn(-(pitch + S.fov_abs_down), S.h_over_fov, (float)P.im_h);
  VoxelBracket b;
  b.d_lo = d * (1.f - kDepthRel);
  b.d_hi = d * (1.f + kDepthRel);
  const float f0 = floorf(rowf - S.eps_row);                 // eps_row < 0.5: floor(rowf + eps_row) is f0 or f0 + 1
  const int r0 = (int)f0;
  const int r1 = rowf + S.eps_row >= f0 + 1.f ? r0 + 1 : r0;
  b.r0 = max(0, min(P.im_h - 1, r0));
  b.r1 = max(0, min(P.im_h - 1, r1));
  if (pitch > S.pitch_hi || pitch < S.pitch_lo) b.r0 = -1;
  return b;
}

// kFresh false: a LATER integration (the volume holds earlier scans).  A voxel behind every candidate pixel's shell,
// outside the field of view or on an empty pixel is skipped as before -- those exits of the kernel string do not look
// at the volume.  A voxel IN FRONT of its candidates (free space, all of them labelled) is left alone only if it has
// never been written (colour 0 and weight 0: then `old_color == new_color` fails and `dist < weight` fails for every
// dist >= 0); its colour and weight are read for that, everything else about it is not.  Nothing is written outside
// the queue.
template <int kVec, bool kFresh>
__global__ void __launch_bounds__(kThreads)
k_tsdf_fresh_shell(float* __restrict__ tsdf_vol, float* __restrict__ weight_vol, float* __restrict__ color_vol,
                   float* __restrict__ rem_vol, const TsdfParams P, const ShellParams S,
                   const float* __restrict__ color_im, const float* __restrict__ depth_im,
                   const float* __restrict__ rem_im, const int* __restrict__ col_px, const float2* __restrict__ shell) {
  __shared__ int s_q[kFastChunk];
  __shared__ int s_nq;
  const int vx = blockIdx.y;
  const int base = blockIdx.x * kFastChunk;
  if (threadIdx.x == 0) s_nq = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned int lt = (1u << lane) - 1;
  const float x = __fmaf_rn((float)vx, P.voxel_size, P.ox);   // :101-104, the values tsdf_voxel sees
#pragma unroll
  for (int k = 0; k < kFastChunk / (kThreads * kVec); ++k) {
    const int rem_i = base + kVec * (k * kThreads + threadIdx.x);   // kVec 4: slab % 4 == 0, so all four are live or none
    const bool live = rem_i < S.slab;
    const int voxel_idx = vx * S.slab + rem_i;
    unsigned int ex = 0;                                            // bit j: voxel rem_i + j takes the exact path
    if (live) {
      if (rem_i < kDecodeWindow || rem_i + kVec > S.slab - kDecodeWindow) {
        ex = (1u << kVec) - 1;
      } else {
        int vy = (int)((float)rem_i * S.inv_dz);
        int vz = rem_i - vy * P.dz;
        if (vz < 0) { vy--; vz += P.dz; } else if (vz >= P.dz) { vy++; vz -= P.dz; }
        const float y = __fmaf_rn((float)vy, P.voxel_size, P.oy);
        const float xy2 = __fmaf_rn(x, x, y * y);
        const int px = __ldg(col_px + vx * P.dy + vy);
        const float2* __restrict__ col = shell + px * P.im_h;
        VoxelBracket b[kVec];
        float2 lh[kVec];
#pragma unroll
        for (int j = 0; j < kVec; ++j)                               // kVec 4: dz % 4 == 0, same z column
          b[j] = shell_bracket(xy2, __fmaf_rn((float)(vz + j), P.voxel_size, P.oz), P, S);
#pragma unroll
        for (int j = 0; j < kVec; ++j) lh[j] = __ldg(col + max(b[j].r0, 0));   // all loads in flight before the first use
        unsigned int front = 0;                                      // kFresh false: free space, decided by the voxel's state
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
          if (b[j].r0 < 0) continue;
          // per candidate pixel: behind its shell (or empty pixel) / in front of it / neither
          bool behind = b[j].d_lo > lh[j].y;
          bool infront = !behind && b[j].d_hi < lh[j].x;
          bool may = !behind && !infront;
          if (b[j].r1 != b[j].r0) {
            const float2 o = __ldg(col + b[j].r1);
            const bool behind2 = b[j].d_lo > o.y, infront2 = !behind2 && b[j].d_hi < o.x;
            may = may || (!behind2 && !infront2);
            infront = infront || infront2;
          }
          if (may) ex |= 1u << j;
          else if (infront) front |= 1u << j;
        }
        if (!kFresh && front) {
          if (kVec == 4) {
            const float4 c = *reinterpret_cast<const float4*>(color_vol + voxel_idx);
            const float4 w = *reinterpret_cast<const float4*>(weight_vol + voxel_idx);
            const unsigned int touched = ((c.x != 0.f || w.x != 0.f) ? 1u : 0u) | ((c.y != 0.f || w.y != 0.f) ? 2u : 0u) |
                                         ((c.z != 0.f || w.z != 0.f) ? 4u : 0u) | ((c.w != 0.f || w.w != 0.f) ? 8u : 0u);
            ex |= front & touched;
          } else if (color_vol[voxel_idx] != 0.f || weight_vol[voxel_idx] != 0.f) {
            ex |= 1u;
          }
        }
      }
      if (!kFresh) {
        // nothing to reset
      } else if (kVec == 4) {
        if (ex == 0) {
          const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(tsdf_vol + voxel_idx) = one;
          *reinterpret_cast<float4*>(weight_vol + voxel_idx) = zero;
          *reinterpret_cast<float4*>(color_vol + voxel_idx) = zero;
          *reinterpret_cast<float4*>(rem_vol + voxel_idx) = zero;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (!(ex >> j & 1)) { tsdf_vol[voxel_idx + j] = 1.f; weight_vol[voxel_idx + j] = 0.f; color_vol[voxel_idx + j] = 0.f; rem_vol[voxel_idx + j] = 0.f; }
        }
      } else if (ex == 0) {
        tsdf_vol[voxel_idx] = 1.f; weight_vol[voxel_idx] = 0.f; color_vol[voxel_idx] = 0.f; rem_vol[voxel_idx] = 0.f;
      }
    }
    if (__any_sync(0xffffffffu, ex != 0)) {                          // queue the exact-path voxels: one atomic per warp
      unsigned int m[kVec];
      int total = 0;
#pragma unroll
      for (int j = 0; j < kVec; ++j) { m[j] = __ballot_sync(0xffffffffu, ex >> j & 1); total += __popc(m[j]); }
      int pos = 0;
      if (lane == 0) pos = atomicAdd(&s_nq, total);
      pos = __shfl_sync(0xffffffffu, pos, 0);
#pragma unroll
      for (int j = 0; j < kVec; ++j) {
        if (ex >> j & 1) s_q[pos + __popc(m[j] & lt)] = voxel_idx + j;
        pos += __popc(m[j]);
      }
    }
  }
  __syncthreads();
  const int nq = s_nq;
  for (int i = threadIdx.x; i < nq; i += kThreads)
    tsdf_voxel<true, kFresh>(s_q[i], tsdf_vol, weight_vol, color_vol, rem_vol, P, color_im, depth_im, rem_im, col_px);
}

// ---- blocked / sparse volumes: per z column, the interval of voxels that exist -------------------------------------
//
// The reference allocates, uploads and sweeps four dense volumes per scan (fusion_lidar.py:45 carries the TODO "use
// larger voxel volume ... by splitting"): 284 M voxels at config 1, of which one integration changes 0.35 %.  Here a
// volume may be SPARSE: hull[column] = [z_lo, z_hi] (packed int16 pair, 4 B per z column) is the interval of the
// column's voxels that are materialised in the four arrays; every voxel outside its column's hull is, by definition,
// in the initial state (tsdf 1, weight / colour / remission 0, fusion_lidar.py:48-52) and its memory is never touched.
//   k_tsdf_rows   per image row: tangents of the row's pitch interval (widened by the arcsine / rounding error)
//   k_tsdf_hull   per z column: the interval of heights at which a voxel can fall into the shell (lo, hi] of a pixel of
//                 the column's image column -- rows are intervals of z / rho, the shell an interval of z^2 -- i.e. a
//                 superset of the voxels the reference would change in a never-written column; new hull = old U that
//   k_tsdf_hull_sweep  the shell sweep of k_tsdf_fresh_shell restricted to the hulls: outside, nothing is computed and
//                 nothing is stored; inside, a voxel that enters the hull with this scan is fresh (initial values
//                 written, or the reference arithmetic on them), one that was in it is integrated like a later scan
//   k_tsdf_densify  writes the initial values outside the hulls (get_volume(), the dense API)
// Config 1: 937 us (reset + integrate, 4.5 GB written) -> ~150 us; bit-identical volumes after k_tsdf_densify
// (tests/test_sparse_tsdf_gpu.py, and against the reference's own CUDA kernel).
constexpr int kHullEmpty = 1;   // lo = 1, hi = 0
constexpr int kRangeBins = 256; // bins of horizontal distance in the per-image-column row-range table
__host__ __device__ __forceinline__ int hull_pack(int lo, int hi) { return (lo & 0xffff) | (hi << 16); }
__device__ __forceinline__ void hull_unpack(int h, int& lo, int& hi) { lo = h & 0xffff; hi = (h >> 16) & 0xffff; }

__global__ void k_tsdf_rows(float4* __restrict__ rows, int H, double fov, double fov_down_abs, double e_p) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= H) return;
  // row r <=> proj_y in [r, r + 1) <=> pitch in (fov (1 - (r + 1) / H) - |fd|, fov (1 - r / H) - |fd|]  (:135-136, 142-144)
  const double p_hi = fov * (1.0 - (double)r / H) - fov_down_abs + e_p;
  const double p_lo = fov * (1.0 - (double)(r + 1) / H) - fov_down_abs - e_p;
  const double t_lo = tan(p_lo), t_hi = tan(p_hi);
  // rho = depth cos(pitch): bounds of the cosine over the row (for the per-image-column range table)
  const double a_max = fmax(fabs(p_lo), fabs(p_hi)), a_min = (p_lo <= 0.0 && p_hi >= 0.0) ? 0.0 : fmin(fabs(p_lo), fabs(p_hi));
  rows[r] = make_float4((float)(t_lo - fabs(t_lo) * 1e-6 - 1e-7), (float)(t_hi + fabs(t_hi) * 1e-6 + 1e-7),
                        (float)(cos(a_max) * (1.0 - 1e-6)), (float)(cos(a_min) * (1.0 + 1e-6)));
}

// Per image column and per bin of horizontal distance rho: the range of image rows whose shell reaches that distance
// (rho = depth cos(pitch) over the shell's depths and the row's pitches).  A z column looks its (image column, rho bin)
// up and visits those rows only -- mostly none (free space, or nothing observed there) or one or two (the ground).
// rmin[px * n_bins + b] (initialised to a large value), rmax1[...] = largest row + 1 (initialised to 0).
__global__ void __launch_bounds__(kThreads)
k_tsdf_range_table(const float2* __restrict__ shell, const float4* __restrict__ rows, int H, int W, float bin_inv, int n_bins,
                   int* __restrict__ rmin, int* __restrict__ rmax1) {
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= H * W) return;
  const int px = p / H, r = p - px * H;               // the shell image is column-major
  const float2 sh = __ldg(shell + p);
  if (sh.y < sh.x) return;
  const float4 t = __ldg(rows + r);
  const float lo = fmaxf(sh.x, 0.f) * (1.f - 3e-5f), hi = sh.y * (1.f + 3e-5f);
  if (!(hi >= 0.f)) return;
  const int b0 = max(0, (int)floorf((lo * t.z - 1e-3f) * bin_inv) - 1);
  const float top = (hi * t.w + 1e-3f) * bin_inv;
  const int b1 = top < (float)(n_bins - 1) ? (int)floorf(top) + 1 : n_bins - 1;    // also inf / NaN
  for (int b = b0; b <= min(b1, n_bins - 1); ++b) {
    atomicMin(rmin + px * n_bins + b, r);
    atomicMax(rmax1 + px * n_bins + b, r + 1);
  }
}

__global__ void __launch_bounds__(kThreads)
k_tsdf_hull(const TsdfParams P, const int* __restrict__ col_px, const float2* __restrict__ shell,
            const float4* __restrict__ rows, int* __restrict__ hull, int* __restrict__ hull_old, int fresh, int window_cols,
            float row_a, float row_b, float bin_inv, int n_bins, const int* __restrict__ rmin, const int* __restrict__ rmax1) {
  const int c = blockIdx.x * kThreads + threadIdx.x;
  if (c >= P.dx * P.dy) return;
  const int vx = c / P.dy, vy = c - vx * P.dy;
  int old = fresh ? kHullEmpty : hull[c];
  hull_old[c] = old;
  const float x = __fmaf_rn((float)vx, P.voxel_size, P.ox), y = __fmaf_rn((float)vy, P.voxel_size, P.oy);
  const float xy2 = __fmaf_rn(x, x, y * y);
  const float rho = sqrtf(xy2);
  int lo = 1, hi = 0;
  // columns whose voxels lie within kDecodeWindow of a slab boundary are evaluated at the (float-decoded) position of
  // another column (:96-98): whole column; so is the column through the sensor (pitch +-90 degrees, NaN at the origin)
  if (vy < window_cols || vy >= P.dy - window_cols || !(rho > 1e-3f)) {
    lo = 0; hi = P.dz - 1;
  } else {
    const float z_bot = P.oz - P.voxel_size, z_top = P.oz + (float)P.dz * P.voxel_size;
    const float m = rho * 2e-6f + 1e-5f;
    float zmin = INFINITY, zmax = -INFINITY;
    const int px = __ldg(col_px + c);
    const float2* __restrict__ col = shell + (size_t)px * P.im_h;
    // rows of this image column whose shell reaches the column's distance at all (k_tsdf_range_table) ...
    const int bin = min(n_bins - 1, (int)(rho * bin_inv));
    int r_first = __ldg(rmin + px * n_bins + bin), r_last = __ldg(rmax1 + px * n_bins + bin) - 1;
    if (r_first <= r_last) {
      // ... and that the column's voxels can fall into: row(z) = H (1 - (atan(z / rho) + |fd|) / fov), two rows of slack
      const float rho_inv = 1.f / rho;
      r_first = max(r_first, (int)floorf(row_a - row_b * atanf(z_top * rho_inv)) - 2);
      r_last = min(r_last, (int)floorf(row_a - row_b * atanf(z_bot * rho_inv)) + 2);
    }
    for (int r = r_first; r <= r_last; ++r) {
      const float2 sh = __ldg(col + r);
      if (sh.y < sh.x) continue;         // empty pixel (lo = +inf, hi = -inf)
      const float4 t = __ldg(rows + r);
      float za = rho * t.x - m, zb = rho * t.y + m;
      if (zb < z_bot) break;             // the rows below look further down still
      if (za > z_top) continue;
      za = fmaxf(za, z_bot); zb = fminf(zb, z_top);
      const float h2 = sh.y * (1.f + 3e-5f);
      const float B = h2 * h2 - xy2;     // depth <= hi  <=>  z^2 <= hi^2 - rho^2
      if (B < 0.f) continue;
      const float l2 = sh.x * (1.f - 3e-5f);
      const float A = l2 > 0.f ? l2 * l2 - xy2 : -1.f;   // depth >= lo  <=>  z^2 >= lo^2 - rho^2: a hole around z = 0
      // cheap rejections in z^2 before any square root: the interval lies inside the hole (free space in front of the
      // surface: most rows of most columns) or beyond the far bound
      const float za2 = za * za, zb2 = zb * zb;
      if (A > 0.f && fmaxf(za2, zb2) * (1.f + 1e-5f) < A) continue;
      if (!(za <= 0.f && zb >= 0.f) && fminf(za2, zb2) > B * (1.f + 1e-5f) + 1e-6f) continue;
      const float sB = sqrtf(B) * (1.f + 1e-6f) + 1e-6f;
      float a = fmaxf(za, -sB), b = fminf(zb, sB);
      if (a > b) continue;
      if (A > 0.f) {
        const float sA = sqrtf(A) * (1.f - 1e-6f) - 1e-6f;
        if (sA > 0.f) {
          const float bl = fminf(b, -sA), ar = fmaxf(a, sA);
          const bool left = a <= bl, right = ar <= b;
          if (!left && !right) continue;
          const float a2 = left ? a : ar, b2 = right ? b : bl;
          a = a2; b = b2;
        }
      }
      zmin = fminf(zmin, a);
      zmax = fmaxf(zmax, b);
    }
    if (zmin <= zmax) {
      const float inv = 1.f / P.voxel_size;
      lo = max(0, (int)floorf((zmin - P.oz) * inv) - 1);
      hi = min(P.dz - 1, (int)ceilf((zmax - P.oz) * inv) + 1);
      if (lo > hi) { lo = 1; hi = 0; }
    }
  }
  int olo, ohi;
  hull_unpack(old, olo, ohi);
  if (olo <= ohi) {                      // union with what the column holds already (an interval: gaps are materialised)
    if (lo > hi) { lo = olo; hi = ohi; }
    else { lo = min(lo, olo); hi = max(hi, ohi); }
  }
  hull[c] = hull_pack(lo, hi);
}

// The sweep of k_tsdf_fresh_shell over the hulls only.  A CTA owns 256 consecutive z columns; the groups of kVec
// z-consecutive voxels that intersect their column's hull are ENUMERATED (block scan of the per-column group counts,
// each thread finds its column by bisection in shared memory), so that every lane has a group inside a hull -- with
// threads tied to fixed voxel positions a warp would run the whole bracket for the 2-3 lanes that sit in the ground's
// hull.  Queue entries carry bit 31 when the voxel is NEW to its hull (fresh: nothing is read, everything is
// written); the others are integrated like a later scan.
constexpr int kHullCols = 256;   // columns per CTA (= kThreads)

#ifndef VL_HULL_MINB
#define VL_HULL_MINB 6   // 40 registers instead of 53: later integrations (which read the volumes) 641 -> 586 us per five scans, a first integration unchanged
#endif
template <int kVec>
__global__ void __launch_bounds__(kThreads, VL_HULL_MINB)
k_tsdf_hull_sweep(float* __restrict__ tsdf_vol, float* __restrict__ weight_vol, float* __restrict__ color_vol,
                  float* __restrict__ rem_vol, const TsdfParams P, const ShellParams S,
                  const float* __restrict__ color_im, const float* __restrict__ depth_im,
                  const float* __restrict__ rem_im, const int* __restrict__ col_px, const float2* __restrict__ shell,
                  const int* __restrict__ hull, const int* __restrict__ hull_old) {
  __shared__ unsigned int s_q[kThreads * kVec];
  __shared__ int s_off[kHullCols + 1];
  __shared__ int s_hull[kHullCols], s_old[kHullCols];
  __shared__ int s_warp[kThreads / 32];
  const int n_cols = P.dx * P.dy;
  const int c0 = blockIdx.x * kHullCols;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned int lt = (1u << lane) - 1;
  // per-column group counts and their exclusive scan
  int ng = 0;
  {
    const int c = c0 + threadIdx.x;
    int h = kHullEmpty, o = kHullEmpty;
    if (c < n_cols) { h = __ldg(hull + c); o = __ldg(hull_old + c); }
    s_hull[threadIdx.x] = h; s_old[threadIdx.x] = o;
    int lo, hi;
    hull_unpack(h, lo, hi);
    if (lo <= hi) ng = hi / kVec - lo / kVec + 1;
  }
  int incl = ng;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  int wbase = 0;
#pragma unroll
  for (int k = 0; k < kThreads / 32; ++k) wbase += k < w ? s_warp[k] : 0;
  s_off[threadIdx.x] = wbase + incl - ng;
  if (threadIdx.x == kThreads - 1) s_off[kHullCols] = wbase + incl;
  __syncthreads();
  const int total = s_off[kHullCols];
  for (int base = w * 32; base < total; base += kThreads) {    // warp-uniform: warp w takes groups [base, base + 32)
    const int i = base + lane;
    unsigned int ex = 0, fresh_m = 0;
    int voxel_idx = 0;
    if (i < total) {
      int j = 0;                                           // the last column whose offset is <= i
#pragma unroll
      for (int step = kHullCols / 2; step > 0; step >>= 1) if (s_off[j + step] <= i) j += step;
      int hlo, hhi, olo, ohi;
      hull_unpack(s_hull[j], hlo, hhi);
      hull_unpack(s_old[j], olo, ohi);
      const int c = c0 + j;
      const int vx = c / P.dy, vy = c - vx * P.dy;
      const int vz = (hlo / kVec + (i - s_off[j])) * kVec;
      const int rem_i = vy * P.dz + vz;
      voxel_idx = c * P.dz + vz;
      unsigned int in_m = 0;
#pragma unroll
      for (int jj = 0; jj < kVec; ++jj) {
        const int z = vz + jj;
        if (z >= hlo && z <= hhi) { in_m |= 1u << jj; if (!(z >= olo && z <= ohi)) fresh_m |= 1u << jj; }
      }
      unsigned int front = 0;
      if (rem_i < kDecodeWindow || rem_i + kVec > S.slab - kDecodeWindow) {
        ex = in_m;
      } else {
        const float x = __fmaf_rn((float)vx, P.voxel_size, P.ox);
        const float y = __fmaf_rn((float)vy, P.voxel_size, P.oy);
        const float xy2 = __fmaf_rn(x, x, y * y);
        const int px = __ldg(col_px + c);
        const float2* __restrict__ col = shell + px * P.im_h;
        VoxelBracket b[kVec];
        float2 lh[kVec];
#pragma unroll
        for (int jj = 0; jj < kVec; ++jj)
          b[jj] = shell_bracket(xy2, __fmaf_rn((float)(vz + jj), P.voxel_size, P.oz), P, S);
#pragma unroll
        for (int jj = 0; jj < kVec; ++jj) lh[jj] = __ldg(col + max(b[jj].r0, 0));
#pragma unroll
        for (int jj = 0; jj < kVec; ++jj) {
          if (b[jj].r0 < 0 || !(in_m >> jj & 1)) continue;
          bool behind = b[jj].d_lo > lh[jj].y;
          bool infront = !behind && b[jj].d_hi < lh[jj].x;
          bool may = !behind && !infront;
          if (b[jj].r1 != b[jj].r0) {
            const float2 o = __ldg(col + b[jj].r1);
            const bool behind2 = b[jj].d_lo > o.y, infront2 = !behind2 && b[jj].d_hi < o.x;
            may = may || (!behind2 && !infront2);
            infront = infront || infront2;
          }
          if (may) ex |= 1u << jj;
          else if (infront) front |= 1u << jj;
        }
        front &= ~fresh_m;                                           // free space matters only for voxels that hold data
        if (front) {
#pragma unroll
          for (int jj = 0; jj < kVec; ++jj)
            if ((front >> jj & 1) && (color_vol[voxel_idx + jj] != 0.f || weight_vol[voxel_idx + jj] != 0.f)) ex |= 1u << jj;
        }
      }
      // voxels that enter the hull without being evaluated get the initial values
      const unsigned int init_m = fresh_m & ~ex;
      if (kVec == 4 && init_m == 15u) {
        const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(tsdf_vol + voxel_idx) = one;
        *reinterpret_cast<float4*>(weight_vol + voxel_idx) = zero;
        *reinterpret_cast<float4*>(color_vol + voxel_idx) = zero;
        *reinterpret_cast<float4*>(rem_vol + voxel_idx) = zero;
      } else if (init_m) {
#pragma unroll
        for (int jj = 0; jj < kVec; ++jj)
          if (init_m >> jj & 1) { tsdf_vol[voxel_idx + jj] = 1.f; weight_vol[voxel_idx + jj] = 0.f; color_vol[voxel_idx + jj] = 0.f; rem_vol[voxel_idx + jj] = 0.f; }
      }
    }
    // the exact-path voxels of this warp's 32 groups, compacted into the warp's own queue and evaluated right away on
    // dense lanes: no CTA barrier inside the loop (a CTA-wide queue cost 8.9 barrier-stall cycles per issue)
    if (__any_sync(0xffffffffu, ex != 0)) {
      unsigned int* q = s_q + w * (32 * kVec);
      unsigned int m[kVec];
      int pos = 0;
#pragma unroll
      for (int jj = 0; jj < kVec; ++jj) {
        m[jj] = __ballot_sync(0xffffffffu, ex >> jj & 1);
        if (ex >> jj & 1) q[pos + __popc(m[jj] & lt)] = (unsigned int)(voxel_idx + jj) | ((fresh_m >> jj & 1) ? 0x80000000u : 0u);
        pos += __popc(m[jj]);
      }
      __syncwarp();
      for (int k = lane; k < pos; k += 32) {
        const unsigned int e = q[k];
        if (e & 0x80000000u)
          tsdf_voxel<true, true>((int)(e & 0x7fffffffu), tsdf_vol, weight_vol, color_vol, rem_vol, P, color_im, depth_im, rem_im, col_px);
        else
          tsdf_voxel<true, false>((int)e, tsdf_vol, weight_vol, color_vol, rem_vol, P, color_im, depth_im, rem_im, col_px);
      }
      __syncwarp();
    }
  }
}

// initial values for every voxel outside its column's hull; afterwards the volume is dense (hull = whole column)
__global__ void __launch_bounds__(kThreads)
k_tsdf_densify(float* __restrict__ tsdf_vol, float* __restrict__ weight_vol, float* __restrict__ color_vol,
               float* __restrict__ rem_vol, int n_cols, int dz, int* __restrict__ hull) {
  // one warp per z column: lanes stride over z
  const int c = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (c >= n_cols) return;
  int lo, hi;
  const int h = hull[c];
  hull_unpack(h, lo, hi);
  const size_t base = (size_t)c * dz;
  for (int z = threadIdx.x & 31; z < dz; z += 32) {
    if (z >= lo && z <= hi) continue;
    tsdf_vol[base + z] = 1.f; weight_vol[base + z] = 0.f; color_vol[base + z] = 0.f; rem_vol[base + z] = 0.f;
  }
  __syncwarp();
  if ((threadIdx.x & 31) == 0) hull[c] = hull_pack(0, dz - 1);
}

__global__ void k_tsdf_hull_fill(int* __restrict__ hull, int n_cols, int value) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_cols) hull[c] = value;
}

int g_tsdf_shell = 1;    // vl_debug_tsdf_shell: 0 off, 1 on, 2 on with one voxel per thread
int g_tsdf_scalar = 0;

}  // namespace

extern "C" void vl_debug_tsdf_shell(int mode) { g_tsdf_shell = mode ? 1 : 0; g_tsdf_scalar = mode == 2; }

extern "C" int vl_tsdf_init(float* d_tsdf, float* d_weight, float* d_color, float* d_rem, long long n_voxels,
                            vl_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_voxels < 0 || !d_tsdf || !d_weight || !d_color || !d_rem) {
    vl_set_error("vl_tsdf_init: invalid argument");
    return VL_EINVAL;
  }
  if (n_voxels == 0) return VL_OK;
  const bool aligned = ((((uintptr_t)d_tsdf) | ((uintptr_t)d_weight) | ((uintptr_t)d_color) | ((uintptr_t)d_rem)) & 15) == 0;
  const long long n4 = aligned ? n_voxels / 4 : 0;
  long long want = (n_voxels / 4 + kThreads - 1) / kThreads;
  int nb = (int)(want < 1 ? 1 : (want > vl_sm_count() * 16LL ? vl_sm_count() * 16LL : want));
  VlProfScope ps(VL_ST_TSDF_INIT, stream);
  k_tsdf_init<<<nb, kThreads, 0, stream>>>(reinterpret_cast<float4*>(d_tsdf), reinterpret_cast<float4*>(d_weight),
                                          reinterpret_cast<float4*>(d_color), reinterpret_cast<float4*>(d_rem), n4,
                                          d_tsdf, d_weight, d_color, d_rem, n_voxels);
  VL_LAUNCH_CHECK("k_tsdf_init");
  return VL_OK;
}

static int tsdf_integrate_impl(float* d_tsdf, float* d_weight, float* d_color, float* d_rem, int dx, int dy, int dz,
                               const float vol_origin[3], float voxel_size, float trunc_margin, float obs_weight,
                               float fov_up_deg, float fov_down_deg, const float* d_color_im, const float* d_depth_im,
                               const float* d_rem_im, int im_h, int im_w, void* d_workspace, size_t workspace_bytes,
                               bool fresh, cudaStream_t stream) {
  const long long n_vox = (long long)dx * dy * dz;
  if (dx <= 0 || dy <= 0 || dz <= 0 || n_vox > 0x7fffffffLL || im_h <= 0 || im_w <= 0 || !vol_origin || !d_tsdf ||
      !d_weight || !d_color || !d_rem || !d_color_im || !d_depth_im || !d_rem_im) {
    vl_set_error("vl_tsdf_integrate: invalid argument (dims %d x %d x %d, image %d x %d)", dx, dy, dz, im_h, im_w);
    return VL_EINVAL;
  }
  if (d_workspace && workspace_bytes < sizeof(int) * (size_t)dx * dy) {
    vl_set_error("vl_tsdf_integrate_ws: workspace too small (%zu < %zu bytes)", workspace_bytes, sizeof(int) * (size_t)dx * dy);
    return VL_ENOSPACE;
  }
  const size_t shell_off = vl_align256(sizeof(int) * (size_t)dx * dy);
  TsdfParams P;
  P.dx = dx; P.dy = dy; P.dz = dz;
  P.ox = vol_origin[0]; P.oy = vol_origin[1]; P.oz = vol_origin[2];
  P.voxel_size = voxel_size; P.trunc_margin = trunc_margin; P.obs_weight = obs_weight;
  P.fov_up = fov_up_deg * VL_PI / 180.0;      // float * double / double -> float, exactly the kernel-string expression
  P.fov_down = fov_down_deg * VL_PI / 180.0;
  P.im_h = im_h; P.im_w = im_w;
  const unsigned int nb = (unsigned int)((n_vox + kThreads - 1) / kThreads);
  VlProfScope ps(VL_ST_TSDF_INTEGRATE, stream);
  if (d_workspace) {
    int* col_px = static_cast<int*>(d_workspace);
    k_tsdf_columns<<<(dx * dy + kThreads - 1) / kThreads, kThreads, 0, stream>>>(col_px, P);
    VL_LAUNCH_CHECK("k_tsdf_columns");
    // the shell sweep needs: room for the shell image, y / z decodable in float, a field of view inside the arcsine
    // series' range, image rows much coarser than its error, a positive truncation margin
    const double fov_rad = fabs((double)P.fov_up) + fabs((double)P.fov_down);
    const double eps_row = fov_rad > 0.0 ? 1.02 * kAsinErr * im_h / fov_rad + 2e-4 + 4e-7 * im_h : 1.0;
    const bool shell_ok = g_tsdf_shell && workspace_bytes >= shell_off + sizeof(float2) * (size_t)im_h * im_w &&
                          (long long)dy * dz <= (1LL << 24) && dx <= 65535 && trunc_margin > 0.f && voxel_size > 0.f &&
                          fabs((double)P.fov_up) <= 0.61 && fabs((double)P.fov_down) <= 0.61 && fov_rad > 0.0 &&
                          eps_row < 0.45;
    if (shell_ok) {
      float2* shell = reinterpret_cast<float2*>(static_cast<char*>(d_workspace) + shell_off);
      k_tsdf_shell<<<(im_h * im_w + kThreads - 1) / kThreads, kThreads, 0, stream>>>(d_depth_im, d_color_im, im_h, im_w,
                                                                                   trunc_margin, shell);
      VL_LAUNCH_CHECK("k_tsdf_shell");
      ShellParams S;
      S.slab = dy * dz;
      S.inv_dz = 1.0f / (float)dz;
      S.fov_abs_down = fabsf(P.fov_down);
      S.h_over_fov = (float)im_h / (fabsf(P.fov_up) + fabsf(P.fov_down));
      S.pitch_hi = P.fov_up + kAsinErr;
      S.pitch_lo = P.fov_down - kAsinErr;
      S.eps_row = (float)eps_row;
      dim3 grid((unsigned int)((S.slab + kFastChunk - 1) / kFastChunk), (unsigned int)dx);
      const bool vec = dz % 4 == 0 && !g_tsdf_scalar &&
                       ((((uintptr_t)d_tsdf) | ((uintptr_t)d_weight) | ((uintptr_t)d_color) | ((uintptr_t)d_rem)) & 15) == 0;
#define VL_SHELL_LAUNCH(V, F) k_tsdf_fresh_shell<V, F><<<grid, kThreads, 0, stream>>>( \
          d_tsdf, d_weight, d_color, d_rem, P, S, d_color_im, d_depth_im, d_rem_im, col_px, shell)
      if (vec) { if (fresh) VL_SHELL_LAUNCH(4, true); else VL_SHELL_LAUNCH(4, false); }
      else { if (fresh) VL_SHELL_LAUNCH(1, true); else VL_SHELL_LAUNCH(1, false); }
#undef VL_SHELL_LAUNCH
    } else if (fresh)
      k_tsdf_integrate<true, true><<<nb, kThreads, 0, stream>>>(d_tsdf, d_weight, d_color, d_rem, P, d_color_im, d_depth_im,
                                                               d_rem_im, n_vox, col_px);
    else
      k_tsdf_integrate<true, false><<<nb, kThreads, 0, stream>>>(d_tsdf, d_weight, d_color, d_rem, P, d_color_im, d_depth_im,
                                                                d_rem_im, n_vox, col_px);
  } else {
    k_tsdf_integrate<false, false><<<nb, kThreads, 0, stream>>>(d_tsdf, d_weight, d_color, d_rem, P, d_color_im, d_depth_im,
                                                               d_rem_im, n_vox, nullptr);
  }
  VL_LAUNCH_CHECK("k_tsdf_integrate");
  return VL_OK;
}

extern "C" int vl_tsdf_integrate(float* d_tsdf, float* d_weight, float* d_color, float* d_rem, int dx, int dy, int dz,
                                 const float vol_origin[3], float voxel_size, float trunc_margin, float obs_weight,
                                 float fov_up_deg, float fov_down_deg, const float* d_color_im, const float* d_depth_im,
                                 const float* d_rem_im, int im_h, int im_w, vl_stream stream_) {
  return tsdf_integrate_impl(d_tsdf, d_weight, d_color, d_rem, dx, dy, dz, vol_origin, voxel_size, trunc_margin, obs_weight,
                             fov_up_deg, fov_down_deg, d_color_im, d_depth_im, d_rem_im, im_h, im_w, nullptr, 0, false,
                             static_cast<cudaStream_t>(stream_));
}

extern "C" size_t vl_tsdf_workspace_bytes(int dx, int dy) {
  return dx > 0 && dy > 0 ? vl_align256(sizeof(int) * (size_t)dx * dy) : 256;
}

extern "C" size_t vl_tsdf_fresh_workspace_bytes(int dx, int dy, int im_h, int im_w) {
  return vl_tsdf_workspace_bytes(dx, dy) + (im_h > 0 && im_w > 0 ? vl_align256(sizeof(float2) * (size_t)im_h * im_w) : 0);
}

extern "C" int vl_tsdf_integrate_ws(float* d_tsdf, float* d_weight, float* d_color, float* d_rem, int dx, int dy, int dz,
                                    const float vol_origin[3], float voxel_size, float trunc_margin, float obs_weight,
                                    float fov_up_deg, float fov_down_deg, const float* d_color_im, const float* d_depth_im,
                                    const float* d_rem_im, int im_h, int im_w, void* d_workspace, size_t workspace_bytes,
                                    vl_stream stream_) {
  if (!d_workspace) { vl_set_error("vl_tsdf_integrate_ws: null workspace"); return VL_EINVAL; }
  return tsdf_integrate_impl(d_tsdf, d_weight, d_color, d_rem, dx, dy, dz, vol_origin, voxel_size, trunc_margin, obs_weight,
                             fov_up_deg, fov_down_deg, d_color_im, d_depth_im, d_rem_im, im_h, im_w, d_workspace,
                             workspace_bytes, false, static_cast<cudaStream_t>(stream_));
}

extern "C" int vl_tsdf_init_integrate(float* d_tsdf, float* d_weight, float* d_color, float* d_rem, int dx, int dy, int dz,
                                      const float vol_origin[3], float voxel_size, float trunc_margin, float obs_weight,
                                      float fov_up_deg, float fov_down_deg, const float* d_color_im, const float* d_depth_im,
                                      const float* d_rem_im, int im_h, int im_w, void* d_workspace, size_t workspace_bytes,
                                      vl_stream stream_) {
  if (!d_workspace) { vl_set_error("vl_tsdf_init_integrate: null workspace"); return VL_EINVAL; }
  return tsdf_integrate_impl(d_tsdf, d_weight, d_color, d_rem, dx, dy, dz, vol_origin, voxel_size, trunc_margin, obs_weight,
                             fov_up_deg, fov_down_deg, d_color_im, d_depth_im, d_rem_im, im_h, im_w, d_workspace,
                             workspace_bytes, true, static_cast<cudaStream_t>(stream_));
}


// ---- sparse volumes (see k_tsdf_hull) -----------------------------------------------------------------------------
extern "C" size_t vl_tsdf_sparse_workspace_bytes(int dx, int dy, int im_h, int im_w) {
  if (dx <= 0 || dy <= 0 || im_h <= 0 || im_w <= 0) return 256;
  return vl_tsdf_fresh_workspace_bytes(dx, dy, im_h, im_w) + vl_align256(sizeof(int) * (size_t)dx * dy) +
         vl_align256(sizeof(float4) * (size_t)im_h) + 2 * vl_align256(sizeof(int) * (size_t)im_w * kRangeBins);
}

extern "C" int vl_tsdf_densify(float* d_tsdf, float* d_weight, float* d_color, float* d_rem, int dx, int dy, int dz,
                               int* d_hull, vl_stream stream_) {
  if (dx <= 0 || dy <= 0 || dz <= 0 || dz > 32767 || !d_tsdf || !d_weight || !d_color || !d_rem || !d_hull) {
    vl_set_error("vl_tsdf_densify: invalid argument");
    return VL_EINVAL;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int n_cols = dx * dy;
  VlProfScope ps(VL_ST_TSDF_INIT, stream);
  k_tsdf_densify<<<(n_cols + kThreads / 32 - 1) / (kThreads / 32), kThreads, 0, stream>>>(d_tsdf, d_weight, d_color, d_rem, n_cols, dz, d_hull);
  VL_LAUNCH_CHECK("k_tsdf_densify");
  return VL_OK;
}

extern "C" int vl_tsdf_sparse_integrate(float* d_tsdf, float* d_weight, float* d_color, float* d_rem, int dx, int dy, int dz,
                                        const float vol_origin[3], float voxel_size, float trunc_margin, float obs_weight,
                                        float fov_up_deg, float fov_down_deg, const float* d_color_im,
                                        const float* d_depth_im, const float* d_rem_im, int im_h, int im_w, int* d_hull,
                                        int flags, void* d_workspace, size_t workspace_bytes, vl_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int fresh = flags & VL_TSDF_FRESH;
  // VL_TSDF_TABLES_VALID: the caller's promise that this workspace still holds the per-column pixel table and the per-row
  // tangents of an earlier call with the same volume geometry, field of view and image size (they depend on nothing else)
  const bool tables_valid = (flags & VL_TSDF_TABLES_VALID) != 0;
  const long long n_vox = (long long)dx * dy * dz;
  if (dx <= 0 || dy <= 0 || dz <= 0 || n_vox > 0x7fffffffLL || im_h <= 0 || im_w <= 0 || !vol_origin || !d_tsdf ||
      !d_weight || !d_color || !d_rem || !d_color_im || !d_depth_im || !d_rem_im || !d_hull || !d_workspace || dz > 32767) {
    vl_set_error("vl_tsdf_sparse_integrate: invalid argument (dims %d x %d x %d, image %d x %d)", dx, dy, dz, im_h, im_w);
    return VL_EINVAL;
  }
  if (workspace_bytes < vl_tsdf_sparse_workspace_bytes(dx, dy, im_h, im_w)) {
    vl_set_error("vl_tsdf_sparse_integrate: workspace too small (%zu < %zu bytes)", workspace_bytes,
                 vl_tsdf_sparse_workspace_bytes(dx, dy, im_h, im_w));
    return VL_ENOSPACE;
  }
  const int n_cols = dx * dy;
  TsdfParams P;
  P.dx = dx; P.dy = dy; P.dz = dz;
  P.ox = vol_origin[0]; P.oy = vol_origin[1]; P.oz = vol_origin[2];
  P.voxel_size = voxel_size; P.trunc_margin = trunc_margin; P.obs_weight = obs_weight;
  P.fov_up = fov_up_deg * VL_PI / 180.0;
  P.fov_down = fov_down_deg * VL_PI / 180.0;
  P.im_h = im_h; P.im_w = im_w;
  const double fov_rad = fabs((double)P.fov_up) + fabs((double)P.fov_down);
  const double eps_row = fov_rad > 0.0 ? 1.02 * kAsinErr * im_h / fov_rad + 2e-4 + 4e-7 * im_h : 1.0;
  const bool ok = g_tsdf_shell && (long long)dy * dz <= (1LL << 24) && dx <= 65535 && trunc_margin > 0.f && voxel_size > 0.f &&
                  fabs((double)P.fov_up) <= 0.61 && fabs((double)P.fov_down) <= 0.61 && fov_rad > 0.0 && eps_row < 0.45 &&
                  P.fov_up >= 0.f && P.fov_down <= 0.f;
  if (!ok) {   // outside the sweep's limits: the dense path, then the volume is dense
    int rc = VL_OK;
    if (!fresh) rc = vl_tsdf_densify(d_tsdf, d_weight, d_color, d_rem, dx, dy, dz, d_hull, stream_);
    if (rc) return rc;
    rc = tsdf_integrate_impl(d_tsdf, d_weight, d_color, d_rem, dx, dy, dz, vol_origin, voxel_size, trunc_margin, obs_weight,
                             fov_up_deg, fov_down_deg, d_color_im, d_depth_im, d_rem_im, im_h, im_w, d_workspace,
                             workspace_bytes, fresh != 0, stream);
    if (rc) return rc;
    k_tsdf_hull_fill<<<(n_cols + 255) / 256, 256, 0, stream>>>(d_hull, n_cols, hull_pack(0, dz - 1));
    VL_LAUNCH_CHECK("k_tsdf_hull_fill");
    return VL_OK;
  }
  char* ws = static_cast<char*>(d_workspace);
  int* col_px = reinterpret_cast<int*>(ws);
  const size_t shell_off = vl_align256(sizeof(int) * (size_t)n_cols);
  float2* shell = reinterpret_cast<float2*>(ws + shell_off);
  const size_t old_off = vl_tsdf_fresh_workspace_bytes(dx, dy, im_h, im_w);
  int* hull_old = reinterpret_cast<int*>(ws + old_off);
  const size_t rows_off = old_off + vl_align256(sizeof(int) * (size_t)n_cols);
  float4* rows = reinterpret_cast<float4*>(ws + rows_off);
  const size_t tab_bytes = vl_align256(sizeof(int) * (size_t)im_w * kRangeBins);
  int* rmin = reinterpret_cast<int*>(ws + rows_off + vl_align256(sizeof(float4) * (size_t)im_h));
  int* rmax1 = reinterpret_cast<int*>(reinterpret_cast<char*>(rmin) + tab_bytes);
  VlProfScope ps(VL_ST_TSDF_INTEGRATE, stream);
  if (!tables_valid) {
    k_tsdf_columns<<<(n_cols + kThreads - 1) / kThreads, kThreads, 0, stream>>>(col_px, P);
    VL_LAUNCH_CHECK("k_tsdf_columns");
  }
  k_tsdf_shell<<<(im_h * im_w + kThreads - 1) / kThreads, kThreads, 0, stream>>>(d_depth_im, d_color_im, im_h, im_w, trunc_margin, shell);
  VL_LAUNCH_CHECK("k_tsdf_shell");
  // pitch slack of a row boundary: the sweep's row tolerance (eps_row rows) as an angle, plus rounding
  const double e_p = (eps_row + 1e-2) * fov_rad / im_h + 2e-5;
  if (!tables_valid) {
    k_tsdf_rows<<<(im_h + 127) / 128, 128, 0, stream>>>(rows, im_h, (double)(fabsf(P.fov_up) + fabsf(P.fov_down)), fabs((double)P.fov_down), e_p);
    VL_LAUNCH_CHECK("k_tsdf_rows");
  }
  // largest horizontal distance of a column from the sensor axis -> bin width
  const double xm = fmax(fabs((double)P.ox), fabs((double)P.ox + dx * (double)voxel_size));
  const double ym = fmax(fabs((double)P.oy), fabs((double)P.oy + dy * (double)voxel_size));
  const float bin_inv = (float)((kRangeBins - 1) / (sqrt(xm * xm + ym * ym) * 1.001 + 1e-3));
  VL_CUDA_CHECK(cudaMemsetAsync(rmin, 0x7f, tab_bytes, stream));
  VL_CUDA_CHECK(cudaMemsetAsync(rmax1, 0, tab_bytes, stream));
  k_tsdf_range_table<<<(im_h * im_w + kThreads - 1) / kThreads, kThreads, 0, stream>>>(shell, rows, im_h, im_w, bin_inv, kRangeBins,
                                                                                      rmin, rmax1);
  VL_LAUNCH_CHECK("k_tsdf_range_table");
  const int window_cols = (kDecodeWindow + dz - 1) / dz + 1;
  // row(z) = H (1 - (atan(z / rho) + |fd|) / fov) = row_a - row_b atan(z / rho)
  const float row_b = (float)(im_h / fov_rad), row_a = (float)(im_h * (1.0 - fabs((double)P.fov_down) / fov_rad));
  k_tsdf_hull<<<(n_cols + kThreads - 1) / kThreads, kThreads, 0, stream>>>(P, col_px, shell, rows, d_hull, hull_old, fresh ? 1 : 0,
                                                                            window_cols, row_a, row_b, bin_inv, kRangeBins, rmin, rmax1);
  VL_LAUNCH_CHECK("k_tsdf_hull");
  ShellParams S;
  S.slab = dy * dz;
  S.inv_dz = 1.0f / (float)dz;
  S.fov_abs_down = fabsf(P.fov_down);
  S.h_over_fov = (float)im_h / (fabsf(P.fov_up) + fabsf(P.fov_down));
  S.pitch_hi = P.fov_up + kAsinErr;
  S.pitch_lo = P.fov_down - kAsinErr;
  S.eps_row = (float)eps_row;
  const unsigned int grid = (unsigned int)((n_cols + kHullCols - 1) / kHullCols);
  const bool vec = dz % 4 == 0 && !g_tsdf_scalar &&
                   ((((uintptr_t)d_tsdf) | ((uintptr_t)d_weight) | ((uintptr_t)d_color) | ((uintptr_t)d_rem)) & 15) == 0;
  if (vec)
    k_tsdf_hull_sweep<4><<<grid, kThreads, 0, stream>>>(d_tsdf, d_weight, d_color, d_rem, P, S, d_color_im, d_depth_im, d_rem_im,
                                                       col_px, shell, d_hull, hull_old);
  else
    k_tsdf_hull_sweep<1><<<grid, kThreads, 0, stream>>>(d_tsdf, d_weight, d_color, d_rem, P, S, d_color_im, d_depth_im, d_rem_im,
                                                       col_px, shell, d_hull, hull_old);
  VL_LAUNCH_CHECK("k_tsdf_hull_sweep");
  return VL_OK;
}
