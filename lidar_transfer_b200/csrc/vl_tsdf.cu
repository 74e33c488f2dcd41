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
  int nb = (int)(want < 1 ? 1 : (want > 148LL * 16 ? 148LL * 16 : want));
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
