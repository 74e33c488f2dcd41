// ref_tsdf_harness.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Runs the reference's own CUDA `integrate` kernel source (auxiliary/fusion_lidar.py:66-229,
// extracted verbatim at build time into oracle/_ref/integrate_kernel.inc by
// oracle/extract_tsdf_kernel.py) on the CPU: a minimal shim supplies the CUDA
// built-ins the kernel body uses.  pycuda and a GPU are absent in the authoring
// container, so this is the closest thing to executing the reference for row (iv);
// it pins oracle/vl_oracle.c:vlo_tsdf_integrate bit for bit (both use the host libm).
// Built with -mfma -ffp-contract=fast to mirror nvcc's default -fmad=true.
#include <math.h>
#include <algorithm>
using std::max;
using std::min;

struct vl_dim3 { int x, y, z; };
static thread_local vl_dim3 blockIdx, threadIdx;
static vl_dim3 gridDim, blockDim;
#define __global__ static
static inline float norm3df(float a, float b, float c) {
  return (float)sqrt((double)a * a + (double)b * b + (double)c * c);
}

#include "integrate_kernel.inc"

// Enumerates voxel_idx = block*1024 + thread over [0, n_vox) exactly like the
// launch at fusion_lidar.py:267-287 (block=(1024,1,1), 3-D grid, host loop index in
// other_params[0]); a single 1-D grid and loop index 0 produce the same voxel_idx set.
// Threads with voxel_idx == n_vox (let through by the kernel's `>` guard, :92) are not run.
extern "C" void ref_tsdf_integrate(float* tsdf_vol, float* weight_vol, float* color_vol,
                                   float* rem_vol, float* vol_dim, float* vol_origin,
                                   float* cam_pose, float* other_params, float* color_im,
                                   float* depth_im, float* rem_im) {
  const long long n_vox = (long long)(int)vol_dim[0] * (int)vol_dim[1] * (int)vol_dim[2];
  blockDim = {1024, 1, 1};
  const int n_blocks = (int)((n_vox + 1023) / 1024);
  gridDim = {n_blocks, 1, 1};
  other_params[0] = 0.0f;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < n_blocks; ++b) {
    blockIdx = {b, 0, 0};
    for (int t = 0; t < 1024; ++t) {
      if ((long long)b * 1024 + t >= n_vox) break;
      threadIdx = {t, 0, 0};
      integrate(tsdf_vol, weight_vol, color_vol, rem_vol, vol_dim, vol_origin, cam_pose,
                other_params, color_im, depth_im, rem_im);
    }
  }
}
