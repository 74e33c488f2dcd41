// ref_tsdf_cuda_harness.cu -- TEST INFRASTRUCTURE ONLY.
//
// The reference's own CUDA `integrate` kernel (the string pycuda's SourceModule JIT-compiles,
// auxiliary/fusion_lidar.py:66-229, extracted verbatim at build time into oracle/_ref/integrate_kernel.inc)
// compiled by nvcc for sm_100a with nvcc's defaults -- what pycuda does (SourceModule: `nvcc --cubin -arch sm_XX`,
// no other flags: -O3, -fmad=true, CUDA's atan2 / asinf / norm3df) -- and launched with the reference's geometry
// (fusion_lidar.py:233-250, 267-287: 1024-thread blocks, a 3-D grid of floor(cbrt) x floor(sqrt) x ceil blocks, a host
// loop of launches with the loop index in other_params[0]).  This is the reference's kernel itself running on the
// B200: the -m gpu tests hold the product's TSDF integration to it BIT FOR BIT at full size.
//
// Volumes: DEVICE pointers to float32[n_voxels + 1] (the kernel's guard is `voxel_idx > n`, fusion_lidar.py:92: the
// thread with voxel_idx == n reads and may write one element past the end -- the caller provides it).
// Small arrays and images: HOST pointers, staged to the device per call like pycuda's cuda.InOut arguments.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include "integrate_kernel.inc"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "ref_tsdf_cuda: %s: %s\n", #x, cudaGetErrorString(e_)); return -1; } } while (0)

extern "C" int ref_tsdf_integrate_cuda(float* d_tsdf, float* d_weight, float* d_color, float* d_rem, const float* h_vol_dim,
                                       const float* h_vol_origin, const float* h_cam_pose16, const float* h_other8,
                                       const float* h_color_im, const float* h_depth_im, const float* h_rem_im,
                                       int im_h, int im_w) {
  const double n_vox = (double)(int)h_vol_dim[0] * (int)h_vol_dim[1] * (int)h_vol_dim[2];
  cudaDeviceProp prop;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaGetDeviceProperties(&prop, dev));
  const int tpb = prop.maxThreadsPerBlock;                                  // gpu_dev.MAX_THREADS_PER_BLOCK
  const long long n_blocks = (long long)ceil(n_vox / (double)tpb);
  long long gx = (long long)floor(cbrt((double)n_blocks));
  if (gx > prop.maxGridSize[0]) gx = prop.maxGridSize[0];
  long long gy = (long long)floor(sqrt((double)n_blocks / (double)gx));
  if (gy > prop.maxGridSize[1]) gy = prop.maxGridSize[1];
  long long gz = (long long)ceil((double)n_blocks / (double)(gx * gy));
  if (gz > prop.maxGridSize[2]) gz = prop.maxGridSize[2];
  const int n_loops = (int)ceil(n_vox / ((double)(gx * gy * gz) * (double)tpb));
  float *d_dim, *d_origin, *d_pose, *d_other, *d_cim, *d_dim_im, *d_rim;
  const size_t npix = (size_t)im_h * im_w;
  CK(cudaMalloc(&d_dim, 3 * 4)); CK(cudaMalloc(&d_origin, 3 * 4)); CK(cudaMalloc(&d_pose, 16 * 4)); CK(cudaMalloc(&d_other, 8 * 4));
  CK(cudaMalloc(&d_cim, npix * 4)); CK(cudaMalloc(&d_dim_im, npix * 4)); CK(cudaMalloc(&d_rim, npix * 4));
  CK(cudaMemcpy(d_dim, h_vol_dim, 12, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_origin, h_vol_origin, 12, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_pose, h_cam_pose16, 64, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_cim, h_color_im, npix * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_dim_im, h_depth_im, npix * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_rim, h_rem_im, npix * 4, cudaMemcpyHostToDevice));
  float other[8];
  for (int k = 0; k < 8; ++k) other[k] = h_other8[k];
  for (int loop = 0; loop < n_loops; ++loop) {
    other[0] = (float)loop;
    CK(cudaMemcpy(d_other, other, 32, cudaMemcpyHostToDevice));
    integrate<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), dim3(tpb, 1, 1)>>>(d_tsdf, d_weight, d_color, d_rem, d_dim, d_origin,
                                                                                   d_pose, d_other, d_cim, d_dim_im, d_rim);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  cudaFree(d_dim); cudaFree(d_origin); cudaFree(d_pose); cudaFree(d_other); cudaFree(d_cim); cudaFree(d_dim_im); cudaFree(d_rim);
  return n_loops;
}
