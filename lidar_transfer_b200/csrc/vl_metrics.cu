// vl_metrics.cu -- identity re-render metrics on the device ("next" row N3), sm_100a.
//
// Replaces compare() (auxiliary/laserscan.py:1181-1301) + iouEval.addBatch (auxiliary/np_ioueval.py:32-45): the
// reference copies eight H x W images around on the host, masks them with numpy, renumbers the labels that occur to
// 0..k-1 (:1217-1223) and accumulates a confusion matrix with np.add.at.  Here the images stay where the ray cast
// left them:
//   k_cmp_mask   per pixel: the source's no-data (black) and background masks (:1200-1210), |colour difference|,
//                squared range / remission differences under the background mask (:1249-1276), presence bits of the
//                masked labels, block-reduced sum of the squared range differences (double)
//   k_cmp_rank   one CTA: prefix count over the 65 536 possible label values -> rank of each label that occurs (the
//                sequential renumbering of :1217-1223 maps the sorted union of the labels to 0..k-1)
//   k_cmp_conf   per pixel: conf[rank(target)][rank(source)] += 1 (rows = prediction, np_ioueval.py:41-45)
// IoU / accuracy are a few hundred flops on the n_classes^2 matrix and stay on the host (np_ioueval.py:47-70).
#include "vl_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kLabelSpace = 65536;   // labels are masked with 0xFFFF when they are read (laserscan.py:588)

struct CmpHeader {
  double range_sq_sum;   // sum over pixels of (source_range - target_range)^2 under the background mask
  int n_present;         // k: labels that occur
  int bad_label;         // a label outside [0, 65536) or more labels than classes: the reference raises IndexError
  int pad[60];
};
static_assert(sizeof(CmpHeader) == 256, "compare header is 256 B");

__global__ void k_cmp_init(CmpHeader* hdr, unsigned int* present, long long* conf, int n_conf) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kLabelSpace / 32; i += stride) present[i] = 0u;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_conf; i += stride) conf[i] = 0ll;
  if (blockIdx.x == 0 && threadIdx.x == 0) { hdr->range_sq_sum = 0.0; hdr->n_present = 0; hdr->bad_label = 0; }
}

__global__ void __launch_bounds__(kThreads)
k_cmp_mask(const float* __restrict__ source_color, const float* __restrict__ target_color,
           const int* __restrict__ source_label, const int* __restrict__ target_label,
           const float* __restrict__ source_range, const float* __restrict__ target_range,
           const float* __restrict__ source_rem, const float* __restrict__ target_rem, int n,
           float* __restrict__ label_diff, float* __restrict__ range_diff, float* __restrict__ rem_diff,
           int* __restrict__ masked_source, int* __restrict__ masked_target, unsigned int* __restrict__ present,
           CmpHeader* hdr) {
  __shared__ double s_sum[kThreads / 32];
  const int i = blockIdx.x * kThreads + threadIdx.x;
  double sq = 0.0;
  if (i < n) {
    const float sr_ = source_color[3 * (size_t)i], sg = source_color[3 * (size_t)i + 1], sb = source_color[3 * (size_t)i + 2];
    const bool black = __fadd_rn(__fadd_rn(sr_, sg), sb) == 0.f;      // :1200 np.sum(source_color, axis=2) == 0
    int sl = black ? 0 : source_label[i];
    int tl = black ? 0 : target_label[i];
    const bool bg = sl == 0;                                           // :1206
    if (bg) tl = 0;
    const bool zero_t = black || bg;
    label_diff[3 * (size_t)i] = fabsf(__fsub_rn(sr_, zero_t ? 0.f : target_color[3 * (size_t)i]));   // :1211
    label_diff[3 * (size_t)i + 1] = fabsf(__fsub_rn(sg, zero_t ? 0.f : target_color[3 * (size_t)i + 1]));
    label_diff[3 * (size_t)i + 2] = fabsf(__fsub_rn(sb, zero_t ? 0.f : target_color[3 * (size_t)i + 2]));
    masked_source[i] = sl;
    masked_target[i] = tl;
    if ((unsigned)sl >= (unsigned)kLabelSpace || (unsigned)tl >= (unsigned)kLabelSpace) hdr->bad_label = 1;
    else { atomicOr(&present[sl >> 5], 1u << (sl & 31)); atomicOr(&present[tl >> 5], 1u << (tl & 31)); }
    const float dr = __fsub_rn(bg ? 0.f : source_range[i], bg ? 0.f : target_range[i]);   // :1249-1252
    const float rd = __fmul_rn(dr, dr);
    range_diff[i] = rd;
    sq = (double)rd;
    const float dm = __fsub_rn(bg ? 0.f : source_rem[i], bg ? 0.f : target_rem[i]);       // :1270-1276
    rem_diff[i] = __fmul_rn(dm, dm);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) t += s_sum[k];
    if (t != 0.0) atomicAdd(&hdr->range_sq_sum, t);
  }
}

// rank[v] = number of present labels < v, for the words' first bits; one CTA of 1024 threads, 2 words each
__global__ void __launch_bounds__(1024)
k_cmp_rank(const unsigned int* __restrict__ present, int* __restrict__ word_rank, CmpHeader* hdr, int n_classes) {
  __shared__ int s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned int a = present[2 * tid], b = present[2 * tid + 1];
  const int v = __popc(a) + __popc(b);
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    const int x = s_warp[lane];
    int xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= d) xi += u; }
    s_warp[lane] = xi - x;
    if (lane == 31) { hdr->n_present = xi; if (xi > n_classes) hdr->bad_label = 1; }
  }
  __syncthreads();
  const int excl = s_warp[w] + incl - v;
  word_rank[2 * tid] = excl;
  word_rank[2 * tid + 1] = excl + __popc(a);
}

__global__ void __launch_bounds__(kThreads)
k_cmp_conf(const int* __restrict__ masked_source, const int* __restrict__ masked_target, int n,
           const unsigned int* __restrict__ present, const int* __restrict__ word_rank, int n_classes,
           unsigned long long* __restrict__ conf, const CmpHeader* __restrict__ hdr) {
  if (hdr->bad_label) return;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int sl = masked_source[i], tl = masked_target[i];
  const int rs = word_rank[sl >> 5] + __popc(present[sl >> 5] & ((1u << (sl & 31)) - 1u));
  const int rt = word_rank[tl >> 5] + __popc(present[tl >> 5] & ((1u << (tl & 31)) - 1u));
  atomicAdd(&conf[(size_t)rt * n_classes + rs], 1ull);   // rows = prediction (the re-rendered scan), cols = source
}

struct CmpLayout { size_t off_present, off_word_rank, off_ms, off_mt, total; };
CmpLayout cmp_layout(int n) {
  CmpLayout L;
  const size_t nn = n > 0 ? (size_t)n : 1;
  size_t off = 256;
  L.off_present = off;   off = vl_align256(off + kLabelSpace / 8);
  L.off_word_rank = off; off = vl_align256(off + 4 * (kLabelSpace / 32));
  L.off_ms = off;        off = vl_align256(off + 4 * nn);
  L.off_mt = off;        off = vl_align256(off + 4 * nn);
  L.total = off;
  return L;
}

}  // namespace

extern "C" size_t vl_compare_workspace_bytes(int n_pixels) { return cmp_layout(n_pixels).total; }

extern "C" int vl_compare(const float* d_source_color, const float* d_target_color, const int* d_source_label,
                          const int* d_target_label, const float* d_source_range, const float* d_target_range,
                          const float* d_source_rem, const float* d_target_rem, int n_pixels, int n_classes,
                          float* d_label_diff, float* d_range_diff, float* d_rem_diff, long long* d_conf,
                          void* d_workspace, size_t workspace_bytes, vl_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_pixels < 0 || n_classes <= 0 || !d_conf || !d_workspace || (((uintptr_t)d_workspace) & 255) ||
      (n_pixels > 0 && (!d_source_color || !d_target_color || !d_source_label || !d_target_label || !d_source_range ||
                        !d_target_range || !d_source_rem || !d_target_rem || !d_label_diff || !d_range_diff || !d_rem_diff))) {
    vl_set_error("vl_compare: invalid argument (n_pixels %d, n_classes %d)", n_pixels, n_classes);
    return VL_EINVAL;
  }
  const CmpLayout L = cmp_layout(n_pixels);
  if (workspace_bytes < L.total) {
    vl_set_error("vl_compare: workspace too small (%zu < %zu bytes)", workspace_bytes, L.total);
    return VL_ENOSPACE;
  }
  char* Wk = static_cast<char*>(d_workspace);
  CmpHeader* hdr = reinterpret_cast<CmpHeader*>(Wk);
  unsigned int* present = reinterpret_cast<unsigned int*>(Wk + L.off_present);
  int* word_rank = reinterpret_cast<int*>(Wk + L.off_word_rank);
  int* ms = reinterpret_cast<int*>(Wk + L.off_ms);
  int* mt = reinterpret_cast<int*>(Wk + L.off_mt);
  VlProfScope ps(VL_ST_COMPARE, stream);
  k_cmp_init<<<32, 256, 0, stream>>>(hdr, present, d_conf, n_classes * n_classes);
  VL_LAUNCH_CHECK("k_cmp_init");
  const int nb = (n_pixels + kThreads - 1) / kThreads;
  if (n_pixels > 0) {
    k_cmp_mask<<<nb, kThreads, 0, stream>>>(d_source_color, d_target_color, d_source_label, d_target_label, d_source_range,
                                           d_target_range, d_source_rem, d_target_rem, n_pixels, d_label_diff, d_range_diff,
                                           d_rem_diff, ms, mt, present, hdr);
    VL_LAUNCH_CHECK("k_cmp_mask");
  }
  k_cmp_rank<<<1, 1024, 0, stream>>>(present, word_rank, hdr, n_classes);
  VL_LAUNCH_CHECK("k_cmp_rank");
  if (n_pixels > 0) {
    k_cmp_conf<<<nb, kThreads, 0, stream>>>(ms, mt, n_pixels, present, word_rank, n_classes,
                                           reinterpret_cast<unsigned long long*>(d_conf), hdr);
    VL_LAUNCH_CHECK("k_cmp_conf");
  }
  return VL_OK;
}

// Synchronises; info: [0] labels that occur (k), [1] bad-label flag; *range_sq_sum = sum of squared range differences.
extern "C" int vl_compare_status(const void* d_workspace, vl_stream stream_, int* info, double* range_sq_sum) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d_workspace) { vl_set_error("vl_compare_status: null workspace"); return VL_EINVAL; }
  CmpHeader h;
  VL_CUDA_CHECK(cudaMemcpyAsync(&h, d_workspace, sizeof(h), cudaMemcpyDeviceToHost, stream));
  VL_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (info) { info[0] = h.n_present; info[1] = h.bad_label; }
  if (range_sq_sum) *range_sq_sum = h.range_sq_sum;
  if (h.bad_label) {
    vl_set_error("vl_compare: a label outside [0, 65536) or more distinct labels (%d) than classes "
                 "(the reference raises IndexError in np.add.at)", h.n_present);
    return VL_EINVAL;
  }
  return VL_OK;
}
