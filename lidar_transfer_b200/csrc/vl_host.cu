// vl_host.cu -- the HOST-pointer side of libvlidar.so: the reference's own entry point `ctrace`
// (auxiliary/raytracer/RayTracer.cpp:116-124, bound by RayTracerCython.pyx:5-7) and the ray normalisation the
// reference's Ray constructor performs (Vector3.h:73-89), evaluated on the host so that the device sees its bits.
//
// ctrace per call: (1) the rays are a per-sensor constant -- the normalised directions and the beam index built from
// them are cached on the device, keyed on (n_rays, height, byte-wise equality with the previous call's rays);
// (2) the caller's pageable mesh arrays are copied into ONE packed pinned staging buffer by a small pool of threads,
// 1 MB at a time, each finished run of chunks leaving for the device at once (the driver's own pageable path is a single
// thread's memcpy); (3) cast (or LBVH build + trace); (4) ONE device->host copy of the packed results; (5) the
// results of the rays that hit are merged into the caller's buffers on the host -- misses leave them untouched
// (RayTracer.cpp:72-90), without the caller's buffers ever travelling to the device.
// Compiled with -ffp-contract=off: the normaliser's roundings are the reference's (canonical -ffp-contract=off build).
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "vl_common.cuh"
#if defined(__SSE__) || defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#include <xmmintrin.h>
#define VL_HAVE_SSE 1
#endif

static thread_local int g_ctrace_status = VL_OK;
extern "C" int vl_ctrace_status(void) { return g_ctrace_status; }

// ---------------------------------------------------------------------------
// normalize(), Vector3.h:73-89: a.w = 0; D = a*a; D = hadd(D, D); D = hadd(D, D)  ->  (x2 + y2) + (z2 + 0);
// r = rsqrtps(D); r = 1.5 r + ((D * -0.5) * r) * (r * r); a * r.  rsqrtss is the same estimate as a lane of rsqrtps.
// ---------------------------------------------------------------------------
extern "C" int vl_normalize_rays(const float* rays, int n_rays, float* out) {
  if (n_rays < 0 || (n_rays > 0 && (!rays || !out))) {
    vl_set_error("vl_normalize_rays: invalid argument (n_rays %d)", n_rays);
    return VL_EINVAL;
  }
#ifdef VL_HAVE_SSE
  for (size_t i = 0; i < (size_t)n_rays; ++i) {
    const float x = rays[3 * i], y = rays[3 * i + 1], z = rays[3 * i + 2];
    const float D = (x * x + y * y) + (z * z + 0.0f);
    float r = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(D)));
    r = (1.5f * r) + (((D * -0.5f) * r) * (r * r));
    out[3 * i] = x * r; out[3 * i + 1] = y * r; out[3 * i + 2] = z * r;
  }
  return VL_OK;
#else
  vl_set_error("vl_normalize_rays: the reference's normalisation is the x86 rsqrtps instruction; this host has none");
  return VL_EINVAL;
#endif
}

// ---------------------------------------------------------------------------
// a small pool of host threads: pageable -> pinned staging, and the merge of the results
// ---------------------------------------------------------------------------
namespace {

// memcpy into the staging buffer with non-temporal stores: the destination is only read by the DMA engine, so it
// should neither be fetched for ownership nor displace the source from the caches
inline void copy_stream(char* dst, const char* src, size_t bytes) {
#ifdef VL_HAVE_SSE
  static const bool nt = getenv("VLIDAR_NO_NT") == nullptr;
  if (nt && (((uintptr_t)dst) & 15) == 0 && bytes >= 4096) {
    const size_t n16 = bytes / 64;
    const __m128i* s = reinterpret_cast<const __m128i*>(src);
    __m128i* d = reinterpret_cast<__m128i*>(dst);
    for (size_t i = 0; i < n16; ++i) {
      const __m128i a = _mm_loadu_si128(s + 4 * i), b = _mm_loadu_si128(s + 4 * i + 1);
      const __m128i c = _mm_loadu_si128(s + 4 * i + 2), e = _mm_loadu_si128(s + 4 * i + 3);
      _mm_stream_si128(d + 4 * i, a); _mm_stream_si128(d + 4 * i + 1, b);
      _mm_stream_si128(d + 4 * i + 2, c); _mm_stream_si128(d + 4 * i + 3, e);
    }
    _mm_sfence();
    const size_t done = n16 * 64;
    if (done < bytes) memcpy(dst + done, src + done, bytes - done);
    return;
  }
#endif
  memcpy(dst, src, bytes);
}

// A handful of detached worker threads that run fn(0) .. fn(n-1) together with the calling thread.  `progress(k)` is
// called on the calling thread, in order, whenever items 0 .. k-1 are all done (at least `batch` new ones, or the tail).
class WorkPool {
 public:
  static WorkPool& get() {
    static WorkPool* p = new WorkPool();   // never destroyed: the workers are detached and outlive static destructors
    return *p;
  }
  // lend: the calling thread takes items too while it waits; without, it only watches for finished runs of items and
  // reports them at once (the staging copy: a finished run has to leave for the device NOW, not after the caller's own
  // next megabyte)
  template <class Fn, class Progress>
  void run(int n, int batch, Fn fn, Progress progress, bool lend = true) {
    if (n <= 0) return;
    if (n_workers_ == 0 || n == 1) {
      for (int i = 0; i < n; ++i) fn(i);
      progress(n);
      return;
    }
    if ((int)done_.size() < n) done_ = std::vector<std::atomic<int>>(n);
    for (int i = 0; i < n; ++i) done_[i].store(0, std::memory_order_relaxed);
    std::function<void(int)> f = fn;
    {
      std::lock_guard<std::mutex> lock(mu_);
      fn_ = &f; n_ = n; next_.store(0); active_ = n_workers_; ++generation_;
    }
    cv_.notify_all();
    int issued = 0;
    while (issued < n) {
      int k = issued;
      while (k < n && done_[k].load(std::memory_order_acquire)) ++k;
      if (k > issued && (k == n || k - issued >= batch)) {
        progress(k);
        issued = k;
      } else if (k < n) {   // lend a hand instead of spinning
        if (lend || n_workers_ < 4) {
          const int i = next_.fetch_add(1);
          if (i < n) { f(i); done_[i].store(1, std::memory_order_release); }
          else std::this_thread::yield();
        } else {
#ifdef VL_HAVE_SSE
          _mm_pause();
#endif
        }
      }
    }
    std::unique_lock<std::mutex> lock(mu_);
    idle_cv_.wait(lock, [&] { return active_ == 0; });
    fn_ = nullptr;
  }

 private:
  WorkPool() {
    // workers beside the calling thread.  Measured on the B200 box's host (16 vCPUs; profiles/r02_experiments.md sections 2
    // and 8, 1 M-triangle scan, packed wire formats): 3 / 5 / 8 / 12 workers -> 1.15 / 0.97 / 0.85 / 0.92 ms per call.
    // Default: the process's share of the host's threads (torchrun exports LOCAL_WORLD_SIZE: one process per GPU), at most 8
    // and at most half of the machine.
    const int hw = (int)std::thread::hardware_concurrency();
    int local_world = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) local_world = atoi(e) > 0 ? atoi(e) : 1;
    int want = hw > 0 ? hw / local_world - 1 : 3;
    if (want > 8) want = 8;
    if (hw > 0 && want > hw / 2) want = hw / 2;
    if (want < 1) want = 1;
    if (const char* e = getenv("VLIDAR_COPY_THREADS")) want = atoi(e);
    if (hw > 0 && want > hw - 1) want = hw - 1;
    if (want < 0) want = 0;
    n_workers_ = want;
    for (int t = 0; t < want; ++t) std::thread([this] { loop(); }).detach();
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lock(mu_);
      cv_.wait(lock, [&] { return generation_ != seen; });
      seen = generation_;
      const std::function<void(int)>* f = fn_;
      const int n = n_;
      lock.unlock();
      for (;;) {
        const int i = next_.fetch_add(1);
        if (i >= n) break;
        (*f)(i);
        done_[i].store(1, std::memory_order_release);
      }
      lock.lock();
      if (--active_ == 0) idle_cv_.notify_all();
    }
  }
  int n_workers_ = 0;
  std::mutex mu_;
  std::condition_variable cv_, idle_cv_;
  unsigned long long generation_ = 0;
  int active_ = 0, n_ = 0;
  const std::function<void(int)>* fn_ = nullptr;
  std::atomic<int> next_{0};
  std::vector<std::atomic<int>> done_;
};

constexpr size_t kChunk = 1 << 20;   // source bytes per staging item

// ---------------------------------------------------------------------------
// compact wire formats of the staging copy.  The calling thread and the pool read the caller's arrays once anyway; what
// they WRITE is what crosses PCIe, so the staging copy packs on the way:
//   faces:   three vertex indices < 2^21 in one 64-bit word (8 B instead of 12 B per face; meshes with more than 2 M
//            vertices travel raw), expanded again on the device by k_unpack_faces;
//   colours: three bytes per vertex (3 B instead of 12 B) when every component is in 0 .. 255 -- what get_mesh produces
//            (fusion_lidar.py:419-423, uint8) and throw_rays_at_mesh widens to int32 (:437); read as they are by the cast's
//            write-back (VL_COLORS_U8), widened by k_unpack_colors for the LBVH path.
// An index or a component that does not fit is noticed by the packing loop itself and the array travels raw instead.
// ---------------------------------------------------------------------------
constexpr int kIdxBits = 21;

__global__ void k_unpack_faces(const unsigned long long* __restrict__ packed, size_t n3, int* __restrict__ faces) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const size_t f = i / 3;
  const int k = (int)(i - 3 * f);
  faces[i] = (int)((__ldg(packed + f) >> (kIdxBits * k)) & ((1ull << kIdxBits) - 1ull));
}

__global__ void k_unpack_colors(const unsigned char* __restrict__ c8, size_t n3, int* __restrict__ colors) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n3) colors[i] = (int)__ldg(c8 + i);
}

// Both packing loops write the staging buffer with non-temporal stores, like copy_stream: the destination is read by the
// DMA engine only; lines left dirty in the caches of several cores make the host->device copy that follows markedly
// slower (measured: profiles/r02_experiments.md section 8).
// returns the OR of every index (bits above 2^21 or the sign bit set = does not fit); dst is 16-byte aligned
#if defined(VL_HAVE_SSE) && defined(__GNUC__)
#include <immintrin.h>
#define VL_HAVE_AVX2_DISPATCH 1
// Eight faces per iteration with AVX2 (selected at run time by __builtin_cpu_supports): three 256-bit loads hold
// a0 b0 c0 a1 b1 c1 ... c7; the stride-3 de-interleave is three lane permutes + two blends per component; the 21-bit
// fields are assembled in 64-bit lanes and leave as 32-byte non-temporal stores.  Returns the number of faces packed
// (a multiple of 8); dst must be 32-byte aligned.
__attribute__((target("avx2"))) static size_t pack_faces_avx2(unsigned long long* dst, const int* src, size_t n_faces,
                                                              unsigned int* seen_out) {
  const __m256i ia0 = _mm256_setr_epi32(0, 3, 6, 0, 0, 0, 0, 0), ia1 = _mm256_setr_epi32(0, 0, 0, 1, 4, 7, 0, 0), ia2 = _mm256_setr_epi32(0, 0, 0, 0, 0, 0, 2, 5);
  const __m256i ib0 = _mm256_setr_epi32(1, 4, 7, 0, 0, 0, 0, 0), ib1 = _mm256_setr_epi32(0, 0, 0, 2, 5, 0, 0, 0), ib2 = _mm256_setr_epi32(0, 0, 0, 0, 0, 0, 3, 6);
  const __m256i ic0 = _mm256_setr_epi32(2, 5, 0, 0, 0, 0, 0, 0), ic1 = _mm256_setr_epi32(0, 0, 0, 3, 6, 0, 0, 0), ic2 = _mm256_setr_epi32(0, 0, 0, 0, 0, 1, 4, 7);
  __m256i acc = _mm256_setzero_si256();
  size_t f = 0;
  for (; f + 8 <= n_faces; f += 8) {
    const __m256i v0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 3 * f));
    const __m256i v1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 3 * f + 8));
    const __m256i v2 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 3 * f + 16));
    acc = _mm256_or_si256(acc, _mm256_or_si256(v0, _mm256_or_si256(v1, v2)));
    // a_i = int 3i: v0[0,3,6] v1[1,4,7] v2[2,5];  b_i = int 3i+1: v0[1,4,7] v1[2,5] v2[0,3,6];  c_i: v0[2,5] v1[0,3,6] v2[1,4,7]
    const __m256i A = _mm256_blend_epi32(_mm256_blend_epi32(_mm256_permutevar8x32_epi32(v0, ia0), _mm256_permutevar8x32_epi32(v1, ia1), 0x38),
                                         _mm256_permutevar8x32_epi32(v2, ia2), 0xC0);
    const __m256i B = _mm256_blend_epi32(_mm256_blend_epi32(_mm256_permutevar8x32_epi32(v0, ib0), _mm256_permutevar8x32_epi32(v1, ib1), 0x18),
                                         _mm256_permutevar8x32_epi32(v2, ib2), 0xE0);
    const __m256i C = _mm256_blend_epi32(_mm256_blend_epi32(_mm256_permutevar8x32_epi32(v0, ic0), _mm256_permutevar8x32_epi32(v1, ic1), 0x1C),
                                         _mm256_permutevar8x32_epi32(v2, ic2), 0xE0);
    const __m256i w0 = _mm256_or_si256(_mm256_cvtepu32_epi64(_mm256_castsi256_si128(A)),
                                       _mm256_or_si256(_mm256_slli_epi64(_mm256_cvtepu32_epi64(_mm256_castsi256_si128(B)), kIdxBits),
                                                       _mm256_slli_epi64(_mm256_cvtepu32_epi64(_mm256_castsi256_si128(C)), 2 * kIdxBits)));
    const __m256i w1 = _mm256_or_si256(_mm256_cvtepu32_epi64(_mm256_extracti128_si256(A, 1)),
                                       _mm256_or_si256(_mm256_slli_epi64(_mm256_cvtepu32_epi64(_mm256_extracti128_si256(B, 1)), kIdxBits),
                                                       _mm256_slli_epi64(_mm256_cvtepu32_epi64(_mm256_extracti128_si256(C, 1)), 2 * kIdxBits)));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + f), w0);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + f + 4), w1);
  }
  _mm_sfence();
  alignas(32) unsigned int lanes[8];
  _mm256_store_si256(reinterpret_cast<__m256i*>(lanes), acc);
  *seen_out = lanes[0] | lanes[1] | lanes[2] | lanes[3] | lanes[4] | lanes[5] | lanes[6] | lanes[7];
  return f;
}
#endif

inline unsigned int pack_faces(unsigned long long* dst, const int* src, size_t n_faces) {
  unsigned int seen = 0;
  auto word = [&](size_t f) {
    const unsigned int a = (unsigned int)src[3 * f], b = (unsigned int)src[3 * f + 1], c = (unsigned int)src[3 * f + 2];
    seen |= a | b | c;
    return (unsigned long long)a | ((unsigned long long)b << kIdxBits) | ((unsigned long long)c << (2 * kIdxBits));
  };
  size_t f = 0;
#ifdef VL_HAVE_AVX2_DISPATCH
  static const bool avx2 = __builtin_cpu_supports("avx2") && getenv("VLIDAR_NO_AVX2") == nullptr;
  if (avx2 && (((uintptr_t)dst) & 31) == 0) {
    unsigned int s8 = 0;
    f = pack_faces_avx2(dst, src, n_faces, &s8);
    seen |= s8;
  }
#endif
#ifdef VL_HAVE_SSE
  if ((((uintptr_t)(dst + f)) & 15) == 0) {
    for (; f + 2 <= n_faces; f += 2) {
      const unsigned long long w0 = word(f), w1 = word(f + 1);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + f), _mm_set_epi64x((long long)w1, (long long)w0));
    }
    _mm_sfence();
  }
#endif
  for (; f < n_faces; ++f) dst[f] = word(f);
  return seen;
}

inline unsigned int pack_colors(unsigned char* dst, const int* src, size_t n) {
  unsigned int seen = 0;
  size_t i = 0;
#ifdef VL_HAVE_SSE
  __m128i acc = _mm_setzero_si128();
  const bool aligned = (((uintptr_t)dst) & 15) == 0;
  for (; i + 16 <= n; i += 16) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 4));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 8));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 12));
    acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(a, b), _mm_or_si128(c, d)));
    // values in 0 .. 255 pass both saturating packs unchanged; anything else is caught by `acc`
    const __m128i v = _mm_packus_epi16(_mm_packs_epi32(a, b), _mm_packs_epi32(c, d));
    if (aligned) _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), v); else _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), v);
  }
  if (aligned) _mm_sfence();
  alignas(16) unsigned int lanes[4];
  _mm_store_si128(reinterpret_cast<__m128i*>(lanes), acc);
  seen = lanes[0] | lanes[1] | lanes[2] | lanes[3];
#endif
  for (; i < n; ++i) { seen |= (unsigned int)src[i]; dst[i] = (unsigned char)src[i]; }
  return seen;
}

struct HostCtx {
  std::mutex mu;
  cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: the colours / remissions travel beside the cast's first kernels
  cudaEvent_t ev_inputs = nullptr;
  char* arena = nullptr;      size_t arena_bytes = 0;    // device: mesh inputs, workspace / blob, packed outputs
  char* pinned = nullptr;     size_t pinned_bytes = 0;   // host staging, same packing
  char* pinned_raw = nullptr; size_t pinned_raw_bytes = 0;   // host staging of an array that could not be packed (rare)
  char* pinned_dirs = nullptr; size_t pinned_dirs_bytes = 0; // host: the unit directions as the device has them
  // per-sensor cache: [beam index][directions f32 x 3 n_rays] in one device allocation + the rays they were made from
  char* cache = nullptr;      size_t cache_bytes = 0;
  std::vector<float> rays_host;
  int c_n_rays = -1, c_height = -1, c_norm = -1;
  bool cache_valid = false;
  long long cache_hits = 0, cache_misses = 0;
  // host wall time of the most recent call's phases (ms): beam index (re)build, staging + H2D issue (incl. the ray
  // comparison), cast + D2H (wait), merge
  double t_ms[4] = {0, 0, 0, 0};
  long long h2d_bytes = 0, d2h_bytes = 0;   // of the most recent call
};
inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
HostCtx g_ctx;
std::atomic<int> g_ctrace_method{0};      // 0 = beam index + scene-streaming cast, 1 = LBVH build + traversal
std::atomic<int> g_ctrace_normalize{0};   // 0 = vl_normalize_rays on the host (the reference's bits), 1 = IEEE on the device
std::atomic<int> g_ctrace_wire{1};        // 1 = packed wire formats + endpoints recomputed on the host, 0 = everything raw

int grow(char** p, size_t* have, size_t want, bool host) {
  if (want <= *have) return VL_OK;
  if (*p) { if (host) VL_CUDA_CHECK(cudaFreeHost(*p)); else VL_CUDA_CHECK(cudaFree(*p)); }
  *p = nullptr; *have = 0;
  const size_t n = want + want / 4;
  if (host) VL_CUDA_CHECK(cudaHostAlloc((void**)p, n, cudaHostAllocDefault)); else VL_CUDA_CHECK(cudaMalloc((void**)p, n));
  *have = n;
  return VL_OK;
}

// the per-sensor part: normalised directions + beam index on the device, (re)built when the rays change
int rebuild_beams(HostCtx& c, const float* rays, int n_rays, int height, int norm) {
  const size_t nr = (size_t)n_rays;
  ++c.cache_misses;
  c.cache_valid = false;
  const size_t beams_bytes = vl_align256(vl_beams_bytes_impl(n_rays, height));
  int rc = grow(&c.cache, &c.cache_bytes, beams_bytes + 12 * nr, false);
  if (rc) return rc;
  rc = grow(&c.pinned_dirs, &c.pinned_dirs_bytes, 12 * nr, true);
  if (rc) return rc;
  c.rays_host.assign(rays, rays + 3 * nr);
  float* stage = reinterpret_cast<float*>(c.pinned_dirs);
  if (norm == 0) {
    rc = vl_normalize_rays(rays, n_rays, stage);
    if (rc) return rc;
  } else {
    memcpy(stage, rays, 12 * nr);
  }
  float* d_dirs = reinterpret_cast<float*>(c.cache + beams_bytes);
  VL_CUDA_CHECK(cudaMemcpyAsync(d_dirs, stage, 12 * nr, cudaMemcpyHostToDevice, c.stream));
  rc = vl_beams_build_launch(d_dirs, n_rays, height, c.cache, norm == 0 ? VL_RAYS_NORMALIZED : 0, c.stream);
  if (rc) return rc;
  VL_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  c.c_n_rays = n_rays; c.c_height = height; c.c_norm = norm;
  c.cache_valid = true;
  return VL_OK;
}

enum ItemKind { IT_COPY = 0, IT_FACES, IT_COLORS, IT_RAYCMP };
struct Item { int kind; size_t off; const char* src; size_t count; size_t end; };   // end: staged bytes complete once this item is

int ctrace_locked(HostCtx& c, const float* rays, const float* origin, const float* verts, const int* faces,
                  const int* colors, const float* rem, int n_rays, int n_verts, int n_faces, int height,
                  float* endpoints, int* endcolors, float* range, float* endrem, int* tri_id) {
  if (!c.stream) VL_CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  if (!c.stream2) VL_CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
  if (!c.ev_inputs) VL_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_inputs, cudaEventDisableTiming));
  cudaStream_t s = c.stream, s2 = c.stream2;
  int norm = g_ctrace_normalize.load();
#ifndef VL_HAVE_SSE
  norm = 1;
#endif
  static const bool direct = getenv("VLIDAR_CTRACE_DIRECT") != nullptr;   // measurement aid: the driver's own pageable path
  const bool wire = g_ctrace_wire.load() != 0 && !direct;
  const size_t nr = (size_t)n_rays, nv = (size_t)n_verts, nf = (size_t)n_faces;
  const double t0 = now_ms();
  // (1) the rays are a per-sensor constant.  A cached beam index of the same shape is taken on trust here and the rays
  // are compared with the cached copy by the staging pool, beside the mesh copy; anything else is rebuilt first.
  const bool same_shape = c.cache_valid && c.c_n_rays == n_rays && c.c_height == height && c.c_norm == norm;
  int rc = VL_OK;
  if (!same_shape) {
    rc = rebuild_beams(c, rays, n_rays, height, norm);
    if (rc) return rc;
  }
  const double t1 = now_ms();
  const size_t beams_bytes = vl_align256(vl_beams_bytes_impl(n_rays, height));
  const float* d_dirs = reinterpret_cast<const float*>(c.cache + beams_bytes);
  const int ray_flags = norm == 0 ? VL_RAYS_NORMALIZED : 0;
  const bool want_lbvh = g_ctrace_method.load() == 1;
  const bool pack_f = wire && nv <= (1ull << kIdxBits);
  const bool pack_c = wire;
  const bool host_ep = wire && norm == 0;   // hit = o + d * t (BVH.cpp:106-107) from the unit directions the host made itself

  // device arena = [wire: origin | verts | faces (packed or raw) | colours (u8 or raw) | rem][unpacked: faces | colours]
  // [outputs: endpoints | endcolors | range | endrem | id | status][workspace or blob]; the pinned buffer mirrors the
  // wire section and the outputs
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = vl_align256(off + bytes); return o; };
  const size_t o_origin = take(12), o_verts = take(12 * nv), o_wf = take(pack_f ? 8 * nf : 12 * nf);
  const size_t split = off;   // what lies before is needed by the cast's first kernels, what follows only by the write-back
  const size_t o_wc = take(pack_c ? 3 * nv : 12 * nv), o_rem = take(4 * nv);
  const size_t in_bytes = off;
  const size_t o_faces = take(12 * nf), o_colors = take(12 * nv);
  const size_t o_ep = take(12 * nr), o_ec = take(12 * nr), o_range = take(4 * nr), o_erem = take(4 * nr), o_id = take(4 * nr), o_st = take(16);
  const size_t io_bytes = off;
  const size_t ws_cast = vl_cast_workspace_bytes_impl(n_rays, n_faces), ws_lbvh = vl_bvh_blob_bytes(n_faces);
  const size_t o_ws = take(want_lbvh ? ws_lbvh : ws_cast);
  rc = grow(&c.arena, &c.arena_bytes, off, false);
  if (rc) return rc;
  rc = grow(&c.pinned, &c.pinned_bytes, io_bytes, true);
  if (rc) return rc;
  char* A = c.arena;
  char* P = c.pinned;

  // (2) mesh: pageable -> (packed) pinned by the thread pool -> device, pipelined
  std::vector<Item> items;
  size_t end = 0;
  if (same_shape)
    for (size_t o = 0; o < 12 * nr; o += kChunk)
      items.push_back({IT_RAYCMP, o, reinterpret_cast<const char*>(rays) + o, 12 * nr - o < kChunk ? 12 * nr - o : kChunk, 0});
  auto add_copy = [&](size_t at, const void* src, size_t bytes) {
    for (size_t o = 0; o < bytes; o += kChunk) {
      const size_t n = bytes - o < kChunk ? bytes - o : kChunk;
      end = at + o + n;
      items.push_back({IT_COPY, at + o, static_cast<const char*>(src) + o, n, end});
    }
  };
  add_copy(o_origin, origin, 12);
  add_copy(o_verts, verts, 12 * nv);
  if (pack_f) {
    const size_t per = (kChunk / 12) & ~(size_t)3;   // a multiple of four: every item's first word is 32-byte aligned
    for (size_t f = 0; f < nf; f += per) {
      const size_t n = nf - f < per ? nf - f : per;
      end = o_wf + 8 * (f + n);
      items.push_back({IT_FACES, o_wf + 8 * f, reinterpret_cast<const char*>(faces + 3 * f), n, end});
    }
  } else {
    add_copy(o_wf, faces, 12 * nf);
  }
  items.back().end = split;   // the alignment gap travels with the last item of the front section
  if (pack_c) {
    const size_t per = kChunk / 4;
    for (size_t i = 0; i < 3 * nv; i += per) {
      const size_t n = 3 * nv - i < per ? 3 * nv - i : per;
      end = o_wc + i + n;
      items.push_back({IT_COLORS, o_wc + i, reinterpret_cast<const char*>(colors + i), n, end});
    }
  } else {
    add_copy(o_wc, colors, 12 * nv);
  }
  add_copy(o_rem, rem, 4 * nv);

  std::atomic<unsigned int> faces_seen{0}, colors_seen{0};
  std::atomic<int> rays_differ{0};
  size_t sent = 0;
  cudaError_t copy_err = cudaSuccess;
  bool front_tried = false, front_done = false;
  int front_rc = VL_OK;
  const int* d_faces = reinterpret_cast<const int*>(pack_f ? A + o_faces : A + o_wf);
  auto launch_front = [&](bool unpack) -> int {
    if (unpack && nf > 0) {
      k_unpack_faces<<<(unsigned)((3 * nf + 255) / 256), 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(A + o_wf), 3 * nf,
                                                                     reinterpret_cast<int*>(A + o_faces));
      VL_LAUNCH_CHECK("k_unpack_faces");
    }
    if (!want_lbvh) {
      int r = vl_cast_launch(c.cache, (const float*)(A + o_verts), d_faces, nullptr, nullptr, n_verts, n_faces,
                             (const float*)(A + o_origin), n_rays, height, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                             A + o_ws, s, VL_CAST_PHASE_FRONT);
      if (r) return r;
    }
    front_done = true;
    return VL_OK;
  };
  auto send_to = [&](size_t upto) {
    while (sent < upto && copy_err == cudaSuccess) {
      const size_t stop = (sent < split && upto > split) ? split : upto;
      copy_err = cudaMemcpyAsync(A + sent, P + sent, stop - sent, cudaMemcpyHostToDevice, sent < split ? s : s2);
      sent = stop;
    }
    // verts and faces are on their way: the cast's first kernels go in behind them, the colours travel beside
    if (sent >= split && !front_tried && copy_err == cudaSuccess) {
      front_tried = true;   // every ray-comparison and face item has completed by now (items complete in order)
      if (!rays_differ.load() && (!pack_f || (faces_seen.load() >> kIdxBits) == 0)) front_rc = launch_front(pack_f);
    }
  };
  if (direct) {
    for (const Item& it : items)
      if (it.kind == IT_COPY && copy_err == cudaSuccess)
        copy_err = cudaMemcpyAsync(A + it.off, it.src, it.count, cudaMemcpyHostToDevice, s);
      else if (it.kind == IT_RAYCMP && memcmp(reinterpret_cast<const char*>(c.rays_host.data()) + it.off, it.src, it.count) != 0)
        rays_differ.store(1);
    sent = in_bytes;
  } else {
    // >= 2 MB of source per DMA, issued by the calling thread the moment a run of items is complete: with four or more
    // workers it copies nothing itself (measured, 8 workers: 0.96 -> 0.86 ms per call against lending a hand with runs of 3)
    static const int batch = getenv("VLIDAR_STAGE_BATCH") ? atoi(getenv("VLIDAR_STAGE_BATCH")) : 2;
    static const bool lend = getenv("VLIDAR_STAGE_LEND") ? atoi(getenv("VLIDAR_STAGE_LEND")) != 0 : false;
    WorkPool::get().run((int)items.size(), batch,
        [&](int i) {
          const Item& it = items[i];
          switch (it.kind) {
            case IT_COPY: copy_stream(P + it.off, it.src, it.count); break;
            case IT_FACES: faces_seen.fetch_or(pack_faces(reinterpret_cast<unsigned long long*>(P + it.off), reinterpret_cast<const int*>(it.src), it.count)); break;
            case IT_COLORS: colors_seen.fetch_or(pack_colors(reinterpret_cast<unsigned char*>(P + it.off), reinterpret_cast<const int*>(it.src), it.count)); break;
            default: if (memcmp(reinterpret_cast<const char*>(c.rays_host.data()) + it.off, it.src, it.count) != 0) rays_differ.store(1);
          }
        },
        [&](int k) { send_to(k == (int)items.size() ? in_bytes : items[k - 1].end); }, lend);
  }
  VL_CUDA_CHECK(copy_err);
  if (front_rc) return front_rc;
  c.h2d_bytes = (long long)in_bytes;
  const double t2 = now_ms();

  // the unusual cases, in the order the device needs them settled
  if (same_shape) {
    if (rays_differ.load()) {   // another sensor: nothing has been launched against the stale index
      rc = rebuild_beams(c, rays, n_rays, height, norm);
      if (rc) return rc;
    } else {
      ++c.cache_hits;
    }
  }
  auto stage_raw = [&](size_t at, const void* src, size_t bytes, cudaStream_t st) -> int {   // an array that does not fit its wire format
    VL_CUDA_CHECK(cudaStreamSynchronize(st));   // pinned_raw may still be in flight from the other array
    int r = grow(&c.pinned_raw, &c.pinned_raw_bytes, bytes, true);
    if (r) return r;
    memcpy(c.pinned_raw, src, bytes);
    VL_CUDA_CHECK(cudaMemcpyAsync(A + at, c.pinned_raw, bytes, cudaMemcpyHostToDevice, st));
    VL_CUDA_CHECK(cudaStreamSynchronize(st));
    c.h2d_bytes += (long long)bytes;
    return VL_OK;
  };
  const bool faces_fit = !pack_f || (faces_seen.load() >> kIdxBits) == 0;
  if (!faces_fit) {   // an index outside [0, 2^21): a bad face (counted by the cast) -- it has to arrive as it is
    rc = stage_raw(o_faces, faces, 12 * nf, s);
    if (rc) return rc;
  }
  if (!front_done) {
    rc = launch_front(pack_f && faces_fit);
    if (rc) return rc;
  }
  const bool colors_fit = pack_c && (colors_seen.load() >> 8) == 0;
  if (pack_c && !colors_fit) {
    rc = stage_raw(o_colors, colors, 12 * nv, s2);
    if (rc) return rc;
  }
  const int* d_colors = reinterpret_cast<const int*>(pack_c ? (colors_fit ? A + o_wc : A + o_colors) : A + o_wc);
  VL_CUDA_CHECK(cudaEventRecord(c.ev_inputs, s2));
  VL_CUDA_CHECK(cudaStreamWaitEvent(s, c.ev_inputs, 0));

  // (3) the rest of the cast; a mesh that needs more work units than the workspace holds is answered through the tree
  const size_t o_back = host_ep ? o_ec : o_ep;   // first output section that travels back
  int st[4] = {0, 0, 0, 0};
  for (int attempt = want_lbvh ? 1 : 0; attempt < 2; ++attempt) {
    const bool lbvh = attempt == 1;
    if (lbvh) {
      if (colors_fit && nv > 0) {
        k_unpack_colors<<<(unsigned)((3 * nv + 255) / 256), 256, 0, s>>>(reinterpret_cast<const unsigned char*>(A + o_wc), 3 * nv,
                                                                        reinterpret_cast<int*>(A + o_colors));
        VL_LAUNCH_CHECK("k_unpack_colors");
      }
      const int* d_c32 = reinterpret_cast<const int*>(pack_c ? A + o_colors : A + o_wc);
      rc = vl_bvh_build_launch((const float*)(A + o_verts), d_faces, d_c32, (const float*)(A + o_rem), n_verts, n_faces, A + o_ws, s);
      if (rc) return rc;
      rc = vl_trace_launch(A + o_ws, n_faces, d_dirs, (const float*)(A + o_origin), n_rays, height, (float*)(A + o_ep),
                           (int*)(A + o_ec), (float*)(A + o_range), (float*)(A + o_erem), (int*)(A + o_id), ray_flags, s);
      if (rc) return rc;
      VL_CUDA_CHECK(cudaMemcpyAsync(A + o_st, A + o_ws + offsetof(VlHeader, n_bad_faces), sizeof(int), cudaMemcpyDeviceToDevice, s));
      VL_CUDA_CHECK(cudaMemsetAsync(A + o_st + 4, 0, 4, s));
    } else {
      rc = vl_cast_launch(c.cache, (const float*)(A + o_verts), d_faces, d_colors, (const float*)(A + o_rem), n_verts, n_faces,
                          (const float*)(A + o_origin), n_rays, height, (float*)(A + o_ep), (int*)(A + o_ec), (float*)(A + o_range),
                          (float*)(A + o_erem), (int*)(A + o_id), colors_fit ? VL_COLORS_U8 : 0, A + o_ws, s, VL_CAST_PHASE_RESOLVE);
      if (rc) return rc;
      VL_CUDA_CHECK(cudaMemcpyAsync(A + o_st, A + o_ws + 16, 16, cudaMemcpyDeviceToDevice, s));   // {n_bad_faces, overflow, ..} as k_cast_resolve left them
    }
    // (4) one copy back
    VL_CUDA_CHECK(cudaMemcpyAsync(P + o_back, A + o_back, io_bytes - o_back, cudaMemcpyDeviceToHost, s));
    VL_CUDA_CHECK(cudaStreamSynchronize(s));
    memcpy(st, P + o_st, 16);
    if (!lbvh && st[1]) {   // overflow: nothing was written; make room for the blob and take the tree
      size_t need = o_ws + ws_lbvh;
      if (need > c.arena_bytes) {
        // growing frees the arena: everything the tree needs is put back in place afterwards
        rc = grow(&c.arena, &c.arena_bytes, need, false);
        if (rc) return rc;
        A = c.arena;
        VL_CUDA_CHECK(cudaMemcpyAsync(A, P, in_bytes, cudaMemcpyHostToDevice, s));
        if (!faces_fit) { rc = stage_raw(o_faces, faces, 12 * nf, s); if (rc) return rc; }
        else if (pack_f && nf > 0) {
          k_unpack_faces<<<(unsigned)((3 * nf + 255) / 256), 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(A + o_wf), 3 * nf,
                                                                         reinterpret_cast<int*>(A + o_faces));
          VL_LAUNCH_CHECK("k_unpack_faces");
        }
        if (pack_c && !colors_fit) { rc = stage_raw(o_colors, colors, 12 * nv, s); if (rc) return rc; }
        d_faces = reinterpret_cast<const int*>(pack_f ? A + o_faces : A + o_wf);
      }
      continue;
    }
    break;
  }
  c.d2h_bytes = (long long)(io_bytes - o_back);
  const double t3 = now_ms();
  // (5) merge the hits into the caller's buffers (misses leave them untouched, RayTracer.cpp:72-90)
  const float* h_ep = reinterpret_cast<const float*>(P + o_ep);
  const int* h_ec = reinterpret_cast<const int*>(P + o_ec);
  const float* h_range = reinterpret_cast<const float*>(P + o_range);
  const float* h_erem = reinterpret_cast<const float*>(P + o_erem);
  const int* h_id = reinterpret_cast<const int*>(P + o_id);
  const float* h_dir = reinterpret_cast<const float*>(c.pinned_dirs);
  const float ox = origin[0], oy = origin[1], oz = origin[2];
  constexpr size_t kMergeBlock = 16384;
  WorkPool::get().run((int)((nr + kMergeBlock - 1) / kMergeBlock), 1 << 30, [&](int b) {
    const size_t r0 = (size_t)b * kMergeBlock, r1 = r0 + kMergeBlock < nr ? r0 + kMergeBlock : nr;
    size_t r = r0;
    while (r < r1) {   // runs of hits leave as block copies
      while (r < r1 && h_id[r] < 0) ++r;
      size_t e = r;
      while (e < r1 && h_id[e] >= 0) ++e;
      if (e > r) {
        if (host_ep) {   // BVH.cpp:106-107 hit = o + d * t, each operation rounded on its own (this file: -ffp-contract=off)
          for (size_t k = r; k < e; ++k) {
            const float t = h_range[k];
            endpoints[3 * k] = ox + h_dir[3 * k] * t;
            endpoints[3 * k + 1] = oy + h_dir[3 * k + 1] * t;
            endpoints[3 * k + 2] = oz + h_dir[3 * k + 2] * t;
          }
        } else {
          memcpy(endpoints + 3 * r, h_ep + 3 * r, 12 * (e - r));
        }
        memcpy(endcolors + 3 * r, h_ec + 3 * r, 12 * (e - r));
        memcpy(range + r, h_range + r, 4 * (e - r));
        memcpy(endrem + r, h_erem + r, 4 * (e - r));
      }
      r = e;
    }
    if (tri_id) memcpy(tri_id + r0, h_id + r0, 4 * (r1 - r0));
  }, [](int) {});
  c.t_ms[0] = t1 - t0; c.t_ms[1] = t2 - t1; c.t_ms[2] = t3 - t2; c.t_ms[3] = now_ms() - t3;
  if (st[0] > 0) {
    vl_set_error("ctrace: %d face(s) reference a vertex outside [0, %d); they were skipped", st[0], n_verts);
    return VL_EBADMESH;
  }
  return VL_OK;
}

}  // namespace

// test hook (host code only, no device needed): the staging copy's packing loops on caller-provided buffers
extern "C" int vl_debug_pack(const int* faces, long long n_faces, unsigned long long* packed_faces, const int* colors, long long n_components,
                             unsigned char* packed_colors, unsigned int* seen2) {
  if (n_faces < 0 || n_components < 0 || !seen2) { vl_set_error("vl_debug_pack: invalid argument"); return VL_EINVAL; }
  seen2[0] = n_faces > 0 ? pack_faces(packed_faces, faces, (size_t)n_faces) : 0u;
  seen2[1] = n_components > 0 ? pack_colors(packed_colors, colors, (size_t)n_components) : 0u;
  return VL_OK;
}

extern "C" void vl_ctrace_method(int method) { g_ctrace_method.store(method == 1 ? 1 : 0); }
extern "C" void vl_ctrace_normalize(int mode) { g_ctrace_normalize.store(mode == 1 ? 1 : 0); }
extern "C" void vl_ctrace_wire(int packed) { g_ctrace_wire.store(packed ? 1 : 0); }
extern "C" void vl_ctrace_traffic(long long* h2d_bytes, long long* d2h_bytes) {
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (h2d_bytes) *h2d_bytes = g_ctx.h2d_bytes;
  if (d2h_bytes) *d2h_bytes = g_ctx.d2h_bytes;
}
extern "C" void vl_ctrace_cache_stats(long long* hits, long long* misses) {
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  if (hits) *hits = g_ctx.cache_hits;
  if (misses) *misses = g_ctx.cache_misses;
}

extern "C" void vl_ctrace_timing(double* ms4) {
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  for (int k = 0; k < 4; ++k) ms4[k] = g_ctx.t_ms[k];
}

extern "C" int vl_ctrace_ids(const float* rays, const float* origin, const float* verts, const int* faces,
                             const int* colors, const float* rem, int n_rays, int n_verts, int n_faces, int height,
                             float* endpoints, int* endcolors, float* range, float* endrem, int* tri_id) {
  if (n_rays < 0 || n_verts < 0 || n_faces < 0 || height <= 0 || !origin ||
      (n_rays > 0 && (!rays || !endpoints || !endcolors || !range || !endrem)) ||
      (n_faces > 0 && (!verts || !faces || !colors || !rem)) || n_faces >= (1 << 28)) {
    vl_set_error("ctrace: invalid argument (n_rays %d, n_verts %d, n_faces %d, height %d)", n_rays, n_verts, n_faces, height);
    return VL_EINVAL;
  }
  int n_dev = 0;
  VL_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
  if (n_dev <= 0) { vl_set_error("ctrace: no CUDA device (libvlidar has no CPU fallback)"); return VL_ECUDA; }
  if (n_rays == 0) return VL_OK;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  const int rc = ctrace_locked(g_ctx, rays, origin, verts, faces, colors, rem, n_rays, n_verts, n_faces, height, endpoints,
                               endcolors, range, endrem, tri_id);
  if (rc != VL_OK && rc != VL_EBADMESH) {
    g_ctx.cache_valid = false;
    // a call that failed half way may have copies in flight out of the staging buffers the next call will overwrite
    if (g_ctx.stream) cudaStreamSynchronize(g_ctx.stream);
    if (g_ctx.stream2) cudaStreamSynchronize(g_ctx.stream2);
    cudaGetLastError();
  }
  return rc;
}

extern "C" void ctrace(float* rays, float* origin, float* verts, int* faces, int* colors, float* rem, int n_rays,
                       int n_verts, int n_faces, int height, float* endpoints, int* endcolors, float* range,
                       float* endrem) {
  g_ctrace_status = vl_ctrace_ids(rays, origin, verts, faces, colors, rem, n_rays, n_verts, n_faces, height, endpoints,
                                  endcolors, range, endrem, nullptr);
  if (g_ctrace_status != VL_OK) fprintf(stderr, "libvlidar ctrace error %d: %s\n", g_ctrace_status, vl_last_error());
}
