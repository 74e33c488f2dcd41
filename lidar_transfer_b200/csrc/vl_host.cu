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

struct Chunk { size_t off; const char* src; size_t bytes; };

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
  int workers() const { return n_workers_; }
  template <class Fn, class Progress>
  void run(int n, int batch, Fn fn, Progress progress) {
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
        const int i = next_.fetch_add(1);
        if (i < n) { f(i); done_[i].store(1, std::memory_order_release); }
        else std::this_thread::yield();
      }
    }
    std::unique_lock<std::mutex> lock(mu_);
    idle_cv_.wait(lock, [&] { return active_ == 0; });
    fn_ = nullptr;
  }

 private:
  WorkPool() {
    // workers beside the calling thread.  Measured on the B200 box's host (profiles/r02_experiments.md, 1 M-triangle scan):
    // 0 / 1 / 2 / 3 / 5 workers -> 3.3 / 2.0 / 1.8 / 1.21 / 1.15 ms per call (the last two with non-temporal stores).
    // Default: the process's share of the host's threads (torchrun exports LOCAL_WORLD_SIZE: one process per GPU), at most 5.
    const int hw = (int)std::thread::hardware_concurrency();
    int local_world = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) local_world = atoi(e) > 0 ? atoi(e) : 1;
    int want = hw > 0 ? hw / local_world - 1 : 3;
    if (want > 5) want = 5;
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

constexpr size_t kChunk = 1 << 20;

struct HostCtx {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  char* arena = nullptr;      size_t arena_bytes = 0;    // device: mesh inputs, workspace / blob, packed outputs
  char* pinned = nullptr;     size_t pinned_bytes = 0;   // host staging, same packing
  // per-sensor cache: [beam index][directions f32 x 3 n_rays] in one device allocation + the rays they were made from
  char* cache = nullptr;      size_t cache_bytes = 0;
  std::vector<float> rays_host;
  int c_n_rays = -1, c_height = -1, c_norm = -1;
  bool cache_valid = false;
  long long cache_hits = 0, cache_misses = 0;
  // host wall time of the most recent call's phases (ms): rays / beam cache, staging + H2D issue, cast + D2H (wait), merge
  double t_ms[4] = {0, 0, 0, 0};
};
inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
HostCtx g_ctx;
std::atomic<int> g_ctrace_method{0};      // 0 = beam index + scene-streaming cast, 1 = LBVH build + traversal
std::atomic<int> g_ctrace_normalize{0};   // 0 = vl_normalize_rays on the host (the reference's bits), 1 = IEEE on the device

int grow(char** p, size_t* have, size_t want, bool host) {
  if (want <= *have) return VL_OK;
  if (*p) { if (host) VL_CUDA_CHECK(cudaFreeHost(*p)); else VL_CUDA_CHECK(cudaFree(*p)); }
  *p = nullptr; *have = 0;
  const size_t n = want + want / 4;
  if (host) VL_CUDA_CHECK(cudaHostAlloc((void**)p, n, cudaHostAllocDefault)); else VL_CUDA_CHECK(cudaMalloc((void**)p, n));
  *have = n;
  return VL_OK;
}

void add_chunks(std::vector<Chunk>& v, size_t off, const void* src, size_t bytes) {
  const char* s = static_cast<const char*>(src);
  for (size_t o = 0; o < bytes; o += kChunk) v.push_back({off + o, s + o, bytes - o < kChunk ? bytes - o : kChunk});
}

// the per-sensor part: normalised directions + beam index on the device, rebuilt only when the rays change
int ensure_beams(HostCtx& c, const float* rays, int n_rays, int height, int norm) {
  const size_t nr = (size_t)n_rays;
  if (c.cache_valid && c.c_n_rays == n_rays && c.c_height == height && c.c_norm == norm &&
      memcmp(c.rays_host.data(), rays, 12 * nr) == 0) {
    ++c.cache_hits;
    return VL_OK;
  }
  ++c.cache_misses;
  c.cache_valid = false;
  const size_t beams_bytes = vl_align256(vl_beams_bytes_impl(n_rays, height));
  int rc = grow(&c.cache, &c.cache_bytes, beams_bytes + 12 * nr, false);
  if (rc) return rc;
  rc = grow(&c.pinned, &c.pinned_bytes, 12 * nr, true);
  if (rc) return rc;
  c.rays_host.assign(rays, rays + 3 * nr);
  float* stage = reinterpret_cast<float*>(c.pinned);
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
  VL_CUDA_CHECK(cudaStreamSynchronize(c.stream));   // the staging buffer is reused below
  c.c_n_rays = n_rays; c.c_height = height; c.c_norm = norm;
  c.cache_valid = true;
  return VL_OK;
}

int ctrace_locked(HostCtx& c, const float* rays, const float* origin, const float* verts, const int* faces,
                  const int* colors, const float* rem, int n_rays, int n_verts, int n_faces, int height,
                  float* endpoints, int* endcolors, float* range, float* endrem, int* tri_id) {
  if (!c.stream) VL_CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  cudaStream_t s = c.stream;
  int norm = g_ctrace_normalize.load();
#ifndef VL_HAVE_SSE
  norm = 1;
#endif
  const double t0 = now_ms();
  int rc = ensure_beams(c, rays, n_rays, height, norm);
  if (rc) return rc;
  const double t1 = now_ms();
  const size_t nr = (size_t)n_rays, nv = (size_t)n_verts, nf = (size_t)n_faces;
  const size_t beams_bytes = vl_align256(vl_beams_bytes_impl(n_rays, height));
  const float* d_dirs = reinterpret_cast<const float*>(c.cache + beams_bytes);
  const int ray_flags = norm == 0 ? VL_RAYS_NORMALIZED : 0;
  const bool want_lbvh = g_ctrace_method.load() == 1;

  // device arena = [inputs: origin | verts | faces | colors | rem][outputs: endpoints | endcolors | range | endrem | id |
  // status][workspace or blob]; the pinned buffer mirrors the first two sections
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = vl_align256(off + bytes); return o; };
  const size_t o_origin = take(12), o_verts = take(12 * nv), o_faces = take(12 * nf), o_colors = take(12 * nv), o_rem = take(4 * nv);
  const size_t in_bytes = off;
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

  // (2) mesh: pageable -> pinned (thread pool) -> device, pipelined
  std::vector<Chunk> chunks;
  add_chunks(chunks, o_origin, origin, 12);
  add_chunks(chunks, o_verts, verts, 12 * nv);
  add_chunks(chunks, o_faces, faces, 12 * nf);
  add_chunks(chunks, o_colors, colors, 12 * nv);
  add_chunks(chunks, o_rem, rem, 4 * nv);
  size_t sent = 0;
  cudaError_t copy_err = cudaSuccess;
  static const bool direct = getenv("VLIDAR_CTRACE_DIRECT") != nullptr;   // measurement aid: the driver's own pageable path
  if (direct) {
    for (const Chunk& ch : chunks)
      if (copy_err == cudaSuccess) copy_err = cudaMemcpyAsync(A + ch.off, ch.src, ch.bytes, cudaMemcpyHostToDevice, s);
  } else {
    WorkPool::get().run((int)chunks.size(), 4,   // >= 4 MB per DMA, or the tail
        [&](int i) { copy_stream(P + chunks[i].off, chunks[i].src, chunks[i].bytes); },
        [&](int k) {
          const size_t end = chunks[k - 1].off + chunks[k - 1].bytes;
          if (copy_err == cudaSuccess) copy_err = cudaMemcpyAsync(A + sent, P + sent, end - sent, cudaMemcpyHostToDevice, s);
          sent = end;
        });
  }
  VL_CUDA_CHECK(copy_err);
  const double t2 = now_ms();

  // (3) cast; a mesh that needs more work units than the workspace holds is answered through the tree instead
  int st[4] = {0, 0, 0, 0};
  for (int attempt = want_lbvh ? 1 : 0; attempt < 2; ++attempt) {
    const bool lbvh = attempt == 1;
    if (lbvh) {
      rc = vl_bvh_build_launch((const float*)(A + o_verts), (const int*)(A + o_faces), (const int*)(A + o_colors),
                               (const float*)(A + o_rem), n_verts, n_faces, A + o_ws, s);
      if (rc) return rc;
      rc = vl_trace_launch(A + o_ws, n_faces, d_dirs, (const float*)(A + o_origin), n_rays, height, (float*)(A + o_ep),
                           (int*)(A + o_ec), (float*)(A + o_range), (float*)(A + o_erem), (int*)(A + o_id), ray_flags, s);
      if (rc) return rc;
      VL_CUDA_CHECK(cudaMemcpyAsync(A + o_st, A + o_ws + offsetof(VlHeader, n_bad_faces), sizeof(int), cudaMemcpyDeviceToDevice, s));
      VL_CUDA_CHECK(cudaMemsetAsync(A + o_st + 4, 0, 4, s));
    } else {
      rc = vl_cast_launch(c.cache, (const float*)(A + o_verts), (const int*)(A + o_faces), (const int*)(A + o_colors),
                          (const float*)(A + o_rem), n_verts, n_faces, (const float*)(A + o_origin), n_rays, height,
                          (float*)(A + o_ep), (int*)(A + o_ec), (float*)(A + o_range), (float*)(A + o_erem), (int*)(A + o_id),
                          0, A + o_ws, s);
      if (rc) return rc;
      VL_CUDA_CHECK(cudaMemcpyAsync(A + o_st, A + o_ws + 16, 16, cudaMemcpyDeviceToDevice, s));   // {n_bad_faces, overflow, ..} as k_cast_resolve left them
    }
    // (4) one copy back
    VL_CUDA_CHECK(cudaMemcpyAsync(P + o_ep, A + o_ep, io_bytes - o_ep, cudaMemcpyDeviceToHost, s));
    VL_CUDA_CHECK(cudaStreamSynchronize(s));
    memcpy(st, P + o_st, 16);
    if (!lbvh && st[1]) {   // overflow: nothing was written; make room for the blob and take the tree
      size_t need = o_ws + ws_lbvh;
      if (need > c.arena_bytes) {
        // growing frees the arena: stage the inputs again afterwards
        rc = grow(&c.arena, &c.arena_bytes, need, false);
        if (rc) return rc;
        A = c.arena;
        VL_CUDA_CHECK(cudaMemcpyAsync(A, P, in_bytes, cudaMemcpyHostToDevice, s));
      }
      continue;
    }
    break;
  }
  const double t3 = now_ms();
  // (5) merge the hits into the caller's buffers (misses leave them untouched, RayTracer.cpp:72-90)
  const float* h_ep = reinterpret_cast<const float*>(P + o_ep);
  const int* h_ec = reinterpret_cast<const int*>(P + o_ec);
  const float* h_range = reinterpret_cast<const float*>(P + o_range);
  const float* h_erem = reinterpret_cast<const float*>(P + o_erem);
  const int* h_id = reinterpret_cast<const int*>(P + o_id);
  constexpr size_t kMergeBlock = 16384;
  WorkPool::get().run((int)((nr + kMergeBlock - 1) / kMergeBlock), 1 << 30, [&](int b) {
    const size_t r0 = (size_t)b * kMergeBlock, r1 = r0 + kMergeBlock < nr ? r0 + kMergeBlock : nr;
    size_t r = r0;
    while (r < r1) {   // runs of hits leave as block copies
      while (r < r1 && h_id[r] < 0) ++r;
      size_t e = r;
      while (e < r1 && h_id[e] >= 0) ++e;
      if (e > r) {
        memcpy(endpoints + 3 * r, h_ep + 3 * r, 12 * (e - r));
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

extern "C" void vl_ctrace_method(int method) { g_ctrace_method.store(method == 1 ? 1 : 0); }
extern "C" void vl_ctrace_normalize(int mode) { g_ctrace_normalize.store(mode == 1 ? 1 : 0); }
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
  if (rc != VL_OK && rc != VL_EBADMESH) g_ctx.cache_valid = false;
  return rc;
}

extern "C" void ctrace(float* rays, float* origin, float* verts, int* faces, int* colors, float* rem, int n_rays,
                       int n_verts, int n_faces, int height, float* endpoints, int* endcolors, float* range,
                       float* endrem) {
  g_ctrace_status = vl_ctrace_ids(rays, origin, verts, faces, colors, rem, n_rays, n_verts, n_faces, height, endpoints,
                                  endcolors, range, endrem, nullptr);
  if (g_ctrace_status != VL_OK) fprintf(stderr, "libvlidar ctrace error %d: %s\n", g_ctrace_status, vl_last_error());
}
