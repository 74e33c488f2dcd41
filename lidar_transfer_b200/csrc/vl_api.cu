// vl_api.cu -- the C ABI of libvlidar.so (declared in include/vlidar.h).
//
// Device-pointer entry points are thin validated wrappers around the launchers.  The
// host-pointer `ctrace` keeps the reference's own signature (auxiliary/raytracer/
// RayTracer.cpp:116-124): it stages the caller's buffers into a grow-only device arena,
// builds the LBVH, traces and copies the four outputs back -- synchronous, like the reference.
#include <stdarg.h>
#include <stddef.h>
#include <string.h>
#include <mutex>
#include "vl_common.cuh"

// ---------------------------------------------------------------------------
// error text
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local int g_ctrace_status = VL_OK;

void vl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------
// launch counter + optional per-stage event timing
// ---------------------------------------------------------------------------
#include <atomic>
#include <vector>
namespace {
std::atomic<long long> g_launches{0};
std::atomic<int> g_prof_on{0};
struct ProfRec { int stage; cudaEvent_t a, b; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
thread_local cudaEvent_t g_prof_open[VL_ST_COUNT];
const char* kStageNames[VL_ST_COUNT] = {"bounds", "morton", "sort_pass", "emit_climb", "top_climb",
                                        "trace", "project_scatter", "project_gather", "tsdf_init", "tsdf_integrate",
                                        "mesh_count", "mesh_scan", "mesh_compact", "mesh_emit",
                                        "beams", "cast_init", "cast_setup", "cast_items", "cast_resolve", "compare"};
cudaEvent_t prof_event() {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void vl_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void vl_prof_begin(int stage, cudaStream_t stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, stream);
  g_prof_open[stage] = e;
}
void vl_prof_end(int stage, cudaStream_t stream) {
  if (!g_prof_on.load(std::memory_order_relaxed) || !g_prof_open[stage]) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, stream);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof_recs.push_back({stage, g_prof_open[stage], e});
  g_prof_open[stage] = nullptr;
}

extern "C" long long vl_launch_count(void) { return g_launches.load(); }
extern "C" int vl_profile_enable(int on) { return g_prof_on.exchange(on ? 1 : 0); }
extern "C" int vl_profile_stage_count(void) { return VL_ST_COUNT; }
extern "C" const char* vl_profile_stage_name(int stage) { return (stage >= 0 && stage < VL_ST_COUNT) ? kStageNames[stage] : ""; }
// Waits for the device, then adds every finished record to stage_ms / stage_launches (arrays of
// vl_profile_stage_count() entries, accumulated into, not cleared) and forgets the records.
extern "C" int vl_profile_collect(double* stage_ms, long long* stage_launches) {
  VL_CUDA_CHECK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (const ProfRec& r : g_prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      if (stage_ms) stage_ms[r.stage] += ms;
      if (stage_launches) stage_launches[r.stage] += 1;
    }
    g_prof_pool.push_back(r.a);
    g_prof_pool.push_back(r.b);
  }
  g_prof_recs.clear();
  return VL_OK;
}

extern "C" int vl_abi_version(void) { return VL_ABI_VERSION; }
extern "C" const char* vl_last_error(void) { return g_err; }
extern "C" int vl_ctrace_status(void) { return g_ctrace_status; }

extern "C" int vl_device_count(void) {
  int n = 0;
  VL_CUDA_CHECK(cudaGetDeviceCount(&n));
  return n;
}

// ---------------------------------------------------------------------------
// device API
// ---------------------------------------------------------------------------
extern "C" size_t vl_bvh_blob_bytes(int n_faces) { return vl_blob_layout(n_faces < 0 ? 0 : n_faces).total; }

extern "C" int vl_bvh_build(const float* d_verts, const int* d_faces, const int* d_colors, const float* d_rem,
                            int n_verts, int n_faces, void* d_blob, size_t blob_bytes, vl_stream stream) {
  if (n_faces < 0 || n_verts < 0 || !d_blob || (((uintptr_t)d_blob) & 255) ||
      (n_faces > 0 && (!d_verts || !d_faces || !d_colors || !d_rem)) || n_faces >= (1 << 28)) {
    vl_set_error("vl_bvh_build: invalid argument (n_verts %d, n_faces %d, blob %p)", n_verts, n_faces, d_blob);
    return VL_EINVAL;
  }
  if (blob_bytes < vl_bvh_blob_bytes(n_faces)) {
    vl_set_error("vl_bvh_build: blob too small (%zu < %zu bytes for %d faces)", blob_bytes, vl_bvh_blob_bytes(n_faces), n_faces);
    return VL_ENOSPACE;
  }
  return vl_bvh_build_launch(d_verts, d_faces, d_colors, d_rem, n_verts, n_faces, d_blob, static_cast<cudaStream_t>(stream));
}

extern "C" int vl_bvh_status(const void* d_blob, int n_faces, vl_stream stream, int* info) {
  (void)n_faces;
  if (!d_blob) { vl_set_error("vl_bvh_status: null blob"); return VL_EINVAL; }
  VlHeader h;
  VL_CUDA_CHECK(cudaMemcpyAsync(&h, d_blob, sizeof(h), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  VL_CUDA_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  if (info) { info[0] = h.n_tris; info[1] = h.root_ref; info[2] = h.n_bad_faces; info[3] = h.max_climb; }
  if (h.n_bad_faces > 0) {
    vl_set_error("mesh has %d face(s) with a vertex index outside [0, n_verts)", h.n_bad_faces);
    return VL_EBADMESH;
  }
  return VL_OK;
}

static int check_trace_args(const char* who, const float* d_rays, const float* d_origin, int n_rays, int height,
                            float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem) {
  if (n_rays < 0 || height <= 0 || !d_origin || (n_rays > 0 && (!d_rays || !d_endpoints || !d_endcolors || !d_range || !d_endrem))) {
    vl_set_error("%s: invalid argument (n_rays %d, height %d)", who, n_rays, height);
    return VL_EINVAL;
  }
  return VL_OK;
}

extern "C" int vl_trace(const void* d_blob, int n_faces, const float* d_rays, const float* d_origin, int n_rays,
                        int height, float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem,
                        int* d_tri_id, int flags, vl_stream stream) {
  int rc = check_trace_args("vl_trace", d_rays, d_origin, n_rays, height, d_endpoints, d_endcolors, d_range, d_endrem);
  if (rc) return rc;
  if (!d_blob || n_faces < 0) { vl_set_error("vl_trace: invalid BVH blob"); return VL_EINVAL; }
  return vl_trace_launch(d_blob, n_faces, d_rays, d_origin, n_rays, height, d_endpoints, d_endcolors, d_range,
                         d_endrem, d_tri_id, flags, static_cast<cudaStream_t>(stream));
}

extern "C" int vl_trace_bruteforce(const float* d_verts, const int* d_faces, const int* d_colors, const float* d_rem,
                                   int n_verts, int n_faces, const float* d_rays, const float* d_origin, int n_rays,
                                   int height, float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem,
                                   int* d_tri_id, vl_stream stream) {
  int rc = check_trace_args("vl_trace_bruteforce", d_rays, d_origin, n_rays, height, d_endpoints, d_endcolors, d_range, d_endrem);
  if (rc) return rc;
  if (n_faces < 0 || n_verts < 0 || (n_faces > 0 && (!d_verts || !d_faces || !d_colors || !d_rem))) {
    vl_set_error("vl_trace_bruteforce: invalid mesh");
    return VL_EINVAL;
  }
  return vl_trace_bruteforce_launch(d_verts, d_faces, d_colors, d_rem, n_verts, n_faces, d_rays, d_origin, n_rays, height,
                                    d_endpoints, d_endcolors, d_range, d_endrem, d_tri_id, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------
// (ii-b) beam index + scene-streaming cast
// ---------------------------------------------------------------------------
extern "C" size_t vl_beams_bytes(int n_rays, int height) {
  return vl_beams_bytes_impl(n_rays < 0 ? 0 : n_rays, height < 1 ? 1 : height);
}

extern "C" int vl_beams_build(const float* d_rays, int n_rays, int height, void* d_beams, size_t beams_bytes,
                              vl_stream stream) {
  if (n_rays < 0 || height <= 0 || !d_beams || (((uintptr_t)d_beams) & 255) || (n_rays > 0 && !d_rays)) {
    vl_set_error("vl_beams_build: invalid argument (n_rays %d, height %d, beams %p)", n_rays, height, d_beams);
    return VL_EINVAL;
  }
  if (beams_bytes < vl_beams_bytes(n_rays, height)) {
    vl_set_error("vl_beams_build: blob too small (%zu < %zu bytes)", beams_bytes, vl_beams_bytes(n_rays, height));
    return VL_ENOSPACE;
  }
  return vl_beams_build_launch(d_rays, n_rays, height, d_beams, static_cast<cudaStream_t>(stream));
}

extern "C" size_t vl_cast_workspace_bytes(int n_rays, int n_faces) { return vl_cast_workspace_bytes_impl(n_rays, n_faces); }

extern "C" int vl_cast(const void* d_beams, const float* d_verts, const int* d_faces, const int* d_colors,
                       const float* d_rem, int n_verts, int n_faces, const float* d_origin, int n_rays, int height,
                       float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem, int* d_tri_id,
                       int flags, void* d_workspace, size_t workspace_bytes, vl_stream stream) {
  if (n_rays < 0 || height <= 0 || !d_origin || !d_beams || !d_workspace || (((uintptr_t)d_workspace) & 255) ||
      (n_rays > 0 && (!d_endpoints || !d_endcolors || !d_range || !d_endrem))) {
    vl_set_error("vl_cast: invalid argument (n_rays %d, height %d)", n_rays, height);
    return VL_EINVAL;
  }
  if (n_faces < 0 || n_verts < 0 || n_faces >= (1 << 28) || (n_faces > 0 && (!d_verts || !d_faces || !d_colors || !d_rem))) {
    vl_set_error("vl_cast: invalid mesh (n_verts %d, n_faces %d; at most 2^28 - 1 faces)", n_verts, n_faces);
    return VL_EINVAL;
  }
  if (workspace_bytes < vl_cast_workspace_bytes(n_rays, n_faces)) {
    vl_set_error("vl_cast: workspace too small (%zu < %zu bytes)", workspace_bytes, vl_cast_workspace_bytes(n_rays, n_faces));
    return VL_ENOSPACE;
  }
  return vl_cast_launch(d_beams, d_verts, d_faces, d_colors, d_rem, n_verts, n_faces, d_origin, n_rays, height,
                        d_endpoints, d_endcolors, d_range, d_endrem, d_tri_id, flags, d_workspace,
                        static_cast<cudaStream_t>(stream));
}

// One scan of a batch, submitted with a single call: make `stream` wait for the producer of the mesh, cast, copy the
// first 16 bytes of the workspace header (n_bad_faces, overflow, ...) to pinned host memory, record `ev_done`.
extern "C" int vl_cast_submit(const void* d_beams, const float* d_verts, const int* d_faces, const int* d_colors,
                              const float* d_rem, int n_verts, int n_faces, const float* d_origin, int n_rays,
                              int height, float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem,
                              int* d_tri_id, int flags, void* d_workspace, size_t workspace_bytes, vl_stream stream,
                              vl_stream producer, void* ev_ready, int* h_status, void* ev_done) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ev_ready) {
    VL_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(ev_ready), static_cast<cudaStream_t>(producer)));
    VL_CUDA_CHECK(cudaStreamWaitEvent(s, static_cast<cudaEvent_t>(ev_ready), 0));
  }
  const int rc = vl_cast(d_beams, d_verts, d_faces, d_colors, d_rem, n_verts, n_faces, d_origin, n_rays, height,
                         d_endpoints, d_endcolors, d_range, d_endrem, d_tri_id, flags, d_workspace, workspace_bytes, stream);
  if (rc) return rc;
  if (h_status) VL_CUDA_CHECK(cudaMemcpyAsync(h_status, d_workspace, 16, cudaMemcpyDeviceToHost, s));
  if (ev_done) VL_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(ev_done), s));
  return VL_OK;
}

// vl_cast_submit as a replayed CUDA graph (see include/vlidar.h)
extern "C" int vl_cast_graph_create(const void* d_beams, const float* d_origin, int n_rays, int height,
                                    float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem,
                                    int* d_tri_id, int flags, void* d_workspace, size_t workspace_bytes, int max_faces,
                                    const void* h_desc, int* h_status, vl_stream stream, void** out_graph) {
  if (n_rays <= 0 || height <= 0 || max_faces < 0 || max_faces >= (1 << 28) || !d_beams || !d_origin || !d_workspace || (((uintptr_t)d_workspace) & 255) ||
      !d_endpoints || !d_endcolors || !d_range || !d_endrem || !h_desc || !out_graph) {
    vl_set_error("vl_cast_graph_create: invalid argument (n_rays %d, height %d, max_faces %d)", n_rays, height, max_faces);
    return VL_EINVAL;
  }
  if (workspace_bytes < vl_cast_workspace_bytes(n_rays, max_faces)) {
    vl_set_error("vl_cast_graph_create: workspace too small (%zu < %zu bytes)", workspace_bytes, vl_cast_workspace_bytes(n_rays, max_faces));
    return VL_ENOSPACE;
  }
  return vl_cast_graph_create_impl(d_beams, d_origin, n_rays, height, d_endpoints, d_endcolors, d_range, d_endrem, d_tri_id,
                                   flags, d_workspace, max_faces, h_desc, h_status, static_cast<cudaStream_t>(stream), out_graph);
}

extern "C" int vl_cast_graph_launch(void* graph, vl_stream stream, vl_stream producer, void* ev_ready, void* ev_done) {
  if (!graph) { vl_set_error("vl_cast_graph_launch: null graph"); return VL_EINVAL; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ev_ready) {
    VL_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(ev_ready), static_cast<cudaStream_t>(producer)));
    VL_CUDA_CHECK(cudaStreamWaitEvent(s, static_cast<cudaEvent_t>(ev_ready), 0));
  }
  const int rc = vl_cast_graph_launch_impl(graph, s);
  if (rc) return rc;
  if (ev_done) VL_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(ev_done), s));
  return VL_OK;
}

extern "C" int vl_cast_graph_destroy(void* graph) { return vl_cast_graph_destroy_impl(graph); }

extern "C" int vl_cast_status(const void* d_workspace, vl_stream stream, int* info) {
  if (!d_workspace) { vl_set_error("vl_cast_status: null workspace"); return VL_EINVAL; }
  return vl_cast_status_read(d_workspace, static_cast<cudaStream_t>(stream), info);
}

// ---------------------------------------------------------------------------
// host-pointer ctrace: grow-only device arena + one stream, guarded by a mutex
// ---------------------------------------------------------------------------
namespace {
struct HostCtx {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  char* arena = nullptr;
  size_t arena_bytes = 0;
};
HostCtx g_ctx;
std::atomic<int> g_ctrace_method{0};   // 0 = beam index + scene-streaming cast, 1 = LBVH build + traversal

int ctx_reserve(HostCtx& c, size_t bytes) {
  if (!c.stream) VL_CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  if (bytes > c.arena_bytes) {
    if (c.arena) VL_CUDA_CHECK(cudaFree(c.arena));
    c.arena = nullptr; c.arena_bytes = 0;
    size_t want = bytes + bytes / 4;
    VL_CUDA_CHECK(cudaMalloc(&c.arena, want));
    c.arena_bytes = want;
  }
  return VL_OK;
}
}  // namespace

extern "C" void vl_ctrace_method(int method) { g_ctrace_method.store(method == 1 ? 1 : 0); }

static int ctrace_run(bool lbvh, bool* overflowed, const float* rays, const float* origin, const float* verts,
                      const int* faces, const int* colors, const float* rem, int n_rays, int n_verts, int n_faces,
                      int height, float* endpoints, int* endcolors, float* range, float* endrem, int* tri_id) {
  HostCtx& c = g_ctx;
  const size_t nr = (size_t)n_rays, nv = (size_t)n_verts, nf = (size_t)n_faces;
  // arena carve-up
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = vl_align256(off + bytes); return o; };
  const size_t o_blob = take(lbvh ? vl_bvh_blob_bytes(n_faces) : vl_beams_bytes(n_rays, height));
  const size_t o_ws = take(lbvh ? 256 : vl_cast_workspace_bytes(n_rays, n_faces));
  const size_t o_rays = take(12 * nr), o_origin = take(12), o_verts = take(12 * nv), o_faces = take(12 * nf);
  const size_t o_colors = take(12 * nv), o_rem = take(4 * nv);
  const size_t o_ep = take(12 * nr), o_ec = take(12 * nr), o_range = take(4 * nr), o_erem = take(4 * nr), o_id = take(4 * nr);
  int rc = ctx_reserve(c, off);
  if (rc) return rc;
  char* A = c.arena;
  cudaStream_t s = c.stream;
  VL_CUDA_CHECK(cudaMemcpyAsync(A + o_rays, rays, 12 * nr, cudaMemcpyHostToDevice, s));
  VL_CUDA_CHECK(cudaMemcpyAsync(A + o_origin, origin, 12, cudaMemcpyHostToDevice, s));
  if (!lbvh) {   // the beam index only needs the rays: it is built while the mesh is still on its way
    rc = vl_beams_build_launch((const float*)(A + o_rays), n_rays, height, A + o_blob, s);
    if (rc) return rc;
  }
  VL_CUDA_CHECK(cudaMemcpyAsync(A + o_verts, verts, 12 * nv, cudaMemcpyHostToDevice, s));
  VL_CUDA_CHECK(cudaMemcpyAsync(A + o_faces, faces, 12 * nf, cudaMemcpyHostToDevice, s));
  VL_CUDA_CHECK(cudaMemcpyAsync(A + o_colors, colors, 12 * nv, cudaMemcpyHostToDevice, s));
  VL_CUDA_CHECK(cudaMemcpyAsync(A + o_rem, rem, 4 * nv, cudaMemcpyHostToDevice, s));
  if (lbvh) {
    rc = vl_bvh_build_launch((const float*)(A + o_verts), (const int*)(A + o_faces), (const int*)(A + o_colors),
                             (const float*)(A + o_rem), n_verts, n_faces, A + o_blob, s);
    if (rc) return rc;
  }
  // misses must leave the caller's buffers untouched (RayTracer.cpp:72-90): round-trip their content
  VL_CUDA_CHECK(cudaMemcpyAsync(A + o_ep, endpoints, 12 * nr, cudaMemcpyHostToDevice, s));
  VL_CUDA_CHECK(cudaMemcpyAsync(A + o_ec, endcolors, 12 * nr, cudaMemcpyHostToDevice, s));
  VL_CUDA_CHECK(cudaMemcpyAsync(A + o_range, range, 4 * nr, cudaMemcpyHostToDevice, s));
  VL_CUDA_CHECK(cudaMemcpyAsync(A + o_erem, endrem, 4 * nr, cudaMemcpyHostToDevice, s));
  if (lbvh)
    rc = vl_trace_launch(A + o_blob, n_faces, (const float*)(A + o_rays), (const float*)(A + o_origin), n_rays, height,
                         (float*)(A + o_ep), (int*)(A + o_ec), (float*)(A + o_range), (float*)(A + o_erem),
                         tri_id ? (int*)(A + o_id) : nullptr, 0, s);
  else
    rc = vl_cast_launch(A + o_blob, (const float*)(A + o_verts), (const int*)(A + o_faces), (const int*)(A + o_colors),
                        (const float*)(A + o_rem), n_verts, n_faces, (const float*)(A + o_origin), n_rays, height,
                        (float*)(A + o_ep), (int*)(A + o_ec), (float*)(A + o_range), (float*)(A + o_erem),
                        tri_id ? (int*)(A + o_id) : nullptr, 0, A + o_ws, s);
  if (rc) return rc;
  VL_CUDA_CHECK(cudaMemcpyAsync(endpoints, A + o_ep, 12 * nr, cudaMemcpyDeviceToHost, s));
  VL_CUDA_CHECK(cudaMemcpyAsync(endcolors, A + o_ec, 12 * nr, cudaMemcpyDeviceToHost, s));
  VL_CUDA_CHECK(cudaMemcpyAsync(range, A + o_range, 4 * nr, cudaMemcpyDeviceToHost, s));
  VL_CUDA_CHECK(cudaMemcpyAsync(endrem, A + o_erem, 4 * nr, cudaMemcpyDeviceToHost, s));
  if (tri_id) VL_CUDA_CHECK(cudaMemcpyAsync(tri_id, A + o_id, 4 * nr, cudaMemcpyDeviceToHost, s));
  int st[2] = {0, 0};   // {n_bad_faces, overflow}: the first two ints of the cast header; n_bad_faces of the LBVH header
  VL_CUDA_CHECK(cudaMemcpyAsync(st, lbvh ? A + o_blob + offsetof(VlHeader, n_bad_faces) : A + o_ws,
                                lbvh ? sizeof(int) : 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  VL_CUDA_CHECK(cudaStreamSynchronize(s));
  if (!lbvh && st[1]) {   // the cast ran out of work units before touching any output: take the tree instead
    *overflowed = true;
    return VL_OK;
  }
  const int n_bad = st[0];
  if (n_bad > 0) {
    vl_set_error("ctrace: %d face(s) reference a vertex outside [0, %d); they were skipped", n_bad, n_verts);
    return VL_EBADMESH;
  }
  return VL_OK;
}


extern "C" int vl_ctrace_ids(const float* rays, const float* origin, const float* verts, const int* faces,
                             const int* colors, const float* rem, int n_rays, int n_verts, int n_faces, int height,
                             float* endpoints, int* endcolors, float* range, float* endrem, int* tri_id) {
  if (n_rays < 0 || n_verts < 0 || n_faces < 0 || height <= 0 || !origin ||
      (n_rays > 0 && (!rays || !endpoints || !endcolors || !range || !endrem)) ||
      (n_faces > 0 && (!verts || !faces || !colors || !rem))) {
    vl_set_error("ctrace: invalid argument (n_rays %d, n_verts %d, n_faces %d, height %d)", n_rays, n_verts, n_faces, height);
    return VL_EINVAL;
  }
  int n_dev = 0;
  VL_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
  if (n_dev <= 0) { vl_set_error("ctrace: no CUDA device (libvlidar has no CPU fallback)"); return VL_ECUDA; }
  if (n_rays == 0) return VL_OK;
  std::lock_guard<std::mutex> lock(g_ctx.mu);
  bool overflowed = false;
  int rc = ctrace_run(g_ctrace_method.load() == 1, &overflowed, rays, origin, verts, faces, colors, rem, n_rays, n_verts,
                      n_faces, height, endpoints, endcolors, range, endrem, tri_id);
  if (rc == VL_OK && overflowed)
    rc = ctrace_run(true, &overflowed, rays, origin, verts, faces, colors, rem, n_rays, n_verts, n_faces, height, endpoints,
                    endcolors, range, endrem, tri_id);
  return rc;
}

extern "C" void ctrace(float* rays, float* origin, float* verts, int* faces, int* colors, float* rem, int n_rays,
                       int n_verts, int n_faces, int height, float* endpoints, int* endcolors, float* range,
                       float* endrem) {
  g_ctrace_status = vl_ctrace_ids(rays, origin, verts, faces, colors, rem, n_rays, n_verts, n_faces, height, endpoints,
                                  endcolors, range, endrem, nullptr);
  if (g_ctrace_status != VL_OK) fprintf(stderr, "libvlidar ctrace error %d: %s\n", g_ctrace_status, g_err);
}
