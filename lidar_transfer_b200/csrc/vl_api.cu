// vl_api.cu -- the device-pointer C ABI of libvlidar.so (declared in include/vlidar.h): thin validated wrappers
// around the launchers, error text, launch counter and the per-stage event profiler.  The host-pointer `ctrace`
// (the reference's own signature) lives in vl_host.cu.
#include <stdarg.h>
#include <stddef.h>
#include <string.h>
#include <mutex>
#include "vl_common.cuh"
#include <nvtx3/nvToolsExt.h>   // header-only: ranges around every stage, visible to nsys / ncu, free when no tool is attached

// ---------------------------------------------------------------------------
// error text
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

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

int vl_sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

void vl_prof_begin(int stage, cudaStream_t stream) {
  nvtxRangePushA(stage >= 0 && stage < VL_ST_COUNT ? kStageNames[stage] : "vlidar");
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, stream);
  g_prof_open[stage] = e;
}
void vl_prof_end(int stage, cudaStream_t stream) {
  nvtxRangePop();
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
                                   int* d_tri_id, int flags, vl_stream stream) {
  int rc = check_trace_args("vl_trace_bruteforce", d_rays, d_origin, n_rays, height, d_endpoints, d_endcolors, d_range, d_endrem);
  if (rc) return rc;
  if (n_faces < 0 || n_verts < 0 || (n_faces > 0 && (!d_verts || !d_faces || !d_colors || !d_rem))) {
    vl_set_error("vl_trace_bruteforce: invalid mesh");
    return VL_EINVAL;
  }
  return vl_trace_bruteforce_launch(d_verts, d_faces, d_colors, d_rem, n_verts, n_faces, d_rays, d_origin, n_rays, height,
                                    d_endpoints, d_endcolors, d_range, d_endrem, d_tri_id, flags, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------
// (ii-b) beam index + scene-streaming cast
// ---------------------------------------------------------------------------
extern "C" size_t vl_beams_bytes(int n_rays, int height) {
  return vl_beams_bytes_impl(n_rays < 0 ? 0 : n_rays, height < 1 ? 1 : height);
}

extern "C" int vl_beams_build(const float* d_rays, int n_rays, int height, void* d_beams, size_t beams_bytes,
                              int flags, vl_stream stream) {
  if (n_rays < 0 || height <= 0 || !d_beams || (((uintptr_t)d_beams) & 255) || (n_rays > 0 && !d_rays)) {
    vl_set_error("vl_beams_build: invalid argument (n_rays %d, height %d, beams %p)", n_rays, height, d_beams);
    return VL_EINVAL;
  }
  if (beams_bytes < vl_beams_bytes(n_rays, height)) {
    vl_set_error("vl_beams_build: blob too small (%zu < %zu bytes)", beams_bytes, vl_beams_bytes(n_rays, height));
    return VL_ENOSPACE;
  }
  return vl_beams_build_launch(d_rays, n_rays, height, d_beams, flags, static_cast<cudaStream_t>(stream));
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
  // d_faces == NULL: a triangle soup, face f = vertices (3f, 3f+1, 3f+2) -- what vl_mesh_emit produces
  if (n_faces < 0 || n_verts < 0 || n_faces >= (1 << 28) || (n_faces > 0 && (!d_verts || !d_colors || !d_rem))) {
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
  // bytes 16..31 of the workspace header: the scan's counters as k_cast_resolve left them for the host
  if (h_status) VL_CUDA_CHECK(cudaMemcpyAsync(h_status, static_cast<char*>(d_workspace) + 16, 16, cudaMemcpyDeviceToHost, s));
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

