/*
 * vlidar.h -- C ABI of libvlidar.so, the B200 (sm_100a) virtual-LiDAR hot path.
 *
 * Drop-in boundary for PRBonn/lidar_transfer (reference file:line in each comment).
 * Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Two layers:
 *   1. `ctrace`  -- the reference's own extern "C" entry point, HOST pointers, same
 *                   signature and semantics (auxiliary/raytracer/RayTracer.cpp:116-124,
 *                   bound by auxiliary/raytracer/RayTracerCython.pyx:5-7).
 *   2. `vl_*`    -- the device-pointer API the Python host side (torch tensors) calls:
 *                   stream-ordered, no hidden allocation, never throws.  Returns 0 on
 *                   success or a negative VL_E* code; vl_last_error() has the text.
 *
 * There is no CPU fallback: every entry point fails loudly when no CUDA device is usable.
 */
#ifndef VLIDAR_H_
#define VLIDAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VL_ABI_VERSION 2

#define VL_OK            0
#define VL_EINVAL       -1   /* bad argument (null pointer, negative size, misaligned blob) */
#define VL_ENOSPACE     -2   /* caller-provided workspace too small */
#define VL_ECUDA        -3   /* CUDA runtime / launch error (text in vl_last_error) */
#define VL_EBADMESH     -4   /* face index outside [0, n_verts) detected on device */

/* cudaStream_t passed as an opaque pointer (torch.cuda.current_stream().cuda_stream). */
typedef void* vl_stream;

int         vl_abi_version(void);
const char* vl_last_error(void);   /* thread-local, never NULL */
int         vl_device_count(void); /* >0 or VL_ECUDA */

/* ------------------------------------------------------------------------------------
 * (0) reference-compatible host entry point.
 *
 * Replaces: void ctrace(...)  auxiliary/raytracer/RayTracer.cpp:116-124 (-> trace(), :19-114).
 * All pointers are HOST pointers to caller-owned contiguous buffers.  Uploads the indexed
 * mesh, casts n_rays rays (width = n_rays / height, RayTracer.cpp:56) from
 * origin[3] and, for HITS ONLY, writes endpoints[3r..], endcolors[3r..] (= colours of the
 * hit triangle's vertex 0), endrem[r] (= mean remission of its three vertices) and
 * range[r]; entries of missing rays are left untouched (the caller zero-fills,
 * auxiliary/fusion_lidar.py:440-447).  n_verts is used to validate face indices (the
 * reference ignores it; a bad index there is undefined behaviour).  Synchronous.
 * Like the reference it returns void; failures are reported through vl_last_error() and
 * vl_ctrace_status().  Unlike the reference it prints nothing to stdout.
 * ---------------------------------------------------------------------------------- */
void ctrace(float* rays, float* origin, float* verts, int* faces, int* colors, float* rem,
            int n_rays, int n_verts, int n_faces, int height,
            float* endpoints, int* endcolors, float* range, float* endrem);
int  vl_ctrace_status(void);       /* status of this thread's most recent ctrace() */

/* Same as ctrace plus a nullable per-ray triangle-id output (-1 = miss, written for every
 * ray) and an explicit return code. */
int vl_ctrace_ids(const float* rays, const float* origin, const float* verts, const int* faces,
                  const int* colors, const float* rem, int n_rays, int n_verts, int n_faces,
                  int height, float* endpoints, int* endcolors, float* range, float* endrem,
                  int* tri_id);

/* ------------------------------------------------------------------------------------
 * Ray normalisation on the HOST, the reference's own arithmetic.
 *
 * Replaces: normalize() auxiliary/raytracer/Vector3.h:73-89 as the Ray constructor applies it (Ray.h:11-12):
 * D = (x*x + y*y) + (z*z + 0) (the two hadd steps), r = rsqrtps(D), one Newton step
 * r = 1.5*r + ((D * -0.5) * r) * (r * r), d = (x, y, z) * r -- every operation rounded separately (no FMA), the
 * reciprocal square root estimate being the x86 instruction itself, so the result is bit for bit what the reference
 * computes on the same host.  rays / out: HOST float32[3*n_rays] (may alias).  The rays are a per-sensor constant
 * (create_rays, auxiliary/laserscan.py:1092-1119): normalise once, upload, and pass VL_RAYS_NORMALIZED to
 * vl_beams_build / vl_trace so that the device's triangle test sees the reference's own unit vectors; without the flag
 * the device normalises with IEEE 1/sqrt (<= 2 ulp away, which decides beams through shared edges differently).
 * Returns VL_OK, or VL_EINVAL on a host without SSE (callers then stay with the device's IEEE normalisation).
 * ---------------------------------------------------------------------------------- */
int vl_normalize_rays(const float* rays, int n_rays, float* out);

/* ------------------------------------------------------------------------------------
 * (i) LBVH build over the per-scan triangle mesh.
 *
 * Replaces: Triangle construction RayTracer.cpp:32-51 + BVH::BVH/build BVH.cpp:112-243.
 * d_* are DEVICE pointers.  The BVH (nodes, sorted triangle records, per-triangle vertex-0
 * colours) and all build temporaries live in ONE caller-provided device blob of at least
 * vl_bvh_blob_bytes(n_faces) bytes, 256-byte aligned.  Input arrays are not referenced
 * after the build kernels finish.
 * ---------------------------------------------------------------------------------- */
size_t vl_bvh_blob_bytes(int n_faces);
int vl_bvh_build(const float* d_verts, const int* d_faces, const int* d_colors, const float* d_rem,
                 int n_verts, int n_faces, void* d_blob, size_t blob_bytes, vl_stream stream);
/* Synchronises the stream and reads the build status word: VL_OK or VL_EBADMESH.
 * info (nullable, int[8]): [0] n_tris [1] root ref [2] n_bad_faces [3] max climb depth. */
int vl_bvh_status(const void* d_blob, int n_faces, vl_stream stream, int* info);

/* ------------------------------------------------------------------------------------
 * (ii) closest-hit traversal + Moller-Trumbore.
 *
 * Replaces: the ray loop RayTracer.cpp:62-92, BVH::getIntersection BVH.cpp:19-110,
 * BBox::intersect BBox.cpp:52-100, Triangle::getIntersection Triangle.h:27-50,
 * normalize Vector3.h:73-89 (IEEE 1/sqrt instead of rsqrtps+NR).
 * d_rays float32[3*n_rays] (not normalised; unit vectors with VL_RAYS_NORMALIZED), d_origin float32[3] on the device.
 * Outputs as in ctrace (hits only) except d_tri_id (nullable): written for every ray,
 * original face index or -1.  Exact-t ties go to the smaller face index.
 * flags: VL_TRACE_ZERO_MISSES writes 0 to all four outputs of a missing ray (what the
 * reference caller obtains by zero-filling first, auxiliary/fusion_lidar.py:440-447).
 * ---------------------------------------------------------------------------------- */
#define VL_TRACE_ZERO_MISSES 1
#define VL_TRACE_PACKET      2   /* warp-packet traversal (8x4 beam tiles share one stack) instead of per-ray stacks */
#define VL_RAYS_NORMALIZED   4   /* d_rays hold unit directions already (vl_normalize_rays): used as given, not re-normalised */
#define VL_TRACE_PERSISTENT 16   /* vl_trace: persistent warps pulling rays from a counter, warp-wide ray compaction, top of the tree staged in shared memory by TMA (cp.async.bulk + mbarrier) */
#define VL_COLORS_U8         8   /* vl_cast only: d_colors is uint8[3*n_verts] (what vl_mesh_emit writes) instead of int32[3*n_verts] */
int vl_trace(const void* d_blob, int n_faces, const float* d_rays, const float* d_origin,
             int n_rays, int height, float* d_endpoints, int* d_endcolors, float* d_range,
             float* d_endrem, int* d_tri_id, int flags, vl_stream stream);

/* ------------------------------------------------------------------------------------
 * (ii-b) beam index + scene-streaming cast: the same reference code as (i) + (ii) together
 * (RayTracer.cpp:32-92, BVH.cpp:19-243, Triangle.h:27-50), for the case the ctrace ABI
 * actually describes -- ALL rays share one origin (RayTracer.cpp:116-124) and every mesh is
 * cast once.  The rays (the sensor's beams, constant across scans) are indexed once in a
 * direction-space cell grid; each scan's triangles are then streamed through that index a
 * single time, candidates tested with the same Moller-Trumbore arithmetic as vl_trace, the
 * closest hit per beam kept by a 64-bit atomicMin on (t, face index).  Identical results to
 * vl_bvh_build + vl_trace (closest hit over all triangles, exact-t ties to the smaller face
 * index), no per-scan tree.
 *
 * vl_beams_build: d_rays f32[3*n_rays] (not normalised; unit vectors with VL_RAYS_NORMALIZED) -> d_beams, a caller-provided device
 * blob of vl_beams_bytes(n_rays, height) bytes, 256-byte aligned; valid for every later
 * vl_cast with the same (n_rays, height), any origin, any mesh.
 * vl_cast: mesh arrays as in vl_bvh_build -- or d_faces = NULL for a triangle SOUP, face f = vertices (3f, 3f+1, 3f+2), the
 * layout vl_mesh_emit produces: no index array is read (or has to be written) at all; outputs / flags as in vl_trace; d_workspace of
 * vl_cast_workspace_bytes(n_rays, n_faces) bytes (8 B per ray + 12 B per face, 256-byte
 * aligned) is scratch for this call.  vl_cast_status synchronises the stream and returns
 * VL_OK, VL_EBADMESH or VL_ENOSPACE (more than 2^36 triangle-cell candidates: results
 * invalid, use the LBVH path) for the most recent vl_cast on that workspace; info (nullable,
 * int[8]): [0] n_bad_faces [1] triangles that can be hit [2],[3] candidate items (low 31
 * bits, high bits).
 * ---------------------------------------------------------------------------------- */
size_t vl_beams_bytes(int n_rays, int height);
int vl_beams_build(const float* d_rays, int n_rays, int height, void* d_beams, size_t beams_bytes,
                   int flags /* 0 or VL_RAYS_NORMALIZED */, vl_stream stream);
size_t vl_cast_workspace_bytes(int n_rays, int n_faces);
int vl_cast(const void* d_beams, const float* d_verts, const int* d_faces, const int* d_colors,
            const float* d_rem, int n_verts, int n_faces, const float* d_origin, int n_rays,
            int height, float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem,
            int* d_tri_id, int flags, void* d_workspace, size_t workspace_bytes, vl_stream stream);
int vl_cast_status(const void* d_workspace, vl_stream stream, int* info);
/* vl_cast for one scan of a batch in a single call (the per-scan host cost is what bounds a batch at this kernel
 * speed): if ev_ready (cudaEvent_t) is given it is recorded on `producer` (the stream that produced the mesh) and
 * `stream` waits for it; then vl_cast; then, if h_status (pinned host int[4]) is given, bytes 16..31 of
 * the workspace header ([0] n_bad_faces, [1] overflow flag of this scan) are copied to it; then ev_done (nullable) is recorded. */
int vl_cast_submit(const void* d_beams, const float* d_verts, const int* d_faces, const int* d_colors,
                   const float* d_rem, int n_verts, int n_faces, const float* d_origin, int n_rays,
                   int height, float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem,
                   int* d_tri_id, int flags, void* d_workspace, size_t workspace_bytes, vl_stream stream,
                   vl_stream producer, void* ev_ready, int* h_status, void* ev_done);
/* The same scan as ONE graph launch.  vl_cast_graph_create captures vl_cast for a stream slot whose beams, origin,
 * outputs and workspace (sized for max_faces) are fixed; the mesh is whatever the 64-byte descriptor at h_desc
 * (PINNED host memory: {const float* verts; const int* faces; const int* colors; const float* rem; int n_verts;
 * int n_faces; 24 bytes unused}, device pointers) holds when the graph's first node copies it to the device -- the caller
 * rewrites it before every vl_cast_graph_launch and not before the previous launch on that slot has finished.
 * h_status (nullable, pinned int[4]) receives the scan's counters as in vl_cast_submit; `stream`
 * must be idle during creation (it is captured).  vl_cast_graph_launch: optional producer wait as in vl_cast_submit,
 * the launch, then ev_done (nullable) is recorded. */
int vl_cast_graph_create(const void* d_beams, const float* d_origin, int n_rays, int height,
                         float* d_endpoints, int* d_endcolors, float* d_range, float* d_endrem,
                         int* d_tri_id, int flags, void* d_workspace, size_t workspace_bytes, int max_faces,
                         const void* h_desc, int* h_status, vl_stream stream, void** out_graph);
int vl_cast_graph_launch(void* graph, vl_stream stream, vl_stream producer, void* ev_ready, void* ev_done);
int vl_cast_graph_destroy(void* graph);

/* Which device path the host-pointer ctrace / vl_ctrace_ids uses: 0 (default) = beam index +
 * vl_cast, 1 = vl_bvh_build + vl_trace.  Process-wide. */
void vl_ctrace_method(int method);
/* How ctrace / vl_ctrace_ids normalise the rays: 0 (default) = vl_normalize_rays on the host (the reference's bits),
 * 1 = IEEE 1/sqrt on the device.  Process-wide. */
void vl_ctrace_normalize(int mode);
/* How often ctrace found the previous call's rays again (beam index reused) / had to rebuild it. */
void vl_ctrace_cache_stats(long long* hits, long long* misses);
/* Host wall time (ms) of the phases of the most recent ctrace: [0] beam index (re)build, [1] staging of the mesh
 * (+ comparison of the rays with the cached sensor) + host->device issue, [2] cast + device->host (wait), [3] merge of the
 * hits into the caller's buffers. */
void vl_ctrace_timing(double* ms4);
/* How the mesh of a ctrace call crosses PCIe: 1 (default) = packed by the staging copy (faces 3 x 21 bits in 8 B when
 * n_verts <= 2^21, colours 3 B per vertex when every component is in 0 .. 255, anything that does not fit travels raw;
 * of the results the end points stay on the device and are recomputed on the host as o + d * t from the unit directions
 * the host made itself, BVH.cpp:106-107), 0 = every array as the caller holds it.  Same results either way.  Process-wide. */
void vl_ctrace_wire(int packed);
/* Test hook, host code only (no device needed): the packing loops of the staging copy on caller-provided host buffers --
 * packed_faces u64[n_faces] (three 21-bit indices per word), packed_colors u8[n_components]; seen2[0] / seen2[1] = OR of every
 * face index / colour component (bits above 2^21 / 2^8 set: the array does not fit and ctrace sends it raw). */
int vl_debug_pack(const int* faces, long long n_faces, unsigned long long* packed_faces, const int* colors, long long n_components,
                  unsigned char* packed_colors, unsigned int* seen2);
/* Bytes the most recent ctrace moved host->device / device->host. */
void vl_ctrace_traffic(long long* h2d_bytes, long long* d2h_bytes);

/* Test aid: same outputs by testing every triangle per ray (no BVH). */
int vl_trace_bruteforce(const float* d_verts, const int* d_faces, const int* d_colors,
                        const float* d_rem, int n_verts, int n_faces, const float* d_rays,
                        const float* d_origin, int n_rays, int height, float* d_endpoints,
                        int* d_endcolors, float* d_range, float* d_endrem, int* d_tri_id,
                        int flags /* 0 or VL_RAYS_NORMALIZED */, vl_stream stream);

/* ------------------------------------------------------------------------------------
 * (iii) spherical range-image projection (atomicMin-on-depth scatter).
 *
 * Replaces: LaserScan.do_range_projection_new auxiliary/laserscan.py:294-442 (method="depth" :369-391; the other two
 * methods through vl_project_select)
 * and SemLaserScan.do_label_projection_new :672-676.
 * d_points float64[3*n] (the reference holds float64 after the pose round trip),
 * d_remissions float32[n], d_labels uint32[n].  fov in degrees.  remove != 0 applies the
 * vertical-FOV filter.  Outputs: d_range f32[H*W] (0 empty), d_index i32[H*W] (-1 empty;
 * index into the KEPT point list, like the reference after remove_points), d_label
 * i32[H*W] (0 empty), d_rem f32[H*W] (-1 empty), d_keep u8[n] (nullable), d_n_kept
 * int[1] (nullable).  Workspace: vl_project_workspace_bytes(n, H, W) device bytes.
 * ---------------------------------------------------------------------------------- */
size_t vl_project_workspace_bytes(long n_points, int H, int W);
int vl_project(const double* d_points, const float* d_remissions, const uint32_t* d_labels,
               long n_points, double fov_up_deg, double fov_down_deg, int H, int W, int remove,
               float* d_range, int32_t* d_index, int32_t* d_label, float* d_rem,
               uint8_t* d_keep, int* d_n_kept, void* d_workspace, size_t workspace_bytes,
               vl_stream stream);
/* The same with the optional `beam_angles` step of auxiliary/laserscan.py:321-327: each point's pitch is replaced by
 * the entry of d_beam_angles f64[n_beam_angles] (device) nearest to it -- first minimum of |pitch - entry|, like
 * numpy's argmin; the pitch in radians is compared with the list as given, as the reference does -- before the image
 * row is computed.  n_beam_angles == 0 is vl_project. */
int vl_project_snap(const double* d_points, const float* d_remissions, const uint32_t* d_labels,
                    long n_points, double fov_up_deg, double fov_down_deg, int H, int W, int remove,
                    const double* d_beam_angles, int n_beam_angles,
                    float* d_range, int32_t* d_index, int32_t* d_label, float* d_rem,
                    uint8_t* d_keep, int* d_n_kept, void* d_workspace, size_t workspace_bytes,
                    vl_stream stream);
/* The same with the reference's `method` argument (auxiliary/laserscan.py:294-295):
 *   VL_PROJECT_DEPTH      'depth'     :369-391  the nearest point per pixel (what deform() uses; = vl_project_snap);
 *   VL_PROJECT_PDIST      'pdist'     :392-416  the point whose image position (proj_y, proj_x) is nearest to the pixel
 *                                     centre -- like 'depth' a float64 quantity compared with the float32 image it was last
 *                                     stored in, so the same single atomicMin key reproduces the sequential loop; d_range holds the
 *                                     winner's DEPTH (:405).  d_rem holds the winner's remission (the reference's 'pdist'
 *                                     never writes proj_remissions; the Python mirror leaves it at -1);
 *   VL_PROJECT_DEPTHFAST  'depthfast' :418-437  the nearest point per pixel by descending argsort + fancy-index assignment:
 *                                     float64 depths compared exactly, empty pixels of d_range are -1 (proj_range's initial
 *                                     value); among EQUAL depths the reference's winner is whatever numpy's unstable
 *                                     argsort leaves last -- here the smallest index.
 * Pinned by tests/golden/golden_methods_v1.npz (the reference's own Python on the fixture scan). */
#define VL_PROJECT_DEPTH     0
#define VL_PROJECT_PDIST     1
#define VL_PROJECT_DEPTHFAST 2
int vl_project_select(const double* d_points, const float* d_remissions, const uint32_t* d_labels,
                      long n_points, double fov_up_deg, double fov_down_deg, int H, int W, int remove,
                      const double* d_beam_angles, int n_beam_angles, int method,
                      float* d_range, int32_t* d_index, int32_t* d_label, float* d_rem,
                      uint8_t* d_keep, int* d_n_kept, void* d_workspace, size_t workspace_bytes,
                      vl_stream stream);

/* Axis-aligned bounds of the points the projection kept: replaces SemLaserScan.get_bnds (auxiliary/laserscan.py:
 * np.amin / np.amax over the points after remove_points), which `mergemesh` clips the volume to (laserscan.py:957-962).
 * d_keep (nullable: every point) is vl_project's keep mask; d_bounds6 f64[6] = min x y z, max x y z (exact). */
int vl_points_bounds(const double* d_points, const uint8_t* d_keep, long n_points, double* d_bounds6, vl_stream stream);

/* Reverse projection of the `cp` adaption: replaces LaserScan.do_reverse_projection_new
 * auxiliary/laserscan.py:475-501.  d_depth_im f32[H*W], d_proj_x / d_proj_y f64[H*W] (image coordinates of each
 * pixel's point in [0, W] / [0, H], float or clamped) -> d_back_points f64[3*H*W], float64 arithmetic like numpy's
 * (sin / cos of the CUDA library: <= 2 ulp from the host's libm). */
int vl_reverse_project(const float* d_depth_im, const double* d_proj_x, const double* d_proj_y, int H, int W,
                       double fov_up_deg, double fov_down_deg, double* d_back_points, vl_stream stream);

/* ------------------------------------------------------------------------------------
 * (iv) class-aware TSDF voxel integration.
 *
 * Replaces: the pycuda `integrate` kernel auxiliary/fusion_lidar.py:66-229 and its launch
 * :252-287; vl_tsdf_init replaces the host-side volume initialisation + upload :48-63.
 * Volumes are float32[dx*dy*dz], C order (z fastest).  d_color_im is the folded single
 * channel image (label * 65536, fusion_lidar.py:259-264).
 * ---------------------------------------------------------------------------------- */
int vl_tsdf_init(float* d_tsdf, float* d_weight, float* d_color, float* d_rem,
                 long long n_voxels, vl_stream stream);
int vl_tsdf_integrate(float* d_tsdf, float* d_weight, float* d_color, float* d_rem,
                      int dx, int dy, int dz, const float vol_origin[3], float voxel_size,
                      float trunc_margin, float obs_weight, float fov_up_deg, float fov_down_deg,
                      const float* d_color_im, const float* d_depth_im, const float* d_rem_im,
                      int im_h, int im_w, vl_stream stream);
/* Same result bit for bit, ~1/3 faster: the arctangent, the double-precision image coordinate and the pixel
 * column (fusion_lidar.py:125, 134-141) depend on a voxel's x and y only and are computed once per z column into
 * d_workspace (vl_tsdf_workspace_bytes(dx, dy) = 4 B per column, scratch for this call). */
size_t vl_tsdf_workspace_bytes(int dx, int dy);
int vl_tsdf_integrate_ws(float* d_tsdf, float* d_weight, float* d_color, float* d_rem,
                         int dx, int dy, int dz, const float vol_origin[3], float voxel_size,
                         float trunc_margin, float obs_weight, float fov_up_deg, float fov_down_deg,
                         const float* d_color_im, const float* d_depth_im, const float* d_rem_im,
                         int im_h, int im_w, void* d_workspace, size_t workspace_bytes, vl_stream stream);
/* vl_tsdf_init followed by vl_tsdf_integrate_ws in ONE pass over the volume (the first integration into a new
 * TSDFVolume, which is every integration of the default `mergemesh` adaption, laserscan.py:968-975): the previous
 * content of the four volumes is neither read nor assumed, every voxel is written.  Same bits as the two calls. */
int vl_tsdf_init_integrate(float* d_tsdf, float* d_weight, float* d_color, float* d_rem,
                           int dx, int dy, int dz, const float vol_origin[3], float voxel_size,
                           float trunc_margin, float obs_weight, float fov_up_deg, float fov_down_deg,
                           const float* d_color_im, const float* d_depth_im, const float* d_rem_im,
                           int im_h, int im_w, void* d_workspace, size_t workspace_bytes, vl_stream stream);
/* Workspace for vl_tsdf_init_integrate's shell sweep: the column table plus 8 B per image pixel.  With at least this
 * much workspace (and |fov| <= 35 deg, dy * dz <= 2^24, dx <= 65535, fewer than ~40000 image rows per radian) the fused
 * first integration brackets every voxel (reciprocal square root, arcsine series to s^15, image row to a few thousandths) against the range image's
 * per-pixel [depth, depth + trunc] shell and runs the reference arithmetic only where the bracket cannot rule out an
 * update; bit-identical to vl_tsdf_init + vl_tsdf_integrate.  With vl_tsdf_workspace_bytes(dx, dy) only, or outside
 * those limits, every voxel takes the reference arithmetic.  vl_tsdf_integrate_ws with this much workspace takes the same
 * sweep for LATER integrations (free space is skipped only for voxels never written).  vl_debug_tsdf_shell(0) turns the shell sweep off, (2) runs
 * it with one voxel per thread instead of four (the variant for dz % 4 != 0); tests compare all of them. */
size_t vl_tsdf_fresh_workspace_bytes(int dx, int dy, int im_h, int im_w);
void vl_debug_tsdf_shell(int mode);

/* SPARSE (blocked) volumes -- the reference's own TODO, auxiliary/fusion_lidar.py:45 ("larger voxel volume ... by
 * splitting"), instead of its four dense arrays initialised per scan (:48-63).  d_hull i32[dx*dy] holds, per z column,
 * the interval [z_lo, z_hi] of voxels that are materialised in the four arrays (z_lo | z_hi << 16; 1 | 0 << 16 = none);
 * a voxel outside its column's hull is in the initial state by definition (tsdf 1, weight / colour / remission 0) and
 * its memory is never read or written.  vl_tsdf_sparse_integrate: the integration of vl_tsdf_integrate into such a
 * volume -- flags & VL_TSDF_FRESH: a NEW volume (the arrays' and d_hull's previous content is ignored: no reset pass at all);
 * flags & VL_TSDF_TABLES_VALID: the caller's promise that d_workspace still holds the geometry tables (pixel column per
 * z column, tangents per image row) an earlier call with the SAME dims, origin, voxel size, field of view and image size
 * left there -- they depend on nothing else and are not rebuilt (20 us per integration at 2000 x 1420 columns);
 * the hull of a column grows to cover every voxel the reference kernel could change (a conservative interval derived
 * from the range image), voxels entering a hull are written (initial or integrated value), voxels already in it are
 * integrated like a later scan, nothing else is touched.  Same values as the dense calls for every voxel, bit for bit
 * (vl_tsdf_densify materialises the rest: afterwards the arrays are the dense volumes and every hull is the whole
 * column).  Outside the sweep's limits (see vl_tsdf_fresh_workspace_bytes; also dz <= 32767, fov_up >= 0 >= fov_down)
 * the call densifies and takes the dense path.  vl_mesh_count_sparse / vl_mesh_emit_sparse read sparse volumes. */
#define VL_TSDF_FRESH        1
#define VL_TSDF_TABLES_VALID 2
size_t vl_tsdf_sparse_workspace_bytes(int dx, int dy, int im_h, int im_w);
int vl_tsdf_sparse_integrate(float* d_tsdf, float* d_weight, float* d_color, float* d_rem,
                             int dx, int dy, int dz, const float vol_origin[3], float voxel_size,
                             float trunc_margin, float obs_weight, float fov_up_deg, float fov_down_deg,
                             const float* d_color_im, const float* d_depth_im, const float* d_rem_im,
                             int im_h, int im_w, int* d_hull, int flags /* VL_TSDF_* */, void* d_workspace, size_t workspace_bytes,
                             vl_stream stream);
int vl_tsdf_densify(float* d_tsdf, float* d_weight, float* d_color, float* d_rem, int dx, int dy, int dz,
                    int* d_hull, vl_stream stream);

/* ------------------------------------------------------------------------------------
 * (v) iso-surface extraction + per-vertex label / remission lookup ("next" row N1).
 *
 * Replaces: TSDFVolume.get_volume + get_mesh, auxiliary/fusion_lidar.py:395-424 (D2H of three
 * volumes, scikit-image marching_cubes_lewiner at level 0 on the CPU, numpy vertex lookup).
 * Two calls because the output size is data dependent: vl_mesh_count sweeps the volume,
 * leaves one bit per voxel (value < level) and per-unit offsets in the workspace and the grand
 * totals in d_totals (device long long[2]: [0] triangles T, [1] active cubes A); the caller reads
 * them, allocates the outputs plus a scratch list of vl_mesh_list_bytes(T, A) bytes (8 B per active
 * cube + 4 B per 256 triangles), and calls vl_mesh_emit with the SAME workspace (n_tris / n_active
 * below the counted totals truncate the output to the first n_tris triangles in cube order; nothing
 * is written beyond the sizes they imply).  Output is a triangle soup in cube order: d_verts f32[9T] (world frame,
 * verts * voxel_size + origin, :412), d_faces i32[3T] = 0..3T-1 (nullable: vl_cast takes a soup without it), d_norms f32[9T] (nullable,
 * flat normals), d_colors u8[9T] = (r, g, b) of the nearest voxel's folded colour with the
 * reference's uint8 wrap (:417-423), d_rem_out f32[3T].
 * ---------------------------------------------------------------------------------- */
size_t vl_mesh_workspace_bytes(int dx, int dy, int dz);   /* ~1 bit per voxel + 24 B per 2048 voxels */
size_t vl_mesh_list_bytes(long long n_tris, long long n_active);
int vl_mesh_count(const float* d_tsdf, int dx, int dy, int dz, float level, void* d_workspace,
                  size_t workspace_bytes, long long* d_totals, vl_stream stream);
int vl_mesh_emit(const float* d_tsdf, const float* d_color, const float* d_rem, int dx, int dy, int dz,
                 float level, float voxel_size, const float vol_origin[3], const void* d_workspace,
                 size_t workspace_bytes, long long n_tris, long long n_active, void* d_active_list,
                 float* d_verts, int* d_faces, float* d_norms, unsigned char* d_colors, float* d_rem_out,
                 vl_stream stream);

/* The same on a SPARSE volume (vl_tsdf_sparse_integrate): d_hull i32[dx*dy] as there; voxels outside their column's hull
 * count as the initial values (tsdf 1, colour / remission 0) and are not read.  Same mesh, bit for bit and in the same
 * order, as vl_tsdf_densify + vl_mesh_count / vl_mesh_emit; the cube sweeps skip the empty part of the volume.
 * Needs level <= 1 and dz < 1984. */
int vl_mesh_count_sparse(const float* d_tsdf, int dx, int dy, int dz, float level, const int* d_hull,
                         void* d_workspace, size_t workspace_bytes, long long* d_totals, vl_stream stream);
int vl_mesh_emit_sparse(const float* d_tsdf, const float* d_color, const float* d_rem, int dx, int dy, int dz,
                        float level, float voxel_size, const float vol_origin[3], const int* d_hull,
                        const void* d_workspace, size_t workspace_bytes, long long n_tris, long long n_active,
                        void* d_active_list, float* d_verts, int* d_faces, float* d_norms, unsigned char* d_colors,
                        float* d_rem_out, vl_stream stream);

/* ------------------------------------------------------------------------------------
 * (vi) identity re-render metrics on the device ("next" row N3).
 *
 * Replaces: compare() auxiliary/laserscan.py:1181-1301 (masks, difference images, label renumbering) and
 * iouEval.addBatch auxiliary/np_ioueval.py:32-45 (confusion matrix).  All images are flat device arrays of
 * n_pixels = H*W entries: colours f32[3*n], labels i32[n] (values in [0, 65536)), ranges / remissions f32[n].
 * Outputs: d_label_diff f32[3*n], d_range_diff f32[n], d_rem_diff f32[n] (the reference's three images) and
 * d_conf i64[n_classes^2] -- conf[p][t] counts pixels whose renumbered TARGET label (the prediction, rows) is p
 * and renumbered SOURCE label is t, labels renumbered to their rank in the sorted union of the masked labels
 * (:1217-1223).  vl_compare_status synchronises and returns info[0] = number of distinct labels, info[1] = bad-label
 * flag (VL_EINVAL: a label outside [0, 65536) or more distinct labels than classes -- the reference raises
 * IndexError there) and the sum of the squared range differences (MSE = sum / n_pixels, :1254).
 * ---------------------------------------------------------------------------------- */
size_t vl_compare_workspace_bytes(int n_pixels);
int vl_compare(const float* d_source_color, const float* d_target_color, const int* d_source_label,
               const int* d_target_label, const float* d_source_range, const float* d_target_range,
               const float* d_source_rem, const float* d_target_rem, int n_pixels, int n_classes,
               float* d_label_diff, float* d_range_diff, float* d_rem_diff, long long* d_conf,
               void* d_workspace, size_t workspace_bytes, vl_stream stream);
int vl_compare_status(const void* d_workspace, vl_stream stream, int* info, double* range_sq_sum);

/* ------------------------------------------------------------------------------------
 * measurement aids (no reference counterpart): a process-wide count of kernels launched by
 * this library, and optional per-stage device timing with CUDA events recorded on the
 * launching stream.  vl_profile_collect synchronises the device and ACCUMULATES into
 * stage_ms[] / stage_launches[] (vl_profile_stage_count() entries each).
 * ---------------------------------------------------------------------------------- */
long long   vl_launch_count(void);
int         vl_profile_enable(int on);          /* returns the previous setting */
int         vl_profile_stage_count(void);
const char* vl_profile_stage_name(int stage);
int         vl_profile_collect(double* stage_ms, long long* stage_launches);
/* Debug: while d_stats is non-NULL every vl_trace launch also writes, per ray, the pair
 * {inner nodes visited, triangles tested} to d_stats[2*r .. 2*r+1] (device int[2*n_rays]). */
void        vl_debug_trace_stats(int* d_stats);
/* Debug: force the traversal variant: 0 auto, 1 per-ray in storage order, 2 per-ray in 16x8 beam tiles,
 * 3 persistent warps + ray compaction + TMA-staged top of the tree, 4/8/16/32 = warp packets of that tile width. */
void        vl_debug_trace_mode(int mode);
/* Debug: force vl_mesh_count's one-cube-per-lane sweep (default: four cubes per lane when dz % 4 == 0). */
void        vl_debug_mesh_scalar(int on);
/* Debug: cell rows per beam row of the beam index (default 1); changes vl_beams_bytes. */
void        vl_debug_cast_cells(int cells_per_beam_row);
/* Debug: 1 = graphs created from now on have no reset kernel (k_cast_resolve re-arms the slot), 0 (default) = k_cast_init per scan. */
void        vl_debug_cast_rearm(int on);
/* Debug: 1 (default) = a triangle's rectangle drops the cell rows at both ends none of whose beams lies inside its sine
 * interval (exact: the comparison k_cast_units makes per beam), 0 = every cell row the interval touches (A/B aid). */
void        vl_debug_cast_row_trim(int on);
/* Debug: 1 = the cast's cull and setup run as two kernels (k_cast_cull at 6 CTAs per SM + k_cast_setup2 on the survivors' list),
 * 0 = one kernel with a shared-memory queue (k_cast_setup).  Same results. */
void        vl_debug_cast_split(int on);
/* Debug: persistent CTAs per SM of the item kernel (default 4). */
void        vl_debug_cast_ctas(int ctas_per_sm);
/* Debug: persistent CTAs per SM of the setup kernel (default 4). */
void        vl_debug_cast_setup_ctas(int ctas_per_sm);
/* Debug (timing only, leaves the blob unusable): 0 full build, 1 / 2 / 3 = stop after bounds / morton / sort. */
void        vl_debug_build_stop(int stage);

#ifdef __cplusplus
}
#endif
#endif /* VLIDAR_H_ */
