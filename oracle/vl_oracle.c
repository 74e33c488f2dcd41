/*
 * vl_oracle.c -- CPU restatement of the lidar_transfer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the reported CPU baseline.  The product path is the CUDA library in
 * lidar_transfer_b200/csrc and has no CPU fallback.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the reference checkout, PRBonn/lidar_transfer @ 661135b).
 *
 * Parity pinning (see tests/test_oracle_pinned.py, tests/golden/):
 *   - vlo_trace (reference-BVH mode, SSE normalise) is checked BIT-EXACT
 *     against the reference C++ ray tracer compiled from its own sources
 *     with -ffp-contract=off (oracle/_ref/libref_raytracer_nofma.so), and
 *     against the known-answer ray/triangle of auxiliary/raytracing.py:229-263.
 *   - vlo_project / vlo_project_snap are checked against the reference's own Python
 *     (auxiliary/laserscan.py imported with stub modules, tests/golden/make_golden.py and
 *     make_golden_beams.py for the beam_angles step).
 *   - vlo_tsdf_integrate is checked BIT-EXACT against the reference's CUDA
 *     kernel string (auxiliary/fusion_lidar.py:66-229) extracted at build
 *     time and compiled for the CPU (oracle/_ref/libref_tsdf.so).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__SSE__)
#include <xmmintrin.h>
#endif

/* ------------------------------------------------------------------ */
/* small float3 helpers with the reference's exact operation order     */
/* ------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;

static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
/* Vector3.h:32-34  dot = x*b.x + y*b.y + z*b.z  (left associative) */
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* Vector3.h:37-46  cross: two lane products and one subtract per lane */
static inline v3 v3_cross(v3 a, v3 b) {
  return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* SSE minps/maxps semantics (Vector3.h:58-66): second operand on NaN */
static inline float sse_min(float a, float b) { return a < b ? a : b; }
static inline float sse_max(float a, float b) { return a > b ? a : b; }
static inline v3 v3_min(v3 a, v3 b) { return v3_make(sse_min(a.x, b.x), sse_min(a.y, b.y), sse_min(a.z, b.z)); }
static inline v3 v3_max(v3 a, v3 b) { return v3_make(sse_max(a.x, b.x), sse_max(a.y, b.y), sse_max(a.z, b.z)); }
static inline float v3_get(v3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

#define VLO_NORMALIZE_SSE 1u   /* Vector3.h:73-89: rsqrtps + one Newton step (x86 only)  */
#define VLO_BRUTE_FORCE   2u   /* all triangles, tie-break (min t, min triangle index)   */
#define VLO_MIN_ID_TIES   4u   /* BVH mode: break exact-t ties by min triangle index     */

/* Vector3.h:73-89 normalize().  mode 0: IEEE 1/sqrt (what the CUDA path computes,
 * bit-reproducible on any IEEE machine); mode SSE: the reference's approximation. */
static inline v3 vlo_normalize(v3 a, unsigned flags) {
  float D = (a.x * a.x + a.y * a.y) + (a.z * a.z + 0.0f); /* hadd(hadd) order, w = 0 */
  float r;
#if defined(__SSE__)
  if (flags & VLO_NORMALIZE_SSE) {
    r = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(D)));
    r = (1.5f * r) + (((D * -0.5f) * r) * (r * r));
  } else
#endif
  {
    r = 1.0f / sqrtf(D);
  }
  return v3_scale(a, r);
}

/* ------------------------------------------------------------------ */
/* triangles  (RayTracer.cpp:32-51, Triangle.h)                         */
/* ------------------------------------------------------------------ */
typedef struct {
  v3 v0, v1, v2;
  int c0[3];   /* colour of vertex 0 (Triangle.h:56-61 getColor(0)) */
  float rem;   /* Triangle.h:63-70 (r0+r1+r2)/3 */
  v3 centroid; /* Triangle.h:78-80 */
  v3 bmin, bmax; /* Triangle.h:72-76 */
} vlo_tri;

/* Triangle.h:27-50 Moller-Trumbore, eps 1e-6, two sided, edges inclusive */
static inline int vlo_tri_hit(const vlo_tri* T, v3 o, v3 d, float* t_out) {
  v3 e1 = v3_sub(T->v1, T->v0);
  v3 e2 = v3_sub(T->v2, T->v0);
  v3 h = v3_cross(d, e2);
  float a = v3_dot(e1, h);
  const float eps = 0.000001f;
  if (a < eps && a > -eps) return 0;
  float inv_a = 1.0f / a;
  v3 s = v3_sub(o, T->v0);
  float u = v3_dot(s, h) * inv_a;
  if (u < 0 || u > 1) return 0;
  v3 q = v3_cross(s, e1);
  float v = v3_dot(d, q) * inv_a;
  if (v < 0 || u + v > 1) return 0;
  float t = v3_dot(e2, q) * inv_a;
  if (t < eps) return 0;
  *t_out = t;
  return 1;
}

/* ------------------------------------------------------------------ */
/* BVH  (BVH.h:12-15 BVHFlatNode, BVH.cpp:143-243 build)               */
/* ------------------------------------------------------------------ */
typedef struct {
  v3 bmin, bmax;
  uint32_t start, nPrims, rightOffset;
} vlo_node;

typedef struct { uint32_t parent, start, end; } vlo_build_entry;

/* BBox.cpp:22-30 maxDimension on extent = max - min */
static inline int vlo_max_dim(v3 mn, v3 mx) {
  v3 e = v3_sub(mx, mn);
  int result = 0;
  if (e.y > e.x) {
    result = 1;
    if (e.z > e.y) result = 2;
  } else if (e.z > e.x) result = 2;
  return result;
}

static vlo_node* vlo_build(const vlo_tri* tris, uint32_t* prims, uint32_t n, uint32_t leafSize,
                           uint32_t* nNodes_out) {
  vlo_build_entry todo[128];
  uint32_t stackptr = 0;
  const uint32_t Untouched = 0xffffffffu, TouchedTwice = 0xfffffffdu;
  uint32_t nNodes = 0;
  vlo_node* nodes = (vlo_node*)malloc(sizeof(vlo_node) * (size_t)(n ? n * 2 : 1));
  todo[stackptr].start = 0; todo[stackptr].end = n; todo[stackptr].parent = 0xfffffffcu;
  stackptr++;
  while (stackptr > 0) {
    vlo_build_entry b = todo[--stackptr];
    uint32_t start = b.start, end = b.end, nPrims = end - start;
    vlo_node node;
    nNodes++;
    node.start = start; node.nPrims = nPrims; node.rightOffset = Untouched;
    /* BVH.cpp:174-181 */
    v3 bbmin = tris[prims[start]].bmin, bbmax = tris[prims[start]].bmax;
    v3 bcmin = tris[prims[start]].centroid, bcmax = bcmin;
    for (uint32_t p = start + 1; p < end; ++p) {
      const vlo_tri* T = &tris[prims[p]];
      bbmin = v3_min(bbmin, T->bmin); bbmax = v3_max(bbmax, T->bmax);
      bcmin = v3_min(bcmin, T->centroid); bcmax = v3_max(bcmax, T->centroid);
    }
    node.bmin = bbmin; node.bmax = bbmax;
    if (nPrims <= leafSize) node.rightOffset = 0;
    nodes[nNodes - 1] = node;
    /* BVH.cpp:194-203 */
    if (b.parent != 0xfffffffcu) {
      nodes[b.parent].rightOffset--;
      if (nodes[b.parent].rightOffset == TouchedTwice)
        nodes[b.parent].rightOffset = nNodes - 1 - b.parent;
    }
    if (node.rightOffset == 0) continue;
    /* BVH.cpp:209-227 */
    int split_dim = vlo_max_dim(bcmin, bcmax);
    float split_coord = .5f * (v3_get(bcmin, split_dim) + v3_get(bcmax, split_dim));
    uint32_t mid = start;
    for (uint32_t i = start; i < end; ++i) {
      if (v3_get(tris[prims[i]].centroid, split_dim) < split_coord) {
        uint32_t tmp = prims[i]; prims[i] = prims[mid]; prims[mid] = tmp;
        ++mid;
      }
    }
    if (mid == start || mid == end) mid = start + (end - start) / 2;
    todo[stackptr].start = mid; todo[stackptr].end = end; todo[stackptr].parent = nNodes - 1; stackptr++;
    todo[stackptr].start = start; todo[stackptr].end = mid; todo[stackptr].parent = nNodes - 1; stackptr++;
  }
  *nNodes_out = nNodes;
  return nodes;
}

/* BBox.cpp:52-100 SSE slab test, lanes x,y,z */
static inline int vlo_box_hit(const vlo_node* nd, v3 o, v3 inv_d, float* tnear, float* tfar) {
  const float pinf = INFINITY, ninf = -INFINITY;
  float lmax[3], lmin[3];
  for (int k = 0; k < 3; ++k) {
    float l1 = (v3_get(nd->bmin, k) - v3_get(o, k)) * v3_get(inv_d, k);
    float l2 = (v3_get(nd->bmax, k) - v3_get(o, k)) * v3_get(inv_d, k);
    float f1a = sse_min(l1, pinf), f2a = sse_min(l2, pinf);
    float f1b = sse_max(l1, ninf), f2b = sse_max(l2, ninf);
    lmax[k] = sse_max(f1a, f2a);
    lmin[k] = sse_min(f1b, f2b);
  }
  float mx = sse_min(sse_min(lmax[0], lmax[1]), lmax[2]);
  float mn = sse_max(sse_max(lmin[0], lmin[1]), lmin[2]);
  *tnear = mn; *tfar = mx;
  return (mx >= 0.0f) & (mx >= mn);
}

/* BVH.cpp:19-110 closest-hit traversal.  Returns the winning primitive index or -1. */
static int vlo_closest(const vlo_node* nodes, const vlo_tri* tris, const uint32_t* prims,
                       v3 o, v3 d, v3 inv_d, unsigned flags, float* t_best) {
  float best_t = 999999999.f;
  int best = -1;
  struct { uint32_t i; float mint; } todo[64];
  int32_t sp = 0;
  float bb[4];
  todo[0].i = 0; todo[0].mint = -9999999.f;
  while (sp >= 0) {
    uint32_t ni = todo[sp].i;
    float near = todo[sp].mint;
    sp--;
    const vlo_node* node = &nodes[ni];
    if (near > best_t) continue;
    if (node->rightOffset == 0) {
      for (uint32_t k = 0; k < node->nPrims; ++k) {
        uint32_t id = prims[node->start + k];
        float t;
        if (vlo_tri_hit(&tris[id], o, d, &t)) {
          if (t < best_t || ((flags & VLO_MIN_ID_TIES) && t == best_t && (int)id < best)) {
            best_t = t; best = (int)id;
          }
        }
      }
    } else {
      int h0 = vlo_box_hit(&nodes[ni + 1], o, inv_d, bb, bb + 1);
      int h1 = vlo_box_hit(&nodes[ni + node->rightOffset], o, inv_d, bb + 2, bb + 3);
      if (h0 && h1) {
        uint32_t closer = ni + 1, other = ni + node->rightOffset;
        if (bb[2] < bb[0]) {
          float s;
          s = bb[0]; bb[0] = bb[2]; bb[2] = s;
          s = bb[1]; bb[1] = bb[3]; bb[3] = s;
          uint32_t u = closer; closer = other; other = u;
        }
        ++sp; todo[sp].i = other; todo[sp].mint = bb[2];
        ++sp; todo[sp].i = closer; todo[sp].mint = bb[0];
      } else if (h0) {
        ++sp; todo[sp].i = ni + 1; todo[sp].mint = bb[0];
      } else if (h1) {
        ++sp; todo[sp].i = ni + node->rightOffset; todo[sp].mint = bb[2];
      }
    }
  }
  *t_best = best_t;
  return best;
}

static int vlo_closest_brute(const vlo_tri* tris, uint32_t n, v3 o, v3 d, float* t_best) {
  float best_t = 999999999.f;
  int best = -1;
  for (uint32_t id = 0; id < n; ++id) {
    float t;
    if (vlo_tri_hit(&tris[id], o, d, &t) && t < best_t) { best_t = t; best = (int)id; }
  }
  *t_best = best_t;
  return best;
}

/*
 * vlo_trace: restates trace()/ctrace (RayTracer.cpp:19-124) with the same argument
 * list, plus a nullable per-ray triangle-id output and a flags word.
 * Outputs are written only for hits (the caller zero-fills; RayTracer.cpp:72-90).
 * stats (nullable): [0] = number of BVH nodes built.
 */
int vlo_trace(const float* rays, const float* origin_in, const float* verts, const int* faces,
              const int* colors, const float* rem, int n_rays, int n_verts, int n_faces, int height,
              float* endpoints, int* endcolors, float* range, float* endrem, int* tri_id,
              unsigned flags, int* stats) {
  (void)n_verts;
  if (n_faces <= 0 || height <= 0) return 0;
  vlo_tri* tris = (vlo_tri*)malloc(sizeof(vlo_tri) * (size_t)n_faces);
  uint32_t* prims = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_faces);
  for (int i = 0; i < n_faces; ++i) {
    vlo_tri* T = &tris[i];
    float r[3];
    int idx = faces[i * 3 + 0] * 3;
    T->v0 = v3_make(verts[idx], verts[idx + 1], verts[idx + 2]);
    /* colours go int -> float (Vector3 ctor) -> int on output (RayTracer.cpp:36,82-84) */
    T->c0[0] = (int)(float)colors[idx]; T->c0[1] = (int)(float)colors[idx + 1]; T->c0[2] = (int)(float)colors[idx + 2];
    r[0] = rem[idx / 3];
    idx = faces[i * 3 + 1] * 3;
    T->v1 = v3_make(verts[idx], verts[idx + 1], verts[idx + 2]);
    r[1] = rem[idx / 3];
    idx = faces[i * 3 + 2] * 3;
    T->v2 = v3_make(verts[idx], verts[idx + 1], verts[idx + 2]);
    r[2] = rem[idx / 3];
    T->rem = (r[0] + r[1] + r[2]) / 3;
    T->bmin = v3_min(T->v0, v3_min(T->v1, T->v2));
    T->bmax = v3_max(T->v0, v3_max(T->v1, T->v2));
    v3 c = v3_add(v3_add(T->v0, T->v1), T->v2);
    T->centroid = v3_make(c.x / 3.0f, c.y / 3.0f, c.z / 3.0f);
    prims[i] = (uint32_t)i;
  }
  vlo_node* nodes = NULL;
  uint32_t nNodes = 0;
  if (!(flags & VLO_BRUTE_FORCE)) nodes = vlo_build(tris, prims, (uint32_t)n_faces, 4, &nNodes);
  if (stats) stats[0] = (int)nNodes;

  const long width = n_rays / height;
  const v3 o = v3_make(origin_in[0], origin_in[1], origin_in[2]);
  /* RayTracer.cpp:62-92 */
#pragma omp parallel for schedule(static)
  for (long i = 0; i < width; ++i) {
    for (int j = 0; j < height; ++j) {
      size_t index = 3 * (size_t)(width * j + i);
      v3 d = vlo_normalize(v3_make(rays[index], rays[index + 1], rays[index + 2]), flags);
      v3 inv_d = v3_make(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); /* Ray.h:11-12 */
      float t;
      int id = (flags & VLO_BRUTE_FORCE) ? vlo_closest_brute(tris, (uint32_t)n_faces, o, d, &t)
                                         : vlo_closest(nodes, tris, prims, o, d, inv_d, flags, &t);
      if (id >= 0) {
        const vlo_tri* T = &tris[id];
        /* BVH.cpp:106-107 hit = o + d*t */
        endpoints[index + 0] = o.x + d.x * t;
        endpoints[index + 1] = o.y + d.y * t;
        endpoints[index + 2] = o.z + d.z * t;
        endcolors[index + 0] = T->c0[0];
        endcolors[index + 1] = T->c0[1];
        endcolors[index + 2] = T->c0[2];
        endrem[index / 3] = T->rem;
        range[width * j + i] = t;
      }
      if (tri_id) tri_id[width * j + i] = id;
    }
  }
  free(nodes); free(prims); free(tris);
  return 0;
}

/* ------------------------------------------------------------------ */
/* spherical range-image projection                                     */
/* ------------------------------------------------------------------ */
/*
 * vlo_project: restates LaserScan.do_range_projection_new(method="depth")
 * (auxiliary/laserscan.py:294-391) + SemLaserScan.do_label_projection_new (:672-676).
 *
 * points float64[n,3] (the reference holds float64 after the pose round trip),
 * remissions float32[n], labels uint32[n].  remove != 0 applies the vertical
 * FOV filter (:337-345).  All angle math float64, sequential per-point loop,
 * comparison of the float64 depth against the float32 image value (:376-378).
 *
 * Outputs: range_image f32[H*W] (0 empty), index i32[H*W] (-1 empty; indices into
 * the KEPT point list like the reference after remove_points), proj_label
 * i32[H*W] (0 empty), proj_rem f32[H*W] (-1 empty); keep u8[n] (nullable)
 * marks which input points survive the depth!=0 and FOV filters.
 * Returns the number of kept points.
 */
/* vlo_project_snap: the same with the `beam_angles` step of :321-327 -- each point's pitch is replaced by the
 * entry of beam_angles[0..n_beam_angles) nearest to it (np.abs(pitch - beam_angles).argmin(): first minimum;
 * the reference compares the pitch in radians with the list as given).  n_beam_angles == 0: no snapping
 * (`if self.beam_angles:` is false for None and for an empty list). */
long vlo_project_snap(const double* points, const float* remissions, const uint32_t* labels, long n,
                      double fov_up_deg, double fov_down_deg, int H, int W, int remove,
                      const double* beam_angles, int n_beam_angles,
                      float* range_image, int32_t* index, int32_t* proj_label, float* proj_rem,
                      uint8_t* keep) {
  const double pi = 3.141592653589793; /* np.pi */
  double fov_up = fov_up_deg / 180.0 * pi;
  double fov_down = fov_down_deg / 180.0 * pi;
  double fov = fabs(fov_down) + fabs(fov_up);
  for (long p = 0; p < (long)H * W; ++p) {
    range_image[p] = 0.0f; index[p] = -1; proj_label[p] = 0; proj_rem[p] = -1.0f;
  }
  long kept = 0;
  for (long i = 0; i < n; ++i) {
    double x = points[3 * i], y = points[3 * i + 1], z = points[3 * i + 2];
    double depth = sqrt((x * x + y * y) + z * z); /* np.linalg.norm(points, 2, axis=1) */
    if (keep) keep[i] = 0;
    if (depth == 0) continue; /* :307-309 */
    double yaw = -atan2(y, x);
    double pitch = asin(z / depth);
    if (n_beam_angles > 0) { /* :321-327 */
      int best = 0;
      double best_d = fabs(pitch - beam_angles[0]);
      for (int k = 1; k < n_beam_angles; ++k) {
        double d = fabs(pitch - beam_angles[k]);
        if (d < best_d) { best_d = d; best = k; }
      }
      pitch = beam_angles[best];
    }
    double proj_x = 0.5 * (yaw / pi + 1.0);
    double proj_y = 1.0 - (pitch + fabs(fov_down)) / fov;
    if (remove && !(proj_y >= 0 && proj_y <= 1)) continue; /* :337-345 */
    proj_x *= W; proj_y *= H;
    double fx = floor(proj_x), fy = floor(proj_y);
    fx = fmin((double)(W - 1), fx); fx = fmax(0.0, fx);
    fy = fmin((double)(H - 1), fy); fy = fmax(0.0, fy);
    int px = (int)fx, py = (int)fy;
    long pix = (long)py * W + px;
    if (keep) keep[i] = 1;
    /* :373-382 */
    if (depth < (double)range_image[pix] || index[pix] == -1) {
      range_image[pix] = (float)depth;
      index[pix] = (int32_t)kept;
      proj_label[pix] = (int32_t)labels[i];
      proj_rem[pix] = remissions[i];
    }
    kept++;
  }
  return kept;
}

long vlo_project(const double* points, const float* remissions, const uint32_t* labels, long n,
                 double fov_up_deg, double fov_down_deg, int H, int W, int remove,
                 float* range_image, int32_t* index, int32_t* proj_label, float* proj_rem,
                 uint8_t* keep) {
  return vlo_project_snap(points, remissions, labels, n, fov_up_deg, fov_down_deg, H, W, remove, 0, 0, range_image,
                          index, proj_label, proj_rem, keep);
}

/* ------------------------------------------------------------------ */
/* TSDF integration                                                     */
/* ------------------------------------------------------------------ */
/*
 * vlo_tsdf_integrate: restates the CUDA `integrate` kernel string
 * (auxiliary/fusion_lidar.py:70-229, class-aware branch `merge == true`) for every
 * voxel_idx in [0, dx*dy*dz) -- the reference's `voxel_idx > N` guard (:92) lets
 * idx == N through, a one-element out-of-bounds access this restatement does not
 * reproduce.  The launch geometry (:233-250, :267-287) only enumerates voxel_idx.
 *
 * other_params mirrors the float32 array built at :277-280:
 *   [1] voxel_size [2] im_h [3] im_w [4] trunc_margin [5] obs_weight [6] fov_up [7] fov_down
 * Expressions keep the reference's float/double mixing; fmaf() marks the places
 * where nvcc's default -fmad=true contracts a float multiply-add (the reference
 * kernel is JIT-compiled by pycuda with nvcc defaults).  Compile with
 * -ffp-contract=off so nothing else is fused.
 */
#define VLO_PI 3.14159265358979323846
void vlo_tsdf_integrate(float* tsdf_vol, float* weight_vol, float* color_vol, float* rem_vol,
                        const float* vol_dim, const float* vol_origin, const float* other_params,
                        const float* color_im, const float* depth_im, const float* rem_im,
                        long long* counters) {
  const int vol_dim_x = (int)vol_dim[0], vol_dim_y = (int)vol_dim[1], vol_dim_z = (int)vol_dim[2];
  const long long n_vox = (long long)vol_dim_x * vol_dim_y * vol_dim_z;
  long long n_vis = 0, n_written = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_vis, n_written)
  for (long long vi = 0; vi < n_vox; ++vi) {
    int voxel_idx = (int)vi;
    /* :96-98 float-precision index decode (can yield x+1, y=-1 for idx > 2^24) */
    float voxel_x = floorf(((float)voxel_idx) / ((float)(vol_dim_y * vol_dim_z)));
    float voxel_y = floorf(((float)(voxel_idx - ((int)voxel_x) * vol_dim_y * vol_dim_z)) / ((float)vol_dim_z));
    float voxel_z = (float)(voxel_idx - ((int)voxel_x) * vol_dim_y * vol_dim_z - ((int)voxel_y) * vol_dim_z);
    float voxel_size = other_params[1];
    float pt_x = fmaf(voxel_x, voxel_size, vol_origin[0]);
    float pt_y = fmaf(voxel_y, voxel_size, vol_origin[1]);
    float pt_z = fmaf(voxel_z, voxel_size, vol_origin[2]);
    float cam_pt_x = pt_x, cam_pt_z = pt_z, cam_pt_y = pt_y;
    int im_h = (int)other_params[2];
    int im_w = (int)other_params[3];
    float fov_up = other_params[6] * VLO_PI / 180.0;
    float fov_down = other_params[7] * VLO_PI / 180.0;
    float fov = fabsf(fov_up) + fabsf(fov_down);
    /* norm3df: correctly rounded sqrt(x^2+y^2+z^2) on the CPU side */
    float depth = (float)sqrt((double)cam_pt_x * cam_pt_x + (double)cam_pt_y * cam_pt_y + (double)cam_pt_z * cam_pt_z);
    float yaw = -atan2f(cam_pt_y, cam_pt_x);
    float pitch = asinf(cam_pt_z / depth);
    if (pitch > fov_up || pitch < fov_down) continue;
    float proj_x = 0.5 * (yaw / VLO_PI + 1.0);
    float proj_y = 1.0 - (pitch + fabsf(fov_down)) / fov;
    proj_x *= im_w;
    proj_y *= im_h;
    int proj_x_cl = (int)floor(proj_x);
    proj_x_cl = im_w - 1 < proj_x_cl ? im_w - 1 : proj_x_cl;
    proj_x_cl = 0 > proj_x_cl ? 0 : proj_x_cl;
    int proj_y_cl = (int)floor(proj_y);
    proj_y_cl = im_h - 1 < proj_y_cl ? im_h - 1 : proj_y_cl;
    proj_y_cl = 0 > proj_y_cl ? 0 : proj_y_cl;
    int pixel_x = proj_x_cl, pixel_y = proj_y_cl;
    if (pixel_x < 0 || pixel_x >= im_w || pixel_y < 0 || pixel_y >= im_h) continue;
    float depth_value = depth_im[pixel_y * im_w + pixel_x];
    if (depth_value == 0) continue;
    float trunc_margin = other_params[4];
    float depth_diff = depth_value - depth;
    if (depth_diff < -trunc_margin) continue;
    n_vis++;
    float dist = fminf(1.0f, depth_diff / trunc_margin);
    float dist_old = weight_vol[voxel_idx]; /* sic: :197 reads the weight volume */
    float old_color = color_vol[voxel_idx];
    float new_color = color_im[pixel_y * im_w + pixel_x];
    if (old_color == new_color) {
      float w_old = weight_vol[voxel_idx];
      float obs_weight = other_params[5];
      float w_new = w_old + obs_weight;
      weight_vol[voxel_idx] = w_new;
      tsdf_vol[voxel_idx] = fmaf(tsdf_vol[voxel_idx], w_old, dist) / w_new;
      float old_rem = rem_vol[voxel_idx];
      float new_rem = rem_im[pixel_y * im_w + pixel_x];
      rem_vol[voxel_idx] = fmaf(old_rem, w_old, new_rem) / w_new;
      n_written++;
    } else if (dist < dist_old) {
      tsdf_vol[voxel_idx] = dist;
      float new_b = floorf(new_color / (256 * 256));
      float new_g = floorf((new_color - new_b * 256 * 256) / 256);
      float new_r = new_color - new_b * 256 * 256 - new_g * 256;
      color_vol[voxel_idx] = new_b * 256 * 256 + new_g * 256 + new_r;
      rem_vol[voxel_idx] = rem_im[pixel_y * im_w + pixel_x];
      n_written++;
    }
  }
  if (counters) { counters[0] = n_vis; counters[1] = n_written; }
}

/* ------------------------------------------------------------------ */
/* iso-surface vertex attribute lookup (TSDFVolume.get_mesh, fusion_lidar.py:408-423) */
/* ------------------------------------------------------------------ */
/*
 * verts_vox float32[n,3] in voxel coordinates (marching-cubes output before the
 * world transform).  np.round = round-half-even (rint).  colours are
 * (r,g,b) = (c - b*65536 - g*256, g, b) -> astype(uint8) (wraps mod 256).
 */
void vlo_mesh_attributes(const float* verts_vox, long n, const float* color_vol, const float* rem_vol,
                         int dx, int dy, int dz, float voxel_size, const float* vol_origin,
                         float* verts_world, uint8_t* colors, float* rem) {
  (void)dx;
  for (long i = 0; i < n; ++i) {
    long ix = (long)rint((double)verts_vox[3 * i]);
    long iy = (long)rint((double)verts_vox[3 * i + 1]);
    long iz = (long)rint((double)verts_vox[3 * i + 2]);
    size_t vi = ((size_t)ix * dy + iy) * dz + iz;
    for (int k = 0; k < 3; ++k) /* :412 verts * voxel_size + origin: float32*py-float + float32 */
      verts_world[3 * i + k] = verts_vox[3 * i + k] * voxel_size + vol_origin[k];
    float rgb = color_vol[vi];
    float b = floorf(rgb / (256 * 256));
    float g = floorf((rgb - b * 256 * 256) / 256);
    float r = rgb - b * 256 * 256 - g * 256;
    colors[3 * i + 0] = (uint8_t)(long)floorf(r);
    colors[3 * i + 1] = (uint8_t)(long)floorf(g);
    colors[3 * i + 2] = (uint8_t)(long)floorf(b);
    rem[i] = rem_vol[vi];
  }
}

/* ------------------------------------------------------------------ */
/* iso-surface extraction (restates lidar_transfer_b200/csrc/vl_mesh.cu; replaces get_mesh,   */
/* auxiliary/fusion_lidar.py:403-424, whose marching cubes lives in un-vendored scikit-image  */
/* -- parity of the topology against skimage is UNPINNED)                     */
/* ------------------------------------------------------------------ */
#include "../lidar_transfer_b200/csrc/vl_mc_table.inc"
static const signed char vlo_tri_table[256][15] = VL_MC_TRI_TABLE;
static const unsigned char vlo_tri_count[256] = VL_MC_TRI_COUNT;
static const unsigned char vlo_edge_corners[12][2] = VL_MC_EDGE_CORNERS;

/*
 * Cubes are visited in voxel-index order (x slowest, z fastest), triangles in table order; the
 * output is a triangle soup (3 vertices per triangle).  Vertex = low corner + t along the edge
 * axis with t = (level - va) / (vb - va) in float32; world = vert * voxel_size + origin (:412);
 * label / remission from the nearest voxel (np.round, :409-415); colours split as :417-423.
 * Returns the number of triangles; writes at most `capacity` of them (capacity 0 = count only).
 */
long long vlo_mesh_extract(const float* tsdf, const float* color_vol, const float* rem_vol, int dx, int dy, int dz,
                           float level, float voxel_size, const float* vol_origin, long long capacity,
                           float* verts, int* faces, uint8_t* colors, float* rem_out) {
  long long n_tris = 0;
  const long long yz = (long long)dy * dz;
  for (int x = 0; x + 1 < dx; ++x)
    for (int y = 0; y + 1 < dy; ++y)
      for (int z = 0; z + 1 < dz; ++z) {
        const long long vi = (long long)x * yz + (long long)y * dz + z;
        float v[8];
        int mask = 0;
        for (int c = 0; c < 8; ++c) {
          v[c] = tsdf[vi + (c & 1) * yz + ((c >> 1) & 1) * dz + ((c >> 2) & 1)];
          if (v[c] < level) mask |= 1 << c;
        }
        const int cnt = vlo_tri_count[mask];
        for (int t = 0; t < cnt; ++t, ++n_tris) {
          if (n_tris >= capacity) continue;
          for (int k = 0; k < 3; ++k) {
            const int e = vlo_tri_table[mask][3 * t + k];
            const int ca = vlo_edge_corners[e][0], cb = vlo_edge_corners[e][1];
            const float tt = (level - v[ca]) / (v[cb] - v[ca]);
            float pv[3] = {(float)(x + (ca & 1)), (float)(y + ((ca >> 1) & 1)), (float)(z + ((ca >> 2) & 1))};
            const int axis = (ca ^ cb) == 1 ? 0 : ((ca ^ cb) == 2 ? 1 : 2);
            pv[axis] = pv[axis] + tt;
            long ix = (long)rint((double)pv[0]), iy = (long)rint((double)pv[1]), iz = (long)rint((double)pv[2]);
            if (ix > dx - 1) ix = dx - 1;
            if (iy > dy - 1) iy = dy - 1;
            if (iz > dz - 1) iz = dz - 1;
            const long long ni = ((long long)ix * dy + iy) * dz + iz;
            const float rgb = color_vol[ni];
            const float b = floorf(rgb / (256 * 256));
            const float g = floorf((rgb - b * 256 * 256) / 256);
            const float r = rgb - b * 256 * 256 - g * 256;
            const long long vtx = 3 * n_tris + k;
            colors[3 * vtx + 0] = (uint8_t)((long long)floorf(r) & 255);
            colors[3 * vtx + 1] = (uint8_t)((long long)floorf(g) & 255);
            colors[3 * vtx + 2] = (uint8_t)((long long)floorf(b) & 255);
            rem_out[vtx] = rem_vol[ni];
            for (int a = 0; a < 3; ++a) verts[3 * vtx + a] = pv[a] * voxel_size + vol_origin[a];
            faces[vtx] = (int)vtx;
          }
        }
      }
  return n_tris;
}

/* The normalisation alone (Vector3.h:73-89) for n rays: what the product's vl_normalize_rays must reproduce bit for
 * bit in SSE mode. */
void vlo_normalize_rays(const float* rays, long n, unsigned flags, float* out) {
  for (long i = 0; i < n; ++i) {
    v3 d = vlo_normalize(v3_make(rays[3 * i], rays[3 * i + 1], rays[3 * i + 2]), flags);
    out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
  }
}

/* One ray against one triangle (Ray.h:11-12 normalise + Triangle.h:27-50): returns 1 and *t on a hit.  Used by the
 * parity tests to PROVE that two triangles the device and the reference disagree on are an exact-t tie
 * (BVH.cpp:59 keeps the first strictly smaller t in traversal order). */
int vlo_ray_triangle(const float* ray, const float* origin, const float* v0, const float* v1, const float* v2,
                     unsigned flags, float* t_out) {
  vlo_tri T;
  memset(&T, 0, sizeof(T));
  T.v0 = v3_make(v0[0], v0[1], v0[2]); T.v1 = v3_make(v1[0], v1[1], v1[2]); T.v2 = v3_make(v2[0], v2[1], v2[2]);
  v3 d = vlo_normalize(v3_make(ray[0], ray[1], ray[2]), flags);
  return vlo_tri_hit(&T, v3_make(origin[0], origin[1], origin[2]), d, t_out);
}

int vlo_abi_version(void) { return 1; }
