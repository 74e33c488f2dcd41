// ref_ids_harness.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Drives the UNMODIFIED reference ray tracer classes (compiled from where they lie,
// -I /root/reference/auxiliary/raytracer, together with the reference's BVH.cpp and
// BBox.cpp) and additionally reports WHICH triangle each ray hit -- the reference's
// own ctrace (RayTracer.cpp:116-124) does not output triangle ids, BASELINE.json's
// north_star asks for them.  The id rides inside the object because BVH::build
// permutes the pointer vector (BVH.cpp:216).
//
// The loop below follows trace() (RayTracer.cpp:19-92) and writes the same outputs.
#include <vector>
#include <cstdio>
#include "BVH.h"
#include "Triangle.h"

struct IdTriangle : public Triangle {
  int id;
  IdTriangle(const Vector3& v0, const Vector3& v1, const Vector3& v2, const Vector3& c0,
             const Vector3& c1, const Vector3& c2, const Vector3& r, int id)
      : Triangle(v0, v1, v2, c0, c1, c2, r), id(id) {}
};

extern "C" void ctrace_ids(float* rays, float* origin_in, float* verts, int* faces, int* colors,
                           float* rem, int n_rays, int n_verts, int n_faces, int height,
                           float* endpoints, int* endcolors, float* range, float* endrem,
                           int* tri_id) {
  (void)n_verts;
  std::vector<Object*> objects;
  objects.reserve(n_faces);
  for (int i = 0; i < n_faces; ++i) {
    Vector3 r(0.0, 0.0, 0.0);
    Vector3 v[3], c[3];
    for (int k = 0; k < 3; ++k) {
      int idx = faces[i * 3 + k] * 3;
      v[k] = Vector3(verts[idx + 0], verts[idx + 1], verts[idx + 2]);
      c[k] = Vector3(colors[idx + 0], colors[idx + 1], colors[idx + 2]);
      r[k] = rem[idx / 3];
    }
    objects.push_back(new IdTriangle(v[0], v[1], v[2], c[0], c[1], c[2], r, i));
  }
  BVH bvh(&objects);
  const unsigned int width = n_rays / height;
  Vector3 origin(origin_in[0], origin_in[1], origin_in[2]);
#pragma omp parallel for
  for (size_t i = 0; i < width; ++i) {
    for (int j = 0; j < height; ++j) {
      size_t index = 3 * (width * j + i);
      Vector3 single_ray(rays[index + 0], rays[index + 1], rays[index + 2]);
      Ray ray(origin, normalize(single_ray));
      IntersectionInfo I;
      bool hit = bvh.getIntersection(ray, &I, false);
      int id = -1;
      if (hit) {
        Vector3 col = I.object->getColor(0);
        endpoints[index + 0] = I.hit[0];
        endpoints[index + 1] = I.hit[1];
        endpoints[index + 2] = I.hit[2];
        endcolors[index + 0] = col.x;
        endcolors[index + 1] = col.y;
        endcolors[index + 2] = col.z;
        endrem[index / 3] = I.object->getRemissions(0);
        range[width * j + i] = I.t;
        id = static_cast<const IdTriangle*>(I.object)->id;
      }
      tri_id[width * j + i] = id;
    }
  }
  for (Object* p : objects) delete p;
}
