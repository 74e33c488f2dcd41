"""Drop-in for the reference's `auxiliary.raytracing.ray_mesh_intersection` (auxiliary/raytracing.py:17-21).

The reference module is imported by fusion_lidar.py:8 but never called; its two implementations disagree with
each other (CUDA: FIRST hit in face order, :150; CPU: LAST valid hit, :28-41) and neither returns the closest
hit.  This shim keeps the name, arguments and return shapes -- (endpoints[R,3] float32, colors[R,3] float32,
zero rows for rays that miss) -- and returns the CLOSEST hit through the BVH engine; the deviation is
deliberate and documented in DESIGN.md."""
import numpy as np

from .. import engine

GPU_MODE = 1


def ray_mesh_intersection(rays, origin, vertices, vertices_colors, faces, H, W):
  rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 3)
  vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
  faces = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
  vertices_colors = np.asarray(vertices_colors).reshape(-1, 3)
  bvh = engine.Bvh(vertices, faces, np.zeros(vertices.shape, np.int32), np.zeros(vertices.shape[0], np.float32))
  out = engine.trace(bvh, rays, np.ascontiguousarray(origin, np.float32), int(H), zero_misses=True)
  tri = out["tri_id"].cpu().numpy()
  endpoints = out["endpoints"].cpu().numpy().reshape(-1, 3)
  colors = np.zeros(rays.shape, np.float32)
  hit = tri >= 0
  colors[hit] = vertices_colors[faces[tri[hit], 0]].astype(np.float32)  # colour of the triangle's vertex 0
  return endpoints, colors


ray_mesh_intersection_CUDA = ray_mesh_intersection
