// Host entry points of the device LBVH build (drt_mesh.cu).
#pragma once
#include <string>
#include <cuda_runtime.h>
#include "../../include/drt.h"

namespace drt {
struct MeshBuffers {
  float4* nodes = nullptr;   // 4 float4 per internal node
  void* tris_f64 = nullptr;  // MeshTri<double>[n_tris]
  void* tris_f32 = nullptr;  // MeshTri<float>[n_tris]
  unsigned short* mat_ids = nullptr;  // per-triangle index into the mesh's material table, or nullptr (one material)
  int n_tris = 0;
  float build_ms = 0;        // device time of the build (CUDA events)
};
// Uploads the mesh, builds the LBVH on the current device.  Returns a drt_status.
int buildMesh(const drt_mesh* mesh, MeshBuffers* out, std::string& err);
void freeMesh(MeshBuffers* m);
}  // namespace drt
