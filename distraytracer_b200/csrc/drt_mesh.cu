// Device LBVH build for drt_mesh (see drt_lbvh.cuh for the kernels).
#include <algorithm>
#include <cmath>
#include <vector>

#include "drt_lbvh.cuh"
#include "drt_mesh.h"

namespace drt {

#define MCK(call)                                                                           \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) { err = std::string(#call) + " failed: " + cudaGetErrorString(e_); goto fail; } \
  } while (0)

void freeMesh(MeshBuffers* m) {
  if (!m) return;
  if (m->nodes) cudaFree(m->nodes);
  if (m->tris_f64) cudaFree(m->tris_f64);
  if (m->tris_f32) cudaFree(m->tris_f32);
  if (m->mat_ids) cudaFree(m->mat_ids);
  *m = MeshBuffers();
}

int buildMesh(const drt_mesh* mesh, MeshBuffers* out, std::string& err) {
  const long long nv = mesh->n_vertices, nt = mesh->n_triangles;
  if (nv < 3 || nt < 1 || !mesh->vertices || !mesh->indices) { err = "empty mesh"; return DRT_ERR_INVALID; }
  if (nt > (1ll << 30)) { err = "mesh too large"; return DRT_ERR_UNSUPPORTED; }
  for (long long i = 0; i < 3 * nt; i++)
    if (mesh->indices[i] < 0 || mesh->indices[i] >= nv) { err = "mesh index out of range"; return DRT_ERR_INVALID; }
  if (mesh->n_materials < 0 || mesh->n_materials > 65535) { err = "mesh material table larger than 65535 entries"; return DRT_ERR_UNSUPPORTED; }
  if (mesh->n_materials > 0) {
    if (!mesh->materials || !mesh->material_ids) { err = "mesh material table without ids"; return DRT_ERR_INVALID; }
    for (long long t = 0; t < nt; t++)
      if (mesh->material_ids[t] < 0 || mesh->material_ids[t] >= mesh->n_materials) { err = "mesh material id out of range"; return DRT_ERR_INVALID; }
  }
  const int n = (int)nt;
  // scene bounds of the centroids' support (host: one pass over the vertices)
  float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
  for (long long v = 0; v < nv; v++)
    for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], mesh->vertices[3 * v + a]); hi[a] = std::max(hi[a], mesh->vertices[3 * v + a]); }
  float3 slo = make_float3(lo[0], lo[1], lo[2]);
  const float sinv = 1.0f / std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), std::max(hi[2] - lo[2], 1e-20f));   // cubic Morton cells

  float *d_verts = nullptr, *d_tc = nullptr; int* d_idx = nullptr;
  float4 *tlo = nullptr, *thi = nullptr, *nlo = nullptr, *nhi = nullptr;
  uint64_t *codes = nullptr, *codes2 = nullptr; int *ids = nullptr, *ids2 = nullptr;
  int *left = nullptr, *right = nullptr, *pin = nullptr, *pleaf = nullptr, *visits = nullptr;
  void* tmp = nullptr; size_t tmp_bytes = 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  MeshBuffers mb;
  mb.n_tris = n;
  const int B = 256, G = (n + B - 1) / B;
  MCK(cudaMalloc(&d_verts, sizeof(float) * 3 * nv));
  MCK(cudaMalloc(&d_idx, sizeof(int) * 3 * nt));
  MCK(cudaMemcpy(d_verts, mesh->vertices, sizeof(float) * 3 * nv, cudaMemcpyHostToDevice));
  MCK(cudaMemcpy(d_idx, mesh->indices, sizeof(int) * 3 * nt, cudaMemcpyHostToDevice));
  if (mesh->texcoords) {
    MCK(cudaMalloc(&d_tc, sizeof(float) * 2 * nv));
    MCK(cudaMemcpy(d_tc, mesh->texcoords, sizeof(float) * 2 * nv, cudaMemcpyHostToDevice));
  }
  MCK(cudaMalloc(&tlo, sizeof(float4) * n)); MCK(cudaMalloc(&thi, sizeof(float4) * n));
  MCK(cudaMalloc(&nlo, sizeof(float4) * std::max(1, n - 1))); MCK(cudaMalloc(&nhi, sizeof(float4) * std::max(1, n - 1)));
  MCK(cudaMalloc(&codes, sizeof(uint64_t) * n)); MCK(cudaMalloc(&codes2, sizeof(uint64_t) * n));
  MCK(cudaMalloc(&ids, sizeof(int) * n)); MCK(cudaMalloc(&ids2, sizeof(int) * n));
  MCK(cudaMalloc(&left, sizeof(int) * std::max(1, n - 1))); MCK(cudaMalloc(&right, sizeof(int) * std::max(1, n - 1)));
  MCK(cudaMalloc(&pin, sizeof(int) * std::max(1, n - 1))); MCK(cudaMalloc(&pleaf, sizeof(int) * n));
  MCK(cudaMalloc(&visits, sizeof(int) * std::max(1, n - 1)));
  MCK(cudaMalloc(&mb.nodes, sizeof(float4) * 8 * std::max(1, n - 1)));
  MCK(cudaMalloc(&mb.tris_f64, sizeof(MeshTri<double>) * n));
  MCK(cudaMalloc(&mb.tris_f32, sizeof(MeshTri<float>) * n));
  if (mesh->n_materials > 0) {
    std::vector<unsigned short> ids((size_t)n);
    for (int t = 0; t < n; t++) ids[t] = (unsigned short)mesh->material_ids[t];
    MCK(cudaMalloc(&mb.mat_ids, sizeof(unsigned short) * n));
    MCK(cudaMemcpy(mb.mat_ids, ids.data(), sizeof(unsigned short) * n, cudaMemcpyHostToDevice));
  }
  MCK(cudaEventCreate(&e0)); MCK(cudaEventCreate(&e1));
  MCK(cudaEventRecord(e0));
  lbvh_tri_setup<<<G, B>>>(n, d_verts, d_idx, slo, sinv, tlo, thi, codes, ids);
  lbvh_tri_records<double><<<G, B>>>(n, d_verts, d_idx, d_tc, (MeshTri<double>*)mb.tris_f64);
  lbvh_tri_records<float><<<G, B>>>(n, d_verts, d_idx, d_tc, (MeshTri<float>*)mb.tris_f32);
  if (n > 1) {
    MCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, codes, codes2, ids, ids2, n, 0, 63));
    MCK(cudaMalloc(&tmp, tmp_bytes));
    MCK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, codes, codes2, ids, ids2, n, 0, 63));
    MCK(cudaMemset(visits, 0, sizeof(int) * (n - 1)));
    lbvh_karras<<<(n - 1 + B - 1) / B, B>>>(n, codes2, left, right, pin, pleaf);
    lbvh_refit<<<G, B>>>(n, ids2, tlo, thi, left, right, pin, pleaf, nlo, nhi, visits);
    lbvh_pack4<<<(n - 1 + B - 1) / B, B>>>(n, ids2, tlo, thi, left, right, pin, nlo, nhi, mb.nodes);
  } else {
    // single triangle: one node
    float4 h_lo, h_hi;
    MCK(cudaMemcpy(&h_lo, tlo, sizeof(float4), cudaMemcpyDeviceToHost));
    MCK(cudaMemcpy(&h_hi, thi, sizeof(float4), cudaMemcpyDeviceToHost));
    float c[3], h[3];
    const float l3[3] = {h_lo.x, h_lo.y, h_lo.z}, u3[3] = {h_hi.x, h_hi.y, h_hi.z};
    for (int a = 0; a < 3; a++) {
      c[a] = (float)(0.5 * ((double)l3[a] + (double)u3[a]));
      const double hd = std::max((double)u3[a] - (double)c[a], (double)c[a] - (double)l3[a]);
      float hf = (float)hd;
      if ((double)hf < hd) hf = nextafterf(hf, INFINITY);
      h[a] = nextafterf(hf, INFINITY);
    }
    const float E = -1e30f;
    const int leaf = -1, none = (int)0x80000000;
    float4 nd[8];
    // slots 0 and 1 both hold the triangle (only slots 2 and 3 may be empty); a second test of it changes nothing
    nd[0] = make_float4(c[0], c[0], c[1], c[1]); nd[1] = make_float4(c[2], c[2], h[0], h[0]); nd[2] = make_float4(h[1], h[1], h[2], h[2]);
    nd[3] = make_float4(0.f, 0.f, 0.f, 0.f); nd[4] = make_float4(0.f, 0.f, E, E); nd[5] = make_float4(E, E, E, E);
    memcpy(&nd[6].x, &leaf, 4); memcpy(&nd[6].y, &leaf, 4); memcpy(&nd[6].z, &none, 4); memcpy(&nd[6].w, &none, 4);
    nd[7] = make_float4(0.f, 0.f, 0.f, 0.f);
    MCK(cudaMemcpy(mb.nodes, nd, sizeof(nd), cudaMemcpyHostToDevice));
  }
  MCK(cudaEventRecord(e1));
  MCK(cudaDeviceSynchronize());
  MCK(cudaGetLastError());
  cudaEventElapsedTime(&mb.build_ms, e0, e1);
  *out = mb;
  {
    void* frees[] = {d_verts, d_tc, d_idx, tlo, thi, nlo, nhi, codes, codes2, ids, ids2, left, right, pin, pleaf, visits, tmp};
    for (void* p : frees) if (p) cudaFree(p);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  return DRT_OK;
fail:
  {
    void* frees[] = {d_verts, d_tc, d_idx, tlo, thi, nlo, nhi, codes, codes2, ids, ids2, left, right, pin, pleaf, visits, tmp};
    for (void* p : frees) if (p) cudaFree(p);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    freeMesh(&mb);
  }
  return DRT_ERR_CUDA;
}

}  // namespace drt
