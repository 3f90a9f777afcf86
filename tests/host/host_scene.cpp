// Exercises the host-side mirror (distraytracer_b200/host/drt_host.h) the way the reference's
// main() drives its renderer: a build*() function fills shapes/lights, then renderImage().
// The scenes restate buildSceneHW4 (scene.h:4451-4477) and buildSceneReflectance
// (scene.h:3668-3694) through the mirrored constructors.
//   host_scene dump   <scene> <out.bin>  : flattened drt_prim / drt_light arrays (no GPU needed)
//   host_scene render <scene> <out.ppm>  : renderImage() on the GPU(s)
#include <cstdlib>
#include <cstring>
#include <fstream>
#include "../../distraytracer_b200/host/drt_host.h"

using namespace drt::host;
using std::make_shared;
using std::shared_ptr;

static void buildSceneHW4(float) {
  Globals& g = globals();
  g.shapes.clear(); g.lights.clear();
  VEC3 c0(-3.5, 0, -10); float r0 = 3; VEC3 rgb0(1, 0.25, 0.25);
  VEC3 c1(3.5, 0, -10); float r1 = 3; VEC3 rgb1(0.25, 0.25, 1);
  VEC3 c2(0, -1000, -10); float r2 = 997; VEC3 rgb2(0.5, 0.5, 0.5);
  g.shapes.push_back(make_shared<Sphere>(c0, r0, rgb0));
  g.shapes.push_back(make_shared<Sphere>(c1, r1, rgb1));
  g.shapes.push_back(make_shared<Sphere>(c2, r2, rgb2));
  g.lights.push_back(make_shared<pointLight>(VEC3(10, 3, -5), VEC3(1, 1, 1)));
  g.lights.push_back(make_shared<pointLight>(VEC3(-10, 3, -7.5), VEC3(0.5, 0, 0)));
}

static void buildSceneReflectance(float framef) {
  int frame = (int)framef;
  Globals& g = globals();
  g.shapes.clear(); g.lights.clear();
  g.shapes.push_back(make_shared<Sphere>(VEC3(3, 0.5, -4), 1, VEC3(0.5, 0.5, 0.5)));
  auto marble = make_shared<Sphere>(VEC3(3, 0.5, -1.5), 1, VEC3(0.5, 0.5, 0.5), "marble", false, "oren-nayar");
  marble->reflect_params.roughness = sqrt(0.2);
  g.shapes.push_back(marble);
  auto metal = make_shared<Sphere>(VEC3(3, 0.5, 1), 1, VEC3(0.5, 0.5, 0.5), "aluminum", false, "cook-torrance");
  metal->reflect_params.roughness = sqrt(0.2); metal->reflect_params.refr = VEC2(0.958, 6.69);
  g.shapes.push_back(metal);
  auto glossy = make_shared<Sphere>(VEC3(3, 0.5, 3.5), 1, VEC3(0.5, 0.5, 0.5), "aluminum", false, "cook-torrance");
  glossy->reflect_params.roughness = sqrt(0.2); glossy->reflect_params.refr = VEC2(0.958, 6.69); glossy->reflect_params.glossy = true;
  g.shapes.push_back(glossy);
  g.shapes.push_back(make_shared<Sphere>(VEC3(-7, 0.5, 4), 3, VEC3(1, 0, 0)));
  VEC3 light_center = VEC3(-6, 5, -10) + VEC3(0, 0, 20) * (float)frame / 150;
  g.lights.push_back(make_shared<pointLight>(light_center, VEC3(1, 1, 1)));
}

// The mocap part of buildSceneChkpt2 (scene.h:3557-3625): bone cylinders of int(frame) + the checkerboard floor,
// lit by a point light (the reference's two orbiting sphere lights are not needed for what this checks).
static Mocap* g_mocap = nullptr;
static void buildSceneMocap(float frame) {
  Globals& g = globals();
  g_mocap->setSkeletonsToSpecifiedFrame(int(frame));
  g.shapes.clear(); g.lights.clear();
  float min_y = 0.251897;
  for (int x = 1; x < g_mocap->totalBones(); x++)
    g.shapes.push_back(make_shared<Cylinder>(g_mocap->leftVertex(x), g_mocap->rightVertex(x), 0.05, VEC3(1, 0, 0)));
  min_y = min_y + 0.05;
  float s = 1;
  VEC3 checker_rgb0(0.58, 0.82, 1), checker_rgb1(1, 0.416, 0.835);
  g.shapes.push_back(make_shared<Checkerboard>(VEC3(-6, min_y, -6), VEC3(-6, min_y, 6), VEC3(6, min_y, 6), VEC3(6, min_y, -6),
                                               checker_rgb0, checker_rgb1, s));
  g.lights.push_back(make_shared<pointLight>(VEC3(-3, 4, 2), VEC3(1, 1, 1)));
  g.eye = VEC3(-6, 0.5, 1) + (VEC3(0.49, 10, 1) - VEC3(-6, 0.5, 1)) * (float)(frame / 320);
}

// host_scene mesh <in.obj> <out.bin> [roughness.ppm]: loadObj + setMesh (column transform of scene.h:296-299), dumps the
// unified arrays: int64 n_vertices, n_triangles, has_uv; float vertices[3V]; int32 indices[3T]; float texcoords[2V];
// with a roughness map (loadTexture: binary PPM) also int64 n_materials; int32 material_ids[T]; double roughness[n_materials]
static int meshMode(const std::string& in, const std::string& out, const char* rough_ppm) {
  std::vector<VEC3> vertices; std::vector<VEC3I> v_inds, t_inds; std::vector<VEC2> texcoords;
  loadObj(in, vertices, v_inds, texcoords, t_inds);
  const double M[12] = {3, 0, 0, 3, 0, 3, 0, -1, 0, 0, 3, 5};
  auto mat = make_shared<Triangle>(VEC3(0, 0, 0), VEC3(1, 0, 0), VEC3(0, 1, 0), VEC3(0.75, 0.75, 0.75), "marble", false, "oren-nayar");
  std::vector<float> rough;
  if (rough_ppm) {
    const int tex = loadTexture(rough_ppm);                          // helpers.h:92-113
    const std::vector<uint8_t>& img = globals().texture_frames[tex];
    rough = faceRoughnessFromMap(texcoords, t_inds, img.data(), img.size(), (int)globals().texture_dims[tex][0], (int)globals().texture_dims[tex][1]);
  }
  setMesh(vertices, v_inds, texcoords, t_inds, mat, M, true, true, rough_ppm ? &rough : nullptr);
  Globals& g = globals();
  FlatScene f; flattenScene(f);
  if (!f.desc.mesh) return 1;
  std::ofstream o(out, std::ios::binary);
  int64_t hdr[3] = {f.mesh.n_vertices, f.mesh.n_triangles, f.mesh.texcoords ? 1 : 0};
  o.write((const char*)hdr, sizeof(hdr));
  o.write((const char*)g.mesh_vertices.data(), g.mesh_vertices.size() * sizeof(float));
  o.write((const char*)g.mesh_indices.data(), g.mesh_indices.size() * sizeof(int32_t));
  o.write((const char*)g.mesh_texcoords.data(), g.mesh_texcoords.size() * sizeof(float));
  if (rough_ppm) {
    int64_t nm = f.mesh.n_materials;
    o.write((const char*)&nm, sizeof(nm));
    o.write((const char*)f.mesh.material_ids, f.mesh.n_triangles * sizeof(int32_t));
    for (int k = 0; k < f.mesh.n_materials; k++) o.write((const char*)&f.mesh.materials[k].roughness, sizeof(double));
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: host_scene dump|render|video hw4|reflectance <out> | mesh <in.obj> <out.bin>\n"); return 2; }
  std::string mode = argv[1], scene = argv[2], out = argv[3];
  if (mode == "mesh") {
    try { return meshMode(scene, out, argc > 4 ? argv[4] : nullptr); } catch (const std::exception& e) { fprintf(stderr, "host_scene: %s\n", e.what()); return 1; }
  }
  Globals& g = globals();
  g.xRes = 160; g.yRes = 120; g.seed = 7;
  if (getenv("DRT_HOST_BLOCKS")) { g.always_blocks = true; g.block_rows = atoi(getenv("DRT_HOST_BLOCKS")); }
  int frame = 0;
  std::function<void(float)> builder;
  std::unique_ptr<Mocap> mocap;
  if (scene == "mocap") {       // host_scene dump|render|video mocap <out> <asf> <amc>   (needs a GPU: the clip is posed on the device)
    if (argc < 6) return 2;
    try { mocap.reset(new Mocap(argv[4], argv[5])); } catch (const std::exception& e) { fprintf(stderr, "host_scene: %s\n", e.what()); return 1; }
    g_mocap = mocap.get();
    g.antialias_samples = 4; frame = 30; builder = buildSceneMocap;
  } else
  if (scene == "hw4") { g.antialias_samples = 1; builder = buildSceneHW4; }
  else if (scene == "reflectance") { g.antialias_samples = 4; frame = 40; builder = buildSceneReflectance; }
  else return 2;
  builder((float)frame);
  try {
    if (mode == "dump") {
      FlatScene f; flattenScene(f);
      drt_settings st = flattenSettings(frame);
      std::ofstream o(out, std::ios::binary);
      int32_t n[2] = {(int32_t)f.prims.size(), (int32_t)f.lights.size()};
      o.write((const char*)n, sizeof(n));
      o.write((const char*)f.prims.data(), f.prims.size() * sizeof(drt_prim));
      o.write((const char*)f.lights.data(), f.lights.size() * sizeof(drt_light));
      o.write((const char*)&st, sizeof(st));
    } else if (mode == "video") {
      // frames 40..43 of the moving-light animation, `out` is a prefix: <out>.0040.ppm ...
      const int f0 = scene == "mocap" ? 30 : 40;
      const int n = renderVideo(f0, f0 + 4, builder, [&](int f) { char b[32]; snprintf(b, sizeof(b), ".%04d.ppm", f); return out + b; });
      if (n != 4) return 1;
    } else {
      renderImage(out, frame, builder);
    }
  } catch (const std::exception& e) { fprintf(stderr, "host_scene: %s\n", e.what()); return 1; }
  return 0;
}
