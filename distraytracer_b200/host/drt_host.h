// Host-side mirror of the reference's scene / settings surface for the per-pixel hot path.
//
// The reference keeps its whole "API" as globals plus one call:
//     vector<shared_ptr<GeoPrimitive>> shapes;  vector<shared_ptr<LightPrimitive>> lights;
//     texture_frames / texture_dims;  xRes, yRes, eye, lookingAt, up, aspect, near, fov,
//     aperture, focal_length, antialias_samples, brdf_samples, blur_samples, ...
//         (render_final_project.cpp:48-138)
//     void renderImage(const string& filename, const int frame,
//                      const function<void(float)> sceneBuilder);   (:965)
// This header offers the same names with the same meaning, so a scene builder written
// against the reference (scene.h) compiles against it with VEC3 spelled drt::host::VEC3, and
// renderImage() hands the frame to the CUDA library through the C ABI of include/drt.h
// instead of tracing on the CPU.  The classes are plain data holders: the constructors take
// the reference's argument lists (geometry.h:87-307) and fill the same members
// (geometry.cpp:83-104, 227-240, 433-445, 621-638, 784-813, 2248-2267, 2344-2364,
// 2563-2586, 2745-2843); intersection, normals, UVs and sampling live on the GPU.
//
// Header-only, C++17, depends only on include/drt.h; link with libdrt.so.
#ifndef DRT_HOST_H
#define DRT_HOST_H

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <vector>

#include "../../include/drt.h"

namespace drt {
namespace host {

struct VEC2 {
  double v[2];
  VEC2(double a = 0, double b = 0) : v{a, b} {}
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
struct VEC3 {
  double v[3];
  VEC3(double a = 0, double b = 0, double c = 0) : v{a, b, c} {}
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  VEC3 operator+(const VEC3& o) const { return VEC3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  VEC3 operator-(const VEC3& o) const { return VEC3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  VEC3 operator*(double s) const { return VEC3(v[0] * s, v[1] * s, v[2] * s); }
  VEC3 operator/(double s) const { return VEC3(v[0] / s, v[1] / s, v[2] / s); }
  double dot(const VEC3& o) const { return v[0] * o.v[0] + (v[1] * o.v[1] + v[2] * o.v[2]); }
  double norm() const { return std::sqrt(dot(*this)); }
  VEC3 normalized() const { double z = dot(*this); return z > 0 ? *this / std::sqrt(z) : *this; }
  VEC3 cross(const VEC3& o) const {
    return VEC3(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
  }
};
inline VEC3 operator*(double s, const VEC3& a) { return a * s; }

struct Reflectance {               // geometry.h:20-25
  std::string material;
  float roughness = 0;
  bool glossy = false;
  VEC2 refr;
};

// ---- primitives (data only) ---------------------------------------------------------
class GeoPrimitive {               // geometry.h:28-85
 public:
  virtual ~GeoPrimitive() {}
  virtual int drtType() const = 0;
  VEC3 color;
  bool light = false;
  bool motion = false;
  bool uv_verts = false;
  int dims = 0;
  std::string name;
  std::string model;
  bool texture = false;
  int tex_frame = -1;
  VEC3 bordercolor = VEC3(0, 0, 0);
  Reflectance reflect_params;
  bool mesh = false;
  VEC3 mesh_normal;
  VEC3 center;
  float radius = 0;
  VEC3 A, B, C, D, E, F, G, H;
  float height = 0, length = 0, width = 0;
  VEC3 c1, c2, axis;
  VEC2 uvA, uvB, uvC, uvD;
  VEC3 velocity;                   // extension: linear motion per frame (drt_prim::velocity)
};

class Sphere : public GeoPrimitive {
 public:
  Sphere() { center = VEC3(0, 0, 0); radius = 1; color = VEC3(1, 1, 1); dims = 3; model = "lambert"; name = "sphere"; }
  Sphere(VEC3 c, float r, VEC3 col, std::string material = "", bool in_motion = false, std::string shader = "lambert") {
    center = c; radius = r; color = col; reflect_params.material = material; dims = 3; motion = in_motion; model = shader; name = "sphere";
  }
  int drtType() const override { return DRT_PRIM_SPHERE; }
};

class Cylinder : public GeoPrimitive {
 public:
  Cylinder(VEC3 v1, VEC3 v2, float r, VEC3 col, std::string material = "", bool in_motion = false, std::string shader = "lambert") {
    c1 = v1; c2 = v2; axis = (v2 - v1).normalized(); radius = r; color = col; reflect_params.material = material; dims = 3;
    motion = in_motion; model = shader; center = (v1 + v2) / 2; name = "cylinder";
  }
  int drtType() const override { return DRT_PRIM_CYLINDER; }
};

class Triangle : public GeoPrimitive {
 public:
  Triangle(VEC3 a, VEC3 b, VEC3 c, VEC3 col, std::string material = "", bool in_motion = false, std::string shader = "lambert") {
    A = a; B = b; C = c; color = col; reflect_params.material = material; dims = 2; motion = in_motion; model = shader;
    center = (A + B + C) / 3; name = "triangle";
  }
  int drtType() const override { return DRT_PRIM_TRIANGLE; }
};

class Rectangle : public GeoPrimitive {
 public:
  Rectangle() {
    color = VEC3(1, 0, 0); length = 1; width = 1; A = VEC3(0, 0, 0); B = VEC3(1, 0, 0); C = VEC3(1, 1, 0); D = VEC3(0, 1, 0);
    dims = 2; model = "lambert"; center = (A + B + C + D) / 4; name = "rectangle";
  }
  // a,b,c,d MUST be clockwise/counter-clockwise (geometry.h:131)
  Rectangle(VEC3 a, VEC3 b, VEC3 c, VEC3 d, VEC3 col, std::string material = "", bool in_motion = false, int texframe = -1,
            std::string shader = "lambert") {
    A = a; B = b; C = c; D = d; length = (float)(b - a).norm(); width = (float)(d - a).norm(); color = col;
    reflect_params.material = material; dims = 2; motion = in_motion; model = shader; center = (A + B + C + D) / 4;
    name = "rectangle"; if (texframe >= 0) tex_frame = texframe;
  }
  int drtType() const override { return DRT_PRIM_RECTANGLE; }
};

class RectPrismV2 : public GeoPrimitive {
 public:
  RectPrismV2(VEC3 a, VEC3 b, VEC3 c, VEC3 d, VEC3 e, VEC3 f, VEC3 g, VEC3 h, VEC3 col, std::string material = "",
              bool in_motion = false, int texframe = -1, std::string shader = "lambert") {
    A = a; B = b; C = c; D = d; E = e; F = f; G = g; H = h;
    length = (float)(b - a).norm(); width = (float)(d - a).norm(); height = (float)(e - a).norm(); color = col;
    reflect_params.material = material; dims = 3; motion = in_motion; model = shader;
    center = (A + B + C + D + E + F + G + H) / 8; name = "rectprism"; if (texframe >= 0) tex_frame = texframe;
  }
  int drtType() const override { return DRT_PRIM_RECTPRISMV2; }
};

// The slab-box prisms (geometry.h:159-217, geometry.cpp:950-2246): the box tested is the WORLD axis-aligned bounding box
// of the eight corners.
class RectPrism : public GeoPrimitive {
 public:
  RectPrism() {}
  RectPrism(VEC3 a, VEC3 b, VEC3 c, VEC3 d, VEC3 e, VEC3 f, VEC3 g, VEC3 h, VEC3 col, std::string material = "",
            bool in_motion = false, int texframe = -1, std::string shader = "lambert") {
    A = a; B = b; C = c; D = d; E = e; F = f; G = g; H = h;
    length = (float)(b - a).norm(); width = (float)(d - a).norm(); height = (float)(e - a).norm(); color = col;
    reflect_params.material = material; dims = 3; motion = in_motion; model = shader;
    center = (A + B + C + D + E + F + G + H) / 8; name = "rectprism"; if (texframe >= 0) tex_frame = texframe;
  }
  int drtType() const override { return DRT_PRIM_RECTPRISM; }
};
class RectPrismWithCylinder : public RectPrism {
 public:
  RectPrismWithCylinder(VEC3 a, VEC3 b, VEC3 c, VEC3 d, VEC3 e, VEC3 f, VEC3 g, VEC3 h, VEC3 col, std::string material = "",
                        bool in_motion = false, int texframe = -1, std::string shader = "lambert")
      : RectPrism(a, b, c, d, e, f, g, h, col, material, in_motion, texframe, shader) { name = "RectPrismWithCylinder"; }
  int drtType() const override { return DRT_PRIM_RECTPRISM_CYL; }
  std::vector<std::shared_ptr<Cylinder>> holes;
};
class RectPrismWithHoles : public RectPrism {
 public:
  RectPrismWithHoles(VEC3 a, VEC3 b, VEC3 c, VEC3 d, VEC3 e, VEC3 f, VEC3 g, VEC3 h, VEC3 col, std::string material = "",
                     bool in_motion = false, int texframe = -1, std::string shader = "lambert")
      : RectPrism(a, b, c, d, e, f, g, h, col, material, in_motion, texframe, shader) {}
  int drtType() const override { return DRT_PRIM_RECTPRISM_HOLES; }
  std::vector<std::shared_ptr<GeoPrimitive>> holes;   // Sphere or Cylinder: the classes with an intersectMax (geometry.h:37)
};

class Checkerboard : public Rectangle {
 public:
  Checkerboard(VEC3 a, VEC3 b, VEC3 c, VEC3 d, VEC3 col1, VEC3 col2, float S_square, std::string material = "",
               bool in_motion = false, std::string shader = "lambert")
      : Rectangle(a, b, c, d, col1, material, in_motion, -1, shader) {
    color1 = col1; color2 = col2; S = S_square; name = "checkerboard";
  }
  int drtType() const override { return DRT_PRIM_CHECKERBOARD; }
  float S = 1;
  VEC3 color1, color2;
};

class CheckerboardWithHole : public Rectangle {
 public:
  CheckerboardWithHole(VEC3 a, VEC3 b, VEC3 c, VEC3 d, VEC3 col1, VEC3 col2, float S_square, std::shared_ptr<Rectangle> shape,
                       std::string material = "", bool in_motion = false, std::string shader = "lambert")
      : Rectangle(a, b, c, d, col1, material, in_motion, -1, shader) {
    color1 = col1; color2 = col2; S = S_square; hole = shape; name = "checkboardhole";
  }
  int drtType() const override { return DRT_PRIM_CHECKERBOARD_HOLE; }
  float S = 1;
  VEC3 color1, color2;
  std::shared_ptr<Rectangle> hole;
  float borderwidth = 0;
};

class CheckerCylinder : public Cylinder {
 public:
  CheckerCylinder(VEC3 v1, VEC3 v2, float r, VEC3 col, float s, std::string material = "", bool in_motion = false,
                  std::string shader = "lambert")
      : Cylinder(v1, v2, r, col, material, in_motion, shader) { S = s; name = "checkercylinder"; }
  int drtType() const override { return DRT_PRIM_CHECKER_CYLINDER; }
  float S = 1;
  float borderwidth = 0;
};

// ---- lights (geometry.h:279-307) ------------------------------------------------------
class LightPrimitive {
 public:
  virtual ~LightPrimitive() {}
  virtual int drtLightType() const = 0;
  VEC3 color;
  VEC3 center;
};
class pointLight : public LightPrimitive {
 public:
  pointLight(const VEC3 c, const VEC3 col) { center = c; color = col; }
  int drtLightType() const override { return DRT_LIGHT_POINT; }
};
class sphereLight : public LightPrimitive, public Sphere {
 public:
  sphereLight(const VEC3 c, const float r, const VEC3 col, std::string material = "", bool in_motion = false) {
    light = true; Sphere::center = c; LightPrimitive::center = c; Sphere::color = col; LightPrimitive::color = col; radius = r;
    reflect_params.material = material; motion = in_motion; name = "spherelight";
  }
  int drtLightType() const override { return DRT_LIGHT_SPHERE; }
  VEC3 baxis = VEC3(0, 0, 0);
};
class rectangleLight : public LightPrimitive, public Rectangle {
 public:
  rectangleLight(VEC3 a, VEC3 b, VEC3 c, VEC3 d, VEC3 col, std::string material = "", bool in_motion = false) {
    light = true; A = a; B = b; C = c; D = d; Rectangle::center = (a + b + c + d) / 4; LightPrimitive::center = (a + b + c + d) / 4;
    Rectangle::color = col; LightPrimitive::color = col; reflect_params.material = material; motion = in_motion; name = "rectanglelight";
  }
  int drtLightType() const override { return DRT_LIGHT_RECT; }
};

// ---- the globals renderImage reads (render_final_project.cpp:48-138) -------------------
struct Globals {
  int xRes = 1920, yRes = 1080;
  VEC3 eye = VEC3(-6, 0.5, 1), lookingAt = VEC3(0.5, 0.5, 1), up = VEC3(0, 1, 0);
  float aspect = (float)1920 / (float)1080, near = 1, fov = 45.0f, aperture = 0.2f, focal_length = 10;
  bool nogloss = false;
  float refr_air = 1, refr_glass = 1.5f;
  int max_depth = 10;
  std::vector<std::shared_ptr<GeoPrimitive>> shapes;
  std::vector<std::shared_ptr<LightPrimitive>> lights;
  float phong = 10;
  int antialias_samples = 10, brdf_samples = 2, blur_samples = 2, frame_range = 1;
  std::vector<std::vector<uint8_t>> texture_frames;   // RGB bytes (the reference stores byte/255 as doubles)
  std::vector<VEC2> texture_dims;
  int frame_prism = 960, frame_cloud = 1952, frame_blur = 1600;
  float move_per_frame = (float)(0.1 / 8), accel_t = (float)(80 / std::pow(360, 3));
  VEC3 sundir = VEC3(0, 0.1, -1);
  bool perlin_cloud = false;
  float saturation = 0.2f, clouddist = 10, cloudhoff = 0.2f;
  VEC3 sun_outer = VEC3(0.9, 0.3, 0.9), sun_inner = VEC3(1.0, 0.7, 0.7), sun_core = VEC3(1, 1, 1), bluesky = VEC3(0.3, 0.55, 0.8),
       redsky = VEC3(0.8, 0.8, 0.6);
  bool reflect = true;
  // extensions
  uint32_t seed = 0;
  int blur_mode = DRT_BLUR_REFERENCE;
  int precision = DRT_PRECISION_REFERENCE;
  int devices = 0;   // 0 = all visible GPUs
  int block_rows = 15;   // multi-GPU single frames: rows per dynamically claimed block (renderFrame); with two streams per
                         // GPU 30-row blocks are too coarse (8 GPUs, 1080p: 57.8 ms vs 34.2 ms with 15 rows)
  bool always_blocks = false;   // cut into blocks even on one GPU (tests)
  int streams_per_gpu = 2;      // renderFrame: scene handles per GPU when the frame is cut into blocks
  // optional indexed triangle mesh (what loadObj + the scene builders' triangle loop produce, scene.h:296-386),
  // kept as arrays and traversed through the device-built LBVH instead of 10^6 Triangle shapes
  std::vector<float> mesh_vertices, mesh_texcoords;
  std::vector<int32_t> mesh_indices;
  std::shared_ptr<GeoPrimitive> mesh_material;
  std::vector<float> mesh_face_roughness;     // optional: Reflectance::roughness per face (faceRoughnessFromMap, scene.h:372-378)
  std::vector<int32_t> mesh_material_ids;     // derived by flattenScene
  // loadTexture(path) decodes binary PPM (P6) itself; any other format goes through this hook (path -> width, height,
  // RGB bytes as stbi_load returns them), e.g. a wrapper around the stb_image.h the reference vendors
  std::function<bool(const std::string&, int&, int&, std::vector<uint8_t>&)> decode_image;
};
inline Globals& globals() { static Globals g; return g; }

// loadTexture's result for already decoded pixels (helpers.h:92-113 stores them the same way)
inline int addTexture(int width, int height, const uint8_t* rgb) {
  Globals& g = globals();
  g.texture_frames.emplace_back(rgb, rgb + (size_t)width * height * 3);
  g.texture_dims.push_back(VEC2(width, height));
  return (int)g.texture_frames.size() - 1;
}

struct VEC3I { int v[3]; int& operator[](int i) { return v[i]; } int operator[](int i) const { return v[i]; } };
// loadTexture (helpers.h:92-113): push one image onto texture_frames / texture_dims.  The reference decodes with stb_image;
// this mirror reads binary PPM itself and leaves every other format to Globals::decode_image.  Prints and throws like
// the reference when the image cannot be loaded.
inline int loadTexture(const std::string& f) {
  Globals& g = globals();
  int width = 0, height = 0;
  std::vector<uint8_t> rgb;
  bool ok = false;
  {
    std::ifstream in(f, std::ios::binary);
    std::string magic;
    if (in && (in >> magic) && magic == "P6") {
      auto next_int = [&]() { int v = -1; for (;;) { in >> std::ws; if (in.peek() == '#') { std::string c; std::getline(in, c); } else break; } in >> v; return v; };
      width = next_int(); height = next_int();
      const int maxval = next_int();
      in.get();                                               // the single whitespace after maxval
      if (width > 0 && height > 0 && maxval == 255) {
        rgb.resize((size_t)width * height * 3);
        in.read((char*)rgb.data(), (std::streamsize)rgb.size());
        ok = (size_t)in.gcount() == rgb.size();
      }
    }
  }
  if (!ok && g.decode_image) ok = g.decode_image(f, width, height, rgb) && width > 0 && height > 0 && rgb.size() == (size_t)width * height * 3;
  if (!ok) { std::cout << "Image loading failed for: " << f << std::endl; throw std::runtime_error("Image loading failed for: " + f); }
  return addTexture(width, height, rgb.data());
}

// The per-face roughness the reference's model builders assign (scene.h:372-378): the roughness map is indexed with the
// ORIGINAL texcoords of the three corners as `map[int(u * int(w-1) + v * int(h-1) * (w-1))]` -- a byte offset into the
// decoded image, channel count ignored, exactly as written there -- and the face gets (r1 + r2 + r3) / (3 * 255).
inline std::vector<float> faceRoughnessFromMap(const std::vector<VEC2>& texcoords, const std::vector<VEC3I>& t_indices,
                                               const uint8_t* map_bytes, size_t map_size, int width, int height) {
  std::vector<float> out;
  for (const VEC3I& t : t_indices) {
    float r[3];
    for (int k = 0; k < 3; k++) {
      const VEC2& uv = texcoords[t[k]];
      const long long at = (long long)(int)(uv[0] * (int)(width - 1) + uv[1] * (int)(height - 1) * (width - 1));
      if (at < 0 || (size_t)at >= map_size) throw std::runtime_error("roughness map lookup outside the image");   // the reference reads out of bounds
      r[k] = map_bytes[at];
    }
    out.push_back((r[0] + r[1] + r[2]) / (3 * 255));
  }
  return out;
}

// ---- flattening ---------------------------------------------------------------------------
inline int materialTag(const std::string& m) {   // refl_materials, render_final_project.cpp:64
  if (m == "glass") return DRT_MAT_GLASS;
  if (m == "steel") return DRT_MAT_STEEL;
  if (m == "aluminum") return DRT_MAT_ALUMINUM;
  if (m == "water") return DRT_MAT_WATER;
  if (m == "linoleum") return DRT_MAT_LINOLEUM;
  return DRT_MAT_NONE;
}
inline int modelTag(const std::string& m) {
  if (m == "oren-nayar") return DRT_MODEL_OREN_NAYAR;
  if (m == "cook-torrance") return DRT_MODEL_COOK_TORRANCE;
  if (m == "raw") return DRT_MODEL_RAW;
  return DRT_MODEL_LAMBERT;
}
inline int nameTag(const std::string& n) {
  if (n == "rectangle") return DRT_NAME_RECTANGLE;
  if (n == "spherelight") return DRT_NAME_SPHERELIGHT;
  if (n == "rectanglelight") return DRT_NAME_RECTANGLELIGHT;
  return DRT_NAME_OTHER;
}
inline void put3(double* o, const VEC3& v) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }

inline drt_prim flattenPrim(const GeoPrimitive& s) {
  drt_prim p;
  drt_prim_default(&p);
  p.type = s.drtType(); p.name = nameTag(s.name); p.material = materialTag(s.reflect_params.material); p.model = modelTag(s.model);
  p.flags = (s.light ? DRT_FLAG_LIGHT : 0) | (s.motion ? DRT_FLAG_MOTION : 0) | (s.texture ? DRT_FLAG_TEXTURE : 0) |
            (s.reflect_params.glossy ? DRT_FLAG_GLOSSY : 0) | (s.mesh ? DRT_FLAG_MESH : 0) | (s.uv_verts ? DRT_FLAG_UV_VERTS : 0);
  p.tex_frame = s.texture ? s.tex_frame : -1;
  put3(p.color, s.color); put3(p.bordercolor, s.bordercolor);
  p.roughness = s.reflect_params.roughness; p.refr[0] = s.reflect_params.refr[0]; p.refr[1] = s.reflect_params.refr[1];
  put3(p.center, s.center); p.radius = s.radius;
  put3(p.A, s.A); put3(p.B, s.B); put3(p.C, s.C); put3(p.D, s.D); put3(p.E, s.E); put3(p.F, s.F); put3(p.G, s.G); put3(p.H, s.H);
  put3(p.c1, s.c1); put3(p.c2, s.c2);
  p.uvA[0] = s.uvA[0]; p.uvA[1] = s.uvA[1]; p.uvB[0] = s.uvB[0]; p.uvB[1] = s.uvB[1]; p.uvC[0] = s.uvC[0]; p.uvC[1] = s.uvC[1];
  put3(p.mesh_normal, s.mesh_normal); put3(p.velocity, s.velocity);
  if (auto* c = dynamic_cast<const Checkerboard*>(&s)) { p.S = c->S; put3(p.color1, c->color1); put3(p.color2, c->color2); }
  if (auto* c = dynamic_cast<const CheckerboardWithHole*>(&s)) {
    p.S = c->S; p.borderwidth = c->borderwidth; put3(p.color1, c->color1); put3(p.color2, c->color2);
    put3(p.hole[0], c->hole->A); put3(p.hole[1], c->hole->B); put3(p.hole[2], c->hole->C); put3(p.hole[3], c->hole->D);
  }
  if (auto* c = dynamic_cast<const CheckerCylinder*>(&s)) { p.S = c->S; p.borderwidth = c->borderwidth; }
  auto put_hole = [&](const GeoPrimitive& h) {
    if (p.n_holes >= DRT_MAX_HOLES) throw std::runtime_error("more holes than DRT_MAX_HOLES");
    drt_hole& o = p.holes[p.n_holes++];
    o.type = h.drtType(); o.radius = h.radius; put3(o.color, h.color);
    if (o.type == DRT_PRIM_SPHERE) put3(o.c1, h.center); else { put3(o.c1, h.c1); put3(o.c2, h.c2); }
  };
  if (auto* c = dynamic_cast<const RectPrismWithCylinder*>(&s)) for (auto& h : c->holes) put_hole(*h);
  if (auto* c = dynamic_cast<const RectPrismWithHoles*>(&s)) for (auto& h : c->holes) put_hole(*h);
  return p;
}

struct FlatScene {
  std::vector<drt_prim> mesh_materials;
  std::vector<drt_prim> prims;
  std::vector<drt_light> lights;
  std::vector<drt_texture> textures;
  drt_mesh mesh;
  drt_scene_desc desc;
};

inline void flattenScene(FlatScene& f) {
  Globals& g = globals();
  f.prims.clear(); f.lights.clear(); f.textures.clear();
  for (auto& s : g.shapes) f.prims.push_back(flattenPrim(*s));
  for (auto& l : g.lights) {
    drt_light o{};
    o.type = l->drtLightType(); o.prim_index = -1;
    put3(o.color, l->color); put3(o.center, l->center);
    // an area light is the same object as one of the shapes (render_final_project.cpp:832-837)
    for (size_t k = 0; k < g.shapes.size(); k++)
      if (dynamic_cast<void*>(g.shapes[k].get()) == dynamic_cast<void*>(l.get())) { o.prim_index = (int)k; break; }
    if (auto* sl = dynamic_cast<sphereLight*>(l.get())) { o.radius = sl->radius; put3(o.baxis, sl->baxis); put3(o.center, sl->Sphere::center); }
    if (auto* rl = dynamic_cast<rectangleLight*>(l.get())) { put3(o.A, rl->A); put3(o.B, rl->B); put3(o.C, rl->C); put3(o.D, rl->D); }
    f.lights.push_back(o);
  }
  for (size_t i = 0; i < g.texture_frames.size(); i++) {
    drt_texture t; t.width = (int)g.texture_dims[i][0]; t.height = (int)g.texture_dims[i][1]; t.rgb = g.texture_frames[i].data();
    f.textures.push_back(t);
  }
  f.desc.abi_version = DRT_ABI_VERSION;
  f.desc.n_prims = (int)f.prims.size(); f.desc.prims = f.prims.data();
  f.desc.n_lights = (int)f.lights.size(); f.desc.lights = f.lights.data();
  f.desc.n_textures = (int)f.textures.size(); f.desc.textures = f.textures.data();
  f.desc.mesh = nullptr;
  if (!g.mesh_indices.empty() && g.mesh_material) {
    f.mesh.n_vertices = (int64_t)g.mesh_vertices.size() / 3; f.mesh.n_triangles = (int64_t)g.mesh_indices.size() / 3;
    f.mesh.vertices = g.mesh_vertices.data(); f.mesh.indices = g.mesh_indices.data();
    f.mesh.texcoords = g.mesh_texcoords.empty() ? nullptr : g.mesh_texcoords.data();
    f.mesh.material = flattenPrim(*g.mesh_material);
    f.mesh.n_materials = 0; f.mesh.materials = nullptr; f.mesh.material_ids = nullptr;
    if (!g.mesh_face_roughness.empty()) {
      // one material per distinct per-face roughness (scene.h:372-378: at most 766 values), triangle -> table index
      if (g.mesh_face_roughness.size() * 3 != g.mesh_indices.size()) throw std::runtime_error("setMesh: one roughness per face expected");
      std::map<float, int> table;
      g.mesh_material_ids.clear(); f.mesh_materials.clear();
      for (float r : g.mesh_face_roughness) {
        auto it = table.find(r);
        if (it == table.end()) {
          it = table.emplace(r, (int)f.mesh_materials.size()).first;
          drt_prim m = f.mesh.material; m.roughness = r;
          f.mesh_materials.push_back(m);
        }
        g.mesh_material_ids.push_back(it->second);
      }
      f.mesh.n_materials = (int32_t)f.mesh_materials.size(); f.mesh.materials = f.mesh_materials.data();
      f.mesh.material_ids = g.mesh_material_ids.data();
    }
    f.desc.mesh = &f.mesh;
  }
}

inline drt_settings flattenSettings(int frame) {
  Globals& g = globals();
  drt_settings s;
  drt_settings_default(&s);
  s.xRes = g.xRes; s.yRes = g.yRes; put3(s.eye, g.eye); put3(s.lookingAt, g.lookingAt); put3(s.up, g.up);
  s.aspect = g.aspect; s.near_plane = g.near; s.fov = g.fov; s.aperture = g.aperture; s.focal_length = g.focal_length;
  s.nogloss = g.nogloss; s.refr_air = g.refr_air; s.refr_glass = g.refr_glass; s.max_depth = g.max_depth; s.phong = g.phong;
  s.antialias_samples = g.antialias_samples; s.brdf_samples = g.brdf_samples; s.blur_samples = g.blur_samples;
  s.frame_range = g.frame_range; s.frame_prism = g.frame_prism; s.frame_cloud = g.frame_cloud; s.frame_blur = g.frame_blur;
  s.move_per_frame = g.move_per_frame; s.accel_t = g.accel_t; put3(s.sundir, g.sundir); s.perlin_cloud = g.perlin_cloud;
  s.saturation = g.saturation; s.clouddist = g.clouddist; s.cloudhoff = g.cloudhoff;
  put3(s.sun_outer, g.sun_outer); put3(s.sun_inner, g.sun_inner); put3(s.sun_core, g.sun_core); put3(s.bluesky, g.bluesky);
  put3(s.redsky, g.redsky);
  s.reflect = g.reflect; s.frame = frame; s.seed = g.seed; s.blur_mode = g.blur_mode; s.precision = g.precision;
  return s;
}

// Renders the frame into `rgb` (xRes*yRes*3 bytes, PPM row order) on all visible GPUs.  The frame is cut into
// blocks of `globals().block_rows` rows that the per-GPU host threads claim from a shared counter: sky rows and
// object rows differ in cost by orders of magnitude, so equal bands would leave GPUs idle.  Every block
// is written straight into `rgb` by its own drt_render (the "final gather" is that device-to-host copy); samples are
// keyed by pixel, so the picture does not depend on how the frame was cut.
inline void renderFrame(int frame, std::vector<uint8_t>& rgb) {
  Globals& g = globals();
  if (g.shapes.size() < 1 && g.mesh_indices.empty()) throw std::runtime_error("No shapes to render!");   // render_final_project.cpp:973-977
  FlatScene f;
  flattenScene(f);
  const drt_settings st = flattenSettings(frame);
  int ndev = drt_device_count();
  if (ndev < 1) throw std::runtime_error("no CUDA device: distraytracer-b200 has no CPU fallback");
  if (g.devices > 0 && g.devices < ndev) ndev = g.devices;
  rgb.assign((size_t)st.xRes * st.yRes * 3, 0);
  if (ndev > 1 && !g.always_blocks) {
    // one launch per GPU, all claiming ~1024-sample units of the frame from one counter in GPU 0's memory and resolving
    // their pixels into GPU 0's frame over NVLink (drt_render_multi); falls through to host-claimed row blocks when the
    // devices have no peer access
    std::vector<drt_scene*> sc(ndev, nullptr);
    std::string err;
    int rc = DRT_OK;
    for (int d = 0; d < ndev && rc == DRT_OK; d++) { rc = drt_scene_create(&f.desc, d, &sc[d]); if (rc != DRT_OK) err = drt_last_error(); }
    if (rc == DRT_OK) {
      const drt_tile whole{0, 0, st.xRes, st.yRes, 0};
      rc = drt_render_multi(sc.data(), ndev, &st, &whole, rgb.data(), nullptr);
      if (rc != DRT_OK) err = drt_last_error();
    }
    for (drt_scene* h : sc) drt_scene_destroy(h);
    if (rc == DRT_OK) return;
    if (rc != DRT_ERR_UNSUPPORTED) throw std::runtime_error(err);
  }
  const int rows = (ndev == 1 && !g.always_blocks) ? st.yRes : std::max(1, std::min(g.block_rows, st.yRes));   // one GPU: one launch sequence
  const int n_blocks = (st.yRes + rows - 1) / rows;
  if (ndev > n_blocks) ndev = n_blocks;
  // Two scene handles (streams) per GPU when the frame is cut: render_wave is a persistent kernel whose last batches
  // drain with few warps busy; with a second stream the SMs a finished block frees start on the next block at once
  // (measured: a 1080p 64 spp frame in 36 blocks on one GPU 248.7 ms with one stream, 208.1 ms with two; whole frame 210).
  const int per_dev = (n_blocks > ndev) ? std::max(1, g.streams_per_gpu) : 1;
  const int nth = ndev * per_dev;
  std::vector<std::string> errs(nth);
  std::atomic<int> next{0};
  std::vector<std::thread> th;
  for (int w = 0; w < nth; w++)
    th.emplace_back([&, w]() {
      const int d = w / per_dev;
      drt_scene* sc = nullptr;
      if (drt_scene_create(&f.desc, d, &sc) != DRT_OK) { errs[w] = drt_last_error(); return; }
      for (int b = next.fetch_add(1); b < n_blocks; b = next.fetch_add(1)) {
        // loop rows [y0,y1) of block b; buffer row 0 of the block is loop row y1-1
        const int y0 = b * rows, y1 = std::min(st.yRes, y0 + rows);
        drt_tile tile{0, y0, st.xRes, y1 - y0, d};
        uint8_t* dst = rgb.data() + (size_t)(st.yRes - y1) * st.xRes * 3;
        if (drt_render(sc, &st, &tile, dst, nullptr) != DRT_OK) { errs[w] = drt_last_error(); break; }
      }
      drt_scene_destroy(sc);
    });
  for (auto& t : th) t.join();
  for (auto& e : errs) if (!e.empty()) throw std::runtime_error(e);
}

// The reference's entry point (render_final_project.cpp:965): render `frame` of the current
// globals and write `filename` as a binary PPM.  `sceneBuilder` is accepted for signature
// compatibility; like the reference's BVH path it is not called.
inline void renderImage(const std::string& filename, const int frame, const std::function<void(float)> sceneBuilder) {
  (void)sceneBuilder;
  std::vector<uint8_t> rgb;
  renderFrame(frame, rgb);
  Globals& g = globals();
  if (drt_write_ppm(filename.c_str(), g.xRes, g.yRes, rgb.data()) != DRT_OK) throw std::runtime_error(drt_last_error());
}

// ---- mesh ingest ------------------------------------------------------

// loadObj (objHelper.h:6-85) without tiny_obj_loader: positions, texcoords and per-face index
// triples, polygons as triangle fans (tiny_obj_loader's default triangulation), indices 0-based,
// -1 where a face corner has no texcoord.  Normals are not read: the path shades flat normals
// (geometry.cpp:588-594) and the reference fills `normals` from positions anyway (quirk Q13).
inline void loadObj(const std::string& path, std::vector<VEC3>& vertices, std::vector<VEC3I>& v_indices,
                    std::vector<VEC2>& texcoords, std::vector<VEC3I>& t_indices) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("loadObj: cannot open " + path);
  std::string line;
  while (std::getline(in, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.resize(hash);
    std::istringstream ls(line);
    std::string key;
    if (!(ls >> key)) continue;
    if (key == "v") { double x, y, z; if (!(ls >> x >> y >> z)) throw std::runtime_error("loadObj: bad vertex"); vertices.push_back(VEC3(x, y, z)); }
    else if (key == "vt") { double u, v = 0; if (!(ls >> u)) throw std::runtime_error("loadObj: bad texcoord"); ls >> v; texcoords.push_back(VEC2(u, v)); }
    else if (key == "f") {
      std::vector<std::pair<int, int>> corners;
      std::string c;
      while (ls >> c) {
        int vi = 0, ti = 0; bool has_t = false;
        const size_t s1 = c.find('/');
        vi = std::stoi(c.substr(0, s1));
        if (s1 != std::string::npos) {
          const size_t s2 = c.find('/', s1 + 1);
          const std::string t = c.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
          if (!t.empty()) { ti = std::stoi(t); has_t = true; }
        }
        vi = vi > 0 ? vi - 1 : (int)vertices.size() + vi;
        ti = !has_t ? -1 : (ti > 0 ? ti - 1 : (int)texcoords.size() + ti);
        if (vi < 0 || vi >= (int)vertices.size() || (has_t && (ti < 0 || ti >= (int)texcoords.size())))
          throw std::runtime_error("loadObj: face index out of range");
        corners.push_back({vi, ti});
      }
      if (corners.size() < 3) throw std::runtime_error("loadObj: face with fewer than 3 corners");
      for (size_t k = 1; k + 1 < corners.size(); k++) {
        v_indices.push_back(VEC3I{{corners[0].first, corners[k].first, corners[k + 1].first}});
        t_indices.push_back(VEC3I{{corners[0].second, corners[k].second, corners[k + 1].second}});
      }
    }
  }
}

// The scene builders' triangle loop (scene.h:296-386) for a whole model at once: positions through
// the 3x4 object transform `M` (row major, or nullptr), texcoords above 1 wrapped by dropping the
// integer part, the bounds check that makes the reference throw, V flipped -- then (position,
// texcoord) index pairs unified into vertices, because drt_mesh carries one texcoord per vertex.
// `material` is the Triangle the reference would copy per face (colour, model, roughness, texture).
inline void setMesh(const std::vector<VEC3>& vertices, const std::vector<VEC3I>& v_indices, const std::vector<VEC2>& texcoords,
                    const std::vector<VEC3I>& t_indices, std::shared_ptr<GeoPrimitive> material, const double* M = nullptr,
                    bool wrap_uv = true, bool flip_v = true, const std::vector<float>* face_roughness = nullptr) {
  Globals& g = globals();
  g.mesh_vertices.clear(); g.mesh_texcoords.clear(); g.mesh_indices.clear();
  g.mesh_face_roughness.clear();
  if (face_roughness) g.mesh_face_roughness = *face_roughness;
  bool has_uv = !texcoords.empty();
  for (const VEC3I& t : t_indices) for (int k = 0; k < 3; k++) if (t[k] < 0) has_uv = false;
  std::vector<VEC2> uv = texcoords;
  if (has_uv) {
    for (VEC2& t : uv) {
      if (wrap_uv) for (int k = 0; k < 2; k++) if (t[k] > 1) t[k] = t[k] - (int)t[k];
    }
    for (const VEC3I& t : t_indices) for (int k = 0; k < 3; k++)
      if (!(uv[t[k]][0] >= 0 && uv[t[k]][1] <= 1)) throw std::runtime_error("Texcoords out of bounds");   // scene.h:343-354
    if (flip_v) for (VEC2& t : uv) t[1] = 1 - t[1];
  }
  std::map<std::pair<int, int>, int> unified;
  for (size_t f = 0; f < v_indices.size(); f++)
    for (int k = 0; k < 3; k++) {
      const std::pair<int, int> key{v_indices[f][k], has_uv ? t_indices[f][k] : -1};
      auto it = unified.find(key);
      if (it == unified.end()) {
        it = unified.emplace(key, (int)(g.mesh_vertices.size() / 3)).first;
        VEC3 p = vertices[key.first];
        if (M) p = VEC3(M[0] * p[0] + M[1] * p[1] + M[2] * p[2] + M[3], M[4] * p[0] + M[5] * p[1] + M[6] * p[2] + M[7],
                        M[8] * p[0] + M[9] * p[1] + M[10] * p[2] + M[11]);
        for (int a = 0; a < 3; a++) g.mesh_vertices.push_back((float)p[a]);
        if (has_uv) { g.mesh_texcoords.push_back((float)uv[key.second][0]); g.mesh_texcoords.push_back((float)uv[key.second][1]); }
      }
      g.mesh_indices.push_back(it->second);
    }
  g.mesh_material = std::move(material);
}

// ---- video (SURVEY.md 8(e) + 8(f)3) -----------------------------------------------------
// Frames [first, last) of an animation: frame f goes to GPU (f - first) mod G; every GPU keeps ONE
// resident scene and only re-uploads the primitives and lights `pose(f)` produced (the re-posed mocap
// bones, scene.h:637-659; a light that moves with the frame, scene.h:3690-3692).  PPM files are written by a separate thread from a queue of finished frames, so
// the 6 MB-per-frame file I/O overlaps the next renders (the reference's shell loop runs build,
// render and writePPM back to back per frame).  `pose` is called under a lock, in frame order per GPU;
// it must leave the number and types of shapes unchanged.  Returns the number of frames written.
inline int renderVideo(int first, int last, const std::function<void(float)>& pose, const std::function<std::string(int)>& filename) {
  Globals& g = globals();
  int ndev = drt_device_count();
  if (ndev < 1) throw std::runtime_error("no CUDA device: distraytracer-b200 has no CPU fallback");
  if (g.devices > 0 && g.devices < ndev) ndev = g.devices;
  if (last - first < ndev) ndev = std::max(1, last - first);
  struct Done { int frame; std::vector<uint8_t> rgb; };
  std::mutex pose_mu, q_mu;
  std::condition_variable q_cv;
  std::deque<Done> queue;
  bool finished = false;
  std::string write_err;
  int written = 0;
  std::thread writer([&]() {
    for (;;) {
      Done d;
      {
        std::unique_lock<std::mutex> lk(q_mu);
        q_cv.wait(lk, [&] { return finished || !queue.empty(); });
        if (queue.empty()) return;
        d = std::move(queue.front()); queue.pop_front();
      }
      q_cv.notify_all();
      if (drt_write_ppm(filename(d.frame).c_str(), g.xRes, g.yRes, d.rgb.data()) != DRT_OK) write_err = drt_last_error();
      else written++;
    }
  });
  std::vector<std::string> errs(ndev);
  std::vector<std::thread> th;
  for (int d = 0; d < ndev; d++)
    th.emplace_back([&, d]() {
      drt_scene* sc = nullptr;
      for (int f = first + d; f < last; f += ndev) {
        FlatScene fs; drt_settings st;
        {
          std::lock_guard<std::mutex> lk(pose_mu);        // the builders mutate the shared globals
          pose((float)f);
          flattenScene(fs); st = flattenSettings(f);
          if (!sc && drt_scene_create(&fs.desc, d, &sc) != DRT_OK) { errs[d] = drt_last_error(); return; }
        }
        if (drt_scene_update_prims(sc, fs.prims.data(), (int)fs.prims.size()) != DRT_OK ||
            drt_scene_update_lights(sc, fs.lights.data(), (int)fs.lights.size()) != DRT_OK) { errs[d] = drt_last_error(); break; }
        Done out; out.frame = f; out.rgb.assign((size_t)st.xRes * st.yRes * 3, 0);
        drt_tile tile{0, 0, st.xRes, st.yRes, d};
        if (drt_render(sc, &st, &tile, out.rgb.data(), nullptr) != DRT_OK) { errs[d] = drt_last_error(); break; }
        {
          std::unique_lock<std::mutex> lk(q_mu);
          q_cv.wait(lk, [&] { return queue.size() < 8; });   // bound the host memory held by unwritten frames
          queue.push_back(std::move(out));
        }
        q_cv.notify_all();
      }
      if (sc) drt_scene_destroy(sc);
    });
  for (auto& t : th) t.join();
  { std::lock_guard<std::mutex> lk(q_mu); finished = true; }
  q_cv.notify_all();
  writer.join();
  for (auto& e : errs) if (!e.empty()) throw std::runtime_error(e);
  if (!write_err.empty()) throw std::runtime_error(write_err);
  return written;
}

// ---- mocap: the `displayer` / `skeleton` / `motion` globals of render_final_project.cpp --------------------
// The reference loads "90.asf" / "90_16_v3.amc" once in main(), then every build*() function calls
// setSkeletonsToSpecifiedFrame(int(frame)) + displayer.ComputeBonePositions() and turns the per-bone
// rotation / scaling / translation / length into cylinder end points (scene.h:109-128, 616-644, 3560-3610).
// Here the clip is posed for ALL frames on the GPU when it is loaded (drt_skeleton_create, kernel skeleton_fk);
// a builder asks for the end points of a frame and constructs its Cylinders as before:
//     Mocap mocap("90.asf", "90_16_v3.amc");
//     mocap.setSkeletonsToSpecifiedFrame(int(frame));
//     for (int x = 1; x < mocap.totalBones(); x++)
//       shapes.push_back(make_shared<Cylinder>(mocap.leftVertex(x), mocap.rightVertex(x), 0.05, VEC3(1,0,0)));
constexpr double MOCAP_SCALE = 0.06;            // types.h:6

class Mocap {
 public:
  Mocap(const std::string& asf_path, const std::string& amc_path, double scale = MOCAP_SCALE, int device = 0) {
    if (drt_skeleton_load(asf_path.c_str(), amc_path.c_str(), scale, device, &skel_) != DRT_OK)
      throw std::runtime_error(drt_last_error());            // the reference throws 1 (skeleton.cpp:575-577)
    int32_t nc = 0, nf = 0;
    drt_skeleton_info(skel_, &nc, &nf, &fk_ms_);
    cylinders_ = nc; frames_ = nf;
    table_.resize((size_t)nf * nc * 6);
    if (drt_skeleton_bones(skel_, 0, nf, table_.data()) != DRT_OK) throw std::runtime_error(drt_last_error());
  }
  ~Mocap() { drt_skeleton_destroy(skel_); }
  Mocap(const Mocap&) = delete;
  Mocap& operator=(const Mocap&) = delete;
  int GetNumFrames() const { return frames_; }               // Motion::GetNumFrames
  int totalBones() const { return cylinders_ + 1; }          // rotations.size(): bone 0 is the origin and is skipped
  float fkMilliseconds() const { return fk_ms_; }
  const drt_skeleton* handle() const { return skel_; }
  // scene.h:109-128: clamps past the last frame, a negative index is fatal
  void setSkeletonsToSpecifiedFrame(int frameIndex) {
    if (frameIndex < 0) throw std::runtime_error("Error in SetSkeletonsToSpecifiedFrame: frameIndex is illegal.");
    frame_ = frameIndex >= frames_ ? frames_ - 1 : frameIndex;
  }
  // end points of bone x (1 <= x < totalBones()) at the current frame: what
  // rotation * scaling * (0,0,0|length,1) + translation evaluates to (scene.h:637-644)
  VEC3 leftVertex(int x) const { const double* p = at(x); return VEC3(p[0], p[1], p[2]); }
  VEC3 rightVertex(int x) const { const double* p = at(x); return VEC3(p[3], p[4], p[5]); }

 private:
  const double* at(int x) const {
    if (x < 1 || x > cylinders_) throw std::out_of_range("bone index");
    return &table_[((size_t)frame_ * cylinders_ + (x - 1)) * 6];
  }
  drt_skeleton* skel_ = nullptr;
  int cylinders_ = 0, frames_ = 0, frame_ = 0;
  float fk_ms_ = 0;
  std::vector<double> table_;
};

}  // namespace host
}  // namespace drt
#endif
