// drt_flatten.h -- the reference's object graph -> the PODs of include/drt.h.
//
// To be included AFTER the reference's own headers / globals (render_final_project.cpp:48-138, geometry.h, scene.h): it
// reads `shapes`, `lights`, `texture_frames`, `texture_dims` and the camera / sampling / switch globals, and is the whole
// host-side work of a drop-in `renderImage`: a dynamic_cast walk over the GeoPrimitive subclasses
// (geometry.h:87-257) and the LightPrimitive subclasses (geometry.h:279-307).
//
// Used by integration/render_drt.cpp (the reference's CLI on the CUDA back end) and by oracle/ref_driver.cpp (which feeds
// the CPU oracle, the compiled reference and the CUDA path the same scene, so this walk is what every parity test runs on).
#pragma once
#include <cstring>
#include <string>
#include <vector>
#include "../include/drt.h"

namespace drt_integration {

inline void v3(double* o, const VEC3& v) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
inline VEC3 V3(const double* p) { return VEC3(p[0], p[1], p[2]); }

inline int materialTag(const std::string& m) {
  if (m == "glass") return DRT_MAT_GLASS;
  if (m == "steel") return DRT_MAT_STEEL;
  if (m == "aluminum") return DRT_MAT_ALUMINUM;
  if (m == "water") return DRT_MAT_WATER;
  if (m == "linoleum") return DRT_MAT_LINOLEUM;
  return DRT_MAT_NONE;
}
inline const char* materialName(int t) {
  switch (t) {
    case DRT_MAT_GLASS: return "glass"; case DRT_MAT_STEEL: return "steel";
    case DRT_MAT_ALUMINUM: return "aluminum"; case DRT_MAT_WATER: return "water";
    case DRT_MAT_LINOLEUM: return "linoleum"; default: return "";
  }
}
inline int modelTag(const std::string& m) {
  if (m == "oren-nayar") return DRT_MODEL_OREN_NAYAR;
  if (m == "cook-torrance") return DRT_MODEL_COOK_TORRANCE;
  if (m == "raw") return DRT_MODEL_RAW;
  return DRT_MODEL_LAMBERT;
}
inline const char* modelName(int t) {
  switch (t) {
    case DRT_MODEL_OREN_NAYAR: return "oren-nayar"; case DRT_MODEL_COOK_TORRANCE: return "cook-torrance";
    case DRT_MODEL_RAW: return "raw"; default: return "lambert";
  }
}
inline int nameTag(const std::string& n) {
  if (n == "rectangle") return DRT_NAME_RECTANGLE;
  if (n == "spherelight") return DRT_NAME_SPHERELIGHT;
  if (n == "rectanglelight") return DRT_NAME_RECTANGLELIGHT;
  return DRT_NAME_OTHER;
}


inline void commonFields(drt_prim& p, GeoPrimitive* s) {
  p.name = nameTag(s->name);
  p.material = materialTag(s->reflect_params.material);
  p.model = modelTag(s->model);
  p.flags = (s->light ? DRT_FLAG_LIGHT : 0) | (s->motion ? DRT_FLAG_MOTION : 0) |
            (s->texture ? DRT_FLAG_TEXTURE : 0) | (s->reflect_params.glossy ? DRT_FLAG_GLOSSY : 0) |
            (s->mesh ? DRT_FLAG_MESH : 0) | (s->uv_verts ? DRT_FLAG_UV_VERTS : 0);
  p.tex_frame = s->texture ? s->tex_frame : -1;
  v3(p.color, s->color); v3(p.bordercolor, s->bordercolor);
  // roughness is an uninitialised float for shapes that never set it; only the
  // Oren-Nayar / Cook-Torrance models read it (render_final_project.cpp:896,925)
  p.roughness = (p.model == DRT_MODEL_OREN_NAYAR || p.model == DRT_MODEL_COOK_TORRANCE) ? s->reflect_params.roughness : 0.0;
  p.refr[0] = s->reflect_params.refr[0]; p.refr[1] = s->reflect_params.refr[1];
  v3(p.center, s->center);
}


// payloads of the exported textures: the pointers handed out by drt_flatten_scene stay valid until its next call
inline std::vector<std::vector<uint8_t>>& tex_bytes() { static std::vector<std::vector<uint8_t>> b; return b; }

// The globals renderImage reads (render_final_project.cpp:48-138) -> drt_settings; `frame`, seed, modes are the caller's.
inline void drt_flatten_settings(drt_settings* s) {
  memset(s, 0, sizeof(*s));
  s->xRes = xRes; s->yRes = yRes;
  v3(s->eye, eye); v3(s->lookingAt, lookingAt); v3(s->up, up);
  s->aspect = aspect; s->near_plane = near; s->fov = fov; s->aperture = aperture; s->focal_length = focal_length;
  s->nogloss = nogloss; s->refr_air = refr_air; s->refr_glass = refr_glass; s->max_depth = max_depth; s->phong = phong;
  s->antialias_samples = antialias_samples; s->brdf_samples = brdf_samples; s->blur_samples = blur_samples;
  s->frame_range = frame_range; s->frame_prism = frame_prism; s->frame_cloud = frame_cloud; s->frame_blur = frame_blur;
  s->move_per_frame = move_per_frame; s->accel_t = accel_t;
  v3(s->sundir, sundir); s->perlin_cloud = perlin_cloud; s->saturation = saturation; s->clouddist = clouddist;
  s->cloudhoff = cloudhoff;
  v3(s->sun_outer, sun_outer); v3(s->sun_inner, sun_inner); v3(s->sun_core, sun_core);
  v3(s->bluesky, bluesky); v3(s->redsky, redsky);
  s->reflect = reflect;
}


// Flatten shapes / lights / textures.  `prims`, `out_lights`, `textures` must have room for shapes.size(), lights.size(),
// texture_frames.size() entries.
inline int drt_flatten_scene(drt_prim* prims, drt_light* out_lights, drt_texture* textures, std::string& err) {
  for (size_t i = 0; i < shapes.size(); i++) {
    GeoPrimitive* s = shapes[i].get();
    drt_prim& p = prims[i];
    memset(&p, 0, sizeof(p));
    commonFields(p, s);
    if (auto* c = dynamic_cast<CheckerboardWithHole*>(s)) {
      p.type = DRT_PRIM_CHECKERBOARD_HOLE;
      v3(p.A, c->A); v3(p.B, c->B); v3(p.C, c->C); v3(p.D, c->D);
      p.S = c->S; p.borderwidth = c->borderwidth; v3(p.color1, c->color1); v3(p.color2, c->color2);
      v3(p.hole[0], c->hole->A); v3(p.hole[1], c->hole->B); v3(p.hole[2], c->hole->C); v3(p.hole[3], c->hole->D);
    } else if (auto* c = dynamic_cast<Checkerboard*>(s)) {
      p.type = DRT_PRIM_CHECKERBOARD;
      v3(p.A, c->A); v3(p.B, c->B); v3(p.C, c->C); v3(p.D, c->D);
      p.S = c->S; v3(p.color1, c->color1); v3(p.color2, c->color2);
    } else if (dynamic_cast<Rectangle*>(s)) {
      p.type = DRT_PRIM_RECTANGLE;
      v3(p.A, s->A); v3(p.B, s->B); v3(p.C, s->C); v3(p.D, s->D);
    } else if (auto* c = dynamic_cast<CheckerCylinder*>(s)) {
      p.type = DRT_PRIM_CHECKER_CYLINDER;
      v3(p.c1, c->c1); v3(p.c2, c->c2); p.radius = c->radius; p.S = c->S; p.borderwidth = c->borderwidth;
    } else if (dynamic_cast<Cylinder*>(s)) {
      p.type = DRT_PRIM_CYLINDER;
      v3(p.c1, s->c1); v3(p.c2, s->c2); p.radius = s->radius;
    } else if (dynamic_cast<Sphere*>(s)) {
      p.type = DRT_PRIM_SPHERE; p.radius = s->radius;
    } else if (dynamic_cast<Triangle*>(s)) {
      p.type = DRT_PRIM_TRIANGLE;
      v3(p.A, s->A); v3(p.B, s->B); v3(p.C, s->C);
      p.uvA[0] = s->uvA[0]; p.uvA[1] = s->uvA[1]; p.uvB[0] = s->uvB[0]; p.uvB[1] = s->uvB[1];
      p.uvC[0] = s->uvC[0]; p.uvC[1] = s->uvC[1];
      if (s->mesh) v3(p.mesh_normal, s->mesh_normal);
    } else if (dynamic_cast<RectPrismV2*>(s)) {
      p.type = DRT_PRIM_RECTPRISMV2;
      v3(p.A, s->A); v3(p.B, s->B); v3(p.C, s->C); v3(p.D, s->D);
      v3(p.E, s->E); v3(p.F, s->F); v3(p.G, s->G); v3(p.H, s->H);
    } else if (dynamic_cast<RectPrism*>(s)) {
      // the slab-box prisms; the derived classes first
      v3(p.A, s->A); v3(p.B, s->B); v3(p.C, s->C); v3(p.D, s->D);
      v3(p.E, s->E); v3(p.F, s->F); v3(p.G, s->G); v3(p.H, s->H);
      auto put_hole = [&](GeoPrimitive* h) -> int {
        if (p.n_holes >= DRT_MAX_HOLES) { err = "more holes than DRT_MAX_HOLES"; return -2; }
        drt_hole& o = p.holes[p.n_holes];
        if (dynamic_cast<Cylinder*>(h)) { o.type = DRT_PRIM_CYLINDER; v3(o.c1, h->c1); v3(o.c2, h->c2); }
        else if (dynamic_cast<Sphere*>(h)) { o.type = DRT_PRIM_SPHERE; v3(o.c1, h->center); }
        else { err = "hole of a class without intersectMax"; return -2; }
        o.radius = h->radius; v3(o.color, h->color);
        p.n_holes++;
        return 0;
      };
      if (auto* c = dynamic_cast<RectPrismWithCylinder*>(s)) {
        p.type = DRT_PRIM_RECTPRISM_CYL;
        for (auto& h : c->holes) { int rc = put_hole(h.get()); if (rc) return rc; }
      } else if (auto* c = dynamic_cast<RectPrismWithHoles*>(s)) {
        p.type = DRT_PRIM_RECTPRISM_HOLES;
        for (auto& h : c->holes) { int rc = put_hole(h.get()); if (rc) return rc; }
      } else p.type = DRT_PRIM_RECTPRISM;
    } else {
      { err = "unsupported primitive class at index " + std::to_string(i) + " (" + s->name + ")"; return -2; }
    }
  }
  for (size_t i = 0; i < lights.size(); i++) {
    LightPrimitive* l = lights[i].get();
    drt_light& o = out_lights[i];
    memset(&o, 0, sizeof(o));
    o.prim_index = -1;
    v3(o.color, l->color); v3(o.center, l->center);
    shared_ptr<void> lv = dynamic_pointer_cast<void>(lights[i]);
    for (size_t k = 0; k < shapes.size(); k++)
      if (dynamic_pointer_cast<void>(shapes[k]) == lv) { o.prim_index = (int)k; break; }
    if (auto* sl = dynamic_cast<sphereLight*>(l)) {
      o.type = DRT_LIGHT_SPHERE; o.radius = sl->radius; v3(o.baxis, sl->baxis);
      v3(o.center, sl->Sphere::center);
    } else if (auto* rl = dynamic_cast<rectangleLight*>(l)) {
      o.type = DRT_LIGHT_RECT; v3(o.A, rl->A); v3(o.B, rl->B); v3(o.C, rl->C); v3(o.D, rl->D);
    } else {
      o.type = DRT_LIGHT_POINT;
    }
  }
  tex_bytes().assign(texture_frames.size(), std::vector<uint8_t>());
  for (size_t i = 0; i < texture_frames.size(); i++) {
    const std::vector<VEC3>& t = texture_frames[i];
    tex_bytes()[i].resize(t.size() * 3);
    for (size_t k = 0; k < t.size(); k++)
      for (int c = 0; c < 3; c++) tex_bytes()[i][3 * k + c] = (uint8_t)lround(t[k][c] * 255.0);
    textures[i].width = (int)texture_dims[i][0]; textures[i].height = (int)texture_dims[i][1];
    textures[i].rgb = tex_bytes()[i].data();
  }
  return 0;
}

}  // namespace drt_integration
