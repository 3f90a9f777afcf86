/*
 * drt.h -- C ABI of distraytracer-b200: the per-pixel distributed ray-tracing
 * hot path of factoryofthesun/distraytracer, rebuilt as CUDA kernels for
 * sm_100a behind plain-C entry points.
 *
 * The reference has no plugin/FFI layer.  Its narrowest stable seam is
 *     renderImage(const string& filename, const int frame,
 *                 const function<void(float)> sceneBuilder)
 *         (render_final_project.cpp:965)
 * plus the global state that call reads:
 *     shapes / lights                      render_final_project.cpp:70-71
 *     texture_frames / texture_dims        render_final_project.cpp:96-97
 *     camera + sampling + switch globals   render_final_project.cpp:48-138
 * This header is that seam with the globals made explicit: a POD scene
 * (drt_prim / drt_light / drt_texture mirror geometry.h:20-85, 279-307) and a POD
 * settings block (drt_settings mirrors the globals).  Everything renderImage does
 * between "generateBVH" (line 980) and "writePPM" (line 1220) happens behind
 * drt_render(); the caller keeps scene construction and PPM output.
 *
 * Conventions: plain pointers and sizes, no C++ / torch types, no exceptions
 * across the boundary; every function returns DRT_OK (0) or a negative
 * drt_status and records a message retrievable with drt_last_error().
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with DRT_ERR_NO_DEVICE.
 */
#ifndef DRT_H
#define DRT_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRT_ABI_VERSION 2

typedef enum drt_status {
  DRT_OK = 0,
  DRT_ERR_INVALID = -1,     /* bad argument / inconsistent description          */
  DRT_ERR_UNSUPPORTED = -2, /* primitive/feature outside the hot-path scope     */
  DRT_ERR_NO_DEVICE = -3,   /* no CUDA device (there is no CPU fallback)        */
  DRT_ERR_CUDA = -4,        /* a CUDA runtime call failed                       */
  DRT_ERR_SCENE = -5        /* the reference would `throw` (e.g. no shapes,     */
                            /* render_final_project.cpp:973-977; gaze == up,    */
                            /* render_final_project.cpp:992-996)                */
} drt_status;

/* ---- primitive type tags: the concrete GeoPrimitive subclasses ------------- */
typedef enum drt_prim_type {
  DRT_PRIM_SPHERE = 0,            /* geometry.h:87   Sphere                    */
  DRT_PRIM_CYLINDER = 1,          /* geometry.h:100  Cylinder (open, no caps)  */
  DRT_PRIM_TRIANGLE = 2,          /* geometry.h:115  Triangle                  */
  DRT_PRIM_RECTANGLE = 3,         /* geometry.h:127  Rectangle                 */
  DRT_PRIM_RECTPRISMV2 = 4,       /* geometry.h:143  RectPrismV2 (6 faces)     */
  DRT_PRIM_CHECKERBOARD = 5,      /* geometry.h:220  Checkerboard              */
  DRT_PRIM_CHECKERBOARD_HOLE = 6, /* geometry.h:232  CheckerboardWithHole      */
  DRT_PRIM_CHECKER_CYLINDER = 7,  /* geometry.h:248  CheckerCylinder           */
  /* The slab-box prisms (geometry.cpp:950-2246).  Their box is the WORLD axis-aligned bounding box of
   * the eight corners (getBounds overwrites the object-space bounds, geometry.cpp:987-988), so only
   * axis-aligned prisms look like prisms -- reproduced as is. */
  DRT_PRIM_RECTPRISM = 8,         /* geometry.h:159  RectPrism                 */
  DRT_PRIM_RECTPRISM_CYL = 9,     /* geometry.h:182  RectPrismWithCylinder: drt_prim.holes are Cylinders   */
  DRT_PRIM_RECTPRISM_HOLES = 10,  /* geometry.h:200  RectPrismWithHoles: holes are Spheres / Cylinders    */
  DRT_PRIM_TYPE_COUNT = 11
} drt_prim_type;

#define DRT_MAX_HOLES 4
/* One entry of RectPrismWithCylinder::holes / RectPrismWithHoles::holes (geometry.h:195, 216). */
typedef struct drt_hole {
  int32_t type;      /* DRT_PRIM_SPHERE or DRT_PRIM_CYLINDER */
  int32_t pad_;
  double c1[3];      /* cylinder end point, or the sphere's centre */
  double c2[3];      /* cylinder end point */
  double radius;
  double color[3];   /* the hole's own colour (a hit on the hole's wall is shaded with it) */
} drt_hole;

/* GeoPrimitive::name (geometry.h:46).  Only these values change behaviour:
 * "spherelight"/"rectanglelight" select the emissive formula
 * (render_final_project.cpp:777,783) and "rectangle" selects the motion-blur
 * translation (render_final_project.cpp:1116,1142). */
typedef enum drt_name {
  DRT_NAME_OTHER = 0,
  DRT_NAME_RECTANGLE = 1,
  DRT_NAME_SPHERELIGHT = 2,
  DRT_NAME_RECTANGLELIGHT = 3
} drt_name;

/* Reflectance::material (geometry.h:21).  refl_materials =
 * {glass, steel, aluminum, water, linoleum} (render_final_project.cpp:64). */
typedef enum drt_material {
  DRT_MAT_NONE = 0, /* "" or any string not in refl_materials (e.g. "marble") */
  DRT_MAT_GLASS = 1,
  DRT_MAT_STEEL = 2,
  DRT_MAT_ALUMINUM = 3,
  DRT_MAT_WATER = 4,
  DRT_MAT_LINOLEUM = 5
} drt_material;

/* GeoPrimitive::model (geometry.h:47; switch at render_final_project.cpp:894-948). */
typedef enum drt_model {
  DRT_MODEL_LAMBERT = 0, /* anything else: Lambert + Phong(phong) */
  DRT_MODEL_OREN_NAYAR = 1,
  DRT_MODEL_COOK_TORRANCE = 2,
  DRT_MODEL_RAW = 3
} drt_model;

enum drt_prim_flags {
  DRT_FLAG_LIGHT = 1 << 0,    /* GeoPrimitive::light   (geometry.h:40)          */
  DRT_FLAG_MOTION = 1 << 1,   /* GeoPrimitive::motion  (geometry.h:41)          */
  DRT_FLAG_TEXTURE = 1 << 2,  /* GeoPrimitive::texture (geometry.h:48)          */
  DRT_FLAG_GLOSSY = 1 << 3,   /* Reflectance::glossy   (geometry.h:23)          */
  DRT_FLAG_MESH = 1 << 4,     /* GeoPrimitive::mesh    (geometry.h:56)          */
  DRT_FLAG_UV_VERTS = 1 << 5, /* GeoPrimitive::uv_verts(geometry.h:42)          */
  DRT_FLAG_VERTEX_MOTION = 1 << 6 /* DRT_BLUR_VELOCITY, cylinders only: the two end points move independently --
                               * c1 by `velocity`, c2 by `velocity2` -- and the axis is re-derived per time sample: a
                               * bone between its poses at frame and frame + 1 */
};

/* One GeoPrimitive, flattened.  Doubles because the reference's host surface is
 * double (SETTINGS.h:13); the library repacks to 16-byte-aligned float SoA on
 * upload.  Fields a type does not use are ignored. */
typedef struct drt_prim {
  int32_t type;      /* drt_prim_type */
  int32_t name;      /* drt_name      */
  int32_t material;  /* drt_material  */
  int32_t model;     /* drt_model     */
  int32_t flags;     /* drt_prim_flags */
  int32_t tex_frame; /* index into textures, used when DRT_FLAG_TEXTURE */
  double color[3];
  double bordercolor[3];
  double roughness;  /* Reflectance::roughness */
  double refr[2];    /* Reflectance::refr (n, k) */
  double center[3];
  double radius;
  double A[3], B[3], C[3], D[3]; /* triangle: A,B,C; rectangle-like: A,B,C,D */
  double E[3], F[3], G[3], H[3]; /* RectPrismV2 second face                 */
  double c1[3], c2[3];           /* cylinder end points (axis = (c2-c1)^)    */
  double uvA[2], uvB[2], uvC[2]; /* triangle per-vertex UV                   */
  double mesh_normal[3];
  double S;                      /* checker square side                      */
  double borderwidth;
  double color1[3], color2[3];   /* checker colours                          */
  double hole[4][3];             /* CheckerboardWithHole::hole  A,B,C,D      */
  /* Linear motion for time sampling: position at time s in [0,frame_range) is
   * p + s * velocity (every vertex).  Zero for static primitives.  This is the
   * generalisation of the reference's two motion semantics named in
   * SURVEY.md 8(f)1; the reference's own "rectangle" translation is selected by
   * drt_settings.blur_mode instead and does not read this field. */
  double velocity[3];
  /* RectPrismWithCylinder / RectPrismWithHoles only (ABI version 2) */
  int32_t n_holes;
  int32_t pad_;
  drt_hole holes[DRT_MAX_HOLES];
  double velocity2[3];           /* DRT_FLAG_VERTEX_MOTION: velocity of c2 (`velocity` is then c1's) */
} drt_prim;

typedef enum drt_light_type {
  DRT_LIGHT_POINT = 0,  /* geometry.h:287 pointLight     */
  DRT_LIGHT_SPHERE = 1, /* geometry.h:294 sphereLight    */
  DRT_LIGHT_RECT = 2    /* geometry.h:302 rectangleLight */
} drt_light_type;

typedef struct drt_light {
  int32_t type;       /* drt_light_type */
  int32_t prim_index; /* index in prims of the SAME object (area lights are both a
                       * light and a shape and never shadow themselves,
                       * render_final_project.cpp:832-837); -1 for point lights */
  double color[3];
  double center[3];
  double radius;      /* sphere light */
  double baxis[3];    /* sphereLight::baxis (geometry.h:299) */
  double A[3], B[3], C[3], D[3]; /* rectangle light */
} drt_light;

/* One entry of texture_frames/texture_dims (helpers.h:92-113): 8-bit RGB, row
 * major, as stbi_load returns it.  Texel value = byte/255. */
typedef struct drt_texture {
  int32_t width, height;
  const uint8_t* rgb; /* width*height*3 bytes */
} drt_texture;

/* Optional indexed triangle mesh (objHelper.h:6-85 output shape): when given,
 * its triangles are appended after `prims` and traversed through the
 * device-built LBVH.  Vertices are single precision (what tiny_obj_loader hands the reference, objHelper.h:40-46; the
 * reference then transforms them in double, scene.h:297-307 -- a mesh that needs that is transformed by the caller).
 * Materials: every triangle uses `material`, unless `materials` is given -- then triangle t uses
 * materials[material_ids[t]].  That is how the reference's per-face roughness (one value per Triangle, looked up in a
 * roughness map, scene.h:372-378: at most 766 distinct values) travels without a material record per triangle.
 * Motion (DRT_BLUR_VELOCITY): a mesh moves as a whole -- DRT_FLAG_MOTION and `velocity` of its material(s), which must
 * agree; the traversal moves the ray instead of the triangles. */
typedef struct drt_mesh {
  int64_t n_vertices, n_triangles;
  const float* vertices;   /* 3 floats per vertex               */
  const int32_t* indices;  /* 3 vertex indices per triangle     */
  const float* texcoords;  /* 2 floats per vertex, or NULL      */
  drt_prim material;       /* type must be DRT_PRIM_TRIANGLE; A/B/C ignored */
  int32_t n_materials;     /* 0: `material` for all triangles; else 1..65535 entries in `materials` */
  int32_t pad_;
  const drt_prim* materials;
  const int32_t* material_ids;  /* n_triangles entries in [0, n_materials) */
} drt_mesh;

typedef struct drt_scene_desc {
  int32_t abi_version; /* DRT_ABI_VERSION */
  int32_t n_prims;
  const drt_prim* prims;
  int32_t n_lights;
  const drt_light* lights;
  int32_t n_textures;
  const drt_texture* textures;
  const drt_mesh* mesh; /* may be NULL */
} drt_scene_desc;

typedef enum drt_sample_mode {
  /* Counter-based per-pixel RNG keyed by (seed,pixel,sample,path,purpose). The
   * same keys drive oracle/ so both sides see identical sample positions. */
  DRT_SAMPLES_KEYED = 0
} drt_sample_mode;

typedef enum drt_precision {
  /* double vectors + float scalars, expression by expression like the reference
   * (SETTINGS.h:13 Real=double; geometry.cpp keeps its scalar temporaries in
   * float).  B200's FP64 pipe runs at half the FP32 rate, which makes this the
   * default: it is what parity is asserted on. */
  DRT_PRECISION_REFERENCE = 0,
  /* fp32 vectors and scalars */
  DRT_PRECISION_FP32 = 1
} drt_precision;

typedef enum drt_blur_mode {
  DRT_BLUR_REFERENCE = 0, /* render_final_project.cpp:1095-1210: re-trace blur_samples
                           * times; primitives named "rectangle" move in y when
                           * frame >= frame_prism */
  DRT_BLUR_VELOCITY = 1   /* re-trace with every primitive displaced by
                           * velocity * (u * frame_range)   */
} drt_blur_mode;

/* The globals renderImage reads (render_final_project.cpp:48-138). */
typedef struct drt_settings {
  int32_t xRes, yRes;          /* :48-49 */
  double eye[3];               /* :52 */
  double lookingAt[3];         /* :53 */
  double up[3];                /* :54 */
  float aspect;                /* :55  NOT recomputed from xRes/yRes (quirk Q2) */
  float near_plane;            /* :56  `near` */
  float fov;                   /* :57 */
  float aperture;              /* :58 */
  float focal_length;          /* :59 */
  int32_t nogloss;             /* :61 */
  float refr_air, refr_glass;  /* :65-66 */
  int32_t max_depth;           /* :67 */
  float phong;                 /* :72 */
  int32_t antialias_samples;   /* :81  spp = int(sqrt(.))^2 (:1046,1061) */
  int32_t brdf_samples;        /* :82 */
  int32_t blur_samples;        /* :83 */
  int32_t frame_range;         /* :84 */
  int32_t frame_prism;         /* :112 */
  int32_t frame_cloud;         /* :113 */
  int32_t frame_blur;          /* :114 */
  float move_per_frame;        /* :121 */
  float accel_t;               /* :123 */
  double sundir[3];            /* :127 */
  int32_t perlin_cloud;        /* :128 */
  float saturation;            /* :129 */
  float clouddist;             /* :130 */
  float cloudhoff;             /* :131 */
  double sun_outer[3], sun_inner[3], sun_core[3], bluesky[3], redsky[3]; /* :132-136 */
  int32_t reflect;             /* :138 */
  int32_t frame;               /* renderImage's `frame` argument */
  uint32_t seed;               /* sample-stream seed */
  int32_t sample_mode;         /* drt_sample_mode */
  int32_t blur_mode;           /* drt_blur_mode */
  int32_t cloud_only;          /* 1 = renderImageCloud (:1224-1279): noise-only frame */
  int32_t precision;           /* drt_precision */
} drt_settings;

/* A rectangle of pixels of the xRes*yRes frame, rendered on one device.  x0,y0
 * are in the reference's loop coordinates (y = 0 is the BOTTOM image row,
 * render_final_project.cpp:1031,1215). */
typedef struct drt_tile {
  int32_t x0, y0, width, height;
  int32_t device; /* CUDA device ordinal */
} drt_tile;

/* Event counters and timing for roofline accounting.  Counters
 * are filled only when `collect` is non-zero (they cost atomics). */
typedef struct drt_counters {
  int32_t collect;
  float kernel_ms;        /* CUDA-event time of the render kernels on their stream */
  int32_t kernel_launches;
  int32_t kernel_variant; /* feature mask of the render_wave instantiation that ran (csrc/drt_launch.h WaveFeat), -1: none */
  uint64_t samples;       /* camera samples (primary rays incl. blur re-traces are separate) */
  uint64_t rays;          /* rayColor invocations that traversed (primary+secondary) */
  uint64_t shadow_rays;
  uint64_t node_tests;    /* AABB slab tests */
  uint64_t prim_tests[DRT_PRIM_TYPE_COUNT];
  uint64_t shade_evals;   /* BRDF evaluations (per unoccluded light) */
  uint64_t noise_evals;   /* ValueNoise_3D evaluations */
} drt_counters;

/* A scene handle owns one CUDA stream, its scratch buffers and its timing events: calls on the SAME handle (render,
 * update, pose) must not overlap -- drive one handle from one thread at a time.  Different handles, on the same or on
 * different devices, are independent. */
typedef struct drt_scene drt_scene;

/* Number of usable CUDA devices (0 if none; never negative). */
int drt_device_count(void);

/* Default values of the globals at render_final_project.cpp:48-138. */
void drt_settings_default(drt_settings* s);
/* Zero-filled primitive with the GeoPrimitive member defaults (geometry.h:39-57). */
void drt_prim_default(drt_prim* p);

/* Validate + repack the scene into device SoA buffers on `device`, upload the
 * textures, and build the BVH there.  Replaces generateBVH
 * (render_final_project.cpp:980, helpers.h:381-472).  The scene is immutable
 * afterwards except through drt_scene_update_prims.
 * Size: any number of analytic primitives (up to 256 intersectable pieces are culled by a
 * filter table in shared memory, larger scenes through a tree over the pieces) plus one
 * optional triangle mesh of up to 2^30 triangles.  Device memory: the scene tables, and per
 * render call the sample records (16 bytes per camera sample of a launch, at most 8 GiB) and
 * the ray pools of the persistent kernel (about 25 MB per SM at max_depth 10, brdf_samples 2). */
int drt_scene_create(const drt_scene_desc* desc, int device, drt_scene** out);
/* Replace the analytic primitives in place (same count and types), e.g. the
 * re-posed bone cylinders of the next mocap frame (scene.h:637-659). */
int drt_scene_update_prims(drt_scene* scene, const drt_prim* prims, int32_t n_prims);
/* Replace the lights (any count), e.g. a light a scene builder moves with the
 * frame number (scene.h:3690-3692).  Area lights keep pointing at their shapes
 * through prim_index. */
int drt_scene_update_lights(drt_scene* scene, const drt_light* lights, int32_t n_lights);
void drt_scene_destroy(drt_scene* scene);

/* ---- mocap skeletons (BASELINE config 4) ------------------------------------------------
 * ASF skeleton + AMC motion -> bone cylinders, for every frame of the clip at once, on the
 * device.  Replaces the host-side chain the reference runs per rendered frame:
 *   Skeleton(asf, scale) / Motion(amc, scale, skeleton)    skeleton.cpp:545-590, motion.cpp:28-38
 *   setSkeletonsToSpecifiedFrame(frame)                    scene.h:109-128
 *   DisplaySkeleton::ComputeBonePositions                  displaySkeleton.cpp:229-270
 *   rotations/scalings/translations/lengths -> end points  scene.h:616-659
 * `scale` is MOCAP_SCALE (types.h:6, 0.06).  The table of end points stays in HBM; results
 * are bit-identical to the reference's. */
typedef struct drt_skeleton drt_skeleton;

int drt_skeleton_create(const char* asf_text, size_t asf_len, const char* amc_text, size_t amc_len,
                        double scale, int device, drt_skeleton** out);
/* Same, reading the two files (the reference opens "90.asf" / "90_16_v3.amc",
 * render_final_project.cpp main). */
int drt_skeleton_load(const char* asf_path, const char* amc_path, double scale, int device, drt_skeleton** out);
/* n_cylinders = bones without the root (the reference skips bone 0, scene.h:631-633);
 * fk_ms = device time of the forward-kinematics kernel over the whole clip.  Any output may be NULL. */
int drt_skeleton_info(const drt_skeleton* skel, int32_t* n_cylinders, int32_t* n_frames, float* fk_ms);
/* Copies frames [frame0, frame0+n_frames) of the device table to `out`:
 * n_frames * n_cylinders * 6 doubles (left x,y,z, right x,y,z per bone). */
int drt_skeleton_bones(const drt_skeleton* skel, int32_t frame0, int32_t n_frames, double* out);
/* Re-pose the n_cylinders DRT_PRIM_CYLINDER primitives prims[first_prim ...] of `scene` to mocap
 * frame `frame` (clamped to the last frame like scene.h:117-121; negative is an error like
 * scene.h:111-115) and re-upload the scene.  Every end point is lowered by `drop_y`
 * (scene.h:646-650 drops the figure by frame-frame_cloud).  set_velocity (DRT_BLUR_VELOCITY):
 *   0  static bones;
 *   1  one translation per bone: velocity = displacement of the bone's midpoint to frame + 1;
 *   2  two poses: c1 and c2 each move to their own position at frame + 1 (DRT_FLAG_VERTEX_MOTION), so a rotating
 *      bone sweeps the pose in between. */
int drt_scene_pose_skeleton(drt_scene* scene, const drt_skeleton* skel, int32_t frame, int32_t first_prim,
                            double drop_y, int32_t set_velocity);
void drt_skeleton_destroy(drt_skeleton* skel);

/* Render `tile` of the frame into `out_rgb`: a tightly packed HOST buffer of
 * tile.width*tile.height*3 bytes laid out like the PPM payload writePPM emits
 * (helpers.h:174-195): rows top to bottom (row 0 of the buffer is loop row
 * y0+height-1, render_final_project.cpp:1215), RGB, value = (unsigned
 * char)(clamp(c)*255).  Includes the device->host copy.  `counters` may be NULL. */
int drt_render(const drt_scene* scene, const drt_settings* settings, const drt_tile* tile,
               uint8_t* out_rgb, drt_counters* counters);

/* Same, but also (or only) returns the un-quantised float image the reference
 * holds in ppmOut before writePPM truncates it (values in [0,255]).  Either
 * output may be NULL. */
int drt_render_float(const drt_scene* scene, const drt_settings* settings, const drt_tile* tile,
                     float* out_rgb_f32, uint8_t* out_rgb, drt_counters* counters);

/* Render into the scene's device-resident frame buffer only (no host copy):
 * the kernel-only path bench.py times for `value`. */
int drt_render_device(const drt_scene* scene, const drt_settings* settings, const drt_tile* tile,
                      drt_counters* counters);

/* Render ONE frame (or tile) on several devices at once: the single-frame partition of BASELINE configs 3 and 5
 * (SURVEY.md 8e).  `scenes` are n_scenes handles holding the SAME scene, normally one per device; scenes[0]'s device
 * gathers.  One persistent kernel per device claims pixel-aligned units of ~1024 camera samples from ONE counter in the
 * gathering device's memory (system-scope atomics over NVLink / NVSwitch peer access), so the devices balance at unit
 * granularity whatever the cost profile of the frame; every device resolves the pixels it rendered straight into the
 * gathering device's frame (peer stores), and one device-to-host copy delivers `out_rgb` (layout as drt_render).  No
 * collective.  The image is byte-identical to drt_render's on one device.  `tile->device` is ignored; `counters` is NULL
 * or an array of n_scenes entries (kernel_ms = each device's own kernel time).  Needs peer access from every device to
 * scenes[0]'s (DRT_ERR_UNSUPPORTED otherwise: cut the frame into tiles and call drt_render per device instead). */
int drt_render_multi(drt_scene* const* scenes, int32_t n_scenes, const drt_settings* settings, const drt_tile* tile,
                     uint8_t* out_rgb, drt_counters* counters);

/* Writes a binary P6 PPM exactly as helpers.h:174-195 does. */
int drt_write_ppm(const char* filename, int32_t width, int32_t height, const uint8_t* rgb);

/* Thread-local message for the last non-OK status. */
const char* drt_last_error(void);

/* ---- diagnostics (used by the test-suite, harmless in production) ----------- */
/* sizeof() of drt_prim, drt_light, drt_scene_desc, drt_settings, drt_tile,
 * drt_counters, for language bindings to verify their struct layout. */
void drt_abi_sizes(int32_t* out6);
/* One uniform of the keyed sample stream: key = child `child` of the root path of
 * camera sample `sample` of pixel `pixel`; evaluated on the host from the same
 * inline functions the kernels use. */
float drt_debug_rng(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t child, uint32_t dim);
/* The keyed form of the reference's per-pixel lens-sample shuffle (helpers.h:270-279): swap target j of step i, and the
 * index of the lens point the shuffle leaves at position `sample` of a pixel's n_lens points. */
int drt_debug_shuffle_j(uint32_t seed, uint32_t pixel, int32_t i);
int drt_debug_lens_index(uint32_t seed, uint32_t pixel, int32_t sample, int32_t n_lens);
/* Primitive indices in the candidate order the library derives from its host-side replay of the
 * reference's generateBVH (helpers.h:381-472); ties of t between shapes resolve in this order. */
int drt_debug_candidate_order(const drt_prim* prims, int32_t n_prims, int32_t* out, int32_t cap);
/* Host half of drt_skeleton_create only (needs no device): bone count including the root, frame count
 * ((non-empty lines - 3) / (moving bones + 1), motion.cpp:113-121), and per bone (up to `cap`) the parent index
 * (-1 for the root) and the degrees of freedom as bits rx ry rz tx ty tz = 1 2 4 8 16 32 (skeleton.cpp:207-231). */
int drt_debug_skeleton_parse(const char* asf_text, size_t asf_len, const char* amc_text, size_t amc_len, double scale,
                             int32_t* n_bones, int32_t* n_frames, int32_t* parents, int32_t* dofs, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* DRT_H */
