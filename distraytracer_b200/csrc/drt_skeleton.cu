// drt_skeleton -- ASF/AMC mocap ingest and forward kinematics on the device.
//
// Replaces, for the bone cylinders of the mocap scenes (BASELINE config 4):
//   Skeleton::Skeleton + readASFfile                 skeleton.cpp:118-293, 545-590
//   Motion::readAMCfile                              motion.cpp:92-213
//   setSkeletonsToSpecifiedFrame + setPosture        scene.h:109-128, skeleton.cpp:502-528
//   DisplaySkeleton::ComputeBonePositions / Traverse / DrawBone   displaySkeleton.cpp:117-270
//   rotation/scaling/translation -> cylinder ends    scene.h:616-659
//
// Design: the reference re-runs a recursive matrix-stack walk on the host for every frame
// it renders.  Here the clip is posed ONCE, for all frames at a time: one thread per
// (frame, bone) multiplies its own ancestor chain root -> bone (depth <= 32) and writes
// the two cylinder end points into a [frame][bone][6] table that stays in HBM; per-frame
// re-posing of a scene (drt_scene_pose_skeleton) is then a table lookup.
//   * per-bone constants (BoneK): the parent-to-bone frame change, the canonical-z ->
//     bone-direction draw rotation (with the aspect scaling folded in), the float-narrowed
//     child offset and length -- computed once at parse time;
//   * per-(frame, bone) inputs (PoseK): the float-narrowed AMC translations and the
//     sin/cos of the three float-narrowed joint angles.  The host evaluates those with the
//     same libm the reference calls, the kernel does every multiply and add of the chain in
//     the reference's order (this unit is compiled with -fmad=false), so the table is
//     bit-identical to the reference's bones, not merely close.
#include <cuda_runtime.h>

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/drt.h"
#include "drt_skeleton.h"

namespace drt {

namespace {

constexpr int kMaxBones = 256;    // types.h:9
constexpr int kMaxDepth = 32;
constexpr double kPi = 3.14159265358979323846;

enum DofBits { DOF_RX = 1, DOF_RY = 2, DOF_RZ = 4, DOF_TX = 8, DOF_TY = 16, DOF_TZ = 32 };

struct BoneK {            // device, per bone
  double A[16];           // rot_parent_current as glMultMatrixd reads it: A(x,y) = m[4x+y]
  double D[16];           // scaling * toMatrix4(rotation^T) of DrawBone (non-root)
  float off[3];           // float(dir * length): the glTranslatef that ends DrawBone
  float len;              // boneLengths[] is a vector<float>
  int parent, dof;
};
struct PoseK {            // device, per (frame, bone)
  double sc[6];           // sin, cos of the float-narrowed radians of rx, ry, rz
  float t[3];             // float(tx), float(ty), float(tz)
  float pad;
};

// ---- 4x4 helpers shared by host precompute and the kernel -----------------------------
// Eigen's fixed-size product reduces each dot product as (p0+p1)+(p2+p3).
__host__ __device__ inline void mul4(const double* a, const double* b, double* r) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      r[4 * i + j] = (a[4 * i] * b[j] + a[4 * i + 1] * b[4 + j]) + (a[4 * i + 2] * b[8 + j] + a[4 * i + 3] * b[12 + j]);
}
__host__ __device__ inline void ident(double* m) { for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
// AngleAxis(angle, axis).toRotationMatrix(), transposed, widened to 4x4 (displaySkeleton.cpp:23-30, 50-61)
__host__ __device__ inline void angleAxisT4(double sn, double c, const double* ax, double* r) {
  const double sa0 = sn * ax[0], sa1 = sn * ax[1], sa2 = sn * ax[2];
  const double k0 = (1.0 - c) * ax[0], k1 = (1.0 - c) * ax[1], k2 = (1.0 - c) * ax[2];
  ident(r);
  double t;
  t = k0 * ax[1]; r[4 * 1 + 0] = t - sa2; r[4 * 0 + 1] = t + sa2;      // res(0,1) -> r(1,0), res(1,0) -> r(0,1)
  t = k0 * ax[2]; r[4 * 2 + 0] = t + sa1; r[4 * 0 + 2] = t - sa1;
  t = k1 * ax[2]; r[4 * 2 + 1] = t - sa0; r[4 * 1 + 2] = t + sa0;
  r[0] = k0 * ax[0] + c; r[5] = k1 * ax[1] + c; r[10] = k2 * ax[2] + c;
}
__host__ __device__ inline void translateInto(double* cur, float x, float y, float z) {   // myTranslatef
  double t[16], r[16]; ident(t); t[12] = x; t[13] = y; t[14] = z;
  mul4(t, cur, r);
  for (int i = 0; i < 16; i++) cur[i] = r[i];
}
__host__ __device__ inline void rotateInto(double* cur, double sn, double c, int axis) {   // myRotatef about a unit axis
  double ax[3] = {0.0, 0.0, 0.0}; ax[axis] = 1.0;
  double m[16], r[16];
  angleAxisT4(sn, c, ax, m);
  mul4(m, cur, r);
  for (int i = 0; i < 16; i++) cur[i] = r[i];
}

// One thread per (frame, bone >= 1): displaySkeleton.cpp:117-270 along the bone's ancestor
// chain, then scene.h:637-644.
__global__ void skeleton_fk(const BoneK* __restrict__ bones, const PoseK* __restrict__ poses, int n_bones, int n_frames,
                            double* __restrict__ out /* [frame][n_bones-1][6] */) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nb1 = n_bones - 1;
  if (gid >= (long long)n_frames * nb1) return;
  const int frame = (int)(gid / nb1), bone = 1 + (int)(gid % nb1);
  int chain[kMaxDepth], depth = 0;
  for (int b = bone; b >= 0; b = bones[b].parent) chain[depth++] = b;
  double cur[16], tmp[16];
  ident(cur);
  // ComputeBonePositions: the skeleton's own offset and orientation (always zero in the reference)
  translateInto(cur, 0.0f, 0.0f, 0.0f);
  rotateInto(cur, 0.0, 1.0, 0); rotateInto(cur, 0.0, 1.0, 1); rotateInto(cur, 0.0, 1.0, 2);
  for (int d = depth - 1; d >= 0; d--) {
    const int b = chain[d];
    const BoneK& K = bones[b];
    const PoseK& P = poses[(size_t)frame * n_bones + b];
    mul4(K.A, cur, tmp);
    for (int i = 0; i < 16; i++) cur[i] = tmp[i];
    if (K.dof & DOF_TZ) translateInto(cur, 0.0f, 0.0f, P.t[2]);
    if (K.dof & DOF_TY) translateInto(cur, 0.0f, P.t[1], 0.0f);
    if (K.dof & DOF_TX) translateInto(cur, P.t[0], 0.0f, 0.0f);
    if (K.dof & DOF_RZ) rotateInto(cur, P.sc[4], P.sc[5], 2);
    if (K.dof & DOF_RY) rotateInto(cur, P.sc[2], P.sc[3], 1);
    if (K.dof & DOF_RX) rotateInto(cur, P.sc[0], P.sc[1], 0);
    if (d == 0) {
      // currentTransform = scaling * rotation * currentMatrix; its last row is the translation, the
      // rest (transposed, last row zeroed) the "rotation"; scaling is applied again on the way out
      double T[16];
      mul4(K.D, cur, T);
      const double tr[3] = {T[12], T[13], T[14]};
      T[12] = 0; T[13] = 0; T[14] = 0;
      double Rt[16], S[16], RS[16];
      for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Rt[4 * i + j] = T[4 * j + i];
      ident(S); S[0] = (b == 0) ? 1.0 : 0.25; S[5] = S[0];          // set_bone_shape, skeleton.cpp:531-543
      mul4(Rt, S, RS);
      const double lv[4] = {0, 0, 0, 1}, rv[4] = {0, 0, (double)K.len, 1};
      double* o = out + ((size_t)frame * nb1 + (bone - 1)) * 6;
      for (int k = 0; k < 3; k++) {
        const double l = (RS[4 * k] * lv[0] + RS[4 * k + 1] * lv[1]) + (RS[4 * k + 2] * lv[2] + RS[4 * k + 3] * lv[3]);
        const double r = (RS[4 * k] * rv[0] + RS[4 * k + 1] * rv[1]) + (RS[4 * k + 2] * rv[2] + RS[4 * k + 3] * rv[3]);
        o[k] = l + tr[k];
        o[3 + k] = r + tr[k];
      }
    } else {
      translateInto(cur, K.off[0], K.off[1], K.off[2]);
    }
  }
}

// ---- host: parse ----------------------------------------------------------------------
struct BoneH {
  std::string name;
  int parent = -1, dof = 0, order[8] = {0};
  bool rot[3] = {false, false, false}, trans[3] = {false, false, false};
  double dir[3] = {0, 0, 0}, axis[3] = {0, 0, 0}, length = 0;
};

struct Tokens {           // lines -> words; the reference splits on ' ' for dof / hierarchy lines
  std::vector<std::vector<std::string>> rows;
  explicit Tokens(const char* text, size_t len) {
    std::string line;
    auto flush = [&]() {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      std::vector<std::string> w; std::string cur;
      for (char ch : line) { if (ch == ' ' || ch == '\t' || ch == '\r') { if (!cur.empty()) w.push_back(cur); cur.clear(); } else cur.push_back(ch); }
      if (!cur.empty()) w.push_back(cur);
      rows.push_back(w); line.clear();
    };
    for (size_t i = 0; i < len; i++) { if (text[i] == '\n') flush(); else line.push_back(text[i]); }
    if (!line.empty()) flush();
  }
};

double num(const std::vector<std::string>& w, size_t i, double keep) { return i < w.size() ? strtod(w[i].c_str(), nullptr) : keep; }

void rotAxis(int axis, double angle, double* R) {          // RotationX/Y/Z, skeleton.cpp:19-56
  const float th = (float)angle;
  const float c = (float)cos((double)th), s = (float)sin((double)th);
  for (int i = 0; i < 16; i++) R[i] = 0;
  R[15] = 1; R[5 * axis] = 1;
  const int a = (axis + 1) % 3, b = (axis + 2) % 3;         // the rotated plane, right-handed
  R[5 * a] = c; R[5 * b] = c; R[4 * a + b] = -s; R[4 * b + a] = s;
}
void euler(const double* axis_deg, bool inverse, double* out) {
  // inverse: Rx(-x) Ry(-y) Rz(-z)   (world -> bone);   else: Rz(z) Ry(y) Rx(x)   (bone -> world)
  double X[16], Y[16], Z[16], t[16];
  const double sg = inverse ? -1.0 : 1.0;
  rotAxis(0, sg * axis_deg[0] * kPi / 180.0, X); rotAxis(1, sg * axis_deg[1] * kPi / 180.0, Y); rotAxis(2, sg * axis_deg[2] * kPi / 180.0, Z);
  if (inverse) { mul4(X, Y, t); mul4(t, Z, out); } else { mul4(Z, Y, t); mul4(t, X, out); }
}

int parseAsf(const char* text, size_t len, double scale, std::vector<BoneH>& bones, std::string& err) {
  Tokens T(text, len);
  size_t r = 0;
  std::string key;
  auto keyword = [&](size_t row) { if (!T.rows[row].empty()) key = T.rows[row][0]; return key; };   // sticky over blank lines (sscanf)
  while (r < T.rows.size() && keyword(r) != ":bonedata") r++;
  if (r >= T.rows.size()) { err = "ASF: no :bonedata section"; return DRT_ERR_INVALID; }
  r += 2;                                                   // the section line and the first "begin"
  bones.clear();
  BoneH root; root.name = "root"; root.length = 0.05; root.dof = 6;               // skeleton.cpp:547-569
  const int root_order[7] = {4, 5, 6, 1, 2, 3, 0};
  for (int i = 0; i < 7; i++) root.order[i] = root_order[i];
  for (int a = 0; a < 3; a++) root.rot[a] = root.trans[a] = true;
  bones.push_back(root);
  double length = 0;                                        // carried from bone to bone like the reference's local
  bool in_hierarchy = false;
  while (!in_hierarchy) {
    if ((int)bones.size() >= kMaxBones) { err = "ASF: too many bones"; return DRT_ERR_UNSUPPORTED; }
    BoneH b;
    bool closed = false;
    for (; r < T.rows.size(); r++) {
      const std::vector<std::string>& w = T.rows[r];
      const std::string k = keyword(r);
      if (k == "end") { closed = true; r++; break; }
      if (k == ":hierarchy") { in_hierarchy = true; r++; break; }
      if (w.empty()) { if (k == "dof") b.dof = 0; continue; }
      if (k == "name" && w.size() > 1) b.name = w[1];
      else if (k == "direction") for (int a = 0; a < 3; a++) b.dir[a] = num(w, 1 + a, b.dir[a]);
      else if (k == "length") length = num(w, 1, length);
      else if (k == "axis") for (int a = 0; a < 3; a++) b.axis[a] = num(w, 1 + a, b.axis[a]);
      else if (k == "dof") {
        b.dof = 0;
        for (size_t i = 1; i < w.size(); i++) {
          static const char* names[7] = {"rx", "ry", "rz", "tx", "ty", "tz", "l"};
          int code = 0;
          for (int c = 0; c < 7; c++) if (w[i] == names[c]) code = c + 1;
          if (code >= 1 && code <= 3) b.rot[code - 1] = true;
          if (code >= 4 && code <= 6) b.trans[code - 4] = true;
          if (b.dof < 7) { b.order[b.dof++] = code; b.order[b.dof] = 0; }
        }
      }
    }
    if (!closed && !in_hierarchy) { err = "ASF: truncated :bonedata"; return DRT_ERR_INVALID; }
    if (closed) { b.length = length * scale; bones.push_back(b); }
  }
  r++;                                                      // "begin"
  for (; r < T.rows.size(); r++) {
    const std::vector<std::string>& w = T.rows[r];
    if (keyword(r) == "end") break;
    int parent = -1;
    for (size_t j = 0; j < w.size(); j++) {
      int idx = -1;
      for (size_t i = 0; i < bones.size(); i++) if (bones[i].name == w[j]) { idx = (int)i; break; }
      if (idx < 0) { err = "ASF: hierarchy names unknown bone '" + w[j] + "'"; return DRT_ERR_INVALID; }
      if (j == 0) parent = idx;
      else if (idx == 0 || bones[idx].parent >= 0) { err = "ASF: bone '" + w[j] + "' has two parents"; return DRT_ERR_INVALID; }
      else bones[idx].parent = parent;
    }
  }
  // every bone must hang off the root through a chain of bounded depth (the reference's recursive
  // walk simply never reaches detached bones; a table of cylinders needs all of them)
  for (size_t i = 1; i < bones.size(); i++) {
    int d = 0, b = (int)i;
    while (b > 0 && d <= kMaxDepth) { b = bones[b].parent; d++; }
    if (b != 0) { err = "ASF: bone '" + bones[i].name + "' is not connected to the root"; return DRT_ERR_INVALID; }
    if (d >= kMaxDepth) { err = "ASF: hierarchy deeper than 32"; return DRT_ERR_UNSUPPORTED; }
  }
  return DRT_OK;
}

// motion.cpp:92-213.  rot/trans: [frame][bone][3]
int parseAmc(const char* text, size_t len, double scale, std::vector<BoneH>& bones, int& n_frames,
             std::vector<double>& rot, std::vector<double>& trans, std::string& err) {
  int lines = 0;
  { size_t start = 0; for (size_t i = 0; i < len; i++) if (text[i] == '\n') { if (i > start) lines++; start = i + 1; } }
  int moving = 0;
  for (const BoneH& b : bones) if (b.dof > 0) moving++;
  n_frames = (lines - 3) / (moving + 1);
  if (n_frames < 1) { err = "AMC: no frames"; return DRT_ERR_INVALID; }
  const size_t nb = bones.size();
  rot.assign((size_t)n_frames * nb * 3, 0.0); trans.assign((size_t)n_frames * nb * 3, 0.0);
  const char* p = text; const char* end = text + len;
  auto word = [&](std::string& w) {
    while (p < end && isspace((unsigned char)*p)) p++;
    if (p >= end) return false;
    const char* q = p; while (q < end && !isspace((unsigned char)*q)) q++;
    w.assign(p, q); p = q; return true;
  };
  std::string w;
  for (;;) {
    if (!word(w)) { err = "AMC: no :DEGREES header"; return DRT_ERR_INVALID; }
    if (w == ":FORCE-ALL-JOINTS-BE-3DOF")                    // skeleton.cpp:467-499
      for (BoneH& b : bones) { if (b.dof == 0) continue; for (int a = 0; a < 3; a++) if (!b.rot[a] && b.dof < 7) { b.rot[a] = true; b.order[b.dof++] = a + 1; b.order[b.dof] = 0; } }
    if (w == ":DEGREES") break;
  }
  for (int f = 0; f < n_frames; f++) {
    if (!word(w)) { err = "AMC: truncated"; return DRT_ERR_INVALID; }        // frame number
    for (int j = 0; j < moving; j++) {
      if (!word(w)) { err = "AMC: truncated"; return DRT_ERR_INVALID; }
      size_t bi = 0; while (bi < nb && bones[bi].name != w) bi++;
      if (bi >= nb) { err = "AMC: unknown bone '" + w + "'"; return DRT_ERR_INVALID; }
      double* R = &rot[((size_t)f * nb + bi) * 3]; double* Tt = &trans[((size_t)f * nb + bi) * 3];
      R[0] = R[1] = R[2] = 0;
      for (int x = 0; x < bones[bi].dof; x++) {
        if (!word(w)) { err = "AMC: truncated"; return DRT_ERR_INVALID; }
        const double v = strtod(w.c_str(), nullptr);
        const int code = bones[bi].order[x];
        if (code >= 1 && code <= 3) R[code - 1] = v; else if (code >= 4 && code <= 6) Tt[code - 4] = v * scale;
      }
    }
  }
  return DRT_OK;
}

thread_local std::string g_skel_err;

}  // namespace

struct Skeleton {
  int device = 0, n_bones = 0, n_frames = 0;
  double* d_table = nullptr;            // [frame][n_bones-1][6], device
  std::vector<double> table;            // host mirror (for re-posing scenes, whose flattening is host side)
  float fk_ms = 0;
};

const std::string& skeletonError() { return g_skel_err; }

// Host half only (no device needed): structure of the parsed clip, for bindings and tests.
int skeletonParseInfo(const char* asf, size_t asf_len, const char* amc, size_t amc_len, double scale, int* n_bones, int* n_frames,
                      int* parents, int* dofs, int cap) {
  std::string& err = g_skel_err;
  if (!asf || !amc) { err = "null argument"; return DRT_ERR_INVALID; }
  std::vector<BoneH> bones;
  int rc = parseAsf(asf, asf_len, scale, bones, err);
  if (rc) return rc;
  int nf = 0;
  std::vector<double> rot, trans;
  rc = parseAmc(amc, amc_len, scale, bones, nf, rot, trans, err);
  if (rc) return rc;
  if (n_bones) *n_bones = (int)bones.size();
  if (n_frames) *n_frames = nf;
  for (int i = 0; i < (int)bones.size() && i < cap; i++) {
    if (parents) parents[i] = bones[i].parent;
    if (dofs) dofs[i] = (bones[i].rot[0] ? DOF_RX : 0) | (bones[i].rot[1] ? DOF_RY : 0) | (bones[i].rot[2] ? DOF_RZ : 0) |
                        (bones[i].trans[0] ? DOF_TX : 0) | (bones[i].trans[1] ? DOF_TY : 0) | (bones[i].trans[2] ? DOF_TZ : 0);
  }
  return DRT_OK;
}

int skeletonCreate(const char* asf, size_t asf_len, const char* amc, size_t amc_len, double scale, int device, Skeleton** out) {
  std::string& err = g_skel_err;
  if (!asf || !amc || !out) { err = "null argument"; return DRT_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); err = "no CUDA device (there is no CPU fallback)"; return DRT_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) { err = "bad device ordinal"; return DRT_ERR_INVALID; }
  std::vector<BoneH> bones;
  int rc = parseAsf(asf, asf_len, scale, bones, err);
  if (rc) return rc;
  int n_frames = 0;
  std::vector<double> rot, trans;
  rc = parseAmc(amc, amc_len, scale, bones, n_frames, rot, trans, err);
  if (rc) return rc;
  const int nb = (int)bones.size();
  if (nb < 2) { err = "ASF: no bones"; return DRT_ERR_INVALID; }

  // ---- per-bone constants --------------------------------------------------------------
  std::vector<BoneK> K(nb);
  for (int i = 0; i < nb; i++) {
    BoneH& b = bones[i];
    BoneK& k = K[i];
    memset(&k, 0, sizeof(k));
    k.parent = b.parent;
    k.dof = (b.rot[0] ? DOF_RX : 0) | (b.rot[1] ? DOF_RY : 0) | (b.rot[2] ? DOF_RZ : 0) | (b.trans[0] ? DOF_TX : 0) |
            (b.trans[1] ? DOF_TY : 0) | (b.trans[2] ? DOF_TZ : 0);
    // direction into the bone's own frame (skeleton.cpp:431-452)
    double dir[3] = {b.dir[0], b.dir[1], b.dir[2]};
    if (i > 0) {
      double inv[16]; euler(b.axis, true, inv);
      const double v[4] = {b.dir[0], b.dir[1], b.dir[2], 1.0};
      for (int a = 0; a < 3; a++) dir[a] = (inv[4 * a] * v[0] + inv[4 * a + 1] * v[1]) + (inv[4 * a + 2] * v[2] + inv[4 * a + 3] * v[3]);
    }
    // frame change parent -> bone (skeleton.cpp:343-424); stored transposed, read back untransposed by
    // glMultMatrixd's (x,y) loop, i.e. A = tmp^T
    double own[16], tmp[16];
    euler(b.axis, false, own);
    if (i == 0) memcpy(tmp, own, sizeof(tmp));
    else { double pinv[16]; euler(bones[b.parent].axis, true, pinv); mul4(pinv, own, tmp); }
    for (int x = 0; x < 4; x++) for (int y = 0; y < 4; y++) k.A[4 * x + y] = tmp[4 * y + x];
    // draw rotation z -> dir (displaySkeleton.cpp:160-183)
    ident(k.D);
    if (i > 0) {
      const double z[3] = {0.0, 0.0, 1.0};
      double ra[3] = {z[1] * dir[2] - z[2] * dir[1], z[2] * dir[0] - z[0] * dir[2], z[0] * dir[1] - z[1] * dir[0]};
      const double dp = z[0] * dir[0] + z[1] * dir[1] + z[2] * dir[2];
      const double rl = sqrt(ra[0] * ra[0] + ra[1] * ra[1] + ra[2] * ra[2]);
      const double theta = atan2(rl, dp);
      const double sq = ra[0] * ra[0] + (ra[1] * ra[1] + ra[2] * ra[2]);
      if (sq > 0) { const double n = std::sqrt(sq); ra[0] /= n; ra[1] /= n; ra[2] /= n; }
      double R[16], S[16];
      angleAxisT4(std::sin(theta), std::cos(theta), ra, R);
      ident(S); S[0] = 0.25; S[5] = 0.25;
      mul4(S, R, k.D);
    }
    for (int a = 0; a < 3; a++) k.off[a] = (float)(dir[a] * b.length);
    k.len = (float)b.length;
  }
  // ---- per-(frame, bone) inputs ----------------------------------------------------------
  std::vector<PoseK> P((size_t)n_frames * nb);
  for (int f = 0; f < n_frames; f++) {
    // setPosture only overwrites the degrees of freedom a bone has; the others keep the value of
    // setBasePosture (0) and are never applied
    for (int i = 0; i < nb; i++) {
      PoseK& p = P[(size_t)f * nb + i];
      const double* R = &rot[((size_t)f * nb + i) * 3]; const double* Tt = &trans[((size_t)f * nb + i) * 3];
      for (int a = 0; a < 3; a++) {
        const float deg = (float)(bones[i].rot[a] ? R[a] : 0.0);
        const float radians = (float)(((double)deg / 360.0) * 2.0 * kPi);     // myRotatef, displaySkeleton.cpp:56
        p.sc[2 * a] = std::sin((double)radians); p.sc[2 * a + 1] = std::cos((double)radians);
        p.t[a] = (float)(bones[i].trans[a] ? Tt[a] : 0.0);
      }
      p.pad = 0;
    }
  }

  Skeleton* s = new Skeleton();
  s->device = device; s->n_bones = nb; s->n_frames = n_frames;
  BoneK* dK = nullptr; PoseK* dP = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const size_t n_out = (size_t)n_frames * (nb - 1) * 6;
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaMalloc(&dK, sizeof(BoneK) * nb);
  if (ce == cudaSuccess) ce = cudaMalloc(&dP, sizeof(PoseK) * P.size());
  if (ce == cudaSuccess) ce = cudaMalloc(&s->d_table, sizeof(double) * n_out);
  if (ce == cudaSuccess) ce = cudaMemcpy(dK, K.data(), sizeof(BoneK) * nb, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(dP, P.data(), sizeof(PoseK) * P.size(), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaEventCreate(&e0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&e1);
  if (ce == cudaSuccess) {
    const long long threads = (long long)n_frames * (nb - 1);
    const int block = 128;
    cudaEventRecord(e0);
    skeleton_fk<<<(unsigned)((threads + block - 1) / block), block>>>(dK, dP, nb, n_frames, s->d_table);
    cudaEventRecord(e1);
    ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaEventSynchronize(e1);
    if (ce == cudaSuccess) cudaEventElapsedTime(&s->fk_ms, e0, e1);
  }
  if (ce == cudaSuccess) { s->table.resize(n_out); ce = cudaMemcpy(s->table.data(), s->d_table, sizeof(double) * n_out, cudaMemcpyDeviceToHost); }
  if (dK) cudaFree(dK);
  if (dP) cudaFree(dP);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (ce != cudaSuccess) {
    err = std::string("skeleton upload / forward kinematics failed: ") + cudaGetErrorString(ce);
    if (s->d_table) cudaFree(s->d_table);
    delete s;
    return DRT_ERR_CUDA;
  }
  *out = s;
  return DRT_OK;
}

void skeletonDestroy(Skeleton* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->d_table) cudaFree(s->d_table);
  delete s;
}

int skeletonCylinders(const Skeleton* s) { return s->n_bones - 1; }
int skeletonFrames(const Skeleton* s) { return s->n_frames; }
float skeletonFkMs(const Skeleton* s) { return s->fk_ms; }
const double* skeletonHostTable(const Skeleton* s) { return s->table.data(); }

int skeletonReadBones(const Skeleton* s, int frame0, int n, double* out) {
  if (!s || !out || frame0 < 0 || n < 0 || frame0 + n > s->n_frames) { g_skel_err = "frame range outside the clip"; return DRT_ERR_INVALID; }
  if (cudaSetDevice(s->device) != cudaSuccess) { g_skel_err = "cudaSetDevice failed"; return DRT_ERR_CUDA; }
  const size_t per = (size_t)(s->n_bones - 1) * 6;
  cudaError_t ce = cudaMemcpy(out, s->d_table + per * frame0, sizeof(double) * per * n, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) { g_skel_err = std::string("cudaMemcpy failed: ") + cudaGetErrorString(ce); return DRT_ERR_CUDA; }
  return DRT_OK;
}

}  // namespace drt
