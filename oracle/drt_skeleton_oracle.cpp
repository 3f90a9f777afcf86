// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product.
//
// CPU restatement of the reference's mocap path: ASF skeleton parse, AMC motion parse,
// forward kinematics and the bone -> cylinder step.  Citations are relative to
// /root/reference/:
//   Skeleton::Skeleton / readASFfile            skeleton.cpp:545-590, 118-293
//   RotationX/Y/Z (float angle, float sin/cos)  skeleton.cpp:19-56
//   RotateBoneDirToLocalCoordSystem             skeleton.cpp:431-452
//   ComputeRotationToParentCoordSystem          skeleton.cpp:370-424
//   Motion::readAMCfile                         motion.cpp:92-213
//   Skeleton::setPosture                        skeleton.cpp:502-528
//   DisplaySkeleton::DrawBone / Traverse /
//   ComputeBonePositions (+ the software GL)    displaySkeleton.cpp:19-73, 117-270
//   setSkeletonsToSpecifiedFrame                scene.h:109-128
//   bone -> cylinder end points                 scene.h:616-659
// Arithmetic follows the reference expression by expression: 4x4 double matrices in the
// row-vector convention of its "software GL", float-narrowed angles and offsets where
// the reference narrows, and Eigen's reduction order for the products ((p0+p1)+(p2+p3);
// p0+(p1+p2) for 3-vectors).
//
// Pin: bit-identical to the compiled reference (oracle/_ref, drtref_mocap_bones) on
// tests/golden/mocap_bones_880_999.npy (frames 0..119 of 90.asf / 90_16_v3.amc) and, when
// the reference tree is present, on frames sampled over the whole clip
// (tests/test_skeleton.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

const int MAX_BONES = 256;   // types.h:9

struct M4 {
  double m[4][4];
  double& operator()(int r, int c) { return m[r][c]; }
  double operator()(int r, int c) const { return m[r][c]; }
};
M4 zero4() { M4 r; memset(&r, 0, sizeof(r)); return r; }
M4 ident4() { M4 r = zero4(); for (int i = 0; i < 4; i++) r(i, i) = 1; return r; }
M4 mul(const M4& a, const M4& b) {                       // Eigen fixed-size product, halves reduction
  M4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      r(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + (a(i, 2) * b(2, j) + a(i, 3) * b(3, j));
  return r;
}
void mulv(const M4& a, const double v[4], double out[4]) {
  for (int i = 0; i < 4; i++) out[i] = (a(i, 0) * v[0] + a(i, 1) * v[1]) + (a(i, 2) * v[2] + a(i, 3) * v[3]);
}
M4 transpose(const M4& a) { M4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r(i, j) = a(j, i); return r; }

// skeleton.cpp:19-56: the angle is narrowed to float, cos/sin resolve to the double
// overloads of <cmath> (no `using namespace std` in that file) and are narrowed again
M4 rotX(float th) { M4 R = zero4(); R(0, 0) = R(3, 3) = 1; float c = cos(th), s = sin(th); R(1, 1) = R(2, 2) = c; R(1, 2) = -s; R(2, 1) = s; return R; }
M4 rotY(float th) { M4 R = zero4(); R(1, 1) = R(3, 3) = 1; float c = cos(th), s = sin(th); R(0, 0) = R(2, 2) = c; R(0, 2) = s; R(2, 0) = -s; return R; }
M4 rotZ(float th) { M4 R = zero4(); R(2, 2) = R(3, 3) = 1; float c = cos(th), s = sin(th); R(0, 0) = R(1, 1) = c; R(0, 1) = -s; R(1, 0) = s; return R; }

// Eigen AngleAxis::toRotationMatrix, then transposeInPlace, then toMatrix4
// (displaySkeleton.cpp:23-30, 50-61)
M4 angleAxisT4(double angle, const double ax[3]) {
  double res[3][3];
  const double sn = std::sin(angle), c = std::cos(angle);
  const double sa[3] = {sn * ax[0], sn * ax[1], sn * ax[2]};
  const double c1[3] = {(1.0 - c) * ax[0], (1.0 - c) * ax[1], (1.0 - c) * ax[2]};
  double tmp;
  tmp = c1[0] * ax[1]; res[0][1] = tmp - sa[2]; res[1][0] = tmp + sa[2];
  tmp = c1[0] * ax[2]; res[0][2] = tmp + sa[1]; res[2][0] = tmp - sa[1];
  tmp = c1[1] * ax[2]; res[1][2] = tmp - sa[0]; res[2][1] = tmp + sa[0];
  res[0][0] = c1[0] * ax[0] + c; res[1][1] = c1[1] * ax[1] + c; res[2][2] = c1[2] * ax[2] + c;
  M4 r = ident4();
  for (int y = 0; y < 3; y++) for (int x = 0; x < 3; x++) r(x, y) = res[y][x];   // transposed
  return r;
}
void normalize3(double v[3]) {                          // Eigen normalize(): zero vector unchanged
  const double z = v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]);
  if (z > 0) { const double n = std::sqrt(z); v[0] /= n; v[1] /= n; v[2] /= n; }
}

struct Bone {
  std::string name;
  int parent = -1;
  std::vector<int> children;       // in hierarchy order (child, then its siblings)
  double dir[3] = {0, 0, 0}, length = 0, axis[3] = {0, 0, 0}, aspx = 0.25, aspy = 0.25;
  int dof = 0, dofr[3] = {0, 0, 0}, doft[3] = {0, 0, 0}, doftl = 0, dofo[8] = {0};
  double rpc[4][4];                // rot_parent_current
  double r[3] = {0, 0, 0}, t[3] = {0, 0, 0};
};

struct Skel {
  std::vector<Bone> bones;
  int num_bones = 1;
  double scale = 0.06;
  // motion: per frame, per bone rotation (degrees) and translation (scaled)
  int n_frames = 0;
  std::vector<double> rot, trans;  // [frame][MAX_BONES][3]
};

std::string g_err;

std::vector<std::string> lines_of(const char* text, size_t len) {
  std::vector<std::string> out; std::string cur;
  for (size_t i = 0; i < len; i++) { if (text[i] == '\n') { out.push_back(cur); cur.clear(); } else cur.push_back(text[i]); }
  if (!cur.empty()) out.push_back(cur);
  for (auto& s : out) if (!s.empty() && s.back() == '\r') s.pop_back();          // removeCR, skeleton.cpp:78-82
  return out;
}
std::string first_word(const std::string& s) {          // sscanf("%s")
  size_t i = 0; while (i < s.size() && isspace((unsigned char)s[i])) i++;
  size_t j = i; while (j < s.size() && !isspace((unsigned char)s[j])) j++;
  return s.substr(i, j - i);
}
std::vector<std::string> split_spaces(const std::string& s) {   // strtok(str, " ")
  std::vector<std::string> out; std::string cur;
  for (char ch : s) { if (ch == ' ') { if (!cur.empty()) out.push_back(cur); cur.clear(); } else cur.push_back(ch); }
  if (!cur.empty()) out.push_back(cur);
  return out;
}
int name2idx(const Skel& sk, const std::string& n) {
  for (int i = 0; i < sk.num_bones; i++) if (sk.bones[i].name == n) return i;
  return -1;
}

// skeleton.cpp:118-293
bool read_asf(Skel& sk, const char* text, size_t len) {
  std::vector<std::string> L = lines_of(text, len);
  size_t ln = 0;
  std::string keyword;
  auto next = [&](std::string& s) { if (ln >= L.size()) return false; s = L[ln++]; return true; };
  std::string str;
  for (;;) {
    if (!next(str)) { g_err = "asf: no :bonedata"; return false; }
    std::string w = first_word(str); if (!w.empty()) keyword = w;
    if (keyword == ":bonedata") break;
  }
  if (!next(str)) { g_err = "asf: truncated"; return false; }
  double length = 0;
  bool done = false;
  for (int i = 1; !done && i < MAX_BONES; i++) {
    sk.bones.push_back(Bone());
    Bone& b = sk.bones[i];
    sk.num_bones++;
    for (;;) {
      if (!next(str)) { g_err = "asf: truncated bonedata"; return false; }
      std::string w = first_word(str); if (!w.empty()) keyword = w;      // sscanf leaves keyword on an empty line
      if (keyword == "end") break;
      if (keyword == ":hierarchy") { sk.num_bones--; done = true; break; }
      if (keyword == "name") { char k[256], nm[256]; if (sscanf(str.c_str(), "%255s %255s", k, nm) == 2) b.name = nm; }
      if (keyword == "direction") { char k[256]; sscanf(str.c_str(), "%255s %lf %lf %lf", k, &b.dir[0], &b.dir[1], &b.dir[2]); }
      if (keyword == "length") { char k[256]; sscanf(str.c_str(), "%255s %lf", k, &length); }
      if (keyword == "axis") { char k[256]; sscanf(str.c_str(), "%255s %lf %lf %lf", k, &b.axis[0], &b.axis[1], &b.axis[2]); }
      if (keyword == "dof") {
        b.dof = 0;
        for (const std::string& tok : split_spaces(str)) {
          int code = 0;
          if (tok == "rx") { b.dofr[0] = 1; code = 1; } else if (tok == "ry") { b.dofr[1] = 1; code = 2; }
          else if (tok == "rz") { b.dofr[2] = 1; code = 3; } else if (tok == "tx") { b.doft[0] = 1; code = 4; }
          else if (tok == "ty") { b.doft[1] = 1; code = 5; } else if (tok == "tz") { b.doft[2] = 1; code = 6; }
          else if (tok == "l") { b.doftl = 1; code = 7; } else if (tok == "dof") continue;
          if (b.dof < 7) { b.dofo[b.dof] = code; b.dof++; b.dofo[b.dof] = 0; }
        }
      }
    }
    b.length = length * sk.scale;
  }
  sk.bones.resize(sk.num_bones);
  if (!next(str)) { g_err = "asf: no hierarchy"; return false; }                  // "begin"
  for (;;) {
    if (!next(str)) { g_err = "asf: truncated hierarchy"; return false; }
    std::string w = first_word(str); if (!w.empty()) keyword = w;
    if (keyword == "end") break;
    std::vector<std::string> toks = split_spaces(str);
    int parent = 0;
    for (size_t j = 0; j < toks.size(); j++) {
      int idx = name2idx(sk, toks[j]);
      if (idx < 0) { g_err = "asf: unknown bone " + toks[j]; return false; }
      if (j == 0) parent = idx;
      else { sk.bones[idx].parent = parent; sk.bones[parent].children.push_back(idx); }
    }
  }
  return true;
}

const double PI = 3.14159265358979323846;

void build_skeleton(Skel& sk) {
  // skeleton.cpp:431-452
  for (int i = 1; i < sk.num_bones; i++) {
    Bone& b = sk.bones[i];
    double dir[4] = {b.dir[0], b.dir[1], b.dir[2], 1}, out[4];
    M4 Rz = rotZ(-b.axis[2] * PI / 180.0), Ry = rotY(-b.axis[1] * PI / 180.0), Rx = rotX(-b.axis[0] * PI / 180.0);
    mulv(mul(mul(Rx, Ry), Rz), dir, out);
    b.dir[0] = out[0]; b.dir[1] = out[1]; b.dir[2] = out[2];
  }
  // skeleton.cpp:370-424 (+ 343-365)
  {
    Bone& r = sk.bones[0];
    M4 t2 = mul(mul(rotZ(r.axis[2] * PI / 180.0), rotY(r.axis[1] * PI / 180.0)), rotX(r.axis[0] * PI / 180.0));
    for (int x = 0; x < 4; x++) for (int y = 0; y < 4; y++) r.rpc[x][y] = t2(y, x);
  }
  for (int i = 0; i < sk.num_bones; i++)
    for (int ch : sk.bones[i].children) {
      const Bone& p = sk.bones[i]; Bone& c = sk.bones[ch];
      M4 t1 = mul(mul(rotX(-p.axis[0] * PI / 180.0), rotY(-p.axis[1] * PI / 180.0)), rotZ(-p.axis[2] * PI / 180.0));
      M4 t2 = mul(mul(rotZ(c.axis[2] * PI / 180.0), rotY(c.axis[1] * PI / 180.0)), rotX(c.axis[0] * PI / 180.0));
      M4 t = mul(t1, t2);
      for (int x = 0; x < 4; x++) for (int y = 0; y < 4; y++) c.rpc[x][y] = t(y, x);
    }
  // skeleton.cpp:531-543
  sk.bones[0].aspx = sk.bones[0].aspy = 1;
}

int reachable(const Skel& sk, int b, bool moving_only) {   // numBonesInSkel / movBonesInSkel, skeleton.cpp:58-101
  int n = (!moving_only || sk.bones[b].dof > 0) ? 1 : 0;
  for (int ch : sk.bones[b].children) n += reachable(sk, ch, moving_only);
  return n;
}

// skeleton.cpp:467-499
void enable_all_rotational(Skel& sk) {
  for (int j = 0; j < sk.num_bones; j++) {
    Bone& b = sk.bones[j];
    if (b.dof == 0) continue;
    for (int a = 0; a < 3; a++)
      if (!b.dofr[a] && b.dof < 7) { b.dofr[a] = 1; b.r[a] = 0; b.dof++; b.dofo[b.dof - 1] = a + 1; b.dofo[b.dof] = 0; }
  }
}

// motion.cpp:92-213
bool read_amc(Skel& sk, const char* text, size_t len) {
  // line count: every getline that does not hit end-of-file and is not empty
  int n = 0;
  {
    size_t start = 0;
    for (size_t i = 0; i < len; i++)
      if (text[i] == '\n') { if (i > start) n++; start = i + 1; }
  }
  const int numbones = reachable(sk, 0, false), movbones = reachable(sk, 0, true);
  n = (n - 3) / (movbones + 1);
  if (n < 0) n = 0;
  sk.n_frames = n;
  sk.rot.assign((size_t)n * MAX_BONES * 3, 0.0);
  sk.trans.assign((size_t)n * MAX_BONES * 3, 0.0);
  // whitespace-separated token stream (operator>>)
  std::vector<std::string> tok; { std::string cur;
    for (size_t i = 0; i < len; i++) { if (isspace((unsigned char)text[i])) { if (!cur.empty()) tok.push_back(cur); cur.clear(); } else cur.push_back(text[i]); }
    if (!cur.empty()) tok.push_back(cur); }
  size_t p = 0;
  for (;;) {
    if (p >= tok.size()) { g_err = "amc: no :DEGREES"; return false; }
    const std::string& s = tok[p++];
    if (s == ":FORCE-ALL-JOINTS-BE-3DOF") enable_all_rotational(sk);
    if (s == ":DEGREES") break;
  }
  for (int i = 0; i < n; i++) {
    if (p >= tok.size()) { g_err = "amc: truncated"; return false; }
    p++;                                                      // frame number
    for (int j = 0; j < movbones; j++) {
      if (p >= tok.size()) { g_err = "amc: truncated"; return false; }
      const std::string name = tok[p++];
      int bi; for (bi = 0; bi < numbones; bi++) if (sk.bones[bi].name == name) break;
      if (bi >= numbones) { g_err = "amc: unknown bone " + name; return false; }
      double* R = &sk.rot[((size_t)i * MAX_BONES + bi) * 3];
      double* T = &sk.trans[((size_t)i * MAX_BONES + bi) * 3];
      R[0] = R[1] = R[2] = 0;
      const Bone& b = sk.bones[bi];
      for (int x = 0; x < b.dof; x++) {
        if (p >= tok.size()) { g_err = "amc: truncated"; return false; }
        const double tmp = strtod(tok[p++].c_str(), nullptr);
        switch (b.dofo[x]) {
          case 1: R[0] = tmp; break; case 2: R[1] = tmp; break; case 3: R[2] = tmp; break;
          case 4: T[0] = tmp * sk.scale; break; case 5: T[1] = tmp * sk.scale; break; case 6: T[2] = tmp * sk.scale; break;
          default: break;
        }
      }
    }
  }
  return true;
}

// ---- the software GL of displaySkeleton.cpp:19-73 ---------------------------------
struct GL {
  M4 cur = ident4();
  std::vector<M4> stack;
  void push() { stack.push_back(cur); }
  void pop() { cur = stack.back(); stack.pop_back(); }
  void translatef(float x, float y, float z) { M4 t = ident4(); t(3, 0) = x; t(3, 1) = y; t(3, 2) = z; cur = mul(t, cur); }
  void rotatef(float degrees, float x, float y, float z) {
    double ax[3] = {x, y, z}; normalize3(ax);
    float radians = (degrees / 360.0) * 2.0 * PI;
    cur = mul(angleAxisT4(radians, ax), cur);
  }
  void multMatrixd(const double* m) { M4 A; int i = 0; for (int x = 0; x < 4; x++) for (int y = 0; y < 4; y++, i++) A(x, y) = m[i]; cur = mul(A, cur); }
};

struct Pose { std::vector<M4> rotations, scalings; std::vector<double> translations; };

// displaySkeleton.cpp:117-204
void draw_bone(GL& gl, const Skel& sk, int bi, Pose& pose) {
  const Bone& b = sk.bones[bi];
  gl.multMatrixd(&b.rpc[0][0]);
  if (b.doft[2]) gl.translatef(0.0f, 0.0f, float(b.t[2]));
  if (b.doft[1]) gl.translatef(0.0f, float(b.t[1]), 0.0f);
  if (b.doft[0]) gl.translatef(float(b.t[0]), 0.0f, 0.0f);
  if (b.dofr[2]) gl.rotatef(float(b.r[2]), 0.0f, 0.0f, 1.0f);
  if (b.dofr[1]) gl.rotatef(float(b.r[1]), 0.0f, 1.0f, 0.0f);
  if (b.dofr[0]) gl.rotatef(float(b.r[0]), 1.0f, 0.0f, 0.0f);
  gl.push();
  const double tx = b.dir[0] * b.length, ty = b.dir[1] * b.length, tz = b.dir[2] * b.length;
  if (bi != 0) {
    static const double z_dir[3] = {0.0, 0.0, 1.0};
    double r_axis[3];
    r_axis[0] = z_dir[1] * b.dir[2] - z_dir[2] * b.dir[1];
    r_axis[1] = z_dir[2] * b.dir[0] - z_dir[0] * b.dir[2];
    r_axis[2] = z_dir[0] * b.dir[1] - z_dir[1] * b.dir[0];
    const double dot_prod = z_dir[0] * b.dir[0] + z_dir[1] * b.dir[1] + z_dir[2] * b.dir[2];
    const double r_axis_len = sqrt(r_axis[0] * r_axis[0] + r_axis[1] * r_axis[1] + r_axis[2] * r_axis[2]);
    const double theta = atan2(r_axis_len, dot_prod);
    double ax[3] = {r_axis[0], r_axis[1], r_axis[2]}; normalize3(ax);
    M4 rotation = angleAxisT4(theta, ax);
    M4 scaling = ident4(); scaling(0, 0) = b.aspx; scaling(1, 1) = b.aspy;
    pose.scalings[bi] = scaling;
    M4 T = mul(mul(scaling, rotation), gl.cur);
    pose.translations[4 * bi + 0] = T(3, 0); pose.translations[4 * bi + 1] = T(3, 1); pose.translations[4 * bi + 2] = T(3, 2);
    pose.translations[4 * bi + 3] = 0;     // VEC4 [3] is never written by the reference and never read (head<3>)
    T(3, 0) = 0; T(3, 1) = 0; T(3, 2) = 0;
    pose.rotations[bi] = transpose(T);
  }
  gl.pop();
  gl.translatef(float(tx), float(ty), float(tz));
}

// displaySkeleton.cpp:211-224: child first, then the siblings, each from the parent's matrix
void traverse(GL& gl, const Skel& sk, int bi, Pose& pose) {
  gl.push();
  draw_bone(gl, sk, bi, pose);
  for (int ch : sk.bones[bi].children) traverse(gl, sk, ch, pose);
  gl.pop();
}

// scene.h:109-128 + skeleton.cpp:502-528 + displaySkeleton.cpp:229-270 + scene.h:616-659
void bones_of_frame(Skel& sk, int frame, double* out /* (num_bones-1) x 6 */) {
  const int pid = frame >= sk.n_frames ? sk.n_frames - 1 : frame;
  for (int j = 0; j < sk.num_bones; j++) {
    Bone& b = sk.bones[j];
    const double* R = &sk.rot[((size_t)pid * MAX_BONES + j) * 3];
    const double* T = &sk.trans[((size_t)pid * MAX_BONES + j) * 3];
    for (int a = 0; a < 3; a++) { if (b.dofr[a]) b.r[a] = R[a]; if (b.doft[a]) b.t[a] = T[a]; }
  }
  Pose pose; pose.rotations.assign(sk.num_bones, ident4()); pose.scalings.assign(sk.num_bones, ident4());
  pose.translations.assign(4 * (size_t)sk.num_bones, 0.0);
  GL gl;
  gl.push();
  gl.cur = ident4();
  gl.push();
  // Skeleton::tx..rz stay 0 (skeleton.cpp:571): GetTranslation / GetRotationAngle
  gl.translatef(float(0.06 * 0.0), float(0.06 * 0.0), float(0.06 * 0.0));
  gl.rotatef(float(0.0), 1.0f, 0.0f, 0.0f);
  gl.rotatef(float(0.0), 0.0f, 1.0f, 0.0f);
  gl.rotatef(float(0.0), 0.0f, 0.0f, 1.0f);
  traverse(gl, sk, 0, pose);
  gl.pop();
  gl.pop();
  for (int x = 1; x < sk.num_bones; x++) {
    const float len = (float)sk.bones[x].length;              // boneLengths is vector<float>
    const double lv[4] = {0, 0, 0, 1}, rv[4] = {0, 0, len, 1};
    const M4 RS = mul(pose.rotations[x], pose.scalings[x]);
    double l[4], r[4];
    mulv(RS, lv, l); mulv(RS, rv, r);
    for (int k = 0; k < 3; k++) {
      out[6 * (x - 1) + k] = l[k] + pose.translations[4 * x + k];
      out[6 * (x - 1) + 3 + k] = r[k] + pose.translations[4 * x + k];
    }
  }
}

}  // namespace

struct drt_oracle_skeleton { Skel sk; };

extern "C" {

const char* drt_oracle_skeleton_error(void) { return g_err.c_str(); }

int drt_oracle_skeleton_create(const char* asf, size_t asf_len, const char* amc, size_t amc_len, double scale,
                               drt_oracle_skeleton** out) {
  drt_oracle_skeleton* s = new drt_oracle_skeleton();
  Skel& sk = s->sk;
  sk.scale = scale;
  // Skeleton::Skeleton, skeleton.cpp:545-573
  sk.bones.push_back(Bone());
  Bone& root = sk.bones[0];
  root.name = "root"; root.length = 0.05; root.dof = 6;
  const int order[7] = {4, 5, 6, 1, 2, 3, 0};
  for (int i = 0; i < 7; i++) root.dofo[i] = order[i];
  for (int a = 0; a < 3; a++) root.dofr[a] = root.doft[a] = 1;
  if (!read_asf(sk, asf, asf_len)) { delete s; return -1; }
  build_skeleton(sk);
  if (!read_amc(sk, amc, amc_len)) { delete s; return -1; }
  *out = s;
  return 0;
}

int drt_oracle_skeleton_info(const drt_oracle_skeleton* s, int* n_cylinders, int* n_frames) {
  *n_cylinders = s->sk.num_bones - 1; *n_frames = s->sk.n_frames; return 0;
}

// parent index and DOF bits (rx ry rz tx ty tz = 1 2 4 8 16 32) per bone, root included
int drt_oracle_skeleton_structure(const drt_oracle_skeleton* s, int* parents, int* dofs, int cap) {
  const Skel& sk = s->sk;
  for (int i = 0; i < sk.num_bones && i < cap; i++) {
    const Bone& b = sk.bones[i];
    parents[i] = b.parent;
    dofs[i] = (b.dofr[0] ? 1 : 0) | (b.dofr[1] ? 2 : 0) | (b.dofr[2] ? 4 : 0) | (b.doft[0] ? 8 : 0) | (b.doft[1] ? 16 : 0) | (b.doft[2] ? 32 : 0);
  }
  return sk.num_bones;
}

// end points of the bone cylinders of `frame`: 6 doubles per bone (left xyz, right xyz)
int drt_oracle_skeleton_bones(drt_oracle_skeleton* s, int frame, double* out) {
  if (frame < 0) { g_err = "frameIndex is illegal"; return -1; }             // scene.h:111-115
  bones_of_frame(s->sk, frame, out);
  return s->sk.num_bones - 1;
}

void drt_oracle_skeleton_destroy(drt_oracle_skeleton* s) { delete s; }

}  // extern "C"
