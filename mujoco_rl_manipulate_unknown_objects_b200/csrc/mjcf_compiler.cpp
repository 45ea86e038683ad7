// Host-side model compiler: MJCF subset + binary STL -> HostModel.
//
// Replaces, for the reference's scenes, what `mujoco.Physics.from_xml_path` does at
// reference simulation/environment/robot_env.py:26 (MuJoCo's XML reader + model compiler + mj_setConst).
// The subset is exactly what reference xmls/*.xml use:
//   compiler(angle, meshdir, inertiafromgeom) option(timestep, iterations, tolerance, impratio, gravity, cone)
//   visual/map/znear  default(+nested classes) for geom/joint/motor/light  asset/{mesh,texture,material}
//   worldbody/{body,geom,joint,freejoint,inertial,camera,light}  actuator/motor
// Everything is written from scratch (own XML reader, STL reader, incremental convex hull, Jacobi
// eigen-solver); it shares no code with oracle/mjcf.py, against which tests/test_model_compiler.py checks it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>

#include "model.h"

namespace grs {
namespace {

// ------------------------------------------------------------------ tiny XML reader
struct XmlNode {
  std::string tag;
  std::vector<std::pair<std::string, std::string>> attr;
  std::vector<std::unique_ptr<XmlNode>> kids;
  const std::string* get(const std::string& k) const {
    for (auto& a : attr)
      if (a.first == k) return &a.second;
    return nullptr;
  }
  const XmlNode* child(const std::string& t) const {
    for (auto& c : kids)
      if (c->tag == t) return c.get();
    return nullptr;
  }
};

struct XmlParser {
  const std::string& s;
  size_t i = 0;
  explicit XmlParser(const std::string& src) : s(src) {}
  [[noreturn]] void fail(const std::string& m) { throw std::runtime_error("MJCF parse error at byte " + std::to_string(i) + ": " + m); }
  void skip_ws() { while (i < s.size() && isspace((unsigned char)s[i])) i++; }
  bool starts(const char* p) { return s.compare(i, strlen(p), p) == 0; }
  void skip_misc() {
    for (;;) {
      skip_ws();
      if (starts("<!--")) { size_t e = s.find("-->", i); if (e == std::string::npos) fail("unterminated comment"); i = e + 3; }
      else if (starts("<?")) { size_t e = s.find("?>", i); if (e == std::string::npos) fail("unterminated declaration"); i = e + 2; }
      else break;
    }
  }
  std::string name() {
    size_t b = i;
    while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '-' || s[i] == ':' || s[i] == '.')) i++;
    if (b == i) fail("expected a name");
    return s.substr(b, i - b);
  }
  std::unique_ptr<XmlNode> element() {
    skip_misc();
    if (i >= s.size() || s[i] != '<') fail("expected '<'");
    i++;
    auto n = std::make_unique<XmlNode>();
    n->tag = name();
    for (;;) {
      skip_ws();
      if (i >= s.size()) fail("unexpected end of file in tag");
      if (s[i] == '/') { if (i + 1 >= s.size() || s[i + 1] != '>') fail("malformed '/>'"); i += 2; return n; }
      if (s[i] == '>') { i++; break; }
      std::string k = name();
      skip_ws();
      if (i >= s.size() || s[i] != '=') fail("expected '=' after attribute " + k);
      i++;
      skip_ws();
      char q = s[i];
      if (q != '"' && q != '\'') fail("expected quoted attribute value");
      size_t e = s.find(q, i + 1);
      if (e == std::string::npos) fail("unterminated attribute value");
      n->attr.emplace_back(k, s.substr(i + 1, e - i - 1));
      i = e + 1;
    }
    for (;;) {
      skip_misc();
      if (i >= s.size()) fail("unexpected end of file inside <" + n->tag + ">");
      if (starts("</")) {
        i += 2;
        std::string t = name();
        if (t != n->tag) fail("mismatched closing tag </" + t + "> for <" + n->tag + ">");
        skip_ws();
        if (i >= s.size() || s[i] != '>') fail("malformed closing tag");
        i++;
        return n;
      }
      if (s[i] == '<') n->kids.push_back(element());
      else i++;  // character data is not used by MJCF
    }
  }
};

std::vector<double> floats(const std::string& s) {
  std::vector<double> v;
  std::istringstream is(s);
  double x;
  while (is >> x) v.push_back(x);
  return v;
}

// ------------------------------------------------------------------ math
using V3 = std::array<double, 3>;
using Q4 = std::array<double, 4>;
using M3 = std::array<double, 9>;

V3 sub(const V3& a, const V3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
V3 add(const V3& a, const V3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
V3 scl(const V3& a, double s) { return {a[0] * s, a[1] * s, a[2] * s}; }
double dot(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
V3 cross(const V3& a, const V3& b) { return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}; }
double norm(const V3& a) { return std::sqrt(dot(a, a)); }
Q4 qmul(const Q4& a, const Q4& b) {
  return {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
          a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
}
Q4 qnormalize(Q4 q) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (auto& x : q) x /= n;
  return q;
}
M3 q2m(const Q4& q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  return {w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
          2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
          2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z};
}
V3 mulv(const M3& R, const V3& v) { return {R[0] * v[0] + R[1] * v[1] + R[2] * v[2], R[3] * v[0] + R[4] * v[1] + R[5] * v[2], R[6] * v[0] + R[7] * v[1] + R[8] * v[2]}; }
V3 multv(const M3& R, const V3& v) { return {R[0] * v[0] + R[3] * v[1] + R[6] * v[2], R[1] * v[0] + R[4] * v[1] + R[7] * v[2], R[2] * v[0] + R[5] * v[1] + R[8] * v[2]}; }
Q4 m2q(const M3& R) {
  double t = R[0] + R[4] + R[8];
  Q4 q;
  if (t > 0) {
    double s = std::sqrt(t + 1.0) * 2;
    q = {0.25 * s, (R[7] - R[5]) / s, (R[2] - R[6]) / s, (R[3] - R[1]) / s};
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0) * 2;
    q[0] = (R[3 * k + j] - R[3 * j + k]) / s;
    q[1 + i] = 0.25 * s;
    q[1 + j] = (R[3 * j + i] + R[3 * i + j]) / s;
    q[1 + k] = (R[3 * k + i] + R[3 * i + k]) / s;
  }
  if (q[0] < 0) for (auto& x : q) x = -x;
  return qnormalize(q);
}
Q4 euler2q(const std::vector<double>& e) {  // MuJoCo default eulerseq "xyz": intrinsic x, y, z
  Q4 q{1, 0, 0, 0};
  for (int ax = 0; ax < 3; ax++) {
    Q4 r{std::cos(e[ax] / 2), 0, 0, 0};
    r[1 + ax] = std::sin(e[ax] / 2);
    q = qmul(q, r);
  }
  return q;
}

// symmetric 3x3 eigen-decomposition by cyclic Jacobi; eigenvalues descending, V columns = axes, det +1
void eig3(const double A_in[9], double w[3], M3& V) {
  double A[9];
  std::memcpy(A, A_in, sizeof A);
  V = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off < 1e-40) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double apq = A[3 * p + q];
        if (std::fabs(apq) < 1e-300) continue;
        double theta = (A[4 * q] - A[4 * p]) / (2 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; k++) {  // A <- A J
          double akp = A[3 * k + p], akq = A[3 * k + q];
          A[3 * k + p] = c * akp - s * akq;
          A[3 * k + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {  // A <- J^T A
          double apk = A[3 * p + k], aqk = A[3 * q + k];
          A[3 * p + k] = c * apk - s * aqk;
          A[3 * q + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = V[3 * k + p], vkq = V[3 * k + q];
          V[3 * k + p] = c * vkp - s * vkq;
          V[3 * k + q] = s * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  double ev[3] = {A[0], A[4], A[8]};
  std::sort(idx, idx + 3, [&](int a, int b) { return ev[a] > ev[b]; });
  M3 Vs;
  for (int c = 0; c < 3; c++) {
    w[c] = ev[idx[c]];
    for (int r = 0; r < 3; r++) Vs[3 * r + c] = V[3 * r + idx[c]];
  }
  V3 c0{Vs[0], Vs[3], Vs[6]}, c1{Vs[1], Vs[4], Vs[7]}, c2{Vs[2], Vs[5], Vs[8]};
  if (dot(cross(c0, c1), c2) < 0) { Vs[2] = -Vs[2]; Vs[5] = -Vs[5]; Vs[8] = -Vs[8]; }
  V = Vs;
}

// ------------------------------------------------------------------ convex hull (incremental, brute-force visibility)
struct Hull {
  std::vector<int> verts;            // indices into the input point set, ascending
  std::vector<std::array<int, 3>> faces;  // input-point indices, outward orientation
};

Hull convex_hull(const std::vector<V3>& P) {
  const int n = (int)P.size();
  if (n < 4) throw std::runtime_error("convex hull needs at least 4 points");
  V3 lo = P[0], hi = P[0];
  for (auto& p : P) for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
  const double scale = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]});
  const double eps = 1e-11 * scale;
  // initial tetrahedron from extreme points
  int i0 = 0, i1 = 0;
  for (int i = 0; i < n; i++) { if (P[i][0] < P[i0][0]) i0 = i; if (P[i][0] > P[i1][0]) i1 = i; }
  int i2 = -1;
  double best = -1;
  for (int i = 0; i < n; i++) { double d = norm(cross(sub(P[i1], P[i0]), sub(P[i], P[i0]))); if (d > best) { best = d; i2 = i; } }
  V3 nrm = cross(sub(P[i1], P[i0]), sub(P[i2], P[i0]));
  int i3 = -1;
  best = -1;
  for (int i = 0; i < n; i++) { double d = std::fabs(dot(nrm, sub(P[i], P[i0]))); if (d > best) { best = d; i3 = i; } }
  if (best < eps * norm(nrm)) throw std::runtime_error("degenerate (flat) mesh: cannot build a convex hull");
  struct Face { int v[3]; V3 n; double d; bool alive; };
  std::vector<Face> F;
  auto mk = [&](int a, int b, int c) {
    Face f;
    f.v[0] = a; f.v[1] = b; f.v[2] = c;
    f.n = cross(sub(P[b], P[a]), sub(P[c], P[a]));
    double l = norm(f.n);
    f.n = scl(f.n, l > 0 ? 1 / l : 0);
    f.d = dot(f.n, P[a]);
    f.alive = true;
    return f;
  };
  if (dot(nrm, sub(P[i3], P[i0])) > 0) std::swap(i1, i2);
  F.push_back(mk(i0, i1, i2)); F.push_back(mk(i0, i3, i1)); F.push_back(mk(i1, i3, i2)); F.push_back(mk(i2, i3, i0));
  // insertion order: farthest from the centroid first (keeps the intermediate hulls small)
  V3 cen = scl(add(add(P[i0], P[i1]), add(P[i2], P[i3])), 0.25);
  std::vector<int> order;
  for (int i = 0; i < n; i++) if (i != i0 && i != i1 && i != i2 && i != i3) order.push_back(i);
  std::vector<double> dist2(n);
  for (int i = 0; i < n; i++) dist2[i] = dot(sub(P[i], cen), sub(P[i], cen));
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return dist2[a] > dist2[b]; });
  std::map<std::pair<int, int>, int> edge2face;
  for (int pi : order) {
    const V3& p = P[pi];
    std::vector<int> vis;
    for (int f = 0; f < (int)F.size(); f++)
      if (F[f].alive && dot(F[f].n, p) - F[f].d > eps) vis.push_back(f);
    if (vis.empty()) continue;
    edge2face.clear();
    for (int f : vis) for (int k = 0; k < 3; k++) edge2face[{F[f].v[k], F[f].v[(k + 1) % 3]}] = f;
    std::vector<std::pair<int, int>> horizon;
    for (int f : vis)
      for (int k = 0; k < 3; k++) {
        int a = F[f].v[k], b = F[f].v[(k + 1) % 3];
        if (!edge2face.count({b, a})) horizon.push_back({a, b});
      }
    for (int f : vis) F[f].alive = false;
    for (auto& e : horizon) F.push_back(mk(e.first, e.second, pi));
    if (F.size() > 4096) {  // compact
      std::vector<Face> G;
      for (auto& f : F) if (f.alive) G.push_back(f);
      F.swap(G);
    }
  }
  Hull h;
  std::set<int> vs;
  for (auto& f : F)
    if (f.alive) { h.faces.push_back({f.v[0], f.v[1], f.v[2]}); for (int k = 0; k < 3; k++) vs.insert(f.v[k]); }
  h.verts.assign(vs.begin(), vs.end());
  return h;
}

// ------------------------------------------------------------------ meshes
std::vector<std::array<V3, 3>> load_stl(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open mesh file '" + path + "'");
  std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (buf.size() < 84) throw std::runtime_error("mesh file '" + path + "' is too short for binary STL");
  uint32_t n;
  std::memcpy(&n, buf.data() + 80, 4);
  if (buf.size() < 84 + (size_t)n * 50) throw std::runtime_error("mesh file '" + path + "' is truncated (binary STL expected)");
  std::vector<std::array<V3, 3>> tri(n);
  for (uint32_t i = 0; i < n; i++) {
    float v[12];
    std::memcpy(v, buf.data() + 84 + (size_t)i * 50, 48);
    for (int c = 0; c < 3; c++) tri[i][c] = {v[3 + 3 * c], v[4 + 3 * c], v[5 + 3 * c]};
  }
  return tri;
}

void finish_hull(HostMesh& m, const std::vector<V3>& pts) {
  Hull h = convex_hull(pts);
  std::map<int, int> remap;
  for (size_t l = 0; l < h.verts.size(); l++) {
    remap[h.verts[l]] = (int)l;
    for (int k = 0; k < 3; k++) m.hull_verts.push_back(pts[h.verts[l]][k]);
  }
  std::vector<std::set<int>> adj(h.verts.size());
  for (auto& f : h.faces) {
    int a = remap[f[0]], b = remap[f[1]], c = remap[f[2]];
    m.hull_faces.insert(m.hull_faces.end(), {a, b, c});
    adj[a].insert(b); adj[b].insert(a); adj[b].insert(c); adj[c].insert(b); adj[a].insert(c); adj[c].insert(a);
  }
  m.adjadr.push_back(0);
  for (auto& s : adj) { m.adj.insert(m.adj.end(), s.begin(), s.end()); m.adjadr.push_back((int)m.adj.size()); }
}

// MuJoCo 2.2.x user_mesh.cc (legacy inertia): pyramids from the area-weighted centre of the face centroids,
// absolute volumes; recentre to the CoM; rotate into the principal frame (moments descending).
HostMesh process_mesh(const std::string& name, const std::vector<std::array<V3, 3>>& tri_in) {
  HostMesh m;
  m.name = name;
  std::vector<std::array<V3, 3>> tri;
  std::vector<V3> nrm, cen;
  std::vector<double> area;
  V3 facecen{0, 0, 0};
  double atot = 0;
  for (auto& t : tri_in) {
    V3 n = cross(sub(t[1], t[0]), sub(t[2], t[0]));
    double a2 = norm(n);
    if (a2 <= 1e-30) continue;
    tri.push_back(t);
    nrm.push_back(scl(n, 1 / a2));
    area.push_back(0.5 * a2);
    cen.push_back(scl(add(add(t[0], t[1]), t[2]), 1.0 / 3));
    facecen = add(facecen, scl(cen.back(), area.back()));
    atot += area.back();
  }
  if (tri.empty()) throw std::runtime_error("mesh '" + name + "' has no non-degenerate faces");
  facecen = scl(facecen, 1 / atot);
  V3 com{0, 0, 0};
  double vol = 0;
  for (size_t i = 0; i < tri.size(); i++) {
    double v = std::fabs(dot(sub(tri[i][0], facecen), nrm[i])) * area[i] / 3;
    com = add(com, scl(add(scl(cen[i], 0.75), scl(facecen, 0.25)), v));
    vol += v;
  }
  com = scl(com, 1 / vol);
  double P[9] = {0};
  double vol2 = 0;
  for (size_t i = 0; i < tri.size(); i++) {
    V3 D = sub(tri[i][0], com), E = sub(tri[i][1], com), F = sub(tri[i][2], com);
    double v = std::fabs(dot(D, nrm[i])) * area[i] / 3;
    vol2 += v;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++)
        P[3 * a + b] += v / 20 * (2 * (D[a] * D[b] + E[a] * E[b] + F[a] * F[b]) + D[a] * E[b] + D[b] * E[a] + D[a] * F[b] + D[b] * F[a] + E[a] * F[b] + E[b] * F[a]);
  }
  double trP = P[0] + P[4] + P[8], I[9];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) I[3 * a + b] = (a == b ? trP : 0) - P[3 * a + b];
  double w[3];
  M3 V;
  eig3(I, w, V);
  Q4 q = m2q(V);
  V = q2m(q);
  m.pos = com;
  m.quat = q;
  m.volume = vol2;
  m.inertia_unit = {w[0], w[1], w[2]};
  // unique vertices in the principal frame
  std::set<V3> uniq;
  for (auto& t : tri_in) for (auto& v : t) uniq.insert(v);
  std::vector<V3> pts;
  for (auto& v : uniq) pts.push_back(multv(V, sub(v, com)));
  finish_hull(m, pts);
  for (auto& t : tri)
    for (auto& v : t) { V3 l = multv(V, sub(v, com)); for (int k = 0; k < 3; k++) m.tri.push_back((float)l[k]); }
  return m;
}

HostMesh box_mesh(const std::string& name, const std::vector<double>& s) {
  HostMesh m;
  m.name = name;
  m.volume = 8 * s[0] * s[1] * s[2];
  m.inertia_unit = {m.volume / 3 * (s[1] * s[1] + s[2] * s[2]), m.volume / 3 * (s[0] * s[0] + s[2] * s[2]), m.volume / 3 * (s[0] * s[0] + s[1] * s[1])};
  std::vector<V3> pts;
  for (int a = -1; a <= 1; a += 2) for (int b = -1; b <= 1; b += 2) for (int c = -1; c <= 1; c += 2) pts.push_back({a * s[0], b * s[1], c * s[2]});
  finish_hull(m, pts);
  for (size_t f = 0; f < m.hull_faces.size(); f += 3)
    for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) m.tri.push_back((float)m.hull_verts[3 * m.hull_faces[f + k] + c]);
  return m;
}

// ------------------------------------------------------------------ defaults
using AttrMap = std::map<std::string, std::string>;
struct Defaults {
  std::map<std::string, std::map<std::string, AttrMap>> cls;  // class -> tag -> attrs
  void walk(const XmlNode* n, const std::string& name, const std::map<std::string, AttrMap>& inherited) {
    auto cur = inherited;
    for (auto& c : n->kids)
      if (c->tag != "default") for (auto& a : c->attr) cur[c->tag][a.first] = a.second;
    cls[name] = cur;
    for (auto& c : n->kids)
      if (c->tag == "default") {
        const std::string* cn = c->get("class");
        if (!cn) throw std::runtime_error("nested <default> without a class name");
        walk(c.get(), *cn, cur);
      }
  }
  AttrMap resolve(const std::string& tag, const XmlNode* e, const std::string& childclass) const {
    std::string cn = "main";
    if (const std::string* c = e->get("class")) cn = *c;
    else if (!childclass.empty()) cn = childclass;
    AttrMap out;
    auto it = cls.find(cn);
    if (it == cls.end() && cn != "main") throw std::runtime_error("unknown default class '" + cn + "'");
    if (it != cls.end()) { auto jt = it->second.find(tag); if (jt != it->second.end()) out = jt->second; }
    for (auto& a : e->attr) out[a.first] = a.second;
    return out;
  }
};
std::string aget(const AttrMap& a, const std::string& k, const std::string& d) { auto it = a.find(k); return it == a.end() ? d : it->second; }

void frame_of(const AttrMap& a, V3& pos, Q4& quat) {
  auto p = floats(aget(a, "pos", "0 0 0"));
  pos = {p[0], p[1], p[2]};
  if (a.count("quat")) { auto q = floats(a.at("quat")); quat = qnormalize({q[0], q[1], q[2], q[3]}); }
  else if (a.count("euler")) quat = euler2q(floats(a.at("euler")));
  else quat = {1, 0, 0, 0};
}
AttrMap amap(const XmlNode* e) { AttrMap m; for (auto& a : e->attr) m[a.first] = a.second; return m; }

// ------------------------------------------------------------------ host kinematics at qpos0 for mj_setConst quantities
struct Kin { std::vector<V3> xpos, xipos, xanchor, xaxis; std::vector<Q4> xquat; std::vector<M3> xmat, ximat; };

Kin kinematics(const HostModel& m, const std::vector<double>& qpos) {
  Kin k;
  k.xpos.resize(m.nbody); k.xquat.resize(m.nbody); k.xmat.resize(m.nbody); k.xipos.resize(m.nbody); k.ximat.resize(m.nbody);
  k.xanchor.resize(m.njnt); k.xaxis.resize(m.njnt);
  k.xpos[0] = {0, 0, 0}; k.xquat[0] = {1, 0, 0, 0}; k.xmat[0] = q2m(k.xquat[0]); k.xipos[0] = k.xpos[0]; k.ximat[0] = k.xmat[0];
  for (int b = 1; b < m.nbody; b++) {
    int p = m.body_parentid[b], ja = m.body_jntadr[b], jn = m.body_jntnum[b];
    V3 pos; Q4 quat;
    if (jn == 1 && m.jnt_type[ja] == JNT_FREE) {
      int qa = m.jnt_qposadr[ja];
      pos = {qpos[qa], qpos[qa + 1], qpos[qa + 2]};
      quat = qnormalize({qpos[qa + 3], qpos[qa + 4], qpos[qa + 5], qpos[qa + 6]});
      k.xanchor[ja] = pos;
      k.xaxis[ja] = {0, 0, 1};
    } else {
      pos = add(k.xpos[p], mulv(k.xmat[p], {m.body_pos[3 * b], m.body_pos[3 * b + 1], m.body_pos[3 * b + 2]}));
      quat = qmul(k.xquat[p], {m.body_quat[4 * b], m.body_quat[4 * b + 1], m.body_quat[4 * b + 2], m.body_quat[4 * b + 3]});
      for (int j = ja; j < ja + jn; j++) {
        V3 jp{m.jnt_pos[3 * j], m.jnt_pos[3 * j + 1], m.jnt_pos[3 * j + 2]}, ax{m.jnt_axis[3 * j], m.jnt_axis[3 * j + 1], m.jnt_axis[3 * j + 2]};
        M3 R = q2m(quat);
        k.xanchor[j] = add(pos, mulv(R, jp));
        k.xaxis[j] = mulv(R, ax);
        double q = qpos[m.jnt_qposadr[j]] - m.qpos0[m.jnt_qposadr[j]];
        if (m.jnt_type[j] == JNT_SLIDE) pos = add(pos, scl(k.xaxis[j], q));
        else {
          Q4 ql{std::cos(q / 2), ax[0] * std::sin(q / 2), ax[1] * std::sin(q / 2), ax[2] * std::sin(q / 2)};
          quat = qmul(quat, ql);
          pos = sub(k.xanchor[j], mulv(q2m(quat), jp));
        }
      }
    }
    quat = qnormalize(quat);
    k.xpos[b] = pos; k.xquat[b] = quat; k.xmat[b] = q2m(quat);
    k.xipos[b] = add(pos, mulv(k.xmat[b], {m.body_ipos[3 * b], m.body_ipos[3 * b + 1], m.body_ipos[3 * b + 2]}));
    k.ximat[b] = q2m(qmul(quat, {m.body_iquat[4 * b], m.body_iquat[4 * b + 1], m.body_iquat[4 * b + 2], m.body_iquat[4 * b + 3]}));
  }
  return k;
}

// translational / rotational Jacobian (3 x nv each) of `point` rigidly attached to `body`
void jacobian(const HostModel& m, const Kin& k, int body, const V3& point, std::vector<double>& jp, std::vector<double>& jr) {
  jp.assign(3 * m.nv, 0); jr.assign(3 * m.nv, 0);
  for (int b = body; b > 0; b = m.body_parentid[b])
    for (int j = m.body_jntadr[b]; j >= 0 && j < m.body_jntadr[b] + m.body_jntnum[b]; j++) {
      int da = m.jnt_dofadr[j];
      if (m.jnt_type[j] == JNT_FREE) {
        for (int c = 0; c < 3; c++) {
          jp[c * m.nv + da + c] = 1;
          V3 ax{k.xmat[b][c], k.xmat[b][3 + c], k.xmat[b][6 + c]};
          V3 lin = cross(ax, sub(point, k.xpos[b]));
          for (int r = 0; r < 3; r++) { jr[r * m.nv + da + 3 + c] = ax[r]; jp[r * m.nv + da + 3 + c] = lin[r]; }
        }
      } else if (m.jnt_type[j] == JNT_SLIDE) {
        for (int r = 0; r < 3; r++) jp[r * m.nv + da] = k.xaxis[j][r];
      } else {
        V3 lin = cross(k.xaxis[j], sub(point, k.xanchor[j]));
        for (int r = 0; r < 3; r++) { jr[r * m.nv + da] = k.xaxis[j][r]; jp[r * m.nv + da] = lin[r]; }
      }
    }
}

void set_const(HostModel& m) {  // engine_setconst.c : mj_setConst (the parts the solver needs)
  Kin k = kinematics(m, m.qpos0);
  const int nv = m.nv;
  std::vector<double> M(nv * nv, 0), jp, jr;
  for (int b = 1; b < m.nbody; b++) {
    if (m.body_mass[b] <= 0) continue;
    jacobian(m, k, b, k.xipos[b], jp, jr);
    const M3& R = k.ximat[b];
    double I[9];
    for (int a = 0; a < 3; a++)
      for (int c = 0; c < 3; c++) {
        I[3 * a + c] = 0;
        for (int d = 0; d < 3; d++) I[3 * a + c] += R[3 * a + d] * m.body_inertia[3 * b + d] * R[3 * c + d];
      }
    for (int i = 0; i < nv; i++)
      for (int j = 0; j < nv; j++) {
        double v = 0;
        for (int r = 0; r < 3; r++) {
          v += m.body_mass[b] * jp[r * nv + i] * jp[r * nv + j];
          for (int c = 0; c < 3; c++) v += jr[r * nv + i] * I[3 * r + c] * jr[c * nv + j];
        }
        M[i * nv + j] += v;
      }
  }
  double tr = 0;
  for (int i = 0; i < nv; i++) { M[i * nv + i] += m.dof_armature[i]; tr += M[i * nv + i]; }
  m.meaninertia = nv ? tr / nv : 1;
  // Minv by Gauss-Jordan (SPD, tiny)
  std::vector<double> A = M, Minv(nv * nv, 0);
  for (int i = 0; i < nv; i++) Minv[i * nv + i] = 1;
  for (int c = 0; c < nv; c++) {
    double piv = A[c * nv + c];
    if (std::fabs(piv) < 1e-300) throw std::runtime_error("singular joint-space inertia at qpos0 (a moving body has no mass/armature)");
    for (int j = 0; j < nv; j++) { A[c * nv + j] /= piv; Minv[c * nv + j] /= piv; }
    for (int r = 0; r < nv; r++) {
      if (r == c) continue;
      double f = A[r * nv + c];
      if (f == 0) continue;
      for (int j = 0; j < nv; j++) { A[r * nv + j] -= f * A[c * nv + j]; Minv[r * nv + j] -= f * Minv[c * nv + j]; }
    }
  }
  m.body_invweight0.assign(2 * m.nbody, 0);
  for (int b = 1; b < m.nbody; b++) {
    if (m.body_weldid[b] == 0) continue;
    jacobian(m, k, b, k.xipos[b], jp, jr);
    double A6[6];
    for (int r = 0; r < 6; r++) {
      const double* J = r < 3 ? &jp[r * nv] : &jr[(r - 3) * nv];
      double v = 0;
      for (int a = 0; a < nv; a++) for (int c = 0; c < nv; c++) v += J[a] * Minv[a * nv + c] * J[c];
      A6[r] = v;
    }
    m.body_invweight0[2 * b] = std::max(1e-15, (A6[0] + A6[1] + A6[2]) / 3);
    m.body_invweight0[2 * b + 1] = std::max(1e-15, (A6[3] + A6[4] + A6[5]) / 3);
  }
  m.dof_invweight0.assign(nv, 0);
  for (int j = 0; j < m.njnt; j++) {
    int da = m.jnt_dofadr[j];
    if (m.jnt_type[j] == JNT_FREE) {
      double t = 0, r = 0;
      for (int c = 0; c < 3; c++) { t += Minv[(da + c) * nv + da + c] / 3; r += Minv[(da + 3 + c) * nv + da + 3 + c] / 3; }
      for (int c = 0; c < 3; c++) { m.dof_invweight0[da + c] = t; m.dof_invweight0[da + 3 + c] = r; }
    } else m.dof_invweight0[da] = Minv[da * nv + da];
  }
  // stat.extent / center from the bounding spheres of the non-plane geoms at qpos0 (engine_setconst.c : set0)
  V3 lo{1e30, 1e30, 1e30}, hi{-1e30, -1e30, -1e30};
  for (int g = 0; g < m.ngeom; g++) {
    if (m.geom_type[g] == GEOM_PLANE) continue;
    int b = m.geom_bodyid[g];
    V3 c = add(k.xpos[b], mulv(k.xmat[b], {m.geom_pos[3 * g], m.geom_pos[3 * g + 1], m.geom_pos[3 * g + 2]}));
    for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], c[a] - m.geom_rbound[g]); hi[a] = std::max(hi[a], c[a] + m.geom_rbound[g]); }
  }
  m.extent = 0;
  for (int a = 0; a < 3; a++) { m.center[a] = 0.5 * (lo[a] + hi[a]); m.extent = std::max(m.extent, 0.5 * (hi[a] - lo[a])); }
}

}  // namespace

// ------------------------------------------------------------------ the compiler proper
HostModel compile_mjcf(const std::string& xml_path) {
  std::ifstream f(xml_path);
  if (!f) throw std::runtime_error("cannot open scene file '" + xml_path + "'");
  std::stringstream ss;
  ss << f.rdbuf();
  std::string src = ss.str();
  XmlParser parser(src);
  auto root = parser.element();
  if (root->tag != "mujoco") throw std::runtime_error("root element must be <mujoco>, got <" + root->tag + ">");
  HostModel m;
  std::string meshdir, inertiafromgeom = "auto";
  if (auto* c = root->child("compiler")) {
    if (auto* a = c->get("angle")) if (*a != "radian") throw std::runtime_error("only <compiler angle=\"radian\"> scenes are supported");
    if (!c->get("angle")) throw std::runtime_error("only <compiler angle=\"radian\"> scenes are supported (MJCF default is degree)");
    if (auto* a = c->get("meshdir")) meshdir = *a;
    if (auto* a = c->get("inertiafromgeom")) inertiafromgeom = *a;
  }
  std::string dir = xml_path.substr(0, xml_path.find_last_of('/') == std::string::npos ? 0 : xml_path.find_last_of('/') + 1);
  std::string meshroot = dir + (meshdir.empty() ? "" : meshdir + "/");
  if (auto* o = root->child("option")) {
    if (auto* a = o->get("timestep")) m.timestep = std::stod(*a);
    if (auto* a = o->get("gravity")) { auto g = floats(*a); for (int k = 0; k < 3; k++) m.gravity[k] = g[k]; }
    if (auto* a = o->get("impratio")) m.impratio = std::stod(*a);
    if (auto* a = o->get("tolerance")) m.tolerance = std::stod(*a);
    if (auto* a = o->get("iterations")) m.iterations = std::stoi(*a);
    if (auto* a = o->get("cone")) m.cone_elliptic = (*a == "elliptic");
    if (auto* a = o->get("integrator")) if (*a != "Euler") throw std::runtime_error("only the Euler integrator is supported");
    if (auto* a = o->get("solver")) if (*a != "Newton") throw std::runtime_error("only the Newton solver is supported");
  }
  if (!m.cone_elliptic) throw std::runtime_error("only <option cone=\"elliptic\"> scenes are supported");
  if (auto* v = root->child("visual")) if (auto* mp = v->child("map")) {
    if (auto* a = mp->get("znear")) m.znear = std::stod(*a);
    if (auto* a = mp->get("zfar")) m.zfar = std::stod(*a);
  }
  Defaults dfl;
  if (auto* d = root->child("default")) dfl.walk(d, "main", {});

  std::map<std::string, AttrMap> materials, textures;
  if (auto* as = root->child("asset"))
    for (auto& e : as->kids) {
      if (e->tag == "mesh") {
        const std::string* fn = e->get("file");
        if (!fn) throw std::runtime_error("<mesh> without file attribute");
        std::string nm = e->get("name") ? *e->get("name") : fn->substr(0, fn->find_last_of('.'));
        m.meshes.push_back(process_mesh(nm, load_stl(meshroot + *fn)));
      } else if (e->tag == "material") {
        if (auto* n = e->get("name")) materials[*n] = amap(e.get());
      } else if (e->tag == "texture") {
        AttrMap t = amap(e.get());
        if (aget(t, "type", "") == "skybox") {
          auto a = floats(aget(t, "rgb1", "0.8 0.8 0.8")), b = floats(aget(t, "rgb2", "0.5 0.5 0.5"));
          for (int k = 0; k < 3; k++) { m.sky_rgb1[k] = a[k]; m.sky_rgb2[k] = b[k]; }
        } else if (auto* n = e->get("name")) textures[*n] = t;
      }
    }

  struct BodyTmp { int parent; V3 pos; Q4 quat; bool has_inertial = false; V3 ipos; Q4 iquat; double imass; V3 iinertia; std::vector<int> geoms; };
  std::vector<BodyTmp> bodies(1);
  bodies[0].parent = 0; bodies[0].pos = {0, 0, 0}; bodies[0].quat = {1, 0, 0, 0};
  m.body_names.push_back("world");
  struct GeomTmp { int body, type, condim, mesh; V3 pos; Q4 quat; double friction[3], margin, gap, solref[2], solimp[5], rgba[4], size[3], mass, inertia[3], matprop[4]; int contype, conaffinity; };
  std::vector<GeomTmp> geoms;
  struct JntTmp { int body, type, limited; V3 pos, axis; double range[2], armature, damping; };
  std::vector<JntTmp> joints;

  std::function<void(const XmlNode*, int, const std::string&)> walk = [&](const XmlNode* elem, int bid, const std::string& childclass) {
    for (auto& ep : elem->kids) {
      const XmlNode* e = ep.get();
      if (e->tag == "geom") {
        AttrMap a = dfl.resolve("geom", e, childclass);
        GeomTmp g{};
        g.body = bid;
        std::string t = aget(a, "type", "sphere");
        if (t == "plane") g.type = GEOM_PLANE; else if (t == "mesh") g.type = GEOM_MESH; else if (t == "box") g.type = GEOM_BOX;
        else throw std::runtime_error("geom type '" + t + "' is not supported (plane, mesh, box only)");
        g.condim = std::stoi(aget(a, "condim", "3"));
        g.contype = std::stoi(aget(a, "contype", "1")); g.conaffinity = std::stoi(aget(a, "conaffinity", "1"));
        g.friction[0] = 1; g.friction[1] = 0.005; g.friction[2] = 0.0001;
        { auto fr = floats(aget(a, "friction", "")); for (size_t k = 0; k < fr.size() && k < 3; k++) g.friction[k] = fr[k]; }
        g.margin = std::stod(aget(a, "margin", "0")); g.gap = std::stod(aget(a, "gap", "0"));
        { auto v = floats(aget(a, "solref", "0.02 1")); g.solref[0] = v[0]; g.solref[1] = v[1]; }
        { double d5[5] = {0.9, 0.95, 0.001, 0.5, 2}; auto v = floats(aget(a, "solimp", "")); for (size_t k = 0; k < v.size() && k < 5; k++) d5[k] = v[k]; std::memcpy(g.solimp, d5, sizeof d5); }
        double rgba[4] = {0.5, 0.5, 0.5, 1};
        g.matprop[0] = 0; g.matprop[1] = 0.5; g.matprop[2] = 0.5; g.matprop[3] = 0;
        std::string matn = aget(a, "material", "");
        if (!matn.empty()) {
          auto it = materials.find(matn);
          if (it == materials.end()) throw std::runtime_error("unknown material '" + matn + "'");
          if (it->second.count("rgba")) { auto v = floats(it->second.at("rgba")); for (int k = 0; k < 4; k++) rgba[k] = v[k]; }
          else { rgba[0] = rgba[1] = rgba[2] = rgba[3] = 1; }
          g.matprop[0] = std::stod(aget(it->second, "emission", "0"));
          g.matprop[1] = std::stod(aget(it->second, "specular", "0.5"));
          g.matprop[2] = std::stod(aget(it->second, "shininess", "0.5"));
          std::string tx = aget(it->second, "texture", "");
          if (!tx.empty()) {
            g.matprop[3] = 1;
            auto tt = textures.find(tx);
            if (tt != textures.end()) {
              auto c1 = floats(aget(tt->second, "rgb1", "0.8 0.8 0.8")), c2 = floats(aget(tt->second, "rgb2", "0.5 0.5 0.5"));
              for (int k = 0; k < 3; k++) { m.tex_rgb1[k] = c1[k]; m.tex_rgb2[k] = c2[k]; }
            }
            auto rp = floats(aget(it->second, "texrepeat", "1 1"));
            m.texrepeat[0] = rp[0]; m.texrepeat[1] = rp[1];
          }
        }
        if (a.count("rgba")) { auto v = floats(a.at("rgba")); for (int k = 0; k < 4; k++) rgba[k] = v[k]; }
        std::memcpy(g.rgba, rgba, sizeof rgba);
        V3 pos; Q4 quat;
        frame_of(a, pos, quat);
        g.mesh = -1;
        if (g.type == GEOM_PLANE) {
          auto s = floats(aget(a, "size", "0 0 0"));
          for (int k = 0; k < 3; k++) g.size[k] = s[k];
          g.pos = pos; g.quat = quat;
        } else {
          if (g.type == GEOM_MESH) {
            std::string mn = aget(a, "mesh", "");
            for (size_t i = 0; i < m.meshes.size(); i++) if (m.meshes[i].name == mn) g.mesh = (int)i;
            if (g.mesh < 0) throw std::runtime_error("geom references unknown mesh '" + mn + "'");
          } else {
            auto s = floats(aget(a, "size", ""));
            if (s.size() < 3) throw std::runtime_error("box geom needs size=\"x y z\"");
            for (int k = 0; k < 3; k++) g.size[k] = s[k];
            m.meshes.push_back(box_mesh("__box" + std::to_string(m.meshes.size()), s));
            g.mesh = (int)m.meshes.size() - 1;
          }
          const HostMesh& mp = m.meshes[g.mesh];
          g.pos = add(pos, mulv(q2m(quat), mp.pos));
          g.quat = qmul(quat, mp.quat);
          double density = std::stod(aget(a, "density", "1000"));
          double mass = density * mp.volume, sc = 1;
          if (a.count("mass")) sc = std::stod(a.at("mass")) / mass;
          g.mass = mass * sc;
          for (int k = 0; k < 3; k++) g.inertia[k] = density * mp.inertia_unit[k] * sc;
        }
        m.geom_names.push_back(aget(a, "name", ""));
        geoms.push_back(g);
        bodies[bid].geoms.push_back((int)geoms.size() - 1);
      } else if (e->tag == "joint" || e->tag == "freejoint") {
        JntTmp j{};
        j.body = bid;
        AttrMap a;
        if (e->tag == "freejoint") { j.type = JNT_FREE; a = amap(e); }
        else {
          a = dfl.resolve("joint", e, childclass);
          std::string t = aget(a, "type", "hinge");
          if (t == "free") j.type = JNT_FREE; else if (t == "slide") j.type = JNT_SLIDE; else if (t == "hinge") j.type = JNT_HINGE;
          else throw std::runtime_error("joint type '" + t + "' is not supported (free, slide, hinge only)");
        }
        auto p = floats(aget(a, "pos", "0 0 0")), ax = floats(aget(a, "axis", "0 0 1"));
        j.pos = {p[0], p[1], p[2]};
        double an = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        j.axis = {ax[0] / an, ax[1] / an, ax[2] / an};
        j.limited = aget(a, "limited", "false") == "true";
        auto r = floats(aget(a, "range", "0 0"));
        j.range[0] = r[0]; j.range[1] = r[1];
        j.armature = j.type == JNT_FREE ? 0 : std::stod(aget(a, "armature", "0"));
        j.damping = j.type == JNT_FREE ? 0 : std::stod(aget(a, "damping", "0"));
        m.jnt_names.push_back(aget(a, "name", ""));
        joints.push_back(j);
      } else if (e->tag == "inertial") {
        AttrMap a = amap(e);
        BodyTmp& b = bodies[bid];
        b.has_inertial = true;
        frame_of(a, b.ipos, b.iquat);
        b.imass = std::stod(aget(a, "mass", "0"));
        auto di = floats(aget(a, "diaginertia", "0 0 0"));
        b.iinertia = {di[0], di[1], di[2]};
      } else if (e->tag == "camera") {
        AttrMap a = amap(e);
        V3 pos; Q4 quat;
        frame_of(a, pos, quat);
        m.cam_names.push_back(aget(a, "name", ""));
        m.cam_bodyid.push_back(bid);
        for (int k = 0; k < 3; k++) m.cam_pos.push_back(pos[k]);
        for (int k = 0; k < 4; k++) m.cam_quat.push_back(quat[k]);
        m.cam_fovy.push_back(std::stod(aget(a, "fovy", "45")));
        m.cam_mode.push_back(aget(a, "mode", "fixed") == "targetbodycom" ? 1 : 0);
        m.cam_target.push_back(-1);
        if (a.count("target")) { m.cam_names.back() += "\t" + a.at("target"); }
      } else if (e->tag == "light") {
        AttrMap a = dfl.resolve("light", e, childclass);
        auto p = floats(aget(a, "pos", "0 0 0")), d = floats(aget(a, "dir", "0 0 -1"));
        auto df = floats(aget(a, "diffuse", "0.7 0.7 0.7")), am = floats(aget(a, "ambient", "0 0 0")), sp = floats(aget(a, "specular", "0.3 0.3 0.3"));
        m.light_bodyid.push_back(bid);
        m.light_directional.push_back(aget(a, "directional", "false") == "true");
        for (int k = 0; k < 3; k++) { m.light_pos.push_back(p[k]); m.light_dir.push_back(d[k]); m.light_diffuse.push_back(df[k]); m.light_ambient.push_back(am[k]); m.light_specular.push_back(sp[k]); }
      } else if (e->tag == "body") {
        AttrMap a = amap(e);
        BodyTmp b;
        b.parent = bid;
        frame_of(a, b.pos, b.quat);
        bodies.push_back(b);
        m.body_names.push_back(aget(a, "name", ""));
        int nb = (int)bodies.size() - 1;
        walk(e, nb, a.count("childclass") ? a.at("childclass") : childclass);
      } else if (e->tag == "site") {
        // sites carry no dynamics
      } else {
        throw std::runtime_error("unsupported element <" + e->tag + "> inside <worldbody>");
      }
    }
  };
  const XmlNode* wb = root->child("worldbody");
  if (!wb) throw std::runtime_error("scene has no <worldbody>");
  walk(wb, 0, "");

  // joints must be stored body by body (MJCF allows interleaving with child bodies; regroup)
  std::stable_sort(joints.begin(), joints.end(), [](const JntTmp& a, const JntTmp& b) { return a.body < b.body; });
  m.nbody = (int)bodies.size(); m.njnt = (int)joints.size(); m.ngeom = (int)geoms.size(); m.nmesh = (int)m.meshes.size();
  m.ncam = (int)m.cam_fovy.size(); m.nlight = (int)m.light_bodyid.size();
  for (int c = 0; c < m.ncam; c++) {  // resolve camera targets
    auto tab = m.cam_names[c].find('\t');
    if (tab != std::string::npos) {
      std::string tg = m.cam_names[c].substr(tab + 1);
      m.cam_names[c] = m.cam_names[c].substr(0, tab);
      for (int b = 0; b < m.nbody; b++) if (m.body_names[b] == tg) m.cam_target[c] = b;
    }
  }
  for (auto& b : bodies) {
    m.body_parentid.push_back(b.parent);
    for (int k = 0; k < 3; k++) m.body_pos.push_back(b.pos[k]);
    for (int k = 0; k < 4; k++) m.body_quat.push_back(b.quat[k]);
  }
  // inertial properties
  m.body_mass.assign(m.nbody, 0); m.body_ipos.assign(3 * m.nbody, 0); m.body_inertia.assign(3 * m.nbody, 0);
  m.body_iquat.assign(4 * m.nbody, 0);
  for (int b = 0; b < m.nbody; b++) m.body_iquat[4 * b] = 1;
  for (int b = 1; b < m.nbody; b++) {
    std::vector<int> solid;
    for (int g : bodies[b].geoms) if (geoms[g].type != GEOM_PLANE) solid.push_back(g);
    bool use = (inertiafromgeom == "true" && !solid.empty()) || (inertiafromgeom == "auto" && !bodies[b].has_inertial && !solid.empty());
    if (use) {
      double mass = 0;
      V3 com{0, 0, 0};
      for (int g : solid) { mass += geoms[g].mass; com = add(com, scl(geoms[g].pos, geoms[g].mass)); }
      com = scl(com, 1 / mass);
      double I[9] = {0};
      for (int g : solid) {
        M3 R = q2m(geoms[g].quat);
        V3 d = sub(geoms[g].pos, com);
        for (int a = 0; a < 3; a++)
          for (int c = 0; c < 3; c++) {
            double v = 0;
            for (int k = 0; k < 3; k++) v += R[3 * a + k] * geoms[g].inertia[k] * R[3 * c + k];
            I[3 * a + c] += v + geoms[g].mass * ((a == c ? dot(d, d) : 0) - d[a] * d[c]);
          }
      }
      double w[3];
      M3 V;
      eig3(I, w, V);
      Q4 q = m2q(V);
      m.body_mass[b] = mass;
      for (int k = 0; k < 3; k++) { m.body_ipos[3 * b + k] = com[k]; m.body_inertia[3 * b + k] = w[k]; }
      for (int k = 0; k < 4; k++) m.body_iquat[4 * b + k] = q[k];
    } else if (bodies[b].has_inertial) {
      m.body_mass[b] = bodies[b].imass;
      for (int k = 0; k < 3; k++) { m.body_ipos[3 * b + k] = bodies[b].ipos[k]; m.body_inertia[3 * b + k] = bodies[b].iinertia[k]; }
      for (int k = 0; k < 4; k++) m.body_iquat[4 * b + k] = bodies[b].iquat[k];
    }
  }
  // joints, dofs, qpos0
  m.body_jntadr.assign(m.nbody, -1); m.body_jntnum.assign(m.nbody, 0); m.body_dofadr.assign(m.nbody, -1); m.body_dofnum.assign(m.nbody, 0);
  std::vector<int> lastdof(m.nbody, -1);
  int jn = 0;
  for (int b = 0; b < m.nbody; b++) {
    int prev = b > 0 ? lastdof[bodies[b].parent] : -1;
    while (jn < m.njnt && joints[jn].body == b) {
      const JntTmp& j = joints[jn];
      if (m.body_jntadr[b] < 0) m.body_jntadr[b] = jn;
      m.body_jntnum[b]++;
      m.jnt_type.push_back(j.type); m.jnt_bodyid.push_back(b); m.jnt_limited.push_back(j.limited);
      m.jnt_qposadr.push_back((int)m.qpos0.size()); m.jnt_dofadr.push_back((int)m.dof_bodyid.size());
      if (m.body_dofadr[b] < 0) m.body_dofadr[b] = (int)m.dof_bodyid.size();
      for (int k = 0; k < 3; k++) { m.jnt_pos.push_back(j.pos[k]); m.jnt_axis.push_back(j.axis[k]); }
      m.jnt_range.push_back(j.range[0]); m.jnt_range.push_back(j.range[1]);
      int nd = 1;
      if (j.type == JNT_FREE) {
        if (bodies[b].parent != 0 || m.body_jntnum[b] != 1) throw std::runtime_error("a free joint must be the only joint of a top-level body");
        for (int k = 0; k < 3; k++) m.qpos0.push_back(bodies[b].pos[k]);
        for (int k = 0; k < 4; k++) m.qpos0.push_back(bodies[b].quat[k]);
        nd = 6;
      } else m.qpos0.push_back(0);
      for (int d = 0; d < nd; d++) {
        m.dof_bodyid.push_back(b); m.dof_jntid.push_back(jn); m.dof_parentid.push_back(prev);
        prev = (int)m.dof_bodyid.size() - 1;
        m.dof_armature.push_back(j.armature); m.dof_damping.push_back(j.damping);
      }
      m.body_dofnum[b] += nd;
      jn++;
    }
    lastdof[b] = prev;
  }
  m.nq = (int)m.qpos0.size(); m.nv = (int)m.dof_bodyid.size();
  m.body_weldid.assign(m.nbody, 0); m.body_rootid.assign(m.nbody, 0);
  for (int b = 1; b < m.nbody; b++) {
    m.body_weldid[b] = m.body_jntnum[b] > 0 ? b : m.body_weldid[m.body_parentid[b]];
    int r = b;
    while (m.body_parentid[r] > 0) r = m.body_parentid[r];
    m.body_rootid[b] = r;
  }
  // geoms
  for (auto& g : geoms) {
    m.geom_type.push_back(g.type); m.geom_bodyid.push_back(g.body); m.geom_meshid.push_back(g.mesh); m.geom_condim.push_back(g.condim);
    for (int k = 0; k < 3; k++) { m.geom_pos.push_back(g.pos[k]); m.geom_friction.push_back(g.friction[k]); m.geom_size.push_back(g.size[k]); }
    for (int k = 0; k < 4; k++) { m.geom_quat.push_back(g.quat[k]); m.geom_rgba.push_back(g.rgba[k]); m.geom_matprop.push_back(g.matprop[k]); }
    m.geom_margin.push_back(g.margin); m.geom_gap.push_back(g.gap); m.geom_mass.push_back(g.mass);
    for (int k = 0; k < 2; k++) m.geom_solref.push_back(g.solref[k]);
    for (int k = 0; k < 5; k++) m.geom_solimp.push_back(g.solimp[k]);
    double rb = 0;
    if (g.mesh >= 0) {
      const auto& hv = m.meshes[g.mesh].hull_verts;
      for (size_t i = 0; i < hv.size(); i += 3) rb = std::max(rb, std::sqrt(hv[i] * hv[i] + hv[i + 1] * hv[i + 1] + hv[i + 2] * hv[i + 2]));
    }
    m.geom_rbound.push_back(rb);
  }
  // candidate pairs: engine_collision_driver.c filter (contype/conaffinity, same weld body, weld parent), MuJoCo order
  struct P { int b1, b2, g1, g2; };
  std::vector<P> pairs;
  for (int g1 = 0; g1 < m.ngeom; g1++)
    for (int g2 = g1 + 1; g2 < m.ngeom; g2++) {
      const GeomTmp &A = geoms[g1], &B = geoms[g2];
      if (!((A.contype & B.conaffinity) || (B.contype & A.conaffinity))) continue;
      int w1 = m.body_weldid[A.body], w2 = m.body_weldid[B.body];
      if (w1 == w2) continue;
      int wp1 = w1 > 0 ? m.body_weldid[m.body_parentid[w1]] : 0, wp2 = w2 > 0 ? m.body_weldid[m.body_parentid[w2]] : 0;
      if (w1 != 0 && w2 != 0 && (w1 == wp2 || w2 == wp1)) continue;
      if (A.type == GEOM_PLANE && B.type == GEOM_PLANE) continue;
      bool sw = A.type > B.type;
      pairs.push_back({std::min(A.body, B.body), std::max(A.body, B.body), sw ? g2 : g1, sw ? g1 : g2});
    }
  std::stable_sort(pairs.begin(), pairs.end(), [](const P& a, const P& b) { return a.b1 != b.b1 ? a.b1 < b.b1 : a.b2 < b.b2; });
  for (auto& p : pairs) { m.pair_geom1.push_back(p.g1); m.pair_geom2.push_back(p.g2); }
  m.npair = (int)pairs.size();
  // actuators
  if (auto* act = root->child("actuator"))
    for (auto& e : act->kids) {
      if (e->tag != "motor") throw std::runtime_error("unsupported actuator <" + e->tag + "> (motor only)");
      AttrMap a = dfl.resolve("motor", e.get(), "");
      std::string jn2 = aget(a, "joint", "");
      int j = -1;
      for (int k = 0; k < m.njnt; k++) if (m.jnt_names[k] == jn2) j = k;
      if (j < 0) throw std::runtime_error("motor references unknown joint '" + jn2 + "'");
      m.act_dofid.push_back(m.jnt_dofadr[j]);
      m.act_gear.push_back(floats(aget(a, "gear", "1"))[0]);
      if (aget(a, "ctrllimited", "false") == "true") { auto r = floats(aget(a, "ctrlrange", "0 0")); m.act_ctrlrange.push_back(r[0]); m.act_ctrlrange.push_back(r[1]); }
      else { m.act_ctrlrange.push_back(-1e30); m.act_ctrlrange.push_back(1e30); }
    }
  m.nu = (int)m.act_dofid.size();
  // named bodies the environment needs (robot_env.py:65,87 ; actuator.py:139-150)
  auto bid = [&](const std::string& n) { for (int b = 0; b < m.nbody; b++) if (m.body_names[b] == n) return b; return -1; };
  m.body_ee = bid("ee"); m.body_object = bid("object");
  m.finger1[0] = bid("left_inner_knuckle"); m.finger1[1] = bid("left_inner_finger");
  m.finger2[0] = bid("right_inner_knuckle"); m.finger2[1] = bid("right_inner_finger");
  set_const(m);
  return m;
}

}  // namespace grs
