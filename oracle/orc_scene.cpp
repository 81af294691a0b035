// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h). Never linked into the product.
// Restatement of the reference's host-side ingest; every routine cites the lines it follows.
#include "orc_scene.h"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>

namespace orc {

// Host arithmetic of the reference is Eigen on the CPU: separate multiplies and adds, and
// Eigen's fixed-size-3 reduction associates as x + (y + z)
// (include/Eigen/src/Core/Redux.h, redux_novec_unroller<..., 0, 3>).
static inline float hdot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
static inline V3 hcross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline V3 hnormalize(V3 a) {
    float n = hdot(a, a);
    if (n > 0.0f) return a / sqrtf(n);
    return a;
}

void finish_triangle(Tri& t) {
    // Triangle.h:26-39
    t.center = ((t.v1 + t.v2) + t.v3) / 3.0f;
    V3 c = hcross(t.v2 - t.v1, t.v3 - t.v1);
    t.normal = hnormalize(c);
    t.hi = V3{std::fmax(std::fmax(t.v1.x, t.v2.x), t.v3.x), std::fmax(std::fmax(t.v1.y, t.v2.y), t.v3.y),
              std::fmax(std::fmax(t.v1.z, t.v2.z), t.v3.z)};
    t.lo = V3{std::fmin(std::fmin(t.v1.x, t.v2.x), t.v3.x), std::fmin(std::fmin(t.v1.y, t.v2.y), t.v3.y),
              std::fmin(std::fmin(t.v1.z, t.v2.z), t.v3.z)};
    t.area = sqrtf(hdot(c, c)) * 0.5f;
    t.area_of_obj = 0.0f;
}

void finish_material(Material& m) {
    const float eps = 0.00001f;  // Global.h:11
    m.has_emit = !(m.ke.x < eps && m.ke.y < eps && m.ke.z < eps);   // Material.h:36-39
    m.mode = m.ns > 1.0f ? SPECULAR : DIFFUSE;                      // Loader.h:107
    if (m.mode == SPECULAR) {
        // Render.cuh:296-300 with Global.h:21-22 (DELTA_THETA = 30*M_PI/180, DELTA_PHI = 120*M_PI/180)
        float e = expf(25.0f / m.ns);
        float c = (float)((double)(e - 1.0f) / (M_E - 1.0));
        m.probe_dtheta = (float)((double)(c * 30.0f) * M_PI / 180.0);
        m.probe_dphi = (float)((double)(c * 120.0f) * M_PI / 180.0);
        // Render.cuh:306-307
        m.probe_shin = (float)((double)log10f(m.ns) * 0.5 + 1.0);
    }
}

void finish_objects(Scene& s) {
    // Object.h:12-26: area_of_obj is the float sum of the group's triangle areas, in order.
    std::vector<float> area(s.n_objects, 0.0f);
    for (auto& t : s.tris) area[t.obj] += t.area;
    for (auto& t : s.tris) t.area_of_obj = area[t.obj];
    s.lights.clear();
    // Scene.h:38-42: light objects in the order they were added.
    std::map<int, int> light_of_obj;
    for (int i = 0; i < (int)s.tris.size(); ++i) {
        const Tri& t = s.tris[i];
        if (!s.mats[t.mat].has_emit) continue;
        auto it = light_of_obj.find(t.obj);
        if (it == light_of_obj.end()) {
            it = light_of_obj.emplace(t.obj, (int)s.lights.size()).first;
            s.lights.push_back(LightObj{{}, area[t.obj]});
        }
        s.lights[it->second].tris.push_back(i);
    }
    // mis estimator tables (DESIGN.md "mis"): weights in double, sequential, then rounded once
    s.light_flat.clear();
    s.light_cdf.clear();
    for (const LightObj& L : s.lights)
        for (int f : L.tris) s.light_flat.push_back(f);
    double W = 0.0;
    for (int f : s.light_flat) W += (double)s.tris[f].area * (double)lumf(s.mats[s.tris[f].mat].ke);
    double acc = 0.0;
    for (int f : s.light_flat) {
        acc += (double)s.tris[f].area * (double)lumf(s.mats[s.tris[f].mat].ke);
        s.light_cdf.push_back(W > 0.0 ? (float)(acc / W) : 1.0f);
    }
    if (!s.light_cdf.empty()) s.light_cdf.back() = 1.0f;
    for (Material& m : s.mats) m.pdf_area = (m.has_emit && W > 0.0) ? (float)((double)lumf(m.ke) / W) : 0.0f;
}

namespace {
struct Shape {                       // OBJLoader.h:12-39
    std::string material_id;
    std::vector<std::vector<uint64_t>> vs;
    float kd[3] = {0, 0, 0}, ks[3] = {0, 0, 0}, ke[3] = {0, 0, 0};
    float ns = 1.0f;                 // reference leaves _ns uninitialised without an Ns line; 1 here
};
}  // namespace

bool load_obj(Scene& s, const std::string& obj_path, const std::string& mtl_dir) {
    std::ifstream obj(obj_path);
    if (!obj.is_open()) { s.error = "Unable to open OBJ file: " + obj_path; return false; }
    std::vector<V3> vertices;
    std::vector<Shape> shapes;
    std::map<std::string, std::vector<uint64_t>> mts;
    std::string mtl_path, line;
    size_t n_vt = 0, n_vn = 0;
    while (std::getline(obj, line)) {              // OBJLoader.h:71-139
        std::istringstream ls(line);
        std::string prefix;
        ls >> prefix;
        if (prefix == "v") {
            V3 v{0, 0, 0};
            ls >> v.x >> v.y >> v.z;
            vertices.push_back(V3{v.x + 0.0f, v.y + 0.0f, v.z + 0.0f});
        } else if (prefix == "vn") {
            ++n_vn;
        } else if (prefix == "vt") {
            ++n_vt;
        } else if (prefix == "f") {
            std::vector<uint64_t> vi;
            std::string tok;
            while (ls >> tok) {                     // OBJLoader.h:98-118: "v/vt/vn", 1-based
                std::string first = tok.substr(0, tok.find('/'));
                uint64_t idx = std::stoull(first);
                vi.push_back(idx > 0 ? idx - 1 : vertices.size() + idx);
            }
            if (!shapes.empty()) shapes.back().vs.push_back(vi);   // faces before usemtl are dropped
        } else if (prefix == "mtllib") {
            std::string name;
            ls >> name;
            mtl_path = mtl_dir + "/" + name;
        } else if (prefix == "usemtl") {
            std::string id;
            ls >> id;
            mts[id].push_back(shapes.size());
            Shape sh;
            sh.material_id = id;
            shapes.push_back(sh);
        }
    }
    std::ifstream mtl(mtl_path);
    if (!mtl.is_open()) { s.error = "Unable to open MTL file: " + mtl_path; return false; }
    std::vector<uint64_t> ids;
    while (std::getline(mtl, line)) {              // OBJLoader.h:154-200
        std::istringstream ls(line);
        std::string prefix;
        ls >> prefix;
        if (prefix == "newmtl") {
            std::string id;
            ls >> id;
            ids = mts[id];
        } else if (prefix == "Kd") {
            float k[3] = {0, 0, 0};
            ls >> k[0] >> k[1] >> k[2];
            for (auto i : ids) { shapes[i].kd[0] = k[0]; shapes[i].kd[1] = k[1]; shapes[i].kd[2] = k[2]; }
        } else if (prefix == "Ks") {     // dropped by the reference (Loader.h:45-47,107); kept for the mis estimator only
            float k[3] = {0, 0, 0};
            ls >> k[0] >> k[1] >> k[2];
            for (auto i : ids) { shapes[i].ks[0] = k[0]; shapes[i].ks[1] = k[1]; shapes[i].ks[2] = k[2]; }
        } else if (prefix == "Ke") {
            float k[3] = {0, 0, 0};
            ls >> k[0] >> k[1] >> k[2];
            for (auto i : ids) { shapes[i].ke[0] = k[0]; shapes[i].ke[1] = k[1]; shapes[i].ke[2] = k[2]; }
        } else if (prefix == "Ns") {
            float ns = 1.0f;
            ls >> ns;
            for (auto i : ids) shapes[i].ns = ns;
        }
        // Ka, Tr, Ni, illum are ignored;
        // map_Kd (Loader.h:55-59,78-105) is out of scope (no shipped scene has one).
    }
    // Loader.h:40-124 + main.cu:131-144: one Object per shape, in shape order.
    for (const Shape& sh : shapes) {
        Material m;
        m.kd = V3{sh.kd[0], sh.kd[1], sh.kd[2]};
        m.ks = V3{sh.ks[0], sh.ks[1], sh.ks[2]};
        m.ke = V3{sh.ke[0], sh.ke[1], sh.ke[2]};
        m.ns = sh.ns;
        m.name = sh.material_id;
        finish_material(m);
        int mat = (int)s.mats.size();
        s.mats.push_back(m);
        if (sh.vs.empty()) continue;                // main.cu:134,139: empty lists add no Object
        int obj = s.n_objects++;
        for (const auto& f : sh.vs) {
            if (f.size() < 3) { s.error = "face with fewer than 3 vertices"; return false; }
            for (int k = 0; k < 3; ++k)
                if (f[k] >= vertices.size()) { s.error = "face index out of range"; return false; }
            Tri t;                                  // Loader.h:62-68: only the first three indices
            t.v1 = vertices[f[0]];
            t.v2 = vertices[f[1]];
            t.v3 = vertices[f[2]];
            t.mat = mat;
            t.obj = obj;
            finish_triangle(t);
            s.tris.push_back(t);
        }
    }
    (void)n_vt; (void)n_vn;
    finish_objects(s);
    return true;
}

void inverse_view_matrix(const float eye[3], const float lookat[3], const float up[3], float out9[9]) {
    // Camera.h:9-36
    V3 e{eye[0], eye[1], eye[2]}, l{lookat[0], lookat[1], lookat[2]}, u0{up[0], up[1], up[2]};
    V3 f = hnormalize(l - e);
    V3 r = hnormalize(hcross(u0, f));
    V3 u = hnormalize(hcross(f, r));
    out9[0] = r.x; out9[1] = u.x; out9[2] = f.x;
    out9[3] = r.y; out9[4] = u.y; out9[5] = f.y;
    out9[6] = r.z; out9[7] = u.z; out9[8] = f.z;
}

}  // namespace orc
