// crt_host.cpp — host ingest of the product: OBJ/MTL, derived triangle/material/object data,
// camera matrix. Semantics follow the reference loader (include/OBJLoader.h:61-203,
// include/Loader.h:40-124); the implementation is a single pass over a memory-mapped file.
// Compiled with -ffp-contract=off: derived values equal the reference's Eigen host arithmetic.
#include "crt_host.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <array>
#include <charconv>
#include <cmath>
#include <cstring>
#include <map>
#include <string_view>

namespace crt {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }

// ---------------------------------------------------------------------------------------------
// memory-mapped text file + tokenizer
// ---------------------------------------------------------------------------------------------
namespace {
struct MappedFile {
    const char* data = nullptr;
    size_t size = 0;
    int fd = -1;
    bool open(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); fd = -1; return false; }
        size = (size_t)st.st_size;
        if (size == 0) { data = ""; return true; }
        void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) { ::close(fd); fd = -1; return false; }
        madvise(p, size, MADV_SEQUENTIAL);
        data = (const char*)p;
        return true;
    }
    ~MappedFile() {
        if (data && size) munmap((void*)data, size);
        if (fd >= 0) ::close(fd);
    }
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

struct LineCursor {
    const char* p;
    const char* end;    // end of line (exclusive)
    std::string_view token() {
        while (p < end && is_space(*p)) ++p;
        const char* b = p;
        while (p < end && !is_space(*p)) ++p;
        return std::string_view(b, (size_t)(p - b));
    }
    float number() {            // like `stream >> float`: 0 when absent or malformed
        std::string_view t = token();
        if (t.empty()) return 0.0f;
        const char* b = t.data();
        const char* e = b + t.size();
        if (*b == '+') ++b;
        float v = 0.0f;
        auto r = std::from_chars(b, e, v);
        if (r.ec != std::errc()) return 0.0f;
        return v;
    }
};

struct Shape {                 // one `usemtl` occurrence (OBJLoader.h:131-137)
    std::string material;
    size_t face_begin = 0, face_end = 0;   // range in the shared index array (3 indices per face)
};

// Loader.h:85-103. The reference passes (&height, &width) to stbi_load's (x, y): its "width" is the image HEIGHT and
// its "height" the image WIDTH; u scales by (image height - 1), v by (image width - 1), and the texel offset is
// (v * image height + u) * channels. Kept as is (it is the usual lookup for square textures). uv of a corner = vt[vertex index].
inline float frac01(float x) {
    float ip;
    return modff(modff(x, &ip) + 1.0f, &ip);
}
bool texel_mean(const std::vector<uint8_t>& tex, int img_w, int img_h, int ch, const std::vector<float>& uv, const uint32_t corner[3],
                float kd[3]) {
    const int width = img_h, height = img_w;             // the reference's names
    float sum[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) {
        const size_t vi = corner[k];
        if (2 * vi + 1 >= uv.size()) return false;
        const int u = (int)(frac01(uv[2 * vi]) * (float)(width - 1));
        const int v = (int)(frac01(uv[2 * vi + 1]) * (float)(height - 1));
        const long long off = ((long long)v * width + u) * ch;
        float t[3];
        for (int c = 0; c < 3; ++c) {
            const long long o = off + c;
            t[c] = (float)((o >= 0 && (size_t)o < tex.size()) ? tex[(size_t)o] : 0) / 255.0f;
        }
        if (k == 0) { sum[0] = t[0]; sum[1] = t[1]; sum[2] = t[2]; }
        else { sum[0] += t[0]; sum[1] += t[1]; sum[2] += t[2]; }
    }
    for (int c = 0; c < 3; ++c) kd[c] = sum[c] / 3.0f;
    return true;
}
}  // namespace

void finish_material(HostMaterial& m) {
    const float eps = 0.00001f;                                   // Global.h:11
    m.has_emit = !(m.ke[0] < eps && m.ke[1] < eps && m.ke[2] < eps);
    m.mode = m.ns > 1.0f ? 1 : 0;
    m.probe_dtheta = m.probe_dphi = 0.0f;
    m.probe_shin = 1.0f;
    if (m.mode == 1) {
        // Render.cuh:296-300: c = (exp(25/ns) - 1)/(e - 1); lobe half-widths c*30deg, c*120deg
        float e = expf(25.0f / m.ns);
        float c = (float)((double)(e - 1.0f) / (M_E - 1.0));
        m.probe_dtheta = (float)((double)(c * 30.0f) * M_PI / 180.0);
        m.probe_dphi = (float)((double)(c * 120.0f) * M_PI / 180.0);
        m.probe_shin = (float)((double)log10f(m.ns) * 0.5 + 1.0);   // Render.cuh:306-307
    }
}

bool push_triangle(HostScene& s, const float vin[9], int mat, int obj) {
    float v[9];
    for (int k = 0; k < 9; ++k) {
        if (!std::isfinite(vin[k])) return false;
        v[k] = vin[k] + 0.0f;                                     // -0 -> +0: min/max stay unambiguous
    }
    // Triangle.h:27,39 with Eigen's host arithmetic: cross by separate products, sum as x+(y+z)
    float ax = v[3] - v[0], ay = v[4] - v[1], az = v[5] - v[2];
    float bx = v[6] - v[0], by = v[7] - v[1], bz = v[8] - v[2];
    float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    float n2 = cx * cx + (cy * cy + cz * cz);
    float nx = cx, ny = cy, nz = cz;
    if (n2 > 0.0f) { float l = sqrtf(n2); nx = cx / l; ny = cy / l; nz = cz / l; }
    s.verts.insert(s.verts.end(), v, v + 9);
    s.normal.push_back(nx); s.normal.push_back(ny); s.normal.push_back(nz);
    s.area.push_back(sqrtf(n2) * 0.5f);
    s.area_of_obj.push_back(0.0f);
    s.mat.push_back(mat);
    s.obj.push_back(obj);
    return true;
}

void finish_objects(HostScene& s) {
    std::vector<float> obj_area(s.n_objects, 0.0f);
    const size_t n = s.n_tris();
    for (size_t t = 0; t < n; ++t) obj_area[s.obj[t]] += s.area[t];          // Object.h:16-19
    for (size_t t = 0; t < n; ++t) s.area_of_obj[t] = obj_area[s.obj[t]];    // Object.h:20-23
    s.lights.clear();
    std::vector<int> light_of_obj(s.n_objects, -1);
    for (size_t t = 0; t < n; ++t) {
        if (!s.mats[s.mat[t]].has_emit) continue;
        int o = s.obj[t];
        if (light_of_obj[o] < 0) {
            light_of_obj[o] = (int)s.lights.size();
            s.lights.emplace_back();
            s.lights.back().area = obj_area[o];
        }
        s.lights[light_of_obj[o]].faces.push_back((int32_t)t);
    }
    // mis estimator tables: weights in double, summed in table order, rounded once
    auto lum = [](const float* c) { return fmaf(0.0722f, c[2], fmaf(0.7152f, c[1], 0.2126f * c[0])); };
    double W = 0.0;
    for (const HostLight& L : s.lights)
        for (int32_t f : L.faces) W += (double)s.area[f] * (double)lum(s.mats[s.mat[f]].ke);
    s.light_cdf.clear();
    double acc = 0.0;
    for (const HostLight& L : s.lights)
        for (int32_t f : L.faces) {
            acc += (double)s.area[f] * (double)lum(s.mats[s.mat[f]].ke);
            s.light_cdf.push_back(W > 0.0 ? (float)(acc / W) : 1.0f);
        }
    if (!s.light_cdf.empty()) s.light_cdf.back() = 1.0f;
    for (HostMaterial& m : s.mats) m.pdf_area = (m.has_emit && W > 0.0) ? (float)((double)lum(m.ke) / W) : 0.0f;
}

int load_obj(HostScene& s, const char* obj_path, const char* mtl_dir) {
    MappedFile f;
    if (!f.open(obj_path)) { set_error(std::string("Unable to open OBJ file: ") + obj_path); return CRT_ERR_IO; }
    std::vector<float> pos;                  // 3 per `v`
    std::vector<float> uv;                   // 2 per `vt` (only read when a material has map_Kd)
    std::vector<uint32_t> idx;               // 3 per kept face
    std::vector<Shape> shapes;
    std::string mtl_name;
    pos.reserve(f.size / 24);
    idx.reserve(f.size / 48);
    const char* p = f.data;
    const char* fend = f.data + f.size;
    size_t line_no = 0;
    while (p < fend) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(fend - p));
        const char* le = nl ? nl : fend;
        ++line_no;
        LineCursor c{p, le};
        std::string_view key = c.token();
        if (key == "v") {
            float x = c.number(), y = c.number(), z = c.number();
            pos.push_back(x); pos.push_back(y); pos.push_back(z);
        } else if (key == "vt") {
            float a = c.number(), b = c.number();
            uv.push_back(a); uv.push_back(b);
        } else if (key == "f") {
            // "v", "v/vt", "v//vn", "v/vt/vn"; only the vertex index matters (Triangle.h:27-28
            // recomputes the normal) and only the first three corners are used (Loader.h:62-68).
            uint32_t corner[3];
            int nc = 0;
            bool bad = false;
            for (;;) {
                std::string_view t = c.token();
                if (t.empty()) break;
                if (nc >= 3) continue;
                long long v = 0;
                const char* b = t.data();
                const char* e = b + t.size();
                const char* slash = (const char*)memchr(b, '/', t.size());
                if (slash) e = slash;
                if (b < e && *b == '+') ++b;
                auto r = std::from_chars(b, e, v);
                if (r.ec != std::errc()) { bad = true; break; }
                long long nv = (long long)(pos.size() / 3);
                long long zero_based = v > 0 ? v - 1 : nv + v;          // OBJLoader.h:106
                if (zero_based < 0 || zero_based >= nv) { bad = true; break; }
                corner[nc++] = (uint32_t)zero_based;
            }
            if (bad || nc < 3) {
                set_error(std::string(obj_path) + ":" + std::to_string(line_no) + ": malformed face");
                return CRT_ERR_IO;
            }
            if (!shapes.empty()) {                                       // OBJLoader.h:120-123
                idx.push_back(corner[0]); idx.push_back(corner[1]); idx.push_back(corner[2]);
                shapes.back().face_end = idx.size() / 3;
            }
        } else if (key == "usemtl") {
            Shape sh;
            sh.material = std::string(c.token());
            sh.face_begin = sh.face_end = idx.size() / 3;
            shapes.push_back(sh);
        } else if (key == "mtllib") {
            mtl_name = std::string(c.token());
        }
        p = nl ? nl + 1 : fend;
    }

    // MTL: newmtl / Kd / Ks / Ke / Ns (OBJLoader.h:154-200); other keys are ignored like there.
    std::map<std::string, HostMaterial> table;
    {
        std::string mtl_path = std::string(mtl_dir) + "/" + mtl_name;    // OBJLoader.h:129
        MappedFile m;
        if (!m.open(mtl_path.c_str())) { set_error("Unable to open MTL file: " + mtl_path); return CRT_ERR_IO; }
        HostMaterial* cur = nullptr;
        const char* q = m.data;
        const char* mend = m.data + m.size;
        while (q < mend) {
            const char* nl = (const char*)memchr(q, '\n', (size_t)(mend - q));
            const char* le = nl ? nl : mend;
            LineCursor c{q, le};
            std::string_view key = c.token();
            if (key == "newmtl") {
                std::string name(c.token());
                cur = &table[name];
                cur->name = name;
            } else if (cur && key == "Kd") { cur->kd[0] = c.number(); cur->kd[1] = c.number(); cur->kd[2] = c.number(); }
            else if (cur && key == "Ks") { cur->ks[0] = c.number(); cur->ks[1] = c.number(); cur->ks[2] = c.number(); }
            else if (cur && key == "Ke") { cur->ke[0] = c.number(); cur->ke[1] = c.number(); cur->ke[2] = c.number(); }
            else if (cur && key == "Ns") { cur->ns = c.number(); }
            else if (cur && key == "map_Kd") { cur->map_kd = std::string(mtl_dir) + "/" + std::string(c.token()); }   // OBJLoader.h:184-193
            q = nl ? nl + 1 : mend;
        }
    }

    // One material slot and (if it has faces) one object per shape, in file order (main.cu:131-144).
    for (const Shape& sh : shapes) {
        HostMaterial m;
        auto it = table.find(sh.material);
        if (it != table.end()) m = it->second;
        m.name = sh.material;
        finish_material(m);
        int mat = (int)s.mats.size();
        s.mats.push_back(m);
        if (sh.face_end == sh.face_begin) continue;
        // map_Kd (Loader.h:55-59,78-105): Kd of a triangle = mean of the three texels at its corners. A texture
        // file that cannot be opened leaves the plain Kd (stbi_load returns null there); one that exists but
        // cannot be decoded is an error.
        int tw = 0, th = 0, tch = 0;
        std::vector<uint8_t> tex;
        std::map<std::array<uint32_t, 3>, int> tex_mats;
        if (!m.map_kd.empty()) {
            int rc = read_image(m.map_kd.c_str(), &tw, &th, &tch, tex);
            if (rc == 2) { set_error("map_Kd: unsupported or corrupt image (PNG and binary PNM are read): " + m.map_kd); return CRT_ERR_IO; }
            if (rc != 0) tex.clear();
        }
        int obj = s.n_objects++;
        for (size_t fi = sh.face_begin; fi < sh.face_end; ++fi) {
            float v[9];
            for (int k = 0; k < 3; ++k) memcpy(v + 3 * k, &pos[3 * (size_t)idx[3 * fi + k]], 3 * sizeof(float));
            int tri_mat = mat;
            if (!tex.empty()) {
                float kd[3];
                if (!texel_mean(tex, tw, th, tch, uv, &idx[3 * fi], kd)) {
                    set_error(std::string(obj_path) + ": map_Kd needs one vt per vertex (the reference indexes vt by the vertex index, Loader.h:81-83)");
                    return CRT_ERR_IO;
                }
                std::array<uint32_t, 3> key;
                memcpy(key.data(), kd, sizeof(kd));
                auto found = tex_mats.find(key);
                if (found == tex_mats.end()) {
                    HostMaterial tm = m;
                    memcpy(tm.kd, kd, sizeof(kd));
                    found = tex_mats.emplace(key, (int)s.mats.size()).first;
                    s.mats.push_back(tm);
                }
                tri_mat = found->second;
            }
            if (!push_triangle(s, v, tri_mat, obj)) {
                set_error(std::string(obj_path) + ": non-finite vertex coordinate");
                return CRT_ERR_IO;
            }
        }
    }
    finish_objects(s);
    return CRT_OK;
}

void inverse_view_matrix(const float eye[3], const float lookat[3], const float up[3], float out9[9]) {
    // Camera.h:9-36 with Eigen's host arithmetic
    auto norm3 = [](float* v) {
        float n2 = v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]);
        if (n2 > 0.0f) { float l = sqrtf(n2); v[0] /= l; v[1] /= l; v[2] /= l; }
    };
    auto cross3 = [](const float* a, const float* b, float* o) {
        o[0] = a[1] * b[2] - a[2] * b[1];
        o[1] = a[2] * b[0] - a[0] * b[2];
        o[2] = a[0] * b[1] - a[1] * b[0];
    };
    float f[3] = {lookat[0] - eye[0], lookat[1] - eye[1], lookat[2] - eye[2]};
    norm3(f);
    float r[3], u[3];
    cross3(up, f, r);
    norm3(r);
    cross3(f, r, u);
    norm3(u);
    for (int k = 0; k < 3; ++k) { out9[3 * k + 0] = r[k]; out9[3 * k + 1] = u[k]; out9[3 * k + 2] = f[k]; }
}

}  // namespace crt
