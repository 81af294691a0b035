// crt_host.cpp — host ingest of the product: OBJ/MTL, derived triangle/material/object data,
// camera matrix. Semantics follow the reference loader (include/OBJLoader.h:61-203,
// include/Loader.h:40-124); the implementation reads a memory-mapped file in newline-aligned chunks on all host
// threads and stitches them in file order, so the result - triangle order, material
// slots, the first error and its line number - is the one a sequential pass gives.
// Compiled with -ffp-contract=off: derived values equal the reference's Eigen host arithmetic.
#include "crt_host.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <string_view>
#include <thread>

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

// One triangle's derived data (Triangle.h:27,39 with Eigen's host arithmetic: cross by separate products, sum as
// x+(y+z)). false: a coordinate is not finite.
static bool triangle_record(const float vin[9], float v[9], float nrm[3], float* area) {
    for (int k = 0; k < 9; ++k) {
        if (!std::isfinite(vin[k])) return false;
        v[k] = vin[k] + 0.0f;                                     // -0 -> +0: min/max stay unambiguous
    }
    float ax = v[3] - v[0], ay = v[4] - v[1], az = v[5] - v[2];
    float bx = v[6] - v[0], by = v[7] - v[1], bz = v[8] - v[2];
    float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    float n2 = cx * cx + (cy * cy + cz * cz);
    nrm[0] = cx; nrm[1] = cy; nrm[2] = cz;
    if (n2 > 0.0f) { float l = sqrtf(n2); nrm[0] = cx / l; nrm[1] = cy / l; nrm[2] = cz / l; }
    *area = sqrtf(n2) * 0.5f;
    return true;
}

bool push_triangle(HostScene& s, const float vin[9], int mat, int obj) {
    float v[9], nrm[3], area;
    if (!triangle_record(vin, v, nrm, &area)) return false;
    s.verts.insert(s.verts.end(), v, v + 9);
    s.normal.push_back(nrm[0]); s.normal.push_back(nrm[1]); s.normal.push_back(nrm[2]);
    s.area.push_back(area);
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

// ---------------------------------------------------------------------------------------------
// OBJ text in newline-aligned chunks, one per host thread
// ---------------------------------------------------------------------------------------------
namespace {
struct FaceRec {
    long long v[3];        // the indices as written (1-based, or negative = relative to the vertices read so far)
    uint32_t line;         // line of the chunk (1-based)
    uint32_t nv_before;    // `v` lines of this chunk before the face
};
struct ShapeMark { std::string material; size_t face_local; };
struct ObjChunk {
    const char* begin = nullptr;
    const char* end = nullptr;
    std::vector<float> pos, uv;
    std::vector<FaceRec> faces;
    std::vector<ShapeMark> marks;
    std::string mtllib;
    bool has_mtllib = false;
    size_t lines = 0;
    size_t bad_line = 0;   // line of the chunk of the first face that does not parse (0: none)
};

void parse_obj_chunk(ObjChunk& ck) {
    const char* p = ck.begin;
    const char* fend = ck.end;
    ck.pos.reserve((size_t)(fend - p) / 24);
    ck.faces.reserve((size_t)(fend - p) / 48);
    while (p < fend) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(fend - p));
        const char* le = nl ? nl : fend;
        ++ck.lines;
        LineCursor c{p, le};
        std::string_view key = c.token();
        if (key == "v") {
            float x = c.number(), y = c.number(), z = c.number();
            ck.pos.push_back(x); ck.pos.push_back(y); ck.pos.push_back(z);
        } else if (key == "vt") {
            float a = c.number(), b = c.number();
            ck.uv.push_back(a); ck.uv.push_back(b);
        } else if (key == "f") {
            // "v", "v/vt", "v//vn", "v/vt/vn"; only the vertex index matters (Triangle.h:27-28
            // recomputes the normal) and only the first three corners are used (Loader.h:62-68).
            FaceRec fr;
            fr.line = (uint32_t)ck.lines;
            fr.nv_before = (uint32_t)(ck.pos.size() / 3);
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
                fr.v[nc++] = v;
            }
            // the first malformed face is remembered; the rest of the chunk is still read, so that the vertex count of the
            // file (which bounds positive indices) does not depend on where the chunks were cut
            if (bad || nc < 3) { if (!ck.bad_line) ck.bad_line = ck.lines; }
            else ck.faces.push_back(fr);
        } else if (key == "usemtl") {
            ck.marks.push_back(ShapeMark{std::string(c.token()), ck.faces.size()});
        } else if (key == "mtllib") {
            ck.mtllib = std::string(c.token());
            ck.has_mtllib = true;
        }
        p = nl ? nl + 1 : fend;
    }
}

unsigned ingest_threads(size_t bytes) {
    unsigned t = std::thread::hardware_concurrency();
    if (const char* e = getenv("CRT_INGEST_THREADS")) t = (unsigned)atoi(e);
    t = std::max(1u, std::min(t, 64u));
    size_t min_chunk = 256u << 10;                         // at least 256 KB of text per thread
    if (const char* e = getenv("CRT_INGEST_MIN_CHUNK")) min_chunk = (size_t)std::max(1L, atol(e));    // tests: chunk tiny files
    const size_t by_size = bytes / min_chunk;
    return (unsigned)std::max<size_t>(1, std::min<size_t>(t, by_size));
}

template <typename F>
void parallel_for(unsigned n, F&& body) {                   // body(k) for k in [0, n), one thread each (n is small)
    if (n <= 1) { if (n == 1) body(0u); return; }
    // an exception in a worker (std::bad_alloc from a chunk's vectors) is carried to the caller: it must not leave a
    // std::thread (std::terminate), and the caller's C-ABI guard turns it into a status
    std::vector<std::exception_ptr> err(n);
    auto run = [&body, &err](unsigned k) {
        try { body(k); } catch (...) { err[k] = std::current_exception(); }
    };
    std::vector<std::thread> th;
    th.reserve(n - 1);
    unsigned started = 1;
    try {
        for (; started < n; ++started) th.emplace_back(run, started);
    } catch (...) {                                          // thread creation failed: the rest runs here
        for (unsigned k = started; k < n; ++k) run(k);
    }
    run(0u);
    for (auto& t : th) t.join();
    for (unsigned k = 0; k < n; ++k) if (err[k]) std::rethrow_exception(err[k]);
}
}  // namespace

bool append_triangles(HostScene& s, const float* verts, const uint32_t* mat_id, const uint32_t* obj_id, size_t n, int mat0, int obj0) {
    const size_t t0 = s.n_tris();
    s.verts.resize(9 * (t0 + n)); s.normal.resize(3 * (t0 + n));
    s.area.resize(t0 + n); s.area_of_obj.resize(t0 + n, 0.0f);
    s.mat.resize(t0 + n); s.obj.resize(t0 + n);
    const unsigned T = ingest_threads(n * 36);
    std::vector<char> bad(T, 0);
    parallel_for(T, [&](unsigned k) {
        for (size_t i = n * k / T, e = n * (k + 1) / T; i < e; ++i) {
            const size_t t = t0 + i;
            if (!triangle_record(verts + 9 * i, &s.verts[9 * t], &s.normal[3 * t], &s.area[t])) { bad[k] = 1; return; }
            s.mat[t] = mat0 + (int)mat_id[i];
            s.obj[t] = obj0 + (int)obj_id[i];
        }
    });
    for (unsigned k = 0; k < T; ++k)
        if (bad[k]) {
            s.verts.resize(9 * t0); s.normal.resize(3 * t0); s.area.resize(t0); s.area_of_obj.resize(t0); s.mat.resize(t0); s.obj.resize(t0);
            return false;
        }
    return true;
}

static int load_obj_body(HostScene& s, const char* obj_path, const char* mtl_dir);

// The scene is left exactly as it was when the file is rejected (or an allocation fails): triangles, material slots,
// object count and the light tables are rolled back.
int load_obj(HostScene& s, const char* obj_path, const char* mtl_dir) {
    const size_t t0 = s.n_tris(), m0 = s.mats.size();
    const int o0 = s.n_objects;
    auto rollback = [&] {
        s.verts.resize(9 * t0); s.normal.resize(3 * t0); s.area.resize(t0); s.area_of_obj.resize(t0); s.mat.resize(t0); s.obj.resize(t0);
        s.mats.resize(m0);
        s.n_objects = o0;
        finish_objects(s);
    };
    int rc;
    try {
        rc = load_obj_body(s, obj_path, mtl_dir);
    } catch (...) {
        rollback();
        throw;
    }
    if (rc != CRT_OK) rollback();
    return rc;
}

static int load_obj_body(HostScene& s, const char* obj_path, const char* mtl_dir) {
    MappedFile f;
    if (!f.open(obj_path)) { set_error(std::string("Unable to open OBJ file: ") + obj_path); return CRT_ERR_IO; }
    // 1. chunks on line boundaries, parsed independently
    const unsigned T = ingest_threads(f.size);
    std::vector<ObjChunk> chunks(T);
    {
        const char* fend = f.data + f.size;
        const char* b = f.data;
        for (unsigned k = 0; k < T; ++k) {
            const char* e = fend;
            if (k + 1 < T) {
                const char* guess = f.data + (size_t)((unsigned long long)f.size * (k + 1) / T);
                if (guess < b) guess = b;
                const char* nl = (const char*)memchr(guess, '\n', (size_t)(fend - guess));
                e = nl ? nl + 1 : fend;
            }
            chunks[k].begin = b; chunks[k].end = e;
            b = e;
        }
    }
    parallel_for(T, [&](unsigned k) { parse_obj_chunk(chunks[k]); });
    // 2. stitch in file order: prefix counts, the first error, the first usemtl
    std::vector<size_t> nv0(T + 1, 0), nuv0(T + 1, 0), nf0(T + 1, 0), line0(T + 1, 0);
    for (unsigned k = 0; k < T; ++k) {
        nv0[k + 1] = nv0[k] + chunks[k].pos.size() / 3;
        nuv0[k + 1] = nuv0[k] + chunks[k].uv.size() / 2;
        nf0[k + 1] = nf0[k] + chunks[k].faces.size();
        line0[k + 1] = line0[k] + chunks[k].lines;
    }
    std::string mtl_name;
    for (unsigned k = 0; k < T; ++k) if (chunks[k].has_mtllib) mtl_name = chunks[k].mtllib;      // the last one wins
    size_t first_kept = (size_t)-1;                          // faces before the first usemtl are dropped (OBJLoader.h:120-123)
    for (unsigned k = 0; k < T && first_kept == (size_t)-1; ++k)
        if (!chunks[k].marks.empty()) first_kept = nf0[k] + chunks[k].marks[0].face_local;
    const size_t n_faces = nf0[T];
    const size_t n_kept = first_kept == (size_t)-1 ? 0 : n_faces - first_kept;
    std::vector<float> pos(3 * nv0[T]);                      // 3 per `v`
    std::vector<float> uv(2 * nuv0[T]);                      // 2 per `vt` (only read when a material has map_Kd)
    std::vector<uint32_t> idx(3 * n_kept);                   // 3 per kept face
    std::vector<size_t> err_line(T, 0);                      // global line of the chunk's first malformed face (0: none)
    parallel_for(T, [&](unsigned k) {
        ObjChunk& ck = chunks[k];
        if (!ck.pos.empty()) memcpy(&pos[3 * nv0[k]], ck.pos.data(), sizeof(float) * ck.pos.size());
        if (!ck.uv.empty()) memcpy(&uv[2 * nuv0[k]], ck.uv.data(), sizeof(float) * ck.uv.size());
        for (size_t fi = 0; fi < ck.faces.size(); ++fi) {
            const FaceRec& fr = ck.faces[fi];
            const long long nv = (long long)(nv0[k] + fr.nv_before);       // vertices read before this face
            const long long nv_all = (long long)nv0[T];
            uint32_t corner[3];
            bool bad = false;
            for (int c = 0; c < 3; ++c) {
                // OBJLoader.h:106 resolves the indices after the whole file is read: a positive index may refer to a
                // vertex defined later in the file; a negative one counts back from the vertices read so far
                const long long zero_based = fr.v[c] > 0 ? fr.v[c] - 1 : nv + fr.v[c];
                if (zero_based < 0 || zero_based >= (fr.v[c] > 0 ? nv_all : nv)) { bad = true; break; }
                corner[c] = (uint32_t)zero_based;
            }
            if (bad) { err_line[k] = line0[k] + fr.line; break; }
            const size_t g = nf0[k] + fi;
            if (first_kept != (size_t)-1 && g >= first_kept) memcpy(&idx[3 * (g - first_kept)], corner, sizeof(corner));
        }
        if (ck.bad_line && (err_line[k] == 0 || line0[k] + ck.bad_line < err_line[k])) err_line[k] = line0[k] + ck.bad_line;
    });
    for (unsigned k = 0; k < T; ++k)
        if (err_line[k]) {
            set_error(std::string(obj_path) + ":" + std::to_string(err_line[k]) + ": malformed face");
            return CRT_ERR_IO;
        }
    std::vector<Shape> shapes;
    for (unsigned k = 0; k < T; ++k)
        for (const ShapeMark& mk : chunks[k].marks) {
            Shape sh;
            sh.material = mk.material;
            sh.face_begin = sh.face_end = nf0[k] + mk.face_local - first_kept;
            if (!shapes.empty()) shapes.back().face_end = sh.face_begin;
            shapes.push_back(sh);
        }
    if (!shapes.empty()) shapes.back().face_end = n_kept;
    chunks.clear();

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
    bool textured = false;
    for (const Shape& sh : shapes) {
        auto it = table.find(sh.material);
        if (it != table.end() && !it->second.map_kd.empty() && sh.face_end > sh.face_begin) textured = true;
    }
    if (!textured) {
        // no map_Kd in use: a triangle's material and object are its shape's, so all triangles are filled in at once
        struct Span { size_t face_begin, face_end; int mat, obj; };
        std::vector<Span> spans;
        for (const Shape& sh : shapes) {
            HostMaterial m;
            auto it = table.find(sh.material);
            if (it != table.end()) m = it->second;
            m.name = sh.material;
            finish_material(m);
            const int mat = (int)s.mats.size();
            s.mats.push_back(m);
            if (sh.face_end == sh.face_begin) continue;
            spans.push_back(Span{sh.face_begin, sh.face_end, mat, s.n_objects++});
        }
        const size_t t0 = s.n_tris();
        s.verts.resize(9 * (t0 + n_kept)); s.normal.resize(3 * (t0 + n_kept));
        s.area.resize(t0 + n_kept); s.area_of_obj.resize(t0 + n_kept, 0.0f);
        s.mat.resize(t0 + n_kept); s.obj.resize(t0 + n_kept);
        std::vector<char> bad(T, 0);
        parallel_for(T, [&](unsigned k) {
            const size_t fb = n_kept * k / T, fe = n_kept * (k + 1) / T;
            size_t sp = 0;
            while (sp < spans.size() && spans[sp].face_end <= fb) ++sp;
            for (size_t fi = fb; fi < fe; ++fi) {
                while (spans[sp].face_end <= fi) ++sp;
                float v[9];
                for (int c = 0; c < 3; ++c) memcpy(v + 3 * c, &pos[3 * (size_t)idx[3 * fi + c]], 3 * sizeof(float));
                const size_t t = t0 + fi;
                if (!triangle_record(v, &s.verts[9 * t], &s.normal[3 * t], &s.area[t])) { bad[k] = 1; return; }
                s.mat[t] = spans[sp].mat;
                s.obj[t] = spans[sp].obj;
            }
        });
        for (unsigned k = 0; k < T; ++k)
            if (bad[k]) {
                s.verts.resize(9 * t0); s.normal.resize(3 * t0); s.area.resize(t0); s.area_of_obj.resize(t0); s.mat.resize(t0); s.obj.resize(t0);
                set_error(std::string(obj_path) + ": non-finite vertex coordinate");
                return CRT_ERR_IO;
            }
        finish_objects(s);
        return CRT_OK;
    }
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
