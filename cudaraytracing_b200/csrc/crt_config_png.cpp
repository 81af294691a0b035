// crt_config_png.cpp — the two file formats at the edges of the path:
//   config.json in (schema of src/main.cu:67-90), RGB8 PNG out (Render.cuh:489-493).
#include "crt_host.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <thread>
#include <vector>

namespace crt {

// ---------------------------------------------------------------------------------------------
// Minimal JSON reader (objects, arrays, strings, numbers, true/false/null), enough for the
// reference's config schema. Missing keys are errors, like nlohmann's operator[] + conversion.
// ---------------------------------------------------------------------------------------------
namespace {
struct JValue {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;
    const JValue* get(const char* key) const {
        for (auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char* p;
    const char* end;
    std::string err;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool fail(const char* m) { if (err.empty()) err = m; return false; }
    bool parse_string(std::string& out) {
        if (p >= end || *p != '"') return fail("expected string");
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("bad escape");
                switch (*p) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (end - p < 5) return fail("bad \\u escape");
                        unsigned cp = (unsigned)strtoul(std::string(p + 1, 4).c_str(), nullptr, 16);
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        p += 4;
                        break;
                    }
                    default: out += *p;      // \" \\ \/
                }
                ++p;
            } else out += *p++;
        }
        if (p >= end) return fail("unterminated string");
        ++p;
        return true;
    }
    int depth = 0;                    // nesting of the value being parsed: a config is 3 levels deep; 64 bounds the recursion
    struct DepthGuard { int& d; explicit DepthGuard(int& x) : d(x) { ++d; } ~DepthGuard() { --d; } };
    bool parse(JValue& v) {
        DepthGuard guard(depth);
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end");
        if (*p == '{') {
            v.kind = JValue::Obj;
            ++p; ws();
            if (p < end && *p == '}') { ++p; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!parse_string(k)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                ++p;
                JValue child;
                if (!parse(child)) return false;
                v.obj.emplace_back(std::move(k), std::move(child));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.kind = JValue::Arr;
            ++p; ws();
            if (p < end && *p == ']') { ++p; return true; }
            for (;;) {
                JValue child;
                if (!parse(child)) return false;
                v.arr.push_back(std::move(child));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') { v.kind = JValue::Str; return parse_string(v.str); }
        if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JValue::Bool; v.b = true; p += 4; return true; }
        if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JValue::Bool; v.b = false; p += 5; return true; }
        if (end - p >= 4 && !strncmp(p, "null", 4)) { v.kind = JValue::Null; p += 4; return true; }
        char* q = nullptr;
        std::string tmp(p, (size_t)std::min<ptrdiff_t>(end - p, 64));
        double d = strtod(tmp.c_str(), &q);
        if (q == tmp.c_str()) return fail("unexpected character");
        v.kind = JValue::Num;
        v.num = d;
        p += (q - tmp.c_str());
        return true;
    }
};

bool need_num(const JValue& o, const char* key, double* out, std::string& err) {
    const JValue* v = o.get(key);
    if (!v || v->kind != JValue::Num) { err = std::string("config: missing or non-numeric key \"") + key + "\""; return false; }
    *out = v->num;
    return true;
}
bool need_vec3(const JValue& o, const char* key, float out[3], std::string& err) {
    const JValue* v = o.get(key);
    if (!v || v->kind != JValue::Obj) { err = std::string("config: missing object \"") + key + "\""; return false; }
    double x, y, z;
    if (!need_num(*v, "x", &x, err) || !need_num(*v, "y", &y, err) || !need_num(*v, "z", &z, err)) return false;
    out[0] = (float)x; out[1] = (float)y; out[2] = (float)z;
    return true;
}
}  // namespace

int load_config(const char* path, crt_config* out) {
    FILE* f = fopen(path, "rb");
    if (!f) { set_error(std::string("config: cannot open ") + path); return CRT_ERR_IO; }
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    fclose(f);
    JParser jp{text.data(), text.data() + text.size(), {}};
    JValue root;
    if (!jp.parse(root) || root.kind != JValue::Obj) { set_error(std::string("config: JSON parse error: ") + (jp.err.empty() ? "not an object" : jp.err)); return CRT_ERR_IO; }
    memset(out, 0, sizeof(*out));
    std::string err;
    const JValue* paths = root.get("OBJ_paths");                      // main.cu:74-77
    if (!paths || paths->kind != JValue::Arr) { set_error("config: missing array \"OBJ_paths\""); return CRT_ERR_IO; }
    if (paths->arr.size() > CRT_MAX_OBJ_PATHS) { set_error("config: too many OBJ_paths"); return CRT_ERR_INVALID; }
    for (const JValue& e : paths->arr) {
        const JValue* op = e.get("OBJ_path");
        const JValue* md = e.get("MTL_dir");
        if (!op || !md || op->kind != JValue::Str || md->kind != JValue::Str) { set_error("config: OBJ_paths entries need OBJ_path and MTL_dir strings"); return CRT_ERR_IO; }
        if (op->str.size() >= CRT_PATH_LEN || md->str.size() >= CRT_PATH_LEN) { set_error("config: path too long"); return CRT_ERR_INVALID; }
        strcpy(out->obj_path[out->n_obj], op->str.c_str());
        strcpy(out->mtl_dir[out->n_obj], md->str.c_str());
        out->n_obj++;
    }
    double d;
    if (!need_vec3(root, "eye_pos", out->eye_pos, err) || !need_vec3(root, "lookat", out->lookat, err) ||
        !need_vec3(root, "up", out->up, err)) { set_error(err); return CRT_ERR_IO; }
    if (!need_num(root, "fov_y", &d, err)) { set_error(err); return CRT_ERR_IO; } out->fov_y = (float)d;
    if (!need_num(root, "width", &d, err)) { set_error(err); return CRT_ERR_IO; } out->width = (uint32_t)d;
    if (!need_num(root, "height", &d, err)) { set_error(err); return CRT_ERR_IO; } out->height = (uint32_t)d;
    if (!need_num(root, "bvh_thresh_n", &d, err)) { set_error(err); return CRT_ERR_IO; } out->bvh_thresh_n = (uint32_t)d;
    if (!need_num(root, "P_RR", &d, err)) { set_error(err); return CRT_ERR_IO; } out->p_rr = (float)d;
    if (!need_num(root, "spp", &d, err)) { set_error(err); return CRT_ERR_IO; } out->spp = (uint32_t)d;
    if (!need_num(root, "light_sample_n", &d, err)) { set_error(err); return CRT_ERR_IO; } out->light_sample_n = (uint32_t)d;
    // optional extensions; shipped configs run unchanged
    if (const JValue* v = root.get("seed")) if (v->kind == JValue::Num) out->seed = (uint32_t)v->num;
    if (const JValue* v = root.get("estimator")) {
        if (v->kind == JValue::Str && v->str == "mis") out->estimator = CRT_ESTIMATOR_MIS;
        else if (v->kind == JValue::Num) out->estimator = (uint32_t)v->num;
    }
    return CRT_OK;
}

// ---------------------------------------------------------------------------------------------
// PNG: 8-bit RGB, filter 0 on every row, one zlib stream.
// ---------------------------------------------------------------------------------------------
namespace {
void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}
void put_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t n) {
    put_be32(out, (uint32_t)n);
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    uint32_t crc = (uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start));
    put_be32(out, crc);
}
}  // namespace

// RGB8 PNG, filter 0 on every row, deflate level 6. The scanlines are cut into bands of about 1 MB (a function of the
// image size only, so the file does not depend on the machine); every band is deflated on its own as a raw stream that
// ends on a byte boundary (Z_SYNC_FLUSH; the last one with Z_FINISH) and the pieces are concatenated behind one zlib
// header, with the Adler-32 of the whole combined from the bands' (the pigz construction). A 3840x2160 frame takes
// ~1.6 s on one thread; the bands are compressed on all host threads.
int write_png(const char* path, const uint8_t* rgb8, uint32_t width, uint32_t height) {
    if (!path || !rgb8 || width == 0 || height == 0) { set_error("write_png: invalid argument"); return CRT_ERR_INVALID; }
    const size_t row = (size_t)width * 3;
    const uint32_t band_rows = (uint32_t)std::max<size_t>(1, (size_t)(1u << 20) / (row + 1));
    const uint32_t n_bands = (height + band_rows - 1) / band_rows;
    struct Band { std::vector<uint8_t> z; uLong adler = 1; size_t raw_len = 0; bool ok = false; };
    std::vector<Band> bands(n_bands);
    std::atomic<uint32_t> next(0);
    auto work = [&]() {
        std::vector<uint8_t> raw;
        for (;;) {
            const uint32_t b = next.fetch_add(1);
            if (b >= n_bands) return;
            const uint32_t y0 = b * band_rows, y1 = std::min(height, y0 + band_rows);
            raw.resize((row + 1) * (size_t)(y1 - y0));
            for (uint32_t y = y0; y < y1; ++y) {
                raw[(row + 1) * (size_t)(y - y0)] = 0;
                memcpy(&raw[(row + 1) * (size_t)(y - y0) + 1], rgb8 + row * y, row);
            }
            Band& bd = bands[b];
            bd.raw_len = raw.size();
            bd.adler = adler32(adler32(0L, Z_NULL, 0), raw.data(), (uInt)raw.size());
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) continue;
            bd.z.resize(deflateBound(&zs, (uLong)raw.size()) + 16);
            zs.next_in = raw.data(); zs.avail_in = (uInt)raw.size();
            zs.next_out = bd.z.data(); zs.avail_out = (uInt)bd.z.size();
            const bool last = b + 1 == n_bands;
            const int rc = deflate(&zs, last ? Z_FINISH : Z_SYNC_FLUSH);
            bd.ok = (last ? rc == Z_STREAM_END : rc == Z_OK) && zs.avail_in == 0;
            bd.z.resize(bd.z.size() - zs.avail_out);
            deflateEnd(&zs);
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (const char* e = getenv("CRT_INGEST_THREADS")) nt = (unsigned)atoi(e);      // one knob for the host-side thread count
    nt = std::max(1u, std::min(std::min(nt, 64u), n_bands));
    {
        std::vector<std::thread> th;
        for (unsigned k = 1; k < nt; ++k) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x9c);                          // zlib header: deflate, 32 KB window, default level
    uLong adler = adler32(0L, Z_NULL, 0);
    for (const Band& bd : bands) {
        if (!bd.ok) { set_error("write_png: deflate failed"); return CRT_ERR_NOMEM; }
        z.insert(z.end(), bd.z.begin(), bd.z.end());
        adler = adler32_combine(adler, bd.adler, (z_off_t)bd.raw_len);
    }
    put_be32(z, (uint32_t)adler);
    std::vector<uint8_t> out;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    out.insert(out.end(), sig, sig + 8);
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width); put_be32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    put_chunk(out, "IHDR", ihdr.data(), ihdr.size());
    put_chunk(out, "IDAT", z.data(), z.size());
    put_chunk(out, "IEND", nullptr, 0);
    FILE* f = fopen(path, "wb");
    if (!f) { set_error(std::string("write_png: cannot open ") + path); return CRT_ERR_IO; }
    size_t w = fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    if (w != out.size()) { set_error("write_png: short write"); return CRT_ERR_IO; }
    return CRT_OK;
}

// ---------------------------------------------------------------------------------------------
// Texture images for map_Kd (the reference calls stbi_load(path, &x, &y, &comp, 0), Loader.h:55-59):
// PNG (non-interlaced; grey / RGB / palette, with or without alpha or tRNS; 1-16 bits) and binary PNM
// (P5 / P6, maxval <= 255). Pixels come back 8 bits per channel, top row first, with the channel count
// stb_image reports for the file (the loader's texel arithmetic depends on it).
// Returns 0 = ok, 1 = cannot open, 2 = format not supported or file corrupt.
// ---------------------------------------------------------------------------------------------
namespace {
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

int read_png(const std::vector<uint8_t>& file, int* w, int* h, int* ch, std::vector<uint8_t>& px) {
    if (file.size() < 8 + 25) return 2;
    size_t pos = 8;
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = 0;
    bool have_ihdr = false, have_trns = false;
    std::vector<uint8_t> idat, plte, trns;
    for (;;) {
        if (pos + 12 > file.size()) return 2;
        const uint32_t len = be32(&file[pos]);
        const char* type = (const char*)&file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) return 2;
        const uint8_t* d = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) return 2;
            W = be32(d); H = be32(d + 4); depth = d[8]; ctype = d[9];
            if (d[10] != 0 || d[11] != 0 || d[12] != 0) return 2;           // compression, filter, interlace (Adam7 not supported)
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(d, d + len);
        else if (!memcmp(type, "tRNS", 4)) { trns.assign(d, d + len); have_trns = true; }
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || W == 0 || H == 0 || W > (1u << 24) || H > (1u << 24) || idat.empty()) return 2;
    int nc;                                                               // channels in the file
    switch (ctype) { case 0: nc = 1; break; case 2: nc = 3; break; case 3: nc = 1; break; case 4: nc = 2; break; case 6: nc = 4; break; default: return 2; }
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) return 2;
    if (ctype == 3 && (depth == 16 || plte.size() < 3)) return 2;
    const size_t bpp_bits = (size_t)nc * depth;
    const size_t stride = ((size_t)W * bpp_bits + 7) / 8;
    const size_t fb = bpp_bits >= 8 ? bpp_bits / 8 : 1;                    // filter byte distance
    // the header is untrusted: deflate expands at most 1032:1, so a W x H the IDAT stream cannot fill is rejected before
    // anything of that size is allocated (and a texture beyond 2^28 samples is not a map_Kd anyone ships)
    if ((stride + 1) > ((size_t)1 << 40) / H || (stride + 1) * H > idat.size() * 1032 + 1024 || (size_t)W * H > ((size_t)1 << 28)) return 2;
    std::vector<uint8_t> raw((stride + 1) * H);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return 2;
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    std::vector<uint8_t> samples((size_t)W * H * nc);                      // one byte per sample (16-bit: high byte; palette: index)
    static const uint8_t scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    for (uint32_t y = 0; y < H; ++y) {
        const uint8_t* in = &raw[(stride + 1) * y];
        const int ft = in[0];
        if (ft > 4) return 2;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= fb ? cur[i - fb] : 0, b = prev[i], c = i >= fb ? prev[i - fb] : 0;
            int x = in[1 + i];
            switch (ft) { case 1: x += a; break; case 2: x += b; break; case 3: x += (a + b) >> 1; break; case 4: x += paeth(a, b, c); break; default: break; }
            cur[i] = (uint8_t)x;
        }
        uint8_t* out = &samples[(size_t)y * W * nc];
        const size_t n = (size_t)W * nc;
        if (depth == 8) memcpy(out, cur.data(), n);
        else if (depth == 16) for (size_t i = 0; i < n; ++i) out[i] = cur[2 * i];
        else for (size_t i = 0; i < n; ++i) {
            const int v = (cur[i * depth / 8] >> (8 - depth - (int)(i * depth % 8))) & ((1 << depth) - 1);
            out[i] = (uint8_t)(ctype == 0 ? v * scale[depth] : v);
        }
        prev.swap(cur);
    }
    const size_t np = (size_t)W * H;
    if (ctype == 3) {                                                      // palette -> RGB, or RGBA when a tRNS chunk is present
        const int oc = have_trns ? 4 : 3;
        px.resize(np * oc);
        for (size_t i = 0; i < np; ++i) {
            const size_t k = samples[i];
            for (int c = 0; c < 3; ++c) px[i * oc + c] = 3 * k + c < plte.size() ? plte[3 * k + c] : 0;
            if (have_trns) px[i * oc + 3] = k < trns.size() ? trns[k] : 255;
        }
        *ch = oc;
    } else if (have_trns && (ctype == 0 || ctype == 2) && trns.size() >= (size_t)2 * nc) {   // colour key -> alpha channel
        uint8_t key[3];
        for (int c = 0; c < nc; ++c) key[c] = depth == 16 ? trns[2 * c] : (uint8_t)(trns[2 * c + 1] * (depth < 8 ? scale[depth] : 1));
        const int oc = nc + 1;
        px.resize(np * oc);
        for (size_t i = 0; i < np; ++i) {
            bool same = true;
            for (int c = 0; c < nc; ++c) { px[i * oc + c] = samples[i * nc + c]; same = same && samples[i * nc + c] == key[c]; }
            px[i * oc + nc] = same ? 0 : 255;
        }
        *ch = oc;
    } else {
        px.swap(samples);
        *ch = nc;
    }
    *w = (int)W; *h = (int)H;
    return 0;
}

int read_pnm(const std::vector<uint8_t>& f, int* w, int* h, int* ch, std::vector<uint8_t>& px) {
    size_t pos = 2;
    auto number = [&](int* out) {
        for (;;) {
            while (pos < f.size() && (f[pos] == ' ' || f[pos] == '\t' || f[pos] == '\n' || f[pos] == '\r')) ++pos;
            if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') ++pos; continue; }
            break;
        }
        if (pos >= f.size() || f[pos] < '0' || f[pos] > '9') return false;
        long v = 0;
        while (pos < f.size() && f[pos] >= '0' && f[pos] <= '9' && v < (1 << 28)) v = v * 10 + (f[pos++] - '0');
        *out = (int)v;
        return true;
    };
    int W, H, maxv;
    if (!number(&W) || !number(&H) || !number(&maxv) || W <= 0 || H <= 0 || maxv <= 0 || maxv > 255) return 2;
    ++pos;                                                                 // the single whitespace after maxval
    const int nc = f[1] == '6' ? 3 : 1;
    const size_t n = (size_t)W * H * nc;
    if (pos + n > f.size()) return 2;
    px.assign(f.begin() + pos, f.begin() + pos + n);
    *w = W; *h = H; *ch = nc;
    return 0;
}
}  // namespace

int read_image(const char* path, int* w, int* h, int* ch, std::vector<uint8_t>& px) {
    FILE* f = fopen(path, "rb");
    if (!f) return 1;
    std::vector<uint8_t> file;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (file.size() >= 8 && !memcmp(file.data(), sig, 8)) return read_png(file, w, h, ch, px);
    if (file.size() >= 7 && file[0] == 'P' && (file[1] == '5' || file[1] == '6')) return read_pnm(file, w, h, ch, px);
    return 2;
}

}  // namespace crt
