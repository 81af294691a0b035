// crt_config_png.cpp — the two file formats at the edges of the path:
//   config.json in (schema of src/main.cu:67-90), RGB8 PNG out (Render.cuh:489-493).
#include "crt_host.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
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
    bool parse(JValue& v) {
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

int write_png(const char* path, const uint8_t* rgb8, uint32_t width, uint32_t height) {
    if (!path || !rgb8 || width == 0 || height == 0) { set_error("write_png: invalid argument"); return CRT_ERR_INVALID; }
    const size_t row = (size_t)width * 3;
    std::vector<uint8_t> raw((row + 1) * height);
    for (uint32_t y = 0; y < height; ++y) {
        raw[(row + 1) * y] = 0;
        memcpy(&raw[(row + 1) * y + 1], rgb8 + row * y, row);
    }
    uLongf bound = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(bound);
    if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), 6) != Z_OK) { set_error("write_png: deflate failed"); return CRT_ERR_NOMEM; }
    std::vector<uint8_t> out;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    out.insert(out.end(), sig, sig + 8);
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width); put_be32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    put_chunk(out, "IHDR", ihdr.data(), ihdr.size());
    put_chunk(out, "IDAT", z.data(), bound);
    put_chunk(out, "IEND", nullptr, 0);
    FILE* f = fopen(path, "wb");
    if (!f) { set_error(std::string("write_png: cannot open ") + path); return CRT_ERR_IO; }
    size_t w = fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    if (w != out.size()) { set_error("write_png: short write"); return CRT_ERR_IO; }
    return CRT_OK;
}

}  // namespace crt
