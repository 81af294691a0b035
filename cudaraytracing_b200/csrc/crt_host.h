// crt_host.h — host-side scene of the product: what the reference keeps in Scene/Object/
// Triangle/Material (include/Scene.h, Object.h, Triangle.h, Material.h), held as flat arrays.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/crt.h"

namespace crt {

void set_error(const std::string& msg);
const char* get_error();

struct HostMaterial {
    float kd[3] = {0, 0, 0}, ks[3] = {0, 0, 0}, ke[3] = {0, 0, 0};
    float ns = 1.0f;
    int has_emit = 0;          // Material.h:36-39
    int mode = 0;              // 0 DIFFUSE, 1 SPECULAR (Loader.h:107)
    float probe_dtheta = 0, probe_dphi = 0, probe_shin = 1;   // Render.cuh:296-300,306-307
    float pdf_area = 0;        // mis estimator: light-sampling density per unit area on triangles of this material
    std::string name;
    std::string map_kd;        // texture path (OBJLoader.h:184-193); per-triangle Kd is derived at load time (Loader.h:78-105)
};

struct HostLight {             // one emissive usemtl group (Scene.h:38-42, DeviceLights.cuh:6-54)
    std::vector<int32_t> faces;
    float area = 0;            // Object.h:15-23
};

struct HostScene {
    // per triangle, scene order (face id = index): Scene::triangles, Scene.h:32-48
    std::vector<float> verts;        // 9 per triangle: v1 v2 v3
    std::vector<float> normal;       // 3 per triangle, Triangle.h:27
    std::vector<float> area;         // Triangle.h:39
    std::vector<float> area_of_obj;  // Object.h:15-23
    std::vector<int32_t> mat, obj;
    std::vector<HostMaterial> mats;
    std::vector<HostLight> lights;
    // mis estimator: CDF over all light triangles (light order, then face order), P ~ area * luminance(Ke)
    std::vector<float> light_cdf;
    int n_objects = 0;
    size_t n_tris() const { return mat.size(); }
};

// Derived per-material constants (has_emit, mode, probe lobe), Material.h:33-40 / Loader.h:107.
void finish_material(HostMaterial& m);
// Append one triangle (derives normal and area like Triangle.h:23-41). Returns false on NaN/inf.
bool push_triangle(HostScene& s, const float v[9], int mat, int obj);
// n triangles at once (verts: 9 floats each), derived data filled in on all host threads; material / object ids are
// offset by mat0 / obj0. false (scene unchanged): a coordinate is not finite.
bool append_triangles(HostScene& s, const float* verts, const uint32_t* mat_id, const uint32_t* obj_id, size_t n, int mat0, int obj0);
// Object areas and the light list (Object.h:12-26, Scene.h:38-48). Call after adding triangles.
void finish_objects(HostScene& s);
// OBJLoader::parse + Loader::load_object semantics over a memory-mapped file.
int load_obj(HostScene& s, const char* obj_path, const char* mtl_dir);

// config.json (main.cu:67-90)
int load_config(const char* path, crt_config* out);
// Camera.h:9-36
void inverse_view_matrix(const float eye[3], const float lookat[3], const float up[3], float out9[9]);
// PNG, RGB8, top row first (replaces stbi_write_png, Render.cuh:492)
int write_png(const char* path, const uint8_t* rgb8, uint32_t width, uint32_t height);
// map_Kd texture files (replaces stbi_load, Loader.h:58): PNG / binary PNM -> 8-bit pixels, top row first, channel count as
// stb_image reports it. 0 = ok, 1 = cannot open, 2 = unsupported format or corrupt file.
int read_image(const char* path, int* width, int* height, int* channels, std::vector<uint8_t>& pixels);

}  // namespace crt
