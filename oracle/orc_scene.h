// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h). Never linked into the product.
//
// orc_scene.h — host scene as the reference builds it before any GPU work:
//   OBJLoader::parse (include/OBJLoader.h:61-203) -> Loader::load_object (include/Loader.h:40-124)
//   -> Triangle (include/Triangle.h:23-41) -> Object (include/Object.h:12-26)
//   -> Scene::add_normal_obj / add_light_obj (include/Scene.h:32-48), in the order of
//   render_view() (src/main.cu:122-145).
#pragma once
#include <string>
#include <vector>
#include "orc_math.h"

namespace orc {

enum Mode { DIFFUSE = 0, SPECULAR = 1 };   // include/Material.h:7-10

struct Material {            // include/Material.h:11-40, as filled by Loader.h:45-47,107
    V3 kd{0, 0, 0};
    V3 ks{0, 0, 0};          // kept for the `mis` estimator only; `compat` ignores it like Loader.h:45-47
    V3 ke{0, 0, 0};
    float pdf_area = 0.0f;   // mis: density (per unit area) of light sampling on a triangle of this material
    float ns = 1.0f;
    int has_emit = 0;        // !(ke.x<eps && ke.y<eps && ke.z<eps), Material.h:36-39
    int mode = DIFFUSE;      // ns > 1 ? SPECULAR : DIFFUSE, Loader.h:107
    // constants of the SPECULAR probe (Render.cuh:296-300,306-308), evaluated once per material
    float probe_dtheta = 0.0f, probe_dphi = 0.0f, probe_shin = 1.0f;
    std::string name;
};

struct Tri {                 // include/Triangle.h:9-41 (+ face id, which the reference lacks)
    V3 v1, v2, v3;
    V3 center;               // (v1+v2+v3)/3                         Triangle.h:26
    V3 normal;               // normalized((v2-v1)x(v3-v1))          Triangle.h:27
    V3 lo, hi;               // per-axis min/max of the vertices     Triangle.h:30-37
    float area;              // |(v2-v1)x(v3-v1)| * 0.5              Triangle.h:39
    float area_of_obj;       // Object.h:15-23
    int mat;                 // index into Scene::mats
    int obj;                 // index of the Object (usemtl group) it came from
};

struct LightObj {            // one emissive Object; DeviceLight (include/DeviceLights.cuh:6-54)
    std::vector<int> tris;   // face ids, in object order
    float area;              // Object area = inv_pdf of DeviceTriangle::sample (DeviceTriangle.cuh:73)
};

struct Scene {
    std::vector<Tri> tris;            // Scene::triangles order (face id = index)
    std::vector<Material> mats;
    std::vector<LightObj> lights;     // Scene::light_objs order
    // mis estimator: all light triangles in (light object, object order), and the CDF that picks one with
    // probability proportional to area * luminance(Ke)
    std::vector<int> light_flat;
    std::vector<float> light_cdf;
    int n_objects = 0;
    std::string error;
};

// Fill derived triangle fields from v1,v2,v3 (Triangle.h:23-41), host arithmetic (no FMA).
void finish_triangle(Tri& t);
void finish_material(Material& m);
// Object.h:12-26 for every usemtl group, then light list (Scene.h:38-48).
void finish_objects(Scene& s);

// Reference-semantics OBJ+MTL ingest; appends to `s`. Returns false and sets s.error on I/O failure.
bool load_obj(Scene& s, const std::string& obj_path, const std::string& mtl_dir);

// Camera.h:9-36 : columns [r u f], row-major 3x3 out.
void inverse_view_matrix(const float eye[3], const float lookat[3], const float up[3], float out9[9]);

}  // namespace orc
