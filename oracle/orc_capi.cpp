// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h). Never linked into the product.
// C entry points for tests/ (ctypes via oracle/orc.py).
#include <cstdio>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "orc_render.h"

using namespace orc;

struct orc_scene {
    Scene s;
    RefBVH ref;
    NewBVH nb;
    Wide8BVH wb;
    bool has_ref = false, has_new = false, has_wide = false;
};

extern "C" {

orc_scene* orc_scene_create() { return new orc_scene(); }
void orc_scene_destroy(orc_scene* h) { delete h; }
const char* orc_scene_error(orc_scene* h) { return h->s.error.c_str(); }

int orc_scene_add_obj(orc_scene* h, const char* obj_path, const char* mtl_dir) {
    return load_obj(h->s, obj_path, mtl_dir) ? 0 : -1;
}

// Synthetic scenes: verts T*9, mat_id T, obj_id T (objects numbered 0..n_obj-1 in first-use order),
// mats n_mat*7 = kd(3), ke(3), ns.
// mats: n_mats rows of mat_cols floats: 7 = (kd, ke, ns), 10 = (kd, ks, ke, ns)
int orc_scene_add_arrays(orc_scene* h, const float* verts, const int* mat_id, const int* obj_id, int n_tris,
                         const float* mats, int n_mats, int mat_cols) {
    Scene& s = h->s;
    int mat0 = (int)s.mats.size(), obj0 = s.n_objects, max_obj = -1;
    for (int m = 0; m < n_mats; ++m) {
        Material mm;
        const float* r = mats + (size_t)mat_cols * m;
        mm.kd = V3{r[0], r[1], r[2]};
        if (mat_cols == 10) { mm.ks = V3{r[3], r[4], r[5]}; mm.ke = V3{r[6], r[7], r[8]}; mm.ns = r[9]; }
        else { mm.ke = V3{r[3], r[4], r[5]}; mm.ns = r[6]; }
        finish_material(mm);
        s.mats.push_back(mm);
    }
    for (int t = 0; t < n_tris; ++t) {
        Tri tr;
        const float* v = verts + 9 * (size_t)t;
        tr.v1 = V3{v[0] + 0.0f, v[1] + 0.0f, v[2] + 0.0f};
        tr.v2 = V3{v[3] + 0.0f, v[4] + 0.0f, v[5] + 0.0f};
        tr.v3 = V3{v[6] + 0.0f, v[7] + 0.0f, v[8] + 0.0f};
        tr.mat = mat0 + mat_id[t];
        tr.obj = obj0 + obj_id[t];
        if (obj_id[t] > max_obj) max_obj = obj_id[t];
        finish_triangle(tr);
        s.tris.push_back(tr);
    }
    s.n_objects = obj0 + max_obj + 1;
    finish_objects(s);
    return 0;
}

int orc_scene_n_tris(orc_scene* h) { return (int)h->s.tris.size(); }
int orc_scene_n_mats(orc_scene* h) { return (int)h->s.mats.size(); }
int orc_scene_n_lights(orc_scene* h) { return (int)h->s.lights.size(); }
int orc_scene_n_objects(orc_scene* h) { return h->s.n_objects; }

// Per-triangle dump: verts T*9, normal T*3, area T, area_of_obj T, mat T, obj T (any may be NULL)
void orc_scene_get_tris(orc_scene* h, float* verts, float* normal, float* area, float* area_of_obj, int* mat, int* obj) {
    const Scene& s = h->s;
    for (size_t t = 0; t < s.tris.size(); ++t) {
        const Tri& tr = s.tris[t];
        if (verts) { float* v = verts + 9 * t; v[0] = tr.v1.x; v[1] = tr.v1.y; v[2] = tr.v1.z; v[3] = tr.v2.x; v[4] = tr.v2.y; v[5] = tr.v2.z; v[6] = tr.v3.x; v[7] = tr.v3.y; v[8] = tr.v3.z; }
        if (normal) { normal[3 * t] = tr.normal.x; normal[3 * t + 1] = tr.normal.y; normal[3 * t + 2] = tr.normal.z; }
        if (area) area[t] = tr.area;
        if (area_of_obj) area_of_obj[t] = tr.area_of_obj;
        if (mat) mat[t] = tr.mat;
        if (obj) obj[t] = tr.obj;
    }
}
// mats n*9: kd(3) ke(3) ns has_emit mode
void orc_scene_get_mats(orc_scene* h, float* out) {
    for (size_t m = 0; m < h->s.mats.size(); ++m) {
        const Material& mm = h->s.mats[m];
        float* o = out + 9 * m;
        o[0] = mm.kd.x; o[1] = mm.kd.y; o[2] = mm.kd.z; o[3] = mm.ke.x; o[4] = mm.ke.y; o[5] = mm.ke.z;
        o[6] = mm.ns; o[7] = (float)mm.has_emit; o[8] = (float)mm.mode;
    }
}
int orc_scene_light_size(orc_scene* h, int li) { return (int)h->s.lights[li].tris.size(); }
float orc_scene_light_area(orc_scene* h, int li) { return h->s.lights[li].area; }
void orc_scene_light_tris(orc_scene* h, int li, int* faces) {
    memcpy(faces, h->s.lights[li].tris.data(), h->s.lights[li].tris.size() * sizeof(int));
}

// ---- reference BVH
int orc_refbvh_build(orc_scene* h, unsigned thresh_n) {
    build_ref_bvh(h->s, thresh_n, h->ref);
    h->has_ref = true;
    return (int)h->ref.nodes.size();
}
int orc_refbvh_root(orc_scene* h) { return h->ref.root; }
void orc_refbvh_get(orc_scene* h, void* nodes40, int* order) {
    if (nodes40) memcpy(nodes40, h->ref.nodes.data(), h->ref.nodes.size() * sizeof(RefNode));
    if (order) memcpy(order, h->ref.order.data(), h->ref.order.size() * sizeof(int));
}

// ---- new BVH
int orc_newbvh_build(orc_scene* h, unsigned thresh_n, int builder) {
    build_new_bvh(h->s, thresh_n, builder, h->nb);
    h->has_new = true;
    return (int)h->nb.nodes.size();
}
void orc_newbvh_get(orc_scene* h, void* nodes64, int* order, uint8_t* last, float* bounds6) {
    if (nodes64) memcpy(nodes64, h->nb.nodes.data(), h->nb.nodes.size() * sizeof(PairNode));
    if (order) memcpy(order, h->nb.order.data(), h->nb.order.size() * sizeof(int));
    if (last) memcpy(last, h->nb.last.data(), h->nb.last.size());
    if (bounds6) { bounds6[0] = h->nb.lo.x; bounds6[1] = h->nb.lo.y; bounds6[2] = h->nb.lo.z; bounds6[3] = h->nb.hi.x; bounds6[4] = h->nb.hi.y; bounds6[5] = h->nb.hi.z; }
}

// ---- 8-wide BVH
int orc_wide8_build2(orc_scene* h, unsigned thresh_n, int builder) {
    build_wide8_bvh(h->s, thresh_n, h->wb, builder);
    h->has_wide = true;
    return (int)h->wb.nodes.size();
}
int orc_wide8_build(orc_scene* h, unsigned thresh_n) {
    build_wide8_bvh(h->s, thresh_n, h->wb);
    h->has_wide = true;
    return (int)h->wb.nodes.size();
}
void orc_wide8_get(orc_scene* h, void* nodes80, int* order, uint8_t* last, float* bounds6) {
    if (nodes80) memcpy(nodes80, h->wb.nodes.data(), h->wb.nodes.size() * sizeof(Wide8Node));
    if (order) memcpy(order, h->wb.order.data(), h->wb.order.size() * sizeof(int));
    if (last) memcpy(last, h->wb.last.data(), h->wb.last.size());
    if (bounds6) { bounds6[0] = h->wb.lo.x; bounds6[1] = h->wb.lo.y; bounds6[2] = h->wb.lo.z; bounds6[3] = h->wb.hi.x; bounds6[4] = h->wb.hi.y; bounds6[5] = h->wb.hi.z; }
}

// ---- tracing. rays n*8 floats: o(3) tmax d(3) pad ; out t[n], face[n]
// which: 0 = new BVH, 1 = reference BVH + reference rule (canonical ties), 2 = reference rule literal
//        ties, 3 = brute force, 4 = 8-wide BVH.  mode: 0 closest, 1 any-hit (which 0, 3 and 4 only).
// stats5 (optional): inner, boxes, tris, max_stack, rays
int orc_trace(orc_scene* h, int which, int mode, const float* rays, int64_t n, float* t_out, int* face_out,
              uint64_t* stats5, int n_threads) {
    if (which == 0 && !h->has_new) return -1;
    if ((which == 1 || which == 2) && !h->has_ref) return -1;
    if (which == 4 && !h->has_wide) return -1;
    if (n_threads < 1) n_threads = 1;
    std::vector<TraceStats> tls(n_threads);
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
    for (int64_t k = 0; k < n; ++k) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        const float* r = rays + 8 * k;
        Ray ray{V3{r[0], r[1], r[2]}, V3{r[4], r[5], r[6]}, r[3]};
        Hit hit;
        if (which == 0) hit = new_intersect(h->s, h->nb, ray, mode, &tls[tid]);
        else if (which == 1) hit = ref_intersect(h->s, h->ref, ray.o, ray.d, true, &tls[tid]);
        else if (which == 2) hit = ref_intersect(h->s, h->ref, ray.o, ray.d, false, &tls[tid]);
        else if (which == 4) hit = wide8_intersect(h->s, h->wb, ray, mode, &tls[tid]);
        else hit = brute_intersect(h->s, ray, mode);
        if (t_out) t_out[k] = hit.t;
        if (face_out) face_out[k] = hit.face;
    }
    if (stats5) {
        memset(stats5, 0, 5 * sizeof(uint64_t));
        for (auto& t : tls) {
            stats5[0] += t.inner; stats5[1] += t.boxes; stats5[2] += t.tris;
            if (t.max_stack > stats5[3]) stats5[3] = t.max_stack;
            stats5[4] += t.rays;
        }
    }
    return 0;
}

float orc_det_log2(float x) { return det_log2(x); }
float orc_det_exp2(float x) { return det_exp2(x); }
float orc_det_pow(float x, float y) { return det_pow(x, y); }

// Canonical triangle test of face faces[k] against ray k (checks a reported any-hit blocker):
// t_out = t of the Moeller-Trumbore statement, inside_out = 1 when the strict-inside rule accepts it.
int orc_tri_test(orc_scene* h, const float* rays, const int* faces, int64_t n, float* t_out, unsigned char* inside_out) {
    for (int64_t k = 0; k < n; ++k) {
        const float* r = rays + 8 * k;
        t_out[k] = FLT_MAX;
        inside_out[k] = 0;
        if (faces[k] < 0 || faces[k] >= (int)h->s.tris.size()) continue;
        float t;
        bool in = tri_test(h->s.tris[faces[k]], V3{r[0], r[1], r[2]}, V3{r[4], r[5], r[6]}, &t);
        t_out[k] = t;
        inside_out[k] = in ? 1 : 0;
    }
    return 0;
}

// Primary rays for an image: jitter == NULL -> pixel centres (u = 0.5); else jitter[2*pixel..]
void orc_primary_rays(const float eye[3], const float M[9], float fovy_rad, int width, int height, float* rays) {
    Camera cam;
    cam.eye = V3{eye[0], eye[1], eye[2]};
    memcpy(cam.M, M, sizeof(cam.M));
    cam.tan_half = tanf(fovy_rad / 2);
    for (int j = 0; j < height; ++j)
        for (int i = 0; i < width; ++i) {
            Ray r = primary_ray(cam, width, height, i, j, 0.5f, 0.5f);
            float* o = rays + 8 * ((size_t)j * width + i);
            o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = FLT_MAX; o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = 0.0f;
        }
}

void orc_inverse_view_matrix(const float eye[3], const float lookat[3], const float up[3], float out9[9]) {
    inverse_view_matrix(eye, lookat, up, out9);
}

// stats12: samples, extend, shadow, probe, closest{inner,tris,rays,max_stack}, any{inner,tris,rays,max_stack}
int orc_render(orc_scene* h, const float eye[3], const float M[9], float fovy_rad, int width, int height,
               uint32_t s_begin, uint32_t s_end, float p_rr, int light_sample_n, uint32_t seed, int estimator, int use_wide,
               int64_t* accum, uint64_t* stats12, int n_threads) {
    if (!h->has_new) return -1;
    if (estimator < ESTIMATOR_COMPAT || estimator > ESTIMATOR_MIS_BSDF_ONLY) return -2;
    Camera cam;
    cam.eye = V3{eye[0], eye[1], eye[2]};
    memcpy(cam.M, M, sizeof(cam.M));
    cam.tan_half = tanf(fovy_rad / 2);                    // Render.cuh:338
    RenderParams p;
    p.width = width; p.height = height; p.s_begin = s_begin; p.s_end = s_end;
    p.p_rr = p_rr; p.light_sample_n = light_sample_n; p.seed = seed; p.estimator = estimator;
    RenderStats st;
    if (use_wide && !h->has_wide) return -1;
    render(h->s, h->nb, cam, p, accum, &st, n_threads, use_wide ? &h->wb : nullptr);
    if (stats12) {
        uint64_t v[12] = {st.samples, st.extend_rays, st.shadow_rays, st.probe_rays,
                          st.closest.inner, st.closest.tris, st.closest.rays, st.closest.max_stack,
                          st.any.inner, st.any.tris, st.any.rays, st.any.max_stack};
        memcpy(stats12, v, sizeof(v));
    }
    return 0;
}

void orc_resolve(const int64_t* accum, int n_pixels, uint32_t spp, float* linear_rgb, uint8_t* rgb8) {
    resolve(accum, n_pixels, spp, linear_rgb, rgb8);
}

// unit-test hooks
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    U4 r = philox4x32_10(U4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
float orc_u01(uint32_t x) { return u01(x); }
void orc_sincos_2pi(float u, float* s, float* c) { sincos_2pi(u, s, c); }
void orc_sincos_rad(float x, float* s, float* c) { sincos_rad(x, s, c); }
int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
