// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h). Never linked into the product.
#include "orc_render.h"
#include <algorithm>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

static const uint32_t kKey1 = 0x43525431u;          // "CRT1"
static const uint32_t kCameraBounce = 0xFFFFFFFFu;
static const float kEps = 0.00001f;                 // Global.h:11
static const float kPi = 3.14159265358979323846f;   // static_cast<float>(M_PI)
static const float kTwoPi = 6.2831853071795864769f; // get_cuda_sphere_sample_inv_pdf(), Global.h:96-99

// Philox draw for (pixel, sample, bounce, dim)
static inline U4 draw(uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t dim, uint32_t seed) {
    return philox4x32_10(U4{pixel, sample, bounce, dim}, seed, kKey1);
}

static inline V3 mat3_mul(const float* M, V3 v) {
    return V3{dot(V3{M[0], M[1], M[2]}, v), dot(V3{M[3], M[4], M[5]}, v), dot(V3{M[6], M[7], M[8]}, v)};
}

Ray primary_ray(const Camera& cam, int width, int height, int i, int j, float u1, float u2) {
    // Render.cuh:338-339,344-347
    float ar = (float)width / (float)height;
    float x = (2.0f * ((float)i + u1) / (float)width - 1.0f) * cam.tan_half * ar;
    float y = (1.0f - 2.0f * ((float)j + u2) / (float)height) * cam.tan_half;
    V3 d = mat3_mul(cam.M, normalize(V3{-x, y, 1.0f}));
    return Ray{cam.eye, normalize(d), FLT_MAX};           // Ray.cuh:12-15 normalises again
}

// Global.h:35-50
static inline V3 to_world(V3 a, V3 N) {
    V3 C;
    if (fabsf(N.x) > fabsf(N.y)) {
        float inv = 1.0f / sqrtf(fmaf(N.z, N.z, N.x * N.x));
        C = V3{N.z * inv, 0.0f, -N.x * inv};
    } else {
        float inv = 1.0f / sqrtf(fmaf(N.z, N.z, N.y * N.y));
        C = V3{0.0f, N.z * inv, -N.y * inv};
    }
    V3 B = cross(C, N);
    return (a.x * B + a.y * C) + a.z * N;
}

// Global.h:57-66 : uniform hemisphere about N, pdf 1/2pi
static inline V3 sample_hemisphere(V3 N, float u1, float u2) {
    float z = fabsf(1.0f - 2.0f * u1);
    float r = sqrtf(1.0f - z * z);
    float sn, cs;
    sincos_2pi(u2, &sn, &cs);
    return to_world(V3{r * cs, r * sn, z}, N);
}

// Global.h:68-94 : box in (theta, phi) around `out`; trigonometry by angle addition so that no
// inverse functions are needed (cos/sin of theta0, phi0 come straight from the components).
static inline V3 sample_probe_lobe(V3 out, float dtheta, float dphi, float u1, float u2) {
    float eta1 = 2.0f * u1 - 1.0f, eta2 = 2.0f * u2 - 1.0f;
    float r = length(out);
    float ct0 = out.z / r;
    ct0 = fminf(1.0f, fmaxf(-1.0f, ct0));
    float st0 = sqrtf(fmaxf(0.0f, 1.0f - ct0 * ct0));
    float cp0, sp0;
    if (fabsf(out.x) < 1e-5f) {                           // Global.h:77-80
        cp0 = 0.0f;
        sp0 = out.y > 0.0f ? 1.0f : -1.0f;
    } else {
        float rho = sqrtf(fmaf(out.y, out.y, out.x * out.x));
        cp0 = out.x / rho;
        sp0 = out.y / rho;
    }
    float sa, ca, sb, cb;
    sincos_rad(eta1 * dtheta, &sa, &ca);
    sincos_rad(eta2 * dphi, &sb, &cb);
    float st = fmaf(st0, ca, ct0 * sa), ct = fmaf(ct0, ca, -(st0 * sa));
    float sp = fmaf(sp0, cb, cp0 * sb), cp = fmaf(cp0, cb, -(sp0 * sb));
    return V3{st * cp, st * sp, ct};
}

static inline int64_t quantize(float c) {
    if (!(fabsf(c) < 1073741824.0f)) return 0;            // drops NaN/inf/absurd values, both sides
    return (int64_t)llrint((double)c * kFixedScale);
}

static inline void add_contrib(int64_t* px, V3 c) {
    px[0] += quantize(c.x);
    px[1] += quantize(c.y);
    px[2] += quantize(c.z);
}

namespace {
struct Ctx {
    const Scene& s;
    const NewBVH& b;
    const Camera& cam;
    const RenderParams& p;
    const Wide8BVH* wide;          // when set, rays go through the 8-wide BVH (same hits, other visit counts)
};
static inline Hit trace(const Ctx& c, const Ray& r, int mode, TraceStats* st) {
    if (getenv("ORC_DEBUG_XCHECK") && c.wide) {                 // development aid: report rays on which the two BVHs disagree
        Hit a = wide8_intersect(c.s, *c.wide, r, mode, nullptr), b = new_intersect(c.s, c.b, r, mode, nullptr), e = brute_intersect(c.s, r, mode);
        bool bad = mode == 0 ? (a.face != b.face || a.t != b.t) : ((a.face >= 0) != (b.face >= 0));
        if (bad)
            fprintf(stderr, "XCHECK mode %d ray o %.9g %.9g %.9g d %.9g %.9g %.9g tmax %.9g | wide %d %.9g pair %d %.9g brute %d %.9g\n", mode, r.o.x, r.o.y,
                    r.o.z, r.d.x, r.d.y, r.d.z, r.tmax, a.face, a.t, b.face, b.t, e.face, e.t);
    }
    return c.wide ? wide8_intersect(c.s, *c.wide, r, mode, st) : new_intersect(c.s, c.b, r, mode, st);
}
}  // namespace

// One camera path of the compat estimator. Reference: cast_ray_v2, Render.cuh:199-328.
// The reference builds the vertex list forward and shades it backward; this is the same sum
// written forward:  L = sum_k T_k (x) D_k,  T_{k+1} = T_k (x) kd_k/pi * cos_k * 2pi / P_RR
// (Render.cuh:288-293), D_k = NEE of vertex k (:262-286) + the SPECULAR probe term (:294-314),
// D_0 = Ke when the first vertex is emissive (:249-255).
static void path_compat(const Ctx& c, uint32_t pixel, int i, int j, uint32_t sample, int64_t* px, RenderStats* st) {
    const Scene& s = c.s;
    const RenderParams& p = c.p;
    U4 r = draw(pixel, sample, kCameraBounce, 0, p.seed);
    Ray ray = primary_ray(c.cam, p.width, p.height, i, j, u01(r.x), u01(r.y));
    V3 T{1.0f, 1.0f, 1.0f};
    bool have_probe = false;
    Ray probe_ray{};
    V3 probe_w{};
    const float lsn_f = (float)p.light_sample_n;
    for (int bnc = 0; bnc < p.max_vertices; ++bnc) {
        Hit h = trace(c, ray, 0, &st->closest);
        st->extend_rays++;
        if (have_probe) {
            // Render.cuh:294-314 — evaluated only when the path continued to a real hit
            if (h.face >= 0) {
                Hit ph = trace(c, probe_ray, 0, &st->closest);
                st->probe_rays++;
                if (ph.face >= 0) {
                    const Material& pm = s.mats[s.tris[ph.face].mat];
                    if (pm.has_emit) add_contrib(px, cmul(probe_w, pm.ke));
                }
            }
            have_probe = false;
        }
        if (h.face < 0) break;                                            // :210
        const Tri& tri = s.tris[h.face];
        const Material& m = s.mats[tri.mat];
        if (m.has_emit) {                                                 // :210,249-255
            if (bnc == 0) add_contrib(px, m.ke);
            break;
        }
        V3 pos = ray.o + h.t * ray.d;                                     // DeviceTriangle.cuh:50
        V3 n = tri.normal;
        V3 f_r = m.kd / kPi;                                              // :259
        V3 Tf = cmul(T, f_r);
        // next-event estimation, :262-286
        for (int li = 0; li < (int)s.lights.size(); ++li) {
            const LightObj& L = s.lights[li];
            for (int sj = 0; sj < p.light_sample_n; ++sj) {
                U4 q = draw(pixel, sample, (uint32_t)bnc, 2u + (uint32_t)(li * p.light_sample_n + sj), p.seed);
                const Tri& lt = s.tris[L.tris[q.x % (uint32_t)L.tris.size()]];   // DeviceLights.cuh:35
                float alpha = u01(q.y);                                   // DeviceTriangle.cuh:69-72
                float beta = u01(q.z) * (1.0f - alpha);
                float gamma = (1.0f - alpha) - beta;
                V3 lp = (alpha * lt.v1 + beta * lt.v2) + gamma * lt.v3;
                V3 dist = lp - pos;
                // dir, t_to_light and the shadow ray's direction keep the reference's operation sequence (three divisions each):
                // its shadow test compares t_to_light - t with an absolute 1e-5 at t ~ 400 (Render.cuh:19-27,272), so how often the
                // light occludes its own samples depends on these roundings (with dist * (1 / d1) the converged cornell-box image
                // left the stated RMSE, r02_s16)
                V3 dir = normalize(dist);
                float d1 = length(dist);
                float d2 = d1 * d1;
                float cos1 = fmaxf(0.0f, dot(dir, n));
                float cos2 = fmaxf(0.0f, -dot(dir, lt.normal));
                const Material& lm = s.mats[lt.mat];
                V3 contrib = cmul(lm.ke, Tf) * (((cos1 * cos2) * L.area) / d2 / lsn_f);   // :274-283, the scalar factor first
                // a sample that cannot contribute needs no visibility test (the reference traces it anyway)
                if (contrib.x == 0.0f && contrib.y == 0.0f && contrib.z == 0.0f) continue;
                float t_to_light = dist.x / dir.x;                        // :272
                if (t_to_light == t_to_light) {                           // NaN => never "blocked" (:19-27)
                    Ray sh{pos, normalize(dir), t_to_light};
                    Hit bh = trace(c, sh, 1, &st->any);
                    st->shadow_rays++;
                    if (bh.face >= 0) continue;
                }
                add_contrib(px, contrib);
            }
        }
        if (bnc == p.max_vertices - 1) break;                             // bounce stack full, :210
        U4 q = draw(pixel, sample, (uint32_t)bnc, 0, p.seed);
        if (u01(q.x) > p.p_rr) break;                                     // :216-221
        V3 wdir = normalize_rcp(normalize_rcp(sample_hemisphere(n, u01(q.y), u01(q.z))));   // :225-227 + Ray ctor
        if (m.mode == SPECULAR) {                                         // :294-303
            V3 in = normalize_rcp(ray.d);
            V3 out = in - (2.0f * dot(in, n)) * n;
            U4 e = draw(pixel, sample, (uint32_t)bnc, 1, p.seed);
            V3 pd = normalize_rcp(normalize_rcp(sample_probe_lobe(out, m.probe_dtheta, m.probe_dphi, u01(e.x), u01(e.y))));
            probe_ray = Ray{pos, pd, FLT_MAX};
            float pc = fmaxf(0.0f, dot(pd, n));
            // :306-312 : (0.5*log10(Ns)+1) * Ke (x) kd * cos * 2pi/8, carried with the path throughput
            probe_w = cmul(T, m.kd) * m.probe_shin * pc * (kTwoPi / 8.0f);
            have_probe = true;
        }
        float cosn = fmaxf(0.0f, dot(wdir, n));
        T = Tf * (cosn * (kTwoPi / p.p_rr));                              // :288-293, the scalar factor first
        ray = Ray{pos, wdir, FLT_MAX};
    }
}

// ---------------------------------------------------------------------------------------------
// `mis` estimator (north star (c): BRDFs from Kd/Ks/Ns, next-event estimation against the emissive
// triangles with multiple importance sampling, Russian roulette at P_RR). Not the reference's
// estimator (the reference has no MIS); its oracle is this function.
//   surface   two-sided: n_s faces the incoming ray; emission is one-sided (front of the light triangle)
//   BRDF      modified Phong: kd/pi + ks (Ns+2)/(2 pi) max(0, r.wi)^Ns, r = mirror direction of wo
//   lights    a triangle is picked with probability ~ area * luminance(Ke) (CDF, binary search), a point
//             uniformly on it; density per area = luminance(Ke) / sum(area * luminance)
//   BSDF      lobe chosen with probability lum(kd) / (lum(kd) + lum(ks)): cosine hemisphere or Phong lobe
//   MIS       power heuristic between light_sample_n light samples and the one BSDF sample
//   RR        continue with probability P_RR after every vertex (as the reference does, Render.cuh:216-221)
// ---------------------------------------------------------------------------------------------
static inline int pick_light(const std::vector<float>& cdf, float u) {
    int lo = 0, hi = (int)cdf.size() - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] >= u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

static inline void phong_eval(const Material& m, float pd, bool has_spec, float cos_s, float ca, V3* f, float* pdf) {
    V3 fd = m.kd / kPi;
    float pdf_d = cos_s / kPi;
    if (has_spec) {
        float pw = det_pow(ca, m.ns);
        *f = fd + m.ks * ((m.ns + 2.0f) / kTwoPi * pw);
        float pdf_s = (m.ns + 1.0f) / kTwoPi * pw;
        *pdf = fmaf(pd, pdf_d, (1.0f - pd) * pdf_s);
    } else {
        *f = fd;
        *pdf = pdf_d;
    }
}

static void path_mis(const Ctx& c, uint32_t pixel, int i, int j, uint32_t sample, int64_t* px, RenderStats* st) {
    const Scene& s = c.s;
    const RenderParams& p = c.p;
    U4 r = draw(pixel, sample, kCameraBounce, 0, p.seed);
    Ray ray = primary_ray(c.cam, p.width, p.height, i, j, u01(r.x), u01(r.y));
    V3 T{1.0f, 1.0f, 1.0f};
    float pdf_b = 0.0f;
    const float lsn_f = (float)p.light_sample_n;
    const int n_lt = (int)s.light_flat.size();
    // validation variants of this oracle only (never on the GPU): the same integral with light samples
    // alone or with BSDF samples alone; all three must converge to the same image (tests/test_oracle.py)
    const int variant = p.estimator;
    for (int bnc = 0; bnc < p.max_vertices; ++bnc) {
        Hit h = trace(c, ray, 0, &st->closest);
        st->extend_rays++;
        if (h.face < 0) break;
        const Tri& tri = s.tris[h.face];
        const Material& m = s.mats[tri.mat];
        V3 n = tri.normal;
        float dn = dot(n, ray.d);
        if (m.has_emit) {
            if (dn < 0.0f) {
                if (bnc == 0) add_contrib(px, m.ke);
                else {
                    float pl = m.pdf_area * (h.t * h.t) / (-dn);
                    float pls = lsn_f * pl;
                    float w = (pdf_b * pdf_b) / fmaf(pdf_b, pdf_b, pls * pls);
                    if (variant == ESTIMATOR_MIS_LIGHT_ONLY) w = 0.0f;
                    if (variant == ESTIMATOR_MIS_BSDF_ONLY) w = 1.0f;
                    add_contrib(px, cmul(T, m.ke) * w);
                }
            }
            break;
        }
        V3 ns = dn > 0.0f ? neg(n) : n;
        V3 wo = neg(ray.d);
        V3 pos = ray.o + h.t * ray.d;
        float off = 1.0e-4f * (1.0f + fmaxf(fmaxf(fabsf(pos.x), fabsf(pos.y)), fabsf(pos.z)));
        V3 org = pos + off * ns;
        float cos_o = dot(ns, wo);
        V3 refl = normalize_rcp((2.0f * cos_o) * ns - wo);
        float lkd = lumf(m.kd), lks = lumf(m.ks);
        float lsum = lkd + lks;
        if (!(lsum > 0.0f)) break;
        float pd = lkd / lsum;
        bool has_spec = lks > 0.0f;
        for (int sj = 0; sj < p.light_sample_n && n_lt > 0; ++sj) {
            U4 q = draw(pixel, sample, (uint32_t)bnc, 2u + (uint32_t)sj, p.seed);
            int k = pick_light(s.light_cdf, u01(q.x));
            const Tri& lt = s.tris[s.light_flat[k]];
            const Material& lm = s.mats[lt.mat];
            float su = sqrtf(u01(q.y));
            float b0 = 1.0f - su, b1 = u01(q.z) * su;
            float b2 = (1.0f - b0) - b1;
            V3 lp = (b0 * lt.v1 + b1 * lt.v2) + b2 * lt.v3;
            V3 dist = lp - org;
            float d2 = dot(dist, dist);
            float d1 = sqrtf(d2);
            V3 wi = dist * (1.0f / d1);
            float cos_s = dot(ns, wi);
            float cos_l = -dot(lt.normal, wi);
            if (!(cos_s > 0.0f && cos_l > 0.0f)) continue;
            float ca = fmaxf(0.0f, dot(refl, wi));
            V3 f;
            float pb;
            phong_eval(m, pd, has_spec, cos_s, ca, &f, &pb);
            float pl = lm.pdf_area * d2 / cos_l;
            float pls = lsn_f * pl;
            float w = (pls * pls) / fmaf(pls, pls, pb * pb);
            if (variant == ESTIMATOR_MIS_LIGHT_ONLY) w = 1.0f;
            if (variant == ESTIMATOR_MIS_BSDF_ONLY) w = 0.0f;
            V3 contrib = cmul(cmul(T, f), lm.ke) * (cos_s * w / pls);
            if (contrib.x == 0.0f && contrib.y == 0.0f && contrib.z == 0.0f) continue;
            Ray sh{org, wi, d1 * 0.999f};
            Hit bh = trace(c, sh, 1, &st->any);
            st->shadow_rays++;
            if (bh.face >= 0) continue;
            add_contrib(px, contrib);
        }
        if (bnc == p.max_vertices - 1) break;
        U4 q = draw(pixel, sample, (uint32_t)bnc, 0, p.seed);
        if (u01(q.x) > p.p_rr) break;
        float u1 = u01(q.z), u2 = u01(q.w);
        float sn, cs;
        sincos_2pi(u2, &sn, &cs);
        V3 wi;
        if (u01(q.y) <= pd) {
            float rr = sqrtf(u1);
            float z = sqrtf(1.0f - u1);
            wi = to_world(V3{rr * cs, rr * sn, z}, ns);
        } else {
            float ca0 = det_pow(u1, 1.0f / (m.ns + 1.0f));
            float sa0 = sqrtf(fmaxf(0.0f, 1.0f - ca0 * ca0));
            wi = to_world(V3{sa0 * cs, sa0 * sn, ca0}, refl);
        }
        wi = normalize_rcp(wi);
        float cos_s = dot(ns, wi);
        if (!(cos_s > 0.0f)) break;
        float ca = fmaxf(0.0f, dot(refl, wi));
        V3 f;
        float pb;
        phong_eval(m, pd, has_spec, cos_s, ca, &f, &pb);
        if (!(pb > 0.0f)) break;
        T = cmul(T, f) * (cos_s / pb / p.p_rr);
        pdf_b = pb;
        ray = Ray{org, wi, FLT_MAX};
    }
}

void render(const Scene& s, const NewBVH& b, const Camera& cam, const RenderParams& p, int64_t* accum,
            RenderStats* stats, int n_threads, const Wide8BVH* wide) {
    Ctx c{s, b, cam, p, wide};
    const int npix = p.width * p.height;
    if (n_threads < 1) n_threads = 1;
    std::vector<RenderStats> tls(n_threads);
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
    for (int pix = 0; pix < npix; ++pix) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        int i = pix % p.width, j = pix / p.width;
        for (uint32_t sm = p.s_begin; sm < p.s_end; ++sm) {
            if (p.estimator != ESTIMATOR_COMPAT) path_mis(c, (uint32_t)pix, i, j, sm, accum + 3 * (size_t)pix, &tls[tid]);
            else path_compat(c, (uint32_t)pix, i, j, sm, accum + 3 * (size_t)pix, &tls[tid]);
            tls[tid].samples++;
        }
    }
    if (stats) {
        for (auto& t : tls) {
            stats->samples += t.samples; stats->extend_rays += t.extend_rays;
            stats->shadow_rays += t.shadow_rays; stats->probe_rays += t.probe_rays;
            auto acc = [](TraceStats& a, const TraceStats& x) {
                a.inner += x.inner; a.boxes += x.boxes; a.tris += x.tris; a.rays += x.rays;
                a.max_stack = std::max(a.max_stack, x.max_stack);
            };
            acc(stats->closest, t.closest);
            acc(stats->any, t.any);
        }
    }
}

void resolve(const int64_t* accum, int n_pixels, uint32_t spp, float* linear_rgb, uint8_t* rgb8) {
    for (int k = 0; k < 3 * n_pixels; ++k) {
        float v = (float)((double)accum[k] / kFixedScale / (double)spp);
        if (linear_rgb) linear_rgb[k] = v;
        if (rgb8) {
            float cl = fmaxf(0.0f, fminf(1.0f, v));                       // Global.h:121-124
            rgb8[k] = (uint8_t)(255.0f * powf(cl, 0.6f));                 // Render.cuh:350
        }
    }
}

}  // namespace orc
