// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h). Never linked into the product.
//
// orc_render.h — CPU statement of the estimators the CUDA wavefront implements.
//   ESTIMATOR_COMPAT : the reference's cast_ray_v2 (Render.cuh:199-328) in expectation, rewritten
//                      as a forward throughput recursion with counter-based Philox draws.
//   ESTIMATOR_MIS    : the north-star estimator (Phong lobe from Kd/Ns + NEE, balance heuristic).
// Output is the fixed-point accumulation buffer (int64, radiance * 2^32 summed over samples), the
// same object the GPU produces, so parity is exact integer equality.
#pragma once
#include "orc_bvh.h"

namespace orc {

// 2 and 3 exist in the oracle only: the mis integral estimated with light samples alone / BSDF samples alone
enum Estimator { ESTIMATOR_COMPAT = 0, ESTIMATOR_MIS = 1, ESTIMATOR_MIS_LIGHT_ONLY = 2, ESTIMATOR_MIS_BSDF_ONLY = 3 };

struct Camera {
    V3 eye;
    float M[9];        // inverse view matrix, row-major, columns [r u f] (Camera.h:25-33)
    float tan_half;    // tanf(fovY/2), evaluated by the caller on the host (Render.cuh:338)
};

struct RenderParams {
    int width, height;
    uint32_t s_begin, s_end;     // sample index range rendered by this call (multi-GPU shards)
    float p_rr;
    int light_sample_n;
    uint32_t seed;
    int estimator;
    int max_vertices = 64;       // BOUNCE_STACK_SIZE (Global.h:18)
};

struct RenderStats {
    uint64_t samples = 0, extend_rays = 0, shadow_rays = 0, probe_rays = 0;
    TraceStats closest, any;     // node / triangle visit counts of the two ray kinds
};

static const double kFixedScale = 4294967296.0;   // 2^32

// Adds the samples [s_begin, s_end) of every pixel into accum[W*H*3].
void render(const Scene& s, const NewBVH& b, const Camera& cam, const RenderParams& p, int64_t* accum,
            RenderStats* stats, int n_threads, const Wide8BVH* wide = nullptr);

// Single primary ray of pixel (i,j) with jitter (u1,u2): Render.cuh:344-347 + Ray.cuh:12-15.
Ray primary_ray(const Camera& cam, int width, int height, int i, int j, float u1, float u2);

// E11 (Render.cuh:350): mean over spp, clamp, pow 0.6, *255, truncate.
void resolve(const int64_t* accum, int n_pixels, uint32_t spp, float* linear_rgb, uint8_t* rgb8);

}  // namespace orc
