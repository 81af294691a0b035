// exp_sort.cuh — included by crt_render.cu only with -DCRT_EXP_SORT (tools/gpu_round4.sh). Result: profiles/r01_ray_sort_experiment.md
// EXPERIMENT (not part of the product build): does ray ordering pay? Sort the shadow queue and the next
// path queue by (origin cell, direction octant) with CUB and gather; stage timings then show the
// traversal kernels on coherent input.
}  // namespace crt
#include <cub/device/device_radix_sort.cuh>
namespace crt {
struct ExpSort {
    uint32_t *k_in = nullptr, *k_out = nullptr, *v_in = nullptr, *v_out = nullptr;
    float4 *t0 = nullptr, *t1 = nullptr, *t2 = nullptr;
    void* temp = nullptr; size_t temp_bytes = 0; uint32_t cap = 0;
};
static ExpSort g_exp;
__global__ void k_exp_keys(const Counters* c, int which, const float4* __restrict__ o, const float4* __restrict__ d, uint32_t cap,
                           float lox, float loy, float loz, float sx, float sy, float sz, uint32_t* keys, uint32_t* vals, int bits) {
    const uint32_t n = which == 0 ? c->n_next : c->n_shadow;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        uint32_t key = 0xffffffffu;
        if (i < n) {
            const float4 oo = o[i], dd = d[i];
            const uint32_t m = (1u << bits) - 1u;
            uint32_t qx = min(m, (uint32_t)max(0.0f, (oo.x - lox) * sx)), qy = min(m, (uint32_t)max(0.0f, (oo.y - loy) * sy)),
                     qz = min(m, (uint32_t)max(0.0f, (oo.z - loz) * sz));
            uint32_t cell = 0;
            for (int b = bits - 1; b >= 0; --b) cell = (cell << 3) | (((qx >> b) & 1u) << 2) | (((qy >> b) & 1u) << 1) | ((qz >> b) & 1u);
            uint32_t oct = (dd.x < 0 ? 4u : 0u) | (dd.y < 0 ? 2u : 0u) | (dd.z < 0 ? 1u : 0u);
            key = (cell << 3) | oct;
        }
        keys[i] = key; vals[i] = i;
    }
}
__global__ void k_exp_gather(const Counters* c, int which, const uint32_t* __restrict__ perm, const float4* __restrict__ a0,
                             const float4* __restrict__ a1, const float4* __restrict__ a2, float4* b0, float4* b1, float4* b2) {
    const uint32_t n = which == 0 ? c->n_next : c->n_shadow;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t j = perm[i];
        b0[i] = a0[j]; b1[i] = a1[j]; b2[i] = a2[j];
    }
}
static void exp_sort(const DeviceScene& ds, Counters* c, int which, float4* a0, float4* a1, float4* a2, uint32_t cap, cudaStream_t st) {
    static int bits = (int)env_u32("CRT_EXP_BITS", 5);
    if (bits == 0) return;
    ExpSort& e = g_exp;
    if (e.cap < cap) {
        cudaFree(e.k_in); cudaFree(e.k_out); cudaFree(e.v_in); cudaFree(e.v_out); cudaFree(e.t0); cudaFree(e.t1); cudaFree(e.t2); cudaFree(e.temp);
        cudaMalloc(&e.k_in, 4ull * cap); cudaMalloc(&e.k_out, 4ull * cap); cudaMalloc(&e.v_in, 4ull * cap); cudaMalloc(&e.v_out, 4ull * cap);
        cudaMalloc(&e.t0, 16ull * cap); cudaMalloc(&e.t1, 16ull * cap); cudaMalloc(&e.t2, 16ull * cap);
        e.temp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, e.temp_bytes, e.k_in, e.k_out, e.v_in, e.v_out, (int)cap, 0, 32, st);
        cudaMalloc(&e.temp, e.temp_bytes);
        e.cap = cap;
    }
    const float* b = ds.bounds;
    const float m = (float)(1 << bits);
    k_exp_keys<<<num_sms() * 8, 256, 0, st>>>(c, which, a0, a1, cap, b[0], b[1], b[2], m / (b[3] - b[0]), m / (b[4] - b[1]), m / (b[5] - b[2]),
                                              e.k_in, e.v_in, bits);
    size_t tb = e.temp_bytes;
    cub::DeviceRadixSort::SortPairs(e.temp, tb, e.k_in, e.k_out, e.v_in, e.v_out, (int)cap, 0, 32, st);
    k_exp_gather<<<num_sms() * 8, 256, 0, st>>>(c, which, e.v_out, a0, a1, a2, e.t0, e.t1, e.t2);
    cudaMemcpyAsync(a0, e.t0, 16ull * cap, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(a1, e.t1, 16ull * cap, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(a2, e.t2, 16ull * cap, cudaMemcpyDeviceToDevice, st);
}
