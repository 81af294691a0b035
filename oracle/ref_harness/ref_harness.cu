// ORACLE — TEST INFRASTRUCTURE ONLY.
// Headless harness around the UNMODIFIED reference headers, included in place from
// /root/reference/include (never copied). Built by oracle/ref_harness/build.sh into
// oracle/_ref/libref.so. It does what src/main.cu:119-146,276,282 and Render.cuh:379-440 do,
// minus GLFW/ImGui/PBO: load OBJ -> Scene -> BVH -> DeviceBVH/DeviceLights -> view_render_kernel.
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "Loader.h"
#include "Scene.h"
#include "Camera.h"
#include "Render.cuh"
#include <chrono>

namespace {
struct RefState {
    Scene* scene = nullptr;
    DeviceBVH* host_bvh = nullptr;
    DeviceBVH* device_bvh = nullptr;
    DeviceLights* host_lights = nullptr;
    DeviceLights* device_lights = nullptr;
    uchar3* device_frame = nullptr;
    DeviceStack<int, BVH_STACK_SIZE>* bvh_stacks = nullptr;
    DeviceStack<HitPayload, BOUNCE_STACK_SIZE>* bounce_stacks = nullptr;
    std::vector<Triangle> scene_order;     // triangles before the BVH sorts them in place
    double ms_parse = 0, ms_load = 0, ms_bvh = 0;
    unsigned width = 0, height = 0;
};
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
void put3(float* o, const Eigen::Vector3f& v) { o[0] = v.x(); o[1] = v.y(); o[2] = v.z(); }
}  // namespace

// closest-hit through the reference's own traversal, one thread per ray (C5 baseline, t cross-check)
__global__ void ref_trace_kernel(DeviceBVH* bvh, DeviceStack<int, BVH_STACK_SIZE>* stacks, const float* rays, long long n,
                                 float* t_out, int stack_slots) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        const float* r = rays + 8 * k;
        Ray ray(Eigen::Vector3f(r[0], r[1], r[2]), Eigen::Vector3f(r[4], r[5], r[6]));
        HitPayload hit = bvh->intersect(ray.get_origin(), ray.get_dir(), ray.get_inv_dir(),
                                        &stacks[(blockIdx.x * (long long)blockDim.x + threadIdx.x) % stack_slots]);
        t_out[k] = hit.t;
    }
}

extern "C" {

// Host path: main.cu:121-146 + :276. Returns an opaque state; no CUDA call is made.
void* ref_host_load(const char* obj_path, const char* mtl_dir, unsigned width, unsigned height, unsigned thresh_n) {
    RefState* st = new RefState();
    st->width = width; st->height = height;
    st->scene = new Scene(width, height);
    Loader loader;
    std::vector<Triangle> triangles, light_triangles;
    double t0 = now_ms();
    loader.read_OBJ(obj_path, mtl_dir);
    double t1 = now_ms();
    for (uint64_t i = 0; i < loader.size(); i++) {
        loader.load_object(i, triangles, light_triangles);
        if (triangles.size() > 0) { Object o(triangles); st->scene->add_normal_obj(o); }
        if (light_triangles.size() > 0) { Object o(light_triangles); st->scene->add_light_obj(o); }
    }
    double t2 = now_ms();
    st->scene_order = st->scene->get_triangles();
    double t3 = now_ms();
    st->scene->set_BVH(thresh_n);
    double t4 = now_ms();
    st->ms_parse = t1 - t0; st->ms_load = t2 - t1; st->ms_bvh = t4 - t3;
    return st;
}
// C4 / C5 (synthetic scenes that exist only as arrays): the triangles go through the reference's public classes the
// way Loader::load_object (Loader.h:107-120) and main.cu:131-144 feed them - Material, Triangle, one Object per
// (contiguous) obj id, add_light_obj when the material emits - and then through Scene::set_BVH, which is timed.
// verts n*9, mat / obj per triangle, mats7 rows of kd(3) ke(3) ns. No OBJ text is involved (ms_parse = 0).
void* ref_host_from_arrays(const float* verts, const int* mat, const int* obj, long long n, const float* mats7, int n_mats,
                           unsigned width, unsigned height) {
    RefState* st = new RefState();
    st->width = width; st->height = height;
    st->scene = new Scene(width, height);
    double t0 = now_ms();
    std::vector<Triangle> group;
    auto flush = [&](bool emits) {
        if (group.empty()) return;
        Object o(group);
        if (emits) st->scene->add_light_obj(o); else st->scene->add_normal_obj(o);
        group.clear();
    };
    bool cur_emits = false;
    for (long long i = 0; i < n; ++i) {
        const float* m7 = mats7 + 7 * (size_t)mat[i];
        Eigen::Vector3f kd(m7[0], m7[1], m7[2]), ke(m7[3], m7[4], m7[5]), zero(0, 0, 0);
        Material m(kd, zero, zero, ke, m7[6], m7[6] > 1 ? SPECULAR : DIFFUSE);
        if (i > 0 && (obj[i] != obj[i - 1] || m.has_emission() != cur_emits)) flush(cur_emits);
        cur_emits = m.has_emission();
        const float* v = verts + 9 * (size_t)i;
        group.push_back(Triangle(Eigen::Vector3f(v[0], v[1], v[2]), Eigen::Vector3f(v[3], v[4], v[5]), Eigen::Vector3f(v[6], v[7], v[8]),
                                 Eigen::Vector3f(0, 1, 0), m));
    }
    flush(cur_emits);
    (void)n_mats;
    st->ms_load = now_ms() - t0;
    return st;
}
// Scene::set_BVH (Scene.h:50-54 -> BVH.h:30-84) on a state made by ref_host_from_arrays; returns the wall-clock ms.
double ref_build_bvh(void* h, unsigned thresh_n) {
    RefState* st = (RefState*)h;
    double t0 = now_ms();
    st->scene->set_BVH(thresh_n);
    st->ms_bvh = now_ms() - t0;
    return st->ms_bvh;
}

void ref_host_times(void* h, double* ms3) { RefState* st = (RefState*)h; ms3[0] = st->ms_parse; ms3[1] = st->ms_load; ms3[2] = st->ms_bvh; }
int ref_n_tris(void* h) { return (int)((RefState*)h)->scene->get_triangles().size(); }
int ref_n_nodes(void* h) { return (int)((RefState*)h)->scene->get_bvh().get_nodes_size(); }
int ref_root(void* h) { return ((RefState*)h)->scene->get_bvh().get_root_index(); }
int ref_n_lights(void* h) { return (int)((RefState*)h)->scene->get_light_objs().size(); }
int ref_sizeof_node() { return (int)sizeof(BVHNode); }
// sorted != 0: after the BVH's in-place sort; else scene order. 20 floats per triangle:
// v1 v2 v3 normal(3) area area_of_obj kd(3) ke(3)  then ns has_emit mode  -> 23 floats
void ref_get_tris(void* h, int sorted, float* out23) {
    RefState* st = (RefState*)h;
    const std::vector<Triangle>& ts = sorted ? st->scene->get_triangles() : st->scene_order;
    for (size_t i = 0; i < ts.size(); ++i) {
        float* o = out23 + 23 * i;
        const Triangle& t = ts[i];
        put3(o, t.get_v1()); put3(o + 3, t.get_v2()); put3(o + 6, t.get_v3()); put3(o + 9, t.get_normal());
        o[12] = t.get_area(); o[13] = t.get_area_of_obj();
        Material m = t.get_material();
        put3(o + 14, m.get_kd()); put3(o + 17, m.get_ke());
        o[20] = m.get_ns(); o[21] = m.has_emission() ? 1.0f : 0.0f; o[22] = (float)m.get_mode();
    }
}
void ref_get_nodes(void* h, void* out) {
    RefState* st = (RefState*)h;
    memcpy(out, st->scene->get_bvh().get_nodes_data(), sizeof(BVHNode) * st->scene->get_bvh().get_nodes_size());
}
void ref_get_light(void* h, int li, int* n_tris, float* area) {
    Object& o = ((RefState*)h)->scene->get_light_objs()[li];
    *n_tris = (int)o.get_triangles().size();
    *area = o.get_triangles()[0].get_area_of_obj();
}
void ref_inverse_view(const float* eye, const float* lookat, const float* up, float* out9) {
    Eigen::Matrix3f m = get_inverse_view_matrix(Eigen::Vector3f(eye[0], eye[1], eye[2]), Eigen::Vector3f(lookat[0], lookat[1], lookat[2]),
                                                Eigen::Vector3f(up[0], up[1], up[2]));
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out9[3 * r + c] = m(r, c);
}

// Device path: Render::Render (Render.cuh:379-433) without the GL parts. Returns 0 on success.
int ref_device_init(void* h) {
    RefState* st = (RefState*)h;
    cudaError_t err;
    st->host_bvh = new DeviceBVH(st->scene->get_bvh());
    if ((err = cudaMalloc((void**)&st->device_bvh, sizeof(DeviceBVH))) != cudaSuccess) return -1;
    cudaMemcpy(st->device_bvh, st->host_bvh, sizeof(DeviceBVH), cudaMemcpyHostToDevice);
    if ((err = cudaMalloc((void**)&st->device_frame, sizeof(uchar3) * st->scene->get_pixels())) != cudaSuccess) return -2;
    st->host_lights = new DeviceLights(st->scene->get_light_objs());
    if ((err = cudaMalloc((void**)&st->device_lights, sizeof(DeviceLights))) != cudaSuccess) return -3;
    cudaMemcpy((void*)st->device_lights, (void*)st->host_lights, sizeof(DeviceLights), cudaMemcpyHostToDevice);
    if ((err = cudaMalloc((void**)&st->bvh_stacks, sizeof(DeviceStack<int, BVH_STACK_SIZE>) * (size_t)st->scene->get_pixels())) != cudaSuccess) return -4;
    if ((err = cudaMalloc((void**)&st->bounce_stacks, sizeof(DeviceStack<HitPayload, BOUNCE_STACK_SIZE>) * (size_t)st->scene->get_pixels())) != cudaSuccess) return -5;
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -6;
}

// Render::run_view (Render.cuh:435-440,464): same 16x16 launch, sync, D2H of the RGB8 frame.
// kernel_ms: cudaEvent time of the kernel; wall_ms: kernel + sync + D2H like main.cu:370-376.
int ref_render(void* h, const float* eye, const float* inv_view9, float fovy_rad, unsigned spp, float p_rr, int lsn,
               unsigned char* rgb8_out, float* kernel_ms, double* wall_ms) {
    RefState* st = (RefState*)h;
    Eigen::Vector3f e(eye[0], eye[1], eye[2]);
    Eigen::Matrix3f m;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m(r, c) = inv_view9[3 * r + c];
    dim3 threadsPerBlock(16, 16);
    dim3 numBlocks((st->width + 15) / 16, (st->height + 15) / 16);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double w0 = now_ms();
    cudaEventRecord(a);
    view_render_kernel<<<numBlocks, threadsPerBlock>>>(st->width, st->height, e, m, fovy_rad, spp, p_rr, lsn, st->device_bvh,
                                                        st->device_frame, st->device_lights, st->bounce_stacks, st->bvh_stacks);
    cudaEventRecord(b);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return -1; }
    if (rgb8_out) cudaMemcpy(rgb8_out, st->device_frame, sizeof(uchar3) * st->scene->get_pixels(), cudaMemcpyDeviceToHost);
    double w1 = now_ms();
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    if (kernel_ms) *kernel_ms = ms;
    if (wall_ms) *wall_ms = w1 - w0;
    cudaEventDestroy(a); cudaEventDestroy(b);
    return 0;
}

// rays: host n*8 floats (o, tmax, d, pad); t_out host n floats. Uses the per-pixel stacks.
int ref_trace(void* h, const float* rays, long long n, float* t_out, float* kernel_ms) {
    RefState* st = (RefState*)h;
    float *d_rays, *d_t;
    if (cudaMalloc(&d_rays, sizeof(float) * 8 * n) != cudaSuccess) return -1;
    if (cudaMalloc(&d_t, sizeof(float) * n) != cudaSuccess) return -1;
    cudaMemcpy(d_rays, rays, sizeof(float) * 8 * n, cudaMemcpyHostToDevice);
    int slots = (int)st->scene->get_pixels();
    int block = 256, grid = slots / block; if (grid < 1) { grid = 1; block = slots; }
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    ref_trace_kernel<<<grid, block>>>(st->device_bvh, st->bvh_stacks, d_rays, n, d_t, grid * block);
    cudaEventRecord(b);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return -2; }
    float ms = 0; cudaEventElapsedTime(&ms, a, b); if (kernel_ms) *kernel_ms = ms;
    cudaMemcpy(t_out, d_t, sizeof(float) * n, cudaMemcpyDeviceToHost);
    cudaFree(d_rays); cudaFree(d_t); cudaEventDestroy(a); cudaEventDestroy(b);
    return 0;
}

void ref_free(void* h) {
    RefState* st = (RefState*)h;
    if (st->host_bvh) { st->host_bvh->free(); cudaFree(st->device_bvh); }
    if (st->host_lights) { st->host_lights->free(); cudaFree(st->device_lights); }
    cudaFree(st->device_frame); cudaFree(st->bvh_stacks); cudaFree(st->bounce_stacks);
    st->scene->free();
    delete st->scene;
    delete st;
}

}  // extern "C"
