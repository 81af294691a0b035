// crt_wide.cuh — traversal of the 8-wide compressed BVH (builder CRT_BUILDER_LBVH8).
//
// Node layout (80 bytes = five 16-byte words; built by crt_bvh_build.cu, stated on the CPU by
// oracle/orc_bvh.cpp build_wide8_bvh, byte-identical):
//   w0: origin p (3 floats), {ex, ey, ez, imask} bytes      scale_k = float with exponent field e_k
//   w1: child_base, tri_base, meta[0..3], meta[4..7]        meta: 0 empty, 0x80 node, 1 + offset leaf
//   w2: qlo_x[0..7], qlo_y[0..7]   w3: qlo_z[0..7], qhi_x[0..7]   w4: qhi_y[0..7], qhi_z[0..7]
// child box k = p + q * scale per axis; the internal child in slot s is node
// child_base + popcount(imask below s); a leaf child starts at triangle slot tri_base + offset and
// runs to the terminator bit of tri_geom.
//
// Per-ray rule = oracle/orc_bvh.cpp wide8_intersect (same visits in the same order, so the oracle's
// node/triangle counts are the algorithmic ones): conservative quantised slabs, children visited
// front to back by slot ^ octant (descending), the leaf children of a node tested before descending,
// one stack entry per node with hit children left. The (t, face) result equals the pair-node BVH's
// and brute force's: conservative boxes make the closest hit independent of the tree.
#pragma once
#include "crt_device.cuh"

namespace crt {

static constexpr int kWideStack = 48;      // one entry per wide-tree level at most; the builder refuses deeper trees

// byte k of w as a float, exactly: PRMT puts the byte under the exponent of 2^23 and one FADD removes the 2^23
// (I2F.U8 is a quarter-rate XU instruction; 48 of them per node step made the wide kernels XU-bound).
CRT_DEV float wide_byte(uint32_t w, int k) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u | (uint32_t)k)) - 8388608.0f; }

// 4-bit mask of the non-zero bytes of w (bit k = byte k != 0)
CRT_DEV uint32_t wide_nonzero_bytes(uint32_t w) {
    const uint32_t t = (w | ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;      // bit 7 of every non-zero byte
    return (((t >> 7) * 0x01020408u) >> 24) & 0xfu;                                // bits 0, 8, 16, 24 -> bits 24..27
}

struct WideStep {
    uint32_t node_hits, leaf_hits;     // priority space: bit (slot ^ octant), visited in descending order
    uint32_t child_base, tri_base, meta_lo, meta_hi, imask;
};

// Box tests of the 8 children of node ni against the ray (o, inv); lim = 1.0001 * current t limit.
CRT_DEV WideStep wide_node_test(const uint4* __restrict__ nodes, uint32_t ni, V3 o, V3 inv, uint32_t oinv, float lim, bool zray) {
    const uint4* p = nodes + 5 * (size_t)ni;
    const uint4 w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3), w4 = __ldg(p + 4);
    const uint32_t ew = w0.w;
    const float sx = __uint_as_float((ew & 0xffu) << 23), sy = __uint_as_float(((ew >> 8) & 0xffu) << 23),
                sz = __uint_as_float(((ew >> 16) & 0xffu) << 23);
    const float ax = (__uint_as_float(w0.x) - o.x) * inv.x, ay = (__uint_as_float(w0.y) - o.y) * inv.y,
                az = (__uint_as_float(w0.z) - o.z) * inv.z;
    const float bx = sx * inv.x, by = sy * inv.y, bz = sz * inv.z;
    float fax = fabsf(ax), fay = fabsf(ay), faz = fabsf(az);
    if (!(fax <= FLT_MAX)) fax = 0.0f;
    if (!(fay <= FLT_MAX)) fay = 0.0f;
    if (!(faz <= FLT_MAX)) faz = 0.0f;
    const float pad = fmaxf(fmaxf(fax, fay), faz) * 4.76837158203125e-07f;        // 2^-21
    // near planes: the low ones when the direction component is non-negative
    const bool px = oinv & 1u, py = oinv & 2u, pz = oinv & 4u;
    const uint32_t nx[2] = {px ? w2.x : w3.z, px ? w2.y : w3.w}, fx[2] = {px ? w3.z : w2.x, px ? w3.w : w2.y};
    const uint32_t ny[2] = {py ? w2.z : w4.x, py ? w2.w : w4.y}, fy[2] = {py ? w4.x : w2.z, py ? w4.y : w2.w};
    const uint32_t nz[2] = {pz ? w3.x : w4.z, pz ? w3.y : w4.w}, fz[2] = {pz ? w4.z : w3.x, pz ? w4.w : w3.y};
    WideStep r;
    r.child_base = w1.x; r.tri_base = w1.y; r.meta_lo = w1.z; r.meta_hi = w1.w; r.imask = ew >> 24;
    uint32_t hits = 0;                                     // slot space: bit c = the box of child c is hit
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int w = c >> 2, k = c & 3;
        const float tnx = fmaf(wide_byte(nx[w], k), bx, ax), tfx = fmaf(wide_byte(fx[w], k), bx, ax);
        const float tny = fmaf(wide_byte(ny[w], k), by, ay), tfy = fmaf(wide_byte(fy[w], k), by, ay);
        const float tnz = fmaf(wide_byte(nz[w], k), bz, az), tfz = fmaf(wide_byte(fz[w], k), bz, az);
        const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
        const float tmax = fminf(fminf(tfx, tfy), fminf(tfz, lim));
        bool hit = tmin <= fmaf(tmax, kSlabSlack, pad);
        if (zray && hit) {                                   // parallel axes: near = lo bytes, far = hi bytes (oinv bit set)
            const float qnx = wide_byte(nx[w], k), qfx = wide_byte(fx[w], k), qny = wide_byte(ny[w], k), qfy = wide_byte(fy[w], k),
                        qnz = wide_byte(nz[w], k), qfz = wide_byte(fz[w], k);
            const float ext = (fabsf(qfx - qnx) * sx + fabsf(qfy - qny) * sy) + fabsf(qfz - qnz) * sz;
            if (inv.x != inv.x) hit = parallel_ok(fmaf(qnx, sx, __uint_as_float(w0.x)), fmaf(qfx, sx, __uint_as_float(w0.x)), o.x, tmax, ext);
            if (hit && inv.y != inv.y) hit = parallel_ok(fmaf(qny, sy, __uint_as_float(w0.y)), fmaf(qfy, sy, __uint_as_float(w0.y)), o.y, tmax, ext);
            if (hit && inv.z != inv.z) hit = parallel_ok(fmaf(qnz, sz, __uint_as_float(w0.z)), fmaf(qfz, sz, __uint_as_float(w0.z)), o.z, tmax, ext);
        }
        if (hit) hits |= 1u << c;
    }
    // children that exist (meta byte != 0), 4 bits per meta word, then node / leaf children from imask
    hits &= wide_nonzero_bytes(w1.z) | (wide_nonzero_bytes(w1.w) << 4);
    // slot space -> priority space (bit i -> bit i ^ oinv) for both masks at once: node hits in bits 0-7, leaf hits in 8-15
    uint32_t both = (hits & r.imask) | ((hits & ~r.imask) << 8);
    if (oinv & 1u) both = ((both & 0x5555u) << 1) | ((both >> 1) & 0x5555u);
    if (oinv & 2u) both = ((both & 0x3333u) << 2) | ((both >> 2) & 0x3333u);
    if (oinv & 4u) both = ((both & 0x0f0fu) << 4) | ((both >> 4) & 0x0f0fu);
    r.node_hits = both & 0xffu;
    r.leaf_hits = both >> 8;
    return r;
}

CRT_DEV uint32_t wide_octant(V3 inv) { return (inv.x < 0.0f ? 0u : 1u) | (inv.y < 0.0f ? 0u : 2u) | (inv.z < 0.0f ? 0u : 4u); }
// first triangle slot of the leaf child with priority bit lp
CRT_DEV int wide_leaf_slot(const WideStep& s, int lp, uint32_t oinv) {
    const uint32_t c = (uint32_t)lp ^ oinv;
    const uint32_t m = ((c & 4u ? s.meta_hi : s.meta_lo) >> (8u * (c & 3u))) & 0xffu;
    return (int)(s.tri_base + m - 1u);
}

// Sequential rule, one ray per call (tail kernel; statement: oracle wide8_intersect).
template <int MODE>
CRT_DEV HitRec traverse_wide(const SceneView& sc, V3 o, V3 d, float tmax) {
    HitRec best;
    best.t = FLT_MAX; best.slot = -1; best.face = -1;
    if (sc.n_nodes == 0) return best;
    const uint4* nodes = (const uint4*)sc.nodes;
    const V3 inv = box_inv3(d);
    const uint32_t oinv = wide_octant(inv);
    const bool zray = has_parallel_axis(inv);
    float tlimit = MODE == 0 ? FLT_MAX : tmax;
    uint2 stack[kWideStack];
    int sp = 0;
    uint32_t g_base = 0, g_bits = (1u << 8) | (1u << oinv);      // imask << 8 | hit mask; the root as slot 0 of a virtual parent
    for (;;) {
        if ((g_bits & 0xffu) == 0u) {
            if (sp == 0) break;
            const uint2 e = stack[--sp];
            g_base = e.x; g_bits = e.y;
        }
        const int pr = 31 - __clz((int)(g_bits & 0xffu));
        g_bits &= ~(1u << pr);
        const uint32_t sl = (uint32_t)pr ^ oinv;
        const uint32_t ni = g_base + (uint32_t)__popc((g_bits >> 8) & ((1u << sl) - 1u));
        WideStep s = wide_node_test(nodes, ni, o, inv, oinv, tlimit * 1.0001f, zray);
        while (s.leaf_hits) {
            const int lp = 31 - __clz((int)s.leaf_hits);
            s.leaf_hits &= ~(1u << lp);
            int slot = wide_leaf_slot(s, lp, oinv);
            for (;; ++slot) {
                V3 tv1, te1, te2;
                const uint32_t fw = load_tri(sc.tri_geom, slot, tv1, te1, te2);
                const int face = (int)(fw & ~kLastBit);
                float t;
                if (tri_test(tv1, te1, te2, o, d, &t) && t > kEps) {
                    if (MODE == 0) {
                        if (t < best.t || (t == best.t && face < best.face)) {
                            best.t = t; best.slot = slot; best.face = face;
                            tlimit = t;
                        }
                    } else if (tmax - t > kEps) {
                        best.t = t; best.slot = slot; best.face = face;
                        return best;
                    }
                }
                if (fw & kLastBit) break;
            }
        }
        if (s.node_hits) {
            if (g_bits & 0xffu) stack[sp++] = make_uint2(g_base, g_bits);
            g_base = s.child_base;
            g_bits = (s.imask << 8) | s.node_hits;
        }
    }
    return best;
}

// ---- persistent lanes, while-while (WSTRAT 0): every lane walks wide nodes until one of them has leaf
// children hit, then the lanes test those leaves. Same per-ray order as traverse_wide.
template <int MODE, typename Load, typename Done>
CRT_DEV void trace_persistent_wide(const SceneView& sc, uint32_t n, uint32_t* fetch, Load load, Done done) {
    const uint4* nodes = (const uint4*)sc.nodes;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint2 stack[kWideStack];
    int sp = 0;
    uint32_t g_base = 0, g_bits = 0, oinv = 0;
    uint32_t idx = 0;
    V3 o = mk3(0, 0, 0), d = mk3(0, 0, 1), inv = mk3(0, 0, 0);
    float tmax = 0.0f, tlimit = 0.0f;
    HitRec best;
    best.t = FLT_MAX; best.slot = -1; best.face = -1;
    WideStep s;
    s.node_hits = s.leaf_hits = s.child_base = s.tri_base = s.meta_lo = s.meta_hi = s.imask = 0;
    bool have = false, exhausted = false, zray = false;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !have);
        if (idle) {
            const int n_idle = __popc(idle);
            if (!exhausted && (n_idle >= kRefillLanes || n_idle == 32)) {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(fetch, (uint32_t)n_idle);
                base = __shfl_sync(0xffffffffu, base, leader);
                if (!have) {
                    const uint32_t i = base + __popc(idle & lt_mask);
                    if (i < n) {
                        idx = i;
                        const bool live = load(i, o, d, tmax);
                        inv = box_inv3(d);
                        oinv = wide_octant(inv);
                        zray = has_parallel_axis(inv);
                        tlimit = MODE == 0 ? FLT_MAX : tmax;
                        best.t = FLT_MAX; best.slot = -1; best.face = -1;
                        sp = 0;
                        g_base = 0;
                        g_bits = (live && sc.n_nodes) ? ((1u << 8) | (1u << oinv)) : 0u;
                        s.leaf_hits = 0;
                        have = true;
                    }
                }
                if (base + (uint32_t)n_idle >= n) exhausted = true;
            }
            if (!__any_sync(0xffffffffu, have)) {
                if (exhausted) break;
                continue;
            }
        }
        if (have) {
            while (s.leaf_hits == 0u) {
                if ((g_bits & 0xffu) == 0u) {
                    if (sp == 0) break;
                    const uint2 e = stack[--sp];
                    g_base = e.x; g_bits = e.y;
                }
                const int pr = 31 - __clz((int)(g_bits & 0xffu));
                g_bits &= ~(1u << pr);
                const uint32_t sl = (uint32_t)pr ^ oinv;
                const uint32_t ni = g_base + (uint32_t)__popc((g_bits >> 8) & ((1u << sl) - 1u));
                s = wide_node_test(nodes, ni, o, inv, oinv, tlimit * 1.0001f, zray);
                if (s.node_hits) {
                    if (g_bits & 0xffu) stack[sp++] = make_uint2(g_base, g_bits);
                    g_base = s.child_base;
                    g_bits = (s.imask << 8) | s.node_hits;
                }
            }
            bool stop = false;
            while (s.leaf_hits) {
                const int lp = 31 - __clz((int)s.leaf_hits);
                s.leaf_hits &= ~(1u << lp);
                int slot = wide_leaf_slot(s, lp, oinv);
                for (;; ++slot) {
                    V3 tv1, te1, te2;
                    const uint32_t fw = load_tri(sc.tri_geom, slot, tv1, te1, te2);
                    const int face = (int)(fw & ~kLastBit);
                    float t;
                    if (tri_test(tv1, te1, te2, o, d, &t) && t > kEps) {
                        if (MODE == 0) {
                            if (t < best.t || (t == best.t && face < best.face)) {
                                best.t = t; best.slot = slot; best.face = face;
                                tlimit = t;
                            }
                        } else if (tmax - t > kEps) {
                            best.t = t; best.slot = slot; best.face = face;
                            stop = true;
                            break;
                        }
                    }
                    if (fw & kLastBit) break;
                }
                if (stop) break;
            }
            if (stop) { s.leaf_hits = 0; g_bits = 0; sp = 0; }
            if (s.leaf_hits == 0u && (g_bits & 0xffu) == 0u && sp == 0) {
                done(idx, best);
                have = false;
            }
        }
    }
}

// ---- persistent lanes + per-warp leaf queue (WSTRAT 2), the wide-node version of
// trace_persistent_queue (crt_device.cuh): the leaf children hit by a node step are appended to a queue
// in shared memory and the lane goes on with its next node; when kWQFlush leaves are queued (or no lane
// has a node left) the whole warp tests them, one entry per lane; owners' best hits are combined with a
// 64-bit shared atomicMin on (t bits, face id). Lanes walk with the t-limit of the last flush.
#ifndef CRT_WQFLUSH
#define CRT_WQFLUSH 12
#endif
#ifndef CRT_WQSTEPS
#define CRT_WQSTEPS 2
#endif
// 1: queue positions from four bit-sliced ballots instead of the shared atomicAdd. Measured slower here (-1 %, B200,
// profiles/r01_s18.md): the wide kernels are bound by ALU issue, not by the L1 data pipe, unlike the pair-node kernels.
#ifndef CRT_WQBALLOT
#define CRT_WQBALLOT 0
#endif
#ifndef CRT_WSSTACK
#define CRT_WSSTACK 4
#endif
// Measured and removed: prefetch.global.L1 of a leaf's triangles when it is queued, -5 % (profiles/r01_s21.md).
static constexpr int kWQFlush = CRT_WQFLUSH;
static constexpr int kWQSteps = CRT_WQSTEPS;
static constexpr int kWQCap = kWQFlush + 32 * 8 * kWQSteps;
struct WarpLeafQueueW {
    float ox[32], oy[32], oz[32], dx[32], dy[32], dz[32], tmax[32];
    unsigned long long best[32];
    int best_slot[32];
    int q_slot[kWQCap];
    unsigned char q_lane[kWQCap];
    int count;
};

template <int MODE, typename Load, typename Done>
CRT_DEV void trace_persistent_wide_queue(const SceneView& sc, uint32_t n, uint32_t* fetch, Load load, Done done) {
    __shared__ WarpLeafQueueW s_wq[4];                     // launched with 128 threads per block
    WarpLeafQueueW& q = s_wq[threadIdx.x >> 5];
    const uint4* nodes = (const uint4*)sc.nodes;
    const unsigned kFull = 0xffffffffu;
    const unsigned long long kNoHit = ((unsigned long long)0x7f7fffffu << 32) | 0x7fffffffull;   // t = FLT_MAX
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
#if CRT_WSSTACK
    // first CRT_WSSTACK entries of every lane's stack in shared memory, [entry][thread]: a push / pop is one
    // conflict-free wavefront and a short-scoreboard wait instead of a local-memory round trip
    uint2 stack[kWideStack - CRT_WSSTACK];
    __shared__ uint2 s_wstack[CRT_WSSTACK][128];
#else
    uint2 stack[kWideStack];
#endif
    int sp = 0;
    uint32_t g_base = 0, g_bits = 0, oinv = 0;
    uint32_t idx = 0;
    V3 o = mk3(0, 0, 0), inv = mk3(0, 0, 0);
    float tlimit = 0.0f;
    int pending = 0;
#if CRT_WQBALLOT
    int qn = 0;                                            // queued leaves, the same value in every lane
#endif
    bool have = false, exhausted = false, zray = false;
    if (lane == 0) q.count = 0;
    __syncwarp();
    for (;;) {
        // A. node steps; leaf children go to the queue
#pragma unroll
        for (int r = 0; r < kWQSteps; ++r) {
            if ((g_bits & 0xffu) == 0u && sp > 0) {
                --sp;
#if CRT_WSSTACK
                const uint2 e = sp < CRT_WSSTACK ? s_wstack[sp][threadIdx.x] : stack[sp - CRT_WSSTACK];
#else
                const uint2 e = stack[sp];
#endif
                g_base = e.x; g_bits = e.y;
            }
#if CRT_WQBALLOT
            int cnt = 0;
            uint32_t leaf_bits = 0;
            WideStep lstep;
            lstep.tri_base = lstep.meta_lo = lstep.meta_hi = 0;
#endif
            if (g_bits & 0xffu) {
                const int pr = 31 - __clz((int)(g_bits & 0xffu));
                g_bits &= ~(1u << pr);
                const uint32_t sl = (uint32_t)pr ^ oinv;
                const uint32_t ni = g_base + (uint32_t)__popc((g_bits >> 8) & ((1u << sl) - 1u));
                WideStep s = wide_node_test(nodes, ni, o, inv, oinv, tlimit * 1.0001f, zray);
#if CRT_WQBALLOT
                cnt = __popc(s.leaf_hits);
                leaf_bits = s.leaf_hits;
                lstep = s;
#else
                if (s.leaf_hits) {
                    const int cnt = __popc(s.leaf_hits);
                    int pos = atomicAdd(&q.count, cnt);
                    pending += cnt;
                    while (s.leaf_hits) {
                        const int lp = 31 - __clz((int)s.leaf_hits);
                        s.leaf_hits &= ~(1u << lp);
                        const int lslot = wide_leaf_slot(s, lp, oinv);
                        q.q_slot[pos] = lslot;
                        q.q_lane[pos] = (unsigned char)lane;
                        ++pos;
                    }
                }
#endif
                if (s.node_hits) {
                    if (g_bits & 0xffu) {
#if CRT_WSSTACK
                        if (sp < CRT_WSSTACK) s_wstack[sp][threadIdx.x] = make_uint2(g_base, g_bits);
                        else stack[sp - CRT_WSSTACK] = make_uint2(g_base, g_bits);
#else
                        stack[sp] = make_uint2(g_base, g_bits);
#endif
                        ++sp;
                    }
                    g_base = s.child_base;
                    g_bits = (s.imask << 8) | s.node_hits;
                }
            }
#if CRT_WQBALLOT
            {   // the queue belongs to this warp and every lane is here: exclusive prefix of cnt (<= 8, four bits) from four
                // ballots instead of a shared atomic that serialises the lanes
                const unsigned b0 = __ballot_sync(kFull, cnt & 1), b1 = __ballot_sync(kFull, cnt & 2),
                               b2 = __ballot_sync(kFull, cnt & 4), b3 = __ballot_sync(kFull, cnt & 8);
                int pos = qn + __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask) + 8 * __popc(b3 & lt_mask);
                qn += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2) + 8 * __popc(b3);
                pending += cnt;
                while (leaf_bits) {
                    const int lp = 31 - __clz((int)leaf_bits);
                    leaf_bits &= ~(1u << lp);
                    q.q_slot[pos] = wide_leaf_slot(lstep, lp, oinv);
                    q.q_lane[pos] = (unsigned char)lane;
                    ++pos;
                }
            }
#endif
        }
        // B. flush the leaf queue when it is full enough, or when no lane has a node in hand
        const unsigned walking = __ballot_sync(kFull, (g_bits & 0xffu) != 0u || sp > 0);
        __syncwarp();
#if CRT_WQBALLOT
        const int q_count = qn;
#else
        const int q_count = q.count;
#endif
        if (q_count >= kWQFlush || (walking == 0 && q_count > 0)) {
            for (int base = 0; base < q_count; base += 32) {
                const int k = base + lane;
                unsigned long long mykey = kNoHit;
                int myslot = -1, owner = 0;
                if (k < q_count) {
                    owner = q.q_lane[k];
                    int slot = q.q_slot[k];
                    const V3 ro = mk3(q.ox[owner], q.oy[owner], q.oz[owner]), rd = mk3(q.dx[owner], q.dy[owner], q.dz[owner]);
                    const float rtmax = q.tmax[owner];
                    for (;; ++slot) {
                        V3 tv1, te1, te2;
                        const uint32_t fw = load_tri(sc.tri_geom, slot, tv1, te1, te2);
                        float t;
                        if (tri_test(tv1, te1, te2, ro, rd, &t) && t > kEps && (MODE == 0 || rtmax - t > kEps)) {
                            const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (fw & ~kLastBit);
                            if (key < mykey) { mykey = key; myslot = slot; }
                        }
                        if (fw & kLastBit) break;
                    }
                    if (myslot >= 0) {
                        if (MODE == 0) atomicMin(&q.best[owner], mykey);
                        else q.best[owner] = mykey;                 // any blocker will do: one of the writers wins (64-bit store)
                    }
                }
                __syncwarp();
                if (myslot >= 0 && q.best[owner] == mykey) q.best_slot[owner] = myslot;
                __syncwarp();
            }
#if CRT_WQBALLOT
            qn = 0;
#else
            if (lane == 0) q.count = 0;
#endif
            pending = 0;
            if (have) {
                const unsigned long long b = q.best[lane];
                if (MODE == 0) tlimit = __uint_as_float((uint32_t)(b >> 32));
                else if (b != kNoHit) { g_bits = 0; sp = 0; }      // blocked: nothing left to learn
            }
            __syncwarp();
        }
        // C. finished rays
        if (have && (g_bits & 0xffu) == 0u && sp == 0 && pending == 0) {
            const unsigned long long b = q.best[lane];
            HitRec h;
            h.t = __uint_as_float((uint32_t)(b >> 32));
            h.face = b == kNoHit ? -1 : (int)(uint32_t)b;
            h.slot = b == kNoHit ? -1 : q.best_slot[lane];
            done(idx, h);
            have = false;
        }
        // D. refill idle lanes from the ray queue
        const unsigned idle = __ballot_sync(kFull, !have);
        if (idle) {
            const int n_idle = __popc(idle);
            if (!exhausted && (n_idle >= kRefillLanes || n_idle == 32)) {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(fetch, (uint32_t)n_idle);
                base = __shfl_sync(kFull, base, leader);
                const uint32_t my_i = base + __popc(idle & lt_mask);
                if (!have) {
                    const uint32_t i = my_i;
                    if (i < n) {
                        idx = i;
                        V3 d;
                        float tmax;
                        const bool live = load(i, o, d, tmax);
                        inv = box_inv3(d);
                        oinv = wide_octant(inv);
                        zray = has_parallel_axis(inv);
                        tlimit = MODE == 0 ? FLT_MAX : tmax;
                        q.ox[lane] = o.x; q.oy[lane] = o.y; q.oz[lane] = o.z;
                        q.dx[lane] = d.x; q.dy[lane] = d.y; q.dz[lane] = d.z;
                        q.tmax[lane] = tmax;
                        q.best[lane] = kNoHit;
                        q.best_slot[lane] = -1;
                        sp = 0;
                        pending = 0;
                        g_base = 0;
                        g_bits = (live && sc.n_nodes) ? ((1u << 8) | (1u << oinv)) : 0u;
                        have = true;
                    }
                }
                if (base + (uint32_t)n_idle >= n) exhausted = true;
                __syncwarp();
            }
            if (idle == kFull && !__any_sync(kFull, have)) {
                if (exhausted) break;
            }
        }
    }
}

template <int MODE, int STRAT, typename Load, typename Done>
CRT_DEV void trace_rays_persistent_wide(const SceneView& sc, uint32_t n, uint32_t* fetch, Load load, Done done) {
    if (STRAT == 2) trace_persistent_wide_queue<MODE>(sc, n, fetch, load, done);
    else trace_persistent_wide<MODE>(sc, n, fetch, load, done);
}

}  // namespace crt
