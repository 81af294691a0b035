// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h). Never linked into the product.
#include "orc_bvh.h"
#include <algorithm>
#include <numeric>

namespace orc {

// =============================================================================================
// Reference host BVH — BVH.h:37-84. Sorting an index array with std::sort and the same
// comparison outcomes yields the same permutation as the reference's std::sort of Triangle
// objects (introsort's moves depend only on comparison results), given the same libstdc++.
// =============================================================================================
namespace {
struct RefBuilder {
    const Scene& s;
    unsigned thresh;
    RefBVH& out;
    int build(int l, int r) {
        if (l >= r) return -1;
        RefNode node;
        node.lc = -1; node.rc = -1; node.it = -1; node.n = 0;
        for (int k = 0; k < 3; ++k) { node.AA[k] = FLT_MAX; node.BB[k] = -FLT_MAX; }
        for (int i = l; i < r; ++i) {                       // BVH.h:43-52
            const Tri& t = s.tris[out.order[i]];
            node.AA[0] = std::min(t.lo.x, node.AA[0]);
            node.AA[1] = std::min(t.lo.y, node.AA[1]);
            node.AA[2] = std::min(t.lo.z, node.AA[2]);
            node.BB[0] = std::max(t.hi.x, node.BB[0]);
            node.BB[1] = std::max(t.hi.y, node.BB[1]);
            node.BB[2] = std::max(t.hi.z, node.BB[2]);
        }
        node.it = l;
        node.n = (unsigned)(r - l);
        if (node.n <= thresh) {                             // BVH.h:57-61
            out.nodes.push_back(node);
            return (int)out.nodes.size() - 1;
        }
        float dx = node.BB[0] - node.AA[0], dy = node.BB[1] - node.AA[1], dz = node.BB[2] - node.AA[2];
        const Scene& sc = s;                                // BVH.h:64-76
        auto b = out.order.begin();
        if (dx >= dy && dx >= dz)
            std::sort(b + l, b + r, [&sc](int a, int c) { return sc.tris[a].center.x < sc.tris[c].center.x; });
        else if (dy >= dx && dy >= dz)
            std::sort(b + l, b + r, [&sc](int a, int c) { return sc.tris[a].center.y < sc.tris[c].center.y; });
        else if (dz >= dx && dz >= dy)
            std::sort(b + l, b + r, [&sc](int a, int c) { return sc.tris[a].center.z < sc.tris[c].center.z; });
        int mid = (l + r) / 2;                              // BVH.h:79-83
        node.lc = build(l, mid);
        node.rc = build(mid, r);
        out.nodes.push_back(node);
        return (int)out.nodes.size() - 1;
    }
};
}  // namespace

void build_ref_bvh(const Scene& s, unsigned thresh_n, RefBVH& out) {
    out.nodes.clear();
    out.order.resize(s.tris.size());
    std::iota(out.order.begin(), out.order.end(), 0);
    RefBuilder rb{s, thresh_n, out};
    out.root = rb.build(0, (int)s.tris.size());
}

// =============================================================================================
// Canonical triangle test — the arithmetic of DeviceTriangle.cuh:39-65 under the contract of
// orc_math.h (dot/cross with explicit fmaf).
// =============================================================================================
bool tri_test(const Tri& tr, V3 o, V3 d, float* t_out) {
    V3 e1 = tr.v2 - tr.v1, e2 = tr.v3 - tr.v1;              // DeviceTriangle.cuh:27-28
    V3 sv = o - tr.v1;
    V3 s1 = cross(d, e2);
    V3 s2 = cross(sv, e1);
    float rcp = 1.0f / dot(s1, e1);
    float beta = dot(s1, sv) * rcp;
    float gamma = dot(s2, d) * rcp;
    float t = dot(s2, e2) * rcp;
    float alpha = (1.0f - beta) - gamma;
    *t_out = t;
    return 0.0f < alpha && alpha < 1.0f && 0.0f < beta && beta < 1.0f && 0.0f < gamma && gamma < 1.0f;
}

static const float kEps = 0.00001f;     // Global.h:11

// DeviceBVH.cuh:87-126, literally (including the NaN behaviour of the ?: min/max).
static inline bool ref_hit_aabb(const RefNode& n, V3 o, V3 d, V3 inv) {
    float tminx = (n.AA[0] - o.x) * inv.x, tminy = (n.AA[1] - o.y) * inv.y, tminz = (n.AA[2] - o.z) * inv.z;
    float tmaxx = (n.BB[0] - o.x) * inv.x, tmaxy = (n.BB[1] - o.y) * inv.y, tmaxz = (n.BB[2] - o.z) * inv.z;
    if (d.x < 0) std::swap(tminx, tmaxx);
    if (d.y < 0) std::swap(tminy, tmaxy);
    if (d.z < 0) std::swap(tminz, tmaxz);
    auto mx = [](float x, float y) { return x > y ? x : y; };
    auto mn = [](float x, float y) { return x < y ? x : y; };
    float t_enter = mx(mx(tminx, tminy), tminz);
    float t_exit = mn(mn(tmaxx, tmaxy), tmaxz);
    return t_enter <= t_exit + kEps && t_exit >= 0;
}

Hit ref_intersect(const Scene& s, const RefBVH& b, V3 o, V3 d, bool canonical, TraceStats* st) {
    Hit best{FLT_MAX, -1};
    if (b.root < 0) return best;
    V3 inv{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};            // Ray.cuh:14
    std::vector<int> stack;
    stack.reserve(64);
    stack.push_back(b.root);
    uint64_t max_stack = 1;
    while (!stack.empty()) {
        int cur = stack.back();
        stack.pop_back();
        if (cur < 0) continue;
        const RefNode& node = b.nodes[cur];
        if (st) st->inner++;
        if (node.lc < 0 && node.rc < 0) {                   // DeviceBVH.cuh:141-148, :31-43
            Hit leaf{FLT_MAX, -1};
            for (int i = node.it; i < (int)(node.it + node.n); ++i) {
                float t;
                int face = b.order[i];
                if (st) st->tris++;
                if (!tri_test(s.tris[face], o, d, &t)) continue;
                if (t > kEps && (t < leaf.t || (canonical && t == leaf.t && face < leaf.face))) leaf = Hit{t, face};
            }
            if (leaf.t < best.t || (canonical && leaf.face >= 0 && leaf.t == best.t && leaf.face < best.face)) best = leaf;
        } else {                                            // DeviceBVH.cuh:149-167
            bool hl = node.lc >= 0 && ref_hit_aabb(b.nodes[node.lc], o, d, inv);
            bool hr = node.rc >= 0 && ref_hit_aabb(b.nodes[node.rc], o, d, inv);
            if (st) st->boxes += 2;
            if (hl && hr) { stack.push_back(node.lc); stack.push_back(node.rc); }
            else if (hl) stack.push_back(node.lc);
            else if (hr) stack.push_back(node.rc);
            max_stack = std::max<uint64_t>(max_stack, stack.size());
        }
    }
    if (st) { st->rays++; st->max_stack = std::max(st->max_stack, max_stack); }
    return best;
}

// =============================================================================================
// New builder, CPU statement.  DESIGN.md §"BVH build" is the specification; the CUDA builder
// must reproduce nodes/order/last byte for byte.
//   1. triangle boxes (exact min/max), scene box = union of triangle boxes;
//   2. key = 63-bit Morton code of the box centre, 21 bits per axis,
//        q = clamp(int((c - lo) * (2^21 / extent)), 0, 2^21 - 1), extent == 0 -> q = 0;
//      x in the most significant interleave position;
//   3. stable sort by key (ties keep face-id order);
//   4. Karras 2012 binary radix tree over the sorted keys, duplicate keys disambiguated by index;
//   5. a subtree with <= thresh_n triangles is a leaf (the reference's rule, BVH.h:57);
//   6. kept nodes are numbered by the rank of their radix-tree index; node 0 is the root.
// =============================================================================================
static inline uint64_t expand21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

static inline int quant21(float c, float lo, float scale) {
    float f = (c - lo) * scale;
    int q = (int)f;                 // f >= 0 always; NaN cannot occur (scale = 0 when extent == 0)
    if (q > 0x1fffff) q = 0x1fffff;
    if (q < 0) q = 0;
    return q;
}

namespace {
struct Radix {
    const std::vector<uint64_t>& key;
    int n;
    int delta(int i, int j) const {
        if (j < 0 || j >= n) return -1;
        uint64_t a = key[i], b = key[j];
        if (a == b) return 64 + __builtin_clz((unsigned)(i ^ j));
        return __builtin_clzll(a ^ b);
    }
};
}  // namespace

void build_new_bvh(const Scene& s, unsigned thresh_n, int builder, NewBVH& out) {
    const int n = (int)s.tris.size();
    out.nodes.clear(); out.order.clear(); out.last.clear();
    out.builder = builder;
    if (thresh_n < 1) thresh_n = 1;
    out.lo = V3{FLT_MAX, FLT_MAX, FLT_MAX};
    out.hi = V3{-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (n == 0) return;
    for (const Tri& t : s.tris) { out.lo = vmin(out.lo, t.lo); out.hi = vmax(out.hi, t.hi); }
    V3 ext = out.hi - out.lo;
    V3 scale{ext.x > 0 ? 2097152.0f / ext.x : 0.0f, ext.y > 0 ? 2097152.0f / ext.y : 0.0f,
             ext.z > 0 ? 2097152.0f / ext.z : 0.0f};
    std::vector<uint64_t> key(n);
    for (int i = 0; i < n; ++i) {
        const Tri& t = s.tris[i];
        V3 c = (t.lo + t.hi) * 0.5f;
        uint64_t qx = quant21(c.x, out.lo.x, scale.x), qy = quant21(c.y, out.lo.y, scale.y),
                 qz = quant21(c.z, out.lo.z, scale.z);
        key[i] = (expand21(qx) << 2) | (expand21(qy) << 1) | expand21(qz);
    }
    out.order.resize(n);
    std::iota(out.order.begin(), out.order.end(), 0);
    std::stable_sort(out.order.begin(), out.order.end(), [&](int a, int b) { return key[a] < key[b]; });
    std::vector<uint64_t> skey(n);
    for (int i = 0; i < n; ++i) skey[i] = key[out.order[i]];
    out.last.assign(n, 0);

    auto tri_box = [&](int slot, V3& lo, V3& hi) { lo = s.tris[out.order[slot]].lo; hi = s.tris[out.order[slot]].hi; };
    auto set_child = [&](PairNode& pn, int which, int ref, int cnt, V3 lo, V3 hi) {
        if (which == 0) { pn.c0 = ref; pn.n0 = cnt; pn.c0lox = lo.x; pn.c0hix = hi.x; pn.c0loy = lo.y; pn.c0hiy = hi.y; pn.c0loz = lo.z; pn.c0hiz = hi.z; }
        else { pn.c1 = ref; pn.n1 = cnt; pn.c1lox = lo.x; pn.c1hix = hi.x; pn.c1loy = lo.y; pn.c1hiy = hi.y; pn.c1loz = lo.z; pn.c1hiz = hi.z; }
    };
    if ((unsigned)n <= thresh_n) {                         // whole scene is one leaf
        PairNode pn{};
        set_child(pn, 0, ~0, n, out.lo, out.hi);
        set_child(pn, 1, kEmptyChild, 0, V3{FLT_MAX, FLT_MAX, FLT_MAX}, V3{-FLT_MAX, -FLT_MAX, -FLT_MAX});
        out.nodes.push_back(pn);
        out.last[n - 1] = 1;
        return;
    }
    // Karras radix tree: internal nodes 0..n-2
    Radix rx{skey, n};
    std::vector<int> left(n - 1), right(n - 1), first(n - 1), lastl(n - 1);
    for (int i = 0; i < n - 1; ++i) {
        int d = (rx.delta(i, i + 1) - rx.delta(i, i - 1)) >= 0 ? 1 : -1;
        int dmin = rx.delta(i, i - d);
        int lmax = 2;
        while (rx.delta(i, i + lmax * d) > dmin) lmax *= 2;
        int l = 0;
        for (int t = lmax / 2; t >= 1; t /= 2)
            if (rx.delta(i, i + (l + t) * d) > dmin) l += t;
        int j = i + l * d;
        int dnode = rx.delta(i, j);
        int sp = 0;
        int t = l;
        do {
            t = (t + 1) >> 1;
            if (rx.delta(i, i + (sp + t) * d) > dnode) sp += t;
        } while (t > 1);
        int gamma = i + sp * d + std::min(d, 0);
        int lo_i = std::min(i, j), hi_i = std::max(i, j);
        first[i] = lo_i; lastl[i] = hi_i;
        left[i] = (lo_i == gamma) ? ~gamma : gamma;            // ~slot = single-triangle leaf
        right[i] = (hi_i == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    }
    // boxes of radix nodes, bottom-up by recursion on an explicit stack (post-order)
    std::vector<V3> blo(n - 1), bhi(n - 1);
    {
        std::vector<std::pair<int, int>> st;   // (node, state)
        st.emplace_back(0, 0);
        while (!st.empty()) {
            auto [nd, state] = st.back();
            if (state == 0) {
                st.back().second = 1;
                if (left[nd] >= 0) st.emplace_back(left[nd], 0);
                if (right[nd] >= 0) st.emplace_back(right[nd], 0);
            } else {
                st.pop_back();
                V3 llo, lhi, rlo, rhi;
                if (left[nd] >= 0) { llo = blo[left[nd]]; lhi = bhi[left[nd]]; } else tri_box(~left[nd], llo, lhi);
                if (right[nd] >= 0) { rlo = blo[right[nd]]; rhi = bhi[right[nd]]; } else tri_box(~right[nd], rlo, rhi);
                blo[nd] = vmin(llo, rlo);
                bhi[nd] = vmax(lhi, rhi);
            }
        }
    }
    // kept nodes: more than thresh_n triangles; rank by radix index
    std::vector<int> rank(n - 1, -1);
    int n_kept = 0;
    for (int i = 0; i < n - 1; ++i)
        if ((unsigned)(lastl[i] - first[i] + 1) > thresh_n) rank[i] = n_kept++;
    out.nodes.resize(n_kept);
    for (int i = 0; i < n - 1; ++i) {
        if (rank[i] < 0) continue;
        PairNode pn{};
        for (int w = 0; w < 2; ++w) {
            int c = w == 0 ? left[i] : right[i];
            V3 lo, hi;
            if (c < 0) {                                     // single triangle
                tri_box(~c, lo, hi);
                set_child(pn, w, ~(~c), 1, lo, hi);
                out.last[~c] = 1;
            } else {
                int cnt = lastl[c] - first[c] + 1;
                lo = blo[c]; hi = bhi[c];
                if (rank[c] >= 0) set_child(pn, w, rank[c], cnt, lo, hi);
                else { set_child(pn, w, ~first[c], cnt, lo, hi); out.last[lastl[c]] = 1; }
            }
        }
        out.nodes[rank[i]] = pn;
    }
}

// =============================================================================================
// New traversal rule (DESIGN.md §"Traversal rule"): conservative slabs, near child first,
// far child pushed, leaves postponed on the stack; t-culling against 1.0001 * current limit.
// =============================================================================================
static inline bool slab(float lox, float hix, float loy, float hiy, float loz, float hiz, V3 o, V3 inv,
                        float limit, float* enter) {
    float tx0 = (lox - o.x) * inv.x, tx1 = (hix - o.x) * inv.x;
    float ty0 = (loy - o.y) * inv.y, ty1 = (hiy - o.y) * inv.y;
    float tz0 = (loz - o.z) * inv.z, tz1 = (hiz - o.z) * inv.z;
    float tmin = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
    float tmax = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), limit));
    *enter = tmin;
    return tmin <= tmax * 1.0000004f;
}

Hit new_intersect(const Scene& s, const NewBVH& b, const Ray& r, int mode, TraceStats* st) {
    Hit best{FLT_MAX, -1};
    if (st) st->rays++;
    if (b.nodes.empty()) return best;
    V3 o = r.o, d = r.d;
    V3 inv{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    float tlimit = mode == 0 ? FLT_MAX : r.tmax;
    int stack[128];
    int sp = 0;
    int cur = 0;
    uint64_t max_sp = 0;
    for (;;) {
        if (cur >= 0) {
            if (cur == kEmptyChild) { if (sp == 0) break; cur = stack[--sp]; continue; }
            const PairNode& n = b.nodes[cur];
            if (st) { st->inner++; st->boxes += 2; }
            float lim = tlimit * 1.0001f;
            float e0, e1;
            bool h0 = slab(n.c0lox, n.c0hix, n.c0loy, n.c0hiy, n.c0loz, n.c0hiz, o, inv, lim, &e0);
            bool h1 = slab(n.c1lox, n.c1hix, n.c1loy, n.c1hiy, n.c1loz, n.c1hiz, o, inv, lim, &e1);
            if (h0 && h1) {
                int nearc = n.c0, farc = n.c1;
                if (e1 < e0) { nearc = n.c1; farc = n.c0; }
                stack[sp++] = farc;
                if ((uint64_t)sp > max_sp) max_sp = sp;
                cur = nearc;
            } else if (h0) cur = n.c0;
            else if (h1) cur = n.c1;
            else { if (sp == 0) break; cur = stack[--sp]; }
        } else {
            int slot = ~cur;
            for (;; ++slot) {
                int face = b.order[slot];
                float t;
                if (st) st->tris++;
                if (tri_test(s.tris[face], o, d, &t) && t > kEps) {
                    if (mode == 0) {
                        if (t < best.t || (t == best.t && face < best.face)) { best = Hit{t, face}; tlimit = t; }
                    } else if (r.tmax - t > kEps) {
                        if (st) st->max_stack = std::max(st->max_stack, max_sp);
                        return Hit{t, face};
                    }
                }
                if (b.last[slot]) break;
            }
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    if (st) st->max_stack = std::max(st->max_stack, max_sp);
    return best;
}

Hit brute_intersect(const Scene& s, const Ray& r, int mode) {
    Hit best{FLT_MAX, -1};
    for (int f = 0; f < (int)s.tris.size(); ++f) {
        float t;
        if (!tri_test(s.tris[f], r.o, r.d, &t) || !(t > kEps)) continue;
        if (mode == 0) { if (t < best.t) best = Hit{t, f}; }     // ascending f: ties keep the lower id
        else if (r.tmax - t > kEps) return Hit{t, f};
    }
    return best;
}

}  // namespace orc
