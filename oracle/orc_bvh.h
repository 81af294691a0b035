// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h). Never linked into the product.
//
// orc_bvh.h — (1) the reference's host BVH and its device traversal rule, restated for the CPU
// with face ids carried through the sort; (2) the CPU build of the NEW builder (the algorithm the
// CUDA builder in cudaraytracing_b200/csrc implements) and the new traversal rule, which the GPU
// must match bit-exactly (node bytes, triangle order, hit ids, visit counts).
#pragma once
#include <vector>
#include "orc_scene.h"

namespace orc {

// ---------------------------------------------------------------- reference BVH (BVH.h:9-84)
struct RefNode {              // byte layout of BVHNode / DeviceBVHNode (40 B)
    int lc, rc;
    unsigned n;
    int it;
    float AA[3], BB[3];
};
struct RefBVH {
    std::vector<RefNode> nodes;   // post-order, root last (BVH.h:57-58,81-82)
    std::vector<int> order;       // order[k] = face id of the k-th triangle after the in-place sorts
    int root = -1;
};
void build_ref_bvh(const Scene& s, unsigned thresh_n, RefBVH& out);

struct Ray { V3 o, d; float tmax; };
struct Hit { float t; int face; };           // face = -1 on a miss, t = FLT_MAX
struct TraceStats { uint64_t inner = 0, boxes = 0, tris = 0, max_stack = 0, rays = 0; };

// Canonical triangle test: Moeller-Trumbore of DeviceTriangle.cuh:39-65
// with the strict-inside rule; returns t or a negative value when outside.
bool tri_test(const Tri& tr, V3 o, V3 d, float* t_out);

// DeviceBVH::intersect (DeviceBVH.cuh:128-170) + hit_AABB (:87-126) + DeviceBVHNode::hit (:31-43).
// canonical_ties=false keeps the reference's "first found, strict <" rule; true breaks t ties by
// the lower face id (the BVH-independent rule the new traversal uses).
Hit ref_intersect(const Scene& s, const RefBVH& b, V3 o, V3 d, bool canonical_ties, TraceStats* st);

// ---------------------------------------------------------------- new BVH
struct PairNode {             // 64 B, four 16-byte words
    float c0lox, c0hix, c0loy, c0hiy;
    float c1lox, c1hix, c1loy, c1hiy;
    float c0loz, c0hiz, c1loz, c1hiz;
    int c0, c1;               // >= 0: node index; < 0: leaf, ~first triangle slot; see kEmpty
    int n0, n1;               // triangles under each child (informational; leaves are sentinel-terminated)
};
static const int kEmptyChild = 0x7fffffff;   // absent child (single-leaf scenes); its box is inverted

struct NewBVH {
    std::vector<PairNode> nodes;      // node 0 is the root
    std::vector<int> order;           // slot -> face id
    std::vector<uint8_t> last;        // slot -> 1 when it is the last triangle of its leaf
    V3 lo, hi;                        // scene bounds
    int builder = 0;
};
// same numbering as the product's crt_builder: bit 0 = 8-wide node layout, bit 1 = PLOC topology
enum Builder { BUILDER_LBVH = 0, BUILDER_LBVH8 = 1, BUILDER_PLOC = 2, BUILDER_PLOC8 = 3 };
void build_new_bvh(const Scene& s, unsigned thresh_n, int builder, NewBVH& out);

// mode 0: closest hit (t > 1e-5, ties -> lower face id); mode 1: any hit with
// (t > 1e-5 && tmax - t > 1e-5), returns face of the first blocker found (or -1).
Hit new_intersect(const Scene& s, const NewBVH& b, const Ray& r, int mode, TraceStats* st);

Hit brute_intersect(const Scene& s, const Ray& r, int mode);

// ---------------------------------------------------------------- 8-wide compressed BVH (DESIGN.md "Wide nodes")
// 80-byte node = five 16-byte words:
//   w0: origin p (3 floats), {ex, ey, ez, imask} bytes          scale_k = float with exponent field e_k
//   w1: child_base, tri_base, meta[0..3], meta[4..7]            meta: 0 empty, 0x80 node, 1 + offset leaf
//   w2: qlo_x[0..7], qlo_y[0..7]   w3: qlo_z[0..7], qhi_x[0..7]   w4: qhi_y[0..7], qhi_z[0..7]
// child box k = p + q * scale (per axis), internal child in slot s = node child_base + popcount(imask below s),
// leaf child = triangle slots tri_base + offset .. up to the terminator in `last`.
struct Wide8Node { uint32_t w[20]; };
struct Wide8BVH {
    std::vector<Wide8Node> nodes;     // breadth-first, node 0 is the root
    std::vector<int> order;           // slot -> face id (leaf triangles of one node are contiguous)
    std::vector<uint8_t> last;
    V3 lo, hi;
};
static const unsigned kWideMaxLeaf = 15;      // a leaf holds at most this many triangles (7-bit offsets)
void build_wide8_bvh(const Scene& s, unsigned thresh_n, Wide8BVH& out, int builder = BUILDER_LBVH8);
Hit wide8_intersect(const Scene& s, const Wide8BVH& b, const Ray& r, int mode, TraceStats* st);

}  // namespace orc
