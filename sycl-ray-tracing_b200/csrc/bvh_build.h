// Host-side builder for the re-laid-out FlattenedBVH (layout documented in include/b200rt.h).
// Replaces BVH::BVH / OctreeNode::insert / compute_volume / flatten of the reference
// (source/bvh.cpp:19-60, include/bvh.h:55-125, :211-250): only the closest-hit RESULT is contractual, not the tree
// shape, so this is a binned-SAH binary BVH whose nodes carry the reference's 7-plane Kay-Kajiya slab volumes
// (include/bounding_volume.h:9-127, plane normals source/bvh.cpp:8-16).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/b200rt.h"

namespace b200rt {

constexpr int kMaxTraversalDepth = 60;   // the device stack holds 64 entries

struct alignas(64) AxisNode
{
    float l_lo[3], l_hi[3], r_lo[3], r_hi[3];
    int l_ref, r_ref, l_count, r_count;
};
static_assert(sizeof(AxisNode) == 64, "axis record is one 64-byte line");

struct alignas(64) DiagNode
{
    float l_near[4], l_far[4], r_near[4], r_far[4];
};
static_assert(sizeof(DiagNode) == 64, "diag record is one 64-byte line");

struct alignas(16) LeafTriangle
{
    float a[3]; int prim;      // vertex a + original triangle index
    float e1[3]; float pad1;   // b - a   (same float subtraction the reference does per test, triangle.h:21)
    float e2[3]; float pad2;   // c - a   (triangle.h:22)
};
static_assert(sizeof(LeafTriangle) == 48, "three float4 per triangle");

// node of the 8-ary quantised BVH (device twin and the plane decoding: csrc/pt_device.cuh, "wide BVH")
struct alignas(16) WideNode
{
    float p[3];                 // frame origin
    uint8_t e[3];               // per-axis cell size 2^(e - 127)
    uint8_t imask;              // bit s set: slot s holds an inner child
    uint32_t child_base;        // node index of the first inner child (inner children are contiguous, in slot order)
    uint32_t tri_base;          // slot of the first triangle of this node's leaf children (contiguous, in slot order)
    uint32_t valid24;           // bit 3s + i: slot s is a leaf child with a triangle i; triangles are stored compactly in that bit order
    uint32_t spare;
    uint8_t lox[8], loy[8], loz[8], hix[8], hiy[8], hiz[8];   // quantised child boxes, one byte per plane
};
static_assert(sizeof(WideNode) == 80, "five float4 per wide node");

// plane byte -> grid value: the float with bits 0x43000000 | q << 16 (128 + q below 128, 2q from 128 on)
inline int wide_grid_value(int q) { return q < 128 ? 128 + q : 2 * q; }
constexpr int kWideMaxLeaf = 3;          // a leaf child's triangle count is unary-coded in 3 bits

struct FlatBVH
{
    std::vector<WideNode> wide;
    std::vector<AxisNode> axis;
    std::vector<DiagNode> diag;
    std::vector<LeafTriangle> tris;
    b200rt_bvh_info info{};
};

// the 4 diagonal plane normals WITHOUT the sqrt(3)/3 factor: distances along (+-1, +-1, 1) give the same slab
// intersection parameters t as the reference's normalised planes (both numerator and denominator scale alike)
inline float diag_dist(int k, float x, float y, float z)
{
    switch (k)
    {
    case 0: return (x + y) + z;
    case 1: return (-x + y) + z;
    case 2: return (-x - y) + z;
    default: return (x - y) + z;
    }
}

void build_flat_bvh(const float* tri9, int n_tri, const b200rt_bvh_options& opts, FlatBVH& out);
int check_flat_bvh(const FlatBVH& bvh, const float* tri9, int n_tri);
double sah_cost_of(const FlatBVH& bvh);
// triangles kept out of the tree become leaf children of extra top-level nodes in front of both layouts' roots (bvh_build.cpp)
void append_top_level(FlatBVH& bvh, const float* tri9, const std::vector<int>& ids, float abs_pad);
// device builder (bvh_build_gpu.cu): 0 = ok, 1 = CUDA error, 2 = input outside what it handles (err says why)
int build_flat_bvh_device(const float* tri9_host, int n_tri, int device, FlatBVH& out, std::string& err);

} // namespace b200rt
