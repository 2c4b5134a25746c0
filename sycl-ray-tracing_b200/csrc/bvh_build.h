// Host-side builder for the re-laid-out FlattenedBVH (layout documented in include/b200rt.h).
// Replaces BVH::BVH / OctreeNode::insert / compute_volume / flatten of the reference
// (source/bvh.cpp:19-60, include/bvh.h:55-125, :211-250): only the closest-hit RESULT is contractual, not the tree
// shape, so this is a binned-SAH binary BVH whose nodes carry the reference's 7-plane Kay-Kajiya slab volumes
// (include/bounding_volume.h:9-127, plane normals source/bvh.cpp:8-16).
#pragma once
#include <cstdint>
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

struct FlatBVH
{
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

} // namespace b200rt
