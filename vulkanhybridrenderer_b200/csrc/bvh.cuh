// bvh.cuh — acceleration-structure layouts shared by the builder (bvh_build.cu) and the ray kernels (trace_kernels.cu).
//
// Replaces the opaque BLAS/TLAS the reference builds with vkCmdBuildAccelerationStructuresKHR
// (/root/reference/src/rendering_backend/resource_manager.cpp:593-801). Geometry semantics follow that code: one
// world-space triangle soup, all opaque, two-sided, geometry index = flat primitive index, primitive id = triangle
// index inside the primitive.
#pragma once
#include <stdint.h>

#include "vhr_common.cuh"

namespace vhr {

// 8-wide node with 8-bit quantised child boxes, 80 bytes = five 128-bit loads.
//   child box i = origin + q{lo,hi}[axis][i] * 2^(e[axis]-127)   (conservative: lo rounded down, hi rounded up)
//   child_base              bits 0-29: index of the first internal child's node; bits 30-31: the axis the slots are sorted along
//                           (box centres ascending; 3 = not sorted). A ray whose direction is negative on that axis visits the
//                           hit children from the highest slot down, so the near side comes first either way.
//   meta[i] == 0            empty slot (its box is inverted: qlo = 255, qhi = 0)
//   slot i internal          bit i of imask (the traversal tells internal from leaf by imask only); meta[i] = 0x80 | ordinal,
//                           node index = child_base + ordinal
//   otherwise               leaf child: triangles [tri_base + (meta[i] & 31), + (meta[i] >> 5)), 1..3 triangles (4 with VHR_MAX_LEAF_TRIS=4)
struct __align__(16) WideNode {
    float origin[3];
    uint8_t e[3];
    uint8_t imask;          // bit i: slot i holds an internal child
    uint32_t child_base;
    uint32_t tri_base;
    uint8_t meta[8];
    uint8_t qlo[3][8];
    uint8_t qhi[3][8];
#ifdef VHR_NODE_PAD_BYTES      // study: 48 pads a node to one 128-byte cache line (measured: see DESIGN.md section 7)
    uint8_t pad[VHR_NODE_PAD_BYTES];
#endif
};
#ifndef VHR_NODE_PAD_BYTES
static_assert(sizeof(WideNode) == 80, "WideNode must be 80 bytes");
#endif

constexpr int kMaxLeafTris = 3;

// Entries of the per-ray traversal stack. An entry is the not-yet-visited rest of one node's hit children, so a ray holds at most
// one entry per level of the wide tree: the builder refuses a tree deeper than this instead of letting a ray drop children.
constexpr int kStackSize = 40;

// Triangle record: three float4, world space. v0.w = geometry index bits, v1.w = primitive id bits, v2.w unused.
struct TriRef {
    float4 v0, v1, v2;
};

}  // namespace vhr
