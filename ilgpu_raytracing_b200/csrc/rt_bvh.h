// rt_bvh.h — host-side flattening of the reference scene arrays into the wide-BVH device layout.
#pragma once
#include <string>
#include <vector>

#include "rt_core.h"
#include "rt_traverse.h"

namespace rtx {

struct HostBvhStats {
    int64_t nPrims = 0, nTris = 0, nSpheres = 0, nWideNodes = 0;
    int maxDepth = 0;
    int depthBounded = 0;            // 1: the SAH tree was deeper than the traversal stack and the depth-bounded rebuild was taken
    float maxInstanceScale = 1.0f;   // max(1, uniformScale of every instance reached from the TLAS)
    float sceneLo[3] = {0, 0, 0}, sceneHi[3] = {0, 0, 0};
};

struct HostBvh {
    std::vector<WideNode> nodes;   // nodes[0] is the root; breadth-first, so every level is one contiguous index range
    std::vector<PrimRec> prims;    // leaf order
    std::vector<int> levelStart;   // levelStart[l] .. levelStart[l + 1] = the nodes of depth l + 1 (for the device-side refit)
    std::vector<double> instBoxXf; // per instance: the 3x4 object-to-world transform its primitive boxes were built with (identity for identity instances)
    HostBvhStats stats;
};

// Where a primsOnly build leaves its output: the caller's (page-locked) staging arrays, written in place by all host threads -
// prims[i] and the padded world box of primitive i as two float4 (lo, hi), in the reference's visiting order.  reserve(n) is
// called once, when the primitive count is known, and must make both arrays hold n entries (false = out of memory).
struct PrimSink {
    PrimRec* prims = nullptr;
    float4* boxes = nullptr;
    bool (*reserve)(PrimSink* self, size_t n) = nullptr;
    void (*recordsDone)(PrimSink* self, size_t n) = nullptr;   // optional: prims[0..n) are final (the boxes are still being padded): the caller may start copying them
    void* user = nullptr;
};

// Validates the reference arrays (every index the device code will follow), derives the reference's
// visiting order and builds the wide BVH.  Returns false with a message on malformed input.
// sink != nullptr ("primsOnly"): stop after the primitive stage (validation, visiting-order ranks, records, padded boxes, scene
// bounds and stats.nPrims) and leave its output in the sink: the tree is then built on the device (rt_build.h).
// maxDepth: deepest wide tree the caller's traversal stack takes (0 = RT_STACK_ENTRIES - 2); a deeper SAH tree is rebuilt
// depth-bounded (median splits, three binary levels per wide node) - the commit never fails for depth.
bool build_wide_bvh(const RtSceneDesc& desc, HostBvh& out, std::string& err, PrimSink* sink = nullptr, int maxDepth = 0);

}   // namespace rtx
