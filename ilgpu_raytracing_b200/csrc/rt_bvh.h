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
    std::vector<float> primBoxes;  // primsOnly builds: padded world box of prims[i], 6 floats (lo, hi), in the reference's visiting order
    HostBvhStats stats;
};

// Validates the reference arrays (every index the device code will follow), derives the reference's
// visiting order and builds the wide BVH.  Returns false with a message on malformed input.
// primsOnly: stop after the primitive stage (validation, visiting-order ranks, records, padded boxes, scene bounds in stats):
// the tree is then built on the device (rt_build.h).
// maxDepth: deepest wide tree the caller's traversal stack takes (0 = RT_STACK_ENTRIES - 2); a deeper SAH tree is rebuilt
// depth-bounded (median splits, three binary levels per wide node) - the commit never fails for depth.
bool build_wide_bvh(const RtSceneDesc& desc, HostBvh& out, std::string& err, bool primsOnly = false, int maxDepth = 0);

}   // namespace rtx
