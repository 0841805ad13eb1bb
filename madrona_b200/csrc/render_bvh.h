// render_bvh.h -- acceleration-structure formats of the batch ray caster, shared
// by the host-side BLAS builder (mesh_bvh.cpp), the ray-cast kernels
// (kernels_render.cu) and, through include/madrona_b200.h, the C ABI.
//
// Layouts are the reference's, field for field, so a CudaBatchRenderConfig
// filled the reference way (render::MeshBVHData / render::MaterialData from
// render::AssetProcessor::makeBVHData) can be handed over unchanged:
//   QBVHNode   == madrona::BVHNodeQuantized<uint32_t, 4> (include/madrona/mesh_bvh.hpp:20-146)
//   MeshBVH    == madrona::MeshBVH data members           (mesh_bvh.hpp:293-307)
//   MeshBVHData / MaterialData == include/madrona/render/cuda_batch_render_assets.hpp
#pragma once

#include <cstdint>
#include <cmath>

namespace mb2 {

constexpr int kBVHWidth = 4;            // MADRONA_BVH_WIDTH
constexpr int kBLASLeafWidth = 2;       // MADRONA_BLAS_LEAF_WIDTH: triangles per BLAS leaf

// 4-wide node with child boxes quantised to 8 bits on a per-node power-of-two
// grid: child box = minPoint + 2^exp * q.  Children: 0xFFFFFFFF = none, bit 31
// set = leaf (TLAS: instance index; BLAS: index of the leaf's first triangle in
// the de-indexed vertex array, triSize triangles long), else node index.
struct QBVHNode {
    float minPoint[3];
    int8_t expX, expY, expZ;
    uint8_t numChildren;
    uint8_t triSize[kBVHWidth];
    uint8_t qMinX[kBVHWidth], qMinY[kBVHWidth], qMinZ[kBVHWidth];
    uint8_t qMaxX[kBVHWidth], qMaxY[kBVHWidth], qMaxZ[kBVHWidth];
    uint32_t childrenIdx[kBVHWidth];
};
static_assert(sizeof(QBVHNode) == 60, "QBVHNode layout");

struct BVHVertex {          // MeshBVH::BVHVertex
    float pos[3];
    float uv[2];
};
static_assert(sizeof(BVHVertex) == 20, "BVHVertex layout");

struct LeafMaterial {       // MeshBVH::LeafMaterial
    int32_t matIDX;
};

struct RenderMaterial {     // madrona::Material
    float color[4];
    int32_t textureIdx;
    float roughness;
    float metalness;
};
static_assert(sizeof(RenderMaterial) == 28, "Material layout");

struct MeshBVH {
    QBVHNode *nodes;
    LeafMaterial *leafMats;
    BVHVertex *vertices;
    float rootAABBMin[3];
    float rootAABBMax[3];
    uint32_t numNodes;
    uint32_t numLeaves;         // triangles (a leaf index addresses a triangle)
    uint32_t numVerts;
    int32_t materialIDX;
    uint32_t magic;
};
static_assert(sizeof(MeshBVH) == 72, "MeshBVH layout");

struct MeshBVHData {
    QBVHNode *nodes;
    uint64_t numNodes;
    LeafMaterial *leafMaterial;
    uint64_t numLeaves;
    BVHVertex *vertices;
    uint64_t numVerts;
    MeshBVH *meshBVHs;
    uint64_t numBVHs;
};

struct MaterialData {
    unsigned long long *textures;      // cudaTextureObject_t *
    uint32_t numTextureBuffers;
    void **textureBuffers;             // cudaArray_t *
    RenderMaterial *materials;
};

// Quantise up to 4 child boxes into a node (same arithmetic as
// BVHNodeQuantized::construct, mesh_bvh.hpp:66-128): exp = ceil(log2(extent /
// 255)), mins floored, maxes ceiled, so the quantised box always contains the
// child box.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void quantizeNode(QBVHNode &node, int num_children, const float (*cmin)[3], const float (*cmax)[3])
{
    float lo[3] = { cmin[0][0], cmin[0][1], cmin[0][2] };
    float hi[3] = { cmax[0][0], cmax[0][1], cmax[0][2] };
    for (int i = 1; i < num_children; i++) {
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(lo[a], cmin[i][a]);
            hi[a] = fmaxf(hi[a], cmax[i][a]);
        }
    }
    int8_t exps[3];
    float inv_scale[3];
    for (int a = 0; a < 3; a++) {
        const float extent = hi[a] - lo[a];
        int e = extent > 0.f ? (int)ceilf(log2f(extent / 255.f)) : -126;
        if (e < -126) e = -126;
        if (e > 126) e = 126;
        // the grid must reach the far side: 255 * 2^e >= extent (log2f may round down)
        while (e < 126 && ldexpf(255.f, e) < extent) e++;
        exps[a] = (int8_t)e;
        inv_scale[a] = ldexpf(1.f, -e);
    }
    node.minPoint[0] = lo[0]; node.minPoint[1] = lo[1]; node.minPoint[2] = lo[2];
    node.expX = exps[0]; node.expY = exps[1]; node.expZ = exps[2];
    node.numChildren = (uint8_t)num_children;
    for (int i = 0; i < kBVHWidth; i++) {
        if (i < num_children) {
            float qlo[3], qhi[3];
            for (int a = 0; a < 3; a++) {
                qlo[a] = floorf((cmin[i][a] - lo[a]) * inv_scale[a]);
                qhi[a] = ceilf((cmax[i][a] - lo[a]) * inv_scale[a]);
                qlo[a] = fminf(fmaxf(qlo[a], 0.f), 255.f);
                qhi[a] = fminf(fmaxf(qhi[a], 0.f), 255.f);
            }
            node.qMinX[i] = (uint8_t)qlo[0]; node.qMinY[i] = (uint8_t)qlo[1]; node.qMinZ[i] = (uint8_t)qlo[2];
            node.qMaxX[i] = (uint8_t)qhi[0]; node.qMaxY[i] = (uint8_t)qhi[1]; node.qMaxZ[i] = (uint8_t)qhi[2];
        } else {
            node.qMinX[i] = node.qMinY[i] = node.qMinZ[i] = 0;
            node.qMaxX[i] = node.qMaxY[i] = node.qMaxZ[i] = 0;
            node.childrenIdx[i] = 0xFFFFFFFFu;
            node.triSize[i] = 0;
        }
    }
}

}
