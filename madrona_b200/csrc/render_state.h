// render_state.h -- layouts shared by the NVRTC-side rendering API
// (device/madrona/render/ecs.hpp) and the ahead-of-time batch ray caster
// (kernels_render.cu).  SURVEY.md 8 rows a13-a15.
#pragma once

#include "mb2_state.h"

namespace mb2 {

struct RVec3 { float x, y, z; };
struct RQuat { float w, x, y, z; };

// One triangle mesh per object ID (the role of render::MeshBVH in the
// reference, include/madrona/mesh_bvh.hpp:20-146; here a flat triangle range --
// fixture meshes are a dozen triangles, the BLAS is a linear scan).
struct MeshDesc {
    u32 firstTriangle;
    u32 numTriangles;
    float aabbMin[3];
    float aabbMax[3];
};

// == render::InstanceData reduced to what the ray caster reads
struct RenderInstance {
    RVec3 position;
    RQuat rotation;
    RVec3 scale;
    i32 objectID;
    u32 color;            // 0xRRGGBB (ColorOverride), white if none
    float aabbMin[3];     // world-space box of the instance (TLBVHNode in the reference)
    float aabbMax[3];
};

// == render::PerspectiveCameraData (include/madrona/render/ecs.hpp:38-46)
struct RenderView {
    RVec3 position;
    RQuat rotation;       // inverse of the viewing entity's rotation
    float xScale;
    float yScale;
    float zNear;
    i32 worldIDX;
    i32 outputRow;        // row of the view's RaycastOutputArchetype entity
};

enum RenderCol : int { RCPosition = 0, RCRotation, RCScale, RCObjectID, RCRenderable, RCCount };
constexpr int kMaxRenderArchetypes = 16;

struct RenderArchetype {
    u32 archetype;
    i32 cols[RCCount];
    i32 colorCol;          // ColorOverride column or -1
};

struct ViewArchetype {
    u32 archetype;
    i32 posCol, rotCol, camCol;
};

struct RenderState {
    // ---- host config (before registerTypes)
    u32 enabled;
    u32 resolution;
    u32 rgbd;              // 1 = RGB + depth, 0 = depth only
    float nearPlane, farPlane;
    const MeshDesc *meshes;
    u32 numMeshes;
    const float *vertices;          // xyz per vertex
    const u32 *indices;             // 3 per triangle
    u32 numTriangles;

    // ---- written by the device-side RenderingSystem::registerTypes
    u32 registered;
    u32 cidRenderable, cidRenderCamera, cidColorOverride;
    u32 cidPosition, cidRotation, cidScale, cidObjectID;
    u32 outputArchetype;   // RaycastOutputArchetype
    u32 cidRGB, cidDepth;

    // ---- filled by the host after registerTypes
    u32 numRenderArchetypes;
    RenderArchetype renderables[kMaxRenderArchetypes];
    u32 numViewArchetypes;
    ViewArchetype viewers[kMaxRenderArchetypes];
    i32 rgbCol, depthCol;

    RenderInstance *instances;      // [numWorlds][maxInstancesPerWorld]
    i32 *instanceCounts;            // [numWorlds]
    i32 maxInstancesPerWorld;
    RenderView *views;              // [capacity of the output archetype]
    i32 maxViews;
};

}
