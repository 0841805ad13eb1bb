// kernels_render.cu -- hot system 3: batch ray-cast renderer (SURVEY.md 8 rows
// a13-a15).  Per step (render-prepare node, inside the step graph): gather
// every world's renderable instances (transform, object, world-space box) and
// every view's camera into flat arrays.  Render graph: one thread per pixel,
// a block covers 256 consecutive pixels of ONE view, so all its rays walk the
// same world's instance list (staged in shared memory) in lock step.
//
// Image formation follows the reference's CUDA ray tracer
// (src/mw/device/bvh_raycast.cpp): ray generation :58-88, object-space ray
// and t rescaling :620-645 / :744-751, watertight ray-triangle test :317-448
// (explicit fmaf kept), depth / RGBA8 output :820-838, 940-1029, unlit colour
// = max(0.2, 0) * colour with zero lights :848-938.  The acceleration structure
// is different by design: the reference sorts render entities by Morton code
// three times per step and builds a per-world LBVH + 4-wide quantised QBVH
// (src/mw/device/bvh.cpp); with tens of instances per world a linear,
// branch-coherent scan of world boxes beats a divergent tree walk, and a hit's
// depth does not depend on the structure that found it.  (Meshes here are flat
// triangle ranges; a mesh BLAS builder is SURVEY 8f N4.)
//
// Parity: the reference can only render on its GPU backend (CPU backend forces
// raycast off, src/render/ecs_system.cpp:684-689), so there is no reference
// output to compare with here -- "parity unpinned"; tests compare against the
// numpy restatement oracle/restate_render.py of the same formulas.
#include "physics_host.hpp"
#include "render_state.h"

#include <madrona/math.hpp>
#include <cfloat>
#include <algorithm>

namespace mb2 {

using madrona::math::Vector3;
using madrona::math::Quat;
using madrona::math::Diag3x3;
using madrona::math::AABB;
using madrona::math::cross;
using madrona::math::dot;

struct RenderHost {
    RenderState *dRender = nullptr;
    RenderState hRender;
    bool active = false;
};

struct RenderCameraComp {     // == madrona::render::RenderCamera
    u32 outGen; i32 outID;
    float fovScale;
    float zNear;
    Vector3 cameraOffset;
};

constexpr int kStagedInstances = 64;

// ---- render-prepare: instances ---------------------------------------------------------
__global__ void __launch_bounds__(128)
renderGatherInstancesKernel(EngineState *Sp)
{
    pdlSync();
    EngineState &S = *Sp;
    RenderState &R = *S.render;
    const int lane = threadIdx.x & 31;
    const i32 w = (i32)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (w >= (i32)S.numWorlds) return;

    RenderInstance *out = R.instances + (size_t)w * R.maxInstancesPerWorld;
    i32 running = 0;
    for (u32 ai = 0; ai < R.numRenderArchetypes; ai++) {
        const RenderArchetype &ra = R.renderables[ai];
        const TableDesc &t = S.tables[ra.archetype];
        const i32 first = t.worldOffsets[w];
        const i32 count = t.worldCounts[w];
        for (i32 base = 0; base < count; base += 32) {
            const i32 row = first + base + lane;
            bool valid = base + lane < count && ((const i32 *)t.columns[1])[row] == w;
            if (valid) {
                const u64 marker = ((const u64 *)t.columns[ra.cols[RCRenderable]])[row];
                valid = (i32)(u32)(marker >> 32) != -1;     // Renderable{Entity::none()} => hidden
            }
            const u32 keep = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                const i32 at = running + __popc(keep & ((1u << lane) - 1u));
                if (at < R.maxInstancesPerWorld) {
                    RenderInstance inst;
                    const Vector3 p = ((const Vector3 *)t.columns[ra.cols[RCPosition]])[row];
                    const Quat q = ((const Quat *)t.columns[ra.cols[RCRotation]])[row];
                    const Diag3x3 s = ((const Diag3x3 *)t.columns[ra.cols[RCScale]])[row];
                    inst.position = RVec3 { p.x, p.y, p.z };
                    inst.rotation = RQuat { q.w, q.x, q.y, q.z };
                    inst.scale = RVec3 { s.d0, s.d1, s.d2 };
                    inst.objectID = ((const i32 *)t.columns[ra.cols[RCObjectID]])[row];
                    inst.color = ra.colorCol >= 0 ? ((const u32 *)t.columns[ra.colorCol])[row] : 0xFFFFFFu;
                    AABB box { { 0, 0, 0 }, { 0, 0, 0 } };
                    if (inst.objectID >= 0 && (u32)inst.objectID < R.numMeshes) {
                        const MeshDesc &m = R.meshes[inst.objectID];
                        box = AABB { { m.aabbMin[0], m.aabbMin[1], m.aabbMin[2] },
                                     { m.aabbMax[0], m.aabbMax[1], m.aabbMax[2] } }.applyTRS(p, q, s);
                    }
                    inst.aabbMin[0] = box.pMin.x; inst.aabbMin[1] = box.pMin.y; inst.aabbMin[2] = box.pMin.z;
                    inst.aabbMax[0] = box.pMax.x; inst.aabbMax[1] = box.pMax.y; inst.aabbMax[2] = box.pMax.z;
                    out[at] = inst;
                }
            }
            running += __popc(keep);
        }
    }
    if (lane == 0) {
        if (running > R.maxInstancesPerWorld) {
            atomicOr(&S.errorFlags, (u32)ErrPhysicsOverflow);
            running = R.maxInstancesPerWorld;
        }
        R.instanceCounts[w] = running;
    }
}

// ---- render-prepare: views (viewTransformUpdate, ecs_system.cpp:275-314) ----------------
__global__ void __launch_bounds__(256)
renderGatherViewsKernel(EngineState *Sp)
{
    pdlSync();
    EngineState &S = *Sp;
    RenderState &R = *S.render;
    if (blockIdx.y >= R.numViewArchetypes) return;
    const ViewArchetype &va = R.viewers[blockIdx.y];
    const TableDesc &t = S.tables[va.archetype];
    const i32 n = t.numRows;
    const i32 *world_col = (const i32 *)t.columns[1];
    for (i32 row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
        const i32 w = world_col[row];
        if (w < 0) continue;
        const RenderCameraComp cam = ((const RenderCameraComp *)t.columns[va.camCol])[row];
        if (cam.outID < 0 || cam.outID >= S.entityCapacity) continue;
        const EntitySlot slot = S.entitySlots[cam.outID];
        if (slot.gen != cam.outGen || (u32)slot.a != R.outputArchetype) continue;
        const i32 out_row = slot.b;
        if (out_row < 0 || out_row >= R.maxViews) continue;
        const Vector3 p = ((const Vector3 *)t.columns[va.posCol])[row];
        const Quat q = ((const Quat *)t.columns[va.rotCol])[row];
        const Vector3 cam_pos = p + cam.cameraOffset;
        const Quat inv = q.inv();
        RenderView v;
        v.position = RVec3 { cam_pos.x, cam_pos.y, cam_pos.z };
        v.rotation = RQuat { inv.w, inv.x, inv.y, inv.z };
        v.xScale = cam.fovScale;          // square output: aspect ratio 1
        v.yScale = -cam.fovScale;
        v.zNear = cam.zNear;
        v.worldIDX = w;
        v.outputRow = out_row;
        R.views[out_row] = v;
    }
}

// ---- ray casting ------------------------------------------------------------------------------

struct RayShear {
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

// Woop et al. 2013: permute so the dominant direction is z, shear onto it
__device__ __forceinline__ RayShear rayShear(Vector3 d, Diag3x3 inv_d)
{
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = (ax > ay && ax > az) ? 0 : (ay > az ? 1 : 2);
    int kx = kz + 1 == 3 ? 0 : kz + 1;
    int ky = kx + 1 == 3 ? 0 : kx + 1;
    if (d[kz] < 0.f) {
        int t = kx; kx = ky; ky = t;
    }
    return RayShear { kx, ky, kz, d[kx] * inv_d[kz], d[ky] * inv_d[kz], inv_d[kz] };
}

__device__ __forceinline__ bool rayTriangle(Vector3 a, Vector3 b, Vector3 c, const RayShear &rs,
                                            Vector3 org, float t_max, float *out_t, Vector3 *out_n)
{
    const Vector3 A = a - org, B = b - org, C = c - org;
    const float a_kz = A[rs.kz], a_kx = A[rs.kx], a_ky = A[rs.ky];
    const float b_kz = B[rs.kz], b_kx = B[rs.kx], b_ky = B[rs.ky];
    const float c_kz = C[rs.kz], c_kx = C[rs.kx], c_ky = C[rs.ky];

    const float Ax = fmaf(-rs.Sx, a_kz, a_kx), Ay = fmaf(-rs.Sy, a_kz, a_ky);
    const float Bx = fmaf(-rs.Sx, b_kz, b_kx), By = fmaf(-rs.Sy, b_kz, b_ky);
    const float Cx = fmaf(-rs.Sx, c_kz, c_kx), Cy = fmaf(-rs.Sy, c_kz, c_ky);

    float U = fmaf(Cx, By, -Cy * Bx);
    float V = fmaf(Ax, Cy, -Ay * Cx);
    float W = fmaf(Bx, Ay, -By * Ax);

    constexpr float eps = 1e-7;
    if (U > -eps && U < eps) U = 0.f;
    if (V > -eps && V < eps) V = 0.f;
    if (W > -eps && W < eps) W = 0.f;

    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;

    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        // edge case: redo the edge functions in double precision
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
        if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    }

    const float det = U + V + W;
    if (det == 0.f) return false;

    const float Az = rs.Sz * a_kz, Bz = rs.Sz * b_kz, Cz = rs.Sz * c_kz;
    const float T = fmaf(U, Az, fmaf(V, Bz, W * Cz));

    const u32 sign = __float_as_uint(det) & 0x80000000u;
    const float xor_T = __uint_as_float(__float_as_uint(T) ^ sign);
    const float abs_det = copysignf(det, 1.f);
    if (xor_T < 0.0f || xor_T > t_max * abs_det) return false;

    const float rcp = 1.0f / det;
    *out_t = T * rcp;
    *out_n = madrona::math::normalize(cross(B - A, C - A));
    return true;
}

// The same test on triangles that were already translated by the (per view, per
// instance) ray origin and staged in shared memory, with the axis permutation
// resolved at compile time: rays of a warp mostly share their dominant axis,
// and the run-time component selects were the single largest cost of the
// generic version (ncu, profiles/r01_ncu_raycast.csv).  Same operations on the
// same operands as rayTriangle -> same bits.
template <int KZ, bool FLIP>
__device__ __forceinline__ void stagedTriangles(const float *tris, const u32 num_tris, const float Sx,
                                                const float Sy, const float Sz, float &t_obj, bool &hit,
                                                Vector3 &n_obj)
{
    constexpr int K1 = (KZ + 1) % 3, K2 = (KZ + 2) % 3;
    constexpr int KX = FLIP ? K2 : K1, KY = FLIP ? K1 : K2;
    for (u32 tri = 0; tri < num_tris; tri++) {
        const float *t9 = tris + tri * 9;
        const float a_kz = t9[KZ], a_kx = t9[KX], a_ky = t9[KY];
        const float b_kz = t9[3 + KZ], b_kx = t9[3 + KX], b_ky = t9[3 + KY];
        const float c_kz = t9[6 + KZ], c_kx = t9[6 + KX], c_ky = t9[6 + KY];

        const float Ax = fmaf(-Sx, a_kz, a_kx), Ay = fmaf(-Sy, a_kz, a_ky);
        const float Bx = fmaf(-Sx, b_kz, b_kx), By = fmaf(-Sy, b_kz, b_ky);
        const float Cx = fmaf(-Sx, c_kz, c_kx), Cy = fmaf(-Sy, c_kz, c_ky);

        float U = fmaf(Cx, By, -Cy * Bx);
        float V = fmaf(Ax, Cy, -Ay * Cx);
        float W = fmaf(Bx, Ay, -By * Ax);

        constexpr float eps = 1e-7;
        if (U > -eps && U < eps) U = 0.f;
        if (V > -eps && V < eps) V = 0.f;
        if (W > -eps && W < eps) W = 0.f;

        if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) continue;

        if (U == 0.0f || V == 0.0f || W == 0.0f) {
            U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
            V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
            W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
            if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) continue;
        }

        const float det = U + V + W;
        if (det == 0.f) continue;

        const float Az = Sz * a_kz, Bz = Sz * b_kz, Cz = Sz * c_kz;
        const float T = fmaf(U, Az, fmaf(V, Bz, W * Cz));

        const u32 sign = __float_as_uint(det) & 0x80000000u;
        const float xor_T = __uint_as_float(__float_as_uint(T) ^ sign);
        const float abs_det = copysignf(det, 1.f);
        if (xor_T < 0.0f || xor_T > t_obj * abs_det) continue;

        const float rcp = 1.0f / det;
        t_obj = T * rcp;
        hit = true;
        const Vector3 A { t9[0], t9[1], t9[2] }, B { t9[3], t9[4], t9[5] }, C { t9[6], t9[7], t9[8] };
        n_obj = madrona::math::normalize(cross(B - A, C - A));
    }
}

constexpr int kTriArena = 640;      // origin-relative triangles staged per block (23 KB)

__global__ void __launch_bounds__(256, 4)
renderRaycastKernel(EngineState *Sp)
{
    pdlSync();
    EngineState &S = *Sp;
    const RenderState &R = *S.render;
    const TableDesc &out_tbl = S.tables[R.outputArchetype];
    const i32 num_views = min(out_tbl.numRows, R.maxViews);
    const u32 res = R.resolution;
    // One block traces a whole view (the per-view staging below is paid once);
    // a warp traces 8 x 4 pixel tiles: coherent rays (same instances entered,
    // same dominant axis) instead of 32-pixel row segments.
    const u32 tiles_x = (res + 7) / 8;
    const u32 num_tiles = tiles_x * ((res + 3) / 4);
    const size_t bytes_per_view = (size_t)res * res * 4;

    __shared__ RenderInstance staged[kStagedInstances];
    __shared__ RenderInstance stage_tmp[kStagedInstances];
    __shared__ float stage_key[kStagedInstances];
    __shared__ int stage_count[2];
    __shared__ float inst_origin[kStagedInstances][3];   // ray origin in the instance's object space
    __shared__ int inst_tris[kStagedInstances];          // arena offset, -1: not staged, -2: skip instance
    __shared__ int inst_nt[kStagedInstances];            // triangles of the instance's mesh
    __shared__ int arena_used;
    __shared__ float tri_arena[kTriArena * 9];

    for (i32 v = blockIdx.y; v < num_views; v += gridDim.y) {
        const RenderView view = R.views[v];
        const i32 w = view.worldIDX;
        const i32 world_inst = min(R.instanceCounts[w], kStagedInstances);

        // camera frame (shared by every pixel of the view)
        const Quat rot { view.rotation.w, view.rotation.x, view.rotation.y, view.rotation.z };
        const Vector3 ray_start { view.position.x, view.position.y, view.position.z };
        const Vector3 look_at = rot.inv().rotateVec({ 0, 1, 0 });
        const float h = 1.0f / (-view.yScale);
        const float viewport = 2 * h;
        const Vector3 forward = look_at.normalize();
        const Vector3 u = rot.inv().rotateVec({ 1, 0, 0 });
        const Vector3 vv = cross(forward, u).normalize();

        // Stage the world's instances, keeping (in order) only those whose box
        // touches the view frustum: no ray of this view could pass their slab
        // test, so dropping them changes nothing but the work per ray.
        __syncthreads();
        if (threadIdx.x < 64) {
            const int k = threadIdx.x;
            bool keep = false;
            RenderInstance inst;
            if (k < world_inst) {
                inst = R.instances[(size_t)w * R.maxInstancesPerWorld + k];
                keep = true;
                const float hm = fabsf(h) * 1.02f + 0.01f;     // widened: conservative
                const Vector3 normals[4] = { forward * hm - u, forward * hm + u,
                                             forward * hm - vv, forward * hm + vv };
                for (int pl = 0; pl < 4; pl++) {
                    const Vector3 n = normals[pl];
                    // box corner furthest along n
                    const Vector3 far_corner {
                        n.x >= 0 ? inst.aabbMax[0] : inst.aabbMin[0],
                        n.y >= 0 ? inst.aabbMax[1] : inst.aabbMin[1],
                        n.z >= 0 ? inst.aabbMax[2] : inst.aabbMin[2] };
                    if (dot(far_corner - ray_start, n) < 0.f) keep = false;
                }
            }
            const u32 kept = __ballot_sync(0xffffffffu, keep);
            if (threadIdx.x == 0) stage_count[0] = __popc(kept);
            if (threadIdx.x == 32) stage_count[1] = __popc(kept);
            // near-to-far key: squared distance from the eye to the instance's box
            // (0 inside).  Visiting near instances first shrinks t_max early, so
            // the boxes of far instances fail their slab test instead of being
            // entered.  Culled instances sort to the end.
            float key = FLT_MAX;
            if (keep) {
                const float dx = fmaxf(fmaxf(inst.aabbMin[0] - ray_start.x, 0.f), ray_start.x - inst.aabbMax[0]);
                const float dy = fmaxf(fmaxf(inst.aabbMin[1] - ray_start.y, 0.f), ray_start.y - inst.aabbMax[1]);
                const float dz = fmaxf(fmaxf(inst.aabbMin[2] - ray_start.z, 0.f), ray_start.z - inst.aabbMax[2]);
                key = dx * dx + dy * dy + dz * dz;
            }
            stage_key[k] = key;
            stage_tmp[k] = inst;
        }
        __syncthreads();
        if (threadIdx.x < 64) {
            // rank sort, ties broken by the original (engine) order
            const int k = threadIdx.x;
            const float mine = stage_key[k];
            int rank = 0;
            for (int j = 0; j < 64; j++) {
                const float other = stage_key[j];
                rank += (other < mine || (other == mine && j < k)) ? 1 : 0;
            }
            if (mine != FLT_MAX) staged[rank] = stage_tmp[k];
        }
        __syncthreads();
        const i32 num_inst = stage_count[0] + stage_count[1];

        // Per (view, instance) work, hoisted out of the per-ray loop: the ray
        // origin in object space and the mesh's triangles relative to it.
        if (threadIdx.x < 64 && threadIdx.x < num_inst) {
            const int k = threadIdx.x;
            const RenderInstance &inst = staged[k];
            int slot = -1;
            if (inst.scale.x == 0.f || inst.scale.y == 0.f || inst.scale.z == 0.f ||
                    inst.objectID < 0 || (u32)inst.objectID >= R.numMeshes) {
                slot = -2;
            } else {
                const Quat q { inst.rotation.w, inst.rotation.x, inst.rotation.y, inst.rotation.z };
                const Diag3x3 inv_scale = Diag3x3 { inst.scale.x, inst.scale.y, inst.scale.z }.inv();
                const Vector3 p { inst.position.x, inst.position.y, inst.position.z };
                const Vector3 o = inv_scale * q.inv().rotateVec(ray_start - p);
                inst_origin[k][0] = o.x; inst_origin[k][1] = o.y; inst_origin[k][2] = o.z;
            }
            inst_tris[k] = slot;
            inst_nt[k] = slot == -2 ? 0 : (int)R.meshes[inst.objectID].numTriangles;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int used = 0;
            for (int k = 0; k < num_inst; k++) {
                if (inst_tris[k] == -2) continue;
                if (used + inst_nt[k] <= kTriArena) {
                    inst_tris[k] = used;
                    used += inst_nt[k];
                }
            }
            arena_used = used;
        }
        __syncthreads();
        // all staged triangle corners in one flat pass (arena slot -> owning instance by scan)
        for (int e = threadIdx.x; e < arena_used * 3; e += blockDim.x) {
            const int tri_slot = e / 3;
            int k = 0;
            for (int j = 0; j < num_inst; j++) {
                if (inst_tris[j] >= 0 && inst_tris[j] <= tri_slot) k = j;
            }
            const MeshDesc &mesh = R.meshes[staged[k].objectID];
            const u32 i = (u32)(e - inst_tris[k] * 3);
            const u32 vi = R.indices[(size_t)mesh.firstTriangle * 3 + i];
            const Vector3 vert { R.vertices[vi * 3], R.vertices[vi * 3 + 1], R.vertices[vi * 3 + 2] };
            const Vector3 rel = vert - Vector3 { inst_origin[k][0], inst_origin[k][1], inst_origin[k][2] };
            float *dst = tri_arena + (size_t)e * 3;
            dst[0] = rel.x; dst[1] = rel.y; dst[2] = rel.z;
        }
        __syncthreads();

        for (u32 tile = threadIdx.x >> 5; tile < num_tiles; tile += blockDim.x >> 5) {
        const u32 px = (tile % tiles_x) * 8 + (threadIdx.x & 7);
        const u32 py = (tile / tiles_x) * 4 + ((threadIdx.x & 31) >> 3);
        const bool in_image = px < res && py < res;
        if (!in_image) continue;

        // ---- primary ray (bvh_raycast.cpp:58-88)
        const Vector3 horizontal = u * viewport;
        const Vector3 vertical = vv * viewport;
        const Vector3 lower_left = ray_start - horizontal / 2 - vertical / 2 + forward;
        const float pu = ((float)px + 0.5f) / (float)res;
        const float pv = ((float)py + 0.5f) / (float)res;
        Vector3 ray_dir = lower_left + pu * horizontal + pv * vertical - ray_start;
        ray_dir = ray_dir.normalize();

        const Diag3x3 inv_dir = Diag3x3::fromVec(ray_dir).inv();
        float t_max = 10000.f;
        int hit_inst = -1;
        Vector3 hit_normal { 0, 0, 0 };

        for (int k = 0; k < num_inst; k++) {
            const RenderInstance &inst = staged[k];
            AABB box { { inst.aabbMin[0], inst.aabbMin[1], inst.aabbMin[2] },
                       { inst.aabbMax[0], inst.aabbMax[1], inst.aabbMax[2] } };
            if (!box.rayIntersects(ray_start, inv_dir, 0.f, t_max)) continue;
            const int first_tri = inst_tris[k];
            if (first_tri == -2) continue;

            // object-space ray; t is rescaled by |d'| while inside the mesh
            const Quat q { inst.rotation.w, inst.rotation.x, inst.rotation.y, inst.rotation.z };
            const Diag3x3 inv_scale = Diag3x3 { inst.scale.x, inst.scale.y, inst.scale.z }.inv();
            const Vector3 o { inst_origin[k][0], inst_origin[k][1], inst_origin[k][2] };
            Vector3 d = inv_scale * q.inv().rotateVec(ray_dir);
            const float t_scale = d.length();
            float t_obj = t_max * t_scale;
            d /= t_scale;
            const Diag3x3 inv_d = Diag3x3::fromVec(d).inv();
            const RayShear rs = rayShear(d, inv_d);

            const MeshDesc &mesh = R.meshes[inst.objectID];
            bool hit_here = false;
            Vector3 n_obj { 0, 0, 0 };
            if (first_tri >= 0) {
                const float *tris = tri_arena + (size_t)first_tri * 9;
                const bool flip = rs.kx != (rs.kz + 1) % 3;
                switch (rs.kz * 2 + (flip ? 1 : 0)) {
                case 0: stagedTriangles<0, false>(tris, mesh.numTriangles, rs.Sx, rs.Sy, rs.Sz, t_obj, hit_here, n_obj); break;
                case 1: stagedTriangles<0, true>(tris, mesh.numTriangles, rs.Sx, rs.Sy, rs.Sz, t_obj, hit_here, n_obj); break;
                case 2: stagedTriangles<1, false>(tris, mesh.numTriangles, rs.Sx, rs.Sy, rs.Sz, t_obj, hit_here, n_obj); break;
                case 3: stagedTriangles<1, true>(tris, mesh.numTriangles, rs.Sx, rs.Sy, rs.Sz, t_obj, hit_here, n_obj); break;
                case 4: stagedTriangles<2, false>(tris, mesh.numTriangles, rs.Sx, rs.Sy, rs.Sz, t_obj, hit_here, n_obj); break;
                default: stagedTriangles<2, true>(tris, mesh.numTriangles, rs.Sx, rs.Sy, rs.Sz, t_obj, hit_here, n_obj); break;
                }
            } else {
                for (u32 tri = 0; tri < mesh.numTriangles; tri++) {
                    const u32 *idx = R.indices + (size_t)(mesh.firstTriangle + tri) * 3;
                    const Vector3 a { R.vertices[idx[0] * 3], R.vertices[idx[0] * 3 + 1], R.vertices[idx[0] * 3 + 2] };
                    const Vector3 b { R.vertices[idx[1] * 3], R.vertices[idx[1] * 3 + 1], R.vertices[idx[1] * 3 + 2] };
                    const Vector3 c { R.vertices[idx[2] * 3], R.vertices[idx[2] * 3 + 1], R.vertices[idx[2] * 3 + 2] };
                    float t;
                    Vector3 n;
                    if (rayTriangle(a, b, c, rs, o, t_obj, &t, &n)) {
                        t_obj = t;
                        hit_here = true;
                        n_obj = n;
                    }
                }
            }
            t_max = t_obj / t_scale;
            if (hit_here) {
                hit_inst = k;
                hit_normal = q.rotateVec(n_obj);
            }
        }

        const size_t off = (size_t)view.outputRow * bytes_per_view + 4 * ((size_t)px + (size_t)py * res);
        float *depth_out = (float *)((char *)out_tbl.columns[R.depthCol] + off);
        *depth_out = hit_inst >= 0 ? t_max : 0.f;
        if (R.rgbd) {
            unsigned char *rgb = (unsigned char *)out_tbl.columns[R.rgbCol] + off;
            Vector3 color { 0.f, 0.f, 0.f };
            if (hit_inst >= 0) {
                const u32 hex = staged[hit_inst].color;
                const Vector3 base { ((hex >> 16) & 0xFF) / 255.0f, ((hex >> 8) & 0xFF) / 255.0f,
                                     (hex & 0xFF) / 255.0f };
                // no lights: max(0.2, 0) * colour, clamped (bvh_raycast.cpp:921-925)
                color = fmaxf(0.2f, 0.f) * base;
                color.x = fminf(1.f, color.x);
                color.y = fminf(1.f, color.y);
                color.z = fminf(1.f, color.z);
            }
            rgb[0] = (unsigned char)(color.x * 255);
            rgb[1] = (unsigned char)(color.y * 255);
            rgb[2] = (unsigned char)(color.z * 255);
            rgb[3] = 255;
        }
        (void)hit_normal;
        }   // tiles of the view
    }
}

// ---- host side --------------------------------------------------------------------------------

struct HostMeshDesc {       // layout of mb2_render_config::mesh_bvhs entries
    u32 firstTriangle;
    u32 numTriangles;
    float aabbMin[3];
    float aabbMax[3];
};
static_assert(sizeof(HostMeshDesc) == sizeof(MeshDesc), "mesh descriptor layout");

bool renderHostCreate(Executor *ex, const mb2_render_config *rc, std::string *err)
{
    RenderHost *rh = new RenderHost();
    ex->render = rh;
    memset(&rh->hRender, 0, sizeof(RenderState));
    if (cudaMalloc((void **)&rh->dRender, sizeof(RenderState)) != cudaSuccess) {
        *err = "render state allocation failed";
        return false;
    }
    ex->allocations.push_back(rh->dRender);
    RenderState &R = rh->hRender;
    if (rc && rc->render_resolution > 0) {
        R.enabled = 1;
        R.resolution = rc->render_resolution;
        R.rgbd = rc->render_mode == 0 ? 1u : 0u;
        R.nearPlane = rc->near_plane;
        R.farPlane = rc->far_plane;
        R.numMeshes = rc->num_mesh_bvhs;
        R.numTriangles = rc->num_triangles;
        auto upload = [&](const void *src, size_t bytes, const void **dst) {
            void *p = nullptr;
            if (bytes == 0) bytes = 16;
            if (cudaMalloc(&p, bytes) != cudaSuccess) return false;
            ex->allocations.push_back(p);
            if (src) cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
            *dst = p;
            return true;
        };
        if (!upload(rc->mesh_bvhs, sizeof(MeshDesc) * rc->num_mesh_bvhs, (const void **)&R.meshes) ||
            !upload(rc->vertices, sizeof(float) * 3 * rc->num_vertices, (const void **)&R.vertices) ||
            !upload(rc->indices, sizeof(u32) * 3 * rc->num_triangles, (const void **)&R.indices)) {
            *err = "render asset upload failed";
            return false;
        }
    }
    cudaMemcpy(rh->dRender, &R, sizeof(RenderState), cudaMemcpyHostToDevice);
    ex->hState->render = rh->dRender;
    return true;
}

bool renderHostAfterRegistry(Executor *ex, std::string *err)
{
    RenderHost *rh = ex->render;
    EngineState &S = *ex->hState;
    RenderState dev;
    cudaMemcpy(&dev, rh->dRender, sizeof(RenderState), cudaMemcpyDeviceToHost);
    RenderState &R = rh->hRender;
    R = dev;
    if (!R.registered) return true;
    if (!R.enabled) {
        // RenderingSystem used without a CudaBatchRenderConfig: the reference then
        // simply does not ray cast (raycastOutputResolution == 0)
        return true;
    }
    rh->active = true;

    auto col = [&](u32 a, u32 cid) -> int {
        return cid < S.numComponents ? S.columnLookup[a][cid] : -1;
    };
    for (u32 a = 0; a < S.numArchetypes; a++) {
        if (!S.archetypes[a].registered) continue;
        if (col(a, R.cidRenderable) >= 0 && col(a, R.cidPosition) >= 0 && col(a, R.cidRotation) >= 0 &&
                col(a, R.cidScale) >= 0 && col(a, R.cidObjectID) >= 0) {
            if (R.numRenderArchetypes >= (u32)kMaxRenderArchetypes) {
                *err = "too many renderable archetypes";
                return false;
            }
            RenderArchetype &ra = R.renderables[R.numRenderArchetypes++];
            ra.archetype = a;
            ra.cols[RCPosition] = col(a, R.cidPosition);
            ra.cols[RCRotation] = col(a, R.cidRotation);
            ra.cols[RCScale] = col(a, R.cidScale);
            ra.cols[RCObjectID] = col(a, R.cidObjectID);
            ra.cols[RCRenderable] = col(a, R.cidRenderable);
            ra.colorCol = col(a, R.cidColorOverride);
        }
        if (col(a, R.cidRenderCamera) >= 0 && col(a, R.cidPosition) >= 0 && col(a, R.cidRotation) >= 0) {
            if (R.numViewArchetypes >= (u32)kMaxRenderArchetypes) {
                *err = "too many viewing archetypes";
                return false;
            }
            ViewArchetype &va = R.viewers[R.numViewArchetypes++];
            va.archetype = a;
            va.posCol = col(a, R.cidPosition);
            va.rotCol = col(a, R.cidRotation);
            va.camCol = col(a, R.cidRenderCamera);
        }
    }
    R.rgbCol = col(R.outputArchetype, R.cidRGB);
    R.depthCol = col(R.outputArchetype, R.cidDepth);
    R.maxInstancesPerWorld = kStagedInstances;
    R.maxViews = S.tables[R.outputArchetype].capacity;

    auto alloc = [&](void **p, size_t bytes) {
        if (cudaMalloc(p, bytes) != cudaSuccess) return false;
        ex->allocations.push_back(*p);
        cudaMemset(*p, 0, bytes);
        return true;
    };
    const size_t W = S.numWorlds;
    if (!alloc((void **)&R.instances, sizeof(RenderInstance) * W * R.maxInstancesPerWorld) ||
        !alloc((void **)&R.instanceCounts, sizeof(i32) * W) ||
        !alloc((void **)&R.views, sizeof(RenderView) * (size_t)R.maxViews)) {
        *err = "render buffers allocation failed";
        return false;
    }
    cudaMemcpy(rh->dRender, &R, sizeof(RenderState), cudaMemcpyHostToDevice);
    return true;
}

void renderHostDestroy(Executor *ex)
{
    delete ex->render;
    ex->render = nullptr;
}

bool renderEnqueuePrepare(Executor *ex, cudaStream_t s, std::string *err)
{
    RenderHost *rh = ex->render;
    if (!rh || !rh->active) return true;   // rendering not configured: nothing to prepare
    (void)err;
    const unsigned W = ex->hState->numWorlds;
    launchK(renderGatherInstancesKernel, dim3((W * 32 + 127) / 128), dim3(128), 0, s, ex->dState);
    int max_cap = 256;
    for (u32 i = 0; i < rh->hRender.numViewArchetypes; i++) {
        max_cap = std::max(max_cap, ex->hState->tables[rh->hRender.viewers[i].archetype].capacity);
    }
    dim3 grid((unsigned)std::min((max_cap + 255) / 256, ex->numSMs * 4),
              std::max(rh->hRender.numViewArchetypes, 1u));
    launchK(renderGatherViewsKernel, dim3(grid), dim3(256), 0, s, ex->dState);
    return true;
}

LaunchGraph *physicsBuildRenderGraph(Executor *ex, std::string *err)
{
    RenderHost *rh = ex->render;
    if (!rh || !rh->active) {
        *err = "buildRenderGraph: no CudaBatchRenderConfig was given (or the simulator never called "
               "RenderingSystem::registerTypes)";
        return nullptr;
    }
    const RenderState &R = rh->hRender;
    LaunchGraph *g = new LaunchGraph();
    g->owner = ex;
    g->name = "render";
    cudaSetDevice(ex->gpu);
    if (cudaStreamBeginCapture(ex->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        *err = "cudaStreamBeginCapture failed";
        delete g;
        return nullptr;
    }
    const unsigned view_blocks = (unsigned)std::max(1, std::min(R.maxViews, 65535));
    launchK(renderRaycastKernel, dim3(1, view_blocks), dim3(256), 0, ex->stream, ex->dState);
    launchStatusCopy(ex, ex->stream);
    cudaError_t e = cudaStreamEndCapture(ex->stream, &g->graph);
    if (e != cudaSuccess || cudaGraphInstantiate(&g->exec, g->graph, 0) != cudaSuccess) {
        *err = std::string("render graph capture failed: ") + cudaGetErrorString(cudaGetLastError());
        if (g->graph) cudaGraphDestroy(g->graph);
        delete g;
        return nullptr;
    }
    g->numKernels = 2;
    return g;
}

uint64_t renderBytesPerFrame(Executor *ex)
{
    RenderHost *rh = ex->render;
    if (!rh || !rh->active) return 0;
    return 0;
}

}
