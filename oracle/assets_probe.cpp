// TEST INFRASTRUCTURE ONLY.  Runs the reference's own
// RigidBodyAssets::processRigidBodyAssets (src/physics/physics_assets.cpp:1268,
// build_convex_hulls = false) on hull meshes / collision objects read from a
// file and writes the resulting arrays in a pointer-free form; tests/
// test_physics_assets.py compares them bit for bit with the arrays
// mb2_process_rigid_body_assets produces from the same input.
//
// input : u32 numHulls { u32 nVerts, u32 nFaces, u32 nIdx, f32 pos[3 nVerts], u32 counts[nFaces], u32 idx[nIdx] }
//         u32 numObjs  { u32 nPrims, f32 invMass, f32 muS, f32 muD, { u32 type, f32 radius, u32 hull }[nPrims] }
// output: u32 {numHalfEdges, numFaces, numVerts, numPrims, numObjs}, halfEdges, faceBase, planes, verts,
//         per prim u32 type + (f32 radius | u32 {heOff, faceOff, vertOff, numHE, numFaces, numVerts}),
//         primAABBs, metadatas, objAABBs, primOffsets, primCounts
#include <madrona/physics_assets.hpp>

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace madrona;
using namespace madrona::phys;

template <typename T>
static T rd(FILE *f)
{
    T v;
    if (fread(&v, sizeof(T), 1, f) != 1) abort();
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *in = fopen(argv[1], "rb");
    FILE *out = fopen(argv[2], "wb");
    if (!in || !out) return 2;
    uint32_t num_hulls = rd<uint32_t>(in);
    std::vector<std::vector<math::Vector3>> pos(num_hulls);
    std::vector<std::vector<uint32_t>> counts(num_hulls), idx(num_hulls);
    std::vector<imp::SourceMesh> meshes(num_hulls);
    for (uint32_t h = 0; h < num_hulls; h++) {
        uint32_t nv = rd<uint32_t>(in), nf = rd<uint32_t>(in), ni = rd<uint32_t>(in);
        pos[h].resize(nv);
        counts[h].resize(nf);
        idx[h].resize(ni);
        if (fread(pos[h].data(), 12, nv, in) != nv) abort();
        if (fread(counts[h].data(), 4, nf, in) != nf) abort();
        if (fread(idx[h].data(), 4, ni, in) != ni) abort();
        meshes[h] = imp::SourceMesh {};
        meshes[h].positions = pos[h].data();
        meshes[h].indices = idx[h].data();
        meshes[h].faceCounts = counts[h].data();
        meshes[h].numVertices = nv;
        meshes[h].numFaces = nf;
    }
    uint32_t num_objs = rd<uint32_t>(in);
    std::vector<std::vector<SourceCollisionPrimitive>> prims(num_objs);
    std::vector<SourceCollisionObject> objs;
    for (uint32_t o = 0; o < num_objs; o++) {
        uint32_t np = rd<uint32_t>(in);
        float inv_mass = rd<float>(in), mu_s = rd<float>(in), mu_d = rd<float>(in);
        prims[o].resize(np);
        for (uint32_t p = 0; p < np; p++) {
            uint32_t type = rd<uint32_t>(in);
            float radius = rd<float>(in);
            uint32_t hull = rd<uint32_t>(in);
            SourceCollisionPrimitive sp {};
            sp.type = (CollisionPrimitive::Type)type;
            if (sp.type == CollisionPrimitive::Type::Sphere) sp.sphere.radius = radius;
            else if (sp.type == CollisionPrimitive::Type::Hull) sp.hullInput.hullIDX = hull;
            prims[o][p] = sp;
        }
        objs.push_back(SourceCollisionObject { Span<const SourceCollisionPrimitive>(prims[o].data(), np),
                                               inv_mass, RigidBodyFrictionData { mu_s, mu_d } });
    }
    StackAlloc tmp;
    RigidBodyAssets assets;
    CountT num_bytes;
    void *buf = RigidBodyAssets::processRigidBodyAssets(
        Span<const imp::SourceMesh>(meshes.data(), num_hulls),
        Span<const SourceCollisionObject>(objs.data(), num_objs), false, tmp, &assets, &num_bytes);
    if (!buf) {
        fprintf(stderr, "processRigidBodyAssets failed\n");
        return 1;
    }
    auto wr = [&](const void *p, size_t n) { fwrite(p, 1, n, out); };
    uint32_t hdr[5] = { assets.hullData.numHalfEdges, assets.hullData.numFaces, assets.hullData.numVerts,
                        assets.totalNumPrimitives, assets.numObjs };
    wr(hdr, sizeof(hdr));
    wr(assets.hullData.halfEdges, sizeof(geo::HalfEdge) * hdr[0]);
    wr(assets.hullData.faceBaseHalfEdges, 4 * hdr[1]);
    wr(assets.hullData.facePlanes, sizeof(geo::Plane) * hdr[1]);
    wr(assets.hullData.vertices, 12 * hdr[2]);
    for (uint32_t p = 0; p < hdr[3]; p++) {
        const CollisionPrimitive &prim = assets.primitives[p];
        uint32_t type = (uint32_t)prim.type;
        wr(&type, 4);
        if (prim.type == CollisionPrimitive::Type::Sphere) {
            wr(&prim.sphere.radius, 4);
        } else if (prim.type == CollisionPrimitive::Type::Hull) {
            const geo::HalfEdgeMesh &m = prim.hull.halfEdgeMesh;
            uint32_t v[6] = { (uint32_t)(m.halfEdges - assets.hullData.halfEdges),
                              (uint32_t)(m.faceBaseHalfEdges - assets.hullData.faceBaseHalfEdges),
                              (uint32_t)(m.vertices - assets.hullData.vertices),
                              m.numHalfEdges, m.numFaces, m.numVertices };
            wr(v, sizeof(v));
        }
    }
    wr(assets.primitiveAABBs, sizeof(math::AABB) * hdr[3]);
    wr(assets.metadatas, sizeof(RigidBodyMetadata) * hdr[4]);
    wr(assets.objAABBs, sizeof(math::AABB) * hdr[4]);
    wr(assets.primOffsets, 4 * hdr[4]);
    wr(assets.primCounts, 4 * hdr[4]);
    fclose(out);
    return 0;
}
