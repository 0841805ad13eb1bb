// mesh_bvh.cpp -- host-side BLAS builder: triangle meshes -> reference-format
// MeshBVHs (4-wide quantised nodes over de-indexed triangles, <= 2 triangles per
// leaf) + upload as a render::MeshBVHData-compatible block.
//
// Role of the reference's MeshBVHBuilder (src/common/mesh_bvh_builder.cpp, which
// drives embree's SAH builder and then packs QBVH nodes) and of
// render::AssetProcessor::makeBVHData (src/render/asset_processor.cpp, which
// concatenates all meshes' arrays and uploads them with cudaMalloc).  embree is
// not available; the builder here is a plain top-down binned-SAH binary build
// collapsed to 4-wide nodes.  Any valid BVH yields the same closest hit, so the
// pixels do not depend on which builder made the tree (tests/test_render_bvh.py
// checks hits against a brute-force scan of all triangles).
#include "../../include/madrona_b200.h"
#include "engine.hpp"
#include "render_bvh.h"

#include <algorithm>
#include <cfloat>
#include <cstring>
#include <vector>

namespace mb2 {

namespace {

struct Box {
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
    float hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    void grow(const float *p)
    {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    void grow(const Box &o)
    {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], o.lo[a]);
            hi[a] = std::max(hi[a], o.hi[a]);
        }
    }
    float area() const
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx < 0 ? 0.f : 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

struct BinNode {
    Box box;
    int left = -1, right = -1;      // children, or
    int first = 0, count = 0;       // triangle range (leaf)
};

struct Builder {
    const float *pos;
    const uint32_t *idx;
    std::vector<uint32_t> order;       // triangle permutation
    std::vector<Box> triBox;
    std::vector<float> centroid;       // 3 per triangle
    std::vector<BinNode> nodes;

    int build(int first, int count)
    {
        const int id = (int)nodes.size();
        nodes.emplace_back();
        Box box, cbox;
        for (int i = first; i < first + count; i++) {
            box.grow(triBox[order[i]]);
            cbox.grow(&centroid[order[i] * 3]);
        }
        nodes[id].box = box;
        if (count <= kBLASLeafWidth) {
            nodes[id].first = first;
            nodes[id].count = count;
            return id;
        }
        // binned SAH over the widest centroid axis (16 bins), median fallback
        int axis = 0;
        float ext = -1.f;
        for (int a = 0; a < 3; a++) {
            if (cbox.hi[a] - cbox.lo[a] > ext) {
                ext = cbox.hi[a] - cbox.lo[a];
                axis = a;
            }
        }
        int mid = first + count / 2;
        if (ext > 0.f) {
            constexpr int kBins = 16;
            Box bin_box[kBins];
            int bin_n[kBins] = {};
            const float scale = kBins / ext;
            auto bin_of = [&](uint32_t t) {
                int b = (int)((centroid[t * 3 + axis] - cbox.lo[axis]) * scale);
                return std::min(std::max(b, 0), kBins - 1);
            };
            for (int i = first; i < first + count; i++) {
                const int b = bin_of(order[i]);
                bin_box[b].grow(triBox[order[i]]);
                bin_n[b]++;
            }
            float right_area[kBins];
            int right_n[kBins];
            Box acc;
            int n = 0;
            for (int b = kBins - 1; b > 0; b--) {
                acc.grow(bin_box[b]);
                n += bin_n[b];
                right_area[b] = acc.area();
                right_n[b] = n;
            }
            Box lacc;
            int ln = 0, best = -1;
            float best_cost = FLT_MAX;
            for (int b = 0; b < kBins - 1; b++) {
                lacc.grow(bin_box[b]);
                ln += bin_n[b];
                if (ln == 0 || right_n[b + 1] == 0) continue;
                const float cost = lacc.area() * ln + right_area[b + 1] * right_n[b + 1];
                if (cost < best_cost) {
                    best_cost = cost;
                    best = b;
                }
            }
            if (best >= 0) {
                auto it = std::stable_partition(order.begin() + first, order.begin() + first + count,
                                                [&](uint32_t t) { return bin_of(t) <= best; });
                mid = (int)(it - order.begin());
            } else {
                std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                                 [&](uint32_t a, uint32_t b) {
                                     return centroid[a * 3 + axis] < centroid[b * 3 + axis];
                                 });
            }
        }
        if (mid <= first || mid >= first + count) mid = first + count / 2;
        const int l = build(first, mid - first);
        const int r = build(mid, first + count - mid);
        nodes[id].left = l;
        nodes[id].right = r;
        return id;
    }
};

struct BuiltMesh {
    std::vector<QBVHNode> nodes;
    std::vector<BVHVertex> verts;
    std::vector<LeafMaterial> leafMats;
    float rootMin[3], rootMax[3];
    int32_t material;
};

void collapse(const Builder &b, int bin_root, BuiltMesh &out)
{
    // breadth-first: wide node k takes a binary node and pulls grandchildren up
    // (largest box first) until it has 4 children or only leaves are left
    struct Pending { int bin; };
    std::vector<Pending> queue { { bin_root } };
    out.nodes.clear();
    out.nodes.emplace_back();
    for (size_t k = 0; k < queue.size(); k++) {
        const BinNode &root = b.nodes[queue[k].bin];
        std::vector<int> kids;
        if (root.left < 0) {
            kids.push_back(queue[k].bin);     // a mesh of <= 2 triangles: the root is a leaf
        } else {
            kids = { root.left, root.right };
            while ((int)kids.size() < kBVHWidth) {
                int pick = -1;
                float best = -1.f;
                for (int i = 0; i < (int)kids.size(); i++) {
                    const BinNode &c = b.nodes[kids[i]];
                    if (c.left >= 0 && c.box.area() > best) {
                        best = c.box.area();
                        pick = i;
                    }
                }
                if (pick < 0) break;
                const BinNode &c = b.nodes[kids[pick]];
                kids[pick] = c.left;
                kids.push_back(c.right);
            }
        }
        float cmin[kBVHWidth][3], cmax[kBVHWidth][3];
        for (int i = 0; i < (int)kids.size(); i++) {
            const BinNode &c = b.nodes[kids[i]];
            for (int a = 0; a < 3; a++) {
                cmin[i][a] = c.box.lo[a];
                cmax[i][a] = c.box.hi[a];
            }
        }
        QBVHNode node;
        memset(&node, 0, sizeof(node));
        quantizeNode(node, (int)kids.size(), cmin, cmax);
        for (int i = 0; i < (int)kids.size(); i++) {
            const BinNode &c = b.nodes[kids[i]];
            if (c.left < 0) {
                node.childrenIdx[i] = 0x80000000u | (uint32_t)c.first;   // triangles keep builder order
                node.triSize[i] = (uint8_t)c.count;
            } else {
                node.childrenIdx[i] = (uint32_t)out.nodes.size();
                node.triSize[i] = 0;
                out.nodes.emplace_back();
                queue.push_back({ kids[i] });
            }
        }
        out.nodes[k] = node;
    }
}

}

struct MeshBVHBundle {
    int gpu = -1;
    // host copies (pointers inside are offsets until upload)
    std::vector<QBVHNode> nodes;
    std::vector<LeafMaterial> leafMats;
    std::vector<BVHVertex> verts;
    std::vector<MeshBVH> meshes;           // host view: pointers into the host vectors
    std::vector<uint32_t> triSource;       // new triangle index -> source triangle (per mesh concatenated)
    std::vector<uint32_t> meshFirstTri;    // first triangle of each mesh in the concatenated arrays
    MeshBVHData device {};                 // device pointers (gpu >= 0)
    MeshBVHData host {};
    std::vector<void *> deviceAllocs;
};

}

using namespace mb2;

extern "C" {

mb2_mesh_bvh_data *mb2_build_mesh_bvhs(const mb2_mesh_source *meshes, uint32_t num_meshes, int gpu_id)
{
    if (!meshes || num_meshes == 0) {
        setError("mb2_build_mesh_bvhs: no meshes");
        return nullptr;
    }
    MeshBVHBundle *bundle = new MeshBVHBundle();
    bundle->gpu = gpu_id;
    std::vector<size_t> node_off, tri_off;
    for (uint32_t m = 0; m < num_meshes; m++) {
        const mb2_mesh_source &src = meshes[m];
        if (!src.positions || !src.indices || src.num_triangles == 0) {
            setError("mb2_build_mesh_bvhs: mesh " + std::to_string(m) + " is empty");
            delete bundle;
            return nullptr;
        }
        Builder b;
        b.pos = src.positions;
        b.idx = src.indices;
        const uint32_t nt = src.num_triangles;
        b.order.resize(nt);
        b.triBox.resize(nt);
        b.centroid.resize((size_t)nt * 3);
        for (uint32_t t = 0; t < nt; t++) {
            b.order[t] = t;
            for (int k = 0; k < 3; k++) {
                const uint32_t vi = src.indices[t * 3 + k];
                if (vi >= src.num_vertices) {
                    setError("mb2_build_mesh_bvhs: index out of range in mesh " + std::to_string(m));
                    delete bundle;
                    return nullptr;
                }
                b.triBox[t].grow(src.positions + (size_t)vi * 3);
            }
            for (int a = 0; a < 3; a++) b.centroid[t * 3 + a] = 0.5f * (b.triBox[t].lo[a] + b.triBox[t].hi[a]);
        }
        b.nodes.reserve((size_t)nt * 2);
        const int root = b.build(0, (int)nt);
        BuiltMesh built;
        collapse(b, root, built);

        node_off.push_back(bundle->nodes.size());
        tri_off.push_back(bundle->verts.size() / 3);
        bundle->meshFirstTri.push_back((uint32_t)(bundle->verts.size() / 3));
        bundle->nodes.insert(bundle->nodes.end(), built.nodes.begin(), built.nodes.end());
        for (uint32_t t = 0; t < nt; t++) {
            const uint32_t s = b.order[t];
            bundle->triSource.push_back(s);
            for (int k = 0; k < 3; k++) {
                const uint32_t vi = src.indices[s * 3 + k];
                BVHVertex v;
                v.pos[0] = src.positions[(size_t)vi * 3];
                v.pos[1] = src.positions[(size_t)vi * 3 + 1];
                v.pos[2] = src.positions[(size_t)vi * 3 + 2];
                v.uv[0] = src.uvs ? src.uvs[(size_t)vi * 2] : 0.f;
                v.uv[1] = src.uvs ? src.uvs[(size_t)vi * 2 + 1] : 0.f;
                bundle->verts.push_back(v);
            }
            bundle->leafMats.push_back(LeafMaterial { src.material_idx });
        }
        MeshBVH mesh;
        memset(&mesh, 0, sizeof(mesh));
        for (int a = 0; a < 3; a++) {
            mesh.rootAABBMin[a] = b.nodes[root].box.lo[a];
            mesh.rootAABBMax[a] = b.nodes[root].box.hi[a];
        }
        mesh.numNodes = (uint32_t)built.nodes.size();
        mesh.numLeaves = nt;
        mesh.numVerts = nt * 3;
        mesh.materialIDX = src.material_idx;
        mesh.magic = 0x69426942u;
        bundle->meshes.push_back(mesh);
    }
    // host view
    for (uint32_t m = 0; m < num_meshes; m++) {
        bundle->meshes[m].nodes = bundle->nodes.data() + node_off[m];
        bundle->meshes[m].leafMats = bundle->leafMats.data() + tri_off[m];
        bundle->meshes[m].vertices = bundle->verts.data() + tri_off[m] * 3;
    }
    bundle->host = MeshBVHData { bundle->nodes.data(), bundle->nodes.size(), bundle->leafMats.data(),
                                 bundle->leafMats.size(), bundle->verts.data(), bundle->verts.size(),
                                 bundle->meshes.data(), bundle->meshes.size() };
    if (gpu_id >= 0) {
        cudaSetDevice(gpu_id);
        auto up = [&](const void *src, size_t bytes) -> void * {
            void *p = nullptr;
            if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) return nullptr;
            cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
            bundle->deviceAllocs.push_back(p);
            return p;
        };
        QBVHNode *d_nodes = (QBVHNode *)up(bundle->nodes.data(), bundle->nodes.size() * sizeof(QBVHNode));
        LeafMaterial *d_mats = (LeafMaterial *)up(bundle->leafMats.data(), bundle->leafMats.size() * sizeof(LeafMaterial));
        BVHVertex *d_verts = (BVHVertex *)up(bundle->verts.data(), bundle->verts.size() * sizeof(BVHVertex));
        std::vector<MeshBVH> dev_meshes = bundle->meshes;
        for (uint32_t m = 0; m < num_meshes; m++) {
            dev_meshes[m].nodes = d_nodes + node_off[m];
            dev_meshes[m].leafMats = d_mats + tri_off[m];
            dev_meshes[m].vertices = d_verts + tri_off[m] * 3;
        }
        MeshBVH *d_meshes = (MeshBVH *)up(dev_meshes.data(), dev_meshes.size() * sizeof(MeshBVH));
        if (!d_nodes || !d_mats || !d_verts || !d_meshes) {
            setError("mb2_build_mesh_bvhs: device allocation failed");
            mb2_mesh_bvh_data_destroy((mb2_mesh_bvh_data *)bundle);
            return nullptr;
        }
        bundle->device = MeshBVHData { d_nodes, bundle->nodes.size(), d_mats, bundle->leafMats.size(), d_verts,
                                       bundle->verts.size(), d_meshes, bundle->meshes.size() };
    }
    return (mb2_mesh_bvh_data *)bundle;
}

const void *mb2_mesh_bvh_data_view(const mb2_mesh_bvh_data *data, int device)
{
    const MeshBVHBundle *b = (const MeshBVHBundle *)data;
    if (!b) return nullptr;
    return device ? (const void *)&b->device : (const void *)&b->host;
}

const uint32_t *mb2_mesh_bvh_triangle_sources(const mb2_mesh_bvh_data *data)
{
    const MeshBVHBundle *b = (const MeshBVHBundle *)data;
    return b ? b->triSource.data() : nullptr;
}

void mb2_mesh_bvh_data_destroy(mb2_mesh_bvh_data *data)
{
    MeshBVHBundle *b = (MeshBVHBundle *)data;
    if (!b) return;
    if (b->gpu >= 0) {
        cudaSetDevice(b->gpu);
        for (void *p : b->deviceAllocs) cudaFree(p);
    }
    delete b;
}

}
