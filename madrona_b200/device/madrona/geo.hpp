// Reference: include/madrona/geo.hpp:7-45 (layouts only; the GJK / ray helper
// routines are engine-internal here).
#pragma once
#include <madrona/math.hpp>
namespace madrona::geo {

struct HalfEdge {
    uint32_t next;
    uint32_t rootVertex;
    uint32_t face;
};

struct Plane {
    math::Vector3 normal; // potentially unnormalized
    float d;
};

struct Segment {
    math::Vector3 p1;
    math::Vector3 p2;
};

struct HalfEdgeMesh {
    template <typename Fn>
    MB2_HD inline void iterateFaceIndices(uint32_t face, Fn &&fn) const
    {
        uint32_t start = faceBaseHalfEdges[face];
        uint32_t cur = start;
        do {
            fn(halfEdges[cur].rootVertex);
            cur = halfEdges[cur].next;
        } while (cur != start);
    }
    // twins are stored adjacently: (2k, 2k+1)
    MB2_HD inline uint32_t twinIDX(uint32_t half_edge_id) const { return half_edge_id ^ 1u; }
    MB2_HD inline uint32_t numEdges() const { return numHalfEdges / 2; }
    MB2_HD inline uint32_t edgeToHalfEdge(uint32_t edge_id) const { return edge_id * 2; }

    HalfEdge *halfEdges;
    uint32_t *faceBaseHalfEdges;
    Plane *facePlanes;
    math::Vector3 *vertices;

    uint32_t numHalfEdges;
    uint32_t numFaces;
    uint32_t numVertices;
};

}
