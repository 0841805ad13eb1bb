// Host hooks for the engine-owned systems (rigid-body physics, ray-cast
// renderer): kernels_physics.cu / kernels_render.cu.
#pragma once
#include "engine.hpp"
#include "../../include/madrona_b200.h"

namespace mb2 {

bool physicsHostCreate(Executor *ex, std::string *err);
bool physicsHostAfterRegistry(Executor *ex, const mb2_render_config *rc, std::string *err);
void physicsHostDestroy(Executor *ex);
// reads device-side facts that select kernel variants; call before stream capture begins
void physicsBeforeGraphCapture(Executor *ex);
bool physicsEnqueueNode(Executor *ex, const NodeRecord &rec, cudaStream_t s, std::string *err);
bool physicsEnqueueNodes(Executor *ex, const NodeRecord *recs, uint32_t count, cudaStream_t s,
                         std::string *err);
LaunchGraph *physicsBuildRenderGraph(Executor *ex, std::string *err);
// batch ray-cast renderer (kernels_render.cu)
bool renderHostCreate(Executor *ex, const mb2_render_config *rc, std::string *err);
bool renderHostAfterRegistry(Executor *ex, std::string *err);
void renderHostDestroy(Executor *ex);
bool renderEnqueuePrepare(Executor *ex, cudaStream_t s, std::string *err);
void *renderDebugHitBuffer(Executor *ex);
void *renderDebugBuffer(Executor *ex, int which, int64_t *stride_out);
uint64_t renderBytesPerFrame(Executor *ex, int64_t num_views);
// algorithmic bytes of one launch of a physics node + a short name (profiling)
uint64_t physicsNodeBytes(Executor *ex, const NodeRecord &rec, const char **name, int64_t *rows);

}
