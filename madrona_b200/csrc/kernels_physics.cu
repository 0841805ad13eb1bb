// kernels_physics.cu -- hot system 2 (placeholder until the XPBD kernels land).
#include "physics_host.hpp"

namespace mb2 {

bool physicsHostCreate(Executor *, std::string *) { return true; }
bool physicsHostAfterRegistry(Executor *, const mb2_render_config *, std::string *) { return true; }
void physicsHostDestroy(Executor *) {}
bool physicsEnqueueNode(Executor *, const NodeRecord &rec, cudaStream_t, std::string *err)
{
    *err = "physics node kind " + std::to_string(rec.kind) + " not available in this build";
    return false;
}
uint64_t physicsNodeBytes(Executor *, const NodeRecord &, const char **name, int64_t *rows)
{
    *name = "physics";
    *rows = 0;
    return 0;
}
LaunchGraph *physicsBuildRenderGraph(Executor *, std::string *err)
{
    *err = "batch ray-cast renderer not available in this build";
    return nullptr;
}

}
