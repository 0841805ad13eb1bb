// Entry points the executor launches once at construction (reference:
// src/mw/device/include/madrona/mw_gpu_entry.hpp:12-91).  They run the
// simulator's own registerTypes / world constructors / setupTasks on the
// device; here they have fixed extern "C" names so the host needs no symbol
// scan for them.
#pragma once
#include <madrona/taskgraph_builder.hpp>

namespace madrona {
namespace mwGPU {
namespace entryKernels {

template <typename ContextT, typename WorldT, typename ConfigT, typename InitT>
inline void initECS()
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    mb2::EngineState &S = engine();
    S.worldDataNeeded = (uint32_t)sizeof(WorldT);
    S.worldDataAlignNeeded = (uint32_t)alignof(WorldT);
    // Entity = component 0, WorldID = component 1 (src/mw/device/state.cpp:150-151)
    StateManager *mgr = getStateManager();
    mgr->registerComponent<Entity>();
    mgr->registerComponent<WorldID>();
    ECSRegistry registry(mgr, nullptr);
    WorldT::registerTypes(registry, *(ConfigT *)S.userConfig);
}

template <typename ContextT, typename WorldT, typename ConfigT, typename InitT>
inline void initWorlds()
{
    mb2::EngineState &S = engine();
    int32_t world_idx = (int32_t)(threadIdx.x + blockDim.x * blockIdx.x);
    if (world_idx >= (int32_t)S.numWorlds) return;

    WorldT *world = (WorldT *)(S.worldData + (size_t)world_idx * S.worldDataStride);
    ContextT ctx(world, WorkerInit { WorldID { world_idx } });
    const InitT *inits = (const InitT *)S.worldInits;
    new (world) WorldT(ctx, *(const ConfigT *)S.userConfig, inits[world_idx]);
}

template <typename ContextT, typename WorldT, typename ConfigT, typename InitT>
inline void initTasks()
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    mb2::EngineState &S = engine();
    TaskGraphManager mgr(S.numTaskGraphs);
    WorldT::setupTasks(mgr, *(ConfigT *)S.userConfig);
}

}
}
}

#define MADRONA_BUILD_MWGPU_ENTRY(ContextT, WorldT, ConfigT, InitT) \
    extern "C" __global__ void mb2_entry_init_ecs() { \
        ::madrona::mwGPU::entryKernels::initECS<ContextT, WorldT, ConfigT, InitT>(); } \
    extern "C" __global__ void mb2_entry_init_worlds() { \
        ::madrona::mwGPU::entryKernels::initWorlds<ContextT, WorldT, ConfigT, InitT>(); } \
    extern "C" __global__ void mb2_entry_init_tasks() { \
        ::madrona::mwGPU::entryKernels::initTasks<ContextT, WorldT, ConfigT, InitT>(); } \
    static_assert(sizeof(WorldT) > 0);
