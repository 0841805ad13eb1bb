// Compile-only: a Manager written against the REFERENCE headers (utils, span,
// optional, render assets) builds unchanged when <madrona/mw_gpu.hpp> resolves to
// this engine's facade (madrona_b200/host before the reference include dir).
// Mirrors what a simulator's mgr.cpp does (e.g. Manager::Impl::init): fill
// StateConfig / CompileConfig / CudaBatchRenderConfig the reference way, build
// launch graphs, run, getExported.
#include <madrona/utils.hpp>
#include <madrona/span.hpp>
#include <madrona/optional.hpp>
#include <madrona/heap_array.hpp>
#include <madrona/mw_gpu.hpp>

#include <array>

using namespace madrona;

struct WorldInit { uint32_t seed; };
struct Config { uint32_t maxSteps; };
enum class TaskGraphID : uint32_t { Step, NumTaskGraphs };

static MWCudaExecutor makeExecutor(int gpu_id, uint32_t num_worlds, const char *sim_src,
                                   const render::MeshBVHData &bvh, const render::MaterialData &mats,
                                   bool enable_render)
{
    CUcontext cu_ctx = MWCudaExecutor::initCUDA(gpu_id);

    HeapArray<WorldInit> world_inits(num_worlds);
    for (CountT i = 0; i < (CountT)num_worlds; i++) world_inits[i] = WorldInit { (uint32_t)i };
    Config cfg { 200 };

    std::array<const char *, 1> sources { sim_src };
    std::array<const char *, 1> flags { "-DEXAMPLE=1" };

    Optional<CudaBatchRenderConfig> render_cfg = Optional<CudaBatchRenderConfig>::none();
    if (enable_render) {
        render_cfg = CudaBatchRenderConfig {
            .renderMode = CudaBatchRenderConfig::RenderMode::RGBD,
            .geoBVHData = bvh,
            .materialData = mats,
            .renderResolution = 64,
            .nearPlane = 0.001f,
            .farPlane = 1000.f,
        };
    }

    return MWCudaExecutor({
        .worldInitPtr = world_inits.data(),
        .numWorldInitBytes = sizeof(WorldInit),
        .userConfigPtr = (void *)&cfg,
        .numUserConfigBytes = sizeof(Config),
        .numWorldDataBytes = 0,
        .worldDataAlignment = 16,
        .numWorlds = num_worlds,
        .numTaskGraphs = (uint32_t)TaskGraphID::NumTaskGraphs,
        .numExportedBuffers = 4,
    }, {
        { sources.data(), (CountT)sources.size() },
        { flags.data(), (CountT)flags.size() },
        CompileConfig::OptMode::LTO,
    }, cu_ctx, render_cfg);
}

int main(int argc, char **argv)
{
    if (argc < 2) return 0;     // the test only builds this file
    render::MeshBVHData bvh {};
    render::MaterialData mats {};
    MWCudaExecutor exec = makeExecutor(0, 16, argv[1], bvh, mats, false);
    MWCudaLaunchGraph step = exec.buildLaunchGraph(TaskGraphID::Step);
    MWCudaLaunchGraph all = exec.buildLaunchGraphAllTaskGraphs();
    exec.run(step);
    exec.runAsync(all, (cudaStream_t)0);
    void *p = exec.getExported(0);
    return p ? 0 : 1;
}
