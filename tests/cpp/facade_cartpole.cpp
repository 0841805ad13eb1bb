// Compiles a reference-style Manager snippet against the C++ facade
// (madrona_b200/host/madrona/mw_gpu.hpp).  With a GPU it runs the cartpole
// fixture for a few steps and prints the first world's state; without one the
// constructor aborts with a message (reference FATAL behaviour), which the
// test expects.
#include <madrona/mw_gpu.hpp>

#include <cstring>
#include <vector>

extern "C" int cudaMemcpy(void *, const void *, size_t, int);

struct Config { uint32_t maxSteps; };
struct WorldInit { uint32_t seed; };

int main(int argc, char **argv)
{
    const char *src = argc > 1 ? argv[1] : "sims/cartpole/sim.cpp";
    const uint32_t num_worlds = 64;
    Config cfg { 200 };
    std::vector<WorldInit> inits(num_worlds);
    for (uint32_t i = 0; i < num_worlds; i++) inits[i].seed = i;

    const char *sources[] = { src };
    const char *flags[] = { "-DCARTPOLE_FACADE_TEST=1" };

    madrona::MWCudaExecutor exec({
        .worldInitPtr = inits.data(),
        .numWorldInitBytes = sizeof(WorldInit),
        .userConfigPtr = &cfg,
        .numUserConfigBytes = sizeof(Config),
        .numWorldDataBytes = 0,
        .worldDataAlignment = 16,
        .numWorlds = num_worlds,
        .numTaskGraphs = 1,
        .numExportedBuffers = 5,
    }, {
        .userSources = madrona::Span<const char * const>(sources, 1),
        .userCompileFlags = madrona::Span<const char * const>(flags, 1),
    }, madrona::MWCudaExecutor::initCUDA(0));

    madrona::MWCudaLaunchGraph step = exec.buildLaunchGraphAllTaskGraphs();
    for (int i = 0; i < 10; i++) exec.run(step);

    float state[4];
    cudaMemcpy(state, exec.getExported(2), sizeof(state), 2 /* DtoH */);
    printf("state %.9g %.9g %.9g %.9g\n", state[0], state[1], state[2], state[3]);
    return 0;
}
