// TEST INFRASTRUCTURE ONLY: sims/cartpole on the reference CPU backend.
#include <madrona/mw_cpu.hpp>
#include "../sims/cartpole/sim.hpp"
#include "harness.hpp"

using namespace cartpole;

int main(int argc, char **argv)
{
    oracle::Args args = oracle::parseArgs(argc, argv);
    Config cfg { (uint32_t)(args.extra[0] ? args.extra[0] : 200) };
    std::vector<WorldInit> inits(args.numWorlds);
    for (int64_t i = 0; i < args.numWorlds; i++) inits[i].seed = (uint32_t)(args.extra[1] + i);

    using Exec = madrona::TaskGraphExecutor<Engine, Sim, Config, WorldInit>;
    Exec exec({
        .numWorlds = (uint32_t)args.numWorlds,
        .numExportedBuffers = (uint32_t)ExportID::NumExports,
        .numWorkers = (uint32_t)args.numWorkers,
    }, cfg, inits.data(), (madrona::CountT)TaskGraphID::NumTaskGraphs);

    size_t W = (size_t)args.numWorlds;
    return oracle::runTrace(exec, args,
        { { (int)ExportID::Reset, 4 }, { (int)ExportID::Action, 4 } },
        { { (int)ExportID::State, [=] { return W * 16; } },
          { (int)ExportID::Reward, [=] { return W * 4; } },
          { (int)ExportID::Done, [=] { return W * 4; } } });
}
