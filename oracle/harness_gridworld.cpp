// TEST INFRASTRUCTURE ONLY: sims/gridworld on the reference CPU backend.
#include <madrona/mw_cpu.hpp>
#include "../sims/gridworld/sim.hpp"
#include "harness.hpp"

using namespace gridworld;

int main(int argc, char **argv)
{
    oracle::Args args = oracle::parseArgs(argc, argv);
    Config cfg {
        (int32_t)(args.extra[0] ? args.extra[0] : 8),
        (int32_t)(args.extra[1] ? args.extra[1] : 50),
        (int32_t)(args.extra[2] ? args.extra[2] : 6),
    };
    std::vector<WorldInit> inits(args.numWorlds);
    for (int64_t i = 0; i < args.numWorlds; i++) inits[i].seed = (uint32_t)(args.extra[3] + i);

    using Exec = madrona::TaskGraphExecutor<Engine, Sim, Config, WorldInit>;
    Exec exec({
        .numWorlds = (uint32_t)args.numWorlds,
        .numExportedBuffers = (uint32_t)ExportID::NumExports,
        .numWorkers = (uint32_t)args.numWorkers,
    }, cfg, inits.data(), (madrona::CountT)TaskGraphID::NumTaskGraphs);

    size_t W = (size_t)args.numWorlds;
    auto total_items = [&exec, W]() {
        const int32_t *counts = (const int32_t *)exec.getExported((int)ExportID::ItemCount);
        size_t n = 0;
        for (size_t i = 0; i < W; i++) n += (size_t)counts[i];
        return n;
    };
    return oracle::runTrace(exec, args,
        { { (int)ExportID::Reset, 4 }, { (int)ExportID::Action, 4 * kNumAgents } },
        { { (int)ExportID::AgentPos, [=] { return W * kNumAgents * 8; } },
          { (int)ExportID::Reward, [=] { return W * kNumAgents * 4; } },
          { (int)ExportID::Obs, [=] { return W * kNumAgents * 16; } },
          { (int)ExportID::ItemCount, [=] { return W * 4; } },
          { (int)ExportID::Done, [=] { return W * 4; } },
          { (int)ExportID::ItemPos, [=] { return total_items() * 8; } },
          { (int)ExportID::ItemEntity, [=] { return total_items() * 8; } },
          { (int)ExportID::ItemKind, [=] { return total_items() * 4; } } });
}
