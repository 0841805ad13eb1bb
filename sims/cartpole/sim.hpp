// Fixture simulator 1 ("Cartpole-like", BASELINE.json configs[0]).
// Written against the public Madrona API only; the SAME sources are compiled
//   (a) by g++ against the reference headers + CPU backend  -> oracle trace
//   (b) by NVRTC against madrona_b200/device              -> B200 engine
// All arithmetic is + - * / (sin/cos are replaced by fixed polynomials that
// are part of this simulator's definition), so with FP contraction disabled on
// both sides the two backends must agree bit-for-bit.
#pragma once

#include <madrona/taskgraph_builder.hpp>
#include <madrona/custom_context.hpp>
#include <madrona/rand.hpp>

namespace cartpole {

using madrona::Entity;
using madrona::CountT;

enum class ExportID : uint32_t {
    Reset,
    Action,
    State,
    Reward,
    Done,
    NumExports,
};

enum class TaskGraphID : uint32_t {
    Step,
    NumTaskGraphs,
};

struct WorldReset {
    int32_t reset;
};

struct Action {
    int32_t push;   // 0 = left, 1 = right
};

struct CartState {
    float x;
    float xDot;
    float theta;
    float thetaDot;
};

struct Reward {
    float v;
};

struct Done {
    int32_t v;
};

struct StepCount {
    uint32_t t;
};

struct Cart : public madrona::Archetype<
    Action, CartState, Reward, Done, StepCount
> {};

struct Config {
    uint32_t maxSteps;
};

struct WorldInit {
    uint32_t seed;
};

class Engine;

struct Sim : public madrona::WorldBase {
    static void registerTypes(madrona::ECSRegistry &registry,
                              const Config &cfg);
    static void setupTasks(madrona::TaskGraphManager &mgr,
                           const Config &cfg);

    Sim(Engine &ctx, const Config &cfg, const WorldInit &init);

    madrona::RNG rng;
    Entity cart;
    uint32_t maxSteps;
    uint32_t episode;
};

class Engine : public madrona::CustomContext<Engine, Sim> {
public:
    using CustomContext::CustomContext;
};

}
