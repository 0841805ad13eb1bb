// Fixture simulator 2 ("pure-ECS grid sim", BASELINE.json configs[4]):
// no physics; entity create/destroy every few steps, compaction + world sort
// every step, Context::get by Entity, per-world queries.  Integer state only,
// so every exported column (including Entity IDs and row order of the dynamic
// Item table) must match the reference CPU backend bit-for-bit.
#pragma once

#include <madrona/taskgraph_builder.hpp>
#include <madrona/custom_context.hpp>
#include <madrona/rand.hpp>

namespace gridworld {

using madrona::Entity;
using madrona::CountT;

constexpr int32_t kNumAgents = 2;
constexpr int32_t kMaxItems = 24;

enum class ExportID : uint32_t {
    Reset,
    Action,
    AgentPos,
    Reward,
    Obs,
    ItemCount,
    ItemPos,
    ItemEntity,
    ItemKind,
    Done,
    NumExports,
};

enum class TaskGraphID : uint32_t {
    Step,
    NumTaskGraphs,
};

struct WorldReset { int32_t reset; };
struct ItemCount { int32_t count; };
struct Done { int32_t v; };

struct Action { int32_t move; };         // 0 stay, 1 +x, 2 -x, 3 +y, 4 -y
struct GridPos { int32_t x; int32_t y; };
struct Reward { float v; };
struct AgentID { int32_t idx; };
struct ItemKind { int32_t kind; };
// nearest item (dx, dy), number of items in the world, items seen of kind 0
struct Obs { int32_t dx; int32_t dy; int32_t numItems; int32_t numKind0; };

struct Agent : public madrona::Archetype<
    GridPos, Action, Reward, AgentID, Obs
> {};

struct Item : public madrona::Archetype<
    GridPos, ItemKind
> {};

struct Config {
    int32_t gridSize;
    int32_t episodeLen;
    int32_t initItems;
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
    Entity agents[kNumAgents];
    Entity items[kMaxItems];
    int32_t numItems;
    int32_t gridSize;
    int32_t episodeLen;
    int32_t initItems;
    int32_t curStep;
};

class Engine : public madrona::CustomContext<Engine, Sim> {
public:
    using CustomContext::CustomContext;
};

}
