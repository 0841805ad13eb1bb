#include "sim.hpp"

#ifdef MADRONA_GPU_MODE
#include <madrona/mw_gpu_entry.hpp>
#endif

using namespace madrona;

namespace gridworld {

void Sim::registerTypes(ECSRegistry &registry, const Config &)
{
    registry.registerComponent<Action>();
    registry.registerComponent<GridPos>();
    registry.registerComponent<Reward>();
    registry.registerComponent<AgentID>();
    registry.registerComponent<ItemKind>();
    registry.registerComponent<Obs>();

    registry.registerSingleton<WorldReset>();
    registry.registerSingleton<ItemCount>();
    registry.registerSingleton<Done>();

    registry.registerArchetype<Agent>(
        ComponentMetadataSelector<> {}, ArchetypeFlags::None, kNumAgents);
    registry.registerArchetype<Item>();

    registry.exportSingleton<WorldReset>((uint32_t)ExportID::Reset);
    registry.exportColumn<Agent, Action>((uint32_t)ExportID::Action);
    registry.exportColumn<Agent, GridPos>((uint32_t)ExportID::AgentPos);
    registry.exportColumn<Agent, Reward>((uint32_t)ExportID::Reward);
    registry.exportColumn<Agent, Obs>((uint32_t)ExportID::Obs);
    registry.exportSingleton<ItemCount>((uint32_t)ExportID::ItemCount);
    registry.exportColumn<Item, GridPos>((uint32_t)ExportID::ItemPos);
    registry.exportColumn<Item, Entity>((uint32_t)ExportID::ItemEntity);
    registry.exportColumn<Item, ItemKind>((uint32_t)ExportID::ItemKind);
    registry.exportSingleton<Done>((uint32_t)ExportID::Done);
}

static inline void spawnItem(Engine &ctx)
{
    Sim &sim = ctx.data();
    if (sim.numItems >= kMaxItems) {
        return;
    }

    Entity item = ctx.makeEntity<Item>();
    ctx.get<GridPos>(item) = GridPos {
        sim.rng.sampleI32(0, sim.gridSize),
        sim.rng.sampleI32(0, sim.gridSize),
    };
    ctx.get<ItemKind>(item).kind = sim.rng.sampleI32(0, 3);
    sim.items[sim.numItems++] = item;
}

static inline void placeAgents(Engine &ctx)
{
    Sim &sim = ctx.data();
    for (int32_t i = 0; i < kNumAgents; i++) {
        ctx.get<GridPos>(sim.agents[i]) = GridPos {
            sim.rng.sampleI32(0, sim.gridSize),
            sim.rng.sampleI32(0, sim.gridSize),
        };
    }
}

static inline void resetEpisode(Engine &ctx)
{
    Sim &sim = ctx.data();
    for (int32_t i = 0; i < sim.numItems; i++) {
        ctx.destroyEntity(sim.items[i]);
    }
    sim.numItems = 0;
    sim.curStep = 0;

    placeAgents(ctx);
    for (int32_t i = 0; i < sim.initItems; i++) {
        spawnItem(ctx);
    }
}

inline void moveSystem(Engine &ctx, GridPos &pos, Action &action, Reward &reward)
{
    const int32_t n = ctx.data().gridSize;
    switch (action.move) {
    case 1: pos.x = pos.x + 1 < n ? pos.x + 1 : pos.x; break;
    case 2: pos.x = pos.x > 0 ? pos.x - 1 : pos.x; break;
    case 3: pos.y = pos.y + 1 < n ? pos.y + 1 : pos.y; break;
    case 4: pos.y = pos.y > 0 ? pos.y - 1 : pos.y; break;
    default: break;
    }
    reward.v = 0.f;
}

// One invocation per world: pickups (agents in order), spawning, episode reset.
inline void worldSystem(Engine &ctx, WorldReset &reset)
{
    Sim &sim = ctx.data();

    for (int32_t a = 0; a < kNumAgents; a++) {
        GridPos apos = ctx.get<GridPos>(sim.agents[a]);
        int32_t i = 0;
        while (i < sim.numItems) {
            Entity item = sim.items[i];
            GridPos ipos = ctx.get<GridPos>(item);
            if (ipos.x == apos.x && ipos.y == apos.y) {
                float value = 1.f + (float)ctx.get<ItemKind>(item).kind;
                ctx.get<Reward>(sim.agents[a]).v += value;
                ctx.destroyEntity(item);
                sim.items[i] = sim.items[sim.numItems - 1];
                sim.numItems -= 1;
            } else {
                i += 1;
            }
        }
    }

    if (sim.rng.sampleI32(0, 4) == 0) {
        spawnItem(ctx);
    }

    sim.curStep += 1;
    int32_t done = 0;
    if (reset.reset != 0 || sim.curStep >= sim.episodeLen) {
        resetEpisode(ctx);
        reset.reset = 0;
        done = 1;
    }
    ctx.singleton<Done>().v = done;
    ctx.singleton<ItemCount>().count = sim.numItems;
}

// Per agent, after the Item table is compacted/sorted: scan this world's items.
inline void obsSystem(Engine &ctx, GridPos &pos, Obs &obs)
{
    int32_t best = 0x7fffffff;
    Obs out { 0, 0, 0, 0 };
    auto q = ctx.query<GridPos, ItemKind>();
    ctx.iterateQuery(q, [&](GridPos &ipos, ItemKind &kind) {
        int32_t dx = ipos.x - pos.x;
        int32_t dy = ipos.y - pos.y;
        int32_t dist = (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy);
        if (dist < best) {
            best = dist;
            out.dx = dx;
            out.dy = dy;
        }
        out.numItems += 1;
        if (kind.kind == 0) {
            out.numKind0 += 1;
        }
    });
    obs = out;
}

void Sim::setupTasks(TaskGraphManager &mgr, const Config &)
{
    TaskGraphBuilder &builder = mgr.init(TaskGraphID::Step);

    auto move = builder.addToGraph<ParallelForNode<Engine, moveSystem,
        GridPos, Action, Reward>>({});
    auto world = builder.addToGraph<ParallelForNode<Engine, worldSystem,
        WorldReset>>({move});
    // removes destroyed rows; on the GPU backend this is the WorldID sort
    auto compact = builder.addToGraph<CompactArchetypeNode<Item>>({world});
#ifdef MADRONA_GPU_MODE
    auto recycle = builder.addToGraph<RecycleEntitiesNode>({compact});
    auto obs_dep = recycle;
#else
    auto obs_dep = compact;
#endif
    builder.addToGraph<ParallelForNode<Engine, obsSystem,
        GridPos, Obs>>({obs_dep});
}

Sim::Sim(Engine &ctx, const Config &cfg, const WorldInit &init)
    : WorldBase(ctx),
      rng(init.seed),
      numItems(0),
      gridSize(cfg.gridSize),
      episodeLen(cfg.episodeLen),
      initItems(cfg.initItems),
      curStep(0)
{
    for (int32_t i = 0; i < kNumAgents; i++) {
        agents[i] = ctx.makeEntity<Agent>();
        ctx.get<Action>(agents[i]).move = 0;
        ctx.get<Reward>(agents[i]).v = 0.f;
        ctx.get<AgentID>(agents[i]).idx = i;
        ctx.get<Obs>(agents[i]) = Obs { 0, 0, 0, 0 };
    }
    ctx.singleton<WorldReset>().reset = 0;
    resetEpisode(ctx);
    ctx.singleton<Done>().v = 0;
    ctx.singleton<ItemCount>().count = numItems;
}

}

#ifdef MADRONA_GPU_MODE
MADRONA_BUILD_MWGPU_ENTRY(gridworld::Engine, gridworld::Sim,
                          gridworld::Config, gridworld::WorldInit);
#endif
