#include "sim.hpp"
#include <madrona/mw_gpu_entry.hpp>

using namespace madrona;

namespace sortcheck {

// key of (world, item) at step t: a pure function, so the test can predict it
static inline uint32_t keyOf(uint32_t seed, uint32_t item, uint32_t t, uint32_t mask)
{
    RandKey k = rand::split_i(rand::initKey(seed), item, t);
    return rand::bits32(k) & mask;
}

void Sim::registerTypes(ECSRegistry &registry, const Config &)
{
    registry.registerComponent<SortKey>();
    registry.registerComponent<Payload>();
    registry.registerComponent<Tag>();
    registry.registerSingleton<StepCounter>();
    registry.registerArchetype<Item>();
    registry.exportColumn<Item, SortKey>((uint32_t)ExportID::Key);
    registry.exportColumn<Item, Payload>((uint32_t)ExportID::Payload);
    registry.exportColumn<Item, Tag>((uint32_t)ExportID::Tag);
    registry.exportSingleton<StepCounter>((uint32_t)ExportID::Step);
}

inline void tickSystem(Engine &, StepCounter &c)
{
    c.t += 1;
}

inline void rekeySystem(Engine &ctx, SortKey &key, Payload &p)
{
    uint32_t t = ctx.singleton<StepCounter>().t;
    key.v = keyOf(ctx.data().seed, p.item, t, ctx.data().keyMask);
}

void Sim::setupTasks(TaskGraphManager &mgr, const Config &)
{
    TaskGraphBuilder &builder = mgr.init(TaskGraphID::Step);
    auto tick = builder.addToGraph<ParallelForNode<Engine, tickSystem, StepCounter>>({});
    auto rekey = builder.addToGraph<ParallelForNode<Engine, rekeySystem, SortKey, Payload>>({tick});
    builder.addToGraph<SortArchetypeNode<Item, SortKey>>({rekey});
}

Sim::Sim(Engine &ctx, const Config &cfg, const WorldInit &init)
    : WorldBase(ctx), seed(init.seed), keyMask(cfg.keyMask)
{
    ctx.singleton<StepCounter>().t = 0;
    for (uint32_t i = 0; i < cfg.itemsPerWorld; i++) {
        Entity e = ctx.makeEntity<Item>();
        ctx.get<SortKey>(e).v = keyOf(seed, i, 0, keyMask);
        ctx.get<Payload>(e) = Payload { (uint32_t)ctx.worldID().idx, i, (float)i * 0.5f, (float)seed };
        Tag tag;
        for (int b = 0; b < 6; b++) tag.bytes[b] = (uint8_t)(i * 7 + b + seed);
        ctx.get<Tag>(e) = tag;
    }
}

}

MADRONA_BUILD_MWGPU_ENTRY(sortcheck::Engine, sortcheck::Sim, sortcheck::Config, sortcheck::WorldInit);
