// GPU-only fixture for SortArchetypeNode with a custom key (4 radix passes, no
// truncation, no world grouping: src/mw/device/sort_archetype.cpp:1431-1440).
// The reference CPU backend's sortArchetype applies the INVERSE permutation
// (SURVEY.md F8), so this path is checked against the numpy restatement
// (oracle/restate.py: stable sort by key) instead of the CPU backend.
#pragma once

#include <madrona/taskgraph_builder.hpp>
#include <madrona/custom_context.hpp>
#include <madrona/rand.hpp>

namespace sortcheck {

using madrona::Entity;

enum class ExportID : uint32_t { Key, Payload, Tag, Step, NumExports };
enum class TaskGraphID : uint32_t { Step, NumTaskGraphs };

struct SortKey { uint32_t v; };
struct Payload { uint32_t world; uint32_t item; float a; float b; };   // 16 bytes
struct Tag { uint8_t bytes[6]; };                                        // odd size
struct StepCounter { uint32_t t; };

struct Item : public madrona::Archetype<SortKey, Payload, Tag> {};

struct Config { uint32_t itemsPerWorld; uint32_t keyMask; };
struct WorldInit { uint32_t seed; };

class Engine;

struct Sim : public madrona::WorldBase {
    static void registerTypes(madrona::ECSRegistry &registry, const Config &cfg);
    static void setupTasks(madrona::TaskGraphManager &mgr, const Config &cfg);
    Sim(Engine &ctx, const Config &cfg, const WorldInit &init);
    uint32_t seed;
    uint32_t keyMask;
};

class Engine : public madrona::CustomContext<Engine, Sim> {
public:
    using CustomContext::CustomContext;
};

}
