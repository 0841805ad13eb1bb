// Context: the per-invocation handle simulator systems receive.  Same surface
// as the reference (include/madrona/context.hpp:24-139; GPU flavour
// src/mw/device/include/madrona/context.{hpp,inl}).
#pragma once
#include <madrona/fwd.hpp>
#include <madrona/ecs.hpp>
#include <madrona/state.hpp>
#include <madrona/registry.hpp>

namespace madrona {

struct WorkerInit {
    WorldID worldID;
};

class Context {
public:
    inline Context(WorldBase *world_data, const WorkerInit &init)
        : data_(world_data), world_id_(init.worldID) {}
    Context(const Context &) = delete;

    template <typename ArchetypeT>
    inline Entity makeEntity() { return makeEntity(TypeTracker::typeID<ArchetypeT>()); }
    inline Entity makeEntity(uint32_t archetype_id);

    template <typename ArchetypeT>
    inline Loc makeTemporary() { return makeTemporary(TypeTracker::typeID<ArchetypeT>()); }
    inline Loc makeTemporary(uint32_t archetype_id);

    inline void destroyEntity(Entity e);

    inline Loc loc(Entity e) const { return mwGPU::lookupLoc(mwGPU::engine(), e); }

    template <typename ComponentT>
    inline ComponentT &get(Entity e) { return get<ComponentT>(loc(e)); }

    template <typename ComponentT>
    inline ComponentT &get(Loc l)
    {
        mb2::EngineState &S = mwGPU::engine();
        int32_t col = S.columnLookup[l.archetype][TypeTracker::typeID<ComponentT>()];
        return ((ComponentT *)S.tables[l.archetype].columns[col])[l.row];
    }

    template <typename ComponentT>
    inline ResultRef<ComponentT> getSafe(Entity e) { return getCheck<ComponentT>(e); }

    template <typename ComponentT>
    inline ResultRef<ComponentT> getCheck(Entity e)
    {
        Loc l = loc(e);
        if (!l.valid()) return ResultRef<ComponentT>(nullptr);
        return getCheck<ComponentT>(l);
    }

    template <typename ComponentT>
    inline ResultRef<ComponentT> getCheck(Loc l)
    {
        mb2::EngineState &S = mwGPU::engine();
        int32_t col = S.columnLookup[l.archetype][TypeTracker::typeID<ComponentT>()];
        if (col < 0) return ResultRef<ComponentT>(nullptr);
        return ResultRef<ComponentT>(
            (ComponentT *)S.tables[l.archetype].columns[col] + l.row);
    }

    template <typename ComponentT>
    inline ComponentT &getDirect(int32_t column_idx, Loc l)
    {
        return ((ComponentT *)mwGPU::engine().tables[l.archetype].columns[column_idx])[l.row];
    }

    template <typename SingletonT>
    inline SingletonT &singleton()
    {
        mb2::EngineState &S = mwGPU::engine();
        uint32_t a = TypeTracker::typeID<SingletonArchetype<SingletonT>>();
        return ((SingletonT *)S.tables[a].columns[2])[world_id_.idx];
    }

    inline void *tmpAlloc(uint64_t num_bytes);

    template <typename... ComponentTs>
    inline Query<ComponentTs...> query()
    {
        auto &q = Query<ComponentTs...>::data();
        if (((volatile int32_t *)&q.resolved)[0] == 0) {
            // idempotent: every racing thread computes the same table
            mwGPU::getStateManager()->template resolveQuery<ComponentTs...>(q);
        }
        return Query<ComponentTs...>();
    }

    // Iterate this world's rows of every archetype matching the query.
    // Needs world-sorted tables (offsets/counts), like the reference
    // (src/mw/device/include/madrona/state.inl:180-252).
    template <typename... ComponentTs, typename Fn>
    inline void iterateQuery(const Query<ComponentTs...> &, Fn &&fn)
    {
        constexpr int N = (int)sizeof...(ComponentTs);
        auto &q = Query<ComponentTs...>::data();
        mb2::EngineState &S = mwGPU::engine();
        for (int qa = 0; qa < q.numArchetypes; qa++) {
            mb2::TableDesc &t = S.tables[q.archetypes[qa]];
            int32_t off = t.worldOffsets[world_id_.idx];
            int32_t cnt = t.worldCounts[world_id_.idx];
            const WorldID *wcol = (const WorldID *)t.columns[1];
            for (int32_t r = off; r < off + cnt; r++) {
                if (wcol[r].idx < 0) continue;
                iterateCall<ComponentTs...>(fn, t, q.cols[qa], r,
                    mwGPU::IntSeq<N> {});
            }
        }
    }

    inline WorldID worldID() const { return world_id_; }
    inline WorldBase &data() const { return *data_; }
    inline StateManager *getStateManager() { return mwGPU::getStateManager(); }

protected:
    WorldBase *data_;

private:
    template <typename... ComponentTs, typename Fn, int... Is>
    inline void iterateCall(Fn &fn, mb2::TableDesc &t, const int32_t *cols,
                            int32_t r, mwGPU::IntList<Is...>)
    {
        fn(((ComponentTs *)t.columns[cols[Is]])[r]...);
    }

    WorldID world_id_;
};

Entity Context::makeEntity(uint32_t archetype_id)
{
    mb2::EngineState &S = mwGPU::engine();
    mb2::TableDesc &tbl = S.tables[archetype_id];

    mb2::IDCache &cache = S.idCaches[world_id_.idx];
    mwGPU::lockCache(cache);
    Entity e = mwGPU::acquireEntityLocked(S, cache);
    mwGPU::unlockCache(cache);

    int32_t row = mwGPU::appendRow(tbl, archetype_id);
    if (row < 0 || e.id < 0) return Entity::none();

    ((Entity *)tbl.columns[0])[row] = e;
    ((WorldID *)tbl.columns[1])[row] = world_id_;
    S.entitySlots[e.id].a = (int32_t)archetype_id;
    S.entitySlots[e.id].b = row;
    return e;
}

Loc Context::makeTemporary(uint32_t archetype_id)
{
    mb2::EngineState &S = mwGPU::engine();
    mb2::TableDesc &tbl = S.tables[archetype_id];
    int32_t row = mwGPU::appendRow(tbl, archetype_id);
    if (row < 0) return Loc::none();
    // CPU backend writes Entity{0,0} for temporaries (state.inl:573-574)
    ((Entity *)tbl.columns[0])[row] = Entity { 0, 0 };
    ((WorldID *)tbl.columns[1])[row] = world_id_;
    return Loc { archetype_id, row };
}

void Context::destroyEntity(Entity e)
{
    mb2::EngineState &S = mwGPU::engine();
    Loc l = mwGPU::lookupLoc(S, e);
    if (!l.valid()) return;
    mb2::TableDesc &tbl = S.tables[l.archetype];
    ((Entity *)tbl.columns[0])[l.row] = Entity::none();
    ((WorldID *)tbl.columns[1])[l.row] = WorldID { -1 };
    tbl.needsSort = 1;

    mb2::IDCache &cache = S.idCaches[world_id_.idx];
    mwGPU::lockCache(cache);
    mwGPU::releaseEntityLocked(S, cache, e.id);
    mwGPU::unlockCache(cache);
}

void *Context::tmpAlloc(uint64_t num_bytes)
{
    mb2::EngineState &S = mwGPU::engine();
    num_bytes = (num_bytes + 255ull) & ~255ull;
    unsigned long long off = atomicAdd((unsigned long long *)&S.tmpOffset,
                                       (unsigned long long)num_bytes);
    if (off + num_bytes > S.tmpCapacity) {
        mwGPU::raiseError(mb2::ErrTmpOverflow);
        return nullptr;
    }
    return S.tmpArena + off;
}

}
