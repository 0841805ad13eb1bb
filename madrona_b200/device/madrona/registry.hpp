// ECSRegistry, same surface as the reference (include/madrona/registry.hpp:25-70).
// Runs on the device inside the 1-thread initECS kernel and only records
// metadata; the host allocates the tables afterwards.
#pragma once
#include <madrona/state.hpp>
namespace madrona {

class ECSRegistry {
public:
    inline ECSRegistry(StateManager *state_mgr, void **export_ptrs)
        : state_mgr_(state_mgr), export_ptrs_(export_ptrs) {}

    template <typename ComponentT>
    void registerComponent(uint32_t num_bytes = 0)
    {
        state_mgr_->registerComponent<ComponentT>(num_bytes);
    }

    template <typename ArchetypeT>
    void registerArchetype()
    {
        state_mgr_->registerArchetype<ArchetypeT>(
            ComponentMetadataSelector<> {}, ArchetypeFlags::None, 0);
    }

    template <typename ArchetypeT, typename... MetadataComponentTs>
    void registerArchetype(
        ComponentMetadataSelector<MetadataComponentTs...> component_metadatas,
        ArchetypeFlags archetype_flags,
        CountT max_num_entities_per_world = 0)
    {
        state_mgr_->registerArchetype<ArchetypeT>(
            component_metadatas, archetype_flags, max_num_entities_per_world);
    }

    template <typename BundleT>
    void registerBundle() { state_mgr_->registerBundle<BundleT>(); }

    template <typename AliasT, typename BundleT>
    void registerBundleAlias() { state_mgr_->registerBundleAlias<AliasT, BundleT>(); }

    template <typename SingletonT>
    void registerSingleton() { state_mgr_->registerSingleton<SingletonT>(); }

    template <typename ArchetypeT, typename ComponentT>
    void exportColumn(int32_t slot)
    {
        recordExport(slot, TypeTracker::typeID<ArchetypeT>(),
                     TypeTracker::typeID<ComponentT>());
    }

    template <typename SingletonT>
    void exportSingleton(int32_t slot)
    {
        exportColumn<SingletonArchetype<SingletonT>, SingletonT>(slot);
    }

    template <typename ArchetypeT, typename ComponentT, EnumType EnumT>
    void exportColumn(EnumT slot) { exportColumn<ArchetypeT, ComponentT>((int32_t)slot); }

    template <typename SingletonT, EnumType EnumT>
    void exportSingleton(EnumT slot) { exportSingleton<SingletonT>((int32_t)slot); }

    inline StateManager *stateManager() { return state_mgr_; }

private:
    inline void recordExport(int32_t slot, uint32_t archetype, uint32_t component)
    {
        mb2::EngineState &S = mwGPU::engine();
        if (slot < 0 || slot >= mb2::kMaxExports || (uint32_t)slot >= S.numExported) {
            mwGPU::raiseError(mb2::ErrRegistry);
            return;
        }
        S.exports[slot].used = 1;
        S.exports[slot].archetype = archetype;
        S.exports[slot].component = component;
    }

    StateManager *state_mgr_;
    void **export_ptrs_;
};

}
