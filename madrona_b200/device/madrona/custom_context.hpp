// Reference: include/madrona/custom_context.hpp:13-29.
#pragma once
#include <madrona/context.hpp>
namespace madrona {
template <typename ContextT, typename DataT>
class CustomContext : public Context {
public:
    inline CustomContext(DataT *world_data, const WorkerInit &worker_init)
        : Context(world_data, worker_init) {}
    inline DataT &data() const { return *static_cast<DataT *>(data_); }
private:
    using WorldDataT = DataT;
};
}
