#pragma once
namespace madrona {
class Context;
class StateManager;
class ECSRegistry;
class TaskGraphBuilder;
class TaskGraphManager;
struct WorkerInit;
}
