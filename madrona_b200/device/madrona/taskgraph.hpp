#pragma once
#include <madrona/taskgraph_builder.hpp>
