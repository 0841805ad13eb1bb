#pragma once
#include <madrona/state.hpp>
