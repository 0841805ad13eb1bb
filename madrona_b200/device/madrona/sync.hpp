#pragma once
#include <madrona/types.hpp>
