#pragma once
#include <cstdint>
