#pragma once
#include <cstddef>
