#pragma once
#include <cfloat>
