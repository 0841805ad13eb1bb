#pragma once
#include <cassert>
