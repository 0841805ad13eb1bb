#pragma once
#include <cstring>
