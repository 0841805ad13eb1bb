#pragma once
#include <cstdio>
