// TEST INFRASTRUCTURE ONLY -- see GridMap.hpp in this directory (the math lives there).
#pragma once
#include <grid_map_core/GridMap.hpp>
