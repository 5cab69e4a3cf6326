// TEST INFRASTRUCTURE ONLY (oracle/_ref build): grid_map_ros pulls in grid_map_core; the converters are not used by the two compiled sources.
#pragma once
#include <grid_map_core/GridMap.hpp>
