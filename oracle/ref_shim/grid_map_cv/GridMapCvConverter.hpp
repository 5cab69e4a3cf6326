// TEST INFRASTRUCTURE ONLY (oracle/_ref build): included by the reference headers, unused by the two compiled sources.
#pragma once
#include <grid_map_core/GridMap.hpp>
