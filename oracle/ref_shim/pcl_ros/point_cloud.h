// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
#pragma once
#include <pcl/point_cloud.h>
