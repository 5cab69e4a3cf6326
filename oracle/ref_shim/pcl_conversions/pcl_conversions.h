// TEST INFRASTRUCTURE ONLY (oracle/_ref build): included by GroundSegmentation.h:35, unused by the two compiled sources.
#pragma once
#include <pcl/point_cloud.h>
