// TEST INFRASTRUCTURE ONLY (oracle/_ref build): included by GroundSegmentation.h:29, unused by the two compiled sources.
#pragma once
#include <std_msgs/Header.h>
