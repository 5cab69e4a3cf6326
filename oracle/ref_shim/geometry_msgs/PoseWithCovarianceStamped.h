// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
#pragma once
#include <geometry_msgs/msgs.h>
