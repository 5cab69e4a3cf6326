#pragma once
#include <geometry_msgs/msgs.h>
