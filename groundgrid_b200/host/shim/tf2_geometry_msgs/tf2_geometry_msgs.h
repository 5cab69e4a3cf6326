#pragma once
#include <tf2_ros/transform_listener.h>
