// TEST INFRASTRUCTURE ONLY (oracle/_ref build): included by GroundGrid.cpp:31, nothing of it is used.
#pragma once
#include <ros/ros.h>
