// TEST INFRASTRUCTURE ONLY (oracle/_ref build): the few roscpp names the two compiled reference sources
// mention (NodeHandle parameter of GroundSegmentation::init, Time, logging macros).  Logging is dropped.
#pragma once
#include <cstdint>
#include <memory>
#include <sstream>
#include <string>

namespace ros {
struct Time {
    uint32_t sec = 0, nsec = 0;
    Time() = default;
    Time(uint32_t s, uint32_t n) : sec(s), nsec(n) {}
};
struct Duration {
    double s = 0.0;
    explicit Duration(double v = 0.0) : s(v) {}
};
class NodeHandle {};
}  // namespace ros

#define GGR_SHIM_NOLOG(...) do { } while (0)
#define ROS_DEBUG_STREAM(x) GGR_SHIM_NOLOG()
#define ROS_INFO_STREAM(x) GGR_SHIM_NOLOG()
#define ROS_WARN_STREAM(x) GGR_SHIM_NOLOG()
#define ROS_DEBUG(...) GGR_SHIM_NOLOG()
#define ROS_INFO(...) GGR_SHIM_NOLOG()
#define ROS_WARN(...) GGR_SHIM_NOLOG()
