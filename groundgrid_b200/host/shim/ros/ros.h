// Minimal stand-in for <ros/ros.h> so that the GroundGrid class surface compiles without ROS
// (ROS is not installable in this image).  On a ROS machine drop this directory from the include
// path and the real headers take over; nothing in the hot path depends on these stubs.
#pragma once
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

namespace ros {
struct Time {
    uint32_t sec = 0, nsec = 0;
    Time() = default;
    Time(uint32_t s, uint32_t n) : sec(s), nsec(n) {}
    static Time now() { return Time(); }
    bool operator==(const Time& o) const { return sec == o.sec && nsec == o.nsec; }
};
struct Duration {
    double s = 0.0;
    explicit Duration(double v = 0.0) : s(v) {}
};
class NodeHandle {};
inline int& gg_shim_log_level() {
    static int level = 1;  // 0 debug, 1 info, 2 warn
    return level;
}
}  // namespace ros

#define GG_SHIM_LOG(lvl, tag, expr)                                              \
    do {                                                                         \
        if (ros::gg_shim_log_level() <= lvl) {                                   \
            std::ostringstream gg_os__;                                          \
            gg_os__ << expr;                                                     \
            std::fprintf(stderr, "[%s] %s\n", tag, gg_os__.str().c_str());       \
        }                                                                        \
    } while (0)
#define ROS_DEBUG_STREAM(x) GG_SHIM_LOG(0, "DEBUG", x)
#define ROS_INFO_STREAM(x) GG_SHIM_LOG(1, "INFO", x)
#define ROS_WARN_STREAM(x) GG_SHIM_LOG(2, "WARN", x)
#define GG_SHIM_LOGF(lvl, tag, ...)                      \
    do {                                                 \
        if (ros::gg_shim_log_level() <= lvl) {           \
            std::fprintf(stderr, "[%s] ", tag);          \
            std::fprintf(stderr, __VA_ARGS__);           \
            std::fprintf(stderr, "\n");                  \
        }                                                \
    } while (0)
#define ROS_DEBUG(...) GG_SHIM_LOGF(0, "DEBUG", __VA_ARGS__)
#define ROS_INFO(...) GG_SHIM_LOGF(1, "INFO", __VA_ARGS__)
#define ROS_WARN(...) GG_SHIM_LOGF(2, "WARN", __VA_ARGS__)
