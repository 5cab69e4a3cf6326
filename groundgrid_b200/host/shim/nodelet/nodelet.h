#pragma once
#include <ros/ros.h>
namespace nodelet {
class Nodelet {
  public:
    virtual ~Nodelet() {}
    void init() { onInit(); }
  protected:
    virtual void onInit() = 0;
    ros::NodeHandle& getNodeHandle() { return nh_; }
    ros::NodeHandle& getPrivateNodeHandle() { return pnh_; }
  private:
    ros::NodeHandle nh_, pnh_;
};
}  // namespace nodelet
