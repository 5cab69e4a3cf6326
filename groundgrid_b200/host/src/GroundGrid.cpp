// groundgrid::GroundGrid over the groundgrid_b200 C-ABI.
#include <groundgrid/GroundGrid.h>

#include <stdexcept>

using namespace groundgrid;

GroundGrid::GroundGrid() : mTf2_listener(mTfBuffer) {}
GroundGrid::~GroundGrid() {}

void GroundGrid::setConfig(groundgrid::GroundGridConfig& config) { config_ = config; }

void GroundGrid::setGeometryOverride(float dimension_m, float resolution, int device, size_t max_points) {
    dim_override_ = dimension_m;
    res_override_ = resolution;
    device_ = device;
    max_points_ = max_points;
}

void GroundGrid::initGroundGrid(const nav_msgs::OdometryConstPtr& inOdom) {
    const float dim = dim_override_ > 0.f ? dim_override_ : mDimension;
    const float res = res_override_ > 0.f ? res_override_ : mResolution;
    mMap_ptr = std::make_shared<grid_map::GridMap>(
        std::vector<std::string>{"points", "ground", "groundpatch", "minGroundHeight", "maxGroundHeight"});
    grid_map::GridMap& map = *mMap_ptr;
    map.setFrameId("map");
    map.setDeviceOptions(device_, max_points_);
    // allocates the device layers; the layer fills of the reference (points 0, ground z, groundpatch 1e-7, min 100, max -100)
    // happen in gg_init_map
    map.setGeometry(grid_map::Length(dim, dim), res, grid_map::Position(inOdom->pose.pose.position.x, inOdom->pose.pose.position.y));
    ROS_INFO("Created map with size %f x %f m (%i x %i cells).", map.getLength().x(), map.getLength().y(), map.getSize()(0), map.getSize()(1));
    if (gg_init_map(map.handle(), map.slot(), inOdom->pose.pose.position.x, inOdom->pose.pose.position.y, inOdom->pose.pose.position.z) != GG_OK)
        throw std::runtime_error(std::string("groundgrid_b200: ") + gg_last_error());
    mLastPose.pose = inOdom->pose;
    mLastPose.header = inOdom->header;
}

std::shared_ptr<grid_map::GridMap> GroundGrid::update(const nav_msgs::OdometryConstPtr& inOdom) {
    if (!mMap_ptr) {
        initGroundGrid(inOdom);
        return mMap_ptr;
    }
    grid_map::GridMap& map = *mMap_ptr;
    try {
        mBaseToMap = mTfBuffer.lookupTransform("base_link", "map", inOdom->header.stamp);
    } catch (tf2::LookupException& e) {
        ROS_WARN("no transform? -> error: %s", e.what());  // potentially degraded performance: the last transform is used
    } catch (tf2::ExtrapolationException& e) {
        ROS_DEBUG("need to extrapolate a transform? -> error: %s", e.what());
    }
    double T[12];
    tf2::toMatrix(mBaseToMap, T);
    int moved = 0;
    if (gg_update_pose(map.handle(), map.slot(), inOdom->pose.pose.position.x, inOdom->pose.pose.position.y, T, &moved) != GG_OK)
        throw std::runtime_error(std::string("groundgrid_b200: ") + gg_last_error());
    if (!moved) return mMap_ptr;  // "We havent moved so we have nothing to do"
    mLastPose.pose = inOdom->pose;
    mLastPose.header = inOdom->header;
    return mMap_ptr;
}
