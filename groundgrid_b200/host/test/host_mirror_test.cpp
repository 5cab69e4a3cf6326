// Drives the nodelet callbacks (onInit / odom_callback / points_callback) from a binary script
// written by tests/test_host_mirror.py and dumps the published clouds + selected layers.
//   script: int32 n_scans, float dim, float res, then per scan:
//     double ox, oy, oz (odometry position), double T_map_base[7] (tx ty tz qx qy qz qw),
//     double T_map_velodyne[7], double T_base_map[7], int32 frame_is_map, int32 point_step, int32 n_points, n_points*point_step bytes
//   output: per scan int32 n_out, n_out * 32 bytes (published cloud), then N*N floats ground, N*N floats groundpatch
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <groundgrid/GroundGridNodelet.h>

static geometry_msgs::TransformStamped make_tf(const char* parent, const char* child, const double* v) {
    geometry_msgs::TransformStamped t;
    t.header.frame_id = parent;
    t.child_frame_id = child;
    t.transform.translation.x = v[0]; t.transform.translation.y = v[1]; t.transform.translation.z = v[2];
    t.transform.rotation.x = v[3]; t.transform.rotation.y = v[4]; t.transform.rotation.z = v[5]; t.transform.rotation.w = v[6];
    return t;
}
int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s script.bin out.bin\n", argv[0]); return 2; }
    FILE* in = std::fopen(argv[1], "rb");
    FILE* out = std::fopen(argv[2], "wb");
    if (!in || !out) return 2;
    int n_scans = 0; float dim = 0, res = 0;
    if (std::fread(&n_scans, 4, 1, in) != 1 || std::fread(&dim, 4, 1, in) != 1 || std::fread(&res, 4, 1, in) != 1) return 2;
    ros::gg_shim_log_level() = 2;
    groundgrid::GroundGridNodelet node;
    node.setGeometryOverride(dim, res, 0, 1u << 19);
    node.init();
    std::vector<uint8_t> published;
    int n_published = -1;
    // image sinks (GroundGridNodelet.cpp:219-228): a colour image of "ground" and the terrain image
    int n_layer_images = 0, n_terrain_images = 0;
    size_t layer_image_bytes = 0, terrain_image_bytes = 0;
    node.layer_pubs_["ground"] = [&](const groundgrid::GroundGridNodelet::LayerImage& im) { ++n_layer_images; layer_image_bytes = im.data.size(); };
    node.terrain_im_pub_ = [&](const groundgrid::GroundGridNodelet::LayerImage& im) { ++n_terrain_images; terrain_image_bytes = im.data.size(); };
    std::shared_ptr<sensor_msgs::PointCloud2> last_map_msg;
    double last_map_odo[3] = {0, 0, 0}, last_map_tmv[7] = {0}, last_map_tmb[7] = {0};
    node.filtered_cloud_pub_ = [&](const sensor_msgs::PointCloud2& m) { published = m.data; n_published = (int)m.width; };

    // a cloud before the first odometry must be dropped silently (GroundGridNodelet.cpp:124-125)
    {
        auto msg = std::make_shared<sensor_msgs::PointCloud2>();
        msg->header.frame_id = "map";
        msg->point_step = 32;
        sensor_msgs::PointField f; f.name = "x"; f.offset = 0; msg->fields.push_back(f); f.name = "y"; f.offset = 4; msg->fields.push_back(f); f.name = "z"; f.offset = 8; msg->fields.push_back(f);
        node.points_callback(msg);
        if (n_published != -1) { std::fprintf(stderr, "scan before odometry was not dropped\n"); return 3; }
    }
    for (int s = 0; s < n_scans; ++s) {
        double odo[3], tmb[7], tmv[7], tbm[7];
        int frame_is_map = 0, step = 0, n = 0;
        if (std::fread(odo, 8, 3, in) != 3 || std::fread(tmb, 8, 7, in) != 7 || std::fread(tmv, 8, 7, in) != 7 || std::fread(tbm, 8, 7, in) != 7) return 2;
        if (std::fread(&frame_is_map, 4, 1, in) != 1 || std::fread(&step, 4, 1, in) != 1 || std::fread(&n, 4, 1, in) != 1) return 2;
        auto msg = std::make_shared<sensor_msgs::PointCloud2>();
        msg->data.resize((size_t)n * step);
        if (n && std::fread(msg->data.data(), (size_t)step, (size_t)n, in) != (size_t)n) return 2;
        msg->header.frame_id = frame_is_map ? "map" : "velodyne";
        msg->header.seq = (uint32_t)s;
        msg->width = (uint32_t)n;
        msg->point_step = (uint32_t)step;
        const char* names[5] = {"x", "y", "z", "intensity", "ring"};
        const uint32_t offs32[5] = {0, 4, 8, 16, 20}, offs18[5] = {0, 4, 8, 12, 16};
        for (int k = 0; k < 5; ++k) { sensor_msgs::PointField f; f.name = names[k]; f.offset = step == 32 ? offs32[k] : offs18[k]; msg->fields.push_back(f); }
        // the /tf topic: map <- base_link, map <- velodyne and the inverse base_link <- map
        tf2_ros::shim_broadcast(make_tf("map", "base_link", tmb));
        tf2_ros::shim_broadcast(make_tf("map", "velodyne", tmv));
        tf2_ros::shim_broadcast(make_tf("base_link", "map", tbm));
        auto odom = std::make_shared<nav_msgs::Odometry>();
        odom->header.frame_id = "map";
        odom->pose.pose.position.x = odo[0]; odom->pose.pose.position.y = odo[1]; odom->pose.pose.position.z = odo[2];
        node.odom_callback(odom);
        if (frame_is_map) {
            last_map_msg = msg;
            for (int q = 0; q < 3; ++q) last_map_odo[q] = odo[q];
            for (int q = 0; q < 7; ++q) { last_map_tmv[q] = tmv[q]; last_map_tmb[q] = tmb[q]; }
        }
        n_published = -1;
        node.points_callback(msg);
        if (n_published < 0) { std::fprintf(stderr, "scan %d produced no cloud\n", s); return 3; }
        std::fwrite(&n_published, 4, 1, out);
        std::fwrite(published.data(), 1, published.size(), out);
        const grid_map::Matrix& G = (*node.map())["ground"];
        std::fwrite(G.data(), 4, G.size(), out);
        const grid_map::Matrix& C = (*node.map())["groundpatch"];
        std::fwrite(C.data(), 4, C.size(), out);
    }
    {
        const size_t N = (size_t)node.map()->getSize()(0);
        if (n_layer_images != n_scans || layer_image_bytes != N * N * 3 || n_terrain_images != n_scans || terrain_image_bytes != N * N * 3 * sizeof(float)) {
            std::fprintf(stderr, "image sinks: %d / %d images, %zu / %zu bytes\n", n_layer_images, n_terrain_images, layer_image_bytes, terrain_image_bytes);
            return 3;
        }
    }
    // The reference's public per-phase methods (GroundSegmentation.h:56-62) against filter_cloud: two fresh
    // GroundGrid + GroundSegmentation pairs, same first odometry, same cloud -- once through filter_cloud, once through
    // insert_cloud -> detect_ground_patches (4 sections) -> spiral_ground_interpolation; terrain and confidence must agree bit for bit.
    if (last_map_msg) {
        pcl::PointCloud<groundgrid::GroundSegmentation::PCLPoint>::Ptr cloud(new pcl::PointCloud<groundgrid::GroundSegmentation::PCLPoint>);
        pcl::fromROSMsg(*last_map_msg, *cloud);
        auto odom = std::make_shared<nav_msgs::Odometry>();
        odom->pose.pose.position.x = last_map_odo[0]; odom->pose.pose.position.y = last_map_odo[1]; odom->pose.pose.position.z = last_map_odo[2];
        groundgrid::GroundSegmentation::PCLPoint origin{};
        origin.x = (float)last_map_tmv[0]; origin.y = (float)last_map_tmv[1]; origin.z = (float)last_map_tmv[2];
        const geometry_msgs::TransformStamped mapToBase = make_tf("map", "base_link", last_map_tmb);
        groundgrid::GroundGridConfig cfg;
        ros::NodeHandle nh;
        groundgrid::GroundGrid gridA, gridB;
        groundgrid::GroundSegmentation segA, segB;
        gridA.setGeometryOverride(dim, res, 0, 1u << 19);
        gridB.setGeometryOverride(dim, res, 0, 1u << 19);
        segA.init(nh, (size_t)dim, res);
        segB.init(nh, (size_t)dim, res);
        segA.setConfig(cfg);
        segB.setConfig(cfg);
        auto mapA = gridA.update(odom), mapB = gridB.update(odom);
        auto outA = segA.filter_cloud(cloud, origin, mapToBase, *mapA);
        std::vector<std::pair<size_t, grid_map::Index>> point_index, ignored;
        std::vector<size_t> outliers;
        segB.insert_cloud(cloud, 0, cloud->points.size(), origin, point_index, ignored, outliers, *mapB);
        for (unsigned short section = 0; section < 4; ++section) segB.detect_ground_patches(*mapB, section);
        segB.spiral_ground_interpolation(*mapB, mapToBase);
        const grid_map::Matrix GA = (*mapA)["ground"], CA = (*mapA)["groundpatch"];
        const grid_map::Matrix& GB = (*mapB)["ground"];
        bool same = true;
        for (size_t q = 0; q < GA.size(); ++q) same = same && (GA.data()[q] == GB.data()[q] || (GA.data()[q] != GA.data()[q] && GB.data()[q] != GB.data()[q]));
        const grid_map::Matrix& CB = (*mapB)["groundpatch"];
        for (size_t q = 0; q < CA.size(); ++q) same = same && (CA.data()[q] == CB.data()[q]);
        // the index lists: kept + ignored + outliers account for every point the output cloud holds (plus border cells)
        if (!same || point_index.size() + ignored.size() + outliers.size() < outA->points.size() || point_index.empty()) {
            std::fprintf(stderr, "per-phase methods disagree with filter_cloud (%zu kept, %zu ignored, %zu outliers, %zu out)\n", point_index.size(), ignored.size(),
                         outliers.size(), outA->points.size());
            return 3;
        }
        try {   // a sub-range is refused (see GroundSegmentation.h)
            segB.insert_cloud(cloud, 1, cloud->points.size(), origin, point_index, ignored, outliers, *mapB);
            std::fprintf(stderr, "insert_cloud accepted a sub-range\n");
            return 3;
        } catch (std::invalid_argument&) {}
        // single-cell methods run
        segB.interpolate_cell(*mapB, 10, 10);
        segB.detect_ground_patch<3>(*mapB, 10, 10);
        segB.detect_ground_patch<5>(*mapB, 10, 10);
    }
    // error behaviour: unknown layer -> std::out_of_range like grid_map
    try { (*node.map())["doesNotExist"]; std::fprintf(stderr, "missing layer did not throw\n"); return 3; } catch (std::out_of_range&) {}
    std::fclose(in); std::fclose(out);
    std::printf("host mirror ok: %d scans\n", n_scans);
    return 0;
}
