// TEST INFRASTRUCTURE ONLY (oracle/_ref build): the PCL macros velodyne_pointcloud/point_types.h:27-66 uses.
// PCL_ADD_POINT4D is PCL's union { float data[4]; struct { float x, y, z; }; } (pcl/impl/point_types.hpp).
#pragma once
#include <Eigen/Core>
#include <cstdint>
#define PCL_ADD_UNION_POINT4D \
    union EIGEN_ALIGN16 {     \
        float data[4];        \
        struct {              \
            float x;          \
            float y;          \
            float z;          \
        };                    \
    };
#define PCL_ADD_POINT4D PCL_ADD_UNION_POINT4D
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fseq)
