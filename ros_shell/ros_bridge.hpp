// ros_bridge.hpp — sensor_msgs/PointCloud2 <-> cones_host::PointCloud2 (1:1 field copy).
// Needs ROS Noetic headers — or, for the tests in the build container (no ROS there), the stand-in headers of
// oracle/ref_shim/include, which provide the same names.
#pragma once
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>

#include "../cones_perception_b200/host/nodes.hpp"

namespace cones_ros {

inline cones_host::PointCloud2 from_ros(const sensor_msgs::PointCloud2& m) {
  cones_host::PointCloud2 o;
  o.header.seq = m.header.seq;
  o.header.stamp_sec = m.header.stamp.sec;
  o.header.stamp_nsec = m.header.stamp.nsec;
  o.header.frame_id = m.header.frame_id;
  o.height = m.height;
  o.width = m.width;
  o.fields.resize(m.fields.size());
  for (size_t i = 0; i < m.fields.size(); ++i) {
    o.fields[i].name = m.fields[i].name;
    o.fields[i].offset = m.fields[i].offset;
    o.fields[i].datatype = m.fields[i].datatype;
    o.fields[i].count = m.fields[i].count;
  }
  o.is_bigendian = m.is_bigendian;
  o.point_step = m.point_step;
  o.row_step = m.row_step;
  o.data = m.data;  // one host copy; pin the subscriber's buffers to make the H2D copy direct
  o.is_dense = m.is_dense;
  return o;
}

inline sensor_msgs::PointCloud2 to_ros(const cones_host::PointCloud2& m) {
  sensor_msgs::PointCloud2 o;
  o.header.seq = m.header.seq;
  o.header.stamp.sec = m.header.stamp_sec;
  o.header.stamp.nsec = m.header.stamp_nsec;
  o.header.frame_id = m.header.frame_id;
  o.height = m.height;
  o.width = m.width;
  o.fields.resize(m.fields.size());
  for (size_t i = 0; i < m.fields.size(); ++i) {
    o.fields[i].name = m.fields[i].name;
    o.fields[i].offset = m.fields[i].offset;
    o.fields[i].datatype = m.fields[i].datatype;
    o.fields[i].count = m.fields[i].count;
  }
  o.is_bigendian = m.is_bigendian;
  o.point_step = m.point_step;
  o.row_step = m.row_step;
  o.data = m.data;
  o.is_dense = m.is_dense;
  return o;
}

// ros::param::get with the reference's "not found, setting to default" notice
template <typename T>
void private_param(const char* name, T& value) {
  if (!ros::param::get(std::string("~") + name, value)) ROS_INFO_STREAM(name << " param not found, setting to default: " << value);
}

}  // namespace cones_ros
