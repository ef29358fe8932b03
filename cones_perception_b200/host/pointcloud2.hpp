// pointcloud2.hpp — ROS-free mirror of sensor_msgs/PointCloud2 (the message type of both
// reference callbacks, src/cone_detection.cpp:130 and src/ground_removal.cpp:50) so the host
// side of the drop-in can be built and tested without ROS.  Field names, types and semantics
// follow sensor_msgs exactly; the ROS shells in ros_shell/ convert 1:1.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/conesgpu.h"

namespace cones_host {

struct PointField {
  enum : uint8_t { INT8 = 1, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 };
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = FLOAT32;
  uint32_t count = 1;
};

struct Header {
  uint32_t seq = 0;
  uint32_t stamp_sec = 0, stamp_nsec = 0;
  std::string frame_id;
};

struct PointCloud2 {
  Header header;
  uint32_t height = 1, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
};

// pcl::fromROSMsg's field resolution: exact name, FLOAT32, count 1
inline int32_t field_offset(const PointCloud2& m, const char* name) {
  for (const auto& f : m.fields)
    if (f.name == name && f.datatype == PointField::FLOAT32 && f.count == 1) return static_cast<int32_t>(f.offset);
  return -1;
}

// perception_handling::intensity_in_cloud, src/perception_handling/utils.cpp:36-43 (name match only)
inline bool intensity_in_cloud(const PointCloud2& m) {
  for (const auto& f : m.fields)
    if (f.name == "intensity") return true;
  return false;
}

inline cp_cloud_view make_view(const PointCloud2& m, bool fake_missing_intensity) {
  cp_cloud_view v{};
  v.data = m.data.data();
  v.width = m.width;
  v.height = m.height;
  v.point_step = m.point_step;
  v.row_step = m.row_step;
  v.off_x = field_offset(m, "x");
  v.off_y = field_offset(m, "y");
  v.off_z = field_offset(m, "z");
  v.off_intensity = field_offset(m, "intensity");
  if (v.off_intensity < 0 && fake_missing_intensity) v.off_intensity = 0;  // src/cone_detection.cpp:142-151
  v.is_bigendian = m.is_bigendian ? 1 : 0;
  v.is_dense = m.is_dense ? 1 : 0;
  return v;
}

// what pcl::toROSMsg emits for a pcl::PointCloud<pcl::PointXYZI>: 32-byte points,
// x@0 y@4 z@8 intensity@16 (FLOAT32), height/width/is_dense from the PCL cloud
inline std::vector<PointField> pcl_xyzi_fields() {
  std::vector<PointField> f(4);
  f[0].name = "x"; f[0].offset = 0;
  f[1].name = "y"; f[1].offset = 4;
  f[2].name = "z"; f[2].offset = 8;
  f[3].name = "intensity"; f[3].offset = 16;
  return f;
}

// PCL point record, include/perception_handling/utils.hpp:13-14 (pcl::PointXYZI)
struct Point {
  float x = 0, y = 0, z = 0, pad = 1.0f;
  float intensity = 0, c1 = 0, c2 = 0, c3 = 0;
};
static_assert(sizeof(Point) == 32, "pcl::PointXYZI is 32 bytes");

// pcl::toROSMsg of a vector of points (height 1, width n, is_dense true)
inline PointCloud2 to_msg(const std::vector<Point>& pts) {
  PointCloud2 m;
  m.height = 1;
  m.width = static_cast<uint32_t>(pts.size());
  m.fields = pcl_xyzi_fields();
  m.point_step = 32;
  m.row_step = 32 * m.width;
  m.is_dense = true;
  m.data.resize(pts.size() * 32);
  if (!pts.empty()) std::memcpy(m.data.data(), pts.data(), pts.size() * 32);
  return m;
}

}  // namespace cones_host
