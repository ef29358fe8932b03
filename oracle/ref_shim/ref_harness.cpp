// ref_harness.cpp — TEST INFRASTRUCTURE.  Builds oracle/_ref/libconesref.so from the reference's own
// node sources, compiled UNMODIFIED from where they lie (/root/reference/src, never copied into this repo),
// against the stand-in ROS / PCL / Eigen surface of ref_shim/include/shim_core.hpp.
//
// What this pins: every line of arithmetic the reference itself wrote —
//   GroundRemover::cloud_handler      src/ground_removal.cpp:50-89   (sector minima, remove_if, zero padding)
//   ConeDetector::cloud_handler       src/cone_detection.cpp:130-187 (faked intensity field, copy, publish)
//   filter_points_position            :189-204                       (level / distance / angle crop)
//   euclidan_cluster                  :206-220                       (tolerance expression, min/max sizes)
//   get_reconstructed_cone            :222-238
//   get_centroid_clouds               :250-340                       (centroid loop, extension, temporal gate)
//   perception_handling::euclidan_dist src/perception_handling/utils.cpp:32-34
// including the C++ overload resolution of its atan2 / pow / floor calls under this compiler and libm.
// What it does NOT pin: PCL's VoxelGrid and EuclideanClusterExtraction, which are not installed; the stand-ins
// below call the oracle's pcl_faithful restatement, so those two stages remain "parity unpinned".
//
// Build flags worth knowing (oracle/Makefile): -ftrivial-auto-var-init=zero makes the reference's uninitialised
// `float x` (src/cone_detection.cpp:264) start at 0, the behaviour the oracle defines.
// The 17th sector slot the reference writes past its 16-float vector (Q2) stays undefined here: the tests and
// golden generators keep azimuths (-8 deg, 0) out of the clouds they feed to this library.
#include "../cones_oracle.h"
#include "shim_core.hpp"

#include "pump_impl.hpp"

// ---------------------------------------------------------------- pcl stand-ins
namespace pcl {
static const sensor_msgs::PointField* find_field(const sensor_msgs::PointCloud2& m, const char* name) {
  for (const auto& f : m.fields)  // exact name, FLOAT32, count 1 (0 is accepted for count 1)
    if (f.name == name && f.datatype == sensor_msgs::PointField::FLOAT32 && (f.count == 1 || f.count == 0)) return &f;
  return nullptr;
}
void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointXYZI>& cloud) {
  cloud.header.seq = msg.header.seq;
  cloud.header.stamp = static_cast<uint64_t>(msg.header.stamp.sec) * 1000000ull + msg.header.stamp.nsec / 1000u;
  cloud.header.frame_id = msg.header.frame_id;
  cloud.width = msg.width;
  cloud.height = msg.height;
  cloud.is_dense = msg.is_dense != 0;
  const size_t n = static_cast<size_t>(msg.width) * msg.height;
  cloud.points.assign(n, PointXYZI());
  const sensor_msgs::PointField* fx = find_field(msg, "x");
  const sensor_msgs::PointField* fy = find_field(msg, "y");
  const sensor_msgs::PointField* fz = find_field(msg, "z");
  const sensor_msgs::PointField* fi = find_field(msg, "intensity");
  for (uint32_t r = 0; r < msg.height; ++r)
    for (uint32_t c = 0; c < msg.width; ++c) {
      const uint8_t* src = msg.data.data() + static_cast<size_t>(r) * msg.row_step + static_cast<size_t>(c) * msg.point_step;
      PointXYZI& p = cloud.points[static_cast<size_t>(r) * msg.width + c];
      if (fx) std::memcpy(&p.x, src + fx->offset, 4);
      if (fy) std::memcpy(&p.y, src + fy->offset, 4);
      if (fz) std::memcpy(&p.z, src + fz->offset, 4);
      if (fi) std::memcpy(&p.intensity, src + fi->offset, 4);
    }
}
void toROSMsg(const PointCloud<PointXYZI>& cloud, sensor_msgs::PointCloud2& msg) {
  uint32_t w = cloud.width, h = cloud.height;
  if (w == 0 && h == 0) {
    w = static_cast<uint32_t>(cloud.points.size());
    h = 1;
  }
  msg.header.seq = cloud.header.seq;
  msg.header.stamp.sec = static_cast<uint32_t>(cloud.header.stamp / 1000000ull);
  msg.header.stamp.nsec = static_cast<uint32_t>(cloud.header.stamp % 1000000ull) * 1000u;
  msg.header.frame_id = cloud.header.frame_id;
  msg.width = w;
  msg.height = h;
  msg.fields.clear();
  const char* names[4] = {"x", "y", "z", "intensity"};
  const uint32_t offs[4] = {0, 4, 8, 16};
  for (int i = 0; i < 4; ++i) {
    sensor_msgs::PointField f;
    f.name = names[i];
    f.offset = offs[i];
    f.datatype = sensor_msgs::PointField::FLOAT32;
    f.count = 1;
    msg.fields.push_back(f);
  }
  msg.is_bigendian = 0;
  msg.point_step = sizeof(PointXYZI);
  msg.row_step = msg.point_step * w;
  msg.is_dense = cloud.is_dense;
  msg.data.resize(cloud.points.size() * sizeof(PointXYZI));
  if (!cloud.points.empty()) std::memcpy(msg.data.data(), cloud.points.data(), msg.data.size());
}

}  // namespace pcl
#include "pcl_delegates.hpp"

// ---------------------------------------------------------------- the reference sources, where they lie
#define main ref_ground_removal_main
#include <src/ground_removal.cpp>
#undef main
#define main ref_cone_detection_main
#include <src/cone_detection.cpp>
#undef main
#include <src/perception_handling/utils.cpp>

// ---------------------------------------------------------------- C surface for the tests / golden scripts
namespace {
using shim_pump::Callback;
using shim_pump::set_params;
using shim_pump::take_callback;
struct GroundNode {
  GroundRemover node;
  Callback cb;
};
struct DetectNode {
  ConeDetector node;
  Callback cb;
};
}  // namespace

extern "C" {

// the real perception_handling::euclidan_dist (utils.cpp:32-34)
float ref_euclidan_dist(float x1, float y1, float z1, float x2, float y2, float z2) {
  return perception_handling::euclidan_dist(x1, y1, z1, x2, y2, z2);
}

void* ref_ground_create(const char* params) {
  set_params(params);
  auto* g = new GroundNode();  // the constructor reads the params and subscribes (src/ground_removal.cpp:31-43)
  g->cb = take_callback();
  return g;
}
void ref_ground_destroy(void* p) { delete static_cast<GroundNode*>(p); }

// One callback of the real node.  out32: width*height PCL points (32 B each) as published on groundless_cloud.
// Returns the number of points in the published cloud, or -1 if nothing was published.
int64_t ref_ground_handle(void* p, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step,
                          uint32_t row_step, int32_t ox, int32_t oy, int32_t oz, int32_t oi, uint8_t* out32,
                          uint32_t* out_point_step, uint32_t* out_n_fields, uint32_t* out_stamp_nsec) {
  return shim_pump::pump_ground(static_cast<GroundNode*>(p)->cb, data, width, height, point_step, row_step, ox, oy, oz,
                                oi, out32, out_point_step, out_n_fields, out_stamp_nsec);
}

// forced_color: < 0 -> the service call fails (ROS_ERROR path); otherwise a deterministic stand-in for the
// color_classifier service: colour = 1 + fnv1a(x,y,z,intensity of the crop) % 3, empty crops skipped
void* ref_detect_create(const char* params, int service_mode) {
  set_params(params);
  auto* d = new DetectNode();
  d->cb = take_callback();
  if (service_mode >= 0) ros::shim::color_service() = shim_pump::hash_color_service;
  else
    ros::shim::color_service() = nullptr;
  return d;
}
void ref_detect_destroy(void* p) { delete static_cast<DetectNode*>(p); }

// One callback of the real node.  out_xy: [4][cap][2] floats (x, y of every published cone per colour topic),
// counts[4].  Returns 0, or 3 if cap is too small.
int ref_detect_handle(void* p, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step,
                      uint32_t row_step, int32_t ox, int32_t oy, int32_t oz, int32_t oi, float* out_xy,
                      uint32_t* counts, uint32_t cap, uint32_t* out_point_step, uint32_t* out_n_fields) {
  return shim_pump::pump_detect(static_cast<DetectNode*>(p)->cb, data, width, height, point_step, row_step, ox, oy, oz,
                                oi, out_xy, counts, cap, out_point_step, out_n_fields);
}

}  // extern "C"
