// pump_impl.hpp — TEST INFRASTRUCTURE.  The one definition of the in-process "ROS" pump declared in shim_core.hpp
// (parameter map, subscriber table, published-message log, colour-service stand-in) plus the helpers a harness
// needs to feed a node one PointCloud2 and read what it published.  Included by exactly one translation unit per
// shared library: oracle/ref_shim/ref_harness.cpp (the reference's own nodes) and ros_shell/shim_harness.cpp (this
// repository's drop-in nodes), so both run behind the same pump code.
#pragma once
#include <cstdlib>
#include <sstream>

#include "shim_core.hpp"

// ---------------------------------------------------------------- ros pump state
namespace ros {
namespace shim {
std::map<std::string, std::string>& params() {
  static std::map<std::string, std::string> m;
  return m;
}
std::map<std::string, std::function<void(const sensor_msgs::PointCloud2ConstPtr&)>>& subscribers() {
  static std::map<std::string, std::function<void(const sensor_msgs::PointCloud2ConstPtr&)>> m;
  return m;
}
std::map<std::string, std::vector<sensor_msgs::PointCloud2>>& published() {
  static std::map<std::string, std::vector<sensor_msgs::PointCloud2>> m;
  return m;
}
std::function<bool(const std::vector<sensor_msgs::PointCloud2>&, std::vector<uint8_t>&)>& color_service() {
  static std::function<bool(const std::vector<sensor_msgs::PointCloud2>&, std::vector<uint8_t>&)> f;
  return f;
}
bool parse(const std::string& s, std::string& v) { v = s; return true; }
bool parse(const std::string& s, int& v) { v = std::atoi(s.c_str()); return true; }
bool parse(const std::string& s, float& v) { v = std::strtof(s.c_str(), nullptr); return true; }
bool parse(const std::string& s, double& v) { v = std::strtod(s.c_str(), nullptr); return true; }
bool parse(const std::string& s, bool& v) { v = (s == "1" || s == "true"); return true; }
}  // namespace shim
}  // namespace ros

namespace shim_pump {
typedef std::function<void(const sensor_msgs::PointCloud2ConstPtr&)> Callback;

inline void set_params(const char* kv) {  // "~name=value;~name=value"
  ros::shim::params().clear();
  if (!kv) return;
  std::stringstream ss(kv);
  std::string item;
  while (std::getline(ss, item, ';')) {
    const size_t eq = item.find('=');
    if (eq != std::string::npos) ros::shim::params()[item.substr(0, eq)] = item.substr(eq + 1);
  }
}
inline Callback take_callback() {
  Callback cb;
  if (!ros::shim::subscribers().empty()) cb = ros::shim::subscribers().begin()->second;
  ros::shim::subscribers().clear();
  return cb;
}
inline sensor_msgs::PointCloud2Ptr make_msg(const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step,
                                     uint32_t row_step, int32_t ox, int32_t oy, int32_t oz, int32_t oi, uint32_t sec,
                                     uint32_t nsec) {
  auto m = std::make_shared<sensor_msgs::PointCloud2>();
  m->header.seq = 1;
  m->header.stamp.sec = sec;
  m->header.stamp.nsec = nsec;
  m->header.frame_id = "cloud";
  m->width = width;
  m->height = height;
  m->point_step = point_step;
  m->row_step = row_step;
  m->is_dense = 1;
  const char* names[4] = {"x", "y", "z", "intensity"};
  const int32_t offs[4] = {ox, oy, oz, oi};
  for (int i = 0; i < 4; ++i)
    if (offs[i] >= 0) {
      sensor_msgs::PointField f;
      f.name = names[i];
      f.offset = static_cast<uint32_t>(offs[i]);
      f.datatype = sensor_msgs::PointField::FLOAT32;
      f.count = 1;
      m->fields.push_back(f);
    }
  m->data.assign(data, data + static_cast<size_t>(row_step) * height);
  return m;
}
inline const char* const kConeTopics[4] = {"cones_cloud_unknowns", "cones_cloud_yellows", "cones_cloud_blues",
                                           "cones_cloud_oranges"};

// the deterministic stand-in for the color_classifier service used by both harnesses:
// colour = 1 + fnv1a(x, y, z, intensity of the crop) % 3, empty crops skipped (the Python service `continue`s)
inline bool hash_color_service(const std::vector<sensor_msgs::PointCloud2>& crops, std::vector<uint8_t>& colors) {
  for (const auto& c : crops) {
    const size_t n = static_cast<size_t>(c.width) * c.height;
    if (n == 0) continue;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) {
      const uint8_t* src = c.data.data() + i * c.point_step;
      float v[4];
      std::memcpy(&v[0], src + 0, 4);
      std::memcpy(&v[1], src + 4, 4);
      std::memcpy(&v[2], src + 8, 4);
      std::memcpy(&v[3], src + 16, 4);
      const uint8_t* b = reinterpret_cast<const uint8_t*>(v);
      for (size_t k = 0; k < sizeof(v); ++k) h = (h ^ b[k]) * 1099511628211ull;
    }
    colors.push_back(static_cast<uint8_t>(1 + h % 3));
  }
  return true;
}

// One callback of a node behind the pump.  Ground node: returns the published point count (or -1).
inline int64_t pump_ground(const Callback& cb, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step,
                           uint32_t row_step, int32_t ox, int32_t oy, int32_t oz, int32_t oi, uint8_t* out32,
                           uint32_t* out_point_step, uint32_t* out_n_fields, uint32_t* out_stamp_nsec) {
  ros::shim::published().clear();
  cb(make_msg(data, width, height, point_step, row_step, ox, oy, oz, oi, 100, 123456789));
  auto& q = ros::shim::published()["groundless_cloud"];
  if (q.empty()) return -1;
  const sensor_msgs::PointCloud2& m = q.back();
  if (out32 && !m.data.empty()) std::memcpy(out32, m.data.data(), m.data.size());
  if (out_point_step) *out_point_step = m.point_step;
  if (out_n_fields) *out_n_fields = static_cast<uint32_t>(m.fields.size());
  if (out_stamp_nsec) *out_stamp_nsec = m.header.stamp.nsec;
  return static_cast<int64_t>(m.width) * m.height;
}
// Detection node: out_xy [4][cap][2] floats, counts[4]; 0, or 3 if cap is too small.
inline int pump_detect(const Callback& cb, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step,
                       uint32_t row_step, int32_t ox, int32_t oy, int32_t oz, int32_t oi, float* out_xy, uint32_t* counts,
                       uint32_t cap, uint32_t* out_point_step, uint32_t* out_n_fields) {
  ros::shim::published().clear();
  cb(make_msg(data, width, height, point_step, row_step, ox, oy, oz, oi, 100, 123456789));
  for (int k = 0; k < 4; ++k) {
    auto& q = ros::shim::published()[kConeTopics[k]];
    counts[k] = 0;
    if (q.empty()) continue;
    const sensor_msgs::PointCloud2& m = q.back();
    const uint32_t n = m.width * m.height;
    counts[k] = n;
    if (n > cap) return 3;
    for (uint32_t i = 0; i < n; ++i)
      std::memcpy(&out_xy[(static_cast<size_t>(k) * cap + i) * 2], m.data.data() + static_cast<size_t>(i) * m.point_step, 8);
    if (out_point_step) *out_point_step = m.point_step;
    if (out_n_fields) *out_n_fields = static_cast<uint32_t>(m.fields.size());
  }
  return 0;
}
}  // namespace shim_pump
