// shim_core.hpp — TEST INFRASTRUCTURE.  Minimal stand-ins for the ROS / PCL / Eigen API surface that
// the reference's node sources touch, so that /root/reference/src/{ground_removal,cone_detection}.cpp and
// src/perception_handling/utils.cpp compile UNMODIFIED (where they lie) into oracle/_ref/libconesref.so.
// The reference's own arithmetic (sector loop, crop lambda, centroid loop, radial extension, temporal gate,
// box gather) then runs for real.  What is NOT the reference's code stays a restatement and is marked so:
//   * pcl::fromROSMsg / toROSMsg / copyPointCloud      -> field-mapped copies, written here
//   * pcl::VoxelGrid, pcl::EuclideanClusterExtraction  -> delegate to the oracle's pcl_faithful restatement
//   * ros::*                                           -> an in-process message pump (ref_harness.cpp)
// Nothing here is copied from ROS, PCL or Eigen; only the names and call signatures the reference uses exist.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

// ---------------------------------------------------------------- messages
namespace ros {
struct Time {
  uint32_t sec = 0, nsec = 0;
};
}  // namespace ros
namespace std_msgs {
struct Header {
  uint32_t seq = 0;
  ros::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs
namespace sensor_msgs {
struct PointField {
  enum { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 0;
};
struct PointCloud2 {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  uint8_t is_bigendian = 0;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  uint8_t is_dense = 0;
  typedef std::shared_ptr<PointCloud2> Ptr;
  typedef std::shared_ptr<const PointCloud2> ConstPtr;
};
typedef std::shared_ptr<PointCloud2> PointCloud2Ptr;
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
}  // namespace sensor_msgs
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Point { double x = 0, y = 0, z = 0; };
struct Pose { Point position; Quaternion orientation; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped {
  std_msgs::Header header;
  std::string child_frame_id;
  Transform transform;
};
}  // namespace geometry_msgs
namespace nav_msgs {
struct Odometry { std_msgs::Header header; };
}  // namespace nav_msgs

// ---------------------------------------------------------------- ros: in-process pump
namespace ros {
namespace shim {
// parameters the harness sets before constructing a node ("~name" keys, as the nodes ask for them)
std::map<std::string, std::string>& params();
// topic -> callback registered by NodeHandle::subscribe
std::map<std::string, std::function<void(const sensor_msgs::PointCloud2ConstPtr&)>>& subscribers();
// topic -> messages published since the harness last cleared it
std::map<std::string, std::vector<sensor_msgs::PointCloud2>>& published();
// the color_classifier service stand-in: crops in, colours out (may answer fewer, like the Python service)
std::function<bool(const std::vector<sensor_msgs::PointCloud2>&, std::vector<uint8_t>&)>& color_service();
bool parse(const std::string& s, std::string& v);
bool parse(const std::string& s, int& v);
bool parse(const std::string& s, float& v);
bool parse(const std::string& s, double& v);
bool parse(const std::string& s, bool& v);
}  // namespace shim

inline void init(int&, char**, const std::string&) {}
inline void spin() {}
inline void shutdown() {}
inline bool ok() { return true; }
namespace param {
template <typename T>
bool get(const std::string& key, T& value) {
  auto it = shim::params().find(key);
  return it != shim::params().end() && shim::parse(it->second, value);
}
}  // namespace param

struct Subscriber {};
struct Publisher {
  std::string topic;
  void publish(const sensor_msgs::PointCloud2& m) const { shim::published()[topic].push_back(m); }
};
struct ServiceClient {
  bool valid = false;
  void waitForExistence() const {}
  template <typename Srv>
  bool call(Srv& srv) const {
    if (!shim::color_service()) return false;
    srv.response.colors.clear();
    return shim::color_service()(srv.request.cones_clouds, srv.response.colors);
  }
};
struct NodeHandle {
  template <typename M, typename T>
  Subscriber subscribe(const std::string& topic, uint32_t, void (T::*fn)(const std::shared_ptr<const M>&), T* obj) {
    shim::subscribers()[topic] = [obj, fn](const sensor_msgs::PointCloud2ConstPtr& m) { (obj->*fn)(m); };
    return Subscriber();
  }
  template <typename M>
  Publisher advertise(const std::string& topic, uint32_t) {
    Publisher p;
    p.topic = topic;
    return p;
  }
  template <typename Srv>
  ServiceClient serviceClient(const std::string&) {
    ServiceClient c;
    c.valid = true;
    return c;
  }
};
}  // namespace ros
#include <sstream>
#define ROS_INFO(...) do { } while (0)
#define ROS_INFO_STREAM(args) do { std::ostringstream ros_shim_os; ros_shim_os << args; } while (0)
#define ROS_FATAL(...) do { std::fprintf(stderr, "[ROS_FATAL] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ref ROS_ERROR] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)

// ---------------------------------------------------------------- Eigen: only what the sources name
namespace Eigen {
template <typename T>
struct aligned_allocator {  // 16-byte aligned allocation, like Eigen's
  typedef T value_type;
  aligned_allocator() = default;
  template <typename U>
  aligned_allocator(const aligned_allocator<U>&) {}
  T* allocate(std::size_t n) {
    void* p = nullptr;
    const std::size_t bytes = n * sizeof(T);
    if (posix_memalign(&p, 16, bytes != 0 ? bytes : 16) != 0) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T* p, std::size_t) { free(p); }
  template <typename U>
  struct rebind { typedef aligned_allocator<U> other; };
  bool operator==(const aligned_allocator&) const { return true; }
  bool operator!=(const aligned_allocator&) const { return false; }
};
// matrix_to_trans (src/perception_handling/utils.cpp:11-31) is never called on the hot path; these
// exist only so that translation unit compiles.
struct Matrix3f {};
struct Matrix4f {
  float m[16] = {0};
  float operator()(int r, int c) const { return m[r * 4 + c]; }
  template <int R, int C>
  Matrix3f block(int, int) const { return Matrix3f(); }
};
struct Quaternionf {
  float qx = 0, qy = 0, qz = 0, qw = 1;
  explicit Quaternionf(const Matrix3f&) {}
  void normalize() {}
  float x() const { return qx; }
  float y() const { return qy; }
  float z() const { return qz; }
  float w() const { return qw; }
};
}  // namespace Eigen

// ---------------------------------------------------------------- pcl: types + conversions
namespace pcl {
template <typename T>
using shared_ptr = std::shared_ptr<T>;  // pcl::shared_ptr (boost::shared_ptr in PCL 1.10; same semantics here)

struct alignas(16) PointXYZI {  // 32 bytes: x y z pad(1.0f) | intensity + 12 bytes (point_types.hpp layout)
  float x = 0.f, y = 0.f, z = 0.f, data3 = 1.f;
  float intensity = 0.f, pad1 = 0.f, pad2 = 0.f, pad3 = 0.f;
};
struct PCLHeader {
  uint32_t seq = 0;
  uint64_t stamp = 0;  // microseconds
  std::string frame_id;
};
struct PointIndices {
  PCLHeader header;
  std::vector<int> indices;
};
template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  PCLHeader header;
  std::vector<PointT, Eigen::aligned_allocator<PointT>> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  void push_back(const PointT& p) {  // PCL: width = size, height = 1
    points.push_back(p);
    width = static_cast<uint32_t>(points.size());
    height = 1;
  }
  void resize(std::size_t n) {  // PCL 1.10: points.resize; width/height adjusted to n x 1 when they disagree
    points.resize(n);
    if (width * height != n) {
      width = static_cast<uint32_t>(n);
      height = 1;
    }
  }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointT& operator[](std::size_t i) { return points[i]; }
};
// restated (PCL conversions.h / pcl_conversions.h): exact-name FLOAT32 count-1 field mapping, header with
// microsecond stamp, 32-byte x/y/z/intensity layout on the way out
void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointXYZI>& cloud);
void toROSMsg(const PointCloud<PointXYZI>& cloud, sensor_msgs::PointCloud2& msg);
inline void copyPointCloud(const PointCloud<PointXYZI>& in, PointCloud<PointXYZI>& out) { out = in; }

namespace search {
template <typename PointT>
struct KdTree {  // built by the reference (src/cone_detection.cpp:207-208); the search itself lives in the
  typedef std::shared_ptr<KdTree<PointT>> Ptr;  // oracle's restatement of extractEuclideanClusters
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr&) {}
};
}  // namespace search

template <typename PointT>
class VoxelGrid {  // delegates to orc_voxel_grid (ORC_PCL_FAITHFUL)
 public:
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { in_ = c; }
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void filter(PointCloud<PointT>& out);
 private:
  typename PointCloud<PointT>::ConstPtr in_;
  float leaf_[3] = {0, 0, 0};
};
template <typename PointT>
class EuclideanClusterExtraction {  // delegates to orc_extract_clusters (ORC_PCL_FAITHFUL)
 public:
  void setClusterTolerance(double t) { tol_ = t; }
  void setMinClusterSize(int n) { min_ = n; }
  void setMaxClusterSize(int n) { max_ = n; }
  void setSearchMethod(const typename search::KdTree<PointT>::Ptr&) {}
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { in_ = c; }
  void extract(std::vector<PointIndices>& clusters);
 private:
  typename PointCloud<PointT>::ConstPtr in_;
  double tol_ = 0;
  int min_ = 1, max_ = 0x7fffffff;
};
}  // namespace pcl

// ---------------------------------------------------------------- generated service header
namespace cones_perception {
struct ClassifyColorSrv {  // srv/ClassifyColorSrv.srv: PointCloud2[] cones_clouds --- uint8[] colors
  struct Request { std::vector<sensor_msgs::PointCloud2> cones_clouds; } request;
  struct Response { std::vector<uint8_t> colors; } response;
};
}  // namespace cones_perception
