// shim_harness.cpp — TEST INFRASTRUCTURE.  The two drop-in node shells of this directory, compiled unmodified
// against the same stand-in ROS surface and in-process message pump the reference's own nodes run behind in
// oracle/_ref (oracle/ref_shim).  A test then feeds one PointCloud2 sequence to the reference node and to the
// drop-in node and compares what each publishes, topic by topic (tests/test_host_shell.py).  ROS itself is not
// installed in the build container; on a Noetic box the same two .cpp files build against the real headers.
#include "pump_impl.hpp"

#define main shell_ground_removal_main
#include "ground_removal_node.cpp"
#undef main
#define main shell_cone_detection_main
#include "cone_detection_node.cpp"
#undef main

namespace {
using shim_pump::Callback;
struct GroundShell {
  GroundRemoverNode node;
  Callback cb;
};
struct DetectShell {
  ConeDetectorNode node;
  Callback cb;
};
std::string g_err;
}  // namespace

extern "C" {

const char* shell_last_error(void) { return g_err.c_str(); }

void* shell_ground_create(const char* params) {
  try {
    shim_pump::set_params(params);
    auto* g = new GroundShell();
    g->cb = shim_pump::take_callback();
    return g;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void shell_ground_destroy(void* p) { delete static_cast<GroundShell*>(p); }
int64_t shell_ground_handle(void* p, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step,
                            uint32_t row_step, int32_t ox, int32_t oy, int32_t oz, int32_t oi, uint8_t* out32,
                            uint32_t* out_point_step, uint32_t* out_n_fields, uint32_t* out_stamp_nsec) {
  return shim_pump::pump_ground(static_cast<GroundShell*>(p)->cb, data, width, height, point_step, row_step, ox, oy, oz,
                                oi, out32, out_point_step, out_n_fields, out_stamp_nsec);
}

void* shell_detect_create(const char* params, int service_mode) {
  try {
    shim_pump::set_params(params);
    auto* d = new DetectShell();
    d->cb = shim_pump::take_callback();
    if (service_mode >= 0) ros::shim::color_service() = shim_pump::hash_color_service;
    else ros::shim::color_service() = nullptr;
    return d;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void shell_detect_destroy(void* p) { delete static_cast<DetectShell*>(p); }
int shell_detect_handle(void* p, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step,
                        uint32_t row_step, int32_t ox, int32_t oy, int32_t oz, int32_t oi, float* out_xy, uint32_t* counts,
                        uint32_t cap, uint32_t* out_point_step, uint32_t* out_n_fields) {
  return shim_pump::pump_detect(static_cast<DetectShell*>(p)->cb, data, width, height, point_step, row_step, ox, oy, oz,
                                oi, out_xy, counts, cap, out_point_step, out_n_fields);
}

}  // extern "C"
