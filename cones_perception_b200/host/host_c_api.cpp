// host_c_api.cpp — a small C surface over the C++ host classes (nodes.hpp) so that pytest can
// drive them through ctypes.  Not part of the drop-in boundary (that is include/conesgpu.h).
#include <cstdio>
#include <map>
#include <string>

#include "nodes.hpp"

using namespace cones_host;

namespace {
std::string g_err;

PointCloud2 make_msg(const float* xyzi, uint32_t n, int with_intensity_field, uint32_t sec, uint32_t nsec) {
  PointCloud2 m;
  m.header.seq = 7;
  m.header.stamp_sec = sec;
  m.header.stamp_nsec = nsec;
  m.header.frame_id = "cloud";
  m.height = 1;
  m.width = n;
  m.point_step = 16;
  m.row_step = 16 * n;
  m.fields.resize(with_intensity_field ? 4 : 3);
  const char* names[4] = {"x", "y", "z", "intensity"};
  for (size_t i = 0; i < m.fields.size(); ++i) {
    m.fields[i].name = names[i];
    m.fields[i].offset = static_cast<uint32_t>(4 * i);
  }
  m.data.resize(static_cast<size_t>(n) * 16);
  if (n) std::memcpy(m.data.data(), xyzi, static_cast<size_t>(n) * 16);
  return m;
}
}  // namespace

extern "C" {

const char* ch_last_error(void) { return g_err.c_str(); }

// ---- ConeTracker (no GPU needed) ---------------------------------------------------------
void* ch_tracker_create(int classify_colors, int use_points_buffer, double match_dist, double extension) {
  auto* t = new ConeTracker();
  t->classify_colors = classify_colors != 0;
  t->use_points_buffer = use_points_buffer != 0;
  t->cones_matching_dist_theshold = match_dist;
  t->cone_position_extension_length = extension;
  return t;
}
void ch_tracker_destroy(void* t) { delete static_cast<ConeTracker*>(t); }

// forced_color >= 0: the colour "service" answers with that colour for every crop
int ch_tracker_update(void* tp, const float* xy, uint32_t n, int forced_color, float* out_xy, uint32_t* counts,
                      uint32_t cap) {
  auto* t = static_cast<ConeTracker*>(tp);
  std::vector<std::pair<float, float>> c(n);
  for (uint32_t i = 0; i < n; ++i) c[i] = {xy[2 * i], xy[2 * i + 1]};
  if (forced_color >= 0)
    t->get_colors = [forced_color](const std::vector<std::vector<Point>>& crops) {
      return std::vector<Color>(crops.size(), static_cast<Color>(forced_color));
    };
  std::vector<Point> whole;
  auto clouds = t->update(c, &whole);
  for (int k = 0; k < kNumberOfColors; ++k) {
    counts[k] = static_cast<uint32_t>(clouds[k].size());
    if (clouds[k].size() > cap) return 3;
    for (size_t i = 0; i < clouds[k].size(); ++i) {
      out_xy[(static_cast<size_t>(k) * cap + i) * 2] = clouds[k][i].x;
      out_xy[(static_cast<size_t>(k) * cap + i) * 2 + 1] = clouds[k][i].y;
    }
  }
  return 0;
}

// ---- ConeDetector / GroundRemover (GPU) -----------------------------------------------------
}  // extern "C"
namespace {
// A deterministic stand-in for the color_classifier service, so the three colour-input modes can be
// compared: colour = 1 + fnv1a(bytes) % 3; empty crops are skipped like the service does
// (scripts/color_classifier_server.py:83-84), which shortens the answer.
uint64_t fnv1a(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
  const uint8_t* b = static_cast<const uint8_t*>(data);
  for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}
struct ColorProbe {
  uint64_t n_inputs = 0, n_points = 0, digest = 1469598103934665603ull;
};
std::vector<Color> probe_crops(ColorProbe* pr, const std::vector<std::vector<Point>>& crops) {
  std::vector<Color> out;
  for (const auto& c : crops) {
    pr->n_inputs++;
    if (c.empty()) continue;
    uint64_t h = 1469598103934665603ull;
    for (const Point& p : c) {
      const float v[4] = {p.x, p.y, p.z, p.intensity};
      h = fnv1a(v, sizeof(v), h);
    }
    pr->n_points += c.size();
    pr->digest = fnv1a(&h, sizeof(h), pr->digest);
    out.push_back(static_cast<Color>(1 + h % 3));
  }
  return out;
}
std::vector<Color> probe_images(ColorProbe* pr, const std::vector<uint8_t>& images, const std::vector<uint32_t>& flags) {
  std::vector<Color> out;
  for (size_t i = 0; i < flags.size(); ++i) {
    pr->n_inputs++;
    if (flags[i] & CP_CONE_EMPTY) continue;
    const uint64_t h = fnv1a(images.data() + i * 180, 180);
    pr->digest = fnv1a(&h, sizeof(h), pr->digest);
    out.push_back(static_cast<Color>(1 + h % 3));
  }
  return out;
}
std::map<void*, ColorProbe> g_probes;
}  // namespace
extern "C" {

// mode: ConeDetector::ColorInputs (0 host crops, 1 GPU crops, 2 GPU crops + range images)
void ch_detector_set_color_inputs(void* dp, int mode) {
  auto* det = static_cast<ConeDetector*>(dp);
  det->color_inputs = static_cast<ConeDetector::ColorInputs>(mode);
  ColorProbe* pr = &g_probes[dp];
  det->get_colors = [pr](const std::vector<std::vector<Point>>& crops) { return probe_crops(pr, crops); };
  det->get_colors_from_images = [pr](const std::vector<uint8_t>& im, const std::vector<uint32_t>& fl) {
    return probe_images(pr, im, fl);
  };
}
// the classifier network on the device (ConeDetector::kGpuColors): model = bytes of the .tflite file
int ch_detector_load_color_model(void* dp, const void* model, uint64_t bytes) {
  try {
    static_cast<ConeDetector*>(dp)->load_color_model(model, static_cast<size_t>(bytes));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 4;
  }
}
void ch_detector_color_probe(void* dp, uint64_t* n_inputs, uint64_t* n_points, uint64_t* digest) {
  const ColorProbe& pr = g_probes[dp];
  *n_inputs = pr.n_inputs;
  *n_points = pr.n_points;
  *digest = pr.digest;
}

void* ch_detector_create(uint64_t max_points, int device, const cp_detect_params* d, int classify_colors,
                         int use_points_buffer, int fused_ground) {
  try {
    auto* det = new ConeDetector(max_points, device, 32);
    det->distance_treshold_max = d->distance_treshold_max;
    det->distance_treshold_min = d->distance_treshold_min;
    det->level_threshold = d->level_threshold;
    det->angle_threshold = d->angle_threshold;
    det->voxel_filter_leaf_size_x = d->voxel_filter_leaf_size_x;
    det->voxel_filter_leaf_size_y = d->voxel_filter_leaf_size_y;
    det->voxel_filter_leaf_size_z = d->voxel_filter_leaf_size_z;
    det->min_cluster_size = d->min_cluster_size;
    det->max_cluster_size = d->max_cluster_size;
    det->classify_colors = classify_colors != 0;
    det->use_points_buffer = use_points_buffer != 0;
    det->fused_ground_removal = fused_ground != 0;
    return det;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void ch_detector_destroy(void* d) {
  g_probes.erase(d);
  delete static_cast<ConeDetector*>(d);
}

// one callback: compact xyzi cloud in, the four published clouds out (x, y per cone) plus the
// layout facts of the published messages
int ch_detector_handle(void* dp, const float* xyzi, uint32_t n, int with_intensity_field, float* out_xy,
                       uint32_t* counts, uint32_t cap, uint32_t* out_point_step, uint32_t* out_n_fields,
                       uint32_t* out_stamp_nsec) {
  try {
    auto* det = static_cast<ConeDetector*>(dp);
    PointCloud2 msg = make_msg(xyzi, n, with_intensity_field, 100, 123456789);
    auto clouds = det->cloud_handler(msg);
    for (int k = 0; k < kNumberOfColors; ++k) {
      counts[k] = clouds[k].width;
      if (clouds[k].width > cap) return 3;
      for (uint32_t i = 0; i < clouds[k].width; ++i) {
        float xy[2];
        std::memcpy(xy, clouds[k].data.data() + static_cast<size_t>(i) * clouds[k].point_step, 8);
        out_xy[(static_cast<size_t>(k) * cap + i) * 2] = xy[0];
        out_xy[(static_cast<size_t>(k) * cap + i) * 2 + 1] = xy[1];
      }
    }
    *out_point_step = clouds[0].point_step;
    *out_n_fields = static_cast<uint32_t>(clouds[0].fields.size());
    *out_stamp_nsec = clouds[0].header.stamp_nsec;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 4;
  }
}

void* ch_ground_create(uint64_t max_points, int device) {
  try {
    return new GroundRemover(max_points, device, 32);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void ch_ground_destroy(void* g) { delete static_cast<GroundRemover*>(g); }

int ch_ground_handle(void* gp, const float* xyzi, uint32_t n, int with_intensity_field, void* out32,
                     uint32_t* kept, uint32_t* out_point_step, uint32_t* out_stamp_nsec, uint32_t* out_n_fields) {
  try {
    auto* gr = static_cast<GroundRemover*>(gp);
    PointCloud2 msg = make_msg(xyzi, n, with_intensity_field, 100, 123456789);
    PointCloud2 out = gr->cloud_handler(msg);
    std::memcpy(out32, out.data.data(), out.data.size());
    *kept = gr->last_kept();
    *out_point_step = out.point_step;
    *out_stamp_nsec = out.header.stamp_nsec;
    *out_n_fields = static_cast<uint32_t>(out.fields.size());
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 4;
  }
}

}  // extern "C"
