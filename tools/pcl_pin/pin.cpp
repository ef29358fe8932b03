// pin.cpp — golden vectors from the REAL pcl::VoxelGrid and pcl::EuclideanClusterExtraction.
//
// The two PCL classes are the only part of the cones_perception hot path whose arithmetic is not in the
// reference repository (find_package(PCL), CMakeLists.txt:16; not vendored), so the CPU oracle of this repo
// restates them from the published algorithm ("parity unpinned").  This program closes that gap on any box that
// has PCL (Ubuntu 20.04 / ROS Noetic ships PCL 1.10): it feeds the crop survivors of the five BASELINE.json
// configs to the real classes, configured exactly as the reference configures them
//     downsample         src/cone_detection.cpp:240-249   VoxelGrid<PointXYZI>, setLeafSize(x, y, z), filter
//     euclidan_cluster   src/cone_detection.cpp:206-220   search::KdTree + EuclideanClusterExtraction,
//                                                         tolerance sqrt(pow(CONE_HEIGHT,2)+pow(CONE_WIDTH,2)),
//                                                         min / max cluster size, extract
// and writes key-free goldens as .npy files:
//     <name>_voxels.npy   float32 [V,4]  the voxel cloud in VoxelGrid's output order (ascending voxel idx)
//     <name>_labels.npy   int32   [V]    canonical label of each voxel: the smallest voxel index of the KEPT cluster
//                                        it belongs to, -1 when its component fails the min/max size filter
//     <name>_order.npy    int32   [K,2]  (min index, size) of the clusters in the order extract() returned them
// tools/pcl_pin/pack.py turns them into tests/golden/pcl_<name>.npz; tests/test_pcl_pin.py consumes those.
//
//   usage: pcl_pin <name> <in.bin (float32 x,y,z,intensity per point)> <leaf_x> <leaf_y> <leaf_z> <min> <max> <out_dir>
//
// In the build container (no PCL) the same source is compiled against the stand-in headers of oracle/ref_shim,
// where the two classes delegate to the oracle: that only checks this file and its output format, it pins nothing.
#include <pcl/filters/voxel_grid.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/search/kdtree.h>
#include <pcl/segmentation/extract_clusters.h>

#include <math.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

typedef pcl::PointXYZI Point;                 // include/perception_handling/utils.hpp:13
typedef pcl::PointCloud<Point> PointCloud;    // :14

static const float CONE_WIDTH = 0.228;        // src/cone_detection.cpp:22-23 (float members, double literals)
static const float CONE_HEIGHT = 0.325;

static bool write_npy(const std::string& path, const char* descr, const std::vector<size_t>& shape, const void* data,
                      size_t bytes) {
  std::string dict = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (";
  for (size_t i = 0; i < shape.size(); ++i) dict += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "");
  dict += "), }";
  while ((10 + dict.size() + 1) % 64 != 0) dict += ' ';
  dict += '\n';
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  const uint16_t hl = static_cast<uint16_t>(dict.size());
  bool ok = fwrite(magic, 1, 8, f) == 8 && fwrite(&hl, 2, 1, f) == 1 && fwrite(dict.data(), 1, dict.size(), f) == dict.size();
  if (bytes) ok = ok && fwrite(data, 1, bytes, f) == bytes;
  fclose(f);
  return ok;
}

int main(int argc, char** argv) {
  if (argc != 9) {
    fprintf(stderr, "usage: %s <name> <in.bin> <leaf_x> <leaf_y> <leaf_z> <min_cluster> <max_cluster> <out_dir>\n", argv[0]);
    return 2;
  }
  const std::string name = argv[1], out_dir = argv[8];
  // the node's members are doubles read from the yaml (src/cone_detection.cpp:36-43)
  const double leaf_x = atof(argv[3]), leaf_y = atof(argv[4]), leaf_z = atof(argv[5]);
  const int min_cluster_size = atoi(argv[6]), max_cluster_size = atoi(argv[7]);

  FILE* f = fopen(argv[2], "rb");
  if (!f) {
    perror(argv[2]);
    return 1;
  }
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<float> raw(static_cast<size_t>(bytes) / 4);
  if (bytes && fread(raw.data(), 1, static_cast<size_t>(bytes), f) != static_cast<size_t>(bytes)) return 1;
  fclose(f);
  const size_t n = raw.size() / 4;

  PointCloud::Ptr cloud(new PointCloud);
  for (size_t i = 0; i < n; ++i) {
    Point p;
    p.x = raw[4 * i];
    p.y = raw[4 * i + 1];
    p.z = raw[4 * i + 2];
    p.intensity = raw[4 * i + 3];
    cloud->push_back(p);
  }

  // ---- downsample, src/cone_detection.cpp:240-249
  PointCloud::Ptr cloud_filtered(new PointCloud);
  {
    pcl::VoxelGrid<Point> vg;
    vg.setInputCloud(cloud);
    vg.setLeafSize(leaf_x, leaf_y, leaf_z);
    vg.filter(*cloud_filtered);
  }

  // ---- euclidan_cluster, src/cone_detection.cpp:206-220
  std::vector<pcl::PointIndices> cluster_indices;
  if (!cloud_filtered->points.empty()) {   // (:166 only calls it on a non-empty cloud)
    pcl::search::KdTree<Point>::Ptr kdtree(new pcl::search::KdTree<Point>);
    kdtree->setInputCloud(cloud_filtered);
    pcl::EuclideanClusterExtraction<Point> ec;
    ec.setClusterTolerance(sqrt(pow(CONE_HEIGHT, 2) + pow(CONE_WIDTH, 2)));
    ec.setMinClusterSize(min_cluster_size);
    ec.setMaxClusterSize(max_cluster_size);
    ec.setSearchMethod(kdtree);
    ec.setInputCloud(cloud_filtered);
    ec.extract(cluster_indices);
  }

  const size_t V = cloud_filtered->points.size();
  std::vector<float> vox(V * 4);
  for (size_t i = 0; i < V; ++i) {
    const Point& p = cloud_filtered->points[i];
    vox[4 * i] = p.x;
    vox[4 * i + 1] = p.y;
    vox[4 * i + 2] = p.z;
    vox[4 * i + 3] = p.intensity;
  }
  std::vector<int32_t> labels(V, -1), order;
  for (size_t c = 0; c < cluster_indices.size(); ++c) {
    const std::vector<int>& idx = cluster_indices[c].indices;
    if (idx.empty()) continue;
    const int mn = *std::min_element(idx.begin(), idx.end());
    for (size_t k = 0; k < idx.size(); ++k) {
      if (labels[idx[k]] != -1) {
        fprintf(stderr, "voxel %d is in two clusters\n", idx[k]);
        return 1;
      }
      labels[idx[k]] = mn;
    }
    order.push_back(mn);
    order.push_back(static_cast<int32_t>(idx.size()));
  }
  bool ok = write_npy(out_dir + "/" + name + "_voxels.npy", "<f4", {V, 4}, vox.data(), vox.size() * 4);
  ok = ok && write_npy(out_dir + "/" + name + "_labels.npy", "<i4", {V}, labels.data(), labels.size() * 4);
  ok = ok && write_npy(out_dir + "/" + name + "_order.npy", "<i4", {order.size() / 2, 2}, order.data(), order.size() * 4);
  if (!ok) {
    fprintf(stderr, "cannot write the .npy files into %s\n", out_dir.c_str());
    return 1;
  }
  printf("%s: %zu points -> %zu voxels -> %zu clusters\n", name.c_str(), n, V, order.size() / 2);
  return 0;
}
