// pcl_delegates.hpp — TEST INFRASTRUCTURE.  The two PCL classes of the stand-in surface (shim_core.hpp) delegate to
// the oracle's pcl_faithful restatement: PCL itself is not installed, so inside oracle/_ref and inside the
// compile check of tools/pcl_pin these stages pin nothing ("parity unpinned").  Include after cones_oracle.h and
// shim_core.hpp, before the first use of VoxelGrid<PointXYZI>::filter / EuclideanClusterExtraction::extract.
#pragma once
#include <cstring>
#include <vector>

namespace pcl {
static_assert(sizeof(PointXYZI) == sizeof(orc_point), "PointXYZI and orc_point share the 32-byte PCL layout");

template <>
void VoxelGrid<PointXYZI>::filter(PointCloud<PointXYZI>& out) {
  const uint32_t n = static_cast<uint32_t>(in_->points.size());
  orc_detect_params d;
  std::memset(&d, 0, sizeof(d));
  d.voxel_filter_leaf_size_x = leaf_[0];  // setLeafSize narrowed the node's doubles to float already
  d.voxel_filter_leaf_size_y = leaf_[1];
  d.voxel_filter_leaf_size_z = leaf_[2];
  std::vector<uint32_t> keys(n ? n : 1), order(n ? n : 1);
  std::vector<orc_point> vox(n ? n : 1);
  uint32_t nv = 0;
  orc_counters ctr;
  orc_voxel_grid(reinterpret_cast<const orc_point*>(in_->points.data()), n, &d, ORC_PCL_FAITHFUL, keys.data(),
                 order.data(), vox.data(), &nv, &ctr);
  out.header = in_->header;
  out.points.assign(nv, PointXYZI());
  if (nv) std::memcpy(out.points.data(), vox.data(), static_cast<size_t>(nv) * sizeof(orc_point));
  out.width = nv;
  out.height = 1;
  out.is_dense = true;
}

template <>
void EuclideanClusterExtraction<PointXYZI>::extract(std::vector<PointIndices>& clusters) {
  clusters.clear();
  const uint32_t n = static_cast<uint32_t>(in_->points.size());
  if (n == 0) return;
  std::vector<orc_cluster> cl(n);
  std::vector<uint32_t> members(n);
  uint32_t k = 0;
  orc_extract_clusters_tol(reinterpret_cast<const orc_point*>(in_->points.data()), n, tol_, min_, max_,
                           ORC_PCL_FAITHFUL, cl.data(), n, &k, members.data());
  size_t start = 0;
  for (uint32_t c = 0; c < k; ++c) {
    PointIndices pi;
    pi.header = in_->header;
    pi.indices.assign(members.begin() + start, members.begin() + start + cl[c].size);
    start += cl[c].size;
    clusters.push_back(pi);
  }
}
}  // namespace pcl
