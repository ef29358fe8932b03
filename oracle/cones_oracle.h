/*
 * cones_oracle.h — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A from-scratch CPU restatement of the point-cloud hot path of
 * dmn-sjk/cones_perception (ground removal -> crop -> VoxelGrid -> Euclidean
 * clustering -> per-cluster centroids).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (libconesgpu.so) never links, loads or calls it.
 *
 * PARITY: the reference ships no tests, golden vectors or fixtures.  What the reference
 * wrote itself is pinned against its own node sources compiled unmodified (oracle/ref_shim -> oracle/_ref,
 * tests/test_reference_pin.py, tests/golden/reference_nodes.npz).  PARITY UNPINNED for pcl::VoxelGrid and
 * pcl::EuclideanClusterExtraction: PCL 1.10 / FLANN 1.9.1 are not vendored and not installed here; those two are
 * restated from the published algorithms at the reference's call sites.
 */
#ifndef CONES_ORACLE_H
#define CONES_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* sensor_msgs/PointCloud2 as the reference's cloud_handlers see it
 * (src/cone_detection.cpp:130, src/ground_removal.cpp:50), already resolved to
 * field byte offsets the way pcl::fromROSMsg resolves them. */
typedef struct orc_view {
  const uint8_t* data;
  uint32_t width, height, point_step, row_step;
  int32_t off_x, off_y, off_z;
  int32_t off_intensity; /* <0: field absent and not faked => intensity = 0 */
  uint8_t is_bigendian, is_dense;
} orc_view;

/* pcl::PointXYZI memory layout (include/perception_handling/utils.hpp:13-14) */
typedef struct orc_point {
  float x, y, z, pad;       /* pad = 1.0f */
  float intensity, c1, c2, c3;
} orc_point;

/* ConeDetector members (src/cone_detection.cpp:22-43), same types */
typedef struct orc_detect_params {
  double distance_treshold_max, distance_treshold_min, level_threshold, angle_threshold;
  double voxel_filter_leaf_size_x, voxel_filter_leaf_size_y, voxel_filter_leaf_size_z;
  int32_t min_cluster_size, max_cluster_size;
  float cone_width, cone_height;
} orc_detect_params;

/* GroundRemover members (src/ground_removal.cpp:18-20) */
typedef struct orc_ground_params {
  int32_t num_of_sectors;
  float default_lowest_point;
} orc_ground_params;

typedef struct orc_cluster {
  float x, y;          /* mean of voxel x / y  (src/cone_detection.cpp:261-273) */
  uint32_t size;       /* voxels in the cluster */
  uint32_t min_index;  /* smallest voxel index = canonical label */
} orc_cluster;

enum { ORC_CANONICAL = 0, ORC_PCL_FAITHFUL = 1 };
enum { ORC_NSECT = 17 };

/* per-stage wall-clock of the last orc_detect / orc_ground_node call, seconds */
typedef struct orc_timing {
  double from_msg, copy_cloud, ground, crop, voxel, cluster, centroid, total;
} orc_timing;

typedef struct orc_counters {
  uint32_t n_points, n_ground_kept, n_cropped, n_voxels, n_components, n_clusters;
  uint32_t key_bits, passthrough;
  int32_t min_b[3], div_b[3];
} orc_counters;

/* A.1  pcl::fromROSMsg (src/cone_detection.cpp:151,153; src/ground_removal.cpp:54) */
int orc_from_msg(const orc_view* v, orc_point* out /* width*height */);

/* the reference's atan2(float, float) = libm atan2f = fdlibm's single-precision routine (not correctly rounded) */
float orc_atan2f(float y, float x);
/* seconds for one pass of atan2f over n pairs: the restated routine (use_libm = 0) or this box's libm (1) */
double orc_time_atan2f(const float* y, const float* x, uint32_t n, int use_libm, float* checksum);

/* A.2  src/ground_removal.cpp:58-68 (pass 1) */
int orc_sector_of(float x, float y);
void orc_ground_minima(const orc_point* p, uint32_t n, float default_lowest, float* low /*17*/);
/* A.2  src/ground_removal.cpp:70-77 (pass 2): keep[i]=1 iff the point survives */
void orc_ground_mask(const orc_point* p, uint32_t n, const float* low, uint8_t* keep);
/* whole handler body :54-79: out has N points (survivors then zero points) */
int orc_ground_node(const orc_view* v, const orc_ground_params* g, orc_point* out,
                    uint32_t* n_kept, float* low /*17 or NULL*/, uint8_t* keep /*N or NULL*/);

/* A.3  src/cone_detection.cpp:189-204 + utils.cpp:32-34 */
void orc_crop_mask(const orc_point* p, uint32_t n, const orc_detect_params* d, uint8_t* keep);

/* A.4  pcl::VoxelGrid<PointXYZI>::applyFilter via src/cone_detection.cpp:240-249.
 * keys_sorted/order have n entries (order[r] = input point index of sorted record r);
 * out_vox has capacity n. Returns 0. */
int orc_voxel_grid(const orc_point* p, uint32_t n, const orc_detect_params* d, int mode,
                   uint32_t* keys_sorted, uint32_t* order, orc_point* out_vox,
                   uint32_t* n_vox, orc_counters* ctr);

/* A.5  pcl::EuclideanClusterExtraction via src/cone_detection.cpp:206-220.
 * labels[v] = min voxel index of v's connected component (all components, before the
 * size filter). clusters: kept components in emitted order; members (optional) are the
 * concatenated ascending index lists, cluster k occupying
 * members[start_k .. start_k+size_k) with start_k = sum of earlier sizes. */
int orc_extract_clusters(const orc_point* vox, uint32_t n_vox, const orc_detect_params* d, int mode,
                int32_t* labels, orc_cluster* clusters, uint32_t cap, uint32_t* n_clusters,
                uint32_t* n_components, uint32_t* members /* n_vox or NULL */);

/* the same with the tolerance handed over as pcl::EuclideanClusterExtraction::setClusterTolerance receives it
 * (used by oracle/ref_shim, where the reference's own code computes that double) */
int orc_extract_clusters_tol(const orc_point* vox, uint32_t n_vox, double cluster_tolerance, int32_t min_size,
                             int32_t max_size, int mode, orc_cluster* clusters, uint32_t cap, uint32_t* n_clusters,
                             uint32_t* members /* n_vox or NULL */);

/* independent O(V^2) labeller used to cross-check orc_extract_clusters */
void orc_label_bruteforce(const orc_point* vox, uint32_t n_vox, const orc_detect_params* d,
                          int32_t* labels);

/* squared search radius handed to FLANN (SURVEY Appendix A constants) */
float orc_r2(const orc_detect_params* d);

/* Whole ConeDetector::cloud_handler hot path (:151-175 up to the centroid mean),
 * optionally preceded by the GroundRemover body (fused, zero padding skipped:
 * padded zeros never survive the crop when distance_treshold_min > 0; when it is
 * <= 0 the zeros are materialised like the two-node chain would). */
int orc_detect(const orc_view* v, const orc_detect_params* d, const orc_ground_params* g /*NULL: off*/,
               int mode, orc_cluster* clusters, uint32_t cap, uint32_t* n_clusters,
               orc_counters* ctr, orc_timing* tm);

/* A.7  host post-processing: radial extension (src/cone_detection.cpp:276-278) */
void orc_extend(float* x, float* y, double extension_length);

/* ---- colour path inputs (first two stages) -------------------------------
 * PINNED: scripts/color_classifier_server.py is plain numpy/scipy, so the reference's own
 * to_image runs in the build container; tests/golden/cone_images.npz holds its output on
 * the 577 real cone crops (tests/golden/make_golden.py) and orc_to_image is checked
 * against it byte for byte. */

/* src/cone_detection.cpp:222-238  get_reconstructed_cone: points of the raw cloud inside the
 * axis-aligned box |x-cx|,|y-cy| <= CONE_WIDTH/1.5 (double arithmetic, both ends inclusive),
 * in cloud order.  out: n matches written as orc_point (pad 1.0f), capacity cap; returns the
 * number of matches (may exceed cap; only cap are written). */
uint32_t orc_reconstruct_cone(const orc_point* cloud, uint32_t n, float cx, float cy, float cone_width,
                              orc_point* out, uint32_t cap);

enum { ORC_IMG_ROWS = 15, ORC_IMG_COLS = 12 };
enum { ORC_CONE_EMPTY = 1, ORC_CONE_BAD_INDEX = 2, ORC_CONE_BAD_INTENSITY = 4 };
/* scripts/color_classifier_server.py:130-156  ColorClassifier.to_image on one cone cloud
 * (x,y,z,intensity float32 widened to double, as pc2.read_points hands them over).
 * img: 15x12 bytes, row-major.  Returns a flag word: EMPTY = the handler skips the cone
 * (:83-84); BAD_INDEX = numpy would raise IndexError (row outside [-15,14]); BAD_INTENSITY =
 * interp1d would raise (value outside [0,255] or NaN).  img is all zero when a flag is set. */
uint32_t orc_to_image(const float* xyzi /* n x 4 */, uint32_t n, uint8_t* img);

#ifdef __cplusplus
}
#endif
#endif
