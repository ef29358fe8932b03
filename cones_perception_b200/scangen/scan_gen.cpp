// scan_gen.cpp — seeded synthetic LiDAR scan generator (host C++, multi-threaded).
//
// Produces the clouds of BASELINE.json's configs: a spinning multi-beam
// sensor above flat ground, traffic cones (base diameter 0.228 m, height 0.325 m — the
// CONE_WIDTH / CONE_HEIGHT constants of the reference, src/cone_detection.cpp:22-23),
// walls and posts, hit by ray casting; rays that hit nothing return a point at max_range.
// Output is the compact device format: float4 {x, y, z, intensity} per point, ring-major
// (index = (sweep * beams + beam) * az + azimuth_step).
//
// Randomness is counter-based (a 64-bit mix of seed, frame and ray index), so frames can
// be generated in parallel and are reproducible across thread counts.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {

typedef struct scan_sensor {
  uint32_t beams, az, sweeps;
  float elev_min_deg, elev_max_deg;
  float sensor_h;     // height above the ground plane (ground at z = -sensor_h)
  float max_range;    // where non-hitting rays are placed
  float noise_sigma;  // range noise, metres
} scan_sensor;

typedef struct scan_scene {
  const float* cones;  // ncones x 2  (x, y) of the cone axis
  uint32_t ncones;
  const float* walls;  // nwalls x 6  (x0, y0, x1, y1, z0, z1) vertical rectangles
  uint32_t nwalls;
  const float* posts;  // nposts x 5  (x, y, radius, z0, z1) vertical cylinders
  uint32_t nposts;
} scan_scene;

}  // extern "C"

namespace {

constexpr float kConeR = 0.114f, kConeH = 0.325f;
constexpr int kBuckets = 720;

inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline double u01(uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

struct Rng {
  uint64_t key;
  uint64_t ctr = 0;
  explicit Rng(uint64_t k) : key(mix64(k)) {}
  uint64_t next() { return mix64(key ^ mix64(ctr++)); }
  double uniform() { return u01(next()); }
  double uniform(double a, double b) { return a + (b - a) * uniform(); }
};

struct Obj {
  int kind;  // 0 cone, 1 wall, 2 post
  double a[6];
};

struct Scene {
  std::vector<Obj> objs;
  std::vector<std::vector<uint32_t>> buckets;  // azimuth bucket -> object ids
  double ground_z;

  void add_span(uint32_t id, double az_lo, double az_hi) {
    // az in radians, possibly wrapping
    const double w = 2 * M_PI / kBuckets;
    int b0 = static_cast<int>(std::floor(az_lo / w)) - 1, b1 = static_cast<int>(std::floor(az_hi / w)) + 1;
    if (b1 - b0 >= kBuckets) b0 = 0, b1 = kBuckets - 1;
    for (int b = b0; b <= b1; ++b) buckets[((b % kBuckets) + kBuckets) % kBuckets].push_back(id);
  }
  void finish() {
    buckets.assign(kBuckets, {});
    for (uint32_t i = 0; i < objs.size(); ++i) {
      const Obj& o = objs[i];
      if (o.kind == 0 || o.kind == 2) {
        double r = o.kind == 0 ? kConeR : o.a[2];
        double d = std::hypot(o.a[0], o.a[1]);
        if (d <= r * 1.01) {
          add_span(i, 0, 2 * M_PI);
          continue;
        }
        double c = std::atan2(o.a[1], o.a[0]), hw = std::asin(std::fmin(1.0, r / d)) * 1.01 + 1e-4;
        add_span(i, c - hw, c + hw);
      } else {
        double a0 = std::atan2(o.a[1], o.a[0]), a1 = std::atan2(o.a[3], o.a[2]);
        double dlt = a1 - a0;
        while (dlt > M_PI) dlt -= 2 * M_PI;
        while (dlt < -M_PI) dlt += 2 * M_PI;
        double lo = dlt >= 0 ? a0 : a0 + dlt, hi = dlt >= 0 ? a0 + dlt : a0;
        add_span(i, lo - 1e-3, hi + 1e-3);
      }
    }
  }
};

// nearest positive hit distance along unit direction (dx,dy,dz) from the origin
inline double cast(const Scene& sc, double dx, double dy, double dz, double az, double max_range) {
  double best = max_range;
  bool hit = false;
  if (dz < -1e-9) {
    double t = sc.ground_z / dz;
    if (t > 0 && t < best) best = t, hit = true;
  }
  double azp = az < 0 ? az + 2 * M_PI : az;
  int b = static_cast<int>(azp / (2 * M_PI / kBuckets));
  if (b >= kBuckets) b = kBuckets - 1;
  for (uint32_t id : sc.buckets[b]) {
    const Obj& o = sc.objs[id];
    if (o.kind == 0) {
      // cone: apex (cx, cy, zg + H), radius at height h above ground = R (1 - h/H)
      const double cx = o.a[0], cy = o.a[1], za = sc.ground_z + kConeH, k = kConeR / kConeH;
      // (x-cx)^2 + (y-cy)^2 = k^2 (za - z)^2, with p = t d
      const double A = dx * dx + dy * dy - k * k * dz * dz;
      const double B = -2 * (dx * cx + dy * cy) - 2 * k * k * za * dz * -1.0;
      const double C = cx * cx + cy * cy - k * k * za * za;
      // derivation: (t dx - cx)^2 + (t dy - cy)^2 - k^2 (za - t dz)^2 = 0
      //  t^2 (dx^2+dy^2-k^2 dz^2) + t (-2 dx cx - 2 dy cy + 2 k^2 za dz) + (cx^2+cy^2-k^2 za^2)
      const double disc = B * B - 4 * A * C;
      if (disc < 0 || std::fabs(A) < 1e-12) continue;
      const double sq = std::sqrt(disc);
      double ts[2] = {(-B - sq) / (2 * A), (-B + sq) / (2 * A)};
      for (double t : ts) {
        if (t <= 0 || t >= best) continue;
        double z = t * dz;
        if (z < sc.ground_z || z > za) continue;
        best = t, hit = true;
      }
    } else if (o.kind == 2) {
      const double cx = o.a[0], cy = o.a[1], r = o.a[2];
      const double A = dx * dx + dy * dy, B = -2 * (dx * cx + dy * cy), C = cx * cx + cy * cy - r * r;
      const double disc = B * B - 4 * A * C;
      if (disc < 0 || A < 1e-12) continue;
      const double sq = std::sqrt(disc);
      double ts[2] = {(-B - sq) / (2 * A), (-B + sq) / (2 * A)};
      for (double t : ts) {
        if (t <= 0 || t >= best) continue;
        double z = t * dz;
        if (z < o.a[3] || z > o.a[4]) continue;
        best = t, hit = true;
      }
    } else {
      // vertical rectangle through (x0,y0)-(x1,y1), z in [z0,z1]
      const double ex = o.a[2] - o.a[0], ey = o.a[3] - o.a[1];
      const double den = dx * ey - dy * ex;
      if (std::fabs(den) < 1e-12) continue;
      const double t = (o.a[0] * ey - o.a[1] * ex) / den;
      const double s = (o.a[0] * dy - o.a[1] * dx) / den;
      if (t <= 0 || t >= best || s < 0 || s > 1) continue;
      const double z = t * dz;
      if (z < o.a[4] || z > o.a[5]) continue;
      best = t, hit = true;
    }
  }
  (void)hit;
  return best;
}

void build_scene(Scene& sc, const scan_sensor& s, const scan_scene& in, Rng* jitter, double yaw) {
  sc.objs.clear();
  sc.ground_z = -static_cast<double>(s.sensor_h);
  const double cy = std::cos(yaw), sy = std::sin(yaw);
  auto rot = [&](double& x, double& y) {
    double nx = cy * x - sy * y, ny = sy * x + cy * y;
    x = nx, y = ny;
  };
  for (uint32_t i = 0; i < in.ncones; ++i) {
    Obj o{0, {in.cones[2 * i], in.cones[2 * i + 1], 0, 0, 0, 0}};
    if (jitter) {
      o.a[0] += jitter->uniform(-0.3, 0.3);
      o.a[1] += jitter->uniform(-0.3, 0.3);
    }
    rot(o.a[0], o.a[1]);
    sc.objs.push_back(o);
  }
  for (uint32_t i = 0; i < in.nwalls; ++i) {
    const float* w = in.walls + 6 * i;
    Obj o{1, {w[0], w[1], w[2], w[3], w[4], w[5]}};
    rot(o.a[0], o.a[1]);
    rot(o.a[2], o.a[3]);
    sc.objs.push_back(o);
  }
  for (uint32_t i = 0; i < in.nposts; ++i) {
    const float* p = in.posts + 5 * i;
    Obj o{2, {p[0], p[1], p[2], p[3], p[4], 0}};
    rot(o.a[0], o.a[1]);
    sc.objs.push_back(o);
  }
  sc.finish();
}

void generate_one(const scan_sensor& s, const scan_scene& in, uint64_t seed, bool jitter, float* out) {
  Rng frame_rng(seed * 0x2545F4914F6CDD1Dull + 17);
  double yaw = 0.0;
  Scene sc;
  if (jitter) {
    yaw = frame_rng.uniform(-M_PI, M_PI);
    build_scene(sc, s, in, &frame_rng, yaw);
  } else {
    build_scene(sc, s, in, nullptr, 0.0);
  }
  const uint64_t key = mix64(seed ^ 0xC0FFEE1234ull);
  const double emin = s.elev_min_deg * M_PI / 180, emax = s.elev_max_deg * M_PI / 180;
  size_t idx = 0;
  for (uint32_t sw = 0; sw < s.sweeps; ++sw) {
    const double az_off = (s.sweeps > 1) ? (static_cast<double>(sw) / s.sweeps) : 0.0;
    for (uint32_t b = 0; b < s.beams; ++b) {
      const double e = s.beams > 1 ? emin + (emax - emin) * b / (s.beams - 1) : emin;
      const double ce = std::cos(e), se = std::sin(e);
      for (uint32_t a = 0; a < s.az; ++a, ++idx) {
        double az = -M_PI + 2 * M_PI * (a + az_off) / s.az;
        const double dx = ce * std::cos(az), dy = ce * std::sin(az), dz = se;
        double t = cast(sc, dx, dy, dz, az, s.max_range);
        // Box-Muller range noise + uniform intensity, counter-based per ray
        uint64_t h0 = mix64(key ^ mix64(idx * 3 + 0)), h1 = mix64(key ^ mix64(idx * 3 + 1)),
                 h2 = mix64(key ^ mix64(idx * 3 + 2));
        double g = std::sqrt(-2.0 * std::log(u01(h0))) * std::cos(2 * M_PI * u01(h1));
        t += s.noise_sigma * g;
        out[4 * idx + 0] = static_cast<float>(t * dx);
        out[4 * idx + 1] = static_cast<float>(t * dy);
        out[4 * idx + 2] = static_cast<float>(t * dz);
        out[4 * idx + 3] = static_cast<float>(100.0 * u01(h2));
      }
    }
  }
}

}  // namespace

extern "C" {

uint64_t scan_points_per_frame(const scan_sensor* s) {
  return static_cast<uint64_t>(s->beams) * s->az * s->sweeps;
}

// frame f uses seed base_seed + f. jitter: per-frame cone jitter (+-0.3 m) and scene yaw.
int scan_generate_batch(const scan_sensor* s, const scan_scene* scene, uint64_t base_seed, uint32_t frames,
                        int jitter, float* out, int nthreads) {
  if (!s || !scene || !out) return 1;
  const size_t per = scan_points_per_frame(s) * 4;
  if (nthreads < 1) nthreads = 1;
  if (static_cast<uint32_t>(nthreads) > frames) nthreads = static_cast<int>(frames ? frames : 1);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([=]() {
      for (uint32_t f = t; f < frames; f += nthreads) generate_one(*s, *scene, base_seed + f, jitter != 0, out + per * f);
    });
  for (auto& x : th) x.join();
  return 0;
}

}  // extern "C"
