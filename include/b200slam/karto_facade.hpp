// karto_facade.hpp — header-only C++ façade over the C ABI (b200slam.h) that keeps the reference's OWN call
// signatures for the hot path, so a lesson6 node switches by changing a namespace:
//
//   reference (lesson6/lib/open_karto/include/open_karto/Mapper.h:1139-1186, Karto.h:5659)        façade
//   ------------------------------------------------------------------------------------------   ----------------------
//   ScanMatcher* ScanMatcher::Create(Mapper*, searchSize, resolution, smearDeviation, rangeThr)   b200slam::ScanMatcher::Create(params, laser)
//   kt_double MatchScan(LocalizedRangeScan*, const LocalizedRangeScanVector&, Pose2&, Matrix3&,   same argument list
//                       kt_bool doPenalize = true, kt_bool doRefineMatch = true)
//   kt_double CorrelateScan(LocalizedRangeScan*, const Pose2&, const Vector2<double>& offset,     same argument list
//                       const Vector2<double>& resolution, double angleOffset, double angleRes,
//                       bool doPenalize, Pose2& mean, Matrix3& cov, bool doingFineMatch)
//   OccupancyGrid* OccupancyGrid::CreateFromScans(const LocalizedRangeScanVector&, resolution)    same argument list
//   kt_bool Mapper::Process(LocalizedRangeScan*) (Mapper.cpp:1999), SetScanSolver (:2220)           b200slam::Mapper
//   hectorslam::HectorSlamProcessor::update(dataContainer, poseHint, map_without_matching)          b200slam::HectorSlamProcessor
//
// Errors: where the reference returns NULL the façade returns nullptr; where it throws (std::runtime_error,
// karto::Exception) the façade throws b200slam::Exception carrying the b2s_status.  Like the reference's
// ScanMatcher a façade object is single-threaded; distinct objects may be used from distinct threads.
// The Mapper tuning values the reference matcher reads through friend access travel in MatcherParams.
#pragma once

#include <array>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../b200slam.h"

namespace b200slam {

struct Exception : std::runtime_error {
  b2s_status status;
  Exception(b2s_status s, const std::string &what) : std::runtime_error(what), status(s) {}
};

inline void check(b2s_status s) {
  if (s != B2S_OK) throw Exception(s, b2s_last_error());
}

struct Pose2 {  // karto::Pose2 (Karto.h:1985-2200)
  double x = 0, y = 0, heading = 0;
  Pose2() = default;
  Pose2(double x_, double y_, double h_) : x(x_), y(y_), heading(h_) {}
  double GetX() const { return x; }
  double GetY() const { return y; }
  double GetHeading() const { return heading; }
};

struct Vector2d {  // karto::Vector2<kt_double>
  double x = 0, y = 0;
  Vector2d() = default;
  Vector2d(double x_, double y_) : x(x_), y(y_) {}
  double GetX() const { return x; }
  double GetY() const { return y; }
};

struct Matrix3 {  // karto::Matrix3 (Karto.h:2330-2700), row-major
  double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double &operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
};

// the part of karto::LaserRangeFinder the hot path reads; defaults = LaserRangeFinder_Hokuyo_UTM_30LX (Karto.h:4048-4065)
inline b2s_laser HokuyoUTM30LX(double range_threshold) {
  const double d = 0.01745329251994329577;  // KT_PI_180
  b2s_laser l;
  std::memset(&l, 0, sizeof(l));
  l.n_readings = 1081;
  l.min_angle = -135 * d;
  l.angular_resolution = 0.25 * d;
  l.min_range = 0.1;
  l.max_range = 30.0;
  l.range_threshold = range_threshold;
  return l;
}

// Mapper defaults for the values the matcher reads (Mapper.cpp:1569-1652)
inline b2s_matcher_params DefaultMatcherParams(double searchSize, double resolution, double smearDeviation,
                                               double rangeThreshold) {
  const double d = 0.01745329251994329577;
  b2s_matcher_params p;
  std::memset(&p, 0, sizeof(p));
  p.search_size = searchSize;
  p.resolution = resolution;
  p.smear_deviation = smearDeviation;
  p.range_threshold = rangeThreshold;
  p.distance_variance_penalty = 0.3 * 0.3;
  p.angle_variance_penalty = (20 * d) * (20 * d);
  p.fine_search_angle_offset = 0.2 * d;
  p.coarse_search_angle_offset = 20 * d;
  p.coarse_angle_resolution = 2 * d;
  p.minimum_angle_penalty = 0.9;
  p.minimum_distance_penalty = 0.5;
  p.use_response_expansion = 0;
  return p;
}

// karto::LocalizedRangeScan as the matcher sees it: readings + (corrected) robot pose (karto_slam.cc:437-440)
class LocalizedRangeScan {
 public:
  LocalizedRangeScan(const std::vector<double> &readings, const Pose2 &pose) : readings_(readings), pose_(pose) {}
  const std::vector<double> &GetRangeReadings() const { return readings_; }
  const Pose2 &GetCorrectedPose() const { return pose_; }
  void SetCorrectedPose(const Pose2 &p) { pose_ = p; }

 private:
  std::vector<double> readings_;
  Pose2 pose_;
};
typedef std::vector<LocalizedRangeScan *> LocalizedRangeScanVector;

class ScanMatcher {
 public:
  // ScanMatcher::Create (Mapper.cpp:126-172): nullptr on invalid parameters
  static ScanMatcher *Create(const b2s_matcher_params &params, const b2s_laser &laser, int maxBaseScans = 128,
                             int device = 0) {
    b2s_matcher *h = nullptr;
    b2s_status s = b2s_matcher_create(&params, &laser, device, 1, maxBaseScans, nullptr, &h);
    if (s == B2S_ERR_BAD_PARAMS) return nullptr;
    check(s);
    return new ScanMatcher(h, laser);
  }
  ~ScanMatcher() { b2s_matcher_destroy(h_); }

  // ScanMatcher::MatchScan (Mapper.cpp:184-291)
  double MatchScan(LocalizedRangeScan *pScan, const LocalizedRangeScanVector &rBaseScans, Pose2 &rMean,
                   Matrix3 &rCovariance, bool doPenalize = true, bool doRefineMatch = true) {
    Upload(pScan, rBaseScans);
    b2s_match_result r;
    check(b2s_matcher_match_scan(h_, doPenalize, doRefineMatch, &r));
    return Unpack(r, rMean, rCovariance);
  }

  // ScanMatcher::CorrelateScan (Mapper.cpp:309-523); uses the grid left by the last MatchScan, like the reference
  double CorrelateScan(LocalizedRangeScan *pScan, const Pose2 &rSearchCenter, const Vector2d &rSearchSpaceOffset,
                       const Vector2d &rSearchSpaceResolution, double searchAngleOffset, double searchAngleResolution,
                       bool doPenalize, Pose2 &rMean, Matrix3 &rCovariance, bool doingFineMatch) {
    (void)pScan;
    b2s_search s;
    s.offset_x = rSearchSpaceOffset.x; s.offset_y = rSearchSpaceOffset.y;
    s.res_x = rSearchSpaceResolution.x; s.res_y = rSearchSpaceResolution.y;
    s.angle_offset = searchAngleOffset; s.angle_res = searchAngleResolution;
    s.do_penalize = doPenalize; s.fine = doingFineMatch;
    const double c[3] = {rSearchCenter.x, rSearchCenter.y, rSearchCenter.heading};
    b2s_match_result r;
    std::memset(&r, 0, sizeof(r));
    std::memcpy(r.cov, rCovariance.m, sizeof(r.cov));  // IN/OUT for the fine stage
    check(b2s_matcher_correlate_scan(h_, c, &s, &r));
    return Unpack(r, rMean, rCovariance);
  }

  b2s_matcher *handle() { return h_; }

 private:
  ScanMatcher(b2s_matcher *h, const b2s_laser &l) : h_(h), laser_(l) {}
  void Upload(LocalizedRangeScan *pScan, const LocalizedRangeScanVector &base) {
    const Pose2 &p = pScan->GetCorrectedPose();
    const double pose[3] = {p.x, p.y, p.heading};
    check(b2s_matcher_set_scans(h_, 1, pScan->GetRangeReadings().data(), pose));
    std::vector<double> br, bp;
    for (const LocalizedRangeScan *s : base) {
      br.insert(br.end(), s->GetRangeReadings().begin(), s->GetRangeReadings().begin() + laser_.n_readings);
      const Pose2 &q = s->GetCorrectedPose();
      bp.push_back(q.x); bp.push_back(q.y); bp.push_back(q.heading);
    }
    check(b2s_matcher_add_scans(h_, (int)base.size(), br.data(), bp.data()));
  }
  static double Unpack(const b2s_match_result &r, Pose2 &mean, Matrix3 &cov) {
    if (r.status != B2S_OK) throw Exception((b2s_status)r.status, "scan match failed");
    mean = Pose2(r.pose[0], r.pose[1], r.pose[2]);
    std::memcpy(cov.m, r.cov, sizeof(r.cov));
    return r.response;
  }
  b2s_matcher *h_;
  b2s_laser laser_;
};

// karto::OccupancyGrid as SlamKarto::updateMap uses it (karto_slam.cc:511-578)
class OccupancyGrid {
 public:
  static OccupancyGrid *CreateFromScans(const LocalizedRangeScanVector &rScans, double resolution, const b2s_laser &laser,
                                        int device = 0) {
    if (rScans.empty()) return nullptr;  // Karto.h:5661-5664
    std::vector<double> r, p;
    for (const LocalizedRangeScan *s : rScans) {
      r.insert(r.end(), s->GetRangeReadings().begin(), s->GetRangeReadings().begin() + laser.n_readings);
      const Pose2 &q = s->GetCorrectedPose();
      p.push_back(q.x); p.push_back(q.y); p.push_back(q.heading);
    }
    b2s_occ_grid *g = nullptr;
    check(b2s_occ_grid_create_from_scans(&laser, (int)rScans.size(), r.data(), p.data(), resolution, device, nullptr, &g));
    return new OccupancyGrid(g);
  }
  ~OccupancyGrid() { b2s_occ_grid_destroy(g_); }
  int GetWidth() const { return info_.width; }
  int GetHeight() const { return info_.height; }
  int GetWidthStep() const { return info_.width_step; }
  Vector2d GetOffset() const { return Vector2d(info_.offset[0], info_.offset[1]); }
  // GridStates: 0 unknown, 100 occupied, 255 free (Karto.h:4193-4198)
  unsigned char GetValue(int x, int y) const { return cells_[(size_t)x + (size_t)y * info_.width_step]; }
  // nav_msgs/OccupancyGrid payload (karto_slam.cc:546-569)
  std::vector<signed char> ToRosData() const {
    std::vector<signed char> out((size_t)info_.width * info_.height);
    check(b2s_occ_grid_copy_ros(g_, reinterpret_cast<int8_t *>(out.data())));
    return out;
  }

 private:
  explicit OccupancyGrid(b2s_occ_grid *g) : g_(g) {
    check(b2s_occ_grid_info_get(g_, &info_));
    cells_.resize((size_t)info_.data_size);
    check(b2s_occ_grid_copy(g_, cells_.data(), nullptr, nullptr));
  }
  b2s_occ_grid *g_;
  b2s_occ_grid_info info_;
  std::vector<unsigned char> cells_;
};

// karto::Mapper as SlamKarto::addScan drives it (karto_slam.cc:407-481; Mapper.cpp:1999-2079): one Process call per
// LaserScan.  The scan object keeps the reference's accessors; Process fills its corrected pose.
class MapperScan {
 public:
  MapperScan(const std::vector<double> &readings) : readings_(readings) {}
  // LocalizedRangeScan(const Name &rSensorName, readings) (Karto.h:5186): scans of different sensors (robots) may feed
  // one Mapper; all of them share the Mapper's b2s_laser
  MapperScan(const std::string &sensorName, const std::vector<double> &readings) : readings_(readings), sensor_(sensorName) {}
  const std::string &GetSensorName() const { return sensor_; }
  void SetOdometricPose(const Pose2 &p) { odometric_ = p; }
  void SetCorrectedPose(const Pose2 &p) { corrected_ = p; }
  void SetTime(double t) { time_ = t; }
  const Pose2 &GetOdometricPose() const { return odometric_; }
  const Pose2 &GetCorrectedPose() const { return corrected_; }
  double GetTime() const { return time_; }
  const std::vector<double> &GetRangeReadings() const { return readings_; }

 private:
  std::vector<double> readings_;
  std::string sensor_ = "laser";
  Pose2 odometric_, corrected_;
  double time_ = 0;
};

class Mapper {
 public:
  // Mapper() + Initialize(rangeThreshold) + the setParam* calls of karto_slam.cc:81-252 (values already in their
  // stored form, see b2s_mapper_params); device matchers are created on the first match
  Mapper(const b2s_mapper_params &params, const b2s_laser &laser, int device = 0) { check(b2s_mapper_create(&params, &laser, device, &h_)); }
  ~Mapper() { b2s_mapper_destroy(h_); }
  Mapper(const Mapper &) = delete;
  Mapper &operator=(const Mapper &) = delete;
  // Mapper::SetScanSolver (Mapper.cpp:2220)
  void SetScanSolver(const b2s_scan_solver *solver) { check(b2s_mapper_set_scan_solver(h_, solver)); }
  // kt_bool Mapper::Process(LocalizedRangeScan*) (Mapper.cpp:1999)
  bool Process(MapperScan *pScan) {
    if (!pScan) return false;
    const Pose2 &o = pScan->GetOdometricPose();
    const double odom[3] = {o.x, o.y, o.heading};
    double corrected[3];
    int32_t ok = 0;
    check(b2s_mapper_process_sensor(h_, pScan->GetSensorName().c_str(), pScan->GetRangeReadings().data(), odom, pScan->GetTime(), &ok,
                                    corrected));
    pScan->SetCorrectedPose(Pose2(corrected[0], corrected[1], corrected[2]));
    return ok != 0;
  }
  // corrected poses of Mapper::GetAllProcessedScans() (Mapper.cpp:2126), as SlamKarto::updateMap reads them
  std::vector<Pose2> GetAllProcessedPoses() const {
    std::vector<double> p(3 * (size_t)b2s_mapper_scan_count(h_));
    if (!p.empty()) check(b2s_mapper_get_poses(h_, p.data()));
    std::vector<Pose2> out;
    for (size_t i = 0; i + 2 < p.size(); i += 3) out.emplace_back(p[i], p[i + 1], p[i + 2]);
    return out;
  }
  b2s_mapper *handle() { return h_; }

 private:
  b2s_mapper *h_ = nullptr;
};

// hectorslam::HectorSlamProcessor (lesson4/include/lesson4/hector_mapping/slam_main/HectorSlamProcessor.h:50-150)
class HectorSlamProcessor {
 public:
  HectorSlamProcessor(float mapResolution, int mapSizeX, int mapSizeY, float startX, float startY, int multi_res_size,
                      int device = 0) {
    check(b2s_hector_slam_create(mapResolution, mapSizeX, mapSizeY, startX, startY, multi_res_size, device, nullptr, &h_));
  }
  ~HectorSlamProcessor() { b2s_hector_slam_destroy(h_); }
  HectorSlamProcessor(const HectorSlamProcessor &) = delete;
  HectorSlamProcessor &operator=(const HectorSlamProcessor &) = delete;
  // update(dataContainer, poseHintWorld, map_without_matching) (:81-108); points = DataContainer entries [n][2]
  void update(const std::vector<float> &points, const float origo[2], const float poseHintWorld[3], bool map_without_matching = false) {
    check(b2s_hector_slam_update(h_, points.data(), (int)(points.size() / 2), origo, poseHintWorld, map_without_matching, pose_, cov_, nullptr));
  }
  void reset() { check(b2s_hector_slam_reset(h_)); }
  const float *getLastScanMatchPose() const { return pose_; }
  const float *getLastScanMatchCovariance() const { return cov_; }
  void setUpdateFactorFree(float f) { free_ = f; check(b2s_hector_slam_set_update_factors(h_, free_, occ_)); }
  void setUpdateFactorOccupied(float f) { occ_ = f; check(b2s_hector_slam_set_update_factors(h_, free_, occ_)); }
  void setMapUpdateMinDistDiff(float d) { dist_ = d; check(b2s_hector_slam_set_map_update_min_diff(h_, dist_, angle_)); }
  void setMapUpdateMinAngleDiff(float a) { angle_ = a; check(b2s_hector_slam_set_map_update_min_diff(h_, dist_, angle_)); }
  b2s_hector_slam *handle() { return h_; }

 private:
  b2s_hector_slam *h_ = nullptr;
  float pose_[3] = {0, 0, 0}, cov_[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  float free_ = 0.4f, occ_ = 0.6f, dist_ = 0.4f, angle_ = 0.13f;
};

}  // namespace b200slam
