// Compile-and-run check of include/b200slam/karto_facade.hpp: the reference's call sequence
// (Create -> MatchScan -> CorrelateScan -> OccupancyGrid::CreateFromScans) through the façade.
// Reads scans from stdin: n_scans, then per scan: pose (3 doubles) + 1081 readings.  Prints results as text.
#include <cstdio>
#include <vector>

#include "b200slam/karto_facade.hpp"

using namespace b200slam;

int main() {
  int n = 0;
  if (scanf("%d", &n) != 1 || n < 2) return 2;
  std::vector<LocalizedRangeScan *> scans;
  for (int s = 0; s < n; s++) {
    double p[3];
    std::vector<double> r(1081);
    if (scanf("%lf %lf %lf", &p[0], &p[1], &p[2]) != 3) return 2;
    for (auto &v : r)
      if (scanf("%lf", &v) != 1) return 2;
    scans.push_back(new LocalizedRangeScan(r, Pose2(p[0], p[1], p[2])));
  }
  b2s_laser laser = HokuyoUTM30LX(9.25);
  if (ScanMatcher::Create(DefaultMatcherParams(1.5, 0.0, 0.03, 9.25), laser) != nullptr) return 3;  // NULL on bad params
  ScanMatcher *m = ScanMatcher::Create(DefaultMatcherParams(1.5, 0.05, 0.03, 9.25), laser);
  if (!m) return 4;
  LocalizedRangeScanVector base(scans.begin(), scans.end() - 1);
  Pose2 mean;
  Matrix3 cov;
  double resp = m->MatchScan(scans.back(), base, mean, cov);
  printf("match %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", resp, mean.x, mean.y, mean.heading, cov(0, 0), cov(1, 1), cov(2, 2));
  const Pose2 &c = scans.back()->GetCorrectedPose();
  resp = m->CorrelateScan(scans.back(), c, Vector2d(0.75, 0.75), Vector2d(0.05, 0.05), 22.5 * 0.01745329251994329577,
                          0.25 * 0.01745329251994329577, true, mean, cov, false);
  printf("corr %.17g %.17g %.17g %.17g\n", resp, mean.x, mean.y, mean.heading);
  OccupancyGrid *g = OccupancyGrid::CreateFromScans(scans, 0.05, laser);
  long occ = 0, fre = 0;
  for (int y = 0; y < g->GetHeight(); y++)
    for (int x = 0; x < g->GetWidth(); x++) {
      occ += g->GetValue(x, y) == 100;
      fre += g->GetValue(x, y) == 255;
    }
  printf("grid %d %d %ld %ld\n", g->GetWidth(), g->GetHeight(), occ, fre);
  delete g;
  delete m;

  // the two front ends, with the reference's own call shapes: karto::Mapper::Process per scan (odometric pose = the
  // pose read in), hectorslam::HectorSlamProcessor::update per scan
  b2s_mapper_params mp;
  b2s_mapper_default_params(&mp, 9.25);
  mp.sequential.search_size = 0.5; mp.sequential.resolution = 0.05;  // small windows keep the demo quick
  mp.loop.search_size = 4.0;
  mp.minimum_travel_distance = 0.0; mp.minimum_travel_heading = 0.0;
  Mapper mapper(mp, laser);
  int processed = 0;
  for (int s = 0; s < n; s++) {
    MapperScan ms(scans[s]->GetRangeReadings());
    ms.SetOdometricPose(scans[s]->GetCorrectedPose());
    ms.SetCorrectedPose(scans[s]->GetCorrectedPose());
    ms.SetTime(0.1 * s);
    processed += mapper.Process(&ms) ? 1 : 0;
  }
  std::vector<Pose2> poses = mapper.GetAllProcessedPoses();
  printf("mapper %d %zu %.17g %.17g %.17g\n", processed, poses.size(), poses.back().x, poses.back().y, poses.back().heading);

  HectorSlamProcessor hector(0.05f, 1024, 1024, 0.5f, 0.5f, 3);
  hector.setUpdateFactorFree(0.4f);
  hector.setUpdateFactorOccupied(0.9f);
  hector.setMapUpdateMinDistDiff(0.0f);
  float hint[3] = {(float)scans[0]->GetCorrectedPose().x, (float)scans[0]->GetCorrectedPose().y, (float)scans[0]->GetCorrectedPose().heading};
  const float origo[2] = {0.f, 0.f};
  for (int s = 0; s < n; s++) {
    std::vector<float> pts;
    const std::vector<double> &r = scans[s]->GetRangeReadings();
    for (int i = 0; i < 1081; i++) {
      const double a = -2.356194490192345 + i * 0.004363323129985824;
      if (!(r[i] > 0.4 && r[i] < 20.0)) continue;
      pts.push_back((float)(r[i] * std::cos(a)) * 20.0f);
      pts.push_back((float)(r[i] * std::sin(a)) * 20.0f);
    }
    hector.update(pts, origo, hint, s == 0);
    for (int k = 0; k < 3; k++) hint[k] = hector.getLastScanMatchPose()[k];
  }
  printf("hector %.9g %.9g %.9g\n", hint[0], hint[1], hint[2]);
  return 0;
}
