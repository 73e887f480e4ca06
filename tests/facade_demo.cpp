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
  return 0;
}
