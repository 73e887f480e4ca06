/*
 * gmapping_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Plain-C restatement of the lesson4 GMapping map update (K2b), paths relative to /root/reference/lesson4:
 *   GridLineTraversal::gridLineCore / gridLine     include/lesson4/gmapping/grid/gridlinetraversal.h:27-207
 *   Map ctor (patch rounding), world2map           include/lesson4/gmapping/grid/map.h:133-140, 171-174
 *   PointAccumulator::update / operator double     include/lesson4/gmapping/grid/map.h:27, 37-48
 *   GMapping::ComputeMap / PublishMap thresholding src/gmapping/gmapping.cc:141-159, 171-242
 * PARITY PINNED against the reference's real headers (oracle/ref_gmapping.cpp -> oracle/_ref/libgmapping_ref.so)
 * by tests/test_oracle_vs_reference.py, and against tests/golden/gmapping.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Map(center, xmin, ymin, xmax, ymax, delta) (map.h:133-140): storage is (cells >> 5) patches of 32 cells */
void orc_gmap_layout(double cx, double cy, double xmin, double ymin, double xmax, double ymax, double delta,
                     int32_t out[4] /* mapSizeX, mapSizeY, sizeX2, sizeY2 */) {
  int xs = (int)ceil((xmax - xmin) / delta), ys = (int)ceil((ymax - ymin) / delta);
  out[0] = (xs >> 5) << 5;
  out[1] = (ys >> 5) << 5;
  out[2] = (int)round((cx - xmin) / delta);
  out[3] = (int)round((cy - ymin) / delta);
}

/* GridLineTraversal::gridLine (gridlinetraversal.h:27-207): cells in order, starting at (x0,y0) */
int orc_gmap_grid_line(int sx, int sy, int ex, int ey, int32_t *out_xy, int cap) {
  int dx = abs(ex - sx), dy = abs(ey - sy), cnt = 0, d, incr1, incr2, x, y, xend, yend, flag;
#define PUSH(px, py) do { if (cnt < cap) { out_xy[2 * cnt] = (px); out_xy[2 * cnt + 1] = (py); } cnt++; } while (0)
  if (dy <= dx) {
    d = 2 * dy - dx; incr1 = 2 * dy; incr2 = 2 * (dy - dx);
    if (sx > ex) { x = ex; y = ey; flag = -1; xend = sx; } else { x = sx; y = sy; flag = 1; xend = ex; }
    PUSH(x, y);
    int up = ((ey - sy) * flag) > 0;
    while (x < xend) {
      x++;
      if (d < 0) d += incr1; else { y += up ? 1 : -1; d += incr2; }
      PUSH(x, y);
    }
  } else {
    d = 2 * dx - dy; incr1 = 2 * dx; incr2 = 2 * (dx - dy);
    if (sy > ey) { y = ey; x = ex; yend = sy; flag = -1; } else { y = sy; x = sx; yend = ey; flag = 1; }
    PUSH(x, y);
    int right = ((ex - sx) * flag) > 0;
    while (y < yend) {
      y++;
      if (d < 0) d += incr1; else { x += right ? 1 : -1; d += incr2; }
      PUSH(x, y);
    }
  }
#undef PUSH
  /* reverse so the list starts at the start point (gridlinetraversal.h:196-206) */
  int n = cnt < cap ? cnt : cap;
  if (n > 0 && (out_xy[0] != sx || out_xy[1] != sy)) {
    for (int i = 0, j = n - 1; i < n / 2; i++, j--) {
      int32_t tx = out_xy[2 * i], ty = out_xy[2 * i + 1];
      out_xy[2 * i] = out_xy[2 * j]; out_xy[2 * i + 1] = out_xy[2 * j + 1];
      out_xy[2 * j] = tx; out_xy[2 * j + 1] = ty;
    }
  }
  return cnt;
}

/* GMapping::ComputeMap (gmapping.cc:171-242) accumulating into n/visits/acc arrays of mapSizeX*mapSizeY cells
 * (index x + y*mapSizeX).  Returns 0, or -1 when a ray leaves the map (the reference asserts there). */
int orc_gmap_compute_map(double cx, double cy, double delta, const int32_t layout[4], const double *ranges,
                         const double *angles, int nb, double laser_x, double laser_y, double max_range,
                         double max_use_range, int32_t *n, int32_t *visits, float *acc_x, float *acc_y) {
  int msx = layout[0], msy = layout[1], sx2 = layout[2], sy2 = layout[3];
  int p0x = (int)round((laser_x - cx) / delta) + sx2, p0y = (int)round((laser_y - cy) / delta) + sy2;
  if (p0x < 0 || p0y < 0 || p0x >= msx || p0y >= msy) return -1;
  int cap = msx + msy + 8;
  int32_t *line = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)cap);
  /* pass 1: validity (nothing is written if any ray leaves the map, like the pre-check in ref_gmapping.cpp) */
  for (int i = 0; i < nb; i++) {
    double d = ranges[i];
    if (d > max_range || d == 0.0 || !isfinite(d)) continue;
    if (d > max_use_range) d = max_use_range;
    double hx = laser_x + d * cos(angles[i]), hy = laser_y + d * sin(angles[i]);
    int p1x = (int)round((hx - cx) / delta) + sx2, p1y = (int)round((hy - cy) / delta) + sy2;
    if (p1x < 0 || p1y < 0 || p1x >= msx || p1y >= msy) { free(line); return -1; }
  }
  for (int i = 0; i < nb; i++) {
    double d = ranges[i];
    if (d > max_range || d == 0.0 || !isfinite(d)) continue;
    if (d > max_use_range) d = max_use_range;
    double hx = laser_x, hy = laser_y;
    hx += d * cos(angles[i]);
    hy += d * sin(angles[i]);
    int p1x = (int)round((hx - cx) / delta) + sx2, p1y = (int)round((hy - cy) / delta) + sy2;
    int cnt = orc_gmap_grid_line(p0x, p0y, p1x, p1y, line, cap);
    for (int k = 0; k < cnt - 1; k++) visits[line[2 * k] + (size_t)line[2 * k + 1] * msx]++;
    if (d < max_use_range) {
      size_t c = (size_t)p1x + (size_t)p1y * msx;
      acc_x[c] += (float)hx;
      acc_y[c] += (float)hy;
      n[c]++;
      visits[c] += 1;
    }
  }
  free(line);
  return 0;
}

/* GMapping::PublishMap thresholding (gmapping.cc:141-159) into a width x height int8 array (MAP_IDX = width*y + x);
 * cells never written keep 0 like the resized std::vector */
void orc_gmap_ros(const int32_t layout[4], const int32_t *n, const int32_t *visits, double occ_thresh, int width,
                  int height, int8_t *out) {
  int msx = layout[0], msy = layout[1];
  memset(out, 0, (size_t)width * height);
  for (int x = 0; x < msx; x++)
    for (int y = 0; y < msy; y++) {
      size_t c = (size_t)x + (size_t)y * msx;
      double occ = visits[c] ? (double)n[c] * 1 / (double)visits[c] : -1;
      size_t o = (size_t)width * y + x;
      if (o >= (size_t)width * height) continue;
      out[o] = occ < 0 ? -1 : (occ > occ_thresh ? 100 : 0);
    }
}
