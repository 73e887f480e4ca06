// ROS-shaped input adapters (host C++, no device work): the conversions the lesson nodes apply to sensor_msgs/LaserScan
// fields before the hot path sees them.  ROS itself is out of scope.
//   SlamKarto::getLaser / addScan                       lesson6/src/karto_slam.cc:323-395, 407-434
//   LaserRangeFinder::SetRangeThreshold / Update        lesson6/lib/open_karto/include/open_karto/Karto.h:3778-3787, 4152-4161
//   HectorMappingRos::rosPointCloudToDataContainer      lesson4/src/hector_mapping/hector_slam.cc:320-362
#include <cmath>

#include "common.cuh"

using namespace b2s;

extern "C" {

b2s_status b2s_ros_karto_laser(const b2s_laser_scan_msg *scan, const double laser_pose_in_base[3], double use_scan_range,
                               b2s_laser *out) {
  if (!scan || !laser_pose_in_base || !out) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  // the setters take kt_double: the float32 message fields widen exactly
  const double min_angle = scan->angle_min, max_angle = scan->angle_max, res = scan->angle_increment;
  const double min_range = scan->range_min, max_range = scan->range_max;
  if (!(res != 0.0)) B2S_FAIL(B2S_ERR_BAD_PARAMS, "angle_increment is zero");
  out->min_angle = min_angle;
  out->angular_resolution = res;
  out->min_range = min_range;
  out->max_range = max_range;
  // math::Clip(rangeThreshold, min, max) (Karto.h:3781)
  out->range_threshold = use_scan_range < min_range ? min_range : (use_scan_range > max_range ? max_range : use_scan_range);
  out->offset_pose[0] = laser_pose_in_base[0];
  out->offset_pose[1] = laser_pose_in_base[1];
  out->offset_pose[2] = laser_pose_in_base[2];
  out->n_readings = (int32_t)cast_u32(kround((max_angle - min_angle) / res));  // Update(): no + 1
  out->reserved = 0;
  if (out->n_readings != scan->n_ranges)
    B2S_FAIL(B2S_ERR_BAD_PARAMS, "LaserScan carries a different number of ranges than Round((angle_max - angle_min) / angle_increment): "
                                 "LaserRangeFinder::Validate rejects such scans");
  return B2S_OK;
}

b2s_status b2s_ros_karto_readings(const b2s_laser_scan_msg *scan, int inverted, double *out) {
  if (!scan || !out || (scan->n_ranges > 0 && !scan->ranges)) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  const int n = scan->n_ranges;
  for (int i = 0; i < n; i++) out[i] = (double)scan->ranges[inverted ? n - 1 - i : i];
  return B2S_OK;
}

b2s_status b2s_ros_hector_points(const float *pts, int n, const float laser_in_base[4], float scale_to_map, float min_dist,
                                 float max_dist, double use_max_scan_range, float z_min, float z_max, float *out_points,
                                 float out_origo[2], int32_t *out_n) {
  if (!laser_in_base || !out_origo || !out_n || n < 0 || (n > 0 && (!pts || !out_points))) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  const float sqr_min = min_dist * min_dist, sqr_max = max_dist * max_dist;  // hector_slam.cc:151-155 (from doubles, cast to float)
  // tf::Transform of the laser in base_link: rotation about z by yaw (tfScalar = double), translation (x, y, z)
  const double lx = laser_in_base[0], ly = laser_in_base[1], lz = laser_in_base[2];
  const double c = cos((double)laser_in_base[3]), s = sin((double)laser_in_base[3]);
  out_origo[0] = (float)lx * scale_to_map;  // Eigen::Vector2f(laserPos.x(), laserPos.y()) * scaleToMap
  out_origo[1] = (float)ly * scale_to_map;
  int k = 0;
  for (int i = 0; i < n; i++) {
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    const float dist_sqr = x * x + y * y;
    if (!((dist_sqr > sqr_min) && (dist_sqr < sqr_max))) continue;
    if ((x < 0.0f) && (dist_sqr < 0.50f)) continue;
    if ((double)dist_sqr > use_max_scan_range * use_max_scan_range) continue;
    const double bx = c * (double)x - s * (double)y + lx, by = s * (double)x + c * (double)y + ly, bz = (double)z + lz;
    const float z_laser = (float)(bz - lz);
    if (z_laser > z_min && z_laser < z_max) {
      out_points[2 * k] = (float)bx * scale_to_map;
      out_points[2 * k + 1] = (float)by * scale_to_map;
      k++;
    }
  }
  *out_n = k;
  return B2S_OK;
}

}  // extern "C"
