/* TEST INFRASTRUCTURE — prototypes of the CPU restatements (see karto_oracle.c header). */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H
#include <stdint.h>
#include "../include/b200slam.h"

double orc_round(double v);
double orc_normalize_angle(double angle);
int orc_matcher_layout(const b2s_matcher_params *p, b2s_grid_info *g);
void orc_smear_kernel(double resolution_param, double smear, int ksize, uint8_t *K);
void orc_sensor_pose(const double robot[3], const double offset[3], double out[3]);
void orc_point_readings(const b2s_laser *l, const double *ranges, const double robot_pose[3], double *out_xy);
void orc_scan_bbox(const b2s_laser *l, const double *ranges, const double robot_pose[3], double bbox[4]);
int orc_find_valid_points(const double *pts_xy, int n, const double viewpoint[2], double *out_xy);
void orc_grid_offset(const b2s_grid_info *g, double resolution_param, const double sensor_pose[3], double off[2]);
void orc_add_scan(const b2s_grid_info *g, double resolution_param, const double grid_off[2], uint8_t *grid,
                  const uint8_t *K, const double *pts_xy, int n, const double viewpoint[2], double *scratch_xy);
void orc_add_scans(const b2s_matcher_params *p, const b2s_laser *l, const b2s_grid_info *g, const double grid_off[2],
                   uint8_t *grid, int n_base, const double *base_ranges, const double *base_poses,
                   const double viewpoint[2]);
int orc_n_steps(double off, double res);
void orc_compute_offsets(const b2s_grid_info *g, double resolution_param, const double grid_off[2],
                         const double *ranges, const double *pts_xy, int n, const double sensor_pose[3],
                         double angle_center, double angle_offset, double angle_res, int32_t *lut);
int orc_response_sums(const b2s_grid_info *g, double resolution_param, const double grid_off[2], const uint8_t *grid,
                      const int32_t *lut, int n, const double center[3], const b2s_search *s, int32_t *out);
int orc_correlate_scan(const b2s_matcher_params *p, const b2s_grid_info *g, const double grid_off[2],
                       const uint8_t *grid, const double *ranges, const double *pts_xy, int n,
                       const double sensor_pose[3], const double center[3], const b2s_search *s,
                       b2s_match_result *result, int32_t *sums_out);
int orc_match_scan(const b2s_matcher_params *p, const b2s_laser *l, const double *ranges, const double robot_pose[3],
                   int n_base, const double *base_ranges, const double *base_poses, int do_penalize, int do_refine,
                   b2s_match_result *result, uint8_t *grid_out, double *grid_off_out);
void orc_occ_dimensions(const b2s_laser *l, int n_scans, const double *ranges, const double *poses,
                        double resolution, b2s_occ_grid_info *info);
int orc_trace_line_cells(int w, int h, int x0, int y0, int x1, int y1, int32_t *out_xy, int cap);
void orc_occ_create_from_scans(const b2s_laser *l, int n_scans, const double *ranges, const double *poses,
                               b2s_occ_grid_info *info, uint32_t *pass, uint32_t *hit, uint8_t *cells);
#endif
