// Shared host/device helpers for libb200slam (sm_100a only).
//
// Arithmetic rule of this library: every double/float expression that feeds an integer grid index
// follows the reference's operation order with no FMA contraction (the library is compiled with
// -fmad=false and the critical expressions use explicit __dmul_rn/__dadd_rn), because the reference's
// x86-64 build has none either (SURVEY.md §7 "libm vs CUDA math").
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/b200slam.h"

namespace b2s {

void set_last_error(const std::string &s);
std::string &last_error_ref();

#define B2S_CUDA_CHECK(expr)                                                                          \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      ::b2s::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +        \
                            __FILE__ + ":" + std::to_string(__LINE__) + ")");                         \
      return B2S_ERR_CUDA;                                                                            \
    }                                                                                                 \
  } while (0)

// same, running `cleanup` (e.g. destroying a half-built handle) before returning
#define B2S_CUDA_CHECK_CLEAN(cleanup, expr)                                                            \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      const std::string _why = std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +      \
                               __FILE__ + ":" + std::to_string(__LINE__) + ")";                       \
      cleanup;                                                                                        \
      ::b2s::set_last_error(_why);                                                                    \
      return B2S_ERR_CUDA;                                                                            \
    }                                                                                                 \
  } while (0)

#define B2S_FAIL(code, msg)        \
  do {                             \
    ::b2s::set_last_error(msg);    \
    return (code);                 \
  } while (0)

constexpr double KT_PI = 3.14159265358979323846;   // Math.h:32
constexpr double KT_2PI = 6.28318530717958647692;  // Math.h:33
constexpr double KT_TOLERANCE = 1e-06;             // Math.h:41
constexpr int32_t INVALID_SCAN = 2147483647;       // Math.h:47
constexpr int GRID_OCCUPIED = 100;                 // Karto.h:4196
constexpr int GRID_FREE = 255;                     // Karto.h:4197
constexpr double MAX_VARIANCE = 500.0;             // Mapper.cpp:36
constexpr double DISTANCE_PENALTY_GAIN = 0.2;      // Mapper.cpp:37
constexpr double ANGLE_PENALTY_GAIN = 0.2;         // Mapper.cpp:38

// math::Round (Math.h:87-90): half away from zero
__host__ __device__ __forceinline__ double kround(double v) {
#ifdef __CUDA_ARCH__
  // identical values: ceil(v - 0.5) == -floor(-v + 0.5) for v < 0 (negation is exact); one rounding op instead of two
  return copysign(floor(fabs(v) + 0.5), v);
#else
  return v >= 0.0 ? floor(v + 0.5) : ceil(v - 0.5);
#endif
}

// static_cast<kt_int32s>(double) as the reference's x86-64 build performs it (cvttsd2si: out-of-range and NaN
// give INT_MIN); CUDA's own conversion saturates instead.
__host__ __device__ __forceinline__ int32_t cast_i32(double d) {
  return (d >= -2147483648.0 && d < 2147483648.0) ? (int32_t)d : (int32_t)0x80000000;
}
// static_cast<kt_int32u>(double): cvttsd2si 64-bit then truncation
__host__ __device__ __forceinline__ uint32_t cast_u32(double d) {
  return (d >= -9223372036854775808.0 && d < 9223372036854775808.0) ? (uint32_t)(long long)d : 0u;
}

// math::DoubleEqual (Math.h:135-139)
__host__ __device__ __forceinline__ bool double_equal(double a, double b) {
  double delta = a - b;
  return delta < 0.0 ? delta >= -KT_TOLERANCE : delta <= KT_TOLERANCE;
}

// math::NormalizeAngle (Math.h:182-211)
__host__ __device__ inline double normalize_angle(double angle) {
  while (angle < -KT_PI) {
    if (angle < -KT_2PI)
      angle += (double)cast_u32(angle / -KT_2PI) * KT_2PI;
    else
      angle += KT_2PI;
  }
  while (angle > KT_PI) {
    if (angle > KT_2PI)
      angle -= (double)cast_u32(angle / KT_2PI) * KT_2PI;
    else
      angle -= KT_2PI;
  }
  return angle;
}

// math::NormalizeAngleDifference (Math.h:221-234)
__host__ __device__ inline double normalize_angle_difference(double minuend, double subtrahend) {
  while (minuend - subtrahend < -KT_PI) minuend += KT_2PI;
  while (minuend - subtrahend > KT_PI) minuend -= KT_2PI;
  return minuend;
}

__host__ __device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }  // math::Maximum

// CoordinateConverter::WorldToGrid, one axis (Karto.h:4237-4252)
__host__ __device__ __forceinline__ int32_t world_to_grid_1(double w, double offset, double scale) {
  return cast_i32(kround((w - offset) * scale));
}

// number of search steps: static_cast<kt_int32u>(math::Round(off * 2.0 / res) + 1) (Mapper.cpp:339-341,361)
__host__ __device__ __forceinline__ int n_steps(double off, double res) {
  return (int)cast_u32(kround(off * 2.0 / res) + 1);
}

// LocalizedRangeScan::GetSensorAt = Transform(robot).TransformPose(offset) (Karto.h:5310-5313, 2860-2935,
// Matrix3::FromAxisAngle 2392-2421 with axis (0,0,1), Matrix3*Pose2 2574-2583)
__host__ __device__ inline void sensor_pose_of(const double robot[3], const double offset[3], double out[3]) {
  double m00, m01, m02, m10, m11, m12, tx, ty, th;
  if (robot[0] == 0.0 && robot[1] == 0.0 && robot[2] == 0.0) {  // rPose1 == rPose2 (Karto.h:2911-2917)
    m00 = 1; m01 = 0; m02 = 0; m10 = 0; m11 = 1; m12 = 0; tx = 0; ty = 0; th = 0;
  } else {
    double radians = robot[2] - 0.0;
    double c = cos(radians), s = sin(radians), omc = 1.0 - c;
    m00 = 0.0 * omc + c;
    m01 = 0.0 * 0.0 * omc - 1.0 * s;
    m02 = 0.0 * 1.0 * omc + 0.0 * s;
    m10 = 0.0 * 0.0 * omc + 1.0 * s;
    m11 = 0.0 * omc + c;
    m12 = 0.0 * 1.0 * omc - 0.0 * s;
    tx = robot[0]; ty = robot[1]; th = robot[2] - 0.0;
  }
  double rx = m00 * offset[0] + m01 * offset[1] + m02 * offset[2];
  double ry = m10 * offset[0] + m11 * offset[1] + m12 * offset[2];
  out[0] = tx + rx;
  out[1] = ty + ry;
  out[2] = normalize_angle(offset[2] + th);
}

struct StreamRef {
  cudaStream_t s = nullptr;
  bool owned = false;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is one value per (kernel, device), shared by every handle of the process.
// Handles of different sizes used from different threads must therefore never LOWER it between another thread's set
// and launch: keep a process-wide running maximum per (kernel, device) and only ever raise the attribute.
template <class K>
inline cudaError_t raise_dyn_smem(K kern, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, size_t> seen;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t &cur = seen[std::make_pair(reinterpret_cast<const void *>(kern), dev)];
  if (bytes <= cur) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

// Calls that stage through cudaMallocAsync / cudaFreeAsync and then synchronise would hand their memory back to the driver
// every time (the default pool's release threshold is 0) and pay a fresh allocation on the next call (~1 ms).  The
// library's create functions raise the threshold of the device's default pool once, so freed blocks stay cached.
inline void keep_pool_memory(int device) {
  static std::mutex mu;
  static std::map<int, bool> done;
  std::lock_guard<std::mutex> lock(mu);
  if (done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long keep = 1ull << 32;  // up to 4 GiB of freed staging blocks stay in the pool
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  done[device] = true;
}

// NVTX range around a host entry point (SURVEY.md §5 tracing): shows up as "b2s:<name>" in Nsight timelines; nvtx3 is
// header-only and costs a predictable branch when no tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};
#define B2S_NVTX(name) ::b2s::NvtxRange _b2s_nvtx_range("b2s:" name)

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace b2s
