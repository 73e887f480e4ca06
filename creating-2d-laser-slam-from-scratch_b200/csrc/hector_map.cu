// K2a + K3 — lesson4 Hector log-odds grid map update and Gauss-Newton scan-to-map alignment on B200 (sm_100a).
// Product code: CUDA only.
//
// Reference behaviour (paths relative to /root/reference/lesson4/include/lesson4/hector_mapping):
//   OccGridMapBase::updateByScan / updateLineBresenhami / bresenham2D / bresenhamCellFree/Occ   map/OccGridMapBase.h:118-168, 220-330
//   LogOddsCell, GridMapLogOddsFunctions                                                       map/GridMapLogOdds.h:37-161
//   GridMapBase ctor, setMapTransformation, getMapCoordsPose, getWorldCoordsPose               map/GridMapBase.h:54-66, 229-286
//   OccGridMapUtil::getCompleteHessianDerivs / interpMapValueWithDerivatives                   map/OccGridMapUtil.h:77-228
//   ScanMatcher::matchData / estimateTransformationLogLh                                       matcher/ScanMatcher.h:60-141
//
// K2a is sequential ACROSS scans (the <50 clamp and the once-per-scan stamps) but parallel WITHIN a scan.  The
// reference's in-scan order dependence is reproduced exactly with two passes over the rays (one warp per beam,
// closed-form Bresenham: minor-axis steps after n major steps = floor((da/2 + n*db) / da)):
//   mark  : every traversed cell records the LOWEST beam index that frees it / ends on it (64-bit atomicMax of
//           (scan epoch << 32 | ~beam));
//   apply : the recorded winner beam applies the cell's update once: free-only cells get += logOddsFree; end cells
//           get the reference's "(v + f) - f" un-free rounding iff a lower-indexed beam had freed them first, then
//           += logOddsOccupied if v < 50 — identical floats to the sequential loop, whatever the execution order.
// The pose's cosf/sinf are taken on the host (glibc) so that all device arithmetic on the integer path is + - *
// in float32 without FMA: the truncated cell coordinates are bit-identical to the CPU.
// K3 runs the whole 1 + maxIterations Gauss-Newton loop of one grid level in ONE kernel launch: per-point terms in
// parallel, then thread 0 accumulates H and dTr in point order (the reference's float32 summation order).
//
// HBM layout per cell: logOdds f32, updateIndex i32 (the reference's LogOddsCell, split SoA), + two u64 stamps.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

using namespace b2s;

struct b2s_hector_map {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sx = 0, sy = 0;
  float cell_length = 0, scale_to_map = 0, off_x = 0, off_y = 0;
  float tw_lin = 0, tw_tx = 0, tw_ty = 0, wt_lin = 0, wt_tx = 0, wt_ty = 0;
  float lo_free = 0, lo_occ = 0;
  int curr_update_index = 0;
  unsigned int epoch = 0;
  float *d_lo = nullptr;
  int32_t *d_ui = nullptr;
  unsigned long long *d_free = nullptr, *d_occ = nullptr;
  float *d_pts = nullptr;
  size_t pts_cap = 0;
  float *d_out = nullptr;  // [3 pose + 9 H]
  unsigned long long *d_visits = nullptr;
  double last_ms[2] = {0, 0};
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // update begin/end, match begin/end
};

namespace b2s {

static float prob_to_log_odds(float prob) {  // GridMapLogOdds.h:153-157
  float odds = prob / (1.0f - prob);
  return logf(odds);
}

struct HcTransform {
  float c, s, mx, my;  // poseTransform = Translation2f(mx, my) * Rotation2Df(theta)
  int just_once;       // updateByScanJustOnce: end = begin + (int)round(p / 0.05) (OccGridMapBase.h:202-203)
};

__device__ __forceinline__ void hc_apply(const HcTransform &t, float px, float py, float &ox, float &oy) {
  ox = __fadd_rn(__fadd_rn(__fmul_rn(t.c, px), __fmul_rn(-t.s, py)), t.mx);
  oy = __fadd_rn(__fadd_rn(__fmul_rn(t.s, px), __fmul_rn(t.c, py)), t.my);
}

struct HcLine {
  bool ok;
  int x0, y0, x1, y1;
  unsigned int da, db;
  int err0, off_a, off_b;
};

// endpoints (OccGridMapBase.h:127-154) + updateLineBresenhami set-up (:220-258)
__device__ __forceinline__ HcLine hc_line(const HcTransform &t, int bx, int by, float px, float py, int sx, int sy) {
  HcLine L;
  float ex, ey;
  hc_apply(t, px, py, ex, ey);
  ex = __fadd_rn(ex, 0.5f);
  ey = __fadd_rn(ey, 0.5f);
  L.x0 = bx; L.y0 = by;
  L.x1 = (int)ex; L.y1 = (int)ey;  // Vector2f::cast<int>(): truncation
  if (t.just_once) {  // points in metres; float / double literal -> double, ::round, (int)
    L.x1 = bx + cast_i32(round((double)px / 0.05));
    L.y1 = by + cast_i32(round((double)py / 0.05));
  }
  L.ok = !(L.x0 == L.x1 && L.y0 == L.y1);
  if ((L.x0 < 0) || (L.x0 >= sx) || (L.y0 < 0) || (L.y0 >= sy)) L.ok = false;
  if ((L.x1 < 0) || (L.x1 >= sx) || (L.y1 < 0) || (L.y1 >= sy)) L.ok = false;
  const int dx = L.x1 - L.x0, dy = L.y1 - L.y0;
  const unsigned int adx = (unsigned int)abs(dx), ady = (unsigned int)abs(dy);
  const int odx = dx > 0 ? 1 : -1, ody = (dy > 0 ? 1 : -1) * sx;  // util::sign: sign(0) = -1
  if (adx >= ady) { L.da = adx; L.db = ady; L.err0 = (int)(adx / 2); L.off_a = odx; L.off_b = ody; }
  else { L.da = ady; L.db = adx; L.err0 = (int)(ady / 2); L.off_a = ody; L.off_b = odx; }
  return L;
}

// pass 1: record, per traversed cell, the lowest beam index that frees it / ends on it
__global__ void __launch_bounds__(256)
    k_hc_mark(const float *__restrict__ pts, int n, HcTransform t, int bx, int by, int sx, int sy,
              unsigned long long epoch_hi, unsigned long long *__restrict__ free_st,
              unsigned long long *__restrict__ occ_st) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp0; i < n; i += nwarps) {
    const HcLine L = hc_line(t, bx, by, pts[2 * i], pts[2 * i + 1], sx, sy);
    if (!L.ok) continue;
    const unsigned long long stamp = epoch_hi | (unsigned long long)(0xffffffffu - (unsigned int)i);
    const int start = L.y0 * sx + L.x0;
    for (unsigned int k = lane; k < L.da; k += 32) {  // bresenham2D: da cells from the start, end excluded
      const unsigned int inc = (unsigned int)(((unsigned long long)L.err0 + (unsigned long long)k * L.db) / L.da);
      const int off = start + (int)k * L.off_a + (int)inc * L.off_b;
      atomicMax(free_st + off, stamp);
    }
    if (lane == 0) atomicMax(occ_st + (L.y1 * sx + L.x1), stamp);
  }
}

// pass 2: the winner beam of each cell applies the reference's update exactly once
__global__ void __launch_bounds__(256)
    k_hc_apply(const float *__restrict__ pts, int n, HcTransform t, int bx, int by, int sx, int sy,
               unsigned long long epoch_hi, const unsigned long long *__restrict__ free_st,
               const unsigned long long *__restrict__ occ_st, float lo_free, float lo_occ, int mark_free, int mark_occ,
               float *__restrict__ lo, int32_t *__restrict__ ui, unsigned long long *__restrict__ visits) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long my_visits = 0;
  for (int i = warp0; i < n; i += nwarps) {
    const HcLine L = hc_line(t, bx, by, pts[2 * i], pts[2 * i + 1], sx, sy);
    if (!L.ok) continue;
    const unsigned long long stamp = epoch_hi | (unsigned long long)(0xffffffffu - (unsigned int)i);
    const int start = L.y0 * sx + L.x0;
    for (unsigned int k = lane; k < L.da; k += 32) {
      const unsigned int inc = (unsigned int)(((unsigned long long)L.err0 + (unsigned long long)k * L.db) / L.da);
      const int off = start + (int)k * L.off_a + (int)inc * L.off_b;
      my_visits++;
      if (free_st[off] == stamp && (occ_st[off] >> 32) != (epoch_hi >> 32)) {
        // freed this scan, never an end point: bresenhamCellFree (OccGridMapBase.h:302-312)
        lo[off] = __fadd_rn(lo[off], lo_free);
        ui[off] = mark_free;
      }
    }
    if (lane == 0) {
      my_visits++;
      const int off = L.y1 * sx + L.x1;
      if (occ_st[off] == stamp) {  // bresenhamCellOcc (OccGridMapBase.h:315-330), first beam ending here
        float v = lo[off];
        const unsigned long long fs = free_st[off];
        if ((fs >> 32) == (epoch_hi >> 32) && fs > stamp) {  // a LOWER beam index freed it first: set free, then unset
          v = __fadd_rn(v, lo_free);
          v = __fsub_rn(v, lo_free);
        }
        if (v < 50.0f) v = __fadd_rn(v, lo_occ);
        lo[off] = v;
        ui[off] = mark_occ;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) my_visits += __shfl_xor_sync(0xffffffffu, my_visits, d);
  if (lane == 0 && my_visits) atomicAdd(visits, my_visits);
}

// ---- K3: Gauss-Newton scan-to-map alignment (one grid level), whole iteration loop in one launch ----

__device__ __forceinline__ float hc_prob(const float *__restrict__ lo, int index) {  // GridMapLogOdds.h:136-140
  const float odds = expf(lo[index]);
  return odds / (odds + 1.0f);
}

__device__ __forceinline__ void hc_interp(const float *__restrict__ lo, int sx, int sy, float x, float y, float out[3]) {
  const float lim_x = (float)sx - 2.0f, lim_y = (float)sy - 2.0f;  // setMapCellDims: dims - 2
  if (x < 0.0f || x > lim_x || y < 0.0f || y > lim_y) { out[0] = out[1] = out[2] = 0.0f; return; }
  const int ix = (int)x, iy = (int)y;
  const float fx = x - (float)ix, fy = y - (float)iy;
  const int index = iy * sx + ix;
  const float i0 = hc_prob(lo, index), i1 = hc_prob(lo, index + 1);
  const float i2 = hc_prob(lo, index + sx), i3 = hc_prob(lo, index + sx + 1);
  const float dx1 = i0 - i1, dx2 = i2 - i3, dy1 = i0 - i2, dy2 = i1 - i3;
  const float xfi = 1.0f - fx, yfi = 1.0f - fy;
  out[0] = ((i0 * xfi + i1 * fx) * yfi) + ((i2 * xfi + i3 * fx) * fy);
  out[1] = -((dx1 * yfi) + (dx2 * fy));
  out[2] = -((dy1 * xfi) + (dy2 * fx));
}

__device__ __forceinline__ void hs_interp(const float *__restrict__ prob, int sx, int sy, float x, float y, float out[3]) {
  const float lim_x = (float)sx - 2.0f, lim_y = (float)sy - 2.0f;  // setMapCellDims: dims - 2
  if (x < 0.0f || x > lim_x || y < 0.0f || y > lim_y) { out[0] = out[1] = out[2] = 0.0f; return; }
  const int ix = (int)x, iy = (int)y;
  const float fx = x - (float)ix, fy = y - (float)iy;
  const int index = iy * sx + ix;
  const float i0 = prob[index], i1 = prob[index + 1], i2 = prob[index + sx], i3 = prob[index + sx + 1];
  const float dx1 = i0 - i1, dx2 = i2 - i3, dy1 = i0 - i2, dy2 = i1 - i3;
  const float xfi = 1.0f - fx, yfi = 1.0f - fy;
  out[0] = ((i0 * xfi + i1 * fx) * yfi) + ((i2 * xfi + i3 * fx) * fy);
  out[1] = -((dx1 * yfi) + (dx2 * fy));
  out[2] = -((dy1 * xfi) + (dy2 * fx));
}

__device__ inline void hc_inv3_mul(const float m[9], const float v[3], float out[3]) {  // Matrix3f::inverse() * v
  const float c00 = m[4] * m[8] - m[5] * m[7], c10 = m[5] * m[6] - m[3] * m[8], c20 = m[3] * m[7] - m[4] * m[6];
  const float det = c00 * m[0] + (c10 * m[1] + c20 * m[2]);  /* Eigen's unrolled 3-term redux: a0 + (a1 + a2) */
  const float invdet = 1.0f / det;
  float inv[9];
  inv[0] = c00 * invdet; inv[3] = c10 * invdet; inv[6] = c20 * invdet;
  inv[1] = (m[2] * m[7] - m[1] * m[8]) * invdet;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) * invdet;
  inv[7] = (m[1] * m[6] - m[0] * m[7]) * invdet;
  inv[2] = (m[1] * m[5] - m[2] * m[4]) * invdet;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) * invdet;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) * invdet;
  for (int r = 0; r < 3; r++) out[r] = inv[3 * r] * v[0] + (inv[3 * r + 1] * v[1] + inv[3 * r + 2] * v[2]);
}

constexpr int GN_THREADS = 512;
constexpr int GN_CHUNK = 1024;  // points staged per pass

__global__ void __launch_bounds__(GN_THREADS)
    k_hc_match(const float *__restrict__ lo, int sx, int sy, const float *__restrict__ pts, int n, float e0, float e1,
               float e2, int iterations, float wt_lin, float wt_tx, float wt_ty, float *__restrict__ out) {
  __shared__ float terms[GN_CHUNK][4];  // t1, t2, rotDeriv, funVal per point
  __shared__ float est[3];
  __shared__ float H[9], dTr[3];
  if (threadIdx.x == 0) { est[0] = e0; est[1] = e1; est[2] = e2; }
  __syncthreads();
  for (int it = 0; it < iterations; it++) {
    const float c = cosf(est[2]), s = sinf(est[2]);
    if (threadIdx.x == 0) {
      for (int q = 0; q < 9; q++) H[q] = 0.0f;
      dTr[0] = dTr[1] = dTr[2] = 0.0f;
    }
    for (int base = 0; base < n; base += GN_CHUNK) {
      const int cnt = min(GN_CHUNK, n - base);
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += GN_THREADS) {
        const float px = pts[2 * (base + i)], py = pts[2 * (base + i) + 1];
        const float tx = (c * px + (-s) * py) + est[0], ty = (s * px + c * py) + est[1];
        float t[3];
        hc_interp(lo, sx, sy, tx, ty, t);
        terms[i][0] = t[1];
        terms[i][1] = t[2];
        terms[i][2] = ((-s * px - c * py) * t[1] + (c * px - s * py) * t[2]);
        terms[i][3] = 1.0f - t[0];
      }
      __syncthreads();
      if (threadIdx.x == 0) {  // getCompleteHessianDerivs accumulation, in point order (OccGridMapUtil.h:99-126)
        for (int i = 0; i < cnt; i++) {
          const float t1 = terms[i][0], t2 = terms[i][1], rot = terms[i][2], fun = terms[i][3];
          dTr[0] += t1 * fun;
          dTr[1] += t2 * fun;
          dTr[2] += rot * fun;
          H[0] += t1 * t1;
          H[4] += t2 * t2;
          H[8] += rot * rot;
          H[1] += t1 * t2;
          H[2] += t1 * rot;
          H[5] += t2 * rot;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      H[3] = H[1]; H[6] = H[2]; H[7] = H[5];
      if (H[0] != 0.0f && H[4] != 0.0f) {  // estimateTransformationLogLh (ScanMatcher.h:107-141)
        float dir[3];
        hc_inv3_mul(H, dTr, dir);
        if (dir[2] > 0.2f) dir[2] = 0.2f;
        else if (dir[2] < -0.2f) dir[2] = -0.2f;
        est[0] += dir[0]; est[1] += dir[1]; est[2] += dir[2];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // util::normalize_angle (UtilFunctions.h:36-48): double fmod, float result
    const double two_pi = 2.0f * 3.14159265358979323846;
    float a = (float)fmod(fmod((double)est[2], two_pi) + two_pi, two_pi);
    if (a > 3.14159265358979323846) a = (float)((double)a - two_pi);  // `a -= 2.0f*M_PI` promotes to double
    // getWorldCoordsPose (GridMapBase.h:229-233)
    out[0] = (wt_lin * est[0] + (-0.0f) * est[1]) + wt_tx;
    out[1] = ((-0.0f) * est[0] + wt_lin * est[1]) + wt_ty;
    out[2] = a;
    for (int q = 0; q < 9; q++) out[3 + q] = H[q];
  }
}

// HectorMappingRos::publishMap (hector_slam.cc:254-317): free -> 0, occupied -> 100, else -1
__global__ void k_hc_ros(const float *__restrict__ lo, int n, int8_t *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = lo[i];
  out[i] = v < 0.0f ? 0 : (v > 0.0f ? 100 : -1);
}


// ---- lesson4 front end (HectorSlamProcessor): all pyramid levels per launch ------------------------------------

struct HsLevel {  // one MapRepMultiMap level as the kernels see it
  float *prob;  // getGridProbability(cell) = e^lo / (e^lo + 1) (GridMapLogOdds.h:136-140), refreshed whenever lo changes —
                // the device-side form of the reference's per-scan GridMapCacheArray (4 expf + 4 divisions per point less)
  float *lo;
  int32_t *ui;
  unsigned long long *free_st, *occ_st;
  const float *pts;  // this level's DataContainer (level 0: the uploaded scan; l > 0: scaled by 1 / 2^l)
  int n, sx, sy;
  float tw_lin, tw_tx, tw_ty, wt_lin, wt_tx, wt_ty;
  int iterations;  // 1 + maxIterations of MapRepMultiMap::matchData (:144-166): 1+5 on level 0, 1+3 above
};
struct HsLevels {
  HsLevel l[B2S_HECTOR_MAX_LEVELS];
  int count;
};
struct HsUpdate {  // per-level update parameters of one scan (host: getMapCoordsPose, glibc cosf/sinf, begin cell)
  HcTransform t[B2S_HECTOR_MAX_LEVELS];
  int bx[B2S_HECTOR_MAX_LEVELS], by[B2S_HECTOR_MAX_LEVELS];
  unsigned long long epoch_hi[B2S_HECTOR_MAX_LEVELS];
  int mark_free[B2S_HECTOR_MAX_LEVELS], mark_occ[B2S_HECTOR_MAX_LEVELS];
};

// MapRepMultiMap::matchData (:144-166) -> ScanMatcher::matchData (ScanMatcher.h:60-98) on every level, coarsest first,
// in ONE launch of one CTA (the problem is a latency chain of 14 dependent Gauss-Newton iterations over ~1000 points,
// not a throughput problem).  The scan is staged in shared memory once; level l reads it scaled by 2^-l on the fly
// (DataPointContainer::setFrom, DataPointContainer.h:46-59: one float multiply per coordinate, exact) and also
// writes that scaled copy to the level's device container for the update kernels.  Per iteration every thread
// evaluates its points (bilinear map value + gradient: 4 cell reads + 4 expf, all loads issued up front), the 9 sums
// of getCompleteHessianDerivs are reduced through shared memory in a fixed order (deterministic run to run; they
// differ from the reference's point-order float sums in the last bits, far inside the 1e-4 contract — device
// cosf/sinf/expf differ from glibc's in the last ulp anyway), and every thread solves the 3x3 system redundantly so
// the new estimate needs no broadcast.
constexpr int HS_THREADS = 1024;
constexpr int HS_PPT = 2;                        // points per thread held in registers
constexpr int HS_SMEM_PTS = HS_THREADS * HS_PPT; // scans up to 2048 points take the staged path
constexpr int HS_SMEM_BYTES = (int)sizeof(float2) * HS_SMEM_PTS + 9 * HS_THREADS * (int)sizeof(float);

__device__ __forceinline__ void hs_point_terms(const float *__restrict__ prob, int sx, int sy, float2 p, float c, float s,
                                               float e0, float e1, float a[9]) {
  const float tx = (c * p.x + (-s) * p.y) + e0, ty = (s * p.x + c * p.y) + e1;
  float t[3];
  hs_interp(prob, sx, sy, tx, ty, t);
  const float rot = ((-s * p.x - c * p.y) * t[1] + (c * p.x - s * p.y) * t[2]);
  const float fun = 1.0f - t[0];
  a[0] += t[1] * fun; a[1] += t[2] * fun; a[2] += rot * fun;    // dTr
  a[3] += t[1] * t[1]; a[4] += t[2] * t[2]; a[5] += rot * rot;  // H00 H11 H22
  a[6] += t[1] * t[2]; a[7] += t[1] * rot; a[8] += t[2] * rot;  // H01 H02 H12
}

__global__ void __launch_bounds__(HS_THREADS)
    k_hs_match(HsLevels L, float w0, float w1, float w2, float *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char hs_smem[];  // HS_SMEM_BYTES, opt-in (> 48 KB)
  float2 *spts = reinterpret_cast<float2 *>(hs_smem);                                                   // [HS_SMEM_PTS]
  float(*part)[HS_THREADS] = reinterpret_cast<float(*)[HS_THREADS]>(hs_smem + sizeof(float2) * HS_SMEM_PTS);  // [9][HS_THREADS]
  __shared__ float tot[9];
  __shared__ float bc[5];  // new estimate + cos / sin of its heading, published by thread 0
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = L.l[0].n;
  const bool staged = n <= HS_SMEM_PTS;
  const float2 *gp = reinterpret_cast<const float2 *>(L.l[0].pts);
  if (staged)
    for (int i = tid; i < n; i += HS_THREADS) spts[i] = gp[i];
  float world0 = w0, world1 = w1, world2 = w2;  // every thread carries the same estimate
  float H[9];
  bool any = false;
#pragma unroll
  for (int q = 0; q < 9; q++) H[q] = 0.0f;
  __syncthreads();
  float factor = 1.0f;
  for (int l = 1; l < L.count; l++) factor *= 0.5f;  // static_cast<float>(1.0 / pow(2.0, l)) is exactly 2^-l
  for (int lv = L.count - 1; lv >= 0; lv--, factor *= 2.0f) {
    const HsLevel m = L.l[lv];
    if (n == 0) break;  // ScanMatcher.h:66,97: no data -> the begin estimate is returned unchanged
    float2 p[HS_PPT];
#pragma unroll
    for (int j = 0; j < HS_PPT; j++) {
      const int i = tid + j * HS_THREADS;
      p[j] = make_float2(0.0f, 0.0f);
      if (staged && i < n) {
        p[j] = make_float2(__fmul_rn(spts[i].x, factor), __fmul_rn(spts[i].y, factor));
        if (lv > 0) reinterpret_cast<float2 *>(const_cast<float *>(m.pts))[i] = p[j];
      }
    }
    if (!staged && lv > 0)
      for (int i = tid; i < n; i += HS_THREADS)
        reinterpret_cast<float2 *>(const_cast<float *>(m.pts))[i] = make_float2(__fmul_rn(gp[i].x, factor), __fmul_rn(gp[i].y, factor));
    // getMapCoordsPose (GridMapBase.h:238-242)
    float e0 = (m.tw_lin * world0 + 0.0f * world1) + m.tw_tx;
    float e1 = (0.0f * world0 + m.tw_lin * world1) + m.tw_ty;
    float e2 = world2;
    float c = cosf(e2), s = sinf(e2);
    for (int it = 0; it < m.iterations; it++) {
      float a[9];
#pragma unroll
      for (int q = 0; q < 9; q++) a[q] = 0.0f;
      if (staged) {
#pragma unroll
        for (int j = 0; j < HS_PPT; j++)
          if (tid + j * HS_THREADS < n) hs_point_terms(m.prob, m.sx, m.sy, p[j], c, s, e0, e1, a);
      } else {
        for (int i = tid; i < n; i += HS_THREADS)
          hs_point_terms(m.prob, m.sx, m.sy, make_float2(__fmul_rn(gp[i].x, factor), __fmul_rn(gp[i].y, factor)), c, s, e0, e1, a);
      }
#pragma unroll
      for (int q = 0; q < 9; q++) part[q][tid] = a[q];
      __syncthreads();
      if (warp < 9) {  // warp q sums quantity q: 32 conflict-free reads per lane, then 5 shuffles
        float v = 0.0f;
#pragma unroll
        for (int j = 0; j < HS_THREADS / 32; j++) v += part[warp][lane + 32 * j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) tot[warp] = v;
      }
      __syncthreads();
      if (tid == 0) {  // one thread solves and publishes the new estimate with its cosine / sine (the kernel is issue-bound:
                       // 32 warps repeating the 3x3 solve and the range reductions cost more than a third barrier)
        const float dTr[3] = {tot[0], tot[1], tot[2]};
        float Hm[9];
        Hm[0] = tot[3]; Hm[4] = tot[4]; Hm[8] = tot[5];
        Hm[1] = Hm[3] = tot[6]; Hm[2] = Hm[6] = tot[7]; Hm[5] = Hm[7] = tot[8];
        float n0 = e0, n1 = e1, n2 = e2;
        if (Hm[0] != 0.0f && Hm[4] != 0.0f) {  // estimateTransformationLogLh (ScanMatcher.h:107-141)
          float dir[3];
          hc_inv3_mul(Hm, dTr, dir);
          if (dir[2] > 0.2f) dir[2] = 0.2f;
          else if (dir[2] < -0.2f) dir[2] = -0.2f;
          n0 += dir[0]; n1 += dir[1]; n2 += dir[2];
        }
        bc[0] = n0; bc[1] = n1; bc[2] = n2; bc[3] = cosf(n2); bc[4] = sinf(n2);
      }
      __syncthreads();
      e0 = bc[0]; e1 = bc[1]; e2 = bc[2]; c = bc[3]; s = bc[4];
    }
    H[0] = tot[3]; H[4] = tot[4]; H[8] = tot[5];  // the Hessian of the level's last iteration (covMatrix = H)
    H[1] = H[3] = tot[6]; H[2] = H[6] = tot[7]; H[5] = H[7] = tot[8];
    {
      const double two_pi = 2.0f * 3.14159265358979323846;  // util::normalize_angle (UtilFunctions.h:36-48)
      float a = (float)fmod(fmod((double)e2, two_pi) + two_pi, two_pi);
      if (a > 3.14159265358979323846) a = (float)((double)a - two_pi);
      world0 = (m.wt_lin * e0 + (-0.0f) * e1) + m.wt_tx;  // getWorldCoordsPose (GridMapBase.h:229-233)
      world1 = ((-0.0f) * e0 + m.wt_lin * e1) + m.wt_ty;
      world2 = a;
      any = true;
    }
  }
  if (tid == 0) {
    out[0] = world0; out[1] = world1; out[2] = world2;
    for (int q = 0; q < 9; q++) out[3 + q] = H[q];  // covMatrix = H of the last level matched (level 0)
    out[12] = any ? 1.0f : 0.0f;
  }
}

__device__ __forceinline__ float hs_prob_of(float lo) {  // getGridProbability (GridMapLogOdds.h:136-140)
  const float odds = expf(lo);
  return odds / (odds + 1.0f);
}

// MapRepMultiMap::updateByScan (:174-191): mark / apply passes of every level in one launch each (blockIdx.y = level)
__global__ void __launch_bounds__(256) k_hs_mark(HsLevels L, HsUpdate U) {
  const int lv = blockIdx.y;
  const HsLevel &m = L.l[lv];
  const HcTransform t = U.t[lv];
  const int bx = U.bx[lv], by = U.by[lv];
  const unsigned long long epoch_hi = U.epoch_hi[lv];
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp0; i < m.n; i += nwarps) {
    const HcLine ln = hc_line(t, bx, by, m.pts[2 * i], m.pts[2 * i + 1], m.sx, m.sy);
    if (!ln.ok) continue;
    const unsigned long long stamp = epoch_hi | (unsigned long long)(0xffffffffu - (unsigned int)i);
    const int start = ln.y0 * m.sx + ln.x0;
    for (unsigned int k = lane; k < ln.da; k += 32) {
      const unsigned int inc = (unsigned int)(((unsigned long long)ln.err0 + (unsigned long long)k * ln.db) / ln.da);
      atomicMax(m.free_st + (start + (int)k * ln.off_a + (int)inc * ln.off_b), stamp);
    }
    if (lane == 0) atomicMax(m.occ_st + (ln.y1 * m.sx + ln.x1), stamp);
  }
}

__global__ void __launch_bounds__(256)
    k_hs_apply(HsLevels L, HsUpdate U, float lo_free, float lo_occ, unsigned long long *__restrict__ visits) {
  const int lv = blockIdx.y;
  const HsLevel &m = L.l[lv];
  const HcTransform t = U.t[lv];
  const int bx = U.bx[lv], by = U.by[lv];
  const unsigned long long epoch_hi = U.epoch_hi[lv];
  const int mark_free = U.mark_free[lv], mark_occ = U.mark_occ[lv];
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long my_visits = 0;
  for (int i = warp0; i < m.n; i += nwarps) {
    const HcLine ln = hc_line(t, bx, by, m.pts[2 * i], m.pts[2 * i + 1], m.sx, m.sy);
    if (!ln.ok) continue;
    const unsigned long long stamp = epoch_hi | (unsigned long long)(0xffffffffu - (unsigned int)i);
    const int start = ln.y0 * m.sx + ln.x0;
    for (unsigned int k = lane; k < ln.da; k += 32) {
      const unsigned int inc = (unsigned int)(((unsigned long long)ln.err0 + (unsigned long long)k * ln.db) / ln.da);
      const int off = start + (int)k * ln.off_a + (int)inc * ln.off_b;
      my_visits++;
      if (m.free_st[off] == stamp && (m.occ_st[off] >> 32) != (epoch_hi >> 32)) {  // bresenhamCellFree (:302-312)
        const float v = __fadd_rn(m.lo[off], lo_free);
        m.lo[off] = v;
        m.prob[off] = hs_prob_of(v);
        m.ui[off] = mark_free;
      }
    }
    if (lane == 0) {
      my_visits++;
      const int off = ln.y1 * m.sx + ln.x1;
      if (m.occ_st[off] == stamp) {  // bresenhamCellOcc (:315-330), first beam ending here
        float v = m.lo[off];
        const unsigned long long fs = m.free_st[off];
        if ((fs >> 32) == (epoch_hi >> 32) && fs > stamp) {
          v = __fadd_rn(v, lo_free);
          v = __fsub_rn(v, lo_free);
        }
        if (v < 50.0f) v = __fadd_rn(v, lo_occ);
        m.lo[off] = v;
        m.prob[off] = hs_prob_of(v);
        m.ui[off] = mark_occ;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) my_visits += __shfl_xor_sync(0xffffffffu, my_visits, d);
  if (lane == 0 && my_visits) atomicAdd(visits, my_visits);
}

__global__ void k_hs_fill(float *__restrict__ p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace b2s

extern "C" void b2s_hector_map_destroy(b2s_hector_map *m);
extern "C" {

b2s_status b2s_hector_map_create(int size_x, int size_y, float resolution, float start_x, float start_y, int device,
                                 void *cuda_stream, b2s_hector_map **out) {
  if (!out || size_x <= 2 || size_y <= 2 || !(resolution > 0.0f)) B2S_FAIL(B2S_ERR_BAD_PARAMS, "b2s_hector_map_create: bad size/resolution");
  *out = nullptr;
  if (b2s_device_count() <= device) B2S_FAIL(B2S_ERR_NO_DEVICE, "no usable CUDA device (the product path has no CPU fallback)");
  B2S_CUDA_CHECK(cudaSetDevice(device));
  b2s_hector_map *m = new (std::nothrow) b2s_hector_map();
  if (!m) B2S_FAIL(B2S_ERR_CUDA, "out of host memory");
  m->device = device;
  m->sx = size_x; m->sy = size_y;
  // MapRepMultiMap ctor (MapRepMultiMap.h:63-67) + GridMapBase::setMapTransformation (GridMapBase.h:270-286)
  const float total_x = resolution * (float)size_x, total_y = resolution * (float)size_y;
  m->off_x = total_x * start_x;
  m->off_y = total_y * start_y;
  m->cell_length = resolution;
  m->scale_to_map = 1.0f / resolution;
  m->tw_lin = m->scale_to_map;  // AlignedScaling2f(s,s) * Translation2f(off): linear diag(s,s), translation s*off
  m->tw_tx = m->scale_to_map * m->off_x;
  m->tw_ty = m->scale_to_map * m->off_y;
  const float det = m->tw_lin * m->tw_lin - 0.0f * 0.0f;  // Affine inverse, cofactor form
  const float invdet = 1.0f / det;
  m->wt_lin = m->tw_lin * invdet;
  const float i01 = -0.0f * invdet;
  m->wt_tx = -(m->wt_lin * m->tw_tx + i01 * m->tw_ty);
  m->wt_ty = -(i01 * m->tw_tx + m->wt_lin * m->tw_ty);
  m->lo_free = prob_to_log_odds(0.4f);  // GridMapLogOdds.h:98-102
  m->lo_occ = prob_to_log_odds(0.6f);
  if (cuda_stream) {
    m->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  } else {
    B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    m->own_stream = true;
  }
  for (auto &e : m->ev) B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaEventCreate(&e));
  const size_t cells = (size_t)size_x * size_y;
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaMalloc(reinterpret_cast<void **>(&m->d_lo), cells * 4));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaMalloc(reinterpret_cast<void **>(&m->d_ui), cells * 4));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaMalloc(reinterpret_cast<void **>(&m->d_free), cells * 8));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaMalloc(reinterpret_cast<void **>(&m->d_occ), cells * 8));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaMalloc(reinterpret_cast<void **>(&m->d_out), 12 * sizeof(float)));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaMalloc(reinterpret_cast<void **>(&m->d_visits), sizeof(unsigned long long)));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaMemsetAsync(m->d_lo, 0, cells * 4, m->stream));     // resetGridCell: logOdds 0
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaMemsetAsync(m->d_ui, 0xff, cells * 4, m->stream));  // updateIndex -1
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaMemsetAsync(m->d_free, 0, cells * 8, m->stream));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaMemsetAsync(m->d_occ, 0, cells * 8, m->stream));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_map_destroy(m), cudaStreamSynchronize(m->stream));
  m->epoch = 0;
  *out = m;
  return B2S_OK;
}

void b2s_hector_map_destroy(b2s_hector_map *m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (void *p : {(void *)m->d_lo, (void *)m->d_ui, (void *)m->d_free, (void *)m->d_occ, (void *)m->d_pts, (void *)m->d_out,
                  (void *)m->d_visits})
    if (p) cudaFree(p);
  for (auto &e : m->ev)
    if (e) cudaEventDestroy(e);
  if (m->own_stream && m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

b2s_status b2s_hector_map_set_factors(b2s_hector_map *m, float update_free, float update_occupied) {
  if (!m) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null handle");
  m->lo_free = prob_to_log_odds(update_free);
  m->lo_occ = prob_to_log_odds(update_occupied);
  return B2S_OK;
}

static b2s_status hc_upload_points(b2s_hector_map *m, const float *points, int n) {
  if ((size_t)n > m->pts_cap) {
    if (m->d_pts) B2S_CUDA_CHECK(cudaFree(m->d_pts));
    m->d_pts = nullptr;
    B2S_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&m->d_pts), sizeof(float) * 2 * (size_t)n));
    m->pts_cap = (size_t)n;
  }
  B2S_CUDA_CHECK(cudaMemcpyAsync(m->d_pts, points, sizeof(float) * 2 * (size_t)n, cudaMemcpyHostToDevice, m->stream));
  return B2S_OK;
}

static b2s_status hc_update(b2s_hector_map *m, const float *points, int n_points, const float origo[2],
                            const float world_pose[3], bool just_once) {
  if (!m || !origo || !world_pose || n_points < 0 || (n_points > 0 && !points)) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  B2S_CUDA_CHECK(cudaSetDevice(m->device));
  const int mark_free = m->curr_update_index + 1, mark_occ = m->curr_update_index + 2;  // OccGridMapBase.h:120-121
  // getMapCoordsPose (GridMapBase.h:238-242)
  float mx = (m->tw_lin * world_pose[0] + 0.0f * world_pose[1]) + m->tw_tx;
  float my = (0.0f * world_pose[0] + m->tw_lin * world_pose[1]) + m->tw_ty;
  float heading = world_pose[2];
  if (just_once) { mx = 800.0f; my = 800.0f; heading = 0.0f; }  // Eigen::Vector3f mapPose(800, 800, 0) (OccGridMapBase.h:182)
  HcTransform t;
  t.c = cosf(heading);  // Rotation2Df: std::cos / std::sin on float, taken on the host (glibc)
  t.s = sinf(heading);
  t.just_once = just_once ? 1 : 0;
  t.mx = mx; t.my = my;
  const float bxf = (t.c * origo[0] + (-t.s) * origo[1]) + mx, byf = (t.s * origo[0] + t.c * origo[1]) + my;
  const int bx = (int)(bxf + 0.5f), by = (int)(byf + 0.5f);  // Vector2i(float, float): truncation
  m->epoch++;
  const unsigned long long epoch_hi = (unsigned long long)m->epoch << 32;
  if (n_points > 0) {
    b2s_status st = hc_upload_points(m, points, n_points);
    if (st) return st;
    B2S_CUDA_CHECK(cudaMemsetAsync(m->d_visits, 0, sizeof(unsigned long long), m->stream));
    const int blocks = std::min(ceil_div(n_points, 8), 148 * 8);
    B2S_CUDA_CHECK(cudaEventRecord(m->ev[0], m->stream));
    k_hc_mark<<<blocks, 256, 0, m->stream>>>(m->d_pts, n_points, t, bx, by, m->sx, m->sy, epoch_hi, m->d_free, m->d_occ);
    k_hc_apply<<<blocks, 256, 0, m->stream>>>(m->d_pts, n_points, t, bx, by, m->sx, m->sy, epoch_hi, m->d_free, m->d_occ,
                                              m->lo_free, m->lo_occ, mark_free, mark_occ, m->d_lo, m->d_ui, m->d_visits);
    B2S_CUDA_CHECK(cudaEventRecord(m->ev[1], m->stream));
    B2S_CUDA_CHECK(cudaGetLastError());
  }
  m->curr_update_index += 3;  // OccGridMapBase.h:167
  return B2S_OK;
}

b2s_status b2s_hector_map_update_by_scan(b2s_hector_map *m, const float *points, int n_points, const float origo[2],
                                         const float world_pose[3]) {
  return hc_update(m, points, n_points, origo, world_pose, false);
}

b2s_status b2s_hector_map_update_by_scan_just_once(b2s_hector_map *m, const float *points, int n_points,
                                                   const float origo[2]) {
  const float unused_pose[3] = {0.0f, 0.0f, 0.0f};
  return hc_update(m, points, n_points, origo, unused_pose, true);
}

b2s_status b2s_hector_map_match_data(b2s_hector_map *m, const float *points, int n_points,
                                     const float begin_world_pose[3], int max_iterations, float out_world_pose[3],
                                     float out_cov[9]) {
  if (!m || !begin_world_pose || !out_world_pose || !out_cov || n_points < 0 || max_iterations < 0 || (n_points > 0 && !points))
    B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  if (n_points == 0) {  // ScanMatcher.h:66,97: no data -> begin estimate returned unchanged
    std::memcpy(out_world_pose, begin_world_pose, 3 * sizeof(float));
    return B2S_OK;
  }
  B2S_CUDA_CHECK(cudaSetDevice(m->device));
  b2s_status st = hc_upload_points(m, points, n_points);
  if (st) return st;
  const float e0 = (m->tw_lin * begin_world_pose[0] + 0.0f * begin_world_pose[1]) + m->tw_tx;
  const float e1 = (0.0f * begin_world_pose[0] + m->tw_lin * begin_world_pose[1]) + m->tw_ty;
  B2S_CUDA_CHECK(cudaEventRecord(m->ev[2], m->stream));
  k_hc_match<<<1, GN_THREADS, 0, m->stream>>>(m->d_lo, m->sx, m->sy, m->d_pts, n_points, e0, e1, begin_world_pose[2],
                                              1 + max_iterations, m->wt_lin, m->wt_tx, m->wt_ty, m->d_out);
  B2S_CUDA_CHECK(cudaEventRecord(m->ev[3], m->stream));
  B2S_CUDA_CHECK(cudaGetLastError());
  float host[12];
  B2S_CUDA_CHECK(cudaMemcpyAsync(host, m->d_out, sizeof(host), cudaMemcpyDeviceToHost, m->stream));
  B2S_CUDA_CHECK(cudaStreamSynchronize(m->stream));
  std::memcpy(out_world_pose, host, 3 * sizeof(float));
  std::memcpy(out_cov, host + 3, 9 * sizeof(float));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, m->ev[2], m->ev[3]) == cudaSuccess) m->last_ms[1] = ms;
  return B2S_OK;
}

b2s_status b2s_hector_map_copy(b2s_hector_map *m, float *log_odds, int32_t *update_index) {
  if (!m) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null handle");
  B2S_CUDA_CHECK(cudaSetDevice(m->device));
  const size_t bytes = (size_t)m->sx * m->sy * 4;
  if (log_odds) B2S_CUDA_CHECK(cudaMemcpyAsync(log_odds, m->d_lo, bytes, cudaMemcpyDeviceToHost, m->stream));
  if (update_index) B2S_CUDA_CHECK(cudaMemcpyAsync(update_index, m->d_ui, bytes, cudaMemcpyDeviceToHost, m->stream));
  B2S_CUDA_CHECK(cudaStreamSynchronize(m->stream));
  return B2S_OK;
}

b2s_status b2s_hector_map_copy_ros(b2s_hector_map *m, int8_t *out) {
  if (!m || !out) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  B2S_CUDA_CHECK(cudaSetDevice(m->device));
  const int n = m->sx * m->sy;
  int8_t *d = nullptr;
  B2S_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&d), (size_t)n, m->stream));
  k_hc_ros<<<ceil_div(n, 256), 256, 0, m->stream>>>(m->d_lo, n, d);
  B2S_CUDA_CHECK(cudaGetLastError());
  B2S_CUDA_CHECK(cudaMemcpyAsync(out, d, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
  B2S_CUDA_CHECK(cudaFreeAsync(d, m->stream));
  B2S_CUDA_CHECK(cudaStreamSynchronize(m->stream));
  return B2S_OK;
}

b2s_status b2s_hector_map_last_timing(b2s_hector_map *m, double out[2]) {
  if (!m || !out) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  B2S_CUDA_CHECK(cudaSetDevice(m->device));
  B2S_CUDA_CHECK(cudaStreamSynchronize(m->stream));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]) == cudaSuccess) m->last_ms[0] = ms;
  out[0] = m->last_ms[0];
  out[1] = m->last_ms[1];
  return B2S_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- HectorSlamProcessor stand-in

struct b2s_hector_slam {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int levels = 0;
  b2s_hector_map *map[B2S_HECTOR_MAX_LEVELS] = {};
  float *d_prob[B2S_HECTOR_MAX_LEVELS] = {};  // per-level probability planes (see HsLevel::prob)
  float *d_pts[B2S_HECTOR_MAX_LEVELS] = {};  // dataContainer (level 0) and dataContainers[l-1]
  int n_pts[B2S_HECTOR_MAX_LEVELS] = {};
  float origo[B2S_HECTOR_MAX_LEVELS][2] = {};
  size_t pts_cap = 0;
  float lo_free = 0, lo_occ = 0;
  float min_dist = 0.4f, min_angle = 0.13f;
  float last_map_update_pose[3], last_scan_match_pose[3], last_cov[9];
  float *d_out = nullptr;   // [3 pose, 9 H, 1 flag]
  float *h_out = nullptr;   // pinned
  float *h_pts = nullptr;   // pinned staging for the scan
  size_t h_pts_cap = 0;
  unsigned long long *d_visits = nullptr;
  double n_matched = 0, n_updated = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_h2d = nullptr;  // completion of the last upload out of h_pts
  bool h2d_pending = false;
  bool ev_match = false, ev_update = false;
};

static void hs_reset_poses(b2s_hector_slam *p) {  // HectorSlamProcessor::reset (:111-116)
  p->last_map_update_pose[0] = p->last_map_update_pose[1] = p->last_map_update_pose[2] = 3.402823466e+38F;
  p->last_scan_match_pose[0] = p->last_scan_match_pose[1] = p->last_scan_match_pose[2] = 0.0f;
}

// util::poseDifferenceLargerThan (UtilFunctions.h:72-90).  Only <cmath> is included there, so the unqualified
// abs(angleDiff) is int abs(int): the difference is truncated to an integer before the compare.
static bool hs_pose_difference_larger_than(const float a[3], const float b[3], float dist_thresh, float angle_thresh) {
  const float dx = a[0] - b[0], dy = a[1] - b[1];
  if (sqrtf(dx * dx + dy * dy) > dist_thresh) return true;
  float d = a[2] - b[2];
  const double pi = 3.14159265358979323846;
  if (d > pi) d = (float)((double)d - pi * 2.0f);
  else if (d < -pi) d = (float)((double)d + pi * 2.0f);
  return (float)std::abs((int)d) > angle_thresh;
}

static HsLevels hs_levels(const b2s_hector_slam *p) {
  HsLevels L;
  L.count = p->levels;
  for (int l = 0; l < p->levels; l++) {
    const b2s_hector_map *m = p->map[l];
    HsLevel &o = L.l[l];
    o.prob = p->d_prob[l];
    o.lo = m->d_lo; o.ui = m->d_ui; o.free_st = m->d_free; o.occ_st = m->d_occ;
    o.pts = p->d_pts[l]; o.n = p->n_pts[l];
    o.sx = m->sx; o.sy = m->sy;
    o.tw_lin = m->tw_lin; o.tw_tx = m->tw_tx; o.tw_ty = m->tw_ty;
    o.wt_lin = m->wt_lin; o.wt_tx = m->wt_tx; o.wt_ty = m->wt_ty;
    o.iterations = 1 + (l == 0 ? 5 : 3);
  }
  return L;
}

extern "C" {

b2s_status b2s_hector_slam_create(float map_resolution, int map_size_x, int map_size_y, float start_x, float start_y,
                                  int levels, int device, void *cuda_stream, b2s_hector_slam **out) {
  if (!out || levels < 1 || levels > B2S_HECTOR_MAX_LEVELS || !(map_resolution > 0.0f) || (map_size_x >> (levels - 1)) <= 2 ||
      (map_size_y >> (levels - 1)) <= 2)
    B2S_FAIL(B2S_ERR_BAD_PARAMS, "b2s_hector_slam_create: bad levels / size / resolution");
  *out = nullptr;
  if (b2s_device_count() <= device) B2S_FAIL(B2S_ERR_NO_DEVICE, "no usable CUDA device (the product path has no CPU fallback)");
  B2S_CUDA_CHECK(cudaSetDevice(device));
  b2s_hector_slam *p = new (std::nothrow) b2s_hector_slam();
  if (!p) B2S_FAIL(B2S_ERR_CUDA, "out of host memory");
  p->device = device;
  p->levels = levels;
  if (cuda_stream) {
    p->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  } else {
    B2S_CUDA_CHECK_CLEAN(b2s_hector_slam_destroy(p), cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    p->own_stream = true;
  }
  // MapRepMultiMap ctor (MapRepMultiMap.h:56-89): one offset for every level, dims halve, cell length doubles
  const float total_x = map_resolution * static_cast<float>(map_size_x), total_y = map_resolution * static_cast<float>(map_size_y);
  const float off_x = total_x * start_x, off_y = total_y * start_y;
  int sx = map_size_x, sy = map_size_y;
  float res = map_resolution;
  for (int l = 0; l < levels; l++) {
    b2s_status st = b2s_hector_map_create(sx, sy, res, 0.0f, 0.0f, device, p->stream, &p->map[l]);
    if (st) { b2s_hector_slam_destroy(p); return st; }
    b2s_hector_map *m = p->map[l];
    m->off_x = off_x; m->off_y = off_y;
    m->tw_tx = m->scale_to_map * off_x; m->tw_ty = m->scale_to_map * off_y;
    const float det = m->tw_lin * m->tw_lin - 0.0f * 0.0f, invdet = 1.0f / det, i01 = -0.0f * invdet;
    m->wt_tx = -(m->wt_lin * m->tw_tx + i01 * m->tw_ty);
    m->wt_ty = -(i01 * m->tw_tx + m->wt_lin * m->tw_ty);
    const int cells = sx * sy;
    B2S_CUDA_CHECK_CLEAN(b2s_hector_slam_destroy(p), cudaMalloc(reinterpret_cast<void **>(&p->d_prob[l]), sizeof(float) * (size_t)cells));
    k_hs_fill<<<ceil_div(cells, 256), 256, 0, p->stream>>>(p->d_prob[l], cells, 0.5f);  // logOdds 0 -> e^0 / (e^0 + 1)
    sx /= 2; sy /= 2;
    res *= 2.0f;
  }
  p->lo_free = prob_to_log_odds(0.4f);
  p->lo_occ = prob_to_log_odds(0.6f);
  hs_reset_poses(p);
  std::memset(p->last_cov, 0, sizeof(p->last_cov));
  for (auto &e : p->ev) B2S_CUDA_CHECK_CLEAN(b2s_hector_slam_destroy(p), cudaEventCreate(&e));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_slam_destroy(p), cudaEventCreateWithFlags(&p->ev_h2d, cudaEventDisableTiming));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_slam_destroy(p),
                       raise_dyn_smem(k_hs_match, HS_SMEM_BYTES));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_slam_destroy(p), cudaMalloc(reinterpret_cast<void **>(&p->d_out), 16 * sizeof(float)));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_slam_destroy(p), cudaMallocHost(reinterpret_cast<void **>(&p->h_out), 16 * sizeof(float)));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_slam_destroy(p), cudaMalloc(reinterpret_cast<void **>(&p->d_visits), sizeof(unsigned long long)));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_slam_destroy(p), cudaMemsetAsync(p->d_visits, 0, sizeof(unsigned long long), p->stream));
  B2S_CUDA_CHECK_CLEAN(b2s_hector_slam_destroy(p), cudaStreamSynchronize(p->stream));
  *out = p;
  return B2S_OK;
}

void b2s_hector_slam_destroy(b2s_hector_slam *p) {
  if (!p) return;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  for (int l = 0; l < B2S_HECTOR_MAX_LEVELS; l++) {
    if (p->map[l]) b2s_hector_map_destroy(p->map[l]);
    if (p->d_pts[l]) cudaFree(p->d_pts[l]);
    if (p->d_prob[l]) cudaFree(p->d_prob[l]);
  }
  if (p->d_out) cudaFree(p->d_out);
  if (p->d_visits) cudaFree(p->d_visits);
  if (p->h_out) cudaFreeHost(p->h_out);
  if (p->h_pts) cudaFreeHost(p->h_pts);
  for (auto &e : p->ev)
    if (e) cudaEventDestroy(e);
  if (p->ev_h2d) cudaEventDestroy(p->ev_h2d);
  if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

b2s_status b2s_hector_slam_set_update_factors(b2s_hector_slam *p, float update_free, float update_occupied) {
  if (!p) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null handle");
  p->lo_free = prob_to_log_odds(update_free);
  p->lo_occ = prob_to_log_odds(update_occupied);
  return B2S_OK;
}

b2s_status b2s_hector_slam_set_map_update_min_diff(b2s_hector_slam *p, float min_dist, float min_angle) {
  if (!p) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null handle");
  p->min_dist = min_dist;
  p->min_angle = min_angle;
  return B2S_OK;
}

b2s_status b2s_hector_slam_reset(b2s_hector_slam *p) {
  if (!p) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null handle");
  B2S_CUDA_CHECK(cudaSetDevice(p->device));
  hs_reset_poses(p);
  for (int l = 0; l < p->levels; l++) {  // GridMapBase::reset -> clear(): resetGridCell on every cell (:93-113)
    b2s_hector_map *m = p->map[l];
    const size_t cells = (size_t)m->sx * m->sy;
    B2S_CUDA_CHECK(cudaMemsetAsync(m->d_lo, 0, cells * 4, p->stream));
    B2S_CUDA_CHECK(cudaMemsetAsync(m->d_ui, 0xff, cells * 4, p->stream));
    k_hs_fill<<<ceil_div((long long)cells, 256), 256, 0, p->stream>>>(p->d_prob[l], (int)cells, 0.5f);
  }
  B2S_CUDA_CHECK(cudaGetLastError());
  return B2S_OK;
}

b2s_status b2s_hector_slam_update(b2s_hector_slam *p, const float *points, int n_points, const float origo[2],
                                  const float pose_hint_world[3], int map_without_matching, float out_pose[3],
                                  float out_cov[9], int *out_map_updated) {
  if (!p || !origo || !pose_hint_world || !out_pose || n_points < 0 || (n_points > 0 && !points))
    B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  B2S_CUDA_CHECK(cudaSetDevice(p->device));
  if ((size_t)n_points > p->pts_cap) {
    const size_t cap = std::max<size_t>(2048, (size_t)n_points);
    for (int l = 0; l < p->levels; l++) {
      float *fresh = nullptr;
      B2S_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&fresh), sizeof(float) * 2 * cap));
      // dataContainers[l-1] persist between calls (a map_without_matching update reuses them): carry them over
      if (p->d_pts[l] && p->n_pts[l] > 0)
        B2S_CUDA_CHECK(cudaMemcpyAsync(fresh, p->d_pts[l], sizeof(float) * 2 * (size_t)p->n_pts[l], cudaMemcpyDeviceToDevice, p->stream));
      B2S_CUDA_CHECK(cudaStreamSynchronize(p->stream));
      if (p->d_pts[l]) B2S_CUDA_CHECK(cudaFree(p->d_pts[l]));
      p->d_pts[l] = fresh;
    }
    p->pts_cap = cap;
  }
  if ((size_t)n_points > p->h_pts_cap) {
    if (p->h2d_pending) B2S_CUDA_CHECK(cudaEventSynchronize(p->ev_h2d));
    if (p->h_pts) B2S_CUDA_CHECK(cudaFreeHost(p->h_pts));
    p->h_pts = nullptr;
    const size_t cap = std::max<size_t>(2048, (size_t)n_points);
    B2S_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void **>(&p->h_pts), sizeof(float) * 2 * cap));
    p->h_pts_cap = cap;
  }
  // level 0 always sees the scan passed in; the coarse levels see the last MATCHED scan (MapRepMultiMap.h:144-191)
  if (n_points > 0) {
    // the pinned staging buffer may still be the source of the PREVIOUS call's upload (a map_without_matching call
    // returns without waiting for the stream): wait for that copy — not for the kernels behind it — before reusing it
    if (p->h2d_pending) B2S_CUDA_CHECK(cudaEventSynchronize(p->ev_h2d));
    std::memcpy(p->h_pts, points, sizeof(float) * 2 * (size_t)n_points);
    B2S_CUDA_CHECK(cudaMemcpyAsync(p->d_pts[0], p->h_pts, sizeof(float) * 2 * (size_t)n_points, cudaMemcpyHostToDevice, p->stream));
    B2S_CUDA_CHECK(cudaEventRecord(p->ev_h2d, p->stream));
    p->h2d_pending = true;
  }
  p->n_pts[0] = n_points;
  p->origo[0][0] = origo[0]; p->origo[0][1] = origo[1];
  float est[3] = {pose_hint_world[0], pose_hint_world[1], pose_hint_world[2]};
  if (!map_without_matching) {
    float factor = 1.0f;
    for (int l = 1; l < p->levels; l++) {  // setFrom(dataContainer, 1 / 2^l)
      factor *= 0.5f;
      p->n_pts[l] = n_points;
      p->origo[l][0] = origo[0] * factor; p->origo[l][1] = origo[1] * factor;
    }
    const HsLevels L = hs_levels(p);
    if (n_points > 0) {
      B2S_CUDA_CHECK(cudaEventRecord(p->ev[0], p->stream));
      k_hs_match<<<1, HS_THREADS, HS_SMEM_BYTES, p->stream>>>(L, est[0], est[1], est[2], p->d_out);
      B2S_CUDA_CHECK(cudaEventRecord(p->ev[1], p->stream));
      p->ev_match = true;
      B2S_CUDA_CHECK(cudaGetLastError());
      B2S_CUDA_CHECK(cudaMemcpyAsync(p->h_out, p->d_out, 13 * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
      B2S_CUDA_CHECK(cudaStreamSynchronize(p->stream));
      std::memcpy(est, p->h_out, sizeof(est));
      if (p->h_out[12] != 0.0f) std::memcpy(p->last_cov, p->h_out + 3, sizeof(p->last_cov));
    }
    p->n_matched += 1;
  }
  std::memcpy(p->last_scan_match_pose, est, sizeof(est));
  const bool do_update =
      hs_pose_difference_larger_than(est, p->last_map_update_pose, p->min_dist, p->min_angle) || map_without_matching;
  if (do_update) {
    const HsLevels L = hs_levels(p);
    HsUpdate U;
    int max_n = 0;
    for (int l = 0; l < p->levels; l++) {
      b2s_hector_map *m = p->map[l];
      const float mx = (m->tw_lin * est[0] + 0.0f * est[1]) + m->tw_tx;  // getMapCoordsPose (GridMapBase.h:238-242)
      const float my = (0.0f * est[0] + m->tw_lin * est[1]) + m->tw_ty;
      HcTransform &t = U.t[l];
      t.c = cosf(est[2]);  // Rotation2Df: std::cos / std::sin on float, on the host (glibc) -> bit-identical cells
      t.s = sinf(est[2]);
      t.mx = mx; t.my = my; t.just_once = 0;
      const float bxf = (t.c * p->origo[l][0] + (-t.s) * p->origo[l][1]) + mx;
      const float byf = (t.s * p->origo[l][0] + t.c * p->origo[l][1]) + my;
      U.bx[l] = (int)(bxf + 0.5f); U.by[l] = (int)(byf + 0.5f);
      m->epoch++;
      U.epoch_hi[l] = (unsigned long long)m->epoch << 32;
      U.mark_free[l] = m->curr_update_index + 1;  // OccGridMapBase.h:120-121
      U.mark_occ[l] = m->curr_update_index + 2;
      m->curr_update_index += 3;                  // :167
      max_n = std::max(max_n, p->n_pts[l]);
    }
    if (max_n > 0) {
      const dim3 grid(std::min(ceil_div(max_n, 8), 148 * 4), p->levels);
      B2S_CUDA_CHECK(cudaEventRecord(p->ev[2], p->stream));
      k_hs_mark<<<grid, 256, 0, p->stream>>>(L, U);
      k_hs_apply<<<grid, 256, 0, p->stream>>>(L, U, p->lo_free, p->lo_occ, p->d_visits);
      B2S_CUDA_CHECK(cudaEventRecord(p->ev[3], p->stream));
      p->ev_update = true;
      B2S_CUDA_CHECK(cudaGetLastError());
    }
    std::memcpy(p->last_map_update_pose, est, sizeof(est));
    p->n_updated += 1;
  }
  std::memcpy(out_pose, est, sizeof(est));
  if (out_cov && !map_without_matching) std::memcpy(out_cov, p->last_cov, sizeof(p->last_cov));
  if (out_map_updated) *out_map_updated = do_update ? 1 : 0;
  return B2S_OK;
}

b2s_status b2s_hector_slam_level_dims(b2s_hector_slam *p, int level, int dims[2], float *cell_length) {
  if (!p || !dims || level < 0 || level >= p->levels) B2S_FAIL(B2S_ERR_BAD_PARAMS, "bad level");
  dims[0] = p->map[level]->sx;
  dims[1] = p->map[level]->sy;
  if (cell_length) *cell_length = p->map[level]->cell_length;
  return B2S_OK;
}

b2s_status b2s_hector_slam_copy_level(b2s_hector_slam *p, int level, float *log_odds, int32_t *update_index) {
  if (!p || level < 0 || level >= p->levels) B2S_FAIL(B2S_ERR_BAD_PARAMS, "bad level");
  return b2s_hector_map_copy(p->map[level], log_odds, update_index);
}

b2s_status b2s_hector_slam_copy_level_ros(b2s_hector_slam *p, int level, int8_t *out) {
  if (!p || level < 0 || level >= p->levels) B2S_FAIL(B2S_ERR_BAD_PARAMS, "bad level");
  return b2s_hector_map_copy_ros(p->map[level], out);
}

b2s_status b2s_hector_slam_stats(b2s_hector_slam *p, double out[5]) {
  if (!p || !out) B2S_FAIL(B2S_ERR_BAD_PARAMS, "null argument");
  B2S_CUDA_CHECK(cudaSetDevice(p->device));
  unsigned long long v = 0;
  B2S_CUDA_CHECK(cudaMemcpyAsync(&v, p->d_visits, sizeof(v), cudaMemcpyDeviceToHost, p->stream));
  B2S_CUDA_CHECK(cudaStreamSynchronize(p->stream));
  out[0] = p->n_matched; out[1] = p->n_updated; out[2] = (double)v;
  float ms = 0;
  out[3] = (p->ev_match && cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]) == cudaSuccess) ? ms : 0.0;
  out[4] = (p->ev_update && cudaEventElapsedTime(&ms, p->ev[2], p->ev[3]) == cudaSuccess) ? ms : 0.0;
  return B2S_OK;
}

}  // extern "C"
