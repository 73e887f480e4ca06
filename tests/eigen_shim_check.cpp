// TEST HELPER (host only): the Eigen stand-in that the reference's lesson4 headers are compiled against
// (oracle/shim/Eigen/mini_eigen.h) must evaluate its primitives exactly as its header says Eigen 3.3 does.  Every check
// spells the expected IEEE-754 float32 operation order out by hand (volatile temporaries: no contraction, no reordering).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <Eigen/Geometry>
#include <Eigen/LU>

static int failures = 0;
static bool same(float a, float b) { return std::memcmp(&a, &b, 4) == 0; }
#define CHECK(name, a, b) do { if (!same((a), (b))) { std::printf("FAIL %s: %.9g vs %.9g\n", name, (double)(a), (double)(b)); failures++; } } while (0)
#define CHECKI(name, a, b) do { if ((a) != (b)) { std::printf("FAIL %s: %d vs %d\n", name, (int)(a), (int)(b)); failures++; } } while (0)

static float mulf(float a, float b) { volatile float r = a * b; return r; }
static float addf(float a, float b) { volatile float r = a + b; return r; }
static float subf(float a, float b) { volatile float r = a - b; return r; }

int main() {
  // awkward values: products and sums round differently under every re-association
  const float tx = 511.99997f, ty = -487.33334f, ang = 2.7182817f;
  const float px = 123.45678f, py = -98.765434f;
  const float c = std::cos(ang), s = std::sin(ang);  // std::cos(float) = cosf
  {  // Translation2f * Rotation2Df applied to a point: (c*x + (-s)*y) + t, two-term row sums left to right
    Eigen::Affine2f T(Eigen::Translation2f(tx, ty) * Eigen::Rotation2Df(ang));
    const Eigen::Vector2f q = T * Eigen::Vector2f(px, py);
    CHECK("affine2 x", q[0], addf(addf(mulf(c, px), mulf(-s, py)), tx));
    CHECK("affine2 y", q[1], addf(addf(mulf(s, px), mulf(c, py)), ty));
  }
  {  // AlignedScaling2f * Translation2f: linear diag(s), translation s * off; its Affine inverse: L' = L^-1, t' = -(L' t)
    const float sc = 1.0f / 0.05f, ox = 25.600002f, oy = 12.799999f;
    Eigen::Affine2f M(Eigen::AlignedScaling2f(sc, sc) * Eigen::Translation2f(ox, oy));
    const Eigen::Vector2f w(1.2345678f, -7.6543207f);
    const Eigen::Vector2f g = M * w;
    CHECK("scaling*translation x", g[0], addf(addf(mulf(sc, w[0]), mulf(0.0f, w[1])), mulf(sc, ox)));
    CHECK("scaling*translation y", g[1], addf(addf(mulf(0.0f, w[0]), mulf(sc, w[1])), mulf(sc, oy)));
    const Eigen::Affine2f Mi = M.inverse();
    const float det = subf(mulf(sc, sc), mulf(0.0f, 0.0f)), invdet = 1.0f / det;
    const float l = mulf(sc, invdet), o01 = mulf(-0.0f, invdet);
    const float itx = -addf(mulf(l, mulf(sc, ox)), mulf(o01, mulf(sc, oy))), ity = -addf(mulf(o01, mulf(sc, ox)), mulf(l, mulf(sc, oy)));
    const Eigen::Vector2f b = Mi * g;
    CHECK("affine inverse x", b[0], addf(addf(mulf(l, g[0]), mulf(o01, g[1])), itx));
    CHECK("affine inverse y", b[1], addf(addf(mulf(o01, g[0]), mulf(l, g[1])), ity));
  }
  {  // float -> int: truncation toward zero, also for negatives; Vector2i(float, float) and cast<int>()
    const Eigen::Vector2f v(-3.9999f, 7.9999f);
    const Eigen::Vector2i a = v.cast<int>();
    CHECKI("cast<int> negative", a[0], -3);
    CHECKI("cast<int> positive", a[1], 7);
    const Eigen::Vector2i b2(-0.99f, 511.5f);
    CHECKI("Vector2i(float) negative", b2[0], 0);
    CHECKI("Vector2i(float) positive", b2[1], 511);
  }
  {  // Matrix3f * Vector3f: three-term rows as p0 + (p1 + p2); Matrix3f::inverse(): cofactors times 1/det, det = c00*m00 + (c10*m01 + c20*m02)
    Eigen::Matrix3f H;
    const float hv[9] = {812.25f, -33.125f, 17.0625f, -33.125f, 640.5f, -91.03125f, 17.0625f, -91.03125f, 2210.75f};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) H(i, j) = hv[3 * i + j];
    const Eigen::Vector3f v(0.33333334f, -1.7320508f, 2.2360680f);
    const Eigen::Vector3f r = H * v;
    CHECK("mat3*vec row0", r[0], addf(mulf(H(0, 0), v[0]), addf(mulf(H(0, 1), v[1]), mulf(H(0, 2), v[2]))));
    CHECK("mat3*vec row2", r[2], addf(mulf(H(2, 0), v[0]), addf(mulf(H(2, 1), v[1]), mulf(H(2, 2), v[2]))));
    const Eigen::Matrix3f Hi = H.inverse();
    const float c00 = subf(mulf(H(1, 1), H(2, 2)), mulf(H(1, 2), H(2, 1)));
    const float c10 = subf(mulf(H(1, 2), H(2, 0)), mulf(H(1, 0), H(2, 2)));
    const float c20 = subf(mulf(H(1, 0), H(2, 1)), mulf(H(1, 1), H(2, 0)));
    const float det = addf(mulf(c00, H(0, 0)), addf(mulf(c10, H(0, 1)), mulf(c20, H(0, 2))));
    const float invdet = 1.0f / det;
    CHECK("inverse(0,0)", Hi(0, 0), mulf(c00, invdet));
    CHECK("inverse(1,0)", Hi(1, 0), mulf(c10, invdet));
    CHECK("inverse(2,0)", Hi(2, 0), mulf(c20, invdet));
    CHECK("inverse(0,1)", Hi(0, 1), mulf(subf(mulf(H(0, 2), H(2, 1)), mulf(H(0, 1), H(2, 2))), invdet));
    CHECK("inverse(2,2)", Hi(2, 2), mulf(subf(mulf(H(0, 0), H(1, 1)), mulf(H(0, 1), H(1, 0))), invdet));
    const Eigen::Vector3f d = Hi * v;  // the Gauss-Newton step H^-1 * dTr
    CHECK("inverse*vec row1", d[1], addf(mulf(Hi(1, 0), v[0]), addf(mulf(Hi(1, 1), v[1]), mulf(Hi(1, 2), v[2]))));
  }
  std::printf("{\"failures\": %d}\n", failures);
  return failures ? 1 : 0;
}
