// math3.cuh - 3x3 float64 linear algebra shared by the RANSAC / refine kernels (host+device so that
// the CPU unit test tests/test_math3_host.py can exercise exactly the code the kernels run).
#pragma once
#include <math.h>

#ifndef RR_HD
#ifdef __CUDACC__
#define RR_HD __host__ __device__ __forceinline__
#else
#define RR_HD inline
#endif
#endif

namespace roreg {

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix, float64.  On return A = U*diag(S) column-wise
// (A[i][j] = U[i][j]*S[j]), V orthogonal with H*V = A.  Singular values are NOT sorted.
RR_HD void svd3_hestenes(const double H[9], double A[9], double V[9], double S[3]) {
  for (int i = 0; i < 9; ++i) A[i] = H[i];
  V[0] = 1; V[1] = 0; V[2] = 0; V[3] = 0; V[4] = 1; V[5] = 0; V[6] = 0; V[7] = 0; V[8] = 1;
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0;
      const int q = (pq == 0) ? 1 : 2;
      double alpha = 0, beta = 0, gamma = 0;
      for (int i = 0; i < 3; ++i) {
        alpha += A[3 * i + p] * A[3 * i + p];
        beta += A[3 * i + q] * A[3 * i + q];
        gamma += A[3 * i + p] * A[3 * i + q];
      }
      if (fabs(gamma) <= 1e-16 * sqrt(alpha * beta) || gamma == 0.0) continue;
      rotated = true;
      const double zeta = (beta - alpha) / (2.0 * gamma);
      const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
      for (int i = 0; i < 3; ++i) {
        const double ap = A[3 * i + p], aq = A[3 * i + q];
        A[3 * i + p] = c * ap - s * aq;
        A[3 * i + q] = s * ap + c * aq;
        const double vp = V[3 * i + p], vq = V[3 * i + q];
        V[3 * i + p] = c * vp - s * vq;
        V[3 * i + q] = s * vp + c * vq;
      }
    }
    if (!rotated) break;
  }
  for (int j = 0; j < 3; ++j)
    S[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
}

RR_HD double det3(const double M[9]) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
         M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// U (orthonormal columns) from A = U*diag(S).  The column of the SMALLEST singular value is always
// rebuilt as the cross product of the other two (exactly orthonormal U even when sigma_min is rounding
// noise, as in the rank-2 3-point Kabsch); its sign follows the computed column when sigma_min is
// meaningful (> 1e-10*sigma_max) and otherwise makes det(U) = det(V), i.e. U*V^T a proper rotation.
// Returns the index of the smallest singular value.
RR_HD int svd3_complete_u(const double A[9], const double V[9], const double S[3], double U[9]) {
  const double smax = fmax(S[0], fmax(S[1], S[2]));
  int jmin = 0;
  if (S[1] < S[jmin]) jmin = 1;
  if (S[2] < S[jmin]) jmin = 2;
  for (int j = 0; j < 3; ++j) {
    const double inv = (S[j] > 0) ? 1.0 / S[j] : 0.0;
    for (int i = 0; i < 3; ++i) U[3 * i + j] = A[3 * i + j] * inv;
  }
  // rank <= 1 (a triplet with a repeated match: 3-point draws are WITH replacement, test/estimator.py:228): two singular values
  // vanish and the cross-product completion below would return zero columns - i.e. a projector, not a rotation.  Complete U with
  // an orthonormal basis instead, as LAPACK's SVD does for the reference (which basis is arbitrary there too).
  {
    int jmax = 0;
    if (S[1] > S[jmax]) jmax = 1;
    if (S[2] > S[jmax]) jmax = 2;
    int nsig = 0;
    for (int j = 0; j < 3; ++j) nsig += (S[j] > 1e-10 * smax) ? 1 : 0;
    if (smax <= 0.0 || nsig <= 1) {
      if (smax <= 0.0) { for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? 1.0 : 0.0; return jmin; }
      const double u0 = U[0 + jmax], u1 = U[3 + jmax], u2 = U[6 + jmax];
      int e = 0;
      if (fabs(u1) < fabs(u0)) e = 1;
      if (fabs(u2) < fabs(e == 0 ? u0 : u1)) e = 2;
      const double ex = (e == 0), ey = (e == 1), ez = (e == 2);
      double v0 = u1 * ez - u2 * ey, v1 = u2 * ex - u0 * ez, v2 = u0 * ey - u1 * ex;        // u x e
      const double nv = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
      v0 /= nv; v1 /= nv; v2 /= nv;
      const double w0 = u1 * v2 - u2 * v1, w1 = u2 * v0 - u0 * v2, w2 = u0 * v1 - u1 * v0;  // u x v: (u, v, w) right-handed
      const int ja = (jmax + 1) % 3, jb = (jmax + 2) % 3;
      U[0 + ja] = v0; U[3 + ja] = v1; U[6 + ja] = v2;
      U[0 + jb] = w0; U[3 + jb] = w1; U[6 + jb] = w2;
      if (jmin == jmax) jmin = ja;
      return jmin;
    }
  }
  const int a = (jmin + 1) % 3, b = (jmin + 2) % 3;
  double c0 = U[3 + a] * U[6 + b] - U[6 + a] * U[3 + b];
  double c1 = U[6 + a] * U[0 + b] - U[0 + a] * U[6 + b];
  double c2 = U[0 + a] * U[3 + b] - U[3 + a] * U[0 + b];
  const double n = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
  if (n > 0) { c0 /= n; c1 /= n; c2 /= n; }
  double sgn;
  if (S[jmin] > 1e-10 * smax) sgn = (A[0 + jmin] * c0 + A[3 + jmin] * c1 + A[6 + jmin] * c2) >= 0 ? 1.0 : -1.0;
  else sgn = (det3(V) >= 0) ? 1.0 : -1.0;      // det([c, u_a, u_b]) = +1 for the cyclic (jmin, a, b)
  U[0 + jmin] = sgn * c0; U[3 + jmin] = sgn * c1; U[6 + jmin] = sgn * c2;
  return jmin;
}

// Orthogonal polar factor R = U V^T of H = U S V^T, WITHOUT reflection fix
// (refiner.SVDR_w, test/estimator.py:39-43: "return np.matmul(U,VT)").
RR_HD void polar_uvt(const double H[9], double R[9]) {
  double A[9], V[9], S[3], U[9];
  svd3_hestenes(H, A, V, S);
  svd3_complete_u(A, V, S, U);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R[3 * i + j] = U[3 * i] * V[3 * j] + U[3 * i + 1] * V[3 * j + 1] + U[3 * i + 2] * V[3 * j + 2];
}

// rotation = V U^T of m = U S V^T (yohoc_ransac.Threepps2Tran, test/estimator.py:139-147), with the
// sign of the smallest singular pair chosen so that det(rotation) = +1.  For the rank-2 3-point case
// the reference's sign is LAPACK rounding noise (50 % reflections); this is the proper
// branch of that coin flip.
RR_HD void kabsch_vut_proper(const double M[9], double R[9]) {
  double A[9], V[9], S[3], U[9];
  svd3_hestenes(M, A, V, S);
  const int jmin = svd3_complete_u(A, V, S, U);
  const double d = (det3(U) * det3(V) < 0) ? -1.0 : 1.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0;
      for (int k = 0; k < 3; ++k) acc += (k == jmin ? d : 1.0) * V[3 * i + k] * U[3 * j + k];
      R[3 * i + j] = acc;
    }
}

// Threepps2Tran: kps0/kps1 are 3 points (row-major [3][3]); T is [3][4] with R k1 + t = k0.
RR_HD void three_point_transform(const double k0[9], const double k1[9], double T[12]) {
  double c0[3], c1[3];
  for (int j = 0; j < 3; ++j) {
    c0[j] = (k0[j] + k0[3 + j] + k0[6 + j]) / 3.0;
    c1[j] = (k1[j] + k1[3 + j] + k1[6 + j]) / 3.0;
  }
  double M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0;
      for (int p = 0; p < 3; ++p) acc += (k1[3 * p + i] - c1[i]) * (k0[3 * p + j] - c0[j]);
      M[3 * i + j] = acc;
    }
  double R[9];
  kabsch_vut_proper(M, R);
  for (int i = 0; i < 3; ++i) {
    T[4 * i + 0] = R[3 * i + 0]; T[4 * i + 1] = R[3 * i + 1]; T[4 * i + 2] = R[3 * i + 2];
    T[4 * i + 3] = c0[i] - (R[3 * i] * c1[0] + R[3 * i + 1] * c1[1] + R[3 * i + 2] * c1[2]);
  }
}

// utils/r_eval.py:90-106 matrix_from_quaternion with float32 products (the reference feeds it a
// float32 row), then @ float32 Rgroup in float64 (test/estimator.py:354-356).
RR_HD void quat_times_anchor(const float q[4], const float Rg[9], double R[9]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  float m[9];
#ifdef __CUDA_ARCH__
  // no FMA contraction: NumPy evaluates each product and sum as a separately rounded float32 op
  m[0] = __fsub_rn(__fsub_rn(1.f, __fmul_rn(__fmul_rn(2.f, y), y)), __fmul_rn(__fmul_rn(2.f, z), z));
  m[1] = __fsub_rn(__fmul_rn(__fmul_rn(2.f, x), y), __fmul_rn(__fmul_rn(2.f, z), w));
  m[2] = __fadd_rn(__fmul_rn(__fmul_rn(2.f, x), z), __fmul_rn(__fmul_rn(2.f, y), w));
  m[3] = __fadd_rn(__fmul_rn(__fmul_rn(2.f, x), y), __fmul_rn(__fmul_rn(2.f, z), w));
  m[4] = __fsub_rn(__fsub_rn(1.f, __fmul_rn(__fmul_rn(2.f, x), x)), __fmul_rn(__fmul_rn(2.f, z), z));
  m[5] = __fsub_rn(__fmul_rn(__fmul_rn(2.f, y), z), __fmul_rn(__fmul_rn(2.f, x), w));
  m[6] = __fsub_rn(__fmul_rn(__fmul_rn(2.f, x), z), __fmul_rn(__fmul_rn(2.f, y), w));
  m[7] = __fadd_rn(__fmul_rn(__fmul_rn(2.f, y), z), __fmul_rn(__fmul_rn(2.f, x), w));
  m[8] = __fsub_rn(__fsub_rn(1.f, __fmul_rn(__fmul_rn(2.f, x), x)), __fmul_rn(__fmul_rn(2.f, y), y));
#else
  volatile float t0, t1;
  t0 = 2.f * y; t0 = t0 * y; t1 = 2.f * z; t1 = t1 * z; t0 = 1.f - t0; m[0] = t0 - t1;
  t0 = 2.f * x; t0 = t0 * y; t1 = 2.f * z; t1 = t1 * w; m[1] = t0 - t1;
  t0 = 2.f * x; t0 = t0 * z; t1 = 2.f * y; t1 = t1 * w; m[2] = t0 + t1;
  t0 = 2.f * x; t0 = t0 * y; t1 = 2.f * z; t1 = t1 * w; m[3] = t0 + t1;
  t0 = 2.f * x; t0 = t0 * x; t1 = 2.f * z; t1 = t1 * z; t0 = 1.f - t0; m[4] = t0 - t1;
  t0 = 2.f * y; t0 = t0 * z; t1 = 2.f * x; t1 = t1 * w; m[5] = t0 - t1;
  t0 = 2.f * x; t0 = t0 * z; t1 = 2.f * y; t1 = t1 * w; m[6] = t0 - t1;
  t0 = 2.f * y; t0 = t0 * z; t1 = 2.f * x; t1 = t1 * w; m[7] = t0 + t1;
  t0 = 2.f * x; t0 = t0 * x; t1 = 2.f * y; t1 = t1 * y; t0 = 1.f - t0; m[8] = t0 - t1;
#endif
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
#ifdef __CUDA_ARCH__
      double acc = __dmul_rn((double)m[3 * i], (double)Rg[j]);
      acc = __dadd_rn(acc, __dmul_rn((double)m[3 * i + 1], (double)Rg[3 + j]));
      acc = __dadd_rn(acc, __dmul_rn((double)m[3 * i + 2], (double)Rg[6 + j]));
      R[3 * i + j] = acc;
#else
      R[3 * i + j] = (double)m[3 * i] * (double)Rg[j] + (double)m[3 * i + 1] * (double)Rg[3 + j] +
                     (double)m[3 * i + 2] * (double)Rg[6 + j];
#endif
    }
}

// ---------------------------------------------------------------------------------------------
// One-shot scoring, float32 pre-filter (score mode 1, kernels_ransac.cuh ransac_score_pre_kernel).  The per-test arithmetic
// lives here so that the host build (tests/test_score_prefilter_bound.py) runs the very expressions the kernel runs.
// ---------------------------------------------------------------------------------------------
// float32 at or below / at or above a float64 value
RR_HD float f32_at_or_below(double x) {
#ifdef __CUDA_ARCH__
  return __double2float_rd(x);
#else
  const float f = (float)x;
  return ((double)f > x) ? nextafterf(f, -INFINITY) : f;
#endif
}
RR_HD float f32_at_or_above(double x) {
#ifdef __CUDA_ARCH__
  return __double2float_ru(x);
#else
  const float f = (float)x;
  return ((double)f < x) ? nextafterf(f, INFINITY) : f;
#endif
}

// Half-width m of the undecided band around r2 for hypothesis T = [R|t] (row-major 3x4, float64) against points whose
// coordinates are bounded by amax (k0 side) and bmax (k1 side); u = 2^-24:
//   E = u (amax + 6 (max_c sum_j |R_cj| * bmax + max_c |t_c|)),  m = 2 (2 sqrt(3) E r + 4 E^2 + 6 u r2)     (derivation: kernels_ransac.cuh)
RR_HD double prefilter_band(const double T[12], double amax, double bmax, double r, double r2) {
  double rrow = 0.0, tmax = 0.0;
  for (int c = 0; c < 3; ++c) {
    rrow = fmax(rrow, fabs(T[4 * c]) + fabs(T[4 * c + 1]) + fabs(T[4 * c + 2]));
    tmax = fmax(tmax, fabs(T[4 * c + 3]));
  }
  const double u = 5.9604644775390625e-8;
  const double E = u * (amax + 6.0 * (rrow * bmax + tmax));
  return 2.0 * (3.4641016151377549 * E * r + 4.0 * E * E + 6.0 * u * r2);
}

// float32 squared distance |a - (R b + t)|^2 with the fixed fma order the bound is derived for
RR_HD float prefilter_dist2(const float T[12], float ax, float ay, float az, float bx, float by, float bz) {
  const float x = fmaf(T[0], bx, fmaf(T[1], by, fmaf(T[2], bz, T[3])));
  const float y = fmaf(T[4], bx, fmaf(T[5], by, fmaf(T[6], bz, T[7])));
  const float z = fmaf(T[8], bx, fmaf(T[9], by, fmaf(T[10], bz, T[11])));
  const float dx = ax - x, dy = ay - y, dz = az - z;
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

}  // namespace roreg
