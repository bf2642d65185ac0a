// Host build of math3.cuh for the CPU unit test (tests/test_math3_host.py): the same source the
// RANSAC / refine kernels compile, exercised without a GPU.
#include "math3.cuh"
extern "C" {
void rr_host_polar_uvt(const double* H, double* R) { roreg::polar_uvt(H, R); }
void rr_host_three_point_transform(const double* k0, const double* k1, double* T) { roreg::three_point_transform(k0, k1, T); }
void rr_host_quat_times_anchor(const float* q, const float* Rg, double* R) { roreg::quat_times_anchor(q, Rg, R); }
// score mode 1 on the host: decisions of the float32 pre-filter for H hypotheses x K points, staged in tiles of 256 points as the
// kernel does (per-tile coordinate maxima).  decision: 0 = outlier, 1 = inlier, 2 = undecided (the kernel repeats it in float64);
// band[h * tiles + tile] = m.
void rr_host_prefilter(const double* T, int H, const double* k0, const double* k1, int K, double r, signed char* decision, double* band) {
  const double r2 = r * r;
  const int tiles = (K + 255) / 256;
  for (int tile = 0; tile < tiles; ++tile) {
    const int k_begin = tile * 256, cnt = (K - k_begin < 256) ? K - k_begin : 256;
    float am = 0.f, bm = 0.f;
    for (int k = k_begin; k < k_begin + cnt; ++k) {
      const float a1 = roreg::f32_at_or_above(fmax(fabs(k0[3 * k]), fmax(fabs(k0[3 * k + 1]), fabs(k0[3 * k + 2]))));
      const float b1 = roreg::f32_at_or_above(fmax(fabs(k1[3 * k]), fmax(fabs(k1[3 * k + 1]), fabs(k1[3 * k + 2]))));
      am = a1 > am ? a1 : am; bm = b1 > bm ? b1 : bm;
    }
    for (int h = 0; h < H; ++h) {
      float Tf[12];
      for (int j = 0; j < 12; ++j) Tf[j] = (float)T[12 * h + j];
      const double m = roreg::prefilter_band(T + 12 * h, (double)am, (double)bm, r, r2);
      band[h * tiles + tile] = m;
      const float lo = roreg::f32_at_or_below(r2 - m), hi = roreg::f32_at_or_above(r2 + m);
      for (int k = k_begin; k < k_begin + cnt; ++k) {
        const float s2 = roreg::prefilter_dist2(Tf, (float)k0[3 * k], (float)k0[3 * k + 1], (float)k0[3 * k + 2], (float)k1[3 * k],
                                                (float)k1[3 * k + 1], (float)k1[3 * k + 2]);
        decision[(long long)h * K + k] = (s2 < lo) ? 1 : ((s2 > hi) ? 0 : 2);
      }
    }
  }
}
}
