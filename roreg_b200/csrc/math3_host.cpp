// Host build of math3.cuh for the CPU unit test (tests/test_math3_host.py): the same source the
// RANSAC / refine kernels compile, exercised without a GPU.
#include "math3.cuh"
extern "C" {
void rr_host_polar_uvt(const double* H, double* R) { roreg::polar_uvt(H, R); }
void rr_host_three_point_transform(const double* k0, const double* k1, double* T) { roreg::three_point_transform(k0, k1, T); }
void rr_host_quat_times_anchor(const float* q, const float* Rg, double* R) { roreg::quat_times_anchor(q, Rg, R); }
}
