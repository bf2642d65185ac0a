/* oracle_c.c - plain C (+pthreads) restatement of the hot loops of RoReg's per-pair path.
 *
 * TEST INFRASTRUCTURE ONLY: this is the multi-threaded CPU baseline bench.py times (cpu_baseline /
 * --impl reference) and a second checker for tests; the product never links or calls it.
 * Each function cites the reference lines it restates; the NumPy oracle (roreg_oracle.py), which is
 * pinned against the reference's own outputs, is the arbiter - tests/test_oracle_c.py compares the two.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

/* minimal parallel-for over pthreads (this image's gcc has no libgomp): dynamic chunks of `grain` */
typedef void (*orc_body)(int lo, int hi, void* arg);
typedef struct { orc_body fn; void* arg; int n, grain; volatile int next; } orc_job;
static void* orc_worker(void* p) {
  orc_job* j = (orc_job*)p;
  for (;;) {
    const int lo = __sync_fetch_and_add(&j->next, j->grain);
    if (lo >= j->n) break;
    j->fn(lo, lo + j->grain < j->n ? lo + j->grain : j->n, j->arg);
  }
  return 0;
}
int orc_threads(void) {
  const char* e = getenv("ORC_THREADS");
  long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}
static void orc_parallel_for(int n, int grain, orc_body fn, void* arg) {
  orc_job job = {fn, arg, n, grain < 1 ? 1 : grain, 0};
  const int nt = orc_threads();
  pthread_t th[256];
  for (int t = 1; t < nt; ++t) pthread_create(&th[t], 0, orc_worker, &job);
  orc_worker(&job);
  for (int t = 1; t < nt; ++t) pthread_join(th[t], 0);
}

/* test/matcher.py:69-72: mean over the 60 group elements, x/(||x||+1e-5).  feats [n,32,60] -> out [n,32] */
typedef struct { const float* feats; int normalise; float* out; } pool_arg;
static void pool_body(int lo, int hi, void* p) {
  pool_arg* A = (pool_arg*)p; const float* feats = A->feats; const int normalise = A->normalise; float* out = A->out;
  for (int i = lo; i < hi; ++i) {
    float m[32]; float ss = 0.f;
    for (int f = 0; f < 32; ++f) {
      const float* p = feats + ((size_t)i * 32 + f) * 60;
      float s = 0.f;
      for (int g = 0; g < 60; ++g) s += p[g];
      m[f] = s / 60.0f; ss += m[f] * m[f];
    }
    const float d = normalise ? (sqrtf(ss) + 1e-5f) : 1.0f;
    for (int f = 0; f < 32; ++f) out[(size_t)i * 32 + f] = m[f] / d;
  }
}
void orc_inv_pool(const float* feats, int n, int normalise, float* out) {
  pool_arg A = {feats, normalise, out};
  orc_parallel_for(n, 64, pool_body, &A);
}

/* utils/knn_search.py:17-21,26-66: for each source row its nearest target row, d = sqrt(sum (a-b)^2 + 1e-7),
 * first minimal index (torch.min).  target [n,f], source [m,f]. */
typedef struct { const float* tgt; int n; const float* src; int f; int32_t* idx; float* dist; } nn_arg;
static void nn_body(int lo, int hi, void* p) {
  nn_arg* A = (nn_arg*)p; const float* tgt = A->tgt; const float* src = A->src; const int n = A->n, f = A->f;
  int32_t* idx = A->idx; float* dist = A->dist;
  for (int i = lo; i < hi; ++i) {
    const float* a = src + (size_t)i * f;
    float best = INFINITY; int bi = -1;
    for (int j = 0; j < n; ++j) {
      const float* b = tgt + (size_t)j * f;
      float d2 = 0.f;
      for (int c = 0; c < f; ++c) { const float d = a[c] - b[c]; d2 += d * d; }
      const float s = sqrtf(d2 + 1e-7f);
      if (s < best) { best = s; bi = j; }
    }
    idx[i] = bi; if (dist) dist[i] = best;
  }
}
void orc_nn(const float* tgt, int n, const float* src, int m, int f, int32_t* idx, float* dist) {
  nn_arg A = {tgt, n, src, f, idx, dist};
  orc_parallel_for(m, 16, nn_body, &A);
}

/* test/estimator.py:85-89 Batch_Des2R_torch: cor[a] = sum_{f,g} X[f,P[a,g]] Y[f,g]; argmax (first).
 * X,Y descriptor arrays [*,32,60]; ix/iy row indices [K]; perm [60][60] int32. */
typedef struct { const float* X; const float* Y; const int32_t* ix; const int32_t* iy; const int32_t* perm; int32_t* out; float* cor_out; } d2r_arg;
static void d2r_body(int lo, int hi, void* p) {
  d2r_arg* A = (d2r_arg*)p; const float* X = A->X; const float* Y = A->Y; const int32_t* ix = A->ix; const int32_t* iy = A->iy;
  const int32_t* perm = A->perm; int32_t* out = A->out; float* cor_out = A->cor_out;
  for (int k = lo; k < hi; ++k) {
    const float* x = X + (size_t)ix[k] * 1920; const float* y = Y + (size_t)iy[k] * 1920;
    float G[60][60];
    for (int h = 0; h < 60; ++h) for (int g = 0; g < 60; ++g) G[h][g] = 0.f;
    for (int f = 0; f < 32; ++f)
      for (int h = 0; h < 60; ++h) {
        const float xv = x[f * 60 + h];
        for (int g = 0; g < 60; ++g) G[h][g] += xv * y[f * 60 + g];
      }
    float bv = -INFINITY; int ba = 0;
    for (int a = 0; a < 60; ++a) {
      float c = 0.f;
      for (int g = 0; g < 60; ++g) c += G[perm[a * 60 + g]][g];
      if (cor_out) cor_out[(size_t)k * 60 + a] = c;
      if (c > bv) { bv = c; ba = a; }
    }
    out[k] = ba;
  }
}
void orc_des2r(const float* X, const float* Y, const int32_t* ix, const int32_t* iy, int K, const int32_t* perm,
               int32_t* out, float* cor_out) {
  d2r_arg A = {X, Y, ix, iy, perm, out, cor_out};
  orc_parallel_for(K, 8, d2r_body, &A);
}

/* test/estimator.py:149-154 / :377-382 overlap_cal for every hypothesis: ov[h] = sum s_i [|k0 - T k1|^2 < ird^2] / K */
typedef struct { const double* k0; const double* k1; const double* scores; int K; const double* hyps; double r2; double* ov; } sc_arg;
static void sc_body(int lo, int hi, void* p) {
  sc_arg* A = (sc_arg*)p; const double* k0 = A->k0; const double* k1 = A->k1; const double* scores = A->scores;
  const int K = A->K; const double* hyps = A->hyps; const double r2 = A->r2; double* ov = A->ov;
  for (int h = lo; h < hi; ++h) {
    const double* T = hyps + (size_t)h * 12;
    double acc = 0.0;
    for (int i = 0; i < K; ++i) {
      const double* b = k1 + 3 * (size_t)i; const double* a = k0 + 3 * (size_t)i;
      const double x = T[0] * b[0] + T[1] * b[1] + T[2] * b[2] + T[3];
      const double y = T[4] * b[0] + T[5] * b[1] + T[6] * b[2] + T[7];
      const double z = T[8] * b[0] + T[9] * b[1] + T[10] * b[2] + T[11];
      const double dx = a[0] - x, dy = a[1] - y, dz = a[2] - z;
      if (dx * dx + dy * dy + dz * dz < r2) acc += scores[i];
    }
    ov[h] = acc / (double)K;
  }
}
void orc_score(const double* k0, const double* k1, const double* scores, int K, const double* hyps, int H, double ird,
               double* ov) {
  sc_arg A = {k0, k1, scores, K, hyps, ird * ird, ov};
  orc_parallel_for(H, 8, sc_body, &A);
}
