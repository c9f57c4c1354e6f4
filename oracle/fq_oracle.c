/*
 * fq_oracle.c -- plain C (+OpenMP) restatement of the reference's CPU arithmetic.
 * TEST INFRASTRUCTURE ONLY: the checker and the all-cores CPU baseline of bench.py.  Never linked
 * into or called by the product (quantization/mxnet_b200).
 *
 * Each function follows the reference lines cited beside it (paths relative to the reference root)
 * and mirrors oracle/fq_oracle.py, against which tests/test_oracle_c.py checks it bit for bit.
 * Parity status: histogram / KL are pinned through the NumPy oracle to the reference's own outputs
 * (tests/golden); the MXNet-op restatements are PARITY UNPINNED (see oracle/fq_oracle.py).
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC  (no -ffast-math: IEEE semantics matter).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* mshadow_op::round == roundf; mshadow_op::clip: compare/select */
static inline float clipf(float x, float lo, float hi) { return x > hi ? hi : (x < lo ? lo : x); }

/* max |x| per row: convert_conv2d.py:56,75,86,92 */
API void fqo_absmax_rows(const float* x, int64_t rows, int64_t len, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    float m = 0.f;
    const float* p = x + r * len;
    for (int64_t i = 0; i < len; ++i) {
      const float a = fabsf(p[i]);
      if (a > m) m = a;
    }
    out[r] = m;
  }
}

/* MXNet CPU mean: Kahan fp32 sum in index order, one fp32 divide (convert_conv2d.py:56 `.mean()`) */
API float fqo_mean_kahan(const float* v, int64_t n) {
  volatile float s = 0.f, c = 0.f;
  for (int64_t i = 0; i < n; ++i) {
    volatile float y = v[i] - c;
    volatile float t = s + y;
    c = (t - s) - y;
    s = t;
  }
  return s / (float)n;
}

/* ste_func.py:41 / :39 with scalar d, s */
API void fqo_fake_quant_scalar(const float* x, int64_t n, float d, float s, float lo, float hi, int use_clip,
                               float* y, float* code) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    float v = x[i];
    if (use_clip) v = clipf(v, lo, hi);
    const float c = roundf(v / d);
    if (code) code[i] = c;
    y[i] = c * s;
  }
}

/* ste_func.py:39 with a per-row scale tensor (weights): d_r = s_r + 1e-10f */
API void fqo_fake_quant_rows(const float* x, int64_t rows, int64_t len, const float* scale, float* y, float* code) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float s = scale[r];
    volatile float d = s + 1e-10f;
    for (int64_t i = r * len; i < (r + 1) * len; ++i) {
      const float c = roundf(x[i] / d);
      if (code) code[i] = c;
      y[i] = c * s;
    }
  }
}

/* distribution_calibrate.py:39-45: counts of the clipped non-zero values, `sc` already rounded to fp32 */
API void fqo_hist_counts(const float* x, int64_t n, float max_, float sc, int bins, int64_t* counts) {
  memset(counts, 0, sizeof(int64_t) * (size_t)(bins + 1));
#pragma omp parallel
  {
    int64_t* local = (int64_t*)calloc((size_t)bins + 1, sizeof(int64_t));
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < n; ++i) {
      float v = x[i];
      v = v < 0.f ? 0.f : v;          /* ndarray.clip(0, max_) */
      v = v > max_ ? max_ : v;
      if (v != 0.f) {
        const int b = (int)(v * sc);
        if (b >= 0 && b <= bins) local[b]++;
      }
    }
#pragma omp critical
    for (int b = 0; b <= bins; ++b) counts[b] += local[b];
    free(local);
  }
}

/* distribution_calibrate.py:138-166 for one candidate i; nep50 != 0: float32 sums, else float64 */
static double kl_one(const float* data, int n_data, int levels, int i, int nep50, double* cand, double* q, float* ref) {
  float last, total;
  if (nep50) {
    volatile float tail = 0.f, tot = 0.f;
    for (int j = i; j < n_data; ++j) tail = tail + data[j];
    last = data[i - 1] + tail;
    for (int j = 0; j < i - 1; ++j) tot = tot + data[j];
    tot = tot + last;
    total = tot;
  } else {
    double tail = 0.0, tot = 0.0;
    for (int j = i; j < n_data; ++j) tail += (double)data[j];
    last = (float)((double)data[i - 1] + tail);
    for (int j = 0; j < i - 1; ++j) tot += (double)data[j];
    tot += (double)last;
    total = (float)tot;
  }
  for (int k = 0; k < levels; ++k) cand[k] = 0.0;
  for (int j = 0; j < i; ++j) cand[(int)(((double)((int64_t)j * levels)) / (double)i)] += (double)data[j];
  double qsum = 0.0;
  for (int j = 0; j < i; ++j) {
    const float p = (j == i - 1 ? last : data[j]) / total;
    ref[j] = p;
    const double t = ((double)((int64_t)j * levels)) / (double)i;
    const int fl = (int)t;
    int ce = (int)ceil(t);
    if (ce > levels - 1) ce = levels - 1;
    double v = (cand[ce] - cand[fl]) * (t - (double)fl) + cand[fl];
    v = v * ((p != 0.f) ? 1.0 : 0.0);
    q[j] = v;
    qsum += v;
  }
  double div = 0.0;
  for (int j = 0; j < i; ++j) {
    const double qn = q[j] / qsum;
    if (qn != 0.0) {
      const double p = (double)ref[j];
      div += p * log(p / qn);
    }
  }
  return div;
}

/* distribution_calibrate.py:117-171: divergences for i in [min_bins, bins) and the first strict arg-min */
API int fqo_kl_calibrate(const float* data, int n_data, int levels, int min_bins, int bins, int nep50, double* div_out) {
#pragma omp parallel
  {
    double* cand = (double*)malloc(sizeof(double) * (size_t)levels);
    double* q = (double*)malloc(sizeof(double) * (size_t)bins);
    float* ref = (float*)malloc(sizeof(float) * (size_t)bins);
#pragma omp for schedule(dynamic, 8)
    for (int i = min_bins; i < bins; ++i) div_out[i] = kl_one(data, n_data, levels, i, nep50, cand, q, ref);
    free(cand);
    free(q);
    free(ref);
  }
  int best = min_bins;
  double best_div = INFINITY;
  for (int i = min_bins; i < bins; ++i)
    if (div_out[i] < best_div) {
      best_div = div_out[i];
      best = i;
    }
  return best;
}

/* ---------------------------------------------------------------------------------------------------
 * Winograd-domain per-channel weight quantisation: convert_conv2d.py:71-83, wino_matrix.py:28-60.
 * U = (G w) G^T per 3x3 kernel, s_o = max|U[o]| / qmax, Uq = roundf(U / (s_o + 1e-10f)) * s_o,
 * w_q = (GI Uq) GTI.  PARITY UNPINNED (nd.dot is a BLAS sgemm): every dot product is the chain
 * acc = a0*b0; acc = fmaf(a_k, b_k, acc) over the contraction index in ascending order -- with the real
 * single-rounding fmaf here, which is what the NumPy oracle's extended-precision emulation is checked against.
 * G [a,3], GI [3,a], GTI [a,3], a <= 8; w, wq: [cout, cin, 3, 3]; scales: [cout]. */
static void wino_u(const float* G, int a, const float* w, float* U) {
  float t[8][3];
  for (int p = 0; p < a; ++p)
    for (int c = 0; c < 3; ++c) {
      float acc = G[p * 3 + 0] * w[0 * 3 + c];
      acc = fmaf(G[p * 3 + 1], w[1 * 3 + c], acc);
      acc = fmaf(G[p * 3 + 2], w[2 * 3 + c], acc);
      t[p][c] = acc;
    }
  for (int p = 0; p < a; ++p)
    for (int q = 0; q < a; ++q) {
      float acc = t[p][0] * G[q * 3 + 0];
      acc = fmaf(t[p][1], G[q * 3 + 1], acc);
      acc = fmaf(t[p][2], G[q * 3 + 2], acc);
      U[p * a + q] = acc;
    }
}

API void fqo_wino_weight(const float* w, int64_t cout, int64_t cin, const float* G, const float* GI, const float* GTI,
                         int a, int bits, float* wq, float* scales) {
  const float qmax = (float)((1 << (bits - 1)) - 1);
#pragma omp parallel for schedule(static)
  for (int64_t o = 0; o < cout; ++o) {
    float U[64], m = 0.f;
    for (int64_t i = 0; i < cin; ++i) {
      wino_u(G, a, w + (o * cin + i) * 9, U);
      for (int e = 0; e < a * a; ++e) {
        const float v = fabsf(U[e]);
        if (v > m) m = v;
      }
    }
    const float s = m / qmax, d = s + 1e-10f;
    scales[o] = s;
    for (int64_t i = 0; i < cin; ++i) {
      float V[3][8];
      wino_u(G, a, w + (o * cin + i) * 9, U);
      for (int e = 0; e < a * a; ++e) U[e] = roundf(U[e] / d) * s;
      for (int r = 0; r < 3; ++r)
        for (int q = 0; q < a; ++q) {
          float acc = GI[r * a + 0] * U[0 * a + q];
          for (int p = 1; p < a; ++p) acc = fmaf(GI[r * a + p], U[p * a + q], acc);
          V[r][q] = acc;
        }
      float* out = wq + (o * cin + i) * 9;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          float acc = V[r][0] * GTI[0 * 3 + c];
          for (int q = 1; q < a; ++q) acc = fmaf(V[r][q], GTI[q * 3 + c], acc);
          out[r * 3 + c] = acc;
        }
    }
  }
}

/* convert_conv2d.py:150-153: per-channel mean and two-pass variance of y [N, C, HW]; sequential Kahan fp32 sums in
 * (n, h, w) order, `/ num` by fl32(num), `** 2` as the fp32 product */
API void fqo_channel_stats(const float* y, int64_t N, int64_t C, int64_t HW, float* mean, float* var) {
  const float num = (float)(N * HW);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < C; ++c) {
    volatile float s = 0.f, r = 0.f;
    for (int64_t n = 0; n < N; ++n) {
      const float* p = y + (n * C + c) * HW;
      for (int64_t i = 0; i < HW; ++i) {
        volatile float a = p[i] - r;
        volatile float t = s + a;
        r = (t - s) - a;
        s = t;
      }
    }
    const float m = s / num;
    mean[c] = m;
    s = 0.f;
    r = 0.f;
    for (int64_t n = 0; n < N; ++n) {
      const float* p = y + (n * C + c) * HW;
      for (int64_t i = 0; i < HW; ++i) {
        volatile float d = p[i] - m;
        volatile float sq = d * d;
        volatile float a = sq - r;
        volatile float t = s + a;
        r = (t - s) - a;
        s = t;
      }
    }
    var[c] = s / num;
  }
}
