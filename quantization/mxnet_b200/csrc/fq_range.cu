// K1: range reductions.  HBM-bound: 4 B/element read once, 128-bit loads, warp-shuffle reductions,
// a grid-level second pass done by the last block to finish (no host round trip, one launch).
//   reference: quantize/convert/convert_conv2d.py:56,75,86,92; convert_dense.py:41,54,60;
//              nn/quantized_conv.py:65,68-69; distribution_calibrate.py:34-35
#include "fq_fused.cuh"

namespace fq {

// ---------------------------------------------------------------------------
// min & max
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) minmax_kernel(const float* __restrict__ x, int64_t n, int64_t per_block,
                                                           Workspace* ws, float* __restrict__ out2) {
  __shared__ float red[64];
  __shared__ unsigned int s_last;
  const int64_t begin = (int64_t)blockIdx.x * per_block;
  const int64_t end = min(n, begin + per_block);
  float lo = INFINITY, hi = -INFINITY;
  for_range<false, false>(
      x, begin, end,
      [&](int64_t, float4 v) {
        lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
        hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
      },
      [&](int64_t, float v) {
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
      });
  lo = warp_min(lo);
  hi = warp_max(hi);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    red[warp] = lo;
    red[32 + warp] = hi;
  }
  __syncthreads();
  if (warp == 0) {
    lo = lane < (kThreads >> 5) ? red[lane] : INFINITY;
    hi = lane < (kThreads >> 5) ? red[32 + lane] : -INFINITY;
    lo = warp_min(lo);
    hi = warp_max(hi);
    if (lane == 0) {
      ws->minmax_part[2 * blockIdx.x] = lo;
      ws->minmax_part[2 * blockIdx.x + 1] = hi;
      __threadfence();
      s_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  lo = INFINITY;
  hi = -INFINITY;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += kThreads) {
    lo = fminf(lo, __ldcg(&ws->minmax_part[2 * b]));
    hi = fmaxf(hi, __ldcg(&ws->minmax_part[2 * b + 1]));
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  __syncthreads();
  if (lane == 0) {
    red[warp] = lo;
    red[32 + warp] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (kThreads >> 5); ++w) {
      lo = fminf(lo, red[w]);
      hi = fmaxf(hi, red[32 + w]);
    }
    out2[0] = lo;
    out2[1] = hi;
    ws->ticket = 0;
  }
}

// ---------------------------------------------------------------------------
// per-row absmax (+ optional Kahan mean and qparams) in one launch
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) rows_absmax_kernel(const float* __restrict__ x, int64_t n, int64_t rows,
                                                                int64_t L, int64_t per_block, Workspace* ws,
                                                                FinishParams fin) {
  __shared__ float red[32];
  __shared__ unsigned int s_last;
  const int64_t begin = (int64_t)blockIdx.x * per_block;
  const int64_t end = min(n, begin + per_block);
  absmax_segments<false>(x, begin, end, L, ws, red);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  finish_rows(ws, rows, fin);
  if (threadIdx.x == 0) ws->ticket = 0;
}

// one block per row of v[rows, n]: out[row] = Kahan mean of the row
__global__ void mean_kahan_kernel(const float* __restrict__ v_all, int64_t n, float* __restrict__ out) {
  __shared__ float stage[1024];
  const float* v = v_all + (int64_t)blockIdx.x * n;
  float s = 0.f, c = 0.f;
  for (int64_t base = 0; base < n; base += 1024) {
    const int m = (int)min((int64_t)1024, n - base);
    for (int i = threadIdx.x; i < m; i += blockDim.x) stage[i] = v[base + i];
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = 0; i < m; ++i) kahan_add(s, c, stage[i]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = __fdiv_rn(s, (float)n);
}

__global__ void scale_from_max_kernel(const float* __restrict__ max_, int bits, int is_signed, int lo_mode,
                                      int promotion, float* __restrict__ qp) {
  if (threadIdx.x == 0 && blockIdx.x == 0) compute_qparams(max_[0], bits, is_signed, lo_mode, promotion, qp);
}

static int launch_rows_absmax(const View& x, int64_t rows, Workspace* ws, const FinishParams& fin, cudaStream_t st) {
  const int64_t n = x.numel;
  const int64_t L = n / rows;
  int64_t per_block;
  const int grid = row_grid(n, L, sm_count() * 8, &per_block);
  rows_absmax_kernel<<<grid, kThreads, 0, st>>>(x.as<const float>(), n, rows, L, per_block, ws, fin);
  FQ_LAUNCH_CHECK("rows_absmax_kernel");
  return 0;
}

}  // namespace fq

using namespace fq;

extern "C" {

int fq_absmax_rows(const DLTensor* x_, int64_t rows, const DLTensor* out_, void* ws, void* stream) {
  View x, out;
  FQ_TRY(view_of(x_, "fq_absmax_rows: x", false, &x));
  FQ_TRY(view_of(out_, "fq_absmax_rows: out", false, &out));
  FQ_REQUIRE(x.is_f32() && out.is_f32(), "fq_absmax_rows: x and out must be float32");
  FQ_REQUIRE(ws != nullptr, "fq_absmax_rows: NULL workspace");
  FQ_REQUIRE(rows >= 1 && rows <= FQ_MAX_ROWS, "fq_absmax_rows: rows=%lld outside [1, %d]", (long long)rows, FQ_MAX_ROWS);
  FQ_REQUIRE(out.numel == rows, "fq_absmax_rows: out has %lld elements, expected rows=%lld", (long long)out.numel,
             (long long)rows);
  FQ_REQUIRE(x.numel % rows == 0, "fq_absmax_rows: numel %lld not divisible by rows %lld", (long long)x.numel,
             (long long)rows);
  cudaStream_t st = (cudaStream_t)stream;
  if (x.numel == 0) {
    FQ_CUDA(cudaMemsetAsync(out.data, 0, sizeof(float) * rows, st));
    return 0;
  }
  FinishParams fin = {};
  fin.out_rows = out.as<float>();
  return launch_rows_absmax(x, rows, (Workspace*)ws, fin, st);
}

int fq_minmax(const DLTensor* x_, const DLTensor* out2_, void* ws, void* stream) {
  View x, out;
  FQ_TRY(view_of(x_, "fq_minmax: x", false, &x));
  FQ_TRY(view_of(out2_, "fq_minmax: out2", false, &out));
  FQ_REQUIRE(x.is_f32() && out.is_f32(), "fq_minmax: x and out2 must be float32");
  FQ_REQUIRE(out.numel == 2, "fq_minmax: out2 must hold 2 floats");
  FQ_REQUIRE(x.numel > 0, "fq_minmax: empty tensor has no min/max");
  FQ_REQUIRE(ws != nullptr, "fq_minmax: NULL workspace");
  int64_t per_block;
  const int grid = slice_grid(x.numel, min(sm_count() * 8, 4096), &per_block);
  minmax_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(x.as<const float>(), x.numel, per_block, (Workspace*)ws,
                                                             out.as<float>());
  FQ_LAUNCH_CHECK("minmax_kernel");
  return 0;
}

int fq_mean_kahan(const DLTensor* v_, const DLTensor* out_, void* stream) {
  View v, out;
  FQ_TRY(view_of(v_, "fq_mean_kahan: v", false, &v));
  FQ_TRY(view_of(out_, "fq_mean_kahan: out", false, &out));
  FQ_REQUIRE(v.is_f32() && out.is_f32(), "fq_mean_kahan: float32 only");
  FQ_REQUIRE(v.numel > 0 && out.numel >= 1, "fq_mean_kahan: empty input or output");
  // v: [n] -> out[0], or [rows, n] -> out[rows] (every layer's per-sample maxima in one launch)
  const int64_t rows = (v_->ndim >= 2) ? v.numel / v_->shape[v_->ndim - 1] : 1;
  FQ_REQUIRE(out.numel >= rows && rows <= 65535, "fq_mean_kahan: out needs one element per row (%lld rows)", (long long)rows);
  mean_kahan_kernel<<<(unsigned)rows, kThreads, 0, (cudaStream_t)stream>>>(v.as<const float>(), v.numel / rows,
                                                                          out.as<float>());
  FQ_LAUNCH_CHECK("mean_kahan_kernel");
  return 0;
}

int fq_input_range(const DLTensor* x_, int64_t n_samples, const DLTensor* per_sample_, const DLTensor* cur_max_,
                   void* ws, void* stream) {
  View x, ps, cur;
  FQ_TRY(view_of(x_, "fq_input_range: x", false, &x));
  FQ_TRY(view_of(per_sample_, "fq_input_range: per_sample", true, &ps));
  FQ_TRY(view_of(cur_max_, "fq_input_range: cur_max", false, &cur));
  FQ_REQUIRE(x.is_f32() && cur.is_f32() && (ps.null || ps.is_f32()), "fq_input_range: float32 only");
  FQ_REQUIRE(ws != nullptr, "fq_input_range: NULL workspace");
  FQ_REQUIRE(n_samples >= 1 && n_samples <= FQ_MAX_ROWS, "fq_input_range: n_samples=%lld outside [1, %d]",
             (long long)n_samples, FQ_MAX_ROWS);
  FQ_REQUIRE(x.numel > 0 && x.numel % n_samples == 0, "fq_input_range: numel %lld not a positive multiple of n_samples %lld",
             (long long)x.numel, (long long)n_samples);
  FQ_REQUIRE(ps.null || ps.numel == n_samples, "fq_input_range: per_sample must have n_samples elements");
  FQ_REQUIRE(cur.numel >= 1, "fq_input_range: cur_max is empty");
  FinishParams fin = {};
  fin.out_rows = ps.null ? nullptr : ps.as<float>();
  fin.out_mean = cur.as<float>();
  return launch_rows_absmax(x, n_samples, (Workspace*)ws, fin, (cudaStream_t)stream);
}

int fq_scale_from_max(const DLTensor* max__, int bits, int is_signed, int lo_mode, int promotion,
                      const DLTensor* qparams_, void* stream) {
  View mx, qp;
  FQ_TRY(view_of(max__, "fq_scale_from_max: max_", false, &mx));
  FQ_TRY(view_of(qparams_, "fq_scale_from_max: qparams", false, &qp));
  FQ_REQUIRE(mx.is_f32() && qp.is_f32(), "fq_scale_from_max: float32 only");
  FQ_REQUIRE(mx.numel >= 1 && qp.numel == 4, "fq_scale_from_max: max_ needs 1 element, qparams 4");
  FQ_TRY(check_quant_args("fq_scale_from_max", bits, lo_mode, promotion));
  scale_from_max_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(mx.as<const float>(), bits, is_signed, lo_mode, promotion,
                                                            qp.as<float>());
  FQ_LAUNCH_CHECK("scale_from_max_kernel");
  return 0;
}

}  // extern "C"
