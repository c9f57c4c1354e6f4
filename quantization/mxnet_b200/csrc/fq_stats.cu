// Per-channel batch statistics of a conv output for the fake-BN EMA.
//   reference: quantize/convert/convert_conv2d.py:148-153
//     mean = y.sum(axis=(0,2,3)) / num ;  var = ((y - mean) ** 2).sum(axis=(0,2,3)) / num
//
// The reference makes six passes over y (sum, broadcast subtract, square, sum, two divides); its CPU sums are
// sequential Kahan-compensated fp32, i.e. within ~1 ULP of the exact sum of the fp32 terms.  Here y is read ONCE
// (4 B/element): every block accumulates, in float64, the shifted moments
//     S1 = sum (y - K),   S2 = sum (y - K)^2,     K = y[0, c, 0, 0] (the channel's first element on this rank)
// and the last block of a channel derives
//     mean = fl32( fl32(S1 + n K) / fl32(n) )
//     var  = fl32( fl32(S2 - 2 (mean - K) S1 + n (mean - K)^2) / fl32(n) )
// which is algebraically sum (y - mean)^2 with the reference's fp32-rounded mean.  With K inside the data the
// three terms are all O(n sigma^2): no cancellation, so the float64 result is exact to ~2^-45 and its fp32
// rounding lands within 2 ULP of the reference's Kahan sums (tests/test_gpu_kernels.py states the bound measured).
// Sums are taken in a fixed order (per thread, shuffle tree, per-block partials in block order): deterministic.
//
// Data parallel: every rank exports its {n, S1, S2, K} records; the records of all ranks are combined by the same
// formula (fq_channel_stats_finish), so R ranks reproduce the single-GPU statistics of the global batch to the
// same bound with ONE collective per step for all layers, instead of two all-reduces per layer.
#include "fq_fused.cuh"

namespace fq {

constexpr int kStatSplitMax = 64;
constexpr int kStatRec = 4;          // {n, S1, S2, K}

struct StatAcc {
  double s1 = 0.0, s2 = 0.0;
  __device__ __forceinline__ void add(float v, double K) {
    const double d = (double)v - K;
    s1 += d;
    s2 = fma(d, d, s2);
  }
  __device__ __forceinline__ void add4(float4 v, double K) {
    add(v.x, K);
    add(v.y, K);
    add(v.z, K);
    add(v.w, K);
  }
};

// {n, S1, S2, K} records of R ranks (stride `rs` doubles between ranks) -> mean, var.
__device__ __forceinline__ void stats_combine(const double* __restrict__ rec, int R, int64_t rs, float* mean, float* var) {
  double n = 0.0, sy = 0.0;
  for (int r = 0; r < R; ++r) {
    const double* q = rec + (int64_t)r * rs;
    n += q[0];
    sy += q[1] + q[0] * q[3];
  }
  // y.sum(axis) is an fp32 NDArray; `/ num_samples` is _div_scalar by fl32(num)
  const float nf = (float)n;
  const float m = __fdiv_rn((float)sy, nf);
  double m2 = 0.0;
  for (int r = 0; r < R; ++r) {
    const double* q = rec + (int64_t)r * rs;
    const double d = (double)m - q[3];
    m2 += q[2] + d * (q[0] * d - 2.0 * q[1]);
  }
  if (m2 < 0.0) m2 = 0.0;
  *mean = m;
  *var = __fdiv_rn((float)m2, nf);
}

// Block (c, s) owns samples [N s / S, N (s+1) / S) of channel c.
__global__ void __launch_bounds__(kThreads) channel_stats_kernel(const float* __restrict__ y, int64_t N, int64_t C,
                                                                 int64_t HW, int S, Workspace* ws,
                                                                 double* __restrict__ parts, float* __restrict__ mean,
                                                                 float* __restrict__ var) {
  __shared__ double red[2][kThreads / 32];
  __shared__ unsigned int s_last;
  const int64_t c = blockIdx.x / S;
  const int s = blockIdx.x % S;
  const int64_t n0 = N * s / S, n1 = N * (s + 1) / S;
  const double K = (double)__ldg(y + c * HW);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nw = kThreads / 32;
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
  StatAcc a, b, e, f;    // independent chains: four 16 B loads in flight per thread in the streaming case
  if (HW >= 1024) {      // the whole block walks one (n, c) plane at a time
    for (int64_t n = n0; n < n1; ++n) {
      const float* p = y + (n * C + c) * HW;
      if (vec) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        const int64_t nvec = HW >> 2;
        int64_t i = threadIdx.x;
        for (; i + 3 * kThreads < nvec; i += 4 * kThreads) {
          const float4 v0 = ld_stream(p4 + i), v1 = ld_stream(p4 + i + kThreads);
          const float4 v2 = ld_stream(p4 + i + 2 * kThreads), v3 = ld_stream(p4 + i + 3 * kThreads);
          a.add4(v0, K);
          b.add4(v1, K);
          e.add4(v2, K);
          f.add4(v3, K);
        }
        for (; i < nvec; i += kThreads) a.add4(ld_stream(p4 + i), K);
      } else {
        int64_t i = threadIdx.x;
        for (; i + kThreads < HW; i += 2 * kThreads) {
          const float v0 = __ldg(p + i), v1 = __ldg(p + i + kThreads);
          a.add(v0, K);
          b.add(v1, K);
        }
        if (i < HW) a.add(__ldg(p + i), K);
      }
    }
  } else {               // small planes (7x7 ... 28x28): one warp per plane, two planes in flight
    for (int64_t n = n0 + warp; n < n1; n += 2 * nw) {
      const float* p = y + (n * C + c) * HW;
      const bool two = n + nw < n1;
      const float* q = two ? p + (int64_t)nw * C * HW : p;
      if (vec) {
        const float4 *p4 = reinterpret_cast<const float4*>(p), *q4 = reinterpret_cast<const float4*>(q);
        const int nvec = (int)(HW >> 2);
        for (int i = lane; i < nvec; i += 32) {
          const float4 v0 = ld_stream(p4 + i);
          if (two) {
            const float4 v1 = ld_stream(q4 + i);
            b.add4(v1, K);
          }
          a.add4(v0, K);
        }
      } else {
        for (int i = lane; i < (int)HW; i += 32) {
          const float v0 = __ldg(p + i);
          if (two) b.add(__ldg(q + i), K);
          a.add(v0, K);
        }
      }
    }
  }
  double s1 = (a.s1 + b.s1) + (e.s1 + f.s1), s2 = (a.s2 + b.s2) + (e.s2 + f.s2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    red[0][warp] = s1;
    red[1][warp] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s1 = red[0][0];
    s2 = red[1][0];
#pragma unroll
    for (int w = 1; w < nw; ++w) {
      s1 += red[0][w];
      s2 += red[1][w];
    }
    double* part = ws->stats_part + 2 * (int64_t)blockIdx.x;
    part[0] = s1;
    part[1] = s2;
    __threadfence();
    s_last = (atomicAdd(&ws->rowmax[c], 1u) == (unsigned)S - 1);      // rowmax[c] doubles as the channel's ticket
    if (s_last) {
      __threadfence();
      const double* all = ws->stats_part + 2 * c * S;
      double rec[kStatRec] = {(double)(N * HW), 0.0, 0.0, K};
      for (int k = 0; k < S; ++k) {
        rec[1] += __ldcg(all + 2 * k);
        rec[2] += __ldcg(all + 2 * k + 1);
      }
      ws->rowmax[c] = 0u;
      if (parts != nullptr) {
#pragma unroll
        for (int k = 0; k < kStatRec; ++k) parts[c * kStatRec + k] = rec[k];
      }
      if (mean != nullptr) stats_combine(rec, 1, 0, mean + c, var + c);
    }
  }
}

__global__ void channel_stats_finish_kernel(const double* __restrict__ parts, int R, int64_t C, float* __restrict__ mean,
                                            float* __restrict__ var) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  stats_combine(parts + c * kStatRec, R, C * kStatRec, mean + c, var + c);
}

}  // namespace fq

using namespace fq;

extern "C" {

int fq_channel_stats(const DLTensor* y_, const DLTensor* mean_, const DLTensor* var_, const DLTensor* parts_, void* ws,
                     void* stream) {
  const char* who = "fq_channel_stats";
  View y, mean, var, parts;
  FQ_TRY(view_of(y_, "fq_channel_stats: y", false, &y));
  FQ_TRY(view_of(mean_, "fq_channel_stats: mean", true, &mean));
  FQ_TRY(view_of(var_, "fq_channel_stats: var", true, &var));
  FQ_TRY(view_of(parts_, "fq_channel_stats: parts", true, &parts));
  FQ_REQUIRE(ws != nullptr, "%s: NULL workspace", who);
  FQ_REQUIRE(y.is_f32(), "%s: y must be float32", who);
  FQ_REQUIRE(y_->ndim >= 2, "%s: y must be [N, C, ...]", who);
  const int64_t N = y_->shape[0], C = y_->shape[1];
  FQ_REQUIRE(N >= 1 && C >= 1 && y.numel > 0, "%s: empty tensor", who);
  const int64_t HW = y.numel / (N * C);
  FQ_REQUIRE(mean.null == var.null, "%s: mean and var must be given together", who);
  FQ_REQUIRE(!mean.null || !parts.null, "%s: nothing to write (mean/var and parts are all NULL)", who);
  FQ_REQUIRE(mean.null || (mean.is_f32() && var.is_f32() && mean.numel == C && var.numel == C),
             "%s: mean and var must be float32 [C=%lld]", who, (long long)C);
  FQ_REQUIRE(parts.null || (parts.code == kDLFloat && parts.bits == 64 && parts.numel == C * kStatRec),
             "%s: parts must be float64 [C=%lld, 4]", who, (long long)C);
  FQ_REQUIRE(C <= FQ_MAX_STAT_BLOCKS, "%s: C=%lld exceeds %d", who, (long long)C, FQ_MAX_STAT_BLOCKS);
  int S = (int)((int64_t)sm_count() * 8 / C);
  if (S < 1) S = 1;
  if (S > kStatSplitMax) S = kStatSplitMax;
  if (S > N) S = (int)N;
  while ((int64_t)C * S > FQ_MAX_STAT_BLOCKS) --S;
  channel_stats_kernel<<<(unsigned)(C * S), kThreads, 0, (cudaStream_t)stream>>>(
      y.as<const float>(), N, C, HW, S, (Workspace*)ws, parts.null ? nullptr : parts.as<double>(),
      mean.null ? nullptr : mean.as<float>(), var.null ? nullptr : var.as<float>());
  FQ_LAUNCH_CHECK("channel_stats_kernel");
  return 0;
}

int fq_channel_stats_finish(const DLTensor* parts_, const DLTensor* mean_, const DLTensor* var_, void* stream) {
  const char* who = "fq_channel_stats_finish";
  View parts, mean, var;
  FQ_TRY(view_of(parts_, "fq_channel_stats_finish: parts", false, &parts));
  FQ_TRY(view_of(mean_, "fq_channel_stats_finish: mean", false, &mean));
  FQ_TRY(view_of(var_, "fq_channel_stats_finish: var", false, &var));
  FQ_REQUIRE(parts.code == kDLFloat && parts.bits == 64 && mean.is_f32() && var.is_f32(),
             "%s: parts float64, mean/var float32", who);
  const int64_t C = mean.numel;
  FQ_REQUIRE(C >= 1 && var.numel == C && parts.numel >= C * kStatRec && parts.numel % (C * kStatRec) == 0,
             "%s: parts must be [R, C=%lld, 4]", who, (long long)C);
  const int64_t R = parts.numel / (C * kStatRec);
  FQ_REQUIRE(R <= 4096, "%s: R=%lld ranks?", who, (long long)R);
  channel_stats_finish_kernel<<<(unsigned)((C + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      parts.as<const double>(), (int)R, C, mean.as<float>(), var.as<float>());
  FQ_LAUNCH_CHECK("channel_stats_finish_kernel");
  return 0;
}

}  // extern "C"
