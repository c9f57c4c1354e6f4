// Backward of the fake-BN fold for MANY blocks in one launch (QAT).
//   reference: quantize/convert/convert_conv2d.py:47-51 differentiated by MXNet's autograd, behind the identity
//   STE of ste_func.py:43-44:
//       W' = (W * gamma) / sd            sd = sqrt(var + 1e-10)
//       b' = (gamma * (b - mean)) / sd + beta
//   =>  dW     = (dW' / sd) * gamma
//       dgamma = sum_row((dW' / sd) * W) + (db' / sd) * (b - mean)
//       db     = (db' / sd) * gamma ,   dbeta = db'
// The op-by-op version is 6-8 framework launches per block (52 blocks in MobileNetV2: ~400 launches per step, a
// millisecond of a 9.7 ms CUDA-graph replay and far more in eager mode); this is one launch per 24 blocks.
// Not a parity kernel in the bit-exact sense (the reference's backward is whatever MXNet's autograd replays; the
// row sum has no defined order): tested against torch autograd of the un-fused formula to fp32 accuracy.
#include "fq_fused.cuh"

namespace fq {

constexpr int kFoldBwdBatch = 24;

struct FoldBwdJobDev {
  const float *dwq, *dbq, *w, *gamma, *mean, *var, *bias;
  float *dw, *dgamma, *dbias, *dbeta;
  int64_t L;
  int cout, first_block;
};

struct FoldBwdBatch {
  FoldBwdJobDev job[kFoldBwdBatch];
  int count;
};

// one warp per output channel
__global__ void __launch_bounds__(kThreads) fold_backward_multi_kernel(const __grid_constant__ FoldBwdBatch tb) {
  int j = 0;
  while (j + 1 < tb.count && (int)blockIdx.x >= tb.job[j + 1].first_block) ++j;
  const FoldBwdJobDev& q = tb.job[j];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = ((int)blockIdx.x - q.first_block) * (kThreads / 32) + warp;
  if (c >= q.cout) return;
  const float sd = __fsqrt_rn(__fadd_rn(__ldg(q.var + c), 1e-10f));
  const float g = __ldg(q.gamma + c);
  const float* dwq = q.dwq + (int64_t)c * q.L;
  const float* w = q.w + (int64_t)c * q.L;
  float* dw = q.dw + (int64_t)c * q.L;
  float acc = 0.f;
  for (int64_t i = lane; i < q.L; i += 32) {
    const float da = __fdiv_rn(__ldg(dwq + i), sd);
    dw[i] = __fmul_rn(da, g);
    acc = __fadd_rn(acc, __fmul_rn(da, __ldg(w + i)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  if (lane == 0) {
    float dgamma = acc;
    if (q.dbq != nullptr) {
      const float dbq = __ldg(q.dbq + c);
      const float dn = __fdiv_rn(dbq, sd);
      const float b = q.bias != nullptr ? __ldg(q.bias + c) : 0.f;
      dgamma = __fadd_rn(dgamma, __fmul_rn(dn, __fsub_rn(b, __ldg(q.mean + c))));
      if (q.dbias != nullptr) q.dbias[c] = __fmul_rn(dn, g);
      if (q.dbeta != nullptr) q.dbeta[c] = dbq;
    } else {
      if (q.dbias != nullptr) q.dbias[c] = 0.f;
      if (q.dbeta != nullptr) q.dbeta[c] = 0.f;
    }
    q.dgamma[c] = dgamma;
  }
}

}  // namespace fq

using namespace fq;

extern "C" {

int fq_fold_backward_multi(const FqFoldBwdJob* jobs, int n_jobs, void* stream) {
  const char* who = "fq_fold_backward_multi";
  FQ_REQUIRE(jobs != nullptr && n_jobs >= 1, "%s: no jobs", who);
  cudaStream_t st = (cudaStream_t)stream;
  int done = 0;
  while (done < n_jobs) {
    FoldBwdBatch tb = {};
    int blocks = 0, cnt = 0;
    for (; done + cnt < n_jobs && cnt < kFoldBwdBatch; ++cnt) {
      const FqFoldBwdJob& jb = jobs[done + cnt];
      View dwq, dbq, w, gamma, mean, var, bias, dw, dgamma, dbias, dbeta;
      FQ_TRY(view_of(jb.dwq, "fq_fold_backward_multi: dwq", false, &dwq));
      FQ_TRY(view_of(jb.dbq, "fq_fold_backward_multi: dbq", true, &dbq));
      FQ_TRY(view_of(jb.w, "fq_fold_backward_multi: w", false, &w));
      FQ_TRY(view_of(jb.gamma, "fq_fold_backward_multi: gamma", false, &gamma));
      FQ_TRY(view_of(jb.mean, "fq_fold_backward_multi: mean", false, &mean));
      FQ_TRY(view_of(jb.var, "fq_fold_backward_multi: var", false, &var));
      FQ_TRY(view_of(jb.bias, "fq_fold_backward_multi: bias", true, &bias));
      FQ_TRY(view_of(jb.dw, "fq_fold_backward_multi: dw", false, &dw));
      FQ_TRY(view_of(jb.dgamma, "fq_fold_backward_multi: dgamma", false, &dgamma));
      FQ_TRY(view_of(jb.dbias, "fq_fold_backward_multi: dbias", true, &dbias));
      FQ_TRY(view_of(jb.dbeta, "fq_fold_backward_multi: dbeta", true, &dbeta));
      FQ_REQUIRE(jb.w->ndim >= 1 && w.is_f32() && dwq.is_f32() && dw.is_f32() && w.numel > 0 && dwq.numel == w.numel &&
                     dw.numel == w.numel,
                 "%s: job %d: w, dwq and dw must be float32 of one shape", who, done + cnt);
      const int64_t cout = jb.w->shape[0];
      FQ_REQUIRE(gamma.is_f32() && mean.is_f32() && var.is_f32() && dgamma.is_f32() && gamma.numel == cout &&
                     mean.numel == cout && var.numel == cout && dgamma.numel == cout,
                 "%s: job %d: gamma, mean, var and dgamma must be float32 [Cout=%lld]", who, done + cnt, (long long)cout);
      FQ_REQUIRE((dbq.null || (dbq.is_f32() && dbq.numel == cout)) && (bias.null || (bias.is_f32() && bias.numel == cout)) &&
                     (dbias.null || (dbias.is_f32() && dbias.numel == cout)) &&
                     (dbeta.null || (dbeta.is_f32() && dbeta.numel == cout)),
                 "%s: job %d: dbq, bias, dbias and dbeta must be float32 [Cout]", who, done + cnt);
      FoldBwdJobDev& q = tb.job[cnt];
      q.dwq = dwq.as<const float>();
      q.dbq = dbq.null ? nullptr : dbq.as<const float>();
      q.w = w.as<const float>();
      q.gamma = gamma.as<const float>();
      q.mean = mean.as<const float>();
      q.var = var.as<const float>();
      q.bias = bias.null ? nullptr : bias.as<const float>();
      q.dw = dw.as<float>();
      q.dgamma = dgamma.as<float>();
      q.dbias = dbias.null ? nullptr : dbias.as<float>();
      q.dbeta = dbeta.null ? nullptr : dbeta.as<float>();
      q.L = w.numel / cout;
      q.cout = (int)cout;
      q.first_block = blocks;
      blocks += (int)((cout + kThreads / 32 - 1) / (kThreads / 32));
    }
    tb.count = cnt;
    fold_backward_multi_kernel<<<blocks, kThreads, 0, st>>>(tb);
    FQ_LAUNCH_CHECK("fold_backward_multi_kernel");
    done += cnt;
  }
  return 0;
}

}  // extern "C"
