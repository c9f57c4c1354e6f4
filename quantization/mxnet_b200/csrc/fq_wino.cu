// Winograd-domain per-channel weight fake-quantisation (gen_conv2d_converter(wino_quantize="F23"|"F43"|"F63")).
//   reference: quantize/convert/convert_conv2d.py:71-83, quantize/convert/wino_matrix.py:28-60
//
//   U    = G w G^T          per 3x3 kernel w[o][i]; G is (a x 3), a = 4 / 6 / 8             (:71-73)
//   M_o  = max |U[o,:,:,:]| ; s_o = M_o / (2^(bits-1)-1) ; Uq = roundf(U / (s_o + 1e-10)) * s_o   (:74-79)
//   w_q  = G+ Uq (G^T)+     with the float32 pseudo-inverses the host computed (np.linalg.pinv)   (:80-83)
//
// One thread owns one 3x3 kernel: 9 loads, four tiny matrix products in registers, 9 stores; the a x a
// transform is produced one row at a time so that only G+ Uq (3 x a) is kept.  Weights are <= 2.4 M elements,
// so the two launches (range, then quantise with the transform recomputed) are latency-bound, like the plain
// weight path.
//
// Arithmetic contract (parity unpinned: MXNet's nd.dot is a BLAS sgemm whose summation order is not defined):
// every dot product runs over its contraction index in ascending order as `acc = a0*b0; acc = fma(ak, bk, acc)`,
// in the order the reference multiplies: (G w) G^T and (G+ Uq) (G^T)+.
#include "fq_fused.cuh"

namespace fq {

constexpr int kWinoMaxA = 8;

struct WinoArgs {
  const float* w;         // [cout, cin, 3, 3]
  float* w_out;
  float* scale_out;       // nullable [cout]
  const float *G, *GI, *GTI;   // [a,3], [3,a], [a,3]
  int a, cin, cout, bits;
  int64_t kernels;        // cout * cin
  Workspace* ws;
};

struct WinoMats {
  float G[kWinoMaxA * 3], GI[3 * kWinoMaxA], GTI[kWinoMaxA * 3];
};

__device__ __forceinline__ void load_mats(WinoMats* m, const float* G, const float* GI, const float* GTI, int a) {
  for (int i = threadIdx.x; i < 3 * a; i += blockDim.x) {
    m->G[i] = __ldg(G + i);
    m->GI[i] = __ldg(GI + i);
    m->GTI[i] = __ldg(GTI + i);
  }
  __syncthreads();
}

// row p of U = (G w) G^T
template <int A>
__device__ __forceinline__ void wino_row(const WinoMats& m, const float (&w)[9], int p, float (&u)[A]) {
  float t[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {          // (G w)[p][c] = sum_r G[p][r] w[r][c]
    float acc = __fmul_rn(m.G[p * 3 + 0], w[0 * 3 + c]);
    acc = __fmaf_rn(m.G[p * 3 + 1], w[1 * 3 + c], acc);
    acc = __fmaf_rn(m.G[p * 3 + 2], w[2 * 3 + c], acc);
    t[c] = acc;
  }
#pragma unroll
  for (int q = 0; q < A; ++q) {          // U[p][q] = sum_c (G w)[p][c] G[q][c]
    float acc = __fmul_rn(t[0], m.G[q * 3 + 0]);
    acc = __fmaf_rn(t[1], m.G[q * 3 + 1], acc);
    acc = __fmaf_rn(t[2], m.G[q * 3 + 2], acc);
    u[q] = acc;
  }
}

template <int A, int PHASE>
__global__ void __launch_bounds__(kThreads) wino_weight_kernel(WinoArgs a) {
  __shared__ WinoMats m;
  __shared__ unsigned int s_last;
  load_mats(&m, a.G, a.GI, a.GTI, A);
  const float qmax = (float)((1 << (a.bits - 1)) - 1);
  if (PHASE == 1 && blockIdx.x == 0 && a.scale_out != nullptr)
    for (int r = threadIdx.x; r < a.cout; r += blockDim.x)
      a.scale_out[r] = __fdiv_rn(__uint_as_float(__ldcg(&a.ws->rowmax[r])), qmax);

  for (int64_t k0 = (int64_t)blockIdx.x * blockDim.x; k0 < a.kernels; k0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = k0 + threadIdx.x;
    const bool live = k < a.kernels;
    const int o = live ? (int)(k / a.cin) : -1;
    float w[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) w[e] = live ? __ldg(a.w + k * 9 + e) : 0.f;

    if (PHASE == 0) {
      float mx = 0.f;
#pragma unroll
      for (int p = 0; p < A; ++p) {
        float u[A];
        wino_row<A>(m, w, p, u);
#pragma unroll
        for (int q = 0; q < A; ++q) mx = fmaxf(mx, fabsf(u[q]));
      }
      // lanes of a warp hold consecutive kernels: usually one output channel
      const int o0 = __shfl_sync(0xffffffffu, o, 0);
      if (__all_sync(0xffffffffu, o == o0)) {
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0 && live) atomicMax(&a.ws->rowmax[o], __float_as_uint(mx));
      } else if (live) {
        atomicMax(&a.ws->rowmax[o], __float_as_uint(mx));
      }
    } else if (live) {
      const float s = __fdiv_rn(__uint_as_float(__ldcg(&a.ws->rowmax[o])), qmax);     // :76
      const QDiv qd = QDiv::make(__fadd_rn(s, 1e-10f));                               // ste_func.py:39
      float v[3][A];                        // G+ Uq, accumulated over the rows p of Uq
#pragma unroll
      for (int p = 0; p < A; ++p) {
        float u[A];
        wino_row<A>(m, w, p, u);
#pragma unroll
        for (int q = 0; q < A; ++q) {
          const float uq = __fmul_rn(qd.code(u[q]), s);
#pragma unroll
          for (int r = 0; r < 3; ++r)
            v[r][q] = (p == 0) ? __fmul_rn(m.GI[r * A + 0], uq) : __fmaf_rn(m.GI[r * A + p], uq, v[r][q]);
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {       // w_q[r][c] = sum_q (G+ Uq)[r][q] (G^T)+[q][c]
          float acc = __fmul_rn(v[r][0], m.GTI[0 * 3 + c]);
#pragma unroll
          for (int q = 1; q < A; ++q) acc = __fmaf_rn(v[r][q], m.GTI[q * 3 + c], acc);
          a.w_out[k * 9 + r * 3 + c] = acc;
        }
    }
  }

  if (PHASE == 0) return;
  // restore the workspace invariant once every block has consumed the row maxima
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&a.ws->ticket2, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    for (int r = threadIdx.x; r < a.cout; r += blockDim.x) a.ws->rowmax[r] = 0u;
    if (threadIdx.x == 0) a.ws->ticket2 = 0;
  }
}

// Straight-through backward, in the order autograd replays the four products:
//   dX = dw_q ((G^T)+)^T ; dUq = (G+)^T dX ; [STE: dU = dUq] ; dT = dU G ; dw = G^T dT
template <int A>
__global__ void __launch_bounds__(kThreads) wino_backward_kernel(const float* __restrict__ dwq, float* __restrict__ dw,
                                                                 const float* G, const float* GI, const float* GTI,
                                                                 int64_t kernels) {
  __shared__ WinoMats m;
  load_mats(&m, G, GI, GTI, A);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < kernels; k += (int64_t)gridDim.x * blockDim.x) {
    float g[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) g[e] = __ldg(dwq + k * 9 + e);
    float dx[3][A];                          // dX[r][q] = sum_c dw_q[r][c] GTI[q][c]
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < A; ++q) {
        float acc = __fmul_rn(g[r * 3 + 0], m.GTI[q * 3 + 0]);
        acc = __fmaf_rn(g[r * 3 + 1], m.GTI[q * 3 + 1], acc);
        acc = __fmaf_rn(g[r * 3 + 2], m.GTI[q * 3 + 2], acc);
        dx[r][q] = acc;
      }
    float out[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) out[e] = 0.f;
#pragma unroll
    for (int p = 0; p < A; ++p) {
      float dt[3];                           // dT[p][c] = sum_q dU[p][q] G[q][c],  dU[p][q] = sum_r GI[r][p] dX[r][q]
#pragma unroll
      for (int c = 0; c < 3; ++c) dt[c] = 0.f;
#pragma unroll
      for (int q = 0; q < A; ++q) {
        float du = __fmul_rn(m.GI[0 * A + p], dx[0][q]);
        du = __fmaf_rn(m.GI[1 * A + p], dx[1][q], du);
        du = __fmaf_rn(m.GI[2 * A + p], dx[2][q], du);
#pragma unroll
        for (int c = 0; c < 3; ++c) dt[c] = (q == 0) ? __fmul_rn(du, m.G[q * 3 + c]) : __fmaf_rn(du, m.G[q * 3 + c], dt[c]);
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)            // dw[r][c] = sum_p G[p][r] dT[p][c]
#pragma unroll
        for (int c = 0; c < 3; ++c)
          out[r * 3 + c] = (p == 0) ? __fmul_rn(m.G[p * 3 + r], dt[c]) : __fmaf_rn(m.G[p * 3 + r], dt[c], out[r * 3 + c]);
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) dw[k * 9 + e] = out[e];
  }
}

static int wino_mats(const char* who, const DLTensor* G_, const DLTensor* GI_, const DLTensor* GTI_, View* G, View* GI,
                     View* GTI, int* a) {
  FQ_TRY(view_of(G_, who, false, G));
  FQ_TRY(view_of(GI_, who, false, GI));
  FQ_TRY(view_of(GTI_, who, false, GTI));
  FQ_REQUIRE(G->is_f32() && GI->is_f32() && GTI->is_f32(), "%s: G, GI and GTI must be float32", who);
  FQ_REQUIRE(G_->ndim == 2 && G_->shape[1] == 3 && (G_->shape[0] == 4 || G_->shape[0] == 6 || G_->shape[0] == 8),
             "%s: G must be [a, 3] with a in {4, 6, 8} (F23 / F43 / F63)", who);
  *a = (int)G_->shape[0];
  FQ_REQUIRE(GI_->ndim == 2 && GI_->shape[0] == 3 && GI_->shape[1] == *a && GTI_->ndim == 2 && GTI_->shape[0] == *a &&
                 GTI_->shape[1] == 3,
             "%s: GI must be [3, %d] and GTI [%d, 3]", who, *a, *a);
  return 0;
}

}  // namespace fq

using namespace fq;

extern "C" {

int fq_quant_weight_wino(const DLTensor* w_, const DLTensor* G_, const DLTensor* GI_, const DLTensor* GTI_, int bits,
                         const DLTensor* w_out_, const DLTensor* scale_out_, void* ws, void* stream) {
  const char* who = "fq_quant_weight_wino";
  View w, out, so, G, GI, GTI;
  int a = 0;
  FQ_TRY(view_of(w_, "fq_quant_weight_wino: w", false, &w));
  FQ_TRY(view_of(w_out_, "fq_quant_weight_wino: w_out", false, &out));
  FQ_TRY(view_of(scale_out_, "fq_quant_weight_wino: scale_out", true, &so));
  FQ_TRY(wino_mats(who, G_, GI_, GTI_, &G, &GI, &GTI, &a) == 0);
  FQ_REQUIRE(ws != nullptr, "%s: NULL workspace", who);
  FQ_REQUIRE(w.is_f32() && out.is_f32() && out.numel == w.numel, "%s: w and w_out must be float32 of equal size", who);
  FQ_REQUIRE(w_->ndim == 4 && w_->shape[2] == 3 && w_->shape[3] == 3, "%s: w must be [Cout, Cin, 3, 3]", who);
  FQ_REQUIRE(bits >= 2 && bits <= 24, "%s: bits=%d outside [2, 24]", who, bits);
  const int64_t cout = w_->shape[0], cin = w_->shape[1];
  FQ_REQUIRE(cout >= 1 && cout <= FQ_MAX_ROWS && cin >= 1 && cin <= INT32_MAX, "%s: Cout=%lld outside [1, %d]", who,
             (long long)cout, FQ_MAX_ROWS);
  FQ_REQUIRE(so.null || (so.is_f32() && so.numel == cout), "%s: scale_out must be float32 [Cout]", who);
  WinoArgs args = {};
  args.w = w.as<const float>();
  args.w_out = out.as<float>();
  args.scale_out = so.null ? nullptr : so.as<float>();
  args.G = G.as<const float>();
  args.GI = GI.as<const float>();
  args.GTI = GTI.as<const float>();
  args.a = a;
  args.cin = (int)cin;
  args.cout = (int)cout;
  args.bits = bits;
  args.kernels = cout * cin;
  args.ws = (Workspace*)ws;
  const int64_t blocks64 = (args.kernels + kThreads - 1) / kThreads;
  const int grid = (int)(blocks64 > sm_count() * 8 ? sm_count() * 8 : blocks64);
  cudaStream_t st = (cudaStream_t)stream;
  switch (a) {
    case 4:
      wino_weight_kernel<4, 0><<<grid, kThreads, 0, st>>>(args);
      wino_weight_kernel<4, 1><<<grid, kThreads, 0, st>>>(args);
      break;
    case 6:
      wino_weight_kernel<6, 0><<<grid, kThreads, 0, st>>>(args);
      wino_weight_kernel<6, 1><<<grid, kThreads, 0, st>>>(args);
      break;
    default:
      wino_weight_kernel<8, 0><<<grid, kThreads, 0, st>>>(args);
      wino_weight_kernel<8, 1><<<grid, kThreads, 0, st>>>(args);
      break;
  }
  FQ_LAUNCH_CHECK("wino_weight_kernel");
  return 0;
}

int fq_wino_backward(const DLTensor* dwq_, const DLTensor* G_, const DLTensor* GI_, const DLTensor* GTI_,
                     const DLTensor* dw_, void* stream) {
  const char* who = "fq_wino_backward";
  View dwq, dw, G, GI, GTI;
  int a = 0;
  FQ_TRY(view_of(dwq_, "fq_wino_backward: dwq", false, &dwq));
  FQ_TRY(view_of(dw_, "fq_wino_backward: dw", false, &dw));
  FQ_TRY(wino_mats(who, G_, GI_, GTI_, &G, &GI, &GTI, &a) == 0);
  FQ_REQUIRE(dwq.is_f32() && dw.is_f32() && dw.numel == dwq.numel && dwq.numel % 9 == 0,
             "%s: dwq and dw must be float32 [.., 3, 3] of equal size", who);
  if (dwq.numel == 0) return 0;
  const int64_t kernels = dwq.numel / 9;
  const int64_t blocks64 = (kernels + kThreads - 1) / kThreads;
  const int grid = (int)(blocks64 > sm_count() * 8 ? sm_count() * 8 : blocks64);
  cudaStream_t st = (cudaStream_t)stream;
  if (a == 4)
    wino_backward_kernel<4><<<grid, kThreads, 0, st>>>(dwq.as<const float>(), dw.as<float>(), G.as<const float>(),
                                                        GI.as<const float>(), GTI.as<const float>(), kernels);
  else if (a == 6)
    wino_backward_kernel<6><<<grid, kThreads, 0, st>>>(dwq.as<const float>(), dw.as<float>(), G.as<const float>(),
                                                        GI.as<const float>(), GTI.as<const float>(), kernels);
  else
    wino_backward_kernel<8><<<grid, kThreads, 0, st>>>(dwq.as<const float>(), dw.as<float>(), G.as<const float>(),
                                                        GI.as<const float>(), GTI.as<const float>(), kernels);
  FQ_LAUNCH_CHECK("wino_backward_kernel");
  return 0;
}

}  // extern "C"
