// QConv2D integer convolution on the 5th-generation tensor cores (outside the north star).
//   reference: nn/quantized_conv.py:106-159 -- pad, quantise input and weight to 8-bit integer codes, im2col in a
//   Python double loop, `F.dot` on float32 casts of the codes, cast to int32, + int32 bias, activation, dequantise.
//
// Here: (1) fq_qconv_pack_input quantises the fp32 NCHW input straight into zero-padded NHWC 8-bit codes (one HBM
// pass, 5 B/element), (2) qconv_igemm_kernel is an implicit GEMM  D[m, co] = sum_k A[m, k] * W[co, k]  with
// m = (n, oh, ow), k = (kh, kw, ci): the loader gathers 16-byte runs of input channels for every (m, kh, kw) directly
// from the NHWC codes into shared memory in the canonical 128-byte-swizzled K-major layout (no im2col buffer),
// `tcgen05.mma.cta_group::1.kind::i8` (M=128, N<=128, K=32; issued by one thread) accumulates exact int32 in TENSOR
// MEMORY, completion is tracked with tcgen05.commit -> mbarrier, and the epilogue reads the accumulators back with
// tcgen05.ld, adds the int32 bias, applies ReLU and dequantises -- fused, the int32 tensor never touches HBM.
// The accumulators are exact integers, so the result equals the reference's float-code dot product bit for bit
// whenever that one is exact (|sum| < 2^24) and is the mathematically right int32 beyond.
//
// Operands are staged with cp.async (LDGSTS, 16 B) rather than TMA: the A operand is an im2col GATHER whose rows
// change with (kh, kw) and with the image border, which a tiled tensor map cannot express without the im2col mode;
// generic-proxy writes are ordered before the tensor core's async-proxy reads with fence.proxy.async.
#include "fq_fused.cuh"

namespace fq {

constexpr int kMmaM = 128;            // output pixels per CTA
constexpr int kMmaK = 128;            // int8 elements per k-block = one 128 B swizzle row
constexpr int kMmaStages = 3;
constexpr int kMmaThreads = 256;

struct QConvArgs {
  const signed char* xq;      // [N, Hp, Wp, C] codes, spatially padded
  const signed char* wq;      // [Cout, KH, KW, Cg] codes
  const int* bias_q;          // [Cout] or NULL
  const float* s_in;          // device scalars
  const float* s_w;
  float* out;                 // [N, Cout, Ho, Wo]
  int N, C, Hp, Wp, Cout, KH, KW, Cg, groups, Ho, Wo, sh, sw, relu, a_unsigned;
  int K;                      // KH * KW * Cg
  int BN;                     // padded output channels per group handled by one CTA (16..128, multiple of 16)
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;          // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], 8-bit integer operands, int32 accumulators
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b),
      "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed (implies fence::before)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor of a K-major operand tile whose rows are 128 B apart and 128 B-swizzled
// (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14), stride byte offset (8 rows x 128 B = 1024) >> 4 in
// bits [32,46), descriptor version 1 in bits [46,48), layout type SWIZZLE_128B = 2 in bits [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 at [4,6), a/b format (1 = signed 8 bit,
// 0 = unsigned) at [7,10) / [10,13), both operands K-major, N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc_i8(int n, int a_unsigned) {
  return (2u << 4) | ((a_unsigned ? 0u : 1u) << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kMmaM >> 4) << 24);
}

// One CTA: 128 output pixels x BN output channels of one group.  grid = (ceil(M / 128), ceil(Cout_g / BN), groups).
__global__ void __launch_bounds__(kMmaThreads, 1) qconv_igemm_kernel(const QConvArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024 B alignment is what the 128 B swizzle atom (8 rows x 128 B) needs
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* smem_a = smem;                                          // [stages][128 rows][128 B]
  unsigned char* smem_b = smem + kMmaStages * kMmaM * kMmaK;             // [stages][128 rows][128 B] (BN rows used)
  __shared__ uint64_t mma_done[kMmaStages];      // stage consumed by the tensor core -> may be refilled
  __shared__ uint64_t acc_ready;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = a.N * a.Ho * a.Wo;
  const int m0 = blockIdx.x * kMmaM;
  const int g = blockIdx.z;
  const int cout_g = a.Cout / a.groups;
  const int co0 = blockIdx.y * a.BN;                     // within the group
  const int nkb = (a.K + kMmaK - 1) / kMmaK;

  if (tid == 0) {
    for (int s = 0; s < kMmaStages; ++s) mbar_init(&mma_done[s], 1);
    mbar_init(&acc_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<128>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_slot;

  // ---- loader: each thread owns 4 (row, 16 B chunk) slots of A and of B per k-block -----------------------
  // slot s = tid + 256 * i: row = s >> 3 (0..127), chunk = s & 7; stored at row * 128 + ((chunk ^ (row & 7)) << 4)
  int a_row[4], a_chunk[4];
  const signed char* a_base[4];        // &xq[n, oh*sh, ow*sw, g*Cg] of the row's output pixel, or NULL beyond M
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int s = tid + kMmaThreads * i;
    a_row[i] = s >> 3;
    a_chunk[i] = s & 7;
    const int m = m0 + a_row[i];
    if (m < M) {
      const int n = m / (a.Ho * a.Wo), r = m % (a.Ho * a.Wo);
      const int oh = r / a.Wo, ow = r % a.Wo;
      a_base[i] = a.xq + (((int64_t)n * a.Hp + (int64_t)oh * a.sh) * a.Wp + (int64_t)ow * a.sw) * a.C + (int64_t)g * a.Cg;
    } else {
      a_base[i] = nullptr;
    }
  }
  auto load_stage = [&](int kb, int stage) {
    const uint32_t sa = smem_u32(smem_a + stage * kMmaM * kMmaK), sb = smem_u32(smem_b + stage * kMmaM * kMmaK);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = a_row[i], chunk = a_chunk[i];
      const int k = kb * kMmaK + chunk * 16;                 // first of 16 consecutive k = (kh, kw, ci..ci+15)
      const uint32_t off = (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
      // A: gather from the padded NHWC codes
      const int khw = k / a.Cg, ci = k % a.Cg;
      const int kh = khw / a.KW, kw = khw % a.KW;
      const bool va = a_base[i] != nullptr && k < a.K;
      const signed char* src_a = va ? a_base[i] + ((int64_t)kh * a.Wp + kw) * a.C + ci : a.xq;
      cp_async16(sa + off, src_a, va);
      // B: weight rows are K-major already
      const int co = co0 + row;
      const bool vb = row < a.BN && co < cout_g && k < a.K;
      const signed char* src_b = vb ? a.wq + ((int64_t)g * cout_g + co) * a.K + k : a.wq;
      if (row < a.BN) cp_async16(sb + off, src_b, vb);
    }
    cp_async_commit();
  };

  const uint32_t idesc = make_idesc_i8(a.BN, a.a_unsigned);
  // prologue
  for (int s = 0; s < kMmaStages - 1; ++s) {
    if (s < nkb) load_stage(s, s);
    else cp_async_commit();
  }
  for (int kb = 0; kb < nkb; ++kb) {
    const int stage = kb % kMmaStages;
    cp_async_wait<kMmaStages - 2>();         // this thread's copies of k-block kb have landed
    fence_proxy_async();                     // ... and are visible to the tensor core's (async proxy) reads
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sa = smem_u32(smem_a + stage * kMmaM * kMmaK), sb = smem_u32(smem_b + stage * kMmaM * kMmaK);
#pragma unroll
      for (int k = 0; k < kMmaK / 32; ++k) {     // UMMA K = 32 int8 = 32 B: advance the start address inside the swizzle row
        umma_i8(tmem_acc, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc, (kb | k) != 0);
      }
      umma_commit(&mma_done[stage]);             // arrives when the MMAs above have finished reading this stage
      if (kb == nkb - 1) umma_commit(&acc_ready);
    }
    // refill the stage consumed by k-block kb - 1 with k-block kb + stages - 1
    const int nxt = kb + kMmaStages - 1;
    if (nxt < nkb) {
      if (kb >= 1) mbar_wait(&mma_done[(kb - 1) % kMmaStages], ((kb - 1) / kMmaStages) & 1);
      load_stage(nxt, nxt % kMmaStages);
    } else {
      cp_async_commit();
    }
  }

  // ---- epilogue: TMEM -> registers -> (+ bias, ReLU, dequantise) -> NCHW float ------------------------------
  mbar_wait(&acc_ready, 0);
  tc_fence_after();
  const float scale = __fmul_rn(__ldg(a.s_in), __ldg(a.s_w));         // nn/quantized_conv.py:158  in_scale * w_scale
  const int row = (warp & 3) * 32 + lane;                             // TMEM lane == tile row; a warp owns its lane quarter
  const int m = m0 + row;
  const int half = warp >> 2;                                         // warps 0-3: columns [0, BN/2), warps 4-7: the rest
  const int ncol = a.BN / 2;
  int64_t out_base = 0;
  if (m < M) {
    const int n = m / (a.Ho * a.Wo), r = m % (a.Ho * a.Wo);
    out_base = ((int64_t)n * a.Cout + (int64_t)g * cout_g) * (a.Ho * a.Wo) + r;
  }
  for (int c0 = half * ncol; c0 < (half + 1) * ncol; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_acc + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
    // BN / 2 may be 8, 24, ...: columns beyond this half belong to the other warps (or are padding)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = c0 + j, co = co0 + c;
      if (m < M && c < (half + 1) * ncol && co < cout_g) {
        int acc = (int)v[j];
        if (a.bias_q != nullptr) acc += __ldg(a.bias_q + g * cout_g + co);
        if (a.relu) acc = max(acc, 0);
        a.out[out_base + (int64_t)co * (a.Ho * a.Wo)] = __fmul_rn((float)acc, scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free<128>(tmem_acc);
}

// ---- fp32 NCHW -> zero-padded NHWC 8-bit codes (pad, clip, divide, round: nn/quantized_conv.py:108-109, 54-61) ----
// One block per (n, padded row): a warp per channel reads that channel's row of W floats (coalesced), quantises with
// the guarded reciprocal (exactly roundf(clip(x) / scale), fq_common.cuh) and writes the byte into a shared-memory tile
// [Wp][C + 4] (the +4 keeps the 32 lanes of a warp on 32 different banks); the tile is then written out as whole
// 32-bit words, coalesced.  Codes of the padding are the code of 0.0 under the same clip, as in the reference, which
// pads first and quantises the padded tensor.
__global__ void __launch_bounds__(kThreads) qconv_pack_input_kernel(const float* __restrict__ x, const float* __restrict__ range2,
                                                                    signed char* __restrict__ xq, float* __restrict__ scale_out,
                                                                    int N, int C, int H, int W, int ph, int pw) {
  extern __shared__ __align__(16) signed char tile[];       // [Wp][C + 4]
  const int Hp = H + 2 * ph, Wp = W + 2 * pw, Cs = C + 4;
  const int n = blockIdx.x / Hp, hp = blockIdx.x % Hp;
  const float lo = __ldg(range2), hi = __ldg(range2 + 1);
  const float scale = (hi == -lo) ? __fdiv_rn(hi, 127.0f) : __fdiv_rn(__fsub_rn(hi, lo), 255.0f);
  const QDiv qd = QDiv::make(scale);
  // (int) -> low 8 bits: int8 codes [-127, 127] and uint8 codes [0, 255] alike
  const signed char pad_code = (signed char)(unsigned char)(int)quant_code(clipf(0.f, lo, hi), scale);
  const int h = hp - ph;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = warp; c < C; c += nw) {
    const float* row = (h >= 0 && h < H) ? x + (((int64_t)n * C + c) * H + h) * W : nullptr;
    for (int w = lane; w < Wp; w += 32) {
      const int ws = w - pw;
      signed char code = pad_code;
      if (row != nullptr && ws >= 0 && ws < W) code = (signed char)(unsigned char)(int)qd.code(clipf(__ldg(row + ws), lo, hi));
      tile[w * Cs + c] = code;
    }
  }
  __syncthreads();
  const int cw = C >> 2;                                     // C % 16 == 0 (host check): whole words
  const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tile);
  uint32_t* dst = reinterpret_cast<uint32_t*>(xq + ((int64_t)n * Hp + hp) * Wp * C);
  for (int i = threadIdx.x; i < Wp * cw; i += blockDim.x) {
    const int w = i / cw, j = i - w * cw;
    dst[i] = t32[w * (cw + 1) + j];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && scale_out != nullptr) scale_out[0] = scale;
}

// weights: fp32 [Cout, Cg, KH, KW] -> int8 codes [Cout, KH, KW, Cg]
__global__ void __launch_bounds__(kThreads) qconv_pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ range2,
                                                                     signed char* __restrict__ wq, float* __restrict__ scale_out,
                                                                     int64_t n, int Cg, int KHW) {
  const float lo = __ldg(range2), hi = __ldg(range2 + 1);
  const float scale = (hi == -lo) ? __fdiv_rn(hi, 127.0f) : __fdiv_rn(__fsub_rn(hi, lo), 255.0f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t co = i / ((int64_t)Cg * KHW);
    const int r = (int)(i % ((int64_t)Cg * KHW));
    const int ci = r / KHW, khw = r % KHW;
    wq[(co * KHW + khw) * Cg + ci] = (signed char)(int)quant_code(clipf(w[i], lo, hi), scale);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && scale_out != nullptr) scale_out[0] = scale;
}

}  // namespace fq

using namespace fq;

extern "C" {

int fq_qconv_pack_input(const DLTensor* x_, const DLTensor* range2_, int pad_h, int pad_w, const DLTensor* xq_,
                        const DLTensor* scale_out_, void* stream) {
  const char* who = "fq_qconv_pack_input";
  View x, rg, xq, so;
  FQ_TRY(view_of(x_, "fq_qconv_pack_input: x", false, &x));
  FQ_TRY(view_of(range2_, "fq_qconv_pack_input: range2", false, &rg));
  FQ_TRY(view_of(xq_, "fq_qconv_pack_input: xq", false, &xq));
  FQ_TRY(view_of(scale_out_, "fq_qconv_pack_input: scale_out", true, &so));
  FQ_REQUIRE(x.is_f32() && x_->ndim == 4 && rg.is_f32() && rg.numel == 2, "%s: x float32 [N, C, H, W], range2 = 2 float32", who);
  FQ_REQUIRE(pad_h >= 0 && pad_w >= 0, "%s: negative padding", who);
  const int64_t N = x_->shape[0], C = x_->shape[1], H = x_->shape[2], W = x_->shape[3];
  const int64_t Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  FQ_REQUIRE(xq.bits == 8 && (xq.code == kDLInt || xq.code == kDLUInt) && xq.numel == N * Hp * Wp * C,
             "%s: xq must be (u)int8 [N, H+2ph, W+2pw, C]", who);
  FQ_REQUIRE(so.null || (so.is_f32() && so.numel >= 1), "%s: scale_out must be float32", who);
  FQ_REQUIRE(C % 4 == 0, "%s: C=%lld must be a multiple of 4", who, (long long)C);
  FQ_REQUIRE(Wp * (C + 4) <= 200 * 1024, "%s: one padded row of %lld x %lld codes does not fit in shared memory", who,
             (long long)Wp, (long long)C);
  FQ_REQUIRE((reinterpret_cast<uintptr_t>(xq.data) & 3u) == 0, "%s: xq must be 4-byte aligned", who);
  if (x.numel == 0) return 0;
  const size_t smem = (size_t)(Wp * (C + 4));
  FQ_CUDA(cudaFuncSetAttribute(qconv_pack_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qconv_pack_input_kernel<<<(unsigned)(N * Hp), kThreads, smem, (cudaStream_t)stream>>>(
      x.as<const float>(), rg.as<const float>(), xq.as<signed char>(), so.null ? nullptr : so.as<float>(), (int)N, (int)C,
      (int)H, (int)W, pad_h, pad_w);
  FQ_LAUNCH_CHECK("qconv_pack_input_kernel");
  return 0;
}

int fq_qconv_pack_weight(const DLTensor* w_, const DLTensor* range2_, const DLTensor* wq_, const DLTensor* scale_out_,
                         void* stream) {
  const char* who = "fq_qconv_pack_weight";
  View w, rg, wq, so;
  FQ_TRY(view_of(w_, "fq_qconv_pack_weight: w", false, &w));
  FQ_TRY(view_of(range2_, "fq_qconv_pack_weight: range2", false, &rg));
  FQ_TRY(view_of(wq_, "fq_qconv_pack_weight: wq", false, &wq));
  FQ_TRY(view_of(scale_out_, "fq_qconv_pack_weight: scale_out", true, &so));
  FQ_REQUIRE(w.is_f32() && w_->ndim == 4 && rg.is_f32() && rg.numel == 2, "%s: w float32 [Cout, Cg, KH, KW], range2 = 2 float32", who);
  FQ_REQUIRE(wq.bits == 8 && wq.code == kDLInt && wq.numel == w.numel, "%s: wq must be int8 with as many elements as w", who);
  FQ_REQUIRE(so.null || (so.is_f32() && so.numel >= 1), "%s: scale_out must be float32", who);
  if (w.numel == 0) return 0;
  const int64_t b = (w.numel + kThreads - 1) / kThreads;
  qconv_pack_weight_kernel<<<(unsigned)(b > 1184 ? 1184 : b), kThreads, 0, (cudaStream_t)stream>>>(
      w.as<const float>(), rg.as<const float>(), wq.as<signed char>(), so.null ? nullptr : so.as<float>(), w.numel,
      (int)w_->shape[1], (int)(w_->shape[2] * w_->shape[3]));
  FQ_LAUNCH_CHECK("qconv_pack_weight_kernel");
  return 0;
}

int fq_qconv_igemm(const DLTensor* xq_, const DLTensor* wq_, const DLTensor* bias_q_, const DLTensor* s_in_,
                   const DLTensor* s_w_, int stride_h, int stride_w, int groups, int relu, const DLTensor* out_,
                   void* stream) {
  const char* who = "fq_qconv_igemm";
  View xq, wq, bq, si, sw, out;
  FQ_TRY(view_of(xq_, "fq_qconv_igemm: xq", false, &xq));
  FQ_TRY(view_of(wq_, "fq_qconv_igemm: wq", false, &wq));
  FQ_TRY(view_of(bias_q_, "fq_qconv_igemm: bias_q", true, &bq));
  FQ_TRY(view_of(s_in_, "fq_qconv_igemm: s_in", false, &si));
  FQ_TRY(view_of(s_w_, "fq_qconv_igemm: s_w", false, &sw));
  FQ_TRY(view_of(out_, "fq_qconv_igemm: out", false, &out));
  FQ_REQUIRE(xq_->ndim == 4 && xq.bits == 8 && (xq.code == kDLInt || xq.code == kDLUInt),
             "%s: xq must be (u)int8 [N, Hp, Wp, C] (fq_qconv_pack_input)", who);
  FQ_REQUIRE(wq_->ndim == 4 && wq.bits == 8 && wq.code == kDLInt, "%s: wq must be int8 [Cout, KH, KW, Cg] (fq_qconv_pack_weight)", who);
  FQ_REQUIRE(si.is_f32() && sw.is_f32() && si.numel >= 1 && sw.numel >= 1 && out.is_f32() && out_->ndim == 4,
             "%s: scales float32, out float32 [N, Cout, Ho, Wo]", who);
  QConvArgs a = {};
  a.N = (int)xq_->shape[0];
  a.Hp = (int)xq_->shape[1];
  a.Wp = (int)xq_->shape[2];
  a.C = (int)xq_->shape[3];
  a.Cout = (int)wq_->shape[0];
  a.KH = (int)wq_->shape[1];
  a.KW = (int)wq_->shape[2];
  a.Cg = (int)wq_->shape[3];
  a.groups = groups;
  a.sh = stride_h;
  a.sw = stride_w;
  FQ_REQUIRE(groups >= 1 && a.C == a.Cg * groups && a.Cout % groups == 0, "%s: C=%d, Cg=%d, Cout=%d do not match groups=%d",
             who, a.C, a.Cg, a.Cout, groups);
  FQ_REQUIRE(a.Cg % 16 == 0, "%s: input channels per group (%d) must be a multiple of 16 (the loader moves 16 B runs of "
             "channels); use the framework convolution otherwise", who, a.Cg);
  FQ_REQUIRE(stride_h >= 1 && stride_w >= 1 && a.Hp >= a.KH && a.Wp >= a.KW, "%s: bad stride or kernel larger than the input", who);
  a.Ho = (a.Hp - a.KH) / stride_h + 1;
  a.Wo = (a.Wp - a.KW) / stride_w + 1;
  FQ_REQUIRE(out_->shape[0] == a.N && out_->shape[1] == a.Cout && out_->shape[2] == a.Ho && out_->shape[3] == a.Wo,
             "%s: out must be [%d, %d, %d, %d]", who, a.N, a.Cout, a.Ho, a.Wo);
  FQ_REQUIRE(bq.null || (bq.code == kDLInt && bq.bits == 32 && bq.numel == a.Cout), "%s: bias_q must be int32 [Cout]", who);
  FQ_REQUIRE((int64_t)a.N * a.Ho * a.Wo < (1LL << 31) && xq.numel < (1LL << 40), "%s: problem too large", who);
  FQ_REQUIRE((reinterpret_cast<uintptr_t>(xq.data) & 15u) == 0 && (reinterpret_cast<uintptr_t>(wq.data) & 15u) == 0,
             "%s: xq and wq must be 16-byte aligned", who);
  a.K = a.KH * a.KW * a.Cg;
  a.xq = xq.as<const signed char>();
  a.wq = wq.as<const signed char>();
  a.bias_q = bq.null ? nullptr : bq.as<const int>();
  a.s_in = si.as<const float>();
  a.s_w = sw.as<const float>();
  a.out = out.as<float>();
  a.relu = relu;
  a.a_unsigned = xq.code == kDLUInt;
  const int cout_g = a.Cout / groups;
  int bn = (cout_g + 15) / 16 * 16;
  if (bn > 128) bn = 128;
  if (bn < 32) bn = 32;              // the two epilogue halves read 16 columns at a time
  a.BN = bn;
  const int64_t M = (int64_t)a.N * a.Ho * a.Wo;
  if (M == 0) return 0;
  const size_t smem = (size_t)2 * kMmaStages * kMmaM * kMmaK + 1024;
  FQ_CUDA(cudaFuncSetAttribute(qconv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((M + kMmaM - 1) / kMmaM), (unsigned)((cout_g + bn - 1) / bn), (unsigned)groups);
  qconv_igemm_kernel<<<grid, kMmaThreads, smem, (cudaStream_t)stream>>>(a);
  FQ_LAUNCH_CHECK("qconv_igemm_kernel");
  return 0;
}

}  // extern "C"
