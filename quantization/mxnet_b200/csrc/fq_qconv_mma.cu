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
#include <cuda.h>      // CUtensorMap (types only: the encoder is resolved through cudaGetDriverEntryPoint)

#include "fq_fused.cuh"

namespace fq {

constexpr int kMmaM = 128;            // output pixels per CTA
constexpr int kMmaK = 128;            // int8 elements per k-block = one 128 B swizzle row

struct QConvArgs {
  const signed char* xq;      // [N, Hp, Wp, C] codes, spatially padded
  const signed char* wq;      // [Cout, KH, KW, Cg] codes
  const int* bias_q;          // [Cout] int32 codes, or NULL
  const float* bias_f;        // [Cout] float bias to be quantised with b_scale = s_in * s_w (:122-127), or NULL
  const float* s_in;          // device scalars
  const float* s_w;
  float* out;                 // [N, Cout, Ho, Wo]
  int N, C, Hp, Wp, Cout, KH, KW, Cg, groups, Ho, Wo, sh, sw, relu, a_unsigned;
  int K;                      // KH * KW * Cg
  int BN;                     // padded output channels per group handled by one CTA (16..256, multiple of 16)
  int tma_a;                  // 1: A arrives by TMA in im2col mode (Cg % 64 == 0), 0: gathered with cp.async
  int kb_bytes;               // bytes of K per k-block = swizzle span: 128, or 64 when Cg is an odd multiple of 64 (TMA only)
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef FQ_MBAR_TEST_WAIT
  uint32_t done = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
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
// the same for rows 64 B apart, 64 B-swizzled: stride byte offset 8 x 64 B = 512, layout type SWIZZLE_64B = 4
__device__ __forceinline__ uint64_t make_smem_desc64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 at [4,6), a/b format (1 = signed 8 bit,
// 0 = unsigned) at [7,10) / [10,13), both operands K-major, N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc_i8(int n, int a_unsigned) {
  return (2u << 4) | ((a_unsigned ? 0u : 1u) << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kMmaM >> 4) << 24);
}

// ---- the implicit GEMM: persistent, warp-specialised --------------------------------------------------------
// One CTA per SM walks over tiles of 128 output pixels x BN output channels (BN <= 256: a whole 256-channel layer
// reads its A tile once).  Three roles, connected by mbarriers only (no __syncthreads in the steady state):
//   warps 4-11 PRODUCERS  k-block `it` of the CTA's stream goes to stage it % 4 and is loaded by ONE PAIR of warps
//              (warps 4+2s, 5+2s for stage s; 64 tile rows each): wait for empty[s], gather A with cp.async -- a
//              16-byte run of input channels per (pixel, kh, kw), straight from the padded NHWC codes into the
//              128-byte-swizzled K-major layout -- then cp.async.wait_group 0 -> fence.proxy.async -> one arrival per
//              warp on full[s].  A pair has four k-blocks of time for this chain (issue ~100 instructions, the L2
//              round trip, the proxy fence), and the four pairs keep four k-blocks in flight.  Lane 0 of the pair's
//              first warp also launches B: BN weight rows x 128 bytes of K as ONE bulk tensor copy (TMA,
//              128 B swizzle), completing on the same barrier.
//   warp  12   MMA        one thread waits for full[s], issues 4 x tcgen05.mma.kind::i8 (M=128, N=BN, K=32) into
//              one of TWO accumulator buffers in tensor memory, and commits to empty[s] (the pair may refill it)
//              and, after a tile's last k-block, to acc_full[buf];
//   warps 0-3  EPILOGUE   wait for acc_full[buf], read the accumulators with tcgen05.ld (warp w owns TMEM lanes
//              32w..32w+31 = tile rows), release the buffer (acc_empty[buf]) and then add the int32 bias, apply
//              ReLU, dequantise and store NCHW floats -- while the tensor core is already busy with the next tile.
// Measured on the way here (3x3, 256 -> 256 channels, 56 x 56, batch 32; profiles/r2_qconv_*): one CTA per tile with
// __syncthreads per k-block 133 us; two CTAs per SM 111 us; this organisation with every producer warp taking part in
// every k-block 99 us (each warp's serial chain -- constant loads, address math, MEMBAR + proxy fence, barrier -- was
// ~900 clocks per k-block against 512 clocks of tensor-core work) and a per-element branchy epilogue; branch-free
// epilogue 70 us.
constexpr int kEpiWarps = 4, kProdWarps = 8;
constexpr int kMmaThreadsV2 = 32 * (kEpiWarps + kProdWarps + 1);
constexpr int kStagesV2 = 4;
constexpr int kMaxBN = 256;
static_assert(kProdWarps == 2 * kStagesV2, "one pair of producer warps per stage");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free_dyn(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// one arrival + `bytes` of expected TMA traffic on a barrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// TMA: one box of a 2-D tensor map -> shared memory (128 B-swizzled as the map says); completion on `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// TMA in im2col mode: 128 consecutive output pixels (walking W, then H, then N inside the map's bounding box, with the
// convolution stride as traversal stride) x 128 channels of ONE filter tap {kw, kh} -> one 128 B-swizzled A stage
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c, int w, int h,
                                                   int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], "
      "{%7, %8};" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

template <int KB>      // bytes of K per k-block: 128, or 64 (TMA path with taps of an odd multiple of 64 channels)
__global__ void __launch_bounds__(kMmaThreadsV2, 1) qconv_igemm_kernel(const QConvArgs a, const __grid_constant__ CUtensorMap tmap_b,
                                                                       const __grid_constant__ CUtensorMap tmap_a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024 B alignment is what the 128 B swizzle atom (8 rows x 128 B) needs
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* smem_a = smem;                                          // [stages][128 rows][128 B]
  unsigned char* smem_b = smem + kStagesV2 * kMmaM * kMmaK;              // [stages][BN rows][128 B]
  __shared__ uint64_t full[kStagesV2], empty[kStagesV2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) int bias_s[2][kMaxBN];                        // this tile's int32 bias codes

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = a.N * a.Ho * a.Wo, HoWo = a.Ho * a.Wo;
  const int cout_g = a.Cout / a.groups;
  const int m_tiles = (M + kMmaM - 1) / kMmaM, n_tiles = (cout_g + a.BN - 1) / a.BN;
  const int tiles = m_tiles * n_tiles * a.groups;
  const int nkb = (a.K + KB - 1) / KB;
  const int my_tiles = blockIdx.x < tiles ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t b_stage_bytes = (uint32_t)a.BN * kMmaK;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)a.BN) tmem_cols <<= 1;               // two accumulator buffers of BN columns

  if (tid == 0) {
    for (int s = 0; s < kStagesV2; ++s) {
      // gathered A: the two gathering warps of the stage + the thread that launches B's TMA; A by TMA: that thread only
      mbar_init(&full[s], a.tma_a ? 1 : 3);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc_dyn(&tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp >= kEpiWarps && warp < kEpiWarps + kProdWarps && a.tma_a) {
    // =========================== PRODUCER, both operands by TMA ===========================
    // Cg % 128 == 0: a k-block is 128 channels of one filter tap, which is exactly what the im2col mode of the tensor
    // map delivers -- no gather, no address arithmetic, no generic->async proxy fence.  ONE thread keeps the ring full.
    if (warp == kEpiWarps && lane == 0) {
      const int cblocks = a.Cg / KB;                     // k-blocks per filter tap
      int it = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int t = (int)blockIdx.x + lt * (int)gridDim.x;
        const int mt = t % m_tiles, rest = t / m_tiles;
        const int nt = rest % n_tiles, g = rest / n_tiles;
        const int b_row = g * cout_g + nt * a.BN;
        const int m0 = mt * kMmaM;
        const int n = m0 / HoWo, r = m0 - n * HoWo;
        const int oh = r / a.Wo, ow = r - oh * a.Wo;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int stage = it % kStagesV2, use = it / kStagesV2;
          if (use > 0) mbar_wait(&empty[stage], (uint32_t)(use - 1) & 1u);
          const int tap = kb / cblocks, c0 = (kb - tap * cblocks) * KB;
          const int kh = tap / a.KW, kw = tap - kh * a.KW;
          // stages keep their 128-byte-row strides; with 64-byte k-blocks a stage is simply half full
          mbar_arrive_expect_tx(&full[stage], (uint32_t)(kMmaM * KB) + (uint32_t)(a.BN * KB));
          tma_load_im2col_4d(smem_u32(smem_a + stage * kMmaM * kMmaK), &tmap_a, &full[stage], g * a.Cg + c0, ow * a.sw,
                             oh * a.sh, n, (uint16_t)kw, (uint16_t)kh);
          tma_load_2d(smem_u32(smem_b) + (uint32_t)stage * b_stage_bytes, &tmap_b, &full[stage], kb * KB, b_row);
        }
      }
    }
  } else if (warp >= kEpiWarps && warp < kEpiWarps + kProdWarps) {
    // =========================== PRODUCERS, A gathered with cp.async ===========================
    const int pw = warp - kEpiWarps;                     // 0..7
    const int stage = pw >> 1, half = pw & 1;            // this warp: k-blocks it = stage (mod 4), tile rows [64 half, +64)
    const int chunk = lane & 7, rsub = lane >> 3;        // copy i of a k-block: row = 64 half + 4 i + rsub, this chunk
    const uint32_t sa0 = smem_u32(smem_a + stage * kMmaM * kMmaK);
    const uint32_t sb0 = smem_u32(smem_b) + (uint32_t)stage * b_stage_bytes;
    const int total = my_tiles * nkb;                    // k-blocks in this CTA's stream
    int cur_tile = -1, kb = 0, b_row = 0;
    const signed char* a_base[16];       // &xq[n, oh*sh, ow*sw, g*Cg] of each row's output pixel, or NULL beyond M
    for (int it = stage; it < total; it += kStagesV2) {
      const int lt = it / nkb;
      kb = it - lt * nkb;
      if (lt != cur_tile) {              // new tile: the 16 pixels of this lane (rows 4 apart: walk (n, oh, ow))
        cur_tile = lt;
        const int t = (int)blockIdx.x + lt * (int)gridDim.x;
        const int mt = t % m_tiles, rest = t / m_tiles;
        const int nt = rest % n_tiles, g = rest / n_tiles;
        b_row = g * cout_g + nt * a.BN;
        int m = mt * kMmaM + 64 * half + rsub;
        int n = m / HoWo, r = m - n * HoWo;
        int oh = r / a.Wo, ow = r - oh * a.Wo;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          a_base[i] = m < M ? a.xq + (((int64_t)n * a.Hp + (int64_t)oh * a.sh) * a.Wp + (int64_t)ow * a.sw) * a.C + (int64_t)g * a.Cg
                            : nullptr;
          m += 4;
          ow += 4;
          while (ow >= a.Wo) {
            ow -= a.Wo;
            if (++oh == a.Ho) {
              oh = 0;
              ++n;
            }
          }
        }
      }
      const int use = it / kStagesV2;                    // how often this stage has been filled before
      if (use > 0) mbar_wait(&empty[stage], (uint32_t)(use - 1) & 1u);
      if (half == 0 && lane == 0) {
        // B: BN weight rows x 128 B of K in ONE bulk tensor copy (rows beyond Cout and bytes beyond K arrive as
        // zeros; rows of the next group, if the tile overhangs its group, feed columns the epilogue drops)
        mbar_arrive_expect_tx(&full[stage], b_stage_bytes);
        tma_load_2d(sb0, &tmap_b, &full[stage], kb * kMmaK, b_row);
      }
      // A: this lane's chunk is k = (kh, kw, ci .. ci + 15)
      const int k = kb * kMmaK + chunk * 16;
      const bool vk = k < a.K;
      const int khw = k / a.Cg, ci = k - khw * a.Cg;
      const int kh = khw / a.KW, kw = khw - kh * a.KW;
      const int64_t a_off = ((int64_t)kh * a.Wp + kw) * a.C + ci;
      const uint32_t dst0 = sa0 + (uint32_t)(64 * half + rsub) * 128u;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int row = 64 * half + 4 * i + rsub;        // row & 7 = (4 i + rsub) & 7
        const bool va = vk && a_base[i] != nullptr;
        cp_async16(dst0 + (uint32_t)i * 512u + (uint32_t)((chunk ^ (row & 7)) << 4), va ? a_base[i] + a_off : a.xq, va);
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();               // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();                      // every lane's copies are in and fenced: ONE arrival per warp
      if (lane == 0) mbar_arrive(&full[stage]);
    }
  } else if (warp == kEpiWarps + kProdWarps) {
    // =========================== MMA ISSUER ===========================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_i8(a.BN, a.a_unsigned);
      int it = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int buf = lt & 1;
        if (lt >= 2) {                         // the epilogue has drained this accumulator buffer
          mbar_wait(&acc_empty[buf], (uint32_t)((lt >> 1) - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t tmem_acc = tmem_base + (uint32_t)buf * (tmem_cols >> 1);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int stage = it % kStagesV2;
          mbar_wait(&full[stage], (uint32_t)(it / kStagesV2) & 1u);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * kMmaM * kMmaK);
          const uint32_t sb = smem_u32(smem_b) + (uint32_t)stage * b_stage_bytes;
#pragma unroll
          for (int k = 0; k < KB / 32; ++k) {    // UMMA K = 32 int8 = 32 B: advance the start address inside the swizzle row
            if (KB == kMmaK) umma_i8(tmem_acc, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc, (kb | k) != 0);
            else umma_i8(tmem_acc, make_smem_desc64(sa + k * 32), make_smem_desc64(sb + k * 32), idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);            // arrives when the MMAs above have finished reading this stage
          if (kb == nkb - 1) umma_commit(&acc_full[buf]);
        }
      }
    }
  } else {
    // =========================== EPILOGUE ===========================
    const float scale = __fmul_rn(__ldg(a.s_in), __ldg(a.s_w));         // nn/quantized_conv.py:158  in_scale * w_scale
    const float b_max = __fmul_rn(scale, 2147483648.0f);                // :123  b_scale * 2^31
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int t = (int)blockIdx.x + lt * (int)gridDim.x;
      const int mt = t % m_tiles, rest = t / m_tiles;
      const int nt = rest % n_tiles, g = rest / n_tiles;
      const int m0 = mt * kMmaM, co0 = nt * a.BN;
      const int buf = lt & 1;
      // this tile's bias as int32 codes: given, or quantised here from the float bias (:122-127)
      for (int c = tid; c < a.BN; c += 32 * kEpiWarps) {
        const int co = co0 + c;
        int bq = 0;
        if (co < cout_g) {
          if (a.bias_q != nullptr) bq = __ldg(a.bias_q + g * cout_g + co);
          else if (a.bias_f != nullptr) bq = (int)quant_code(clipf(__ldg(a.bias_f + g * cout_g + co), -b_max, b_max), scale);
        }
        bias_s[buf][c] = bq;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");   // epilogue warps only
      const int row = warp * 32 + lane;                                  // TMEM lane == tile row
      const int m = m0 + row;
      int64_t out_base = 0;
      if (m < M) {
        const int n = m / HoWo, r = m - n * HoWo;
        out_base = ((int64_t)n * a.Cout + (int64_t)g * cout_g + co0) * HoWo + r;
      }
      mbar_wait(&acc_full[buf], (uint32_t)(lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * (tmem_cols >> 1);
      const int ncols = min(a.BN, cout_g - co0);
      float* out_row = a.out + out_base;                              // + c * HoWo per output channel
      const uint32_t cstride = (uint32_t)HoWo * 4u;                   // bytes between channels (HoWo < 2^29: host check)
      for (int c0 = 0; c0 < a.BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld16_nowait(tmem_acc + (uint32_t)c0, v);
        if (c0 + 16 < a.BN) tmem_ld16_nowait(tmem_acc + (uint32_t)(c0 + 16), v + 16);
        tmem_ld_wait();
        if (c0 + 32 >= a.BN) {                 // last read of this buffer: hand it back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (m >= M) continue;
        char* o = reinterpret_cast<char*>(out_row) + (uint64_t)(uint32_t)c0 * cstride;
        if (c0 + 32 <= ncols) {
          // whole chunk inside the layer: straight-line code, the 32 bias codes as 8 broadcast 128-bit loads
          int b[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int4 t4 = *reinterpret_cast<const int4*>(&bias_s[buf][c0 + 4 * q]);
            b[4 * q] = t4.x; b[4 * q + 1] = t4.y; b[4 * q + 2] = t4.z; b[4 * q + 3] = t4.w;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            int acc = (int)v[j] + b[j];
            if (a.relu) acc = max(acc, 0);
            *reinterpret_cast<float*>(o + (uint64_t)(uint32_t)j * cstride) = __fmul_rn((float)acc, scale);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (c0 + j < ncols) {
              int acc = (int)v[j] + bias_s[buf][c0 + j];
              if (a.relu) acc = max(acc, 0);
              *reinterpret_cast<float*>(o + (uint64_t)(uint32_t)j * cstride) = __fmul_rn((float)acc, scale);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free_dyn(tmem_base, tmem_cols);
}

// ---- the same GEMM on SM PAIRS: tcgen05 cta_group::2 --------------------------------------------------------------
// With one SM per tile every MMA reads 12 KB of operands from shared memory for 128 clocks of math while TMA writes
// 48 KB per k-block into it: ~190 B/clock against the 128 B/clock a shared memory delivers, so the tensor pipe stalls
// (53 % busy, measured).  A 2-CTA cluster (the two SMs of a TPC) computes a 256-pixel x BN tile instead: each CTA
// holds ITS 128 rows of A and HALF of B's rows, the leader CTA's single MMA thread issues M = 256 instructions that
// read both shared memories, and each tensor core accumulates its own 128 rows in its own tensor memory.  Per SM and
// k-block that is 32 KB written + 32 KB read -- and 32 KB stages make room for a 6-deep ring.
//   producer (1 thread per CTA): waits for ITS empty[s], launches its im2col TMA for A and its tiled TMA for half of B;
//     both complete on the LEADER's full[s] (mbarrier address with the peer bit cleared), for which the leader's
//     producer has announced the bytes of both CTAs;
//   MMA (leader only): waits for full[s], issues 4 x tcgen05.mma.cta_group::2, commits with a 2-CTA multicast to
//     empty[s] -- and to acc_full[buf] after a tile's last k-block -- in BOTH CTAs;
//   epilogue (4 warps per CTA): as above on the CTA's own 128 rows; the buffer is handed back by arriving on the
//     LEADER's acc_empty[buf] (8 warp arrivals).
// Used when both operands come by TMA (Cin/groups % 128 == 0) and BN is a multiple of 32.
constexpr int kStages2 = 6;
constexpr int kThreads2 = 32 * (kEpiWarps + 2);
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;        // shared::cluster address of the same word in the pair's even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on this barrier's twin in the even CTA
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cluster.b64 _, [%0];\n\t}" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          dst),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c, int w, int h,
                                                    int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
      "[%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
    qconv_igemm2_kernel(const QConvArgs a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* smem_a = smem;                                          // [stages][128 rows][128 B]
  unsigned char* smem_b = smem + kStages2 * kMmaM * kMmaK;               // [stages][BN / 2 rows][128 B]
  __shared__ uint64_t full[kStages2], empty[kStages2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) int bias_s[2][kMaxBN];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = (int)blockIdx.x >> 1, n_clusters = (int)gridDim.x >> 1;
  const int M = a.N * a.Ho * a.Wo, HoWo = a.Ho * a.Wo;
  const int cout_g = a.Cout / a.groups;
  const int m_tiles = (M + 2 * kMmaM - 1) / (2 * kMmaM), n_tiles = (cout_g + a.BN - 1) / a.BN;
  const int tiles = m_tiles * n_tiles * a.groups;
  const int nkb = (a.K + kMmaK - 1) / kMmaK;
  const int my_tiles = cluster_id < tiles ? (tiles - 1 - cluster_id) / n_clusters + 1 : 0;
  const uint32_t b_stage_bytes = (uint32_t)(a.BN >> 1) * kMmaK;          // this CTA's half of B
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)a.BN) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full[s], 1);            // the leader's producer (which announces both CTAs' bytes)
      mbar_init(&empty[s], 1);           // one multicast commit per use
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 2 * kEpiWarps);   // the epilogue warps of both CTAs (leader's copy is the one in use)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                    // both CTAs' barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == kEpiWarps) {
    // =========================== PRODUCER (one thread per CTA) ===========================
    if (lane == 0) {
      const int cblocks = a.Cg / kMmaK;
      const uint32_t stage_tx = 2u * ((uint32_t)(kMmaM * kMmaK) + b_stage_bytes);
      int it = 0;
#ifdef FQ_QCONV_PROFILE
      long long prof_wait_empty = 0;
      const long long prof_t0 = clock64();
#endif
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int t = cluster_id + lt * n_clusters;
        const int mt = t % m_tiles, rest = t / m_tiles;
        const int nt = rest % n_tiles, g = rest / n_tiles;
        const int b_row = g * cout_g + nt * a.BN + rank * (a.BN >> 1);
        int m0 = mt * 2 * kMmaM + rank * kMmaM;
        if (m0 >= M) m0 = 0;             // a half tile entirely beyond M: load something valid, the epilogue drops it
        const int n = m0 / HoWo, r = m0 - n * HoWo;
        const int oh = r / a.Wo, ow = r - oh * a.Wo;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int stage = it % kStages2, use = it / kStages2;
#ifdef FQ_QCONV_PROFILE
          const long long tp0 = clock64();
#endif
          if (use > 0) mbar_wait(&empty[stage], (uint32_t)(use - 1) & 1u);
#ifdef FQ_QCONV_PROFILE
          prof_wait_empty += clock64() - tp0;
#endif
          const int tap = kb / cblocks, c0 = (kb - tap * cblocks) * kMmaK;
          const int kh = tap / a.KW, kw = tap - kh * a.KW;
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], stage_tx);
          tma2_load_im2col_4d(smem_u32(smem_a + stage * kMmaM * kMmaK), &tmap_a, &full[stage], g * a.Cg + c0, ow * a.sw,
                              oh * a.sh, n, (uint16_t)kw, (uint16_t)kh);
          tma2_load_2d(smem_u32(smem_b) + (uint32_t)stage * b_stage_bytes, &tmap_b, &full[stage], kb * kMmaK, b_row);
        }
      }
#ifdef FQ_QCONV_PROFILE
      if (blockIdx.x < 2) printf("cta %d producer: total %lld, waiting for empty %lld, k-blocks %d\n", (int)blockIdx.x, clock64() - prof_t0, prof_wait_empty, it);
#endif
    }
  } else if (warp == kEpiWarps + 1) {
    // =========================== MMA ISSUER (leader CTA only) ===========================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = (2u << 4) | ((a.a_unsigned ? 0u : 1u) << 7) | (1u << 10) | ((uint32_t)(a.BN >> 3) << 17) |
                             ((uint32_t)((2 * kMmaM) >> 4) << 24);
      int it = 0;
#ifdef FQ_QCONV_PROFILE
      long long prof_full = 0, prof_acc = 0;
      const long long prof_t0 = clock64();
#endif
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int buf = lt & 1;
        if (lt >= 2) {
#ifdef FQ_QCONV_PROFILE
          const long long tp0 = clock64();
#endif
          mbar_wait(&acc_empty[buf], (uint32_t)((lt >> 1) - 1) & 1u);
#ifdef FQ_QCONV_PROFILE
          prof_acc += clock64() - tp0;
#endif
          tc_fence_after();
        }
        const uint32_t tmem_acc = tmem_base + (uint32_t)buf * (tmem_cols >> 1);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int stage = it % kStages2;
#ifdef FQ_QCONV_PROFILE
          const long long tp0 = clock64();
#endif
          mbar_wait(&full[stage], (uint32_t)(it / kStages2) & 1u);
#ifdef FQ_QCONV_PROFILE
          prof_full += clock64() - tp0;
#endif
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * kMmaM * kMmaK);
          const uint32_t sb = smem_u32(smem_b) + (uint32_t)stage * b_stage_bytes;
#pragma unroll
          for (int k = 0; k < kMmaK / 32; ++k)
            umma2_i8(tmem_acc, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc, (kb | k) != 0);
          umma2_commit_both(&empty[stage]);
          if (kb == nkb - 1) umma2_commit_both(&acc_full[buf]);
        }
      }
#ifdef FQ_QCONV_PROFILE
      if (blockIdx.x == 0) printf("cta 0 mma: total %lld, waiting for full %lld, for acc_empty %lld, tiles %d\n", clock64() - prof_t0, prof_full, prof_acc, my_tiles);
#endif
    }
  } else {
    // =========================== EPILOGUE (this CTA's 128 rows) ===========================
    const float scale = __fmul_rn(__ldg(a.s_in), __ldg(a.s_w));
    const float b_max = __fmul_rn(scale, 2147483648.0f);
#ifdef FQ_QCONV_PROFILE
    long long prof_wait = 0;
    const long long prof_t0 = clock64();
#endif
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int t = cluster_id + lt * n_clusters;
      const int mt = t % m_tiles, rest = t / m_tiles;
      const int nt = rest % n_tiles, g = rest / n_tiles;
      const int m0 = mt * 2 * kMmaM + rank * kMmaM, co0 = nt * a.BN;
      const int buf = lt & 1;
      for (int c = tid; c < a.BN; c += 32 * kEpiWarps) {
        const int co = co0 + c;
        int bq = 0;
        if (co < cout_g) {
          if (a.bias_q != nullptr) bq = __ldg(a.bias_q + g * cout_g + co);
          else if (a.bias_f != nullptr) bq = (int)quant_code(clipf(__ldg(a.bias_f + g * cout_g + co), -b_max, b_max), scale);
        }
        bias_s[buf][c] = bq;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      const int m = m0 + warp * 32 + lane;
      int64_t out_base = 0;
      if (m < M) {
        const int n = m / HoWo, r = m - n * HoWo;
        out_base = ((int64_t)n * a.Cout + (int64_t)g * cout_g + co0) * HoWo + r;
      }
#ifdef FQ_QCONV_PROFILE
      const long long tp0 = clock64();
#endif
      mbar_wait(&acc_full[buf], (uint32_t)(lt >> 1) & 1u);
#ifdef FQ_QCONV_PROFILE
      prof_wait += clock64() - tp0;
#endif
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)buf * (tmem_cols >> 1);
      const int ncols = min(a.BN, cout_g - co0);
      float* out_row = a.out + out_base;
      const uint32_t cstride = (uint32_t)HoWo * 4u;
      for (int c0 = 0; c0 < a.BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld16_nowait(tmem_acc + (uint32_t)c0, v);
        tmem_ld16_nowait(tmem_acc + (uint32_t)(c0 + 16), v + 16);      // BN % 32 == 0 on this path
        tmem_ld_wait();
        if (c0 + 32 >= a.BN) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&acc_empty[buf]);
        }
        if (m >= M) continue;
        char* o = reinterpret_cast<char*>(out_row) + (uint64_t)(uint32_t)c0 * cstride;
        if (c0 + 32 <= ncols) {
          int b[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int4 t4 = *reinterpret_cast<const int4*>(&bias_s[buf][c0 + 4 * q]);
            b[4 * q] = t4.x; b[4 * q + 1] = t4.y; b[4 * q + 2] = t4.z; b[4 * q + 3] = t4.w;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            int acc = (int)v[j] + b[j];
            if (a.relu) acc = max(acc, 0);
            *reinterpret_cast<float*>(o + (uint64_t)(uint32_t)j * cstride) = __fmul_rn((float)acc, scale);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (c0 + j < ncols) {
              int acc = (int)v[j] + bias_s[buf][c0 + j];
              if (a.relu) acc = max(acc, 0);
              *reinterpret_cast<float*>(o + (uint64_t)(uint32_t)j * cstride) = __fmul_rn((float)acc, scale);
            }
          }
        }
      }
    }
#ifdef FQ_QCONV_PROFILE
    if (blockIdx.x == 0 && tid == 0) printf("cta 0 epilogue warp 0: total %lld, waiting for acc_full %lld\n", clock64() - prof_t0, prof_wait);
#endif
  }
  tc_fence_before();
  cluster_sync_all();                    // nobody signals the peer or reads tensor memory after this point
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

// ---- fp32 NCHW -> zero-padded NHWC 8-bit codes (pad, clip, divide, round: nn/quantized_conv.py:108-109, 54-61) ----
// HBM-bound: 4 B read + 1 B written per element.  One block per (n, padded row).  A work item is (16 channels, one
// pixel): its thread issues the 16 scalar loads up front (lanes run along W, so each load instruction of a warp is one
// coalesced 128 B line of one channel row), quantises with the guarded reciprocal (exactly roundf(clip(x) / scale),
// fq_common.cuh), packs the 16 codes into one 128-bit shared-memory store into a tile [W][C + 16] (row stride 4 banks
// mod 32: the 8 lanes of a store phase cover the 32 banks exactly once), and the tile leaves as whole 16-byte vectors,
// coalesced.  Codes of the padding are the code of 0.0 under the same clip, as in the reference, which pads first and
// quantises the padded tensor; padding rows and columns never touch shared memory.
__device__ __forceinline__ uint32_t pack4_codes(float c0, float c1, float c2, float c3) {
  // (int) -> low 8 bits: int8 codes [-127, 127] and uint8 codes [0, 255] alike; three byte permutes
  const uint32_t lo = __byte_perm((uint32_t)(int)c0, (uint32_t)(int)c1, 0x0040);
  const uint32_t hi = __byte_perm((uint32_t)(int)c2, (uint32_t)(int)c3, 0x0040);
  return __byte_perm(lo, hi, 0x5410);
}
// the by-the-book code of one element, out of line: taken by a few elements per million (QDiv's guard band)
__device__ __noinline__ float slow_code(float v, float d) { return quant_code(v, d); }

__global__ void __launch_bounds__(kThreads, 4) qconv_pack_input_kernel(const float* __restrict__ x, const float* __restrict__ range2,
                                                                    signed char* __restrict__ xq, float* __restrict__ scale_out,
                                                                    int N, int C, int H, int W, int ph, int pw) {
  extern __shared__ __align__(16) unsigned char tile[];       // [W][C + 16]
  const int Hp = H + 2 * ph, Wp = W + 2 * pw, Cs = C + 16, c16n = C >> 4;
  // rows are taken from the END of the tensor backwards: the range pass that precedes this kernel (max |x|) read x
  // front to back, so its tail is what still sits in L2
  const int bid = (int)(gridDim.x - 1 - blockIdx.x);
  const int n = bid / Hp, hp = bid % Hp;
  const float lo = __ldg(range2), hi = __ldg(range2 + 1);
  const float scale = (hi == -lo) ? __fdiv_rn(hi, 127.0f) : __fdiv_rn(__fsub_rn(hi, lo), 255.0f);
  const QDiv qd = QDiv::make(scale);
  const uint32_t pad1 = (uint32_t)(int)quant_code(clipf(0.f, lo, hi), scale) & 0xffu;
  const uint32_t pad4 = pad1 * 0x01010101u;
  const int h = hp - ph;
  const bool interior = h >= 0 && h < H;
  if (interior) {
    const float* plane = x + ((int64_t)n * C * H + h) * W;      // channel c of this row: plane + c * H * W
    const uint32_t cstride = (uint32_t)(H * W);                 // H * W < 2^27 (host check): 32-bit offsets
    const uint64_t magic_w = 0xffffffffull / (uint32_t)W + 1ull;   // exact quotients for item < 2^16 (host check)
#pragma unroll 1
    for (int item = threadIdx.x; item < c16n * W; item += blockDim.x) {
      const int c16 = (int)(((uint64_t)(uint32_t)item * magic_w) >> 32), w = item - c16 * W;
      const char* src = reinterpret_cast<const char*>(plane + (int64_t)(c16 * 16) * cstride + w);
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)       // base + u32 * u32 in one wide multiply-add per load
        v[j] = __ldg(reinterpret_cast<const float*>(src + (uint64_t)(uint32_t)j * (uint64_t)(cstride * 4u)));
      float t[16];
      float worst = 0.f;                 // NaN-propagating max of QDiv::near_tie's left-hand side over the 16 elements
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        v[j] = clipf(v[j], lo, hi);
        const float q0 = __fmul_rn(v[j], qd.r);
        t[j] = rintf(q0);
        const float m = __fmaf_rn(fabsf(q0), qd.guard, fabsf(__fsub_rn(q0, t[j])));
        asm("max.NaN.f32 %0, %0, %1;" : "+f"(worst) : "f"(m));
      }
      if (!(worst < 0.5f)) {             // some element sits in the guard band (or is NaN): by the book for those
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (qd.near_tie(__fmul_rn(v[j], qd.r), t[j])) t[j] = slow_code(v[j], qd.d);
      }
      *reinterpret_cast<uint4*>(tile + (size_t)w * Cs + c16 * 16) =
          make_uint4(pack4_codes(t[0], t[1], t[2], t[3]), pack4_codes(t[4], t[5], t[6], t[7]),
                     pack4_codes(t[8], t[9], t[10], t[11]), pack4_codes(t[12], t[13], t[14], t[15]));
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(xq + ((int64_t)n * Hp + hp) * Wp * C);
  const uint4 padv = make_uint4(pad4, pad4, pad4, pad4);
  const uint64_t magic_c = 0xffffffffull / (uint32_t)c16n + 1ull;   // 2^32 when c16n == 1
#pragma unroll 2
  for (int i = threadIdx.x; i < Wp * c16n; i += blockDim.x) {
    const int wp_ = (int)(((uint64_t)(uint32_t)i * magic_c) >> 32), j = i - wp_ * c16n;
    const int w = wp_ - pw;
    dst[i] = (interior && w >= 0 && w < W) ? *reinterpret_cast<const uint4*>(tile + (size_t)w * Cs + j * 16) : padv;
  }
  if (bid == 0 && threadIdx.x == 0 && scale_out != nullptr) scale_out[0] = scale;
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

// cuTensorMapEncodeTiled through the runtime's driver entry point query: no link-time dependency on libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeIm2colFn encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  }
  return fn;
}
// FQ_QCONV_TMA_A=0 (environment, read per call) keeps the cp.async gather for A on every shape (A/B measurements, tests)
static bool tma_a_disabled() {
  const char* env = getenv("FQ_QCONV_TMA_A");
  return env != nullptr && env[0] == '0';
}
// FQ_QCONV_2CTA (environment, read per call so that tests can switch it): "0" one SM per tile, "1" SM pairs whenever
// the shape allows, unset: SM pairs when there are at least two waves of 256-pixel tiles
static int two_cta_mode() {
  const char* env = getenv("FQ_QCONV_2CTA");
  return env == nullptr ? -1 : (env[0] == '0' ? 0 : 1);
}

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
  FQ_REQUIRE(C % 16 == 0, "%s: C=%lld must be a multiple of 16 (the tensor-core loader moves 16-byte runs of channels)",
             who, (long long)C);
  FQ_REQUIRE(W * (C + 16) <= 200 * 1024 && Wp * (C / 16) < 65536,
             "%s: one row of %lld x %lld codes does not fit in shared memory", who, (long long)W, (long long)C);
  FQ_REQUIRE((reinterpret_cast<uintptr_t>(xq.data) & 15u) == 0, "%s: xq must be 16-byte aligned", who);
  FQ_REQUIRE(H * W < (1LL << 27), "%s: one channel plane of %lld x %lld is too large", who, (long long)H, (long long)W);
  if (x.numel == 0) return 0;
  const size_t smem = (size_t)(W * (C + 16));
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
  FQ_REQUIRE(bq.null || (((bq.code == kDLInt && bq.bits == 32) || bq.is_f32()) && bq.numel == a.Cout),
             "%s: bias must be int32 codes or float32 [Cout]", who);
  FQ_REQUIRE((int64_t)a.N * a.Ho * a.Wo < (1LL << 31) && (int64_t)a.Ho * a.Wo < (1LL << 29) && xq.numel < (1LL << 40),
             "%s: problem too large", who);
  FQ_REQUIRE((reinterpret_cast<uintptr_t>(xq.data) & 15u) == 0 && (reinterpret_cast<uintptr_t>(wq.data) & 15u) == 0,
             "%s: xq and wq must be 16-byte aligned", who);
  a.K = a.KH * a.KW * a.Cg;
  a.xq = xq.as<const signed char>();
  a.wq = wq.as<const signed char>();
  a.bias_q = (bq.null || bq.is_f32()) ? nullptr : bq.as<const int>();
  a.bias_f = (!bq.null && bq.is_f32()) ? bq.as<const float>() : nullptr;
  a.s_in = si.as<const float>();
  a.s_w = sw.as<const float>();
  a.out = out.as<float>();
  a.relu = relu;
  a.a_unsigned = xq.code == kDLUInt;
  const int cout_g = a.Cout / groups;
  int bn = (cout_g + 15) / 16 * 16;
  if (bn > kMaxBN) bn = kMaxBN;
  a.BN = bn;
  const int64_t M = (int64_t)a.N * a.Ho * a.Wo;
  if (M == 0) return 0;
  const int64_t tiles = ((M + kMmaM - 1) / kMmaM) * ((cout_g + bn - 1) / bn) * groups;
  FQ_REQUIRE(tiles < (1LL << 31), "%s: problem too large", who);
  const size_t smem = (size_t)kStagesV2 * (kMmaM * kMmaK + (size_t)bn * kMmaK) + 1024;

  // A by TMA needs a k-block to be channels of ONE filter tap: 128-byte k-blocks when Cg % 128 == 0, 64-byte
  // k-blocks (64 B swizzle, two MMAs per k-block) when Cg is an odd multiple of 64; otherwise A is gathered
  EncodeIm2colFn encode_im2col = encode_im2col_fn();
  const bool tma_a_ok = a.Cg % 64 == 0 && encode_im2col != nullptr && a.KW <= 128 && a.KH <= 128 && a.sw <= 8 && a.sh <= 8 &&
                        !tma_a_disabled();
  a.kb_bytes = (tma_a_ok && a.Cg % kMmaK != 0) ? 64 : kMmaK;
  const CUtensorMapSwizzle swz = a.kb_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  // B = the [Cout, K] int8 code matrix, fetched by TMA in boxes of bn rows x one k-block of K, swizzled like A
  EncodeTiledFn encode = encode_tiled_fn();
  FQ_REQUIRE(encode != nullptr, "%s: the driver does not export cuTensorMapEncodeTiled", who);
  CUtensorMap tmap_b;
  const cuuint64_t gdim[2] = {(cuuint64_t)a.K, (cuuint64_t)a.Cout};
  const cuuint64_t gstride[1] = {(cuuint64_t)a.K};                        // bytes between rows (K % 16 == 0)
  const cuuint32_t box[2] = {(cuuint32_t)a.kb_bytes, (cuuint32_t)bn};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult er = encode(&tmap_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<signed char*>(a.wq), gdim, gstride, box,
                             estride, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FQ_REQUIRE(er == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed with %d", who, (int)er);
  // A = the padded NHWC codes [N, Hp, Wp, C] read through an im2col tensor map when a k-block is 128 (or 64) channels
  // of one filter tap.  The padding is materialised, so the bounding box of the window origins starts at 0 and
  // ends KW-1 / KH-1 short of the far edges; the convolution stride is the traversal stride.
  CUtensorMap tmap_a = tmap_b;
  a.tma_a = 0;
  if (tma_a_ok) {
    const cuuint64_t adim[4] = {(cuuint64_t)a.C, (cuuint64_t)a.Wp, (cuuint64_t)a.Hp, (cuuint64_t)a.N};
    const cuuint64_t astride[3] = {(cuuint64_t)a.C, (cuuint64_t)a.Wp * a.C, (cuuint64_t)a.Hp * a.Wp * a.C};
    const int lower[2] = {0, 0};
    const int upper[2] = {-(a.KW - 1), -(a.KH - 1)};
    const cuuint32_t astep[4] = {1, (cuuint32_t)a.sw, (cuuint32_t)a.sh, 1};
    const CUresult ar = encode_im2col(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<signed char*>(a.xq), adim, astride,
                                      lower, upper, (cuuint32_t)a.kb_bytes, (cuuint32_t)kMmaM, astep, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (ar == CUDA_SUCCESS) {
      // drivers up to CUDA 13.1 set a descriptor bit for tensors below 128 KB that the im2col mode must not carry
      // (the same fix-up CUTLASS applies in make_im2col_tma_copy_desc)
      int drv = 0;
      if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && xq.numel < 131072)
        reinterpret_cast<uint64_t*>(&tmap_a)[1] &= ~(1ull << 21);
      a.tma_a = 1;
    } else {
      FQ_REQUIRE(a.kb_bytes == kMmaK, "%s: cuTensorMapEncodeIm2col failed with %d", who, (int)ar);
      tmap_a = tmap_b;                                   // not representable: keep the cp.async gather
    }
  } else {
    FQ_REQUIRE(a.kb_bytes == kMmaK, "%s: internal: 64-byte k-blocks without TMA", who);
  }
  const int64_t sms = sm_count();
  const int64_t tiles2 = ((M + 2 * kMmaM - 1) / (2 * kMmaM)) * ((cout_g + bn - 1) / bn) * groups;
  const int mode2 = two_cta_mode();
  // pairs pay a cluster launch and two cluster-wide barriers: worth it from two waves of tiles on (measured)
  if (a.tma_a && a.kb_bytes == kMmaK && bn % 32 == 0 && (mode2 == 1 || (mode2 < 0 && tiles2 >= 2 * (sms / 2)))) {
    // SM pairs (cta_group::2): 256 pixels x bn channels per cluster, each CTA fetches half of B's rows
    CUtensorMap tmap_b2;
    const cuuint32_t box2[2] = {(cuuint32_t)kMmaK, (cuuint32_t)(bn / 2)};
    const CUresult er2 = encode(&tmap_b2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<signed char*>(a.wq), gdim, gstride, box2,
                                estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FQ_REQUIRE(er2 == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed with %d", who, (int)er2);
    const size_t smem2 = (size_t)kStages2 * (kMmaM * kMmaK + (size_t)(bn / 2) * kMmaK) + 1024;
    FQ_CUDA(cudaFuncSetAttribute(qconv_igemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    const int64_t pairs = sms / 2;
    const unsigned grid2 = 2u * (unsigned)(tiles2 < pairs ? tiles2 : pairs);
    qconv_igemm2_kernel<<<grid2, kThreads2, smem2, (cudaStream_t)stream>>>(a, tmap_b2, tmap_a);
    FQ_LAUNCH_CHECK("qconv_igemm2_kernel");
    return 0;
  }
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);            // persistent: one CTA per SM
  auto kern1 = a.kb_bytes == 64 ? qconv_igemm_kernel<64> : qconv_igemm_kernel<kMmaK>;
  FQ_CUDA(cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern1<<<grid, kMmaThreadsV2, smem, (cudaStream_t)stream>>>(a, tmap_b, tmap_a);
  FQ_LAUNCH_CHECK("qconv_igemm_kernel");
  return 0;
}

}  // extern "C"
