// Helpers shared by the range kernels and the fused (single-launch) input / weight paths.
#pragma once
#include "fq_common.cuh"

namespace fq {

// what the last block does once every row maximum is known
struct FinishParams {
  float* out_rows;          // nullable: [rows] maxima
  float* out_mean;          // nullable: Kahan mean of the maxima (current_input_max)
  float* qparams;           // nullable: {d, s, lo, hi} derived from input_max or the mean
  const float* input_max;   // nullable: offline range (convert_conv2d.py:58)
  int bits, is_signed, lo_mode, promotion;
};

// Each block owns a contiguous slice of per_block elements (a multiple of 4 * kThreads * kUnroll so that
// every slice starts 16 B aligned).  Returns the grid size.
inline int slice_grid(int64_t n, int max_blocks, int64_t* per_block) {
  // small tensors are latency-bound: give them more, smaller slices (one float4 per thread)
  const int64_t quantum = (n < (1LL << 20)) ? 4LL * kThreads : 4LL * kThreads * kUnroll;
  int64_t blocks = (n + quantum - 1) / quantum;
  if (blocks < 1) blocks = 1;
  if (blocks > max_blocks) blocks = max_blocks;
  int64_t pb = (n + blocks - 1) / blocks;
  pb = (pb + quantum - 1) / quantum * quantum;
  *per_block = pb;
  return (int)((n + pb - 1) / pb > 0 ? (n + pb - 1) / pb : 1);
}

// Slices for kernels that walk [rows, L]: short rows (< 2048 elements, taken one per warp with scalar loads, so
// no alignment requirement) get whole rows per block and as many blocks as there are rows -- these tensors are
// tiny and latency-bound, and every row costs a dependent chain of loads; long rows use the 16 B aligned slices.
inline int row_grid(int64_t n, int64_t L, int max_blocks, int64_t* per_block) {
  if (L >= 2048 || L <= 0) return slice_grid(n, max_blocks, per_block);
  const int64_t rows = n / L;
  const int64_t k = (rows + max_blocks - 1) / max_blocks;
  *per_block = k * L;
  return (int)((rows + k - 1) / k);
}

// Grid for the tile-interleaved elementwise kernels: one block per tile up to max_blocks.
inline int tile_grid(int64_t n, int max_blocks) {
  int64_t tiles = (n + kTileElems - 1) / kTileElems;
  if (tiles < 1) tiles = 1;
  return (int)(tiles > max_blocks ? max_blocks : tiles);
}

inline bool check_quant_args(const char* who, int bits, int lo_mode, int promotion) {
  if (bits < 2 || bits > 24) {
    set_error("%s: bits=%d outside [2, 24]", who, bits);
    return false;
  }
  if (lo_mode != FQ_LO_ZERO && lo_mode != FQ_LO_NEG_MAX) {
    set_error("%s: bad lo_mode %d", who, lo_mode);
    return false;
  }
  if (promotion != FQ_PROMOTION_LEGACY && promotion != FQ_PROMOTION_NEP50) {
    set_error("%s: bad promotion %d", who, promotion);
    return false;
  }
  return true;
}

// y = roundf(clip(x, lo, hi) / d) * s with {d, s, lo, hi} read from device memory (fq_quant.cu).  `reverse`
// makes the first-scheduled blocks take the END of the tensor: after a forward range pass those are the
// lines most likely still in L2.
int launch_forward_scalar_dev(const DLTensor* x, const float* qp_dev, const DLTensor* y, const DLTensor* codes,
                              bool reverse, void* stream);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#ifdef __CUDACC__

// mshadow::red::sum::Reduce(dst, src, residual): Kahan step
__device__ __forceinline__ void kahan_add(float& s, float& c, float a) {
  const float y = __fsub_rn(a, c);
  const float t = __fadd_rn(s, y);
  c = __fsub_rn(__fsub_rn(t, s), y);
  s = t;
}

// Host scalar math of convert_conv2d.py:57-64 + ste_func.py:41 replayed on the device.
__device__ __forceinline__ void compute_qparams(float max_, int bits, int is_signed, int lo_mode, int promotion,
                                                float* qp) {
  const int qmax = is_signed ? ((1 << (bits - 1)) - 1) : ((1 << bits) - 1);
  float d, s;
  if (promotion == FQ_PROMOTION_LEGACY) {
    const double s64 = (double)max_ / (double)qmax;     // numpy.float32 / int -> float64
    d = (float)(s64 + 1e-10);                           // float64 + 1e-10, then DType(scalar)
    s = (float)s64;
  } else {
    s = __fdiv_rn(max_, (float)qmax);
    d = __fadd_rn(s, 1e-10f);
  }
  qp[FQ_QP_D] = d;
  qp[FQ_QP_S] = s;
  qp[FQ_QP_LO] = (lo_mode == FQ_LO_NEG_MAX) ? -max_ : 0.0f;
  qp[FQ_QP_HI] = max_;
}

__device__ __forceinline__ float absmax4(float m, float4 v) {
  return fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fabsf(v.y), fmaxf(fabsf(v.z), fabsf(v.w))));
}

// max |x| of every row segment inside the block's slice [begin, end) -> atomicMax on ws->rowmax[row].
// Rows shorter than 2048 elements are taken one per warp, longer ones by the whole block.
template <bool KEEP>
__device__ __forceinline__ void absmax_segments(const float* __restrict__ x, int64_t begin, int64_t end, int64_t L,
                                                Workspace* ws, float* red) {
  if (begin >= end) return;
  const int64_t r0 = begin / L;
  const int64_t r1 = (end - 1) / L;
  if (L < 2048) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int64_t r = r0 + warp; r <= r1; r += nw) {
      const int64_t lo = max(begin, r * L), hi = min(end, (r + 1) * L);
      float m = 0.f;
      for (int64_t i = lo + lane; i < hi; i += 32) m = fmaxf(m, fabsf(__ldg(x + i)));
      m = warp_max(m);
      if (lane == 0) atomicMax(&ws->rowmax[r], __float_as_uint(m));
    }
  } else {
    for (int64_t r = r0; r <= r1; ++r) {
      const int64_t lo = max(begin, r * L), hi = min(end, (r + 1) * L);
      float m = 0.f;
      for_range<false, KEEP>(
          x, lo, hi, [&](int64_t, float4 v) { m = absmax4(m, v); }, [&](int64_t, float v) { m = fmaxf(m, fabsf(v)); });
      m = block_max(m, red);
      if (threadIdx.x == 0) atomicMax(&ws->rowmax[r], __float_as_uint(m));
    }
  }
}

// Whole-block epilogue run by exactly one block after all atomicMax have landed:
// rows -> out_rows, rowmax <- 0, sequential Kahan mean, qparams.
__device__ __forceinline__ void finish_rows(Workspace* ws, int64_t rows, const FinishParams& fin) {
  __shared__ float stage[1024];
  float s = 0.f, c = 0.f;
  const bool want_mean = fin.out_mean != nullptr || (fin.qparams != nullptr && fin.input_max == nullptr);
  for (int64_t base = 0; base < rows; base += 1024) {
    const int m = (int)min((int64_t)1024, rows - base);
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const float v = __uint_as_float(__ldcg(&ws->rowmax[base + i]));
      ws->rowmax[base + i] = 0u;
      stage[i] = v;
      if (fin.out_rows) fin.out_rows[base + i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0 && want_mean)
      for (int i = 0; i < m; ++i) kahan_add(s, c, stage[i]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float mean = want_mean ? __fdiv_rn(s, (float)rows) : 0.f;
    if (fin.out_mean) fin.out_mean[0] = mean;
    if (fin.qparams) {
      const float max_ = fin.input_max ? fin.input_max[0] : mean;
      compute_qparams(max_, fin.bits, fin.is_signed, fin.lo_mode, fin.promotion, fin.qparams);
    }
    __threadfence();
  }
}

#endif  // __CUDACC__
}  // namespace fq
