// Fused single-launch paths of one converted block.
//
//  fq_forward_online : per-sample absmax -> Kahan mean -> scale -> clip/quantise   (inputs)
//      reference: quantize/convert/convert_conv2d.py:56-66, convert_dense.py:41-49, ste_func.py:41
//      (abs, max, mean, .asscalar() sync, clip, div, round, mul = 7 passes + a host round trip there;
//       here: a range launch whose last block does the mean and the scale, then the streaming quantiser
//       fed from device memory and walking the tensor backwards so that it re-reads from L2:
//       12 B/element worst case, 8 B/element of HBM traffic when the tensor fits in L2.)
//  fq_quant_weight   : optional BN fold -> per-row absmax -> scale -> quantise      (weights)
//      reference: convert_conv2d.py:47-51, 70-95; convert_dense.py:52-63; merge_bn.py:65-74
#include <stdlib.h>

#include "fq_fused.cuh"

namespace fq {

enum InputMode { kRangeOnly = 0, kOfflineTrack = 2 };
constexpr int64_t kL2ReuseElems = 24LL << 20;        // 96 MB of fp32: what may still sit in the 126 MB L2

struct InputArgs {
  const float* x;
  float* y;
  void* codes;
  int code_kind;          // 0 none, 1 i8, 2 u8, 3 i16, 4 u16, 5 i32, 6 f32
  int64_t n, rows, L, per_block;
  Workspace* ws;
  FinishParams fin;
  int defer_finish;       // range kernel: leave the maxima in ws->rowmax, online_quant_small_kernel finishes
  const float* maxima;    // online_quant_small_kernel: per-sample maxima given by the caller (data parallel: the
                          // all-gathered ones) instead of ws->rowmax; no workspace clean-up then
};

__device__ __forceinline__ void put_code1(void* p, int kind, int64_t i, float c) {
  switch (kind) {
    case 1: ((signed char*)p)[i] = (signed char)(int)c; break;
    case 2: ((unsigned char*)p)[i] = (unsigned char)(int)c; break;
    case 3: ((short*)p)[i] = (short)(int)c; break;
    case 4: ((unsigned short*)p)[i] = (unsigned short)(int)c; break;
    case 5: ((int*)p)[i] = (int)c; break;
    case 6: ((float*)p)[i] = c; break;
    default: break;
  }
}
__device__ __forceinline__ void put_code4(void* p, int kind, int64_t i, float4 c) {
  switch (kind) {
    case 1:
    case 2:
      *reinterpret_cast<uchar4*>((unsigned char*)p + i) =
          make_uchar4((unsigned char)(int)c.x, (unsigned char)(int)c.y, (unsigned char)(int)c.z, (unsigned char)(int)c.w);
      break;
    case 3:
    case 4:
      *reinterpret_cast<ushort4*>((unsigned short*)p + i) = make_ushort4(
          (unsigned short)(int)c.x, (unsigned short)(int)c.y, (unsigned short)(int)c.z, (unsigned short)(int)c.w);
      break;
    case 5: *reinterpret_cast<int4*>((int*)p + i) = make_int4((int)c.x, (int)c.y, (int)c.z, (int)c.w); break;
    case 6: *reinterpret_cast<float4*>((float*)p + i) = c; break;
    default: break;
  }
}

template <bool REVERSE>
__device__ __forceinline__ void quantise_slice(const InputArgs& a, int64_t begin, int64_t end, float d, float s,
                                               float lo, float hi) {
  const QDiv qd = QDiv::make(d);
  auto one = [&](float v, float& c) {
    c = qd.code(clipf(v, lo, hi));
    return __fmul_rn(c, s);
  };
  for_range<REVERSE, false>(
      a.x, begin, end,
      [&](int64_t i, float4 v) {
        const float4 c = qd.code4(make_float4(clipf(v.x, lo, hi), clipf(v.y, lo, hi), clipf(v.z, lo, hi), clipf(v.w, lo, hi)));
        st_stream(reinterpret_cast<float4*>(a.y + i),
                  make_float4(__fmul_rn(c.x, s), __fmul_rn(c.y, s), __fmul_rn(c.z, s), __fmul_rn(c.w, s)));
        if (a.code_kind) put_code4(a.codes, a.code_kind, i, c);
      },
      [&](int64_t i, float v) {
        float c;
        a.y[i] = one(v, c);
        if (a.code_kind) put_code1(a.codes, a.code_kind, i, c);
      });
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 6) input_path_kernel(InputArgs a) {
  __shared__ float red[32];
  __shared__ unsigned int s_last;
  if (MODE == kRangeOnly) pdl_launch_dependents();       // the quantiser behind us may start loading x (fq_common.cuh)
  const int64_t begin = (int64_t)blockIdx.x * a.per_block;
  const int64_t end = min(a.n, begin + a.per_block);

  if (MODE == kOfflineTrack) {
    // the range is tracked for the EMA but the quantiser uses input_max: a single pass, 8 B/element
    __shared__ float qp[4];
    if (threadIdx.x == 0)
      compute_qparams(a.fin.input_max[0], a.fin.bits, a.fin.is_signed, a.fin.lo_mode, a.fin.promotion, qp);
    __syncthreads();
    const float d = qp[0], s = qp[1], lo = qp[2], hi = qp[3];
    const QDiv qd = QDiv::make(d);
    if (a.L < 2048) {
      // short rows: per-warp row maxima first, then the slice is quantised out of L1/L2
      absmax_segments<true>(a.x, begin, end, a.L, a.ws, red);
      quantise_slice<true>(a, begin, end, d, s, lo, hi);
    } else if (begin < end) {
      // absmax and quantisation share the loads: walk row segments, quantise as we go
      for (int64_t r = begin / a.L; r * a.L < end; ++r) {
        const int64_t b0 = max(begin, r * a.L), b1 = min(end, (r + 1) * a.L);
        float m = 0.f;
        auto one = [&](float v, float& c) {
          m = fmaxf(m, fabsf(v));
          c = qd.code(clipf(v, lo, hi));
          return __fmul_rn(c, s);
        };
        for_range<false, false>(
            a.x, b0, b1,
            [&](int64_t i, float4 v) {
              m = absmax4(m, v);
              const float4 c =
                  qd.code4(make_float4(clipf(v.x, lo, hi), clipf(v.y, lo, hi), clipf(v.z, lo, hi), clipf(v.w, lo, hi)));
              st_stream(reinterpret_cast<float4*>(a.y + i),
                        make_float4(__fmul_rn(c.x, s), __fmul_rn(c.y, s), __fmul_rn(c.z, s), __fmul_rn(c.w, s)));
              if (a.code_kind) put_code4(a.codes, a.code_kind, i, c);
            },
            [&](int64_t i, float v) {
              float c;
              a.y[i] = one(v, c);
              if (a.code_kind) put_code1(a.codes, a.code_kind, i, c);
            });
        m = block_max(m, red);
        if (threadIdx.x == 0) atomicMax(&a.ws->rowmax[r], __float_as_uint(m));
      }
    }
  } else {
    absmax_segments<true>(a.x, begin, end, a.L, a.ws, red);     // plain loads: leave the lines in L2 for the quantiser
    if (a.defer_finish) return;                                 // the dependent quantiser derives mean and scale itself
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&a.ws->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  finish_rows(a.ws, a.rows, a.fin);
  if (threadIdx.x == 0) a.ws->ticket = 0;
}

// Offline range + tracking for long rows (L >= one tile, so a tile meets at most two rows).  ONE tile per block like
// the plain quantiser (no loop, loads issued before anything else) and NO block-wide reduction: each warp reduces
// its maxima with shuffles, folds them into two shared-memory slots with shared atomics and takes a ticket; the LAST
// warp of the block to do so owns the block's maxima and touches the global row slot -- and only if the block's
// maximum beats what the slot already holds (a plain L2 read first).  No warp ever waits for another one.
// History (2^30 elements; profiles/README.md): one tile per block with a block-wide reduce (two __syncthreads) and an
// atomicMax per tile 5.9 TB/s; up to 8 consecutive tiles per block with the running row maximum in registers and one
// block reduce per row change 6.4 (prefetching the next tile: the same); one RED.MAX per WARP straight to the row
// slot 2.1 -- two million same-address global atomics serialise at ~170 ns each; this version at 5 blocks/SM (48
// registers) 6.76 (6.57 at 2^28); with the one-row tile path separated and 6 blocks/SM (40 registers, 8 bytes of
// spill) 6.93 (6.79 at 2^28) -- the plain quantiser's speed.
__global__ void __launch_bounds__(kThreads, 6) offline_track_tiles_kernel(InputArgs a) {
  __shared__ float qp[4];
  __shared__ int64_t row_s;
  __shared__ unsigned int blk_max[2], ticket;
  const int64_t nvec = a.n >> 2;
  const float4* p4 = reinterpret_cast<const float4*>(a.x);
  const int64_t tile = blockIdx.x;
  const int64_t v0 = tile * (kTileElems / 4) + threadIdx.x;
  float4 v[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t j = v0 + u * kThreads;
    v[u] = j < nvec ? ld_stream(p4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) {
    compute_qparams(a.fin.input_max[0], a.fin.bits, a.fin.is_signed, a.fin.lo_mode, a.fin.promotion, qp);
    row_s = (tile * kTileElems) / a.L;               // one 64-bit division per block
    blk_max[0] = blk_max[1] = 0u;
    ticket = 0u;
  }
  __syncthreads();
  const float s = qp[1], lo = qp[2], hi = qp[3];
  const QDiv qd = QDiv::make(qp[0]);
  const int64_t row = row_s;
  const int64_t rel64 = (row + 1) * a.L - tile * kTileElems;         // row boundary relative to the tile start
  const int rel = rel64 > (int64_t)kTileElems ? (int)kTileElems + 8 : (int)rel64;
  const bool two = rel < (int)kTileElems;            // block-uniform: this tile holds the start of the next row
  float m_cur = 0.f, m_nxt = 0.f;
  if (!two) {                                        // the common case: the whole tile lies in one row
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) m_cur = absmax4(m_cur, v[u]);
  } else {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int ir = 4 * ((int)threadIdx.x + u * kThreads);
      const float m = absmax4(0.f, v[u]);
      if (ir < rel) m_cur = fmaxf(m_cur, m);           // selects, not branches
      else m_nxt = fmaxf(m_nxt, m);
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t j = v0 + u * kThreads;
    if (j < nvec) {
      const float4 c = qd.code4(make_float4(clipf(v[u].x, lo, hi), clipf(v[u].y, lo, hi), clipf(v[u].z, lo, hi),
                                            clipf(v[u].w, lo, hi)));
      st_stream(reinterpret_cast<float4*>(a.y) + j,
                make_float4(__fmul_rn(c.x, s), __fmul_rn(c.y, s), __fmul_rn(c.z, s), __fmul_rn(c.w, s)));
      if (a.code_kind) put_code4(a.codes, a.code_kind, 4 * j, c);
    }
  }
  m_cur = warp_max(m_cur);
  if (two) m_nxt = warp_max(m_nxt);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&blk_max[0], __float_as_uint(m_cur));
    if (two) atomicMax(&blk_max[1], __float_as_uint(m_nxt));
    __threadfence_block();
    if (atomicAdd(&ticket, 1u) == kThreads / 32 - 1) {           // the last warp: every warp's maxima are in
      const unsigned int b0 = atomicMax(&blk_max[0], 0u), b1 = atomicMax(&blk_max[1], 0u);
      if (row < a.rows && b0 > __ldcg(&a.ws->rowmax[row])) atomicMax(&a.ws->rowmax[row], b0);
      if (two && row + 1 < a.rows && b1 > __ldcg(&a.ws->rowmax[row + 1])) atomicMax(&a.ws->rowmax[row + 1], b1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Online input path of a latency-bound tensor (<= 1024 tiles): quantiser that finishes the range itself.
// ---------------------------------------------------------------------------------------------
// Launched as a programmatic dependent of input_path_kernel<range> running with defer_finish: that kernel ends as
// soon as its per-sample atomicMax'es are out -- no fence, no "last block" ticket, no Kahan mean on ITS tail.  Every
// block here loads its tile of x first (the range kernel only reads x), waits for the range grid, reads the N
// maxima, and runs the reference's sequential Kahan mean and scale math itself: redundant across blocks, but a
// block never waits for a value another block has to compute, publish and fence.  That takes the ticket atomic, two
// fences and one L2 round trip (publish qparams -> read qparams) off a chain that is nothing but round trips.
// Block 0 writes cur_max / qparams / per_sample; the last block to have read the maxima zeroes the workspace.
constexpr int kSelfFinishRowsMax = 2048;

__global__ void __launch_bounds__(kThreads, 4) online_quant_small_kernel(InputArgs a) {
  __shared__ float stage[kSelfFinishRowsMax];
  __shared__ float qp[4];
  const unsigned int nvec = (unsigned int)(a.n >> 2);
  const float4* p4 = reinterpret_cast<const float4*>(a.x);
  const unsigned int v0 = blockIdx.x * (unsigned int)(kTileElems / 4) + threadIdx.x;
  float4 v[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const unsigned int j = v0 + u * kThreads;
    v[u] = j < nvec ? ld_stream(p4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_wait();                                   // the range grid has completed and its atomics are visible
  const int rows = (int)a.rows;
  if (a.maxima != nullptr) {
    for (int i = threadIdx.x; i < rows; i += kThreads) stage[i] = __ldcg(a.maxima + i);
  } else {
    for (int i = threadIdx.x; i < rows; i += kThreads) stage[i] = __uint_as_float(__ldcg(&a.ws->rowmax[i]));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, c = 0.f;
    for (int i = 0; i < rows; ++i) kahan_add(s, c, stage[i]);
    const float mean = __fdiv_rn(s, (float)rows);                  // MXNet mean: Kahan sum / fp32(N)
    compute_qparams(mean, a.fin.bits, a.fin.is_signed, a.fin.lo_mode, a.fin.promotion, qp);
    if (blockIdx.x == 0) {
      a.fin.out_mean[0] = mean;
#pragma unroll
      for (int k = 0; k < 4; ++k) a.fin.qparams[k] = qp[k];
    }
    if (a.maxima == nullptr && atomicAdd(&a.ws->ticket, 1u) == gridDim.x - 1) {   // everyone has read the maxima
      for (int i = 0; i < rows; ++i) a.ws->rowmax[i] = 0u;
      a.ws->ticket = 0u;
    }
  }
  if (blockIdx.x == 0 && a.fin.out_rows != nullptr)
    for (int i = threadIdx.x; i < rows; i += kThreads) a.fin.out_rows[i] = stage[i];
  __syncthreads();
  const float s = qp[1], lo = qp[2], hi = qp[3];
  const QDiv qd = QDiv::make(qp[0]);
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const unsigned int j = v0 + u * kThreads;
    if (j < nvec) {
      const float4 c = qd.code4(make_float4(clipf(v[u].x, lo, hi), clipf(v[u].y, lo, hi), clipf(v[u].z, lo, hi),
                                            clipf(v[u].w, lo, hi)));
      st_stream(reinterpret_cast<float4*>(a.y) + j,
                make_float4(__fmul_rn(c.x, s), __fmul_rn(c.y, s), __fmul_rn(c.z, s), __fmul_rn(c.w, s)));
      if (a.code_kind) put_code4(a.codes, a.code_kind, 4 * (int64_t)j, c);
    }
  }
}

// FQ_ONLINE_MODE (environment, read once): how fq_forward_online runs a latency-bound tensor.
//   0  range kernel with last-block finish + dependent streaming quantiser (round 1; still the path of large,
//      ragged or > 2048-sample tensors)
//   2  range kernel without finish + dependent quantiser that finishes itself (default)
// Measured under CUDA-graph replay (profiles/r2_graph_probe_online_path.txt): 5.5 / 6.4 / 6.9 / 8.2 us per call at
// 2^14 / 2^18 / 2^20 / 2^21 elements against 7.1 / 7.9 / 8.4 / 10.0 us for mode 0.  A third variant -- ONE launch
// with a spinning grid barrier, x kept in registers, 8 B/element -- was built and measured slower than mode 2
// everywhere (6.1 / 7.5 / 8.6 / 10.9 us): the barrier's two extra L2 round trips cost more than the launch saved.
static int online_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* env = getenv("FQ_ONLINE_MODE");
    mode = (env != nullptr && env[0] == '0') ? 0 : 2;
  }
  return mode;
}

__global__ void __launch_bounds__(kThreads) finish_rows_kernel(Workspace* ws, int64_t rows, FinishParams fin) {
  finish_rows(ws, rows, fin);
}

// ---------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------
struct WeightArgs {
  const float* w;
  float* w_out;
  void* codes;
  int code_kind;
  int64_t n, cout, Lc, rows, per_block;   // Lc = elements per output channel; rows = quantisation rows
  int bits;                               // <= 0: fold only
  const float *gamma, *beta, *mean, *var, *bias;   // all NULL = no fold
  float* bias_out;
  float* scale_out;
  Workspace* ws;
  unsigned int* rowmax;                   // this tensor's row-max slots inside ws->rowmax
};

struct RowQ {      // per-row quantiser: multiplier s_r and the divisor context of d_r = s_r + 1e-10
  float s;
  QDiv q;
};

// per-channel fold factors: W' = (W * gamma) / sqrtf(var + 1e-10)      convert_conv2d.py:50
__device__ __forceinline__ void fold_factors(const WeightArgs& a, int64_t c, float& g, float& sd) {
  g = __ldg(a.gamma + c);
  sd = __fsqrt_rn(__fadd_rn(__ldg(a.var + c), 1e-10f));
}

template <bool FOLD, class CF, class VF, class SF>
__device__ __forceinline__ void for_channels(const WeightArgs& a, int64_t begin, int64_t end, CF cf, VF vf, SF sf,
                                             bool reverse) {
  // cf(c) -> per-channel context; VF(ctx, i, float4 folded) / SF(ctx, i, float folded).
  // Channel rows shorter than 2048 elements go one per warp, longer ones to the whole block.
  if (begin >= end) return;
  const int64_t c0 = begin / a.Lc, c1 = (end - 1) / a.Lc;
  if (a.Lc < 2048) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int64_t c = c0 + warp; c <= c1; c += nw) {
      float g = 1.f, sd = 1.f;
      if (FOLD) fold_factors(a, c, g, sd);
      const auto ctx = cf(c);
      const int64_t lo = max(begin, c * a.Lc), hi = min(end, (c + 1) * a.Lc);
      // 8 independent loads in flight per lane before any of them is used: these rows are tiny and the
      // loop is otherwise one L2 round trip per element
      for (int64_t i0 = lo + lane; i0 < hi; i0 += 32 * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i0 + 32 * u < hi) ? __ldg(a.w + i0 + 32 * u) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (i0 + 32 * u < hi) sf(ctx, i0 + 32 * u, FOLD ? __fdiv_rn(__fmul_rn(v[u], g), sd) : v[u]);
        }
      }
    }
  } else {
    for (int64_t c = c0; c <= c1; ++c) {
      float g = 1.f, sd = 1.f;
      if (FOLD) fold_factors(a, c, g, sd);
      const auto ctx = cf(c);
      const int64_t lo = max(begin, c * a.Lc), hi = min(end, (c + 1) * a.Lc);
      auto f1 = [&](float v) { return FOLD ? __fdiv_rn(__fmul_rn(v, g), sd) : v; };
      auto v4 = [&](int64_t i, float4 v) { vf(ctx, i, make_float4(f1(v.x), f1(v.y), f1(v.z), f1(v.w))); };
      auto s1 = [&](int64_t i, float v) { sf(ctx, i, f1(v)); };
      if (reverse)
        for_range<true, true>(a.w, lo, hi, v4, s1);
      else
        for_range<false, true>(a.w, lo, hi, v4, s1);
    }
  }
}

// PHASE 0: folded bias, then per-row max |W'| (or, bits <= 0, the folded weights themselves); the last block
// to finish publishes the scales.  PHASE 1: quantise with those maxima, then restore the workspace.
// Two plain launches: stream order is the barrier between them.
// blk / nblk: this block's index and the number of blocks working on THIS tensor.  MULTI (several tensors
// in one launch): no tickets -- block 0 of the tensor writes the scales in phase 1 and the host clears the
// row-max slots with a memset node afterwards.
template <bool FOLD, int PHASE, bool MULTI>
__device__ __forceinline__ void weight_phase(const WeightArgs& a, int blk, int nblk) {
  __shared__ float red[32];
  __shared__ unsigned int s_last;
  const int64_t begin = (int64_t)blk * a.per_block;
  const int64_t end = min(a.n, begin + a.per_block);
  const int64_t ch_per_row = a.cout / a.rows;

  if (PHASE == 0) {
  // folded bias: b' = ((gamma * (b - mean)) / sqrtf(var + 1e-10)) + beta      convert_conv2d.py:51
  if (FOLD && a.bias_out != nullptr) {
    for (int64_t c = (int64_t)blk * blockDim.x + threadIdx.x; c < a.cout; c += (int64_t)nblk * blockDim.x) {
      const float b = a.bias ? __ldg(a.bias + c) : 0.f;
      const float sd = __fsqrt_rn(__fadd_rn(__ldg(a.var + c), 1e-10f));
      a.bias_out[c] =
          __fadd_rn(__fdiv_rn(__fmul_rn(__ldg(a.gamma + c), __fsub_rn(b, __ldg(a.mean + c))), sd), __ldg(a.beta + c));
    }
  }

  if (a.bits <= 0) {   // fold only (merge_bn.py:65-66)
    for_channels<FOLD>(
        a, begin, end, [&](int64_t) { return 0; },
        [&](int, int64_t i, float4 v) { st_stream(reinterpret_cast<float4*>(a.w_out + i), v); },
        [&](int, int64_t i, float v) { a.w_out[i] = v; }, false);
    return;
  }

  // phase 1: per-row max of |W'|
  if (begin < end) {
    if (a.Lc < 2048) {
      // warp-per-channel: every lane keeps its own running max, merged per channel
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
      const int64_t c0 = begin / a.Lc, c1 = (end - 1) / a.Lc;
      for (int64_t c = c0 + warp; c <= c1; c += nw) {
        float g = 1.f, sd = 1.f;
        if (FOLD) fold_factors(a, c, g, sd);
        const int64_t lo = max(begin, c * a.Lc), hi = min(end, (c + 1) * a.Lc);
        float m = 0.f;
        for (int64_t i = lo + lane; i < hi; i += 32) {
          float v = __ldg(a.w + i);
          if (FOLD) v = __fdiv_rn(__fmul_rn(v, g), sd);
          m = fmaxf(m, fabsf(v));
        }
        m = warp_max(m);
        if (lane == 0) atomicMax(&a.rowmax[c / ch_per_row], __float_as_uint(m));
      }
    } else {
      const int64_t c0 = begin / a.Lc, c1 = (end - 1) / a.Lc;
      for (int64_t c = c0; c <= c1; ++c) {
        float g = 1.f, sd = 1.f;
        if (FOLD) fold_factors(a, c, g, sd);
        const int64_t lo = max(begin, c * a.Lc), hi = min(end, (c + 1) * a.Lc);
        auto f1 = [&](float v) { return FOLD ? __fdiv_rn(__fmul_rn(v, g), sd) : v; };
        float m = 0.f;
        for_range<false, true>(
            a.w, lo, hi,
            [&](int64_t, float4 v) { m = absmax4(m, make_float4(f1(v.x), f1(v.y), f1(v.z), f1(v.w))); },
            [&](int64_t, float v) { m = fmaxf(m, fabsf(f1(v))); });
        m = block_max(m, red);
        if (threadIdx.x == 0) atomicMax(&a.rowmax[c / ch_per_row], __float_as_uint(m));
      }
    }
  }

  if (!MULTI && a.scale_out != nullptr) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&a.ws->ticket, 1u) == (unsigned)nblk - 1);
    __syncthreads();
    if (s_last) {
      __threadfence();
      const float qmax = (float)((1 << (a.bits - 1)) - 1);
      for (int64_t r = threadIdx.x; r < a.rows; r += blockDim.x)
        a.scale_out[r] = __fdiv_rn(__uint_as_float(__ldcg(&a.rowmax[r])), qmax);
      if (threadIdx.x == 0) a.ws->ticket = 0;
    }
  }
  return;
  }   // PHASE 0

  if (a.bits <= 0) return;               // fold-only job inside a mixed launch: phase 0 did everything
  if (MULTI && blk == 0 && a.scale_out != nullptr) {
    const float qm = (float)((1 << (a.bits - 1)) - 1);
    for (int64_t r = threadIdx.x; r < a.rows; r += blockDim.x)
      a.scale_out[r] = __fdiv_rn(__uint_as_float(__ldcg(&a.rowmax[r])), qm);
  }

  // PHASE 1: quantise, newest lines first
  const float qmax = (float)((1 << (a.bits - 1)) - 1);
  for_channels<FOLD>(
      a, begin, end,
      [&](int64_t c) {   // {s_r, d_r}: convert_conv2d.py:76 / ste_func.py:39
        const float s = __fdiv_rn(__uint_as_float(__ldcg(&a.rowmax[c / ch_per_row])), qmax);
        RowQ rq;
        rq.s = s;
        rq.q = QDiv::make(__fadd_rn(s, 1e-10f));
        return rq;
      },
      [&](const RowQ& rq, int64_t i, float4 v) {
        const float4 q = rq.q.code4(v);
        st_stream(reinterpret_cast<float4*>(a.w_out + i),
                  make_float4(__fmul_rn(q.x, rq.s), __fmul_rn(q.y, rq.s), __fmul_rn(q.z, rq.s), __fmul_rn(q.w, rq.s)));
        if (a.code_kind) put_code4(a.codes, a.code_kind, i, q);
      },
      [&](const RowQ& rq, int64_t i, float v) {
        const float q = rq.q.code(v);
        a.w_out[i] = __fmul_rn(q, rq.s);
        if (a.code_kind) put_code1(a.codes, a.code_kind, i, q);
      },
      true);

  // restore the workspace invariant once every block has consumed the row maxima
  if (MULTI) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&a.ws->ticket2, 1u) == (unsigned)nblk - 1);
  __syncthreads();
  if (s_last) {
    for (int64_t r = threadIdx.x; r < a.rows; r += blockDim.x) a.rowmax[r] = 0u;
    if (threadIdx.x == 0) a.ws->ticket2 = 0;
  }
}

template <bool FOLD, int PHASE>
__global__ void __launch_bounds__(kThreads, 4) weight_path_kernel(WeightArgs a) {
  weight_phase<FOLD, PHASE, false>(a, (int)blockIdx.x, (int)gridDim.x);
}

// ---- every converted block's weights in one launch per phase ------------------------------------------
constexpr int kWeightBatch = 30;       // jobs per launch: the table travels as a kernel parameter (< 4 KB)

struct WeightJobDev {
  const float *w, *gamma, *beta, *mean, *var, *bias;
  float *w_out, *bias_out, *scale_out;
  int64_t n, per_block;
  int cout, rows, bits, first_block, rowmax_off, pad;
};

struct WeightBatch {
  WeightJobDev job[kWeightBatch];
  int count, total_blocks;
  Workspace* ws;
};

template <int PHASE>
__global__ void __launch_bounds__(kThreads, 4) weight_multi_kernel(const __grid_constant__ WeightBatch tb) {
  int j = 0;
  while (j + 1 < tb.count && (int)blockIdx.x >= tb.job[j + 1].first_block) ++j;
  const WeightJobDev& q = tb.job[j];
  WeightArgs a;
  a.w = q.w;
  a.w_out = q.w_out;
  a.codes = nullptr;
  a.code_kind = 0;
  a.n = q.n;
  a.cout = q.cout;
  a.Lc = q.n / q.cout;
  a.rows = q.rows;
  a.per_block = q.per_block;
  a.bits = q.bits;
  a.gamma = q.gamma;
  a.beta = q.beta;
  a.mean = q.mean;
  a.var = q.var;
  a.bias = q.bias;
  a.bias_out = q.bias_out;
  a.scale_out = q.scale_out;
  a.ws = tb.ws;
  a.rowmax = tb.ws->rowmax + q.rowmax_off;
  const int blk = (int)blockIdx.x - q.first_block;
  const int nblk = ((j + 1 < tb.count) ? tb.job[j + 1].first_block : tb.total_blocks) - q.first_block;
  if (q.gamma != nullptr)
    weight_phase<true, PHASE, true>(a, blk, nblk);
  else
    weight_phase<false, PHASE, true>(a, blk, nblk);
}

static int code_kind_of(const char* who, const View& codes, int64_t n, int* kind) {
  *kind = 0;
  if (codes.null) return 0;
  FQ_REQUIRE(codes.numel == n, "%s: codes has %lld elements, expected %lld", who, (long long)codes.numel, (long long)n);
  if (codes.code == kDLInt && codes.bits == 8) *kind = 1;
  else if (codes.code == kDLUInt && codes.bits == 8) *kind = 2;
  else if (codes.code == kDLInt && codes.bits == 16) *kind = 3;
  else if (codes.code == kDLUInt && codes.bits == 16) *kind = 4;
  else if (codes.code == kDLInt && codes.bits == 32) *kind = 5;
  else if (codes.code == kDLFloat && codes.bits == 32) *kind = 6;
  else FQ_REQUIRE(false, "%s: codes dtype (code %d, %d bits) unsupported", who, codes.code, codes.bits);
  return 0;
}

}  // namespace fq

using namespace fq;

extern "C" {

int fq_forward_online(const DLTensor* x_, int64_t n_samples, int bits, int is_signed, int lo_mode, int promotion,
                      const DLTensor* input_max_, const DLTensor* y_, const DLTensor* codes_,
                      const DLTensor* cur_max_, const DLTensor* qparams_, const DLTensor* per_sample_, void* ws,
                      void* stream) {
  const char* who = "fq_forward_online";
  View x, imax, y, codes, cur, qp, ps;
  FQ_TRY(view_of(x_, "fq_forward_online: x", false, &x));
  FQ_TRY(view_of(input_max_, "fq_forward_online: input_max", true, &imax));
  FQ_TRY(view_of(y_, "fq_forward_online: y", true, &y));
  FQ_TRY(view_of(codes_, "fq_forward_online: codes", true, &codes));
  FQ_TRY(view_of(cur_max_, "fq_forward_online: cur_max", false, &cur));
  FQ_TRY(view_of(qparams_, "fq_forward_online: qparams", true, &qp));
  FQ_TRY(view_of(per_sample_, "fq_forward_online: per_sample", true, &ps));
  FQ_REQUIRE(ws != nullptr, "%s: NULL workspace", who);
  FQ_REQUIRE(x.is_f32() && cur.is_f32() && cur.numel >= 1, "%s: x and cur_max must be float32", who);
  FQ_REQUIRE(n_samples >= 1 && n_samples <= FQ_MAX_ROWS, "%s: n_samples=%lld outside [1, %d]", who,
             (long long)n_samples, FQ_MAX_ROWS);
  FQ_REQUIRE(x.numel > 0 && x.numel % n_samples == 0, "%s: numel %lld not a positive multiple of n_samples %lld", who,
             (long long)x.numel, (long long)n_samples);
  FQ_REQUIRE(ps.null || (ps.is_f32() && ps.numel == n_samples), "%s: per_sample must be float32 [n_samples]", who);
  FQ_REQUIRE(imax.null || (imax.is_f32() && imax.numel >= 1), "%s: input_max must be float32", who);
  FQ_TRY(check_quant_args(who, bits, lo_mode, promotion));
  cudaStream_t st = (cudaStream_t)stream;

  InputArgs a = {};
  a.x = x.as<const float>();
  a.n = x.numel;
  a.rows = n_samples;
  a.L = x.numel / n_samples;
  a.ws = (Workspace*)ws;
  a.fin.out_rows = ps.null ? nullptr : ps.as<float>();
  a.fin.out_mean = cur.as<float>();
  a.fin.input_max = imax.null ? nullptr : imax.as<const float>();
  a.fin.bits = bits;
  a.fin.is_signed = is_signed;
  a.fin.lo_mode = lo_mode;
  a.fin.promotion = promotion;
  a.fin.qparams = qp.null ? nullptr : qp.as<float>();

  if (y.null) {   // range tracking only
    FQ_REQUIRE(codes.null, "%s: codes without y", who);
    const int grid = row_grid(a.n, a.L, sm_count() * 8, &a.per_block);
    input_path_kernel<kRangeOnly><<<grid, kThreads, 0, st>>>(a);
    FQ_LAUNCH_CHECK("input_path_kernel<range>");
    return 0;
  }
  FQ_REQUIRE(y.is_f32() && y.numel == x.numel, "%s: y must be float32 like x", who);
  FQ_REQUIRE(!qp.null && qp.is_f32() && qp.numel == 4, "%s: qparams must be 4 float32", who);
  FQ_REQUIRE(aligned16(x.data) && aligned16(y.data) && (codes.null || aligned16(codes.data)),
             "%s: x, y and codes must be 16-byte aligned", who);
  a.y = y.as<float>();
  a.codes = codes.null ? nullptr : codes.data;
  FQ_TRY(code_kind_of(who, codes, x.numel, &a.code_kind) == 0);

  if (!imax.null && a.L >= kTileElems && a.L % 4 == 0) {   // offline range + tracking, long rows: streaming tile kernel
    const int64_t ntiles = (a.n + kTileElems - 1) / kTileElems;
    FQ_REQUIRE(ntiles < (1LL << 31), "%s: tensor too large", who);
    offline_track_tiles_kernel<<<(unsigned)ntiles, kThreads, 0, st>>>(a);
    FQ_LAUNCH_CHECK("offline_track_tiles_kernel");
    finish_rows_kernel<<<1, kThreads, 0, st>>>(a.ws, a.rows, a.fin);
    FQ_LAUNCH_CHECK("finish_rows_kernel");
    return 0;
  }
  if (!imax.null) {   // offline range, current range still tracked: one pass
    const int grid = row_grid(a.n, a.L, sm_count() * 8, &a.per_block);
    input_path_kernel<kOfflineTrack><<<grid, kThreads, 0, st>>>(a);
    FQ_LAUNCH_CHECK("input_path_kernel<offline>");
    return 0;
  }
  {
    const int64_t ntiles = (a.n + kTileElems - 1) / kTileElems;
    // latency-bound tensors: the dependent quantiser finishes the range itself (see online_quant_small_kernel)
    if (online_mode() == 2 && a.n % 4 == 0 && a.rows <= kSelfFinishRowsMax && ntiles <= 1024) {
      a.defer_finish = 1;
      const int grid = row_grid(a.n, a.L, sm_count() * 8, &a.per_block);
      input_path_kernel<kRangeOnly><<<grid, kThreads, 0, st>>>(a);
      FQ_LAUNCH_CHECK("input_path_kernel<range, deferred>");
      FQ_CUDA(launch_dependent(online_quant_small_kernel, dim3((unsigned)ntiles), dim3(kThreads), 0, st, a));
      FQ_LAUNCH_CHECK("online_quant_small_kernel");
      return 0;
    }
  }
  // Online: the range pass (per-sample absmax, then the last block's Kahan mean and scale math) followed by
  // the streaming quantiser reading its qparams from device memory -- no host round trip.  Two plain launches
  // beat a cooperative single launch at every size (profiles/README.md): the activation is re-read from L2
  // either way, and a cooperative launch costs more than the second launch it saves.
  const int grid = row_grid(a.n, a.L, sm_count() * 8, &a.per_block);
  input_path_kernel<kRangeOnly><<<grid, kThreads, 0, st>>>(a);
  FQ_LAUNCH_CHECK("input_path_kernel<range>");
  return launch_forward_scalar_dev(x_, a.fin.qparams, y_, codes_, a.n <= kL2ReuseElems, stream);
}

// ---- persistent argument block of one block's input path ------------------------------------------------------
// A converted block calls fq_forward_online with the same arguments on every forward except the addresses of the
// activation and of its quantised copy.  The plan keeps private copies of every descriptor (shapes included), so a
// call is four machine words across the language boundary instead of seven freshly built DLTensor structs
// ("per-block call plan"; 20 us -> ~8 us of host time per call from Python).
struct FqInputPlan {
  DLTensor x, input_max, cur_max, qparams, per_sample;
  int64_t x_shape[8], s_shape[4][8];
  int has_input_max, has_qparams, has_per_sample, quantize;
  int64_t n_samples;
  int bits, is_signed, lo_mode, promotion;
};

static int plan_copy(const char* who, const DLTensor* src, DLTensor* dst, int64_t* shape) {
  FQ_REQUIRE(src->ndim >= 0 && src->ndim <= 8, "%s: at most 8 dimensions", who);
  if (src->strides != nullptr) {
    int64_t expect = 1;
    for (int i = src->ndim - 1; i >= 0; --i) {
      FQ_REQUIRE(src->shape[i] == 1 || src->strides[i] == expect, "%s: plan tensors must be compact row-major", who);
      expect *= src->shape[i];
    }
  }
  *dst = *src;
  for (int i = 0; i < src->ndim; ++i) shape[i] = src->shape[i];
  dst->shape = shape;
  dst->strides = nullptr;
  return 0;
}

int fq_input_plan_create(const DLTensor* x_like, int64_t n_samples, int bits, int is_signed, int lo_mode, int promotion,
                         const DLTensor* input_max, int quantize, const DLTensor* cur_max, const DLTensor* qparams,
                         const DLTensor* per_sample, FqInputPlan** out) {
  const char* who = "fq_input_plan_create";
  FQ_REQUIRE(out != nullptr && x_like != nullptr && cur_max != nullptr, "%s: x_like, cur_max and out are required", who);
  FQ_REQUIRE(!quantize || qparams != nullptr, "%s: a quantising plan needs qparams", who);
  FQ_TRY(check_quant_args(who, bits, lo_mode, promotion));
  FqInputPlan* p = (FqInputPlan*)calloc(1, sizeof(FqInputPlan));
  FQ_REQUIRE(p != nullptr, "%s: out of memory", who);
  int rc = plan_copy(who, x_like, &p->x, p->x_shape);
  if (rc == 0) rc = plan_copy(who, cur_max, &p->cur_max, p->s_shape[0]);
  if (rc == 0 && input_max != nullptr) rc = plan_copy(who, input_max, &p->input_max, p->s_shape[1]);
  if (rc == 0 && qparams != nullptr) rc = plan_copy(who, qparams, &p->qparams, p->s_shape[2]);
  if (rc == 0 && per_sample != nullptr) rc = plan_copy(who, per_sample, &p->per_sample, p->s_shape[3]);
  if (rc != 0) {
    free(p);
    return rc;
  }
  p->has_input_max = input_max != nullptr;
  p->has_qparams = qparams != nullptr;
  p->has_per_sample = per_sample != nullptr;
  p->quantize = quantize != 0;
  p->n_samples = n_samples;
  p->bits = bits;
  p->is_signed = is_signed;
  p->lo_mode = lo_mode;
  p->promotion = promotion;
  *out = p;
  return 0;
}

int fq_input_plan_run(const FqInputPlan* p, const void* x_data, void* y_data, void* ws, void* stream) {
  const char* who = "fq_input_plan_run";
  FQ_REQUIRE(p != nullptr && x_data != nullptr, "%s: NULL plan or x", who);
  FQ_REQUIRE(!p->quantize || y_data != nullptr, "%s: this plan quantises: y is required", who);
  DLTensor x = p->x, y = p->x;            // descriptors on the stack: a plan may be run from several threads
  x.data = const_cast<void*>(x_data);
  x.byte_offset = 0;
  y.data = y_data;
  y.byte_offset = 0;
  return fq_forward_online(&x, p->n_samples, p->bits, p->is_signed, p->lo_mode, p->promotion,
                           p->has_input_max ? &p->input_max : nullptr, p->quantize ? &y : nullptr, nullptr, &p->cur_max,
                           p->has_qparams ? &p->qparams : nullptr, p->has_per_sample ? &p->per_sample : nullptr, ws,
                           stream);
}

int fq_input_plan_destroy(FqInputPlan* p) {
  free(p);
  return 0;
}

int fq_forward_from_maxima(const DLTensor* x_, const DLTensor* maxima_, int bits, int is_signed, int lo_mode,
                           int promotion, const DLTensor* y_, const DLTensor* codes_, const DLTensor* cur_max_,
                           const DLTensor* qparams_, void* stream) {
  const char* who = "fq_forward_from_maxima";
  View x, mx, y, codes, cur, qp;
  FQ_TRY(view_of(x_, "fq_forward_from_maxima: x", false, &x));
  FQ_TRY(view_of(maxima_, "fq_forward_from_maxima: maxima", false, &mx));
  FQ_TRY(view_of(y_, "fq_forward_from_maxima: y", false, &y));
  FQ_TRY(view_of(codes_, "fq_forward_from_maxima: codes", true, &codes));
  FQ_TRY(view_of(cur_max_, "fq_forward_from_maxima: cur_max", false, &cur));
  FQ_TRY(view_of(qparams_, "fq_forward_from_maxima: qparams", false, &qp));
  FQ_REQUIRE(x.is_f32() && y.is_f32() && mx.is_f32() && cur.is_f32() && qp.is_f32() && y.numel == x.numel &&
                 cur.numel >= 1 && qp.numel == 4 && mx.numel >= 1,
             "%s: float32 tensors; y like x, qparams [4]", who);
  FQ_TRY(check_quant_args(who, bits, lo_mode, promotion));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ntiles = (x.numel + kTileElems - 1) / kTileElems;
  const bool small = x.numel > 0 && x.numel % 4 == 0 && mx.numel <= kSelfFinishRowsMax && ntiles <= 1024 &&
                     aligned16(x.data) && aligned16(y.data) && (codes.null || aligned16(codes.data));
  if (!small) {   // large or ragged tensors: mean, scale and the streaming quantiser as three launches
    FQ_TRY(fq_mean_kahan(maxima_, cur_max_, stream) == 0);
    FQ_TRY(fq_scale_from_max(cur_max_, bits, is_signed, lo_mode, promotion, qparams_, stream) == 0);
    return x.numel == 0 ? 0 : launch_forward_scalar_dev(x_, qp.as<const float>(), y_, codes_, false, stream);
  }
  InputArgs a = {};
  a.x = x.as<const float>();
  a.y = y.as<float>();
  a.codes = codes.null ? nullptr : codes.data;
  FQ_TRY(code_kind_of(who, codes, x.numel, &a.code_kind) == 0);
  a.n = x.numel;
  a.rows = mx.numel;
  a.L = x.numel / mx.numel;
  a.maxima = mx.as<const float>();
  a.fin.out_mean = cur.as<float>();
  a.fin.qparams = qp.as<float>();
  a.fin.bits = bits;
  a.fin.is_signed = is_signed;
  a.fin.lo_mode = lo_mode;
  a.fin.promotion = promotion;
  online_quant_small_kernel<<<(unsigned)ntiles, kThreads, 0, st>>>(a);     // plain launch: griddepcontrol.wait is a no-op
  FQ_LAUNCH_CHECK("online_quant_small_kernel");
  return 0;
}

int fq_quant_weight(const DLTensor* w_, int64_t rows, int bits, const DLTensor* gamma_, const DLTensor* beta_,
                    const DLTensor* mean_, const DLTensor* var_, const DLTensor* bias_, const DLTensor* w_out_,
                    const DLTensor* bias_out_, const DLTensor* scale_out_, const DLTensor* codes_, void* ws,
                    void* stream) {
  const char* who = "fq_quant_weight";
  View w, gamma, beta, mean, var, bias, w_out, bias_out, scale_out, codes;
  FQ_TRY(view_of(w_, "fq_quant_weight: w", false, &w));
  FQ_TRY(view_of(gamma_, "fq_quant_weight: gamma", true, &gamma));
  FQ_TRY(view_of(beta_, "fq_quant_weight: beta", true, &beta));
  FQ_TRY(view_of(mean_, "fq_quant_weight: mean", true, &mean));
  FQ_TRY(view_of(var_, "fq_quant_weight: var", true, &var));
  FQ_TRY(view_of(bias_, "fq_quant_weight: bias", true, &bias));
  FQ_TRY(view_of(w_out_, "fq_quant_weight: w_out", false, &w_out));
  FQ_TRY(view_of(bias_out_, "fq_quant_weight: bias_out", true, &bias_out));
  FQ_TRY(view_of(scale_out_, "fq_quant_weight: scale_out", true, &scale_out));
  FQ_TRY(view_of(codes_, "fq_quant_weight: codes", true, &codes));
  FQ_REQUIRE(ws != nullptr, "%s: NULL workspace", who);
  FQ_REQUIRE(w.is_f32() && w_out.is_f32() && w.numel == w_out.numel && w.numel > 0,
             "%s: w and w_out must be non-empty float32 of equal size", who);
  FQ_REQUIRE(w_->ndim >= 1 && w_->shape[0] >= 1, "%s: w needs a leading output-channel axis", who);
  const int64_t cout = w_->shape[0];
  const bool fold = !gamma.null;
  FQ_REQUIRE(fold == !beta.null && fold == !mean.null && fold == !var.null,
             "%s: gamma, beta, mean and var must be given together", who);
  if (fold) {
    FQ_REQUIRE(gamma.is_f32() && beta.is_f32() && mean.is_f32() && var.is_f32() && gamma.numel == cout &&
                   beta.numel == cout && mean.numel == cout && var.numel == cout,
               "%s: BN vectors must be float32 [Cout=%lld]", who, (long long)cout);
    FQ_REQUIRE(bias.null || (bias.is_f32() && bias.numel == cout), "%s: bias must be float32 [Cout]", who);
    FQ_REQUIRE(bias_out.null || (bias_out.is_f32() && bias_out.numel == cout), "%s: bias_out must be float32 [Cout]", who);
  }
  if (bits > 0) {
    FQ_REQUIRE(bits >= 2 && bits <= 24, "%s: bits=%d outside [2, 24]", who, bits);
    FQ_REQUIRE(rows >= 1 && rows <= FQ_MAX_ROWS && cout % rows == 0, "%s: rows=%lld must divide Cout=%lld", who,
               (long long)rows, (long long)cout);
    FQ_REQUIRE(scale_out.null || (scale_out.is_f32() && scale_out.numel == rows), "%s: scale_out must be float32 [rows]", who);
  } else {
    rows = 1;
    FQ_REQUIRE(fold, "%s: bits <= 0 (fold only) needs the BN vectors", who);
  }
  FQ_REQUIRE(aligned16(w.data) && aligned16(w_out.data) && (codes.null || aligned16(codes.data)),
             "%s: w, w_out and codes must be 16-byte aligned", who);

  WeightArgs a = {};
  a.w = w.as<const float>();
  a.w_out = w_out.as<float>();
  a.codes = codes.null ? nullptr : codes.data;
  FQ_TRY(code_kind_of(who, codes, w.numel, &a.code_kind) == 0);
  a.n = w.numel;
  a.cout = cout;
  a.Lc = w.numel / cout;
  a.rows = rows;
  a.bits = bits;
  a.gamma = fold ? gamma.as<const float>() : nullptr;
  a.beta = fold ? beta.as<const float>() : nullptr;
  a.mean = fold ? mean.as<const float>() : nullptr;
  a.var = fold ? var.as<const float>() : nullptr;
  a.bias = bias.null ? nullptr : bias.as<const float>();
  a.bias_out = bias_out.null ? nullptr : bias_out.as<float>();
  a.scale_out = scale_out.null ? nullptr : scale_out.as<float>();
  a.ws = (Workspace*)ws;
  a.rowmax = a.ws->rowmax;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = row_grid(a.n, a.Lc, sm_count() * 8, &a.per_block);
  if (fold) {
    weight_path_kernel<true, 0><<<grid, kThreads, 0, st>>>(a);
    FQ_LAUNCH_CHECK("weight_path_kernel<fold, range>");
    if (bits > 0) {
      weight_path_kernel<true, 1><<<grid, kThreads, 0, st>>>(a);
      FQ_LAUNCH_CHECK("weight_path_kernel<fold, quantise>");
    }
    return 0;
  }
  weight_path_kernel<false, 0><<<grid, kThreads, 0, st>>>(a);
  FQ_LAUNCH_CHECK("weight_path_kernel<range>");
  weight_path_kernel<false, 1><<<grid, kThreads, 0, st>>>(a);
  FQ_LAUNCH_CHECK("weight_path_kernel<quantise>");
  return 0;
}

int fq_quant_weight_multi(const FqWeightJob* jobs, int n_jobs, const DLTensor* w_out_flat_,
                          const DLTensor* bias_out_flat_, const DLTensor* scale_out_flat_, void* ws, void* stream) {
  const char* who = "fq_quant_weight_multi";
  View wf, bf, sf;
  FQ_REQUIRE(jobs != nullptr && n_jobs >= 1, "%s: no jobs", who);
  FQ_REQUIRE(ws != nullptr, "%s: NULL workspace", who);
  FQ_TRY(view_of(w_out_flat_, "fq_quant_weight_multi: w_out_flat", false, &wf));
  FQ_TRY(view_of(bias_out_flat_, "fq_quant_weight_multi: bias_out_flat", true, &bf));
  FQ_TRY(view_of(scale_out_flat_, "fq_quant_weight_multi: scale_out_flat", true, &sf));
  FQ_REQUIRE(wf.is_f32() && aligned16(wf.data) && (bf.null || bf.is_f32()) && (sf.null || sf.is_f32()),
             "%s: flat outputs must be float32 (w_out_flat 16-byte aligned)", who);
  cudaStream_t st = (cudaStream_t)stream;
  const int budget = sm_count() * 8;
  int done = 0;
  while (done < n_jobs) {
    WeightBatch tb = {};
    tb.ws = (Workspace*)ws;
    int blocks = 0, rowmax_used = 0, cnt = 0;
    bool any_quant = false;
    for (; done + cnt < n_jobs && cnt < kWeightBatch; ++cnt) {
      const FqWeightJob& jb = jobs[done + cnt];
      View w, gamma, beta, mean, var, bias;
      FQ_TRY(view_of(jb.w, "fq_quant_weight_multi: w", false, &w));
      FQ_TRY(view_of(jb.gamma, "fq_quant_weight_multi: gamma", true, &gamma));
      FQ_TRY(view_of(jb.beta, "fq_quant_weight_multi: beta", true, &beta));
      FQ_TRY(view_of(jb.mean, "fq_quant_weight_multi: mean", true, &mean));
      FQ_TRY(view_of(jb.var, "fq_quant_weight_multi: var", true, &var));
      FQ_TRY(view_of(jb.bias, "fq_quant_weight_multi: bias", true, &bias));
      FQ_REQUIRE(w.is_f32() && w.numel > 0 && aligned16(w.data) && jb.w->ndim >= 1, "%s: job %d: bad weight tensor", who, done + cnt);
      const int64_t cout = jb.w->shape[0];
      const bool fold = !gamma.null;
      FQ_REQUIRE(fold == !beta.null && fold == !mean.null && fold == !var.null, "%s: job %d: gamma, beta, mean and var must be given together", who, done + cnt);
      if (fold)
        FQ_REQUIRE(gamma.numel == cout && beta.numel == cout && mean.numel == cout && var.numel == cout &&
                       (bias.null || bias.numel == cout) && jb.bias_off >= 0 && !bf.null && jb.bias_off + cout <= bf.numel,
                   "%s: job %d: BN vectors / bias output do not match Cout=%lld", who, done + cnt, (long long)cout);
      int64_t rows = jb.rows;
      if (jb.bits > 0) {
        FQ_REQUIRE(jb.bits >= 2 && jb.bits <= 24 && rows >= 1 && rows <= FQ_MAX_ROWS && cout % rows == 0,
                   "%s: job %d: bits=%d rows=%lld Cout=%lld", who, done + cnt, jb.bits, (long long)rows, (long long)cout);
        any_quant = true;
      } else {
        FQ_REQUIRE(fold, "%s: job %d: bits <= 0 (fold only) needs the BN vectors", who, done + cnt);
        rows = 1;
      }
      FQ_REQUIRE(jb.w_off >= 0 && jb.w_off % 4 == 0 && jb.w_off + w.numel <= wf.numel, "%s: job %d: bad w_off", who, done + cnt);
      FQ_REQUIRE(jb.scale_off < 0 || (!sf.null && jb.scale_off + rows <= sf.numel), "%s: job %d: bad scale_off", who, done + cnt);
      if (rowmax_used + rows > FQ_MAX_ROWS) break;        // next launch
      WeightJobDev& q = tb.job[cnt];
      q.w = w.as<const float>();
      q.gamma = fold ? gamma.as<const float>() : nullptr;
      q.beta = fold ? beta.as<const float>() : nullptr;
      q.mean = fold ? mean.as<const float>() : nullptr;
      q.var = fold ? var.as<const float>() : nullptr;
      q.bias = bias.null ? nullptr : bias.as<const float>();
      q.w_out = wf.as<float>() + jb.w_off;
      q.bias_out = fold ? bf.as<float>() + jb.bias_off : nullptr;
      q.scale_out = (jb.scale_off >= 0 && jb.bits > 0) ? sf.as<float>() + jb.scale_off : nullptr;
      q.n = w.numel;
      q.cout = (int)cout;
      q.rows = (int)rows;
      q.bits = jb.bits;
      q.first_block = blocks;
      q.rowmax_off = rowmax_used;
      blocks += row_grid(w.numel, w.numel / cout, budget, &q.per_block);
      rowmax_used += (int)rows;
    }
    FQ_REQUIRE(cnt > 0, "%s: a single job needs more than %d row slots", who, FQ_MAX_ROWS);
    tb.count = cnt;
    tb.total_blocks = blocks;
    weight_multi_kernel<0><<<blocks, kThreads, 0, st>>>(tb);
    FQ_LAUNCH_CHECK("weight_multi_kernel<range>");
    if (any_quant) {
      weight_multi_kernel<1><<<blocks, kThreads, 0, st>>>(tb);
      FQ_LAUNCH_CHECK("weight_multi_kernel<quantise>");
      FQ_CUDA(cudaMemsetAsync(tb.ws->rowmax, 0, sizeof(unsigned int) * rowmax_used, st));
    }
    done += cnt;
  }
  return 0;
}

}  // extern "C"
