/*
 * fq.h -- C ABI of the B200-native fake-quantization hot path (libfq_b200.so).
 *
 * Drop-in boundary for hey-yahei/Quantization.MXNet.  The reference has no FFI
 * of its own for this path: every entry point below replaces a run of MXNet
 * NDArray ops (or NumPy code) issued from the reference's Python, cited as
 * <file>:<lines> relative to the reference root.  The only C-ABI convention the
 * reference uses is libmxnet's (quantize/freeze/freeze.py:32,67-76):
 * int return code (0 = ok), message from a *GetLastError() call, out-params by
 * pointer.  This header follows the same convention.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from fq_last_error() (thread local);
 *   - tensors are borrowed `const DLTensor*` (DLPack), device kDLCUDA,
 *     compact row-major, lanes == 1; NULL is allowed only where stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     nothing here allocates, synchronises the host or changes the device;
 *   - `ws` is a caller-owned device workspace of fq_workspace_bytes() bytes that
 *     was zero-filled once with fq_workspace_init(); kernels leave it zeroed.
 *     One workspace serves one stream at a time;
 *   - `promotion` selects how the reference's HOST scalar math is replayed:
 *     FQ_PROMOTION_LEGACY  numpy.float32 (op) python-number -> float64 (NumPy 1.x,
 *                          what an MXNet 1.x install computes),
 *     FQ_PROMOTION_NEP50   stays float32 (NumPy >= 2).
 */
#ifndef FQ_B200_FQ_H_
#define FQ_B200_FQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- minimal DLPack (ABI-compatible with dlpack/dlpack.h v0.8 / v1.x) ---- */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3, kDLCUDAManaged = 13 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0, kDLUInt = 1, kDLFloat = 2 } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides;     /* NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;
#endif

#if defined(__GNUC__)
#define FQ_API __attribute__((visibility("default")))
#else
#define FQ_API
#endif

#define FQ_PROMOTION_LEGACY 0
#define FQ_PROMOTION_NEP50 1

#define FQ_STE_IDENTITY 0   /* ste_func.py:43-44: dx = dy (the reference)          */
#define FQ_STE_CLIP_MASK 1  /* extension: dx = dy * 1[lo <= x <= hi]               */

#define FQ_LO_ZERO 0        /* unsigned inputs, and ALL Dense inputs (convert_dense.py:49) */
#define FQ_LO_NEG_MAX 1     /* signed Conv2D inputs (convert_conv2d.py:60)          */

#define FQ_QP_D 0           /* qparams[4] = {divisor d, multiplier s, clip lo, clip hi} */
#define FQ_QP_S 1
#define FQ_QP_LO 2
#define FQ_QP_HI 3

#define FQ_MAX_BATCH 64     /* tensors one multi-tensor launch takes */
#define FQ_MAX_ROWS 65536   /* rows (samples / channels / groups) a fused kernel accepts */
#define FQ_MAX_STAT_BLOCKS 32768 /* channels fq_channel_stats accepts */

/* ---- library ---- */
FQ_API int fq_version(void);
FQ_API const char* fq_last_error(void);                 /* cf. MXGetLastError, freeze.py:32 */
FQ_API size_t fq_workspace_bytes(void);
FQ_API int fq_workspace_init(void* ws, size_t bytes, void* stream);
FQ_API int fq_sm_count(int* out);

/* ---- K1 range reductions ------------------------------------------------ */
/* out[r] = max |x[r, :]| for x viewed as [rows, numel/rows].
 * convert_conv2d.py:56 (rows=N), :75 (Cout), :86 (G), :92 (1); convert_dense.py:41,54,60;
 * nn/quantized_conv.py:65. */
FQ_API int fq_absmax_rows(const DLTensor* x, int64_t rows, const DLTensor* out, void* ws, void* stream);
/* out2 = {min x, max x}.  nn/quantized_conv.py:68-69; distribution_calibrate.py:34-35. */
FQ_API int fq_minmax(const DLTensor* x, const DLTensor* out2, void* ws, void* stream);
/* out[0] = MXNet CPU mean: sequential Kahan fp32 sum / fp32(n).  convert_conv2d.py:56 `.mean()`.
 * v may also be [rows, n] with out [rows]: one mean per row in a single launch. */
FQ_API int fq_mean_kahan(const DLTensor* v, const DLTensor* out, void* stream);
/* cur_max[0] = mean_n max_chw |x| in ONE launch; per_sample (NULL or [n_samples]) receives the maxima.
 * convert_conv2d.py:56; convert_dense.py:41. */
FQ_API int fq_input_range(const DLTensor* x, int64_t n_samples, const DLTensor* per_sample,
                   const DLTensor* cur_max, void* ws, void* stream);
/* qparams[4] from a device-resident max_: the host scalar math of
 * convert_conv2d.py:57-64 / convert_dense.py:42-47 + ste_func.py:41 without the .asscalar() sync. */
FQ_API int fq_scale_from_max(const DLTensor* max_, int bits, int is_signed, int lo_mode, int promotion,
                      const DLTensor* qparams, void* stream);

/* ---- K2 fake-quant forward ---------------------------------------------- */
/* y = roundf(clip(x, lo, hi) / d) * s, qparams on the device.  ste_func.py:41.
 * codes: NULL, or int8/uint8/int16/uint16/int32/float32 tensor receiving the rounded quotient. */
FQ_API int fq_forward_scalar(const DLTensor* x, const DLTensor* qparams, const DLTensor* y,
                      const DLTensor* codes, void* stream);
/* Same with host scalars (LinearQuantizeSTE called with Python floats). use_clip=0 -> ste_func.py:39. */
FQ_API int fq_forward_scalar_host(const DLTensor* x, float d, float s, float lo, float hi, int use_clip,
                           const DLTensor* y, const DLTensor* codes, void* stream);
/* y[r,:] = roundf(x[r,:] / (scale[r] + 1e-10f)) * scale[r].  ste_func.py:39 with an NDArray scale
 * (convert_conv2d.py:77-79, 88-90, 93-95; convert_dense.py:56-58, 61-63). */
FQ_API int fq_forward_rows(const DLTensor* x, int64_t rows, const DLTensor* scale, const DLTensor* y,
                    const DLTensor* codes, void* stream);
/* Online input path without a host round trip: a range launch (per-sample absmax; its last block does the
 * Kahan mean and the scale math on the device) and the streaming quantiser, which walks the tensor backwards
 * so that it re-reads from L2.  On latency-bound tensors (<= 4 Mi elements) the quantiser is a programmatic
 * dependent launch of the range kernel (it loads its tile before griddepcontrol.wait); both launches are plain
 * stream work and may be captured in a CUDA graph.
 * convert_conv2d.py:56-66 / convert_dense.py:41-49.  cur_max[0] and qparams[4] are written.
 * input_max != NULL selects the offline range (`input_max.asscalar()`, :58) while cur_max is still tracked;
 * y == NULL tracks the range only (quantize_input disabled, :55-57). */
FQ_API int fq_forward_online(const DLTensor* x, int64_t n_samples, int bits, int is_signed, int lo_mode,
                      int promotion, const DLTensor* input_max, const DLTensor* y, const DLTensor* codes,
                      const DLTensor* cur_max, const DLTensor* qparams, const DLTensor* per_sample,
                      void* ws, void* stream);
/* Persistent argument block ("call plan") of one converted block's input path: everything fq_forward_online takes
 * except the addresses of the activation and of its quantised copy is captured once (the plan keeps its own copies
 * of the descriptors; the state tensors cur_max / qparams / input_max / per_sample must stay where they are while
 * the plan lives).  fq_input_plan_run(plan, x, y, ws, stream) is then exactly
 * fq_forward_online(x_like with data = x, ..., y_like with data = y, ...) -- same kernels, same checks, same results --
 * at four machine words per call.  quantize == 0: range tracking only (y ignored).  A plan is immutable after
 * creation; it may be run concurrently with different workspaces / streams.
 * Replaces the per-forward Python of convert_conv2d.py:55-66 / convert_dense.py:41-49. */
typedef struct FqInputPlan FqInputPlan;
FQ_API int fq_input_plan_create(const DLTensor* x_like, int64_t n_samples, int bits, int is_signed, int lo_mode,
                                int promotion, const DLTensor* input_max, int quantize, const DLTensor* cur_max,
                                const DLTensor* qparams, const DLTensor* per_sample, FqInputPlan** out);
FQ_API int fq_input_plan_run(const FqInputPlan* plan, const void* x_data, void* y_data, void* ws, void* stream);
FQ_API int fq_input_plan_destroy(FqInputPlan* plan);
/* The online input path when the per-sample maxima come from outside (data parallel: the all-gathered maxima of
 * the GLOBAL batch, [N] in sample order): Kahan mean -> cur_max[0], scale math -> qparams[4], quantise -- one launch
 * for latency-bound tensors (every block derives the mean itself), three for large or ragged ones. */
FQ_API int fq_forward_from_maxima(const DLTensor* x, const DLTensor* maxima, int bits, int is_signed, int lo_mode,
                                  int promotion, const DLTensor* y, const DLTensor* codes, const DLTensor* cur_max,
                                  const DLTensor* qparams, void* stream);
/* Weight path, two launches and no host round trip: optional BN fold -> per-row absmax -> scale -> quantise.
 * rows in {1 (layer), G (group), Cout (channel)}; bits <= 0 folds only (merge_bn.py:65-74).
 * gamma/beta/mean/var all NULL = no fold; bias may be NULL (treated as zeros, initialize.py:65-70).
 * convert_conv2d.py:47-51, 70-95; convert_dense.py:52-63. */
FQ_API int fq_quant_weight(const DLTensor* w, int64_t rows, int bits, const DLTensor* gamma, const DLTensor* beta,
                    const DLTensor* mean, const DLTensor* var, const DLTensor* bias, const DLTensor* w_out,
                    const DLTensor* bias_out, const DLTensor* scale_out, const DLTensor* codes,
                    void* ws, void* stream);

/* Winograd-domain per-channel weight fake-quantisation (wino_quantize = "F23" | "F43" | "F63";
 * convert_conv2d.py:71-83, wino_matrix.py:28-60):
 *   U = G w G^T per 3x3 kernel;  M_o = max |U[o]|;  s_o = M_o / (2^(bits-1)-1);
 *   Uq = roundf(U / (s_o + 1e-10f)) * s_o;  w_out = GI Uq GTI.
 * w, w_out: float32 [Cout, Cin, 3, 3]; G [a, 3], GI = pinv(G) [3, a], GTI = pinv(G^T) [a, 3], float32, a in {4, 6, 8}
 * (the caller computes the pseudo-inverses, as the reference does with numpy); scale_out: optional [Cout].
 * Every dot product runs over its contraction index in ascending order as acc = a0*b0, acc = fma(ak, bk, acc). */
FQ_API int fq_quant_weight_wino(const DLTensor* w, const DLTensor* G, const DLTensor* GI, const DLTensor* GTI, int bits,
                                const DLTensor* w_out, const DLTensor* scale_out, void* ws, void* stream);
/* Its straight-through backward: dw = G^T ((GI^T (dwq GTI^T)) G), the four products autograd replays. */
FQ_API int fq_wino_backward(const DLTensor* dwq, const DLTensor* G, const DLTensor* GI, const DLTensor* GTI,
                            const DLTensor* dw, void* stream);

/* The weight paths of MANY blocks (a whole network) in one launch per phase: launch-bound nets spend more time
 * launching 50 tiny kernels than running them.  Outputs go to caller-owned flat float32 buffers at the given
 * element offsets (w_off a multiple of 4; bias_off / scale_off < 0 = not wanted), so a cached job table stays
 * valid when only the output allocation changes.  Fold-only jobs (bits <= 0) and jobs with and without BN fold
 * can be mixed.  Results are bit-identical to calling fq_quant_weight per block. */
typedef struct {
  const DLTensor *w, *gamma, *beta, *mean, *var, *bias;   /* gamma..var NULL together = no fold; bias may be NULL */
  int64_t rows;                                           /* 1 | G | Cout */
  int32_t bits;                                           /* <= 0: fold only */
  int32_t reserved;
  int64_t w_off, bias_off, scale_off;
} FqWeightJob;
FQ_API int fq_quant_weight_multi(const FqWeightJob* jobs, int n_jobs, const DLTensor* w_out_flat,
                                 const DLTensor* bias_out_flat, const DLTensor* scale_out_flat, void* ws,
                                 void* stream);

/* Backward of the fake-BN fold (convert_conv2d.py:47-51 behind the identity STE) for many blocks in one launch:
 *   dw = (dwq / sd) * gamma;  dgamma = sum_row((dwq / sd) * w) + (dbq / sd) * (bias - mean);
 *   dbias = (dbq / sd) * gamma;  dbeta = dbq;   sd = sqrt(var + 1e-10).
 * dbq / bias / dbias / dbeta may be NULL.  Replaces 6-8 framework launches per block per QAT step. */
typedef struct {
  const DLTensor *dwq, *dbq, *w, *gamma, *mean, *var, *bias;   /* inputs  */
  const DLTensor *dw, *dgamma, *dbias, *dbeta;                 /* outputs */
} FqFoldBwdJob;
FQ_API int fq_fold_backward_multi(const FqFoldBwdJob* jobs, int n_jobs, void* stream);

/* ---- K3 straight-through estimator backward ----------------------------- */
FQ_API int fq_ste_backward(const DLTensor* dy, const DLTensor* x, const DLTensor* qparams, const DLTensor* dx,
                    int mode, void* stream);

/* ---- K4 EMA   convert.py:66-78 ------------------------------------------ */
/* scalar_cur != 0: `cur` plays the host numpy.float32 of convert.py:70 (promotion applies);
 * scalar_cur == 0: `cur` is an NDArray (running_mean / running_var, convert.py:75-78). */
FQ_API int fq_ema_update(const DLTensor* state, const DLTensor* cur, double momentum, int scalar_cur,
                  int promotion, void* stream);

/* Per-channel batch statistics of a conv output y [N, C, H, W] for the fake-BN EMA (convert_conv2d.py:148-153):
 *   mean[c] = sum_{n,h,w} y / (N*H*W);  var[c] = sum (y - mean[c])^2 / (N*H*W).
 * ONE pass over y (4 B/element): float64 shifted moments {n, S1 = sum(y-K), S2 = sum(y-K)^2, K = y[0,c,0,0]} per
 * channel, from which mean and var follow algebraically with the reference's fp32-rounded mean; within 2 ULP of the
 * reference's Kahan-compensated fp32 sums, deterministic.  mean/var (together) and parts (float64 [C, 4], the
 * records) are each optional.  C <= FQ_MAX_STAT_BLOCKS. */
FQ_API int fq_channel_stats(const DLTensor* y, const DLTensor* mean, const DLTensor* var, const DLTensor* parts,
                            void* ws, void* stream);
/* Data parallel: parts float64 [R, C, 4] = the records of R ranks (all-gathered; C may span every fake-BN layer of
 * a network) -> mean/var [C] of the GLOBAL batch, same formula, same bound. */
FQ_API int fq_channel_stats_finish(const DLTensor* parts, const DLTensor* mean, const DLTensor* var, void* stream);

/* ---- K5 KL calibration   quantize/distribution_calibrate.py ------------- */
/* counts[bin] += 1 for every clipped non-zero element; counts is uint64/int64 [bins+1].  :39-45
 * bad_flag (NULL or int32 [1]): |= 1 when the tensor holds a negative value or a NaN, or max_ is not > 0 -- the
 * reference's asserts of :35-36, which it evaluates on every batch; the kernel itself drops such elements. */
FQ_API int fq_hist_nonzero(const DLTensor* x, const DLTensor* max_, int bins, int promotion,
                    const DLTensor* counts, const DLTensor* bad_flag, void* stream);
/* The same for n_tensors (<= FQ_MAX_BATCH per launch, more are chunked) layer inputs in ONE launch: the
 * thread blocks are shared out in proportion to the tensor sizes, so small layers cost no launch of their own.
 * xs[i]: float32, 16-byte aligned; its frozen max is maxes[i * max_stride + max_offset];
 * counts: (u)int64 [n_tensors, bins+1]; bad_flags: NULL or int32 [n_tensors] (see fq_hist_nonzero). */
FQ_API int fq_hist_nonzero_multi(const DLTensor* const* xs, int n_tensors, const DLTensor* maxes, int max_stride,
                                 int max_offset, int bins, int promotion, const DLTensor* counts,
                                 const DLTensor* bad_flags, void* stream);
/* hist = (first ? 0 : hist) + float32(counts); counts <- 0; seen_last[0] |= counts[bins] != 0.  :47,103-104
 * hist: float32 [n]; counts: (u)int64 [steps, n] (steps >= 1) -- the batches are folded in order, one float32
 * add per batch as the reference does, so a data-parallel run may all-reduce the counts of many batches at once. */
FQ_API int fq_hist_accumulate_f32(const DLTensor* counts, const DLTensor* hist, int first,
                           const DLTensor* seen_last, void* stream);
/* best[l] = first strict arg-min of the KL divergence over i in [min_bins, bins).  :117-171
 * hist: float32 [n_data] or [layers, n_data] with n_data in {bins, bins+1}; best: int32 [layers];
 * divergence: caller-provided float64 [layers, bins] scratch that receives D_i (entries below
 * min_bins are left untouched); margin: NULL or float64 [layers] receiving (runner-up - best) / |best| -- below
 * ~1e-12 the reference's own choice depends on the last place of its libm's log, so callers flag small margins. */
FQ_API int fq_kl_search(const DLTensor* hist, int levels, int min_bins, int bins, int promotion,
                 const DLTensor* best, const DLTensor* divergence, const DLTensor* margin, void* stream);
/* input_max[0] = (best + 0.5) * (fm_max / bins).  examples/simulate_quantization.py:310,314 */
FQ_API int fq_kl_threshold(const DLTensor* best, const DLTensor* fm_max, int bins, const DLTensor* input_max,
                    void* stream);

/* ---- K6 integer export -------------------------------------------------- */
/* MXNet contrib.quantize(out_type="int8"), zero centred.  freeze.py:100-103.
 * range2 = {min_range, max_range} on the device; out_range2 receives {-real, +real}. */
FQ_API int fq_quantize_int8_export(const DLTensor* w, const DLTensor* range2, const DLTensor* out_i8,
                            const DLTensor* out_range2, void* stream);
/* nn/quantized_conv.py:54-61: int32 codes + scale_out[0]; range2 = {min, max} on the device. */
FQ_API int fq_qconv_quantize(const DLTensor* x, const DLTensor* range2, const DLTensor* codes_i32,
                      const DLTensor* scale_out, void* stream);
/* nn/quantized_conv.py:74-76: y = float(acc) * (s_in * s_w). */
FQ_API int fq_qconv_dequantize(const DLTensor* acc_i32, const DLTensor* s_in, const DLTensor* s_w,
                        const DLTensor* y, void* stream);

/* ---- QConv2D integer convolution on the tensor cores (nn/quantized_conv.py:106-159) -------
 * fq_qconv_pack_input: fp32 NCHW x -> spatially zero-padded NHWC 8-bit codes xq [N, H+2ph, W+2pw, C] with the
 *   arithmetic of fq_qconv_quantize (clip, divide by scale, roundf); the padding holds the code of 0.0, because the
 *   reference pads first and quantises the padded tensor (:108-116).  int8 xq for symmetric ranges, uint8 xq for
 *   [0, max] ranges (the caller guarantees that the codes fit).  C % 16 == 0, xq 16-byte aligned.  scale_out[0]
 *   receives the scale.
 * fq_qconv_pack_weight: fp32 [Cout, Cg, KH, KW] -> int8 codes [Cout, KH, KW, Cg] (K-major for the GEMM).
 * fq_qconv_igemm: implicit GEMM on tcgen05 (kind::i8, int32 accumulators in tensor memory), fused epilogue
 *   out = float(max?(acc + bias_q)) * (s_in * s_w)  (:143-158).  Cg % 16 == 0; out float32 [N, Cout, Ho, Wo]. */
FQ_API int fq_qconv_pack_input(const DLTensor* x, const DLTensor* range2, int pad_h, int pad_w, const DLTensor* xq,
                               const DLTensor* scale_out, void* stream);
FQ_API int fq_qconv_pack_weight(const DLTensor* w, const DLTensor* range2, const DLTensor* wq, const DLTensor* scale_out,
                                void* stream);
FQ_API int fq_qconv_igemm(const DLTensor* xq, const DLTensor* wq, const DLTensor* bias_q, const DLTensor* s_in,
                          const DLTensor* s_w, int stride_h, int stride_w, int groups, int relu, const DLTensor* out,
                          void* stream);

/* ---- data-parallel collectives for hosts that own an ncclComm_t ------------------------
 * One process per GPU, batch sharded by sample, weights replicated.  `nccl_comm` is the host's ncclComm_t passed as
 * void*; libnccl is resolved at run time from the libraries the process already carries (so that the communicator
 * and the functions belong to the same NCCL instance), else from fq_nccl_load(path) / libnccl.so.2 -- it is not a
 * link-time dependency of libfq_b200.so.  Everything is enqueued on `stream`; nothing synchronises the host.
 * A torch host uses torch.distributed instead (quantization/mxnet_b200/dist.py). */
#define FQ_REDUCE_SUM 0
#define FQ_REDUCE_MAX 1
#define FQ_REDUCE_MIN 2
FQ_API int fq_nccl_load(const char* path);            /* NULL: whatever NCCL is already loaded, else libnccl.so.2 */
/* in place; float32/float64/(u)int32/(u)int64.  KL first-batch maxima: MAX on [L]; QAT gradients: SUM on the bucket. */
FQ_API int fq_dist_all_reduce(const DLTensor* t, int op, void* nccl_comm, void* stream);
/* out = [world x in] in rank order. */
FQ_API int fq_dist_all_gather(const DLTensor* in, const DLTensor* out, void* nccl_comm, void* stream);
/* current_input_max of the GLOBAL batch (convert_conv2d.py:56): all-gather of the shard's per-sample maxima
 * ([N/R] -> per_sample_all [N], sample order) + the reference's Kahan mean -> cur_max[0], on every rank. */
FQ_API int fq_dist_input_range(const DLTensor* per_sample_local, const DLTensor* per_sample_all,
                               const DLTensor* cur_max, void* nccl_comm, void* stream);
/* Histogram counts of `steps` batches ([steps, n], 32- or 64-bit): integer SUM over the ranks, then
 * fq_hist_accumulate_f32 -- every rank ends with the float32 histograms of a single-GPU run (:47,103-104). */
FQ_API int fq_dist_hist_fold(const DLTensor* counts, const DLTensor* hist, int first, const DLTensor* seen_last,
                             void* nccl_comm, void* stream);
/* fake-BN batch statistics of the GLOBAL batch (convert_conv2d.py:150-153): all-gather of the ranks' float64 records
 * (parts_local [C, 4] from fq_channel_stats -> parts_all [R, C, 4]) + fq_channel_stats_finish. */
FQ_API int fq_dist_channel_stats(const DLTensor* parts_local, const DLTensor* parts_all, const DLTensor* mean,
                                 const DLTensor* var, void* nccl_comm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FQ_B200_FQ_H_ */
