// Data-parallel collectives for hosts that own an ncclComm_t ("NCCL-fused variants taking an
// ncclComm_t"; 8e: which quantity travels with which collective).  A torch host uses torch.distributed
// (quantization/mxnet_b200/dist.py); an MXNet / C++ host that already has a communicator calls these instead.
//
// libnccl is NOT a link-time dependency: the symbols are resolved at run time, first among the libraries the
// process has already loaded (so that the communicator handed in and the functions called belong to the SAME NCCL
// instance -- a torch process carries its own copy), then from the path given to fq_nccl_load() / libnccl.so.2.
// Every message on this path is tiny (<= a few MB) and latency-bound, so the "fusion" is the pairing of each
// exchange with the kernel that consumes it on the same stream, with no host synchronisation in between.
#include <dlfcn.h>
#include <nccl.h>

#include "fq_common.cuh"

namespace fq {

struct NcclApi {
  ncclResult_t (*all_reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*all_gather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*comm_count)(const ncclComm_t, int*) = nullptr;
  const char* (*error_string)(ncclResult_t) = nullptr;
  void* handle = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;

static bool nccl_resolve(void* h) {
  NcclApi a;
  a.all_reduce = reinterpret_cast<decltype(a.all_reduce)>(dlsym(h, "ncclAllReduce"));
  a.all_gather = reinterpret_cast<decltype(a.all_gather)>(dlsym(h, "ncclAllGather"));
  a.comm_count = reinterpret_cast<decltype(a.comm_count)>(dlsym(h, "ncclCommCount"));
  a.error_string = reinterpret_cast<decltype(a.error_string)>(dlsym(h, "ncclGetErrorString"));
  if (a.all_reduce == nullptr || a.all_gather == nullptr || a.comm_count == nullptr) return false;
  a.handle = h;
  a.ok = true;
  g_nccl = a;
  return true;
}

static bool nccl_ready() {
  if (g_nccl.ok) return true;
  if (nccl_resolve(RTLD_DEFAULT)) return true;               // whatever NCCL the process already carries
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h != nullptr && nccl_resolve(h)) return true;
  set_error("NCCL is not loaded in this process and libnccl.so.2 was not found: call fq_nccl_load(path) first");
  return false;
}

static bool nccl_type(const View& v, ncclDataType_t* t) {
  if (v.code == kDLFloat && v.bits == 32) *t = ncclFloat32;
  else if (v.code == kDLFloat && v.bits == 64) *t = ncclFloat64;
  else if (v.code == kDLInt && v.bits == 32) *t = ncclInt32;
  else if (v.code == kDLUInt && v.bits == 32) *t = ncclUint32;
  else if (v.code == kDLInt && v.bits == 64) *t = ncclInt64;
  else if (v.code == kDLUInt && v.bits == 64) *t = ncclUint64;
  else {
    set_error("fq_dist: dtype (code %d, %d bits) has no NCCL equivalent here", v.code, v.bits);
    return false;
  }
  return true;
}

#define FQ_NCCL(call)                                                                                   \
  do {                                                                                                  \
    ncclResult_t r__ = (call);                                                                          \
    if (r__ != ncclSuccess) {                                                                           \
      ::fq::set_error("%s failed: %s", #call, g_nccl.error_string ? g_nccl.error_string(r__) : "NCCL error"); \
      return -1;                                                                                        \
    }                                                                                                   \
  } while (0)

}  // namespace fq

using namespace fq;

extern "C" {

int fq_nccl_load(const char* path) {
  if (path == nullptr) return nccl_ready() ? 0 : -1;
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);     // an already loaded library is returned as the same instance
  FQ_REQUIRE(h != nullptr, "fq_nccl_load: dlopen(%s) failed: %s", path, dlerror());
  FQ_REQUIRE(nccl_resolve(h), "fq_nccl_load: %s does not export ncclAllReduce / ncclAllGather / ncclCommCount", path);
  return 0;
}

int fq_dist_all_reduce(const DLTensor* t_, int op, void* nccl_comm, void* stream) {
  View t;
  FQ_TRY(view_of(t_, "fq_dist_all_reduce: tensor", false, &t));
  FQ_REQUIRE(nccl_comm != nullptr, "fq_dist_all_reduce: NULL communicator");
  FQ_REQUIRE(op == FQ_REDUCE_SUM || op == FQ_REDUCE_MAX || op == FQ_REDUCE_MIN, "fq_dist_all_reduce: bad op %d", op);
  FQ_TRY(nccl_ready());
  ncclDataType_t dt;
  FQ_TRY(nccl_type(t, &dt));
  if (t.numel == 0) return 0;
  const ncclRedOp_t rop = op == FQ_REDUCE_SUM ? ncclSum : (op == FQ_REDUCE_MAX ? ncclMax : ncclMin);
  FQ_NCCL(g_nccl.all_reduce(t.data, t.data, (size_t)t.numel, dt, rop, (ncclComm_t)nccl_comm, (cudaStream_t)stream));
  return 0;
}

int fq_dist_all_gather(const DLTensor* in_, const DLTensor* out_, void* nccl_comm, void* stream) {
  View in, out;
  FQ_TRY(view_of(in_, "fq_dist_all_gather: in", false, &in));
  FQ_TRY(view_of(out_, "fq_dist_all_gather: out", false, &out));
  FQ_REQUIRE(nccl_comm != nullptr, "fq_dist_all_gather: NULL communicator");
  FQ_TRY(nccl_ready());
  int world = 0;
  FQ_NCCL(g_nccl.comm_count((ncclComm_t)nccl_comm, &world));
  FQ_REQUIRE(in.code == out.code && in.bits == out.bits && out.numel == in.numel * world,
             "fq_dist_all_gather: out must hold %d x the %lld elements of in, same dtype", world, (long long)in.numel);
  ncclDataType_t dt;
  FQ_TRY(nccl_type(in, &dt));
  if (in.numel == 0) return 0;
  FQ_NCCL(g_nccl.all_gather(in.data, out.data, (size_t)in.numel, dt, (ncclComm_t)nccl_comm, (cudaStream_t)stream));
  return 0;
}

// current_input_max over the GLOBAL batch (convert_conv2d.py:56 under data parallelism): all-gather of the shard's
// per-sample maxima in rank (= sample) order, then the reference's Kahan mean on every rank.
int fq_dist_input_range(const DLTensor* per_sample_local, const DLTensor* per_sample_all, const DLTensor* cur_max,
                        void* nccl_comm, void* stream) {
  if (fq_dist_all_gather(per_sample_local, per_sample_all, nccl_comm, stream) != 0) return -1;
  return fq_mean_kahan(per_sample_all, cur_max, stream);
}

// Histogram counts of `steps` batches: integer sum over the ranks, then the per-batch float32 adds in batch order
// (distribution_calibrate.py:47,103-104) -- every rank ends with the histograms of a single-GPU run.
int fq_dist_hist_fold(const DLTensor* counts, const DLTensor* hist, int first, const DLTensor* seen_last, void* nccl_comm,
                      void* stream) {
  if (fq_dist_all_reduce(counts, FQ_REDUCE_SUM, nccl_comm, stream) != 0) return -1;
  return fq_hist_accumulate_f32(counts, hist, first, seen_last, stream);
}

// fake-BN batch statistics of the global batch (convert_conv2d.py:150-153): all-gather of the ranks' float64
// {n, S1, S2, K} records (fq_channel_stats), then the same combination a single GPU applies to its own record.
int fq_dist_channel_stats(const DLTensor* parts_local, const DLTensor* parts_all, const DLTensor* mean,
                          const DLTensor* var, void* nccl_comm, void* stream) {
  if (fq_dist_all_gather(parts_local, parts_all, nccl_comm, stream) != 0) return -1;
  return fq_channel_stats_finish(parts_all, mean, var, stream);
}

}  // extern "C"
