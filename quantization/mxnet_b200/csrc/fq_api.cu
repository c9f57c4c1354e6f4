// Library plumbing: error string, DLTensor validation, workspace, device properties.
#include <stdarg.h>
#include <string.h>

#include <string>

#include "fq_common.cuh"

namespace fq {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool view_of(const DLTensor* t, const char* name, bool allow_null, View* out) {
  *out = View();
  if (t == nullptr) {
    if (allow_null) return true;
    set_error("%s: NULL tensor", name);
    return false;
  }
  if (t->device.device_type != kDLCUDA && t->device.device_type != kDLCUDAManaged) {
    set_error("%s: device_type %d is not CUDA (this library has no CPU path)", name, (int)t->device.device_type);
    return false;
  }
  int dev = -1;
  cudaGetDevice(&dev);
  if (dev != t->device.device_id) {
    set_error("%s: tensor lives on cuda:%d but the current device is cuda:%d", name, (int)t->device.device_id, dev);
    return false;
  }
  if (t->dtype.lanes != 1) {
    set_error("%s: lanes != 1", name);
    return false;
  }
  int64_t n = 1;
  for (int i = 0; i < t->ndim; ++i) {
    if (t->shape[i] < 0) {
      set_error("%s: negative extent", name);
      return false;
    }
    n *= t->shape[i];
  }
  if (t->strides != nullptr && n > 1) {
    int64_t expect = 1;
    for (int i = t->ndim - 1; i >= 0; --i) {
      if (t->shape[i] != 1 && t->strides[i] != expect) {
        set_error("%s: tensor is not compact row-major (stride[%d]=%lld, expected %lld)", name, i,
                  (long long)t->strides[i], (long long)expect);
        return false;
      }
      expect *= t->shape[i];
    }
  }
  if (n > 0 && t->data == nullptr) {
    set_error("%s: NULL data pointer", name);
    return false;
  }
  out->data = static_cast<char*>(t->data) + t->byte_offset;
  out->numel = n;
  out->code = t->dtype.code;
  out->bits = t->dtype.bits;
  out->null = false;
  return true;
}

}  // namespace fq

extern "C" {

int fq_version(void) { return 100; }

const char* fq_last_error(void) { return fq::g_last_error.c_str(); }

size_t fq_workspace_bytes(void) { return sizeof(fq::Workspace); }

int fq_workspace_init(void* ws, size_t bytes, void* stream) {
  FQ_REQUIRE(ws != nullptr, "fq_workspace_init: NULL workspace");
  FQ_REQUIRE(bytes >= sizeof(fq::Workspace), "fq_workspace_init: workspace has %zu bytes, need %zu", bytes,
             sizeof(fq::Workspace));
  FQ_CUDA(cudaMemsetAsync(ws, 0, sizeof(fq::Workspace), (cudaStream_t)stream));
  return 0;
}

int fq_sm_count(int* out) {
  FQ_REQUIRE(out != nullptr, "fq_sm_count: NULL out");
  *out = fq::sm_count();
  return 0;
}

}  // extern "C"
