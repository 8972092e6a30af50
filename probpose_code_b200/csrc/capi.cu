// C-ABI glue: error string, version, launch counter.
#include "engine_ops.cuh"

#include <string.h>

#include <mutex>
#include <vector>

namespace pp {

static thread_local char g_error[1024] = "";
thread_local int64_t g_launch_count = 0;
thread_local cudaAccessPolicyWindow L2Window::win = {};

bool pdl_enabled() {
  static const bool on = getenv("PP_NO_PDL") == nullptr;
  return on;
}

bool attention_use_tc() {
  static const bool on = !(getenv("PP_ATTENTION") && strcmp(getenv("PP_ATTENTION"), "mma") == 0);
  return on;
}

int device_sm_count() {
  static std::atomic<int> cache[256] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 256) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

namespace {
std::vector<OverflowReader>& overflow_readers() {
  static std::vector<OverflowReader> v;
  return v;
}
std::mutex& overflow_mutex() {
  static std::mutex m;
  return m;
}
}  // namespace

void register_overflow_reader(OverflowReader fn) {
  std::lock_guard<std::mutex> lk(overflow_mutex());
  overflow_readers().push_back(fn);
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

}  // namespace pp

extern "C" const char* pp_last_error(void) { return pp::g_error; }
extern "C" const char* pp_version(void) { return "probpose_b200 0.1.0 sm_100a"; }

extern "C" int pp_operand_overflow(int32_t clear, int32_t* flagged) {
  using namespace pp;
  PP_REQUIRE(flagged != nullptr, PP_ERR_INVALID, "pp_operand_overflow: flagged is NULL");
  *flagged = 0;
  PP_CHECK_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(overflow_mutex());
  for (OverflowReader fn : overflow_readers()) {
    unsigned v = 0;
    PP_REQUIRE(fn(clear, &v) == 0, PP_ERR_CUDA, "pp_operand_overflow: could not read a device flag: %s",
               cudaGetErrorString(cudaGetLastError()));
    *flagged += v ? 1 : 0;
  }
  return PP_OK;
}

extern "C" int pp_attention(int32_t precision, const void* qkv_op, int32_t batch, int32_t tokens, int32_t heads,
                            int32_t head_dim, void* out_op, int32_t impl, void* stream) {
  using namespace pp;
  PP_REQUIRE(batch >= 0 && heads >= 1, PP_ERR_INVALID, "pp_attention: bad batch %d / heads %d", batch, heads);
  PP_REQUIRE(batch == 0 || (qkv_op && out_op), PP_ERR_INVALID, "pp_attention: qkv_op and out_op must be non-NULL");
  PP_REQUIRE(precision == PP_PREC_FP16X3 || precision == PP_PREC_BF16 || precision == PP_PREC_FP16, PP_ERR_UNSUPPORTED,
             "pp_attention: precision %d is not a tensor-core mode", precision);
  PP_REQUIRE(impl >= 0 && impl <= 2, PP_ERR_INVALID, "pp_attention: impl %d (0 default, 1 mma.sync, 2 tcgen05)", impl);
  const bool tc = impl == 2 || (impl == 0 && attention_use_tc());
  return tc ? launch_attention_tc(precision, qkv_op, batch, tokens, heads, head_dim, out_op, (cudaStream_t)stream)
            : launch_attention_mma(precision, qkv_op, batch, tokens, heads, head_dim, out_op, (cudaStream_t)stream);
}
