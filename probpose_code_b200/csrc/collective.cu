// pp_allgather: the path's ONE exchange step - every rank contributes its (B_local, K, 7) fp32 records
// and receives all of them in rank order - as a single ncclAllGather on the caller's communicator and stream.
// Replaces mmengine's end-of-epoch pickled `collect_results` (mmengine/evaluator, reached from tools/test.py:136).
//
// The library does not link NCCL: the communicator belongs to the host framework (torch.distributed's
// ProcessGroupNCCL hands out its ncclComm_t), so the entry point is resolved at run time from the libnccl that is
// already loaded into the process - the one that created the communicator.
#include "common.cuh"

#include <dlfcn.h>

#include <mutex>

namespace pp {
namespace {

typedef int (*AllGatherFn)(const void*, void*, size_t, int /*ncclDataType_t*/, void* /*ncclComm_t*/, cudaStream_t);
typedef const char* (*ErrStrFn)(int);
constexpr int kNcclFloat32 = 7;  // ncclFloat32 in nccl.h (stable since NCCL 2.0)

struct Nccl {
  AllGatherFn all_gather = nullptr;
  ErrStrFn err_str = nullptr;
};

const Nccl& nccl() {
  static Nccl api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {  // already loaded by the framework that owns the communicator
      h = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
      if (h) break;
    }
    if (!h) {  // statically linked into the framework, or loaded under another name: search the global scope
      if (dlsym(RTLD_DEFAULT, "ncclAllGather")) h = RTLD_DEFAULT;
    }
    if (!h) return;
    api.all_gather = reinterpret_cast<AllGatherFn>(dlsym(h, "ncclAllGather"));
    api.err_str = reinterpret_cast<ErrStrFn>(dlsym(h, "ncclGetErrorString"));
  });
  return api;
}

}  // namespace
}  // namespace pp

extern "C" int pp_allgather(void* nccl_comm, const float* send, float* recv, int64_t floats_per_rank, void* stream) {
  using namespace pp;
  PP_REQUIRE(nccl_comm != nullptr, PP_ERR_INVALID, "pp_allgather: communicator is NULL");
  PP_REQUIRE(floats_per_rank >= 0, PP_ERR_INVALID, "pp_allgather: negative count %lld", (long long)floats_per_rank);
  if (floats_per_rank == 0) return PP_OK;
  PP_REQUIRE(send && recv, PP_ERR_INVALID, "pp_allgather: send / recv must be non-NULL");
  const Nccl& api = nccl();
  PP_REQUIRE(api.all_gather != nullptr, PP_ERR_UNSUPPORTED,
             "pp_allgather: no NCCL library is loaded in this process (the communicator's owner must have loaded libnccl.so.2)");
  const int rc = api.all_gather(send, recv, (size_t)floats_per_rank, kNcclFloat32, nccl_comm, (cudaStream_t)stream);
  PP_REQUIRE(rc == 0, PP_ERR_CUDA, "ncclAllGather failed: %s", api.err_str ? api.err_str(rc) : "unknown NCCL error");
  count_launch();
  return PP_OK;
}
