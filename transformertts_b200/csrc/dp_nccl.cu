// Data-parallel gradient exchange of the training step over NCCL, behind the C ABI (include/ttsb.h: ttsb_dp_*).
// The reference has no distributed code; BASELINE.json asks for "NCCL allreduce over NVLink on gradient
// buckets only".  NCCL is resolved at run time with dlopen (the process that hosts torch already has libnccl.so.2 mapped;
// TTSB_NCCL_LIB names another copy), so libttsb.so itself has no link-time dependency on it.
//
// ncclAllReduce(sum) in place on contiguous slices of the flat fp32 gradient buffer, enqueued on the caller's stream:
// stream-ordered like every other entry point, so the host overlaps a bucket with the remaining backward kernels by giving
// it a second stream and two events (transformertts_b200/utils/data_parallel.py: NativeGradSync).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ttsb.h"
#include "ttsb_host.h"

namespace ttsb {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat = 7, kNcclSum = 0;

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};

static NcclApi* nccl() {
  static NcclApi api{};
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void* h = nullptr;
  if (const char* e = getenv("TTSB_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already mapped by the host process (torch)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
  return api.ok ? &api : nullptr;
}

static int nccl_fail(NcclApi* a, ncclResult_t r, const char* what) {
  set_last_error("%s failed: NCCL error %d (%s)", what, (int)r, a->GetErrorString(r));
  return TTSB_ERR_CUDA;
}

}  // namespace ttsb

using namespace ttsb;

struct ttsb_dp_comm {
  ncclComm_t comm;
  int rank, world;
};

extern "C" int ttsb_dp_unique_id(void* id_out) {
  if (!id_out) { set_last_error("ttsb_dp_unique_id: id_out is NULL"); return TTSB_ERR_INVALID_ARGUMENT; }
  NcclApi* a = nccl();
  if (!a) { set_last_error("ttsb_dp: libnccl.so.2 could not be loaded (set TTSB_NCCL_LIB)"); return TTSB_ERR_UNSUPPORTED; }
  ncclUniqueId id;
  ncclResult_t r = a->GetUniqueId(&id);
  if (r) return nccl_fail(a, r, "ncclGetUniqueId");
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

extern "C" int ttsb_dp_init(const void* unique_id, int rank, int world, ttsb_dp_comm** out) {
  if (!unique_id || !out || world < 1 || rank < 0 || rank >= world) { set_last_error("ttsb_dp_init: bad arguments"); return TTSB_ERR_INVALID_ARGUMENT; }
  NcclApi* a = nccl();
  if (!a) { set_last_error("ttsb_dp: libnccl.so.2 could not be loaded (set TTSB_NCCL_LIB)"); return TTSB_ERR_UNSUPPORTED; }
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ttsb_dp_comm* c = new ttsb_dp_comm{nullptr, rank, world};
  ncclResult_t r = a->CommInitRank(&c->comm, world, id, rank);   // binds to the calling thread's current CUDA device
  if (r) { delete c; return nccl_fail(a, r, "ncclCommInitRank"); }
  *out = c;
  return 0;
}

extern "C" int ttsb_dp_allreduce_bucket(ttsb_dp_comm* comm, float* buf, int64_t count, void* stream) {
  if (!comm || !buf || count < 0) { set_last_error("ttsb_dp_allreduce_bucket: bad arguments"); return TTSB_ERR_INVALID_ARGUMENT; }
  if (count == 0 || comm->world == 1) return 0;
  NcclApi* a = nccl();
  ncclResult_t r = a->AllReduce(buf, buf, (size_t)count, kNcclFloat, kNcclSum, comm->comm, static_cast<cudaStream_t>(stream));
  if (r) return nccl_fail(a, r, "ncclAllReduce");
  return 0;
}

extern "C" int ttsb_dp_destroy(ttsb_dp_comm* comm) {
  if (!comm) return 0;
  NcclApi* a = nccl();
  ncclResult_t r = a ? a->CommDestroy(comm->comm) : 0;
  delete comm;
  if (r) return nccl_fail(a, r, "ncclCommDestroy");
  return 0;
}
