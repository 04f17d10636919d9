// comm.cu — the one collective of the hot path behind the C ABI: an NCCL all-gather of the per-channel
// {bit_errors, bits} counters, so that a C# host with one process (or thread) per GPU can collect the BER table of a
// channel set sharded over 2 / 4 / 8 GPUs without any Python.  There is no other exchange anywhere on the path: channels
// are independent, every rank demodulates its own contiguous block (shard.py: channel c -> rank floor(c*G/C)).
//
// NCCL is resolved at run time with dlopen("libnccl.so.2"): libqpskcuda.so keeps no link-time dependency on it (a
// single-GPU host needs none), and inside a process that already carries an NCCL (torch's) the same copy is reused.
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; every call goes through the table below

#include <mutex>

#include "common.cuh"

namespace qpsk {

struct NcclApi {
  void* so = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.so) break;
    }
    if (!api.so) return;
    auto sym = [&](const char* n) { return dlsym(api.so, n); };
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.CommCount = (decltype(api.CommCount))sym("ncclCommCount");
    api.CommUserRank = (decltype(api.CommUserRank))sym("ncclCommUserRank");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = api.GetVersion && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.CommCount && api.CommUserRank &&
             api.AllGather && api.GetErrorString;
  });
  return api;
}

void set_text_error(const char* text);   // core.cu: fills qpsk_last_cuda_error()

static int nccl_fail(ncclResult_t r, const char* what) {
  char buf[256];
  snprintf(buf, sizeof buf, "%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
  set_text_error(buf);
  return QPSK_ERR_CUDA;
}

#define QPSK_NCCL_TRY(expr)                           \
  do {                                                \
    ncclResult_t _r = (expr);                         \
    if (_r != ncclSuccess) return nccl_fail(_r, #expr); \
  } while (0)

}  // namespace qpsk

using namespace qpsk;

struct qpsk_comm {
  ncclComm_t comm = nullptr;
  int n_ranks = 0, rank = 0, device = 0;
  cudaStream_t stream = nullptr;
  DevBuf<uint32_t> d_pad, d_all;     // padded local block / gathered table for the host entry point
};

extern "C" {

int qpsk_comm_unique_id(uint8_t* id, int cap) {
  if (!id) return QPSK_ERR_NULL;
  if (cap < (int)sizeof(ncclUniqueId)) return QPSK_ERR_CAPACITY;
  NcclApi& a = nccl();
  if (!a.ok) {
    set_text_error("libnccl.so.2 could not be loaded");
    return QPSK_ERR_UNSUPPORTED;
  }
  ncclUniqueId u;
  QPSK_NCCL_TRY(a.GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return QPSK_OK;
}

int qpsk_comm_create(const uint8_t* id, int n_ranks, int rank, qpsk_comm** out) {
  if (!id || !out) return QPSK_ERR_NULL;
  *out = nullptr;
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return QPSK_ERR_RANGE;
  NcclApi& a = nccl();
  if (!a.ok) {
    set_text_error("libnccl.so.2 could not be loaded");
    return QPSK_ERR_UNSUPPORTED;
  }
  QPSK_TRY(ensure_device());
  qpsk_comm* c = new (std::nothrow) qpsk_comm();
  if (!c) return QPSK_ERR_NOMEM;
  c->device = current_device();
  c->n_ranks = n_ranks;
  c->rank = rank;
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclResult_t r = a.CommInitRank(&c->comm, n_ranks, u, rank);   // collective: every rank of the job calls it
  if (r != ncclSuccess) {
    delete c;
    return nccl_fail(r, "ncclCommInitRank");
  }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    a.CommDestroy(c->comm);
    delete c;
    return QPSK_ERR_CUDA;
  }
  *out = c;
  return QPSK_OK;
}

int qpsk_comm_destroy(qpsk_comm* c) {
  if (!c) return QPSK_OK;
  cudaSetDevice(c->device);
  if (c->stream) {
    cudaStreamSynchronize(c->stream);
    cudaStreamDestroy(c->stream);
  }
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
  return QPSK_OK;
}

int qpsk_comm_info(qpsk_comm* c, int* n_ranks, int* rank, int* nccl_version) {
  if (!c) return QPSK_ERR_NULL;
  int n = 0, r = 0, v = 0;
  QPSK_NCCL_TRY(nccl().CommCount(c->comm, &n));      // asked of the communicator itself, not of the creation arguments
  QPSK_NCCL_TRY(nccl().CommUserRank(c->comm, &r));
  QPSK_NCCL_TRY(nccl().GetVersion(&v));
  if (n_ranks) *n_ranks = n;
  if (rank) *rank = r;
  if (nccl_version) *nccl_version = v;
  return QPSK_OK;
}

// d_counters: this rank's uint32 {errors, bits}[channels_local] (qpsk_ber_count_dev); d_all: [n_ranks][channels_max][2] on
// every rank, rank r's block at r*channels_max, rows beyond a rank's own count zero.
int qpsk_ber_gather_dev(qpsk_comm* c, const uint32_t* d_counters, int channels_local, int channels_max, uint32_t* d_all,
                        void* stream) {
  if (!c || !d_all) return QPSK_ERR_NULL;
  if (channels_local < 0 || channels_max < 1 || channels_local > channels_max) return QPSK_ERR_RANGE;
  if (channels_local > 0 && !d_counters) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device(c->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  const uint32_t* src = d_counters;
  if (channels_local != channels_max) {              // ragged blocks: pad to the common width with {0, 0} rows
    QPSK_TRY(c->d_pad.ensure((size_t)2 * channels_max));
    QPSK_CUDA_TRY(cudaMemsetAsync(c->d_pad.p, 0, sizeof(uint32_t) * 2 * channels_max, s));
    if (channels_local > 0)
      QPSK_CUDA_TRY(cudaMemcpyAsync(c->d_pad.p, d_counters, sizeof(uint32_t) * 2 * channels_local, cudaMemcpyDeviceToDevice, s));
    src = c->d_pad.p;
  }
  QPSK_NCCL_TRY(nccl().AllGather(src, d_all, (size_t)2 * channels_max, ncclUint32, c->comm, s));
  return QPSK_OK;
}

int qpsk_ber_gather(qpsk_comm* c, const uint32_t* d_counters, int channels_local, int channels_max, uint32_t* all_host) {
  if (!c || !all_host) return QPSK_ERR_NULL;
  if (channels_max < 1) return QPSK_ERR_RANGE;
  QPSK_TRY(ensure_device(c->device));
  const size_t n = (size_t)2 * channels_max * c->n_ranks;
  QPSK_TRY(c->d_all.ensure(n));
  QPSK_TRY(qpsk_ber_gather_dev(c, d_counters, channels_local, channels_max, c->d_all.p, c->stream));
  QPSK_CUDA_TRY(cudaMemcpyAsync(all_host, c->d_all.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  QPSK_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return QPSK_OK;
}

}  // extern "C"
