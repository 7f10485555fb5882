// NCCL bound at run time (dlopen) so that libisokann_b200.so loads, and single-GPU contexts
// work, on machines without NCCL.  Only the collectives of SURVEY section 8(e) are used:
// all-reduce(SUM) of the flat gradient (+ packed loss) per optimiser step and all-gather of
// the sharded chi / K-chi vectors once per iteration.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace ik {

struct Nccl {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr;  // optional (NCCL >= 2.18)
};

static Nccl *g_nccl = nullptr;

Nccl *nccl_load(std::string &err) {
  if (g_nccl) return g_nccl;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  // $ISOKANN_NCCL_LIB first; a libnccl already mapped into the process (e.g. by torch) is found by soname
  const char *names[] = {getenv("ISOKANN_NCCL_LIB"), "libnccl.so.2", "libnccl.so",
                         "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  void *h = nullptr;
  for (const char *nm : names) {
    if (!nm || !nm[0]) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    err = std::string("cannot dlopen libnccl: ") + dlerror();
    return nullptr;
  }
  Nccl *n = new Nccl;
  n->handle = h;
#define IK_SYM(field, name)                                      \
  *(void **)(&n->field) = dlsym(h, name);                         \
  if (!n->field) {                                                \
    err = std::string("missing NCCL symbol ") + name;             \
    delete n;                                                     \
    return nullptr;                                               \
  }
  IK_SYM(GetUniqueId, "ncclGetUniqueId")
  IK_SYM(CommInitRank, "ncclCommInitRank")
  IK_SYM(CommDestroy, "ncclCommDestroy")
  IK_SYM(AllReduce, "ncclAllReduce")
  IK_SYM(AllGather, "ncclAllGather")
  IK_SYM(GetErrorString, "ncclGetErrorString")
#undef IK_SYM
  *(void **)(&n->CommSplit) = dlsym(h, "ncclCommSplit");
  g_nccl = n;
  return n;
}

int nccl_get_unique_id(Nccl *n, void *id128, std::string &err) {
  ncclUniqueId id;
  ncclResult_t r = n->GetUniqueId(&id);
  if (r != ncclSuccess) {
    err = std::string("ncclGetUniqueId: ") + n->GetErrorString(r);
    return ISOKANN_ERR_NCCL;
  }
  std::memcpy(id128, &id, 128);
  return ISOKANN_OK;
}

void *nccl_comm_init(Nccl *n, int world, int rank, const void *id128, std::string &err) {
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = n->CommInitRank(&comm, world, id, rank);
  if (r != ncclSuccess) {
    err = std::string("ncclCommInitRank: ") + n->GetErrorString(r);
    return nullptr;
  }
  return (void *)comm;
}

// a second communicator over the same ranks whose kernels use at most max_ctas SMs (ncclCommSplit + ncclConfig_t):
// the collectives that run beside the GEMMs of the backward pass go through it.  nullptr if unsupported.
void *nccl_comm_split_limited(Nccl *n, void *comm, int rank, int max_ctas) {
  if (!n || !comm || !n->CommSplit) return nullptr;
  ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
  cfg.maxCTAs = max_ctas;
  cfg.minCTAs = 1;
  ncclComm_t out = nullptr;
  if (n->CommSplit((ncclComm_t)comm, 0, rank, &out, &cfg) != ncclSuccess) return nullptr;
  return (void *)out;
}

void nccl_comm_destroy(Nccl *n, void *comm) {
  if (n && comm) n->CommDestroy((ncclComm_t)comm);
}

int nccl_allreduce_sum_f32(Nccl *n, void *comm, float *buf, size_t count, cudaStream_t s, std::string &err) {
  ncclResult_t r = n->AllReduce(buf, buf, count, ncclFloat32, ncclSum, (ncclComm_t)comm, s);
  if (r != ncclSuccess) {
    err = std::string("ncclAllReduce: ") + n->GetErrorString(r);
    return ISOKANN_ERR_NCCL;
  }
  return ISOKANN_OK;
}

int nccl_allgather_f32(Nccl *n, void *comm, const float *send, float *recv, size_t count_per_rank, cudaStream_t s,
                       std::string &err) {
  ncclResult_t r = n->AllGather(send, recv, count_per_rank, ncclFloat32, (ncclComm_t)comm, s);
  if (r != ncclSuccess) {
    err = std::string("ncclAllGather: ") + n->GetErrorString(r);
    return ISOKANN_ERR_NCCL;
  }
  return ISOKANN_OK;
}

}  // namespace ik
