// Multi-GPU plumbing: one process per GPU, slab partition along the slowest axis.
// NCCL is dlopen()ed (torch's bundled libnccl.so.2 when the host passes its path) so the library
// loads on a single GPU box without it.  Halo planes are contiguous (natural z-slowest layout), so
// the exchange is a pair of ncclSend/ncclRecv per neighbour and component with no pack kernel; the
// PCG scalars are reduced in place on the device with ncclAllReduce (no host round-trip).
#include <dlfcn.h>

#include <cstring>

#include "device.cuh"

typedef struct { char internal[128]; } NcclUid;
typedef int (*fn_GetUniqueId)(NcclUid*);
typedef int (*fn_CommInitRank)(void**, int, NcclUid, int);
typedef int (*fn_CommDestroy)(void*);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_Send)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_Recv)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_Group)(void);
typedef const char* (*fn_ErrStr)(int);

struct NcclApi {
  void* handle = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_AllReduce AllReduce = nullptr;
  fn_Send Send = nullptr;
  fn_Recv Recv = nullptr;
  fn_Group GroupStart = nullptr, GroupEnd = nullptr;
  fn_ErrStr GetErrorString = nullptr;
};

static NcclApi g_api;

static int load_nccl(const char* path) {
  if (g_api.handle) return 0;
  const char* cand[] = {path, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* p : cand) {
    if (!p || !*p) continue;
    h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) PDE_FAIL(std::string("cannot dlopen libnccl: ") + (dlerror() ? dlerror() : "?"));
#define LOAD(name)                                                   \
  g_api.name = (decltype(g_api.name))dlsym(h, "nccl" #name);         \
  if (!g_api.name) PDE_FAIL("libnccl is missing nccl" #name);
  LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(AllReduce) LOAD(Send) LOAD(Recv)
  LOAD(GroupStart) LOAD(GroupEnd) LOAD(GetErrorString)
#undef LOAD
  g_api.handle = h;
  return 0;
}

#define NCCL_OK(call)                                                                       \
  do {                                                                                      \
    int r__ = (call);                                                                       \
    if (r__ != 0) PDE_FAIL(std::string("NCCL error: ") + g_api.GetErrorString(r__) + " at " #call); \
  } while (0)

extern "C" int pde_nccl_unique_id(const char* libnccl_path, void* id128) {
  PDE_OK(load_nccl(libnccl_path));
  NcclUid id;
  NCCL_OK(g_api.GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int pde_comm_init(pde_ctx* c, int rank, int world, const void* id128, const char* libnccl_path) {
  if (!c) PDE_FAIL("null context");
  if (world < 1 || rank < 0 || rank >= world) PDE_FAIL("bad rank/world");
  c->rank = rank;
  c->world = world;
  if (world == 1) return 0;
  PDE_OK(load_nccl(libnccl_path));
  CUDA_OK(cudaSetDevice(c->device));
  NcclUid id;
  std::memcpy(&id, id128, sizeof(id));
  NCCL_OK(g_api.CommInitRank(&c->nccl_comm, world, id, rank));
  c->nccl = &g_api;
  return 0;
}

int comm_destroy(pde_ctx* c) {
  if (c->nccl_comm && c->nccl) c->nccl->CommDestroy(c->nccl_comm);
  c->nccl_comm = nullptr;
  return 0;
}

int comm_allreduce_scal(pde_ctx* c, int slot, int count) {
  if (c->world == 1) return 0;
  NCCL_OK(c->nccl->AllReduce(c->scal + slot, c->scal + slot, (size_t)count, /*ncclFloat64*/ 8, /*ncclSum*/ 0,
                             c->nccl_comm, c->stream));
  return 0;
}

int comm_allreduce_buf(pde_ctx* c, double* buf, size_t count) {
  if (c->world == 1 || count == 0) return 0;
  NCCL_OK(c->nccl->AllReduce(buf, buf, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl_comm, c->stream));
  return 0;
}

int comm_halo_exchange(pde_ctx* c, const Grid& g, int ncomp, double* f, int depth) {
  if (c->world == 1 || g.nzl == g.nzg) return 0;   // single GPU, or a replicated (global) multigrid level
  if (depth < 1 || depth > PDE_NG) PDE_FAIL("halo depth out of range");
  if (depth > g.nzl) PDE_FAIL("halo deeper than the slab");
  const size_t n = (size_t)g.plane * depth;   // `depth` consecutive planes are contiguous
  NCCL_OK(c->nccl->GroupStart());
  for (int i = 0; i < ncomp; ++i) {
    double* b = f + (size_t)i * g.comp_stride;
    if (c->rank > 0) {
      NCCL_OK(c->nccl->Send(b, n, 8, c->rank - 1, c->nccl_comm, c->stream));                          // my first planes
      NCCL_OK(c->nccl->Recv(b - (size_t)depth * g.plane, n, 8, c->rank - 1, c->nccl_comm, c->stream));  // lower ghosts
    }
    if (c->rank < c->world - 1) {
      NCCL_OK(c->nccl->Send(b + (size_t)(g.nzl - depth) * g.plane, n, 8, c->rank + 1, c->nccl_comm, c->stream));
      NCCL_OK(c->nccl->Recv(b + (size_t)g.nzl * g.plane, n, 8, c->rank + 1, c->nccl_comm, c->stream));
    }
  }
  NCCL_OK(c->nccl->GroupEnd());
  return 0;
}
