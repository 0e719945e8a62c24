// Multi-GPU plumbing: one process per GPU, slab partition along the slowest axis.
// NCCL is dlopen()ed (torch's bundled libnccl.so.2 when the host passes its path) so the library
// loads on a single GPU box without it.  Halo planes are contiguous (natural z-slowest layout), so
// the exchange needs no pack kernel.  Default path: ONE kernel per exchange that stores this rank's boundary planes
// straight into the neighbours' mailboxes over NVLink (cudaIpc-mapped peer memory), raises a sequence flag there,
// waits for its own flags and moves the arrived planes into the ghost planes (k_halo_p2p, ~10 us against 23-53 us
// for a grouped ncclSend/ncclRecv of the same planes, which is pure latency at 2 MB).  NCCL send/recv remains as the
// fallback (PDE_B200_HALO=nccl, or when peer mapping is unavailable).  The PCG scalars are reduced in place on the
// device with ncclAllReduce (no host round-trip).
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

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

// ---- peer-memory halo exchange -----------------------------------------------------------------------
// Mailbox of a rank (device memory, mapped by both z-neighbours through cudaIpc):
//   u64 flags[32]:  ARR0 / ARR1 = sequence number of the newest planes that arrived from below / above,
//                   ACK0 / ACK1 = newest sequence number the rank below / above has finished reading from ITS inbox
//   inbox[2][2]  :  [from below, from above] x [two slots, alternating by sequence number]
// An exchange with sequence number s: wait until the neighbours have consumed exchange s-2 (slot reuse), store the
// boundary planes into their inboxes, fence, raise ARR there; wait for the own ARR flags, copy the inbox slots
// into the ghost planes, raise ACK at the neighbours.  All ranks issue the same exchanges in the same order on
// their compute streams, so a waiting kernel is always waiting for work its neighbour issues no later than it.
enum { F_ARR0 = 0, F_ARR1 = 1, F_ACK0 = 2, F_ACK1 = 3, F_WORDS = 32 };

struct P2pHalo {
  bool tried = false, ok = false;
  size_t slot_bytes = 0;
  char* mine = nullptr;
  char* peer[2] = {nullptr, nullptr};   // mailbox of rank-1 / rank+1
  unsigned long long seq = 0;
  unsigned* counters = nullptr;         // [2] last-CTA tickets
  int* err_host = nullptr;              // mapped pinned: set by a kernel whose wait timed out
  int* err_dev = nullptr;
  unsigned long long* xbuf = nullptr;   // handle exchange buffer [world][8]
};

struct HaloArgs {
  const double* src_lo; const double* src_hi;       // my first / last `depth` planes (component 0)
  double* ghost_lo; double* ghost_hi;                // my ghost planes (component 0)
  double* out_lo; double* out_hi;                    // slot in the inbox of rank-1 (its "from above") / rank+1; null: no neighbour
  const double* in_lo; const double* in_hi;          // slots of my inbox
  unsigned long long* my_flags; unsigned long long* lo_flags; unsigned long long* hi_flags;
  long long n, comp_stride;                          // doubles per component and exchange; field component stride
  int ncomp;
  unsigned long long seq;
  unsigned* counters;
  int* err;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// returns false if the wait gave up (timeout, or another wait had already timed out)
__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned long long want, int* err) {
  // bounded: a lost neighbour must not hang the GPU.  After about a minute the error flag is raised (the host
  // fails at its next poll) and every later wait returns at once.  `err` is mapped HOST memory: it is only looked at
  // every 1024 polls of a wait that is not being served.
  for (long long it = 0; it < (1LL << 26); ++it) {
    if (ld_acquire_sys(p) >= want) return true;
    if ((it & 1023) == 1023 && *(volatile int*)err) return false;
    __nanosleep(40);
  }
  *(volatile int*)err = 1;
  return false;
}

// grid-stride copy of n2 double2 with four independent loads in flight per thread
template <bool CG>
__device__ __forceinline__ void copy_d2(double2* __restrict__ d, const double2* __restrict__ s, long long n2, long long g0,
                                        long long gsz) {
  long long i = g0;
  for (; i + 3 * gsz < n2; i += 4 * gsz) {
    double2 v0, v1, v2, v3;
    if (CG) { v0 = __ldcg(s + i); v1 = __ldcg(s + i + gsz); v2 = __ldcg(s + i + 2 * gsz); v3 = __ldcg(s + i + 3 * gsz); }
    else { v0 = s[i]; v1 = s[i + gsz]; v2 = s[i + 2 * gsz]; v3 = s[i + 3 * gsz]; }
    d[i] = v0; d[i + gsz] = v1; d[i + 2 * gsz] = v2; d[i + 3 * gsz] = v3;
  }
  for (; i < n2; i += gsz) d[i] = CG ? __ldcg(s + i) : s[i];
}

__global__ void __launch_bounds__(256)
k_halo_p2p(const __grid_constant__ HaloArgs a) {
  __shared__ int s_dead;
  const bool lo = a.out_lo != nullptr, hi = a.out_hi != nullptr;
  const long long n2 = a.n >> 1;   // planes are multiples of 4 doubles
  const long long gsz = (long long)gridDim.x * blockDim.x, g0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (threadIdx.x == 0 && a.seq > 2) {
    if (lo) wait_flag(a.my_flags + F_ACK0, a.seq - 2, a.err);
    if (hi) wait_flag(a.my_flags + F_ACK1, a.seq - 2, a.err);
  }
  __syncthreads();
  for (int c = 0; c < a.ncomp; ++c) {
    if (lo) {
      copy_d2<false>(reinterpret_cast<double2*>(a.out_lo + c * a.n),
                     reinterpret_cast<const double2*>(a.src_lo + c * a.comp_stride), n2, g0, gsz);
    }
    if (hi) {
      copy_d2<false>(reinterpret_cast<double2*>(a.out_hi + c * a.n),
                     reinterpret_cast<const double2*>(a.src_hi + c * a.comp_stride), n2, g0, gsz);
    }
  }
  // one system-scope fence per CTA (after the CTA barrier it covers the stores of all its threads: fences are
  // cumulative over what the barrier made visible); a fence in every thread costs several microseconds here
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned t = atomicAdd(a.counters, 1u);
    if (t == gridDim.x - 1) {   // every CTA's planes are out (and fenced): announce them
      a.counters[0] = 0;
      __threadfence_system();
      if (lo) st_release_sys(a.lo_flags + F_ARR1, a.seq);
      if (hi) st_release_sys(a.hi_flags + F_ARR0, a.seq);
    }
    bool ok = true;
    if (lo) ok = wait_flag(a.my_flags + F_ARR0, a.seq, a.err) && ok;
    if (hi) ok = wait_flag(a.my_flags + F_ARR1, a.seq, a.err) && ok;
    s_dead = ok ? 0 : 1;
  }
  __syncthreads();
  // a wait that gave up leaves stale inbox slots: do not move them into the ghost planes (the host fails at its next
  // poll; nothing computed from them may reach an output that is fetched without one)
  const bool dead = s_dead != 0;
  for (int c = 0; c < a.ncomp && !dead; ++c) {
    if (lo) {
      copy_d2<true>(reinterpret_cast<double2*>(a.ghost_lo + c * a.comp_stride),
                    reinterpret_cast<const double2*>(a.in_lo + c * a.n), n2, g0, gsz);
    }
    if (hi) {
      copy_d2<true>(reinterpret_cast<double2*>(a.ghost_hi + c * a.comp_stride),
                    reinterpret_cast<const double2*>(a.in_hi + c * a.n), n2, g0, gsz);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(a.counters + 1, 1u);
    if (t == gridDim.x - 1) {   // the inbox slots have been read by every CTA: the neighbours may reuse them
      a.counters[1] = 0;
      __threadfence_system();
      if (lo) st_release_sys(a.lo_flags + F_ACK1, a.seq);
      if (hi) st_release_sys(a.hi_flags + F_ACK0, a.seq);
    }
  }
}

static void p2p_release(pde_ctx* c, bool all) {
  P2pHalo* h = (P2pHalo*)c->p2p;
  if (!h) return;
  for (int d = 0; d < 2; ++d) {
    if (h->peer[d]) cudaIpcCloseMemHandle(h->peer[d]);
    h->peer[d] = nullptr;
  }
  if (h->mine) cudaFree(h->mine);
  h->mine = nullptr;
  h->slot_bytes = 0;
  if (all) {
    if (h->counters) cudaFree(h->counters);
    if (h->xbuf) cudaFree(h->xbuf);
    if (h->err_host) cudaFreeHost(h->err_host);
    delete h;
    c->p2p = nullptr;
  }
}

// Collective: (re)allocate the mailboxes for slots of at least `need` bytes and map the neighbours' ones.
static int p2p_ensure(pde_ctx* c, size_t need) {
  if (!c->p2p) {
    P2pHalo* h = new P2pHalo();
    c->p2p = h;
    const char* e = getenv("PDE_B200_HALO");
    if (e && std::string(e) == "nccl") h->tried = true;   // forced fallback
  }
  P2pHalo* h = (P2pHalo*)c->p2p;
  if (h->tried && !h->ok) return 0;
  if (h->ok && h->slot_bytes >= need) return 0;
  h->tried = true;
  h->ok = false;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  // This function is collective.  A rank whose local set-up fails must still reach every all-reduce below, or the others
  // block forever: local failures are only recorded and the ranks agree on the outcome through the (always present) scalar
  // slots before the exchange buffer is used.
  int setup_fail = 0;
  if (!h->counters) {
    if (cudaMalloc(&h->counters, 2 * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(h->counters, 0, 2 * sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc(&h->xbuf, (size_t)c->world * 8 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaHostAlloc(&h->err_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess) {
      setup_fail = 1;
      cudaGetLastError();
    } else {
      *h->err_host = 0;
      if (cudaHostGetDevicePointer((void**)&h->err_dev, h->err_host, 0) != cudaSuccess) { setup_fail = 1; cudaGetLastError(); }
    }
  }
  {
    double sf = (double)setup_fail;
    CUDA_OK(cudaMemcpyAsync(c->scal + (S_NSLOTS - 1), &sf, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NCCL_OK(c->nccl->AllReduce(c->scal + (S_NSLOTS - 1), c->scal + (S_NSLOTS - 1), 1, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl_comm, c->stream));
    CUDA_OK(cudaMemcpyAsync(&sf, c->scal + (S_NSLOTS - 1), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (sf != 0.0) {   // somebody could not set up: NCCL send/recv everywhere
      if (setup_fail) {
        if (h->counters) cudaFree(h->counters);
        if (h->xbuf) cudaFree(h->xbuf);
        if (h->err_host) cudaFreeHost(h->err_host);
        h->counters = nullptr; h->xbuf = nullptr; h->err_host = nullptr; h->err_dev = nullptr;
      }
      return 0;
    }
  }
  // nobody may still be writing into a mailbox that is about to disappear: a collective acts as the barrier
  CUDA_OK(cudaMemsetAsync(h->xbuf, 0, (size_t)c->world * 8 * sizeof(unsigned long long), c->stream));
  NCCL_OK(c->nccl->AllReduce(h->xbuf, h->xbuf, (size_t)c->world * 8, /*ncclUint64*/ 5, 0, c->nccl_comm, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  p2p_release(c, false);
  const size_t slot = ((need + need / 4 + 4095) / 4096) * 4096;
  const size_t total = F_WORDS * sizeof(unsigned long long) + 4 * slot;
  int fail = 0;
  cudaIpcMemHandle_t mh;
  if (cudaMalloc(&h->mine, total) != cudaSuccess) { fail = 1; h->mine = nullptr; cudaGetLastError(); }
  if (!fail && cudaMemset(h->mine, 0, total) != cudaSuccess) fail = 1;
  if (!fail && cudaIpcGetMemHandle(&mh, h->mine) != cudaSuccess) { fail = 1; cudaGetLastError(); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  // all-gather of the handles: every rank fills its row of a zeroed table, integer sum completes it
  std::vector<unsigned long long> tab((size_t)c->world * 8, 0ULL);
  if (!fail) std::memcpy(&tab[(size_t)c->rank * 8], &mh, 64);
  CUDA_OK(cudaMemcpyAsync(h->xbuf, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, c->stream));
  NCCL_OK(c->nccl->AllReduce(h->xbuf, h->xbuf, tab.size(), 5, 0, c->nccl_comm, c->stream));
  CUDA_OK(cudaMemcpyAsync(tab.data(), h->xbuf, tab.size() * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int d = 0; d < 2 && !fail; ++d) {
    const int nb = d == 0 ? c->rank - 1 : c->rank + 1;
    if (nb < 0 || nb >= c->world) continue;
    cudaIpcMemHandle_t ph;
    std::memcpy(&ph, &tab[(size_t)nb * 8], 64);
    bool zero = true;
    for (int q = 0; q < 8; ++q) zero = zero && tab[(size_t)nb * 8 + q] == 0ULL;
    void* pp = nullptr;
    if (zero || cudaIpcOpenMemHandle(&pp, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { fail = 1; cudaGetLastError(); }
    else h->peer[d] = (char*)pp;
  }
  // agree on the outcome (and make sure every mailbox is zeroed before anybody stores into it)
  unsigned long long f = (unsigned long long)fail;
  CUDA_OK(cudaMemcpyAsync(h->xbuf, &f, 8, cudaMemcpyHostToDevice, c->stream));
  NCCL_OK(c->nccl->AllReduce(h->xbuf, h->xbuf, 1, 5, 0, c->nccl_comm, c->stream));
  CUDA_OK(cudaMemcpyAsync(&f, h->xbuf, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  if (getenv("PDE_B200_HALO_DEBUG"))
    fprintf(stderr, "[pde_b200] rank %d: peer-memory halo mailboxes %s (slot %zu bytes)\n", c->rank,
            f != 0 ? "unavailable, using NCCL send/recv" : "mapped", slot);
  if (f != 0) { p2p_release(c, false); return 0; }   // somebody could not map: NCCL path everywhere
  h->slot_bytes = slot;
  h->seq = 0;
  h->ok = true;
  return 0;
}

int comm_check_error(pde_ctx* c) {
  P2pHalo* h = (P2pHalo*)c->p2p;
  if (h && h->err_host && *(volatile int*)h->err_host) PDE_FAIL("peer-memory halo exchange timed out waiting for a neighbour");
  return 0;
}

int comm_destroy(pde_ctx* c) {
  p2p_release(c, true);
  if (c->nccl_comm && c->nccl) c->nccl->CommDestroy(c->nccl_comm);
  c->nccl_comm = nullptr;
  return 0;
}

// what the multi-GPU layer actually does (bench.py reports it instead of assuming): halo path 0 = none yet / single
// GPU, 1 = NCCL send/recv, 2 = peer-memory mailbox kernel; exchanges and all-reduces issued so far
extern "C" int pde_comm_info(pde_ctx* c, int32_t* halo_path, int64_t* halo_exchanges, int64_t* allreduces) {
  if (!c) PDE_FAIL("null context");
  P2pHalo* h = (P2pHalo*)c->p2p;
  if (halo_path) *halo_path = c->world == 1 ? 0 : (h && h->ok ? 2 : (h && h->tried ? 1 : 0));
  if (halo_exchanges) *halo_exchanges = c->n_halo;
  if (allreduces) *allreduces = c->n_allreduce;
  return 0;
}

int comm_allreduce_scal(pde_ctx* c, int slot, int count) {
  if (c->world == 1) return 0;
  c->n_allreduce++;
  NCCL_OK(c->nccl->AllReduce(c->scal + slot, c->scal + slot, (size_t)count, /*ncclFloat64*/ 8, /*ncclSum*/ 0,
                             c->nccl_comm, c->stream));
  return 0;
}

int comm_allreduce_buf(pde_ctx* c, double* buf, size_t count) {
  if (c->world == 1 || count == 0) return 0;
  c->n_allreduce++;
  NCCL_OK(c->nccl->AllReduce(buf, buf, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl_comm, c->stream));
  return 0;
}

int comm_halo_exchange(pde_ctx* c, const Grid& g, int ncomp, double* f, int depth) {
  if (c->world == 1 || g.nzl == g.nzg) return 0;   // single GPU, or a replicated (global) multigrid level
  if (depth < 1 || depth > PDE_NG) PDE_FAIL("halo depth out of range");
  if (depth > g.nzl) PDE_FAIL("halo deeper than the slab");
  const size_t n = (size_t)g.plane * depth;   // `depth` consecutive planes are contiguous
  c->n_halo++;
  PDE_OK(p2p_ensure(c, n * ncomp * sizeof(double)));
  P2pHalo* h = (P2pHalo*)c->p2p;
  if (h && h->ok) {
    h->seq += 1;
    const size_t slot = h->slot_bytes, fb = F_WORDS * sizeof(unsigned long long);
    const int sl = (int)(h->seq & 1);
    HaloArgs a{};
    a.src_lo = f; a.src_hi = f + (size_t)(g.nzl - depth) * g.plane;
    a.ghost_lo = f - (size_t)depth * g.plane; a.ghost_hi = f + (size_t)g.nzl * g.plane;
    // rank-1 receives my first planes in ITS inbox "from above" (index 1); rank+1 in its inbox "from below" (0)
    a.out_lo = h->peer[0] ? (double*)(h->peer[0] + fb + (size_t)(2 * 1 + sl) * slot) : nullptr;
    a.out_hi = h->peer[1] ? (double*)(h->peer[1] + fb + (size_t)(2 * 0 + sl) * slot) : nullptr;
    a.in_lo = (const double*)(h->mine + fb + (size_t)(2 * 0 + sl) * slot);
    a.in_hi = (const double*)(h->mine + fb + (size_t)(2 * 1 + sl) * slot);
    a.my_flags = (unsigned long long*)h->mine;
    a.lo_flags = (unsigned long long*)h->peer[0];
    a.hi_flags = (unsigned long long*)h->peer[1];
    a.n = (long long)n; a.comp_stride = g.comp_stride; a.ncomp = ncomp; a.seq = h->seq;
    a.counters = h->counters; a.err = h->err_dev;
    long long want = ((long long)n / 2 + 255) / 256;
    static const int max_blocks = getenv("PDE_B200_HALO_BLOCKS") ? atoi(getenv("PDE_B200_HALO_BLOCKS")) : 0;
    const long long cap = max_blocks > 0 ? max_blocks : c->sm_count;   // one CTA per SM at most: all resident, they wait on each other's tickets
    const int blocks = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    k_halo_p2p<<<blocks, 256, 0, c->stream>>>(a);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
  }
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
