// acb_comm.cu -- communicator of the sharded search (acb_comm.hpp): NCCL for the 8-byte control
// traffic, cudaIpc peer mapping (NVLink / NVSwitch) for the match records.
#include "acb_comm.hpp"

#include <chrono>
#include <thread>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/acb200.h"

#ifndef ACB_EMULATE
#include <dlfcn.h>
#endif

namespace {

#define CKC(expr)                                                                          \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      std::fprintf(stderr, "acb200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_),   \
                   __FILE__, __LINE__, cudaGetErrorString(e_));                            \
      return ACG_E_CUDA;                                                                   \
    }                                                                                      \
  } while (0)

#ifndef ACB_EMULATE
// ---- the handful of NCCL entry points the control traffic needs, resolved at run time ----------
// (signatures as in nccl.h 2.27 / 2.28; ncclUniqueId is 128 opaque bytes passed by value)
struct NcclId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclSuccess = 0 };
enum { kNcclInt8 = 0, kNcclUint8 = 1, kNcclInt32 = 2, kNcclUint64 = 5 };
enum { kNcclMin = 3 };
struct NcclApi {
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // prefer the copy the process already carries (torch bundles its own libnccl.so.2)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      std::fprintf(stderr, "acb200: libnccl.so.2 not found: %s\n", dlerror());
      return;
    }
    bool all = true;
    auto sym = [&](auto& fp, const char* name) {
      fp = reinterpret_cast<std::remove_reference_t<decltype(fp)>>(dlsym(h, name));
      if (!fp) { all = false; std::fprintf(stderr, "acb200: %s missing from libnccl\n", name); }
    };
    sym(api.GetUniqueId, "ncclGetUniqueId");
    sym(api.CommInitRank, "ncclCommInitRank");
    sym(api.CommDestroy, "ncclCommDestroy");
    sym(api.AllGather, "ncclAllGather");
    sym(api.AllReduce, "ncclAllReduce");
    sym(api.Broadcast, "ncclBroadcast");
    sym(api.Send, "ncclSend");
    sym(api.Recv, "ncclRecv");
    sym(api.GroupStart, "ncclGroupStart");
    sym(api.GroupEnd, "ncclGroupEnd");
    sym(api.GetErrorString, "ncclGetErrorString");
    sym(api.GetVersion, "ncclGetVersion");
    api.ok = all;
  });
  return api;
}

#define CKN(expr)                                                                           \
  do {                                                                                      \
    int r_ = (expr);                                                                        \
    if (r_ != kNcclSuccess) {                                                               \
      std::fprintf(stderr, "acb200: NCCL error at %s:%d: %s\n", __FILE__, __LINE__,         \
                   nccl().GetErrorString ? nccl().GetErrorString(r_) : "?");                \
      return ACG_E_CUDA;                                                                    \
    }                                                                                       \
  } while (0)

#else  // ACB_EMULATE ------------------------------------------------------------------------------
// Dry-run fabric (tests/emu): the "ranks" are threads of one process, "device memory" is host
// memory, so a peer mapping is the pointer itself.  Lets the CPU suite drive
// acg_find_overlapping_sharded end to end (slice plan, ownership by end offset, offsets of the
// global list, growth of the receive buffer) without GPUs or NCCL.
}  // namespace
#include <condition_variable>
#include <map>
#include <string>
namespace {
struct Fabric {
  std::mutex mu;
  std::condition_variable cv;
  int nranks = 0, joined = 0, refs = 0;
  int arrived = 0;
  uint64_t generation = 0;
  std::vector<uint64_t> slots;
  uint8_t* recv = nullptr;
  void barrier(std::unique_lock<std::mutex>& lk) {
    const uint64_t gen = generation;
    if (++arrived == nranks) { arrived = 0; ++generation; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != gen; });
  }
};
std::mutex g_fab_mu;
std::map<std::string, Fabric*> g_fabrics;
uint64_t g_next_id = 1;
#endif

}  // namespace

namespace acb {

int comm_unique_id(uint8_t* id128) {
  if (!id128) return ACG_E_INVALID_ARG;
#ifndef ACB_EMULATE
  if (!nccl().ok) return ACG_E_NO_DEVICE;
  NcclId id;
  CKN(nccl().GetUniqueId(&id));
  std::memcpy(id128, id.internal, 128);
#else
  std::lock_guard<std::mutex> lk(g_fab_mu);
  std::memset(id128, 0, 128);
  const uint64_t v = g_next_id++;
  std::memcpy(id128, &v, 8);
#endif
  return ACG_OK;
}

int comm_create(const uint8_t* id128, int rank, int nranks, acg_comm** out) {
  if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return ACG_E_INVALID_ARG;
  *out = nullptr;
  acg_comm* c = new (std::nothrow) acg_comm();
  if (!c) return ACG_E_NOMEM;
  c->rank = rank;
  c->nranks = nranks;
  c->counts.assign(size_t(nranks), 0);
  auto fail = [&](int rc) { comm_destroy(c); return rc; };
#ifndef ACB_EMULATE
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(ACG_E_NO_DEVICE); }
  if (!nccl().ok) return fail(ACG_E_NO_DEVICE);
  if (cudaGetDevice(&c->device) != cudaSuccess) return fail(ACG_E_CUDA);
  NcclId id;
  std::memcpy(id.internal, id128, 128);
  NcclComm comm = nullptr;
  int r = nccl().CommInitRank(&comm, nranks, id, rank);
  if (r != kNcclSuccess) {
    std::fprintf(stderr, "acb200: ncclCommInitRank failed: %s\n", nccl().GetErrorString(r));
    return fail(ACG_E_CUDA);
  }
  c->nccl = comm;
#else
  c->device = 0;
  {
    std::lock_guard<std::mutex> lk(g_fab_mu);
    Fabric*& f = g_fabrics[std::string(reinterpret_cast<const char*>(id128), 128)];
    if (!f) { f = new Fabric(); f->nranks = nranks; f->slots.assign(size_t(nranks) + 1, 0); }
    if (f->nranks != nranks) return fail(ACG_E_INVALID_ARG);
    ++f->refs;
    c->nccl = f;
  }
#endif
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(ACG_E_CUDA);
  if (cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) return fail(ACG_E_CUDA);
  for (auto& st : c->steps)
    if (cudaEventCreate(&st.done) != cudaSuccess || cudaEventCreate(&st.begun) != cudaSuccess) return fail(ACG_E_CUDA);
  if (cudaMalloc(&c->d_counts, (size_t(nranks) + 1) * 8) != cudaSuccess) return fail(ACG_E_CUDA);
  if (cudaMallocHost(&c->h_counts, (size_t(nranks) + 1) * 8) != cudaSuccess) return fail(ACG_E_CUDA);
  if (cudaMalloc(&c->d_handle, 64) != cudaSuccess) return fail(ACG_E_CUDA);
  if (cudaMallocHost(&c->h_flag_src, 16) != cudaSuccess || cudaMalloc(&c->d_flag_src, 16) != cudaSuccess ||
      cudaMallocHost(&c->h_flags, size_t(nranks) * 8) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->poll, cudaStreamNonBlocking) != cudaSuccess)
    return fail(ACG_E_CUDA);
  c->transport = ACG_TRANSPORT_PEER;  // until a mapping fails (comm_ensure_recv)
  int rc = comm_ensure_recv(c, 1 << 16);
  if (rc) return fail(rc);
  *out = c;
  return ACG_OK;
}

void comm_destroy(acg_comm* c) {
  if (!c) return;
#ifndef ACB_EMULATE
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->recv_peer) cudaIpcCloseMemHandle(c->recv_peer);
  if (c->nccl) nccl().CommDestroy(static_cast<NcclComm>(c->nccl));
#else
  if (c->nccl) {
    std::lock_guard<std::mutex> lk(g_fab_mu);
    Fabric* f = static_cast<Fabric*>(c->nccl);
    if (--f->refs == 0) {
      for (auto it = g_fabrics.begin(); it != g_fabrics.end(); ++it)
        if (it->second == f) { g_fabrics.erase(it); break; }
      delete f;
    }
  }
#endif
  cudaFree(c->recv_own);
  cudaFree(c->send_buf);
  cudaFree(c->d_counts);
  cudaFree(c->d_handle);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->h_view) cudaFreeHost(c->h_view);
  if (c->h_flag_src) cudaFreeHost(c->h_flag_src);
  if (c->h_flags) cudaFreeHost(c->h_flags);
  cudaFree(c->d_flag_src);
  if (c->poll) cudaStreamDestroy(c->poll);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  for (auto& m : c->mark) if (m) cudaEventDestroy(m);
  for (auto& st : c->steps) {
    if (st.done) cudaEventDestroy(st.done);
    if (st.begun) cudaEventDestroy(st.begun);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int comm_exchange_counts(acg_comm* c, uint64_t mine, uint64_t* total, uint64_t* my_offset) {
  const int n = c->nranks;
#ifndef ACB_EMULATE
  c->h_counts[n] = mine;
  CKC(cudaMemcpyAsync(c->d_counts + n, c->h_counts + n, 8, cudaMemcpyHostToDevice, c->stream));
  CKN(nccl().AllGather(c->d_counts + n, c->d_counts, 1, kNcclUint64, static_cast<NcclComm>(c->nccl), c->stream));
  CKC(cudaMemcpyAsync(c->h_counts, c->d_counts, size_t(n) * 8, cudaMemcpyDeviceToHost, c->stream));
  CKC(cudaStreamSynchronize(c->stream));
#else
  Fabric* f = static_cast<Fabric*>(c->nccl);
  {
    std::unique_lock<std::mutex> lk(f->mu);
    f->slots[size_t(c->rank)] = mine;
    f->barrier(lk);
    for (int r = 0; r < n; ++r) c->h_counts[r] = f->slots[size_t(r)];
    f->barrier(lk);
  }
#endif
  uint64_t tot = 0, off = 0;
  for (int r = 0; r < n; ++r) {
    c->counts[size_t(r)] = c->h_counts[r];
    if (r < c->rank) off += c->h_counts[r];
    tot += c->h_counts[r];
  }
  *total = tot;
  *my_offset = off;
  return ACG_OK;
}

int comm_ensure_recv(acg_comm* c, uint64_t total) {
  if (total <= c->recv_cap) return ACG_OK;
  // every rank sees the same totals, so all of them take this branch together and agree on the size.
  // The buffer cannot move under a step that is still writing into it: the caller waits first.
  if (c->steps[0].active || c->steps[1].active) return ACG_E_OVERFLOW;
  const uint64_t cap = total + total / 8 + 1024;
  // two halves + the completion flags of the begin / wait form (8 bytes per rank, comm_flags)
  const size_t flag_bytes = (size_t(c->nranks) * 8 + 255) & ~size_t(255);
  const size_t bytes = 2 * size_t(cap) * sizeof(acg_match) + flag_bytes;
#ifndef ACB_EMULATE
  NcclComm comm = static_cast<NcclComm>(c->nccl);
  if (c->recv_peer) { cudaIpcCloseMemHandle(c->recv_peer); c->recv_peer = nullptr; }
  cudaIpcMemHandle_t handle;
  std::memset(&handle, 0, sizeof(handle));
  int ok = 1;
  if (c->rank == 0) {
    if (c->recv_own) { cudaFree(c->recv_own); c->recv_own = nullptr; }
    CKC(cudaMalloc(&c->recv_own, bytes));
    CKC(cudaMemset(c->recv_own + (bytes - flag_bytes), 0, flag_bytes));  // sequence numbers only grow
    if (c->transport == ACG_TRANSPORT_PEER && c->nranks > 1) {
      if (cudaIpcGetMemHandle(&handle, c->recv_own) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
  }
  if (c->nranks > 1 && c->transport == ACG_TRANSPORT_PEER) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (c->rank == 0) CKC(cudaMemcpyAsync(c->d_handle, &handle, 64, cudaMemcpyHostToDevice, c->stream));
    CKN(nccl().Broadcast(c->d_handle, c->d_handle, 64, kNcclUint8, 0, comm, c->stream));
    if (c->rank != 0) {
      CKC(cudaMemcpyAsync(&handle, c->d_handle, 64, cudaMemcpyDeviceToHost, c->stream));
      CKC(cudaStreamSynchronize(c->stream));
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      } else {
        c->recv_peer = static_cast<uint8_t*>(p);
      }
    }
    // all ranks must agree: one failed mapping sends everybody down the NCCL payload path
    int* flag = reinterpret_cast<int*>(c->h_counts + c->nranks);
    *flag = ok;
    CKC(cudaMemcpyAsync(c->d_counts + c->nranks, flag, 4, cudaMemcpyHostToDevice, c->stream));
    CKN(nccl().AllReduce(c->d_counts + c->nranks, c->d_counts + c->nranks, 1, kNcclInt32, kNcclMin, comm, c->stream));
    CKC(cudaMemcpyAsync(flag, c->d_counts + c->nranks, 4, cudaMemcpyDeviceToHost, c->stream));
    CKC(cudaStreamSynchronize(c->stream));
    if (*flag == 0) {
      if (c->recv_peer) { cudaIpcCloseMemHandle(c->recv_peer); c->recv_peer = nullptr; }
      c->transport = ACG_TRANSPORT_NCCL;
      if (c->rank == 0)
        std::fprintf(stderr, "acb200: peer mapping of the receive buffer unavailable; gathering with ncclSend/ncclRecv\n");
    }
  }
#else
  Fabric* f = static_cast<Fabric*>(c->nccl);
  if (c->rank == 0) {
    if (c->recv_own) { cudaFree(c->recv_own); c->recv_own = nullptr; }
    CKC(cudaMalloc(&c->recv_own, bytes));
    CKC(cudaMemset(c->recv_own + (bytes - flag_bytes), 0, flag_bytes));
  }
  {
    std::unique_lock<std::mutex> lk(f->mu);
    if (c->rank == 0) f->recv = c->recv_own;
    f->barrier(lk);
    if (c->rank != 0) c->recv_peer = f->recv;
    f->barrier(lk);
  }
#endif
  c->recv_cap = cap;
  return ACG_OK;
}

int comm_record_target(acg_comm* c, int slot, uint64_t my_offset, uint64_t mine, bool staged, uint8_t** target) {
  const bool direct = c->rank == 0 || (c->transport == ACG_TRANSPORT_PEER && !staged);
  if (direct) {
    uint8_t* base = c->rank == 0 ? c->recv_own : c->recv_peer;
    if (!base) return ACG_E_CUDA;
    *target = base + (size_t(slot) * size_t(c->recv_cap) + my_offset) * sizeof(acg_match);
    return ACG_OK;
  }
  if (c->transport == ACG_TRANSPORT_PEER && !c->recv_peer) return ACG_E_CUDA;
  if (mine > c->send_cap) {
    // (cudaFree waits for the device: a copy of the previous step that still reads the buffer ends first)
    if (c->send_buf) { cudaFree(c->send_buf); c->send_buf = nullptr; }
    const uint64_t cap = mine + mine / 8 + 1024;
    CKC(cudaMalloc(&c->send_buf, size_t(cap) * sizeof(acg_match)));
    c->send_cap = cap;
  }
  *target = c->send_buf;
  return ACG_OK;
}

int comm_enqueue_close(acg_comm* c, int slot, uint64_t my_offset, uint64_t mine, bool staged, uint64_t seq,
                       bool* flagged) {
  *flagged = false;
  if (staged && c->transport == ACG_TRANSPORT_PEER && c->nranks > 1) {
    if (c->rank != 0) {
      // copy-engine payload: the records expanded into the local staging buffer travel to their place
      // in rank 0's buffer as one device-to-device copy over NVLink -- no SM is held while they do --
      // and the step's sequence number follows them into this rank's flag word on the same stream
      if (mine) {
        uint8_t* dst = c->recv_peer + (size_t(slot) * size_t(c->recv_cap) + my_offset) * sizeof(acg_match);
        CKC(cudaMemcpyAsync(dst, c->send_buf, size_t(mine) * sizeof(acg_match), cudaMemcpyDeviceToDevice, c->stream));
      }
      c->h_flag_src[slot] = seq;
      CKC(cudaMemcpyAsync(c->d_flag_src + slot, c->h_flag_src + slot, 8, cudaMemcpyHostToDevice, c->stream));
      CKC(cudaMemcpyAsync(comm_flags(c) + c->rank, c->d_flag_src + slot, 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    *flagged = true;
    return ACG_OK;
  }
#ifndef ACB_EMULATE
  NcclComm comm = static_cast<NcclComm>(c->nccl);
  if (c->nranks > 1) {
    if (c->transport == ACG_TRANSPORT_NCCL) {
      CKN(nccl().GroupStart());
      if (c->rank == 0) {
        uint8_t* half = comm_half(c, slot);
        uint64_t off = c->counts[0];
        for (int r = 1; r < c->nranks; ++r) {
          if (c->counts[size_t(r)])
            CKN(nccl().Recv(half + off * sizeof(acg_match), size_t(c->counts[size_t(r)]) * sizeof(acg_match),
                            kNcclUint8, r, comm, c->stream));
          off += c->counts[size_t(r)];
        }
      } else if (mine) {
        CKN(nccl().Send(c->send_buf, size_t(mine) * sizeof(acg_match), kNcclUint8, 0, comm, c->stream));
      }
      CKN(nccl().GroupEnd());
    }
    // closing barrier: a kernel's peer stores are complete when it retires, and this collective is
    // ordered behind the expand kernel on every rank's stream
    CKN(nccl().AllGather(c->d_counts + c->nranks, c->d_counts, 1, kNcclUint64, comm, c->stream));
  }
#else
  (void)mine; (void)slot;
  Fabric* f = static_cast<Fabric*>(c->nccl);
  CKC(cudaStreamSynchronize(c->stream));
  std::unique_lock<std::mutex> lk(f->mu);
  f->barrier(lk);
#endif
  return ACG_OK;
}

int comm_wait_flags(acg_comm* c, uint64_t seq) {
  if (c->rank != 0 || c->nranks < 2) return ACG_OK;
  const uint64_t* flags = comm_flags(c);
  if (!flags) return ACG_E_CUDA;
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    CKC(cudaMemcpyAsync(c->h_flags, flags, size_t(c->nranks) * 8, cudaMemcpyDeviceToHost, c->poll));
    CKC(cudaStreamSynchronize(c->poll));
    bool all = true;
    for (int r = 1; r < c->nranks; ++r) all = all && c->h_flags[r] >= seq;
    if (all) return ACG_OK;
    // a peer that died must not hang this rank forever
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) {
      std::fprintf(stderr, "acb200: a rank's records of step %llu did not arrive within 60 s\n", (unsigned long long)seq);
      return ACG_E_CUDA;
    }
#ifdef ACB_EMULATE
    std::this_thread::yield();
#endif
  }
}

}  // namespace acb
