// acb_comm.hpp -- multi-GPU plumbing behind the C ABI: one process per GPU,
// haystack slices, and the gather of match buffers to rank 0.
//
// Transport of the gather, chosen once per communicator:
//   peer   -- rank 0 owns the receive buffer and exports it with cudaIpc; every rank maps it and its
//             expand kernel stores the acg_match records of its slice straight into rank 0's HBM over
//             NVLink / NVSwitch at its offset of the global list (fused compute + collective: the
//             24-byte records are produced by the kernel that ships them, no staging copy, no NCCL
//             payload call).  NCCL carries only the 8-byte counts and the closing barrier.
//   nccl   -- fallback when the peer mapping is unavailable (no P2P between the devices, IPC refused):
//             local expand, then grouped ncclSend / ncclRecv of exactly the bytes each rank holds.
// NCCL is loaded with dlopen at first use so that the library shares whatever libnccl.so.2 the host
// process already carries (e.g. the one bundled with torch) instead of forcing a second copy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

struct acg_comm {
  int rank = 0, nranks = 1, device = -1;
  void* nccl = nullptr;            // ncclComm_t
  cudaStream_t stream = nullptr;   // NCCL calls + expand kernels of the sharded search
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int transport = 0;               // ACG_TRANSPORT_*
  // receive buffer: owned by rank 0 (recv_own), mapped by the others (recv_peer)
  uint8_t* recv_own = nullptr;
  uint8_t* recv_peer = nullptr;
  uint64_t recv_cap = 0;           // records; identical on every rank
  // nccl transport: local staging of this rank's records
  uint8_t* send_buf = nullptr;
  uint64_t send_cap = 0;
  unsigned long long* d_counts = nullptr;  // [nranks + 1]: slot nranks = this rank's count
  unsigned long long* h_counts = nullptr;  // pinned mirror
  uint8_t* d_handle = nullptr;             // 64-byte cudaIpcMemHandle_t in transit
  std::vector<uint64_t> counts;            // last call: records per rank
  uint8_t* h_view = nullptr;               // pinned host copy of the gathered records (acg_comm_fetch_view)
  uint64_t h_view_cap = 0;                 // records
  float last_gather_ms = 0;
};

namespace acb {

// 0 on success, ACG_E_* otherwise (ACG_E_CUDA for NCCL failures; the text goes to stderr).
int comm_unique_id(uint8_t* id128);
int comm_create(const uint8_t* id128, int rank, int nranks, acg_comm** out);
void comm_destroy(acg_comm* c);
// All ranks: publish `mine` records, learn everybody's; returns total and this rank's offset.
int comm_exchange_counts(acg_comm* c, uint64_t mine, uint64_t* total, uint64_t* my_offset);
// All ranks: make the receive buffer hold at least `total` records (collective when it must grow).
int comm_ensure_recv(acg_comm* c, uint64_t total);
// Where this rank's expand kernel writes record `my_offset` (peer transport), or its local staging
// buffer (nccl transport; comm_ship_records then moves it).
int comm_record_target(acg_comm* c, uint64_t my_offset, uint64_t mine, uint8_t** target);
// After the expand kernel was enqueued on c->stream: move the payload if the transport needs it and
// close the step (every rank's records are in rank 0's buffer when this returns).
int comm_finish_gather(acg_comm* c, uint64_t my_offset, uint64_t mine);

}  // namespace acb
