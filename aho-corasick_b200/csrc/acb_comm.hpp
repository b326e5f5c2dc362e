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
  uint8_t* recv_own = nullptr;     // 2 * recv_cap records: two halves, one per step in flight
  uint8_t* recv_peer = nullptr;
  uint64_t recv_cap = 0;           // records per half; identical on every rank
  // nccl transport: local staging of this rank's records
  uint8_t* send_buf = nullptr;
  uint64_t send_cap = 0;
  unsigned long long* d_counts = nullptr;  // [nranks + 1]: slot nranks = this rank's count
  unsigned long long* h_counts = nullptr;  // pinned mirror
  uint8_t* d_handle = nullptr;             // 64-byte cudaIpcMemHandle_t in transit
  std::vector<uint64_t> counts;            // last call: records per rank
  uint8_t* h_view = nullptr;               // pinned host copy of the gathered records (acg_comm_fetch_view)
  uint64_t h_view_cap = 0;                 // records
  // Steps in flight (acg_find_overlapping_sharded_begin / _wait): at most two, slot = step parity.
  // The scan of step k + 1 overlaps the transfer of step k's records into rank 0's buffer.
  struct Step {
    bool active = false;
    void* lease = nullptr;                 // the workspace of the handle whose tuples the expand kernel reads
    const void* dfa = nullptr;
    uint64_t mine = 0, total = 0;
    cudaEvent_t done = nullptr;            // behind the closing barrier of the step on `stream`
    cudaEvent_t begun = nullptr;           // before the count exchange of the step
    float scan_ms = 0, order_ms = 0;
    uint64_t seq = 0;                      // 1-based sequence number of the step
    bool flagged = false;                  // closed by completion flags instead of the NCCL barrier
    uint64_t candidates = 0;
    int launches = 0;
  } steps[2];
  uint64_t step_seq = 0;                   // steps begun
  const uint8_t* last_result = nullptr;    // rank 0: records of the step that was waited for last
  uint64_t last_total = 0;
  float last_gather_ms = 0;
  cudaEvent_t mark[2] = {nullptr, nullptr};  // acg_comm_mark: device timestamps around a stream of steps
  // Completion flags of the begin / wait form (peer transport): one 64-bit word per rank behind the two
  // halves of rank 0's buffer.  A rank's copy engine writes the step's sequence number there right
  // behind its records; rank 0 polls the words with small device-to-host copies.  No kernel takes part,
  // so nothing sits on an SM waiting for a peer while the next step's scan wants that SM.
  uint64_t* h_flag_src = nullptr;   // pinned, [2]: the value in transit, by step parity
  uint64_t* d_flag_src = nullptr;   // device, [2]
  uint64_t* h_flags = nullptr;      // pinned, [nranks]: rank 0's poll target
  cudaStream_t poll = nullptr;
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
// Where this rank's expand kernel writes record `my_offset` of half `slot` (peer transport), or its
// local staging buffer (nccl transport, or peer transport with `staged`; comm_enqueue_close then moves
// it: ncclSend / a copy-engine copy into the mapped buffer).  `staged` is what a stream of steps uses:
// the scan of the next step owns the SMs while the records travel.
int comm_record_target(acg_comm* c, int slot, uint64_t my_offset, uint64_t mine, bool staged, uint8_t** target);
// After the expand kernel was enqueued on c->stream: enqueue the payload transfer if the transport
// needs one and the closing barrier; nothing is waited for (the step's `done` event is recorded by
// the caller behind it).
int comm_enqueue_close(acg_comm* c, int slot, uint64_t my_offset, uint64_t mine, bool staged, uint64_t seq,
                       bool* flagged);
// Rank 0, step closed by flags: wait until every rank's flag word has reached `seq`.
int comm_wait_flags(acg_comm* c, uint64_t seq);
// The flag words (rank 0: own memory; others: the mapping), behind the two halves.
inline uint64_t* comm_flags(acg_comm* c) {
  uint8_t* base = c->rank == 0 ? c->recv_own : c->recv_peer;
  return base ? reinterpret_cast<uint64_t*>(base + 2 * size_t(c->recv_cap) * 24) : nullptr;
}
// Rank 0's half `slot` of the receive buffer.
inline uint8_t* comm_half(acg_comm* c, int slot) { return c->recv_own + size_t(slot) * size_t(c->recv_cap) * 24; }

}  // namespace acb
