// tests/emu/acb_ptx_emu.cuh -- CPU dry-run counterpart of aho-corasick_b200/csrc/acb_ptx.cuh
// (selected through ACB_PTX_HEADER by tests/emu/build_emu.py).  Same names, plain C++: shared
// addresses are offsets into the emulated dynamic shared memory, a bulk copy completes at once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ACB_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__)
#define ACB_DYNAMIC_SMEM(name) unsigned char* name = emu::g_dyn_smem

namespace acb {
namespace ptx {

inline unsigned char* smem_ptr(uint32_t a, size_t bytes) {
  if ((size_t)a + bytes > emu::g_dyn_smem_bytes) emu::die("shared-memory access out of range");
  return emu::g_dyn_smem + a;
}
inline uint32_t smem_addr(const void* p) {
  const size_t off = (size_t)((const unsigned char*)p - emu::g_dyn_smem);
  if (off >= emu::g_dyn_smem_bytes) emu::die("smem_addr of a pointer outside dynamic shared memory");
  return (uint32_t)off;
}
// mbarrier word = number of completed phases; with one arrival + transaction bytes per phase the
// phase completes when the bulk copy lands, which in this model is immediately
inline void mbar_init(uint32_t bar, uint32_t) { *reinterpret_cast<uint64_t*>(smem_ptr(bar, 8)) = 0; }
inline void mbar_init_fence() {}
inline void fence_proxy_async() {}
inline void mbar_arrive_expect_tx(uint32_t bar, uint32_t) { (void)smem_ptr(bar, 8); }
inline bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  const uint64_t done = *reinterpret_cast<uint64_t*>(smem_ptr(bar, 8));
  if ((done & 1) != (parity & 1)) return true;  // the phase with this parity has completed
  emu::yield("mbarrier wait");
  return false;
}
inline void tma_load_1d(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
  if ((bytes & 15) || (smem_dst & 15) || ((uintptr_t)gmem_src & 15)) emu::die("bulk copy needs 16-byte alignment and size");
  std::memcpy(smem_ptr(smem_dst, bytes), gmem_src, bytes);
  ++*reinterpret_cast<uint64_t*>(smem_ptr(bar, 8));
  emu::g_cta->progress++;
}
inline uint4 lds128(uint32_t a) {
  if (a & 15) emu::die("misaligned 16-byte shared load");
  uint4 v;
  std::memcpy(&v, smem_ptr(a, 16), 16);
  return v;
}
inline uint32_t lds32(uint32_t a) {
  if (a & 3) emu::die("misaligned 4-byte shared load");
  uint32_t v;
  std::memcpy(&v, smem_ptr(a, 4), 4);
  return v;
}
inline uint32_t lds8(uint32_t a) { return *smem_ptr(a, 1); }
inline uint32_t atoms_add(uint32_t a, uint32_t v) {
  uint32_t* p = reinterpret_cast<uint32_t*>(smem_ptr(a, 4));
  const uint32_t old = *p;
  *p = old + v;
  return old;
}
inline uint64_t atoms_add64(uint32_t a, uint64_t v) {
  uint64_t* p = reinterpret_cast<uint64_t*>(smem_ptr(a, 8));
  const uint64_t old = *p;
  *p = old + v;
  return old;
}
inline void atoms_exch64(uint32_t a, uint64_t v) { *reinterpret_cast<uint64_t*>(smem_ptr(a, 8)) = v; }
inline uint32_t lds32_volatile(uint32_t a) {
  emu::yield_poll("spin on a 32-bit shared word");   // a spin on this value must let the thread that changes it run
  return *reinterpret_cast<uint32_t*>(smem_ptr(a, 4));
}
inline void sts32_volatile(uint32_t a, uint32_t v) { *reinterpret_cast<uint32_t*>(smem_ptr(a, 4)) = v; }
inline uint64_t lds64_volatile(uint32_t a) {
  emu::yield_poll("spin on a 64-bit shared word");   // a spin on this value must let the thread that changes it run
  return *reinterpret_cast<uint64_t*>(smem_ptr(a, 8));
}
inline void keep_in_registers(uint32_t&, uint32_t&, uint32_t&) {}
inline uint4 ld_nc_u4(const void* p) {
  if ((uintptr_t)p & 15) emu::die("misaligned 16-byte global load");
  uint4 v;
  std::memcpy(&v, p, 16);
  return v;
}

}  // namespace ptx
}  // namespace acb
