// Common macros and the carry-chain instruction layer.
//
// Device build: thin wrappers over the PTX extended-precision integer instructions
// (add.cc / addc / mad.lo.cc / madc.hi.cc ...); ptxas fuses the lo/hi pairs of a chain
// into IMAD.WIDE.U32(.X) on sm_100a.
// Host build (g++, used by tests/host_ff_check.cpp only): the same entry points emulated
// with a thread-local carry flag so that the exact limb schedules in ff.cuh / ec.cuh can be
// checked bit-for-bit on a machine without a GPU.  The host build is a test vehicle, not a
// CPU fallback: nothing in the product library calls field code on the host.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
// nvcc: field / curve code is device-only (the product never runs it on the host)
#define BZ_HDI __device__ __forceinline__
#define BZ_DI __device__ __forceinline__
#else
#define BZ_HDI inline
#define BZ_DI inline
#endif

namespace bz {
namespace cc {

#ifdef __CUDACC__
BZ_DI uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
BZ_DI uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
BZ_DI uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
BZ_DI uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
BZ_DI uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
BZ_DI uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
BZ_DI uint32_t mul_lo(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
BZ_DI uint32_t mul_hi(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
BZ_DI uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
BZ_DI uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
BZ_DI uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
BZ_DI uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
BZ_DI uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
BZ_DI uint64_t mul_wide(uint32_t a, uint32_t b) {
  uint64_t r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
}
BZ_DI uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) {
  uint64_t r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
BZ_DI uint64_t add_cc64(uint64_t a, uint64_t b) {
  uint64_t r; asm volatile("add.cc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
BZ_DI uint64_t addc_cc64(uint64_t a, uint64_t b) {
  uint64_t r; asm volatile("addc.cc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
BZ_DI uint64_t addc64(uint64_t a, uint64_t b) {
  uint64_t r; asm volatile("addc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
// m * (2^32 - 1) = (m << 32) - m without the multiplier (asm: the compiler would turn the C form back into a product)
BZ_DI uint64_t mul_2p32m1(uint32_t m) {
  uint64_t r;
  asm("{\n\t.reg .u32 lo, hi;\n\tsub.cc.u32 lo, 0, %1;\n\tsubc.u32 hi, %1, 0;\n\tmov.b64 %0, {lo, hi};\n\t}" : "=l"(r) : "r"(m));
  return r;
}
// a + (CC << 32): consume the carry flag into the HIGH 32-bit half of a 64-bit word (one IADD3.X)
BZ_DI uint64_t addc_hi32(uint64_t a) {
  uint64_t r;
  asm volatile("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\taddc.u32 hi, hi, 0;\n\tmov.b64 %0, {lo, hi};\n\t}"
               : "=l"(r) : "l"(a));
  return r;
}
#else
// ---- host emulation (test vehicle) ----
inline uint32_t& flag() { static thread_local uint32_t f = 0; return f; }
inline uint32_t add3(uint32_t a, uint32_t b, uint32_t cin, bool set) {
  uint64_t t = (uint64_t)a + b + cin;
  if (set) flag() = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t sub3(uint32_t a, uint32_t b, uint32_t bin, bool set) {
  uint64_t t = (uint64_t)a - b - bin;
  if (set) flag() = (uint32_t)((t >> 32) & 1);   // borrow
  return (uint32_t)t;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { return add3(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return add3(a, b, flag(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return add3(a, b, flag(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return sub3(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return sub3(a, b, flag(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return sub3(a, b, flag(), false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_lo(a, b), c, 0, true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_lo(a, b), c, flag(), true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, flag(), true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, flag(), false); }
inline uint64_t mul_wide(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
inline uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) { return (uint64_t)a * b + c; }
inline uint64_t add3_64(uint64_t a, uint64_t b, uint32_t cin, bool set) {
  unsigned __int128 t = (unsigned __int128)a + b + cin;
  if (set) flag() = (uint32_t)(t >> 64);
  return (uint64_t)t;
}
inline uint64_t add_cc64(uint64_t a, uint64_t b) { return add3_64(a, b, 0, true); }
inline uint64_t addc_cc64(uint64_t a, uint64_t b) { return add3_64(a, b, flag(), true); }
inline uint64_t addc64(uint64_t a, uint64_t b) { return add3_64(a, b, flag(), false); }
inline uint64_t addc_hi32(uint64_t a) { return a + ((uint64_t)flag() << 32); }
inline uint64_t mul_2p32m1(uint32_t m) { return ((uint64_t)m << 32) - (uint64_t)m; }
#endif

}  // namespace cc
}  // namespace bz
