// Internal (non-ABI) declarations shared by ntt.cu and ntt_api.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace bz {

#define NTT_TILE 2048         // elements per CTA tile: lanes V = NTT_TILE / R adjacent work items (>= 4, i.e. >= 128-byte runs)
#define NTT_MAX_PEERS 8

struct NttTables {
  const uint4* lo;     // w^x,            x < 2^min(lo_bits, log_root)      (Montgomery form)
  const uint4* hi;     // w^(x 2^lo_bits), x < 2^(log_root - lo_bits)
  const uint4* ninv;   // (2^log_root)^-1
  int lo_bits;
  int log_root;        // w is a primitive 2^log_root-th root of unity
};

// One global-memory pass.  Work item q in [0, Q) splits as q0 = q % Q0, q1 = (q / Q0) % Q1,
// q2 = q / (Q0 Q1); element r of its input is at in + q0 in_s0 + q1 in_s1 + q2 in_s2 + r in_sr and
// output k goes to out + q0 out_s0 + q1 out_s1 + q2 out_s2 + k out_sr (all in elements of 32 B).
struct NttPassParams {
  const uint4* in;
  uint4* out;
  // exchange pass (last column pass of a multi-GPU four-step): output k of a work item goes to peer
  // k / peer_rows at peer_out[peer] + item offset + (k % peer_rows) * out_sr; stores travel over NVLink
  uint4* peer_out[NTT_MAX_PEERS];
  uint32_t peer_rows;               // 0 = not an exchange pass; else a power of two
  int lpeer_rows;
  int lr;                           // log2 of the pass radix R (1..9)
  int lv;                           // log2 of the lanes per CTA, V = max(4, NTT_TILE / R) (set by ntt_launch_pass)
  uint64_t Q;
  uint64_t Q0, Q1;                  // powers of two
  int lq0, lq1;                     // their logs (index split by shift/mask, no 64-bit division)
  uint64_t in_s0, in_s1, in_s2, in_sr;
  uint64_t out_s0, out_s1, out_s2, out_sr;
  int tw_sel;                       // input twiddle w^(r * q[tw_sel] * tw_scale); -1 = none
  uint64_t tw_scale;
  const uint4* tw_full;             // optional: the same twiddles precomputed in the pass' INPUT layout (one per
                                    // element, coalesced like the data) -- saves the two-level table product
  uint4* tw_full_out;               // table-generation mode: write the twiddle of every element here and return
  // output twiddle w^((row * col mod 2^log_root) * otw_scale), row = q[otw_rsel] * otw_ra + k * otw_rb,
  // col = otw_base + q0 (the four-step twiddle between the column and the row transforms); otw_rsel < 0 = none
  int otw_rsel;
  uint64_t otw_ra, otw_rb, otw_base, otw_scale;
  int store_k_fastest;              // store loop order (k fastest when a work item's outputs are contiguous)
  int scale_ninv;                   // multiply outputs by (2^log_root)^-1 (last pass of an inverse transform)
  int* err;                         // first pass of a client transform: set to 1 when an input element is not canonical
                                    // (>= r); null = no check
  NttTables tab;
};

void ntt_gen_tables(int field, NttTables& t, int log_root, int inverse, cudaStream_t st);
cudaError_t ntt_launch_pass(int field, const NttPassParams& P, cudaStream_t st);
int ntt_two_adicity(int field);

}  // namespace bz
