// Microbenchmark: issue rate of the integer multiplier instructions used by ff.cuh on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/imad_mb scripts/imad_microbench.cu
// Prints warp-instructions per cycle per SM sub-partition for several instruction mixes.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256

template <int MODE>
__global__ void __launch_bounds__(1024) k(uint64_t* out, uint32_t seed, int iters) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint64_t acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = a + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < REP / 8; r++) {
      if (MODE == 0) {   // 8 independent mad.wide (no carry)
#pragma unroll
        for (int i = 0; i < 8; i++)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[(i + 3) & 7]), "r"(b));
      } else if (MODE == 1) {   // two carry chains of 4: mul.wide + add.cc.u64 (fused IMAD.WIDE.X)
        uint64_t t[8];
#pragma unroll
        for (int i = 0; i < 8; i++) asm("mul.wide.u32 %0, %1, %2;" : "=l"(t[i]) : "r"((uint32_t)acc[(i + 3) & 7]), "r"(b));
        asm volatile("add.cc.u64 %0, %0, %1;" : "+l"(acc[0]) : "l"(t[0]));
        asm volatile("addc.cc.u64 %0, %0, %1;" : "+l"(acc[1]) : "l"(t[1]));
        asm volatile("addc.cc.u64 %0, %0, %1;" : "+l"(acc[2]) : "l"(t[2]));
        asm volatile("addc.u64 %0, %0, %1;" : "+l"(acc[3]) : "l"(t[3]));
        asm volatile("add.cc.u64 %0, %0, %1;" : "+l"(acc[4]) : "l"(t[4]));
        asm volatile("addc.cc.u64 %0, %0, %1;" : "+l"(acc[5]) : "l"(t[5]));
        asm volatile("addc.cc.u64 %0, %0, %1;" : "+l"(acc[6]) : "l"(t[6]));
        asm volatile("addc.u64 %0, %0, %1;" : "+l"(acc[7]) : "l"(t[7]));
      } else if (MODE == 2) {   // 32-bit mad.lo (IMAD), 8 independent
        uint32_t* w = reinterpret_cast<uint32_t*>(acc);
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(w[i]) : "r"(w[(i + 3) & 7]), "r"(b));
      } else if (MODE == 3) {   // mad.hi (IMAD.HI), 8 independent
        uint32_t* w = reinterpret_cast<uint32_t*>(acc);
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(w[i]) : "r"(w[(i + 3) & 7] | 0x80000000u), "r"(b | 0xf0000000u));
      } else if (MODE == 5) {   // DFMA, 8 independent chains
        double* w = reinterpret_cast<double*>(acc);
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(w[i]) : "d"(w[(i + 3) & 7]), "d"(1.0000001));
      } else if (MODE == 6) {   // 4 mad.wide + 4 DFMA interleaved (do the two pipes overlap?)  counts all 8
#pragma unroll
        for (int i = 0; i < 4; i++)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[(i + 1) & 3]), "r"(b));
        double* w = reinterpret_cast<double*>(acc + 4);
#pragma unroll
        for (int i = 0; i < 4; i++) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(w[i]) : "d"(w[(i + 1) & 3]), "d"(1.0000001));
      } else if (MODE == 4) {   // 4 mad.wide + 4 independent IADD3-class adds (mix like the field code)
#pragma unroll
        for (int i = 0; i < 4; i++)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[(i + 1) & 3]), "r"(b));
        uint32_t* w = reinterpret_cast<uint32_t*>(acc + 4);
#pragma unroll
        for (int i = 0; i < 4; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(w[(i + 1) & 3]));
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps_per_smsp) {
  int threads = 32 * 4 * warps_per_smsp;   // per block = per SM (1 block/SM)
  int blocks = 148;
  uint64_t* out;
  cudaMalloc(&out, sizeof(uint64_t) * blocks * threads);
  int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 1, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, 1, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double warp_instr_per_smsp = (double)iters * REP * warps_per_smsp;   // counted instructions of interest
  double cycles = ms * 1e-3 * clk_khz * 1e3;
  printf("%-44s warps/SMSP=%d  %.3f ms  -> %.3f counted warp-instr / cycle / SMSP (at %d MHz nominal)\n", name,
         warps_per_smsp, ms, warp_instr_per_smsp / cycles, clk_khz / 1000);
  cudaFree(out);
}

int main() {
  for (int w : {2, 4, 8}) {
    run<0>("IMAD.WIDE.U32 independent", w);
    run<1>("IMAD.WIDE.U32.X carry chains (2 x 4)", w);
    run<2>("IMAD (mad.lo) independent", w);
    run<3>("IMAD.HI independent", w);
    run<4>("4 IMAD.WIDE + 4 IADD (counts all 8)", w);
    run<5>("DFMA independent", w);
    run<6>("4 IMAD.WIDE + 4 DFMA (counts all 8)", w);
  }
  return 0;
}
