// Hardware probe (not part of the product): sustained cycles per tcgen05.mma (M=128, K=16, kind::f16) issued back to back
// by one thread with both operands resident in shared memory, as a function of N, of the accumulator pattern (same
// columns / sliding window of columns like the d-tap-fused convolution) and of the A layout (no-swizzle K-major with the
// conv kernel's brick strides, or the same bytes laid out for SWIZZLE_32B/64B/128B).  Nothing is loaded while timing:
// this is the tensor-pipe + operand-read floor the kernels are measured against.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on one B200.
#include "../../3d-unet-renal-anatomy-extraction_b200/csrc/common.cuh"
#include <stdio.h>
#include <stdlib.h>
using namespace u3d;

struct Params {
  long long* cycles;   // [grid]
  int n;               // MMA width
  int reps;            // MMAs timed
  int pattern;         // 0: same accumulator; 1: sliding (col offset cycles 0, n/3, 2n/3 ...); 2: alternate two accumulators
  int a_mode;          // 0: no-swizzle K-major, LBO = 2944, SBO = 160 (conv brick); 1: no-swizzle canonical (LBO = 2048... SBO = 128);
                       // 2/4/6: swizzle 128/64/32 B, SBO = 10 rows
  int a_step;          // A start address advances by this many 16-byte units per MMA (tap / plane walk), modulo 64
  int b_mn;            // 1: both operands MN-major (weight-gradient style descriptors)
};

__global__ void __launch_bounds__(128, 1) rate(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 96 * 1024;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar2), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_base), 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  volatile int abort_flag = 0;
  if (threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc_bf16(128, p.n, p.b_mn, p.b_mn, 1, 1);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      uint64_t ad0 = 0, bd = 0;
      if (p.b_mn) {
        ad0 = umma_desc(a_smem, 160, 2944);
        bd = umma_desc(b_smem, 128, 2048);
      } else if (p.a_mode == 0) {
        ad0 = umma_desc(a_smem, 2944, 160);
        bd = umma_desc(b_smem, (uint32_t)p.n * 16, 128);
      } else if (p.a_mode == 1) {
        ad0 = umma_desc(a_smem, 2048, 128);
        bd = umma_desc(b_smem, (uint32_t)p.n * 16, 128);
      } else {
        const uint32_t row = p.a_mode == 2 ? 128 : (p.a_mode == 4 ? 64 : 32);
        ad0 |= (uint64_t)((a_smem >> 4) & 0x3FFF);
        ad0 |= (uint64_t)1 << 16;
        ad0 |= (uint64_t)((10 * row >> 4) & 0x3FFF) << 32;
        ad0 |= (uint64_t)1 << 46;
        ad0 |= (uint64_t)(p.a_mode & 7) << 61;
        bd = umma_desc(b_smem, (uint32_t)p.n * 16, 128);
      }
      // warm up
      for (int i = 0; i < 32; ++i) tc_mma_bf16(tmem_base, ad0, bd, idesc, i ? 1u : 0u);
      tc_commit(smem_u32(&bar));
      while (!mbar_try_wait(smem_u32(&bar), 0)) {}
      t0 = clock64();
      const int third = p.n >= 96 ? p.n / 3 : p.n;
      uint32_t walk = 0;
      if (p.pattern >= 3) {
        // the d-tap-fused convolution's sequence for Dt = 8: per tap the widths n/3, 2n/3, n x 6, 2n/3, n/3 on a sliding
        // accumulator window; pattern 4 issues the same ten MMAs all at full width n (what a fixed instruction
        // descriptor would cost); pattern 5 adds a commit every 90 MMAs like the kernel's slab / ring hand-back
        const uint32_t i1 = umma_idesc_bf16(128, p.n / 3, 0, 0, 1, 1), i2 = umma_idesc_bf16(128, 2 * p.n / 3, 0, 0, 1, 1);
        for (int i = 0; i < p.reps; i += 10) {
#pragma unroll
          for (int pl = 0; pl < 10; ++pl) {
            const int hi = pl < 2 ? pl : 2, lo = pl - 7 > 0 ? pl - 7 : 0, cnt = hi - lo + 1;
            const uint32_t id = (p.pattern == 4 || cnt == 3) ? idesc : (cnt == 2 ? i2 : i1);
            tc_mma_bf16(tmem_base + (uint32_t)((pl - hi) * third), ad0 + (uint32_t)(pl * 6), bd, id, 1u);
          }
          if (p.pattern == 5 && (i % 90) == 80) tc_commit(smem_u32(&bar2));
        }
      } else
      for (int i = 0; i < p.reps; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t acc = tmem_base;
          if (p.pattern == 1) acc += (uint32_t)(((i + j) % 6) * third);        // slides like out planes pl-2..pl
          if (p.pattern == 2) acc += (uint32_t)(((i + j) & 1) * 256);
          tc_mma_bf16(acc, ad0 + walk, bd, idesc, 1u);
          walk = (walk + (uint32_t)p.a_step) & 63u;
        }
      }
      tc_commit(smem_u32(&bar));
      while (!mbar_try_wait(smem_u32(&bar), 1)) {}
      t1 = clock64();
      p.cycles[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  (void)abort_flag;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 4500;
  struct Case { int n, pattern, a_mode, a_step, b_mn; const char* what; };
  const Case cases[] = {
      {32, 0, 0, 0, 0, "N=32  same acc, conv brick A"},   {64, 0, 0, 0, 0, "N=64  same acc, conv brick A"},
      {96, 0, 0, 0, 0, "N=96  same acc, conv brick A"},   {128, 0, 0, 0, 0, "N=128 same acc, conv brick A"},
      {192, 0, 0, 0, 0, "N=192 same acc, conv brick A"},  {256, 0, 0, 0, 0, "N=256 same acc, conv brick A"},
      {96, 1, 0, 0, 0, "N=96  sliding acc, conv brick A"}, {192, 1, 0, 0, 0, "N=192 sliding acc, conv brick A"},
      {96, 2, 0, 0, 0, "N=96  two accs alternating"},
      {96, 1, 0, 1, 0, "N=96  sliding acc, A start walks +16 B"}, {96, 1, 0, 10, 0, "N=96  sliding acc, A start walks +160 B"},
      {96, 0, 1, 0, 0, "N=96  same acc, canonical no-swizzle A (SBO 128)"},
      {96, 0, 2, 0, 0, "N=96  same acc, SW128 A"}, {96, 0, 4, 0, 0, "N=96  same acc, SW64 A"}, {96, 0, 6, 0, 0, "N=96  same acc, SW32 A"},
      {96, 1, 4, 4, 0, "N=96  sliding acc, SW64 A, start walks one row"},
      {192, 0, 2, 0, 0, "N=192 same acc, SW128 A"}, {256, 0, 2, 0, 0, "N=256 same acc, SW128 A"},
      {96, 3, 0, 0, 0, "conv tap sequence 32,64,96x6,64,32 (Dt=8)"}, {96, 4, 0, 0, 0, "same ten MMAs, all N=96"},
      {96, 5, 0, 0, 0, "conv tap sequence + commit every 90 MMAs"}, {192, 3, 0, 0, 0, "conv tap sequence 64,128,192x6,128,64"},
      {32, 0, 0, 0, 1, "N=32  MN-major A and B (wgrad)"}, {64, 0, 0, 0, 1, "N=64  MN-major A and B (wgrad)"},
      {128, 0, 0, 0, 1, "N=128 MN-major A and B (wgrad)"},
  };
  for (int grid : {1, 148}) {
    printf("---- %d CTA(s), %d MMAs (M=128, K=16) each; floor = N/2 cycles, operand read model = 32 + N/4\n", grid, reps);
    for (const Case& c : cases) {
      Params p{d, c.n, reps, c.pattern, c.a_mode, c.a_step, c.b_mn};
      rate<<<grid, 128, 200 * 1024>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%-52s : CUDA error %s\n", c.what, cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double mx = 0, mn = 1e30;
      for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
      printf("%-52s : %6.1f cycles/MMA (min CTA %6.1f)   floor %5.1f  model %5.1f\n", c.what, mx / reps, mn / reps, c.n / 2.0,
             32.0 + c.n / 4.0);
    }
  }
  return 0;
}
