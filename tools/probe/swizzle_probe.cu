// Hardware probe (not part of the product): does a UMMA shared-memory descriptor with a SWIZZLED K-major layout read the
// right rows when its start address is shifted by an arbitrary number of rows inside a TMA-written brick, and with a
// line pitch (SBO) that is not a multiple of the swizzle atom?  D = A * I is read back and compared with the rows
// the shifted window should contain.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o swizzle_probe swizzle_probe.cu
#include "../../3d-unet-renal-anatomy-extraction_b200/csrc/common.cuh"
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
using namespace u3d;

struct Params {
  CUtensorMap map;
  float* out;       // [128][16]
  int row_bytes;    // 32 / 64 / 128
  int layout;       // 6 = SW32, 4 = SW64, 2 = SW128
  int shift_rows;   // start shift in rows
  int kbyte;        // K offset inside the row (bytes)
  int base_off;     // descriptor base_offset field
  int wb, hb;       // brick lines: hb lines of wb rows
};

__global__ void probe(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, mbar;
  __shared__ uint32_t tmem_base;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 64 * 1024;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&mbar), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_base), 32); tmem_relinquish(); }
  // B = identity 16x16, no-swizzle K-major: [chunk 2][row 16][8 elems]
  bf16* b = reinterpret_cast<bf16*>(smem + (b_smem - smem_u32(smem)));
  for (int i = threadIdx.x; i < 2 * 16 * 8; i += blockDim.x) {
    const int chunk = i / 128, n = (i / 8) % 16, k = chunk * 8 + i % 8;
    b[i] = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar), (uint32_t)p.wb * p.hb * p.row_bytes);
    tma_load_5d(a_smem, &p.map, smem_u32(&bar), 0, 0, 0, 0, 0);
  }
  volatile int abort_flag = 0;
  mbar_wait(smem_u32(&bar), 0, &abort_flag, nullptr, 0);
  tc_fence_after();
  if (threadIdx.x < 32 && elect_one()) {
    uint64_t ad = 0;
    const uint32_t start = a_smem + p.shift_rows * p.row_bytes + p.kbyte;
    ad |= (uint64_t)((start >> 4) & 0x3FFF);
    ad |= (uint64_t)1 << 16;                                            // LBO (ignored for swizzled K-major)
    ad |= (uint64_t)(((uint32_t)p.wb * p.row_bytes >> 4) & 0x3FFF) << 32; // SBO = one brick line
    ad |= (uint64_t)1 << 46;
    ad |= (uint64_t)(p.base_off & 7) << 49;
    ad |= (uint64_t)(p.layout & 7) << 61;
    const uint64_t bd = umma_desc(b_smem, 16 * 16, 128);
    tc_mma_bf16(tmem_base, ad, bd, umma_idesc_bf16(128, 16, 0, 0), 0);
    tc_commit(smem_u32(&mbar));
  }
  mbar_wait(smem_u32(&mbar), 0, &abort_flag, nullptr, 0);
  tc_fence_after();
  if (threadIdx.x < 128) {
    const int q = threadIdx.x >> 5;
    uint32_t v[32];
    // 16 columns: use x32 load over 32 allocated columns, keep 16
    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) p.out[threadIdx.x * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 32);
}

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fnp);
  const int WB = 10, HB = 18, HG = 40, WG = 10;
  int n_fail_total = 0;
  for (int mode = 0; mode < 3; ++mode) {
    const int C = mode == 0 ? 16 : (mode == 1 ? 32 : 64);
    const int row_bytes = C * 2;
    const int layout = mode == 0 ? 6 : (mode == 1 ? 4 : 2);
    const CUtensorMapSwizzle sw = mode == 0 ? CU_TENSOR_MAP_SWIZZLE_32B : (mode == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    const int rows = HG * WG;
    std::vector<__nv_bfloat16> h(rows * C);
    // value = (row % 64) * 4 + (channel / 16 % 4): exact in bf16 (<= 255); channel-in-16 tested by the identity columns
    for (int r = 0; r < rows; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = __float2bfloat16((float)((r % 61) * 4 + (c / 16) % 4) * ((c % 16) == (r % 16) ? 1.f : 0.5f));
    __nv_bfloat16* dx; float* dout;
    cudaMalloc(&dx, h.size() * 2); cudaMalloc(&dout, 128 * 16 * 4);
    cudaMemcpy(dx, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    Params p;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)WG, (cuuint64_t)HG, 1, 1};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)WG * C * 2, (cuuint64_t)HG * WG * C * 2, (cuuint64_t)HG * WG * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)C, WB, HB, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    p.out = dout; p.row_bytes = row_bytes; p.layout = layout; p.wb = WB; p.hb = HB;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int shifts[] = {0, 1, 2, 8, 11, 12, 22, 7};
    for (int bo_mode = 0; bo_mode < 2; ++bo_mode)
      for (int kb = 0; kb < row_bytes; kb += 32)
        for (int si = 0; si < 8; ++si) {
          p.shift_rows = shifts[si]; p.kbyte = kb;
          // base_offset variant: 0, or (address >> 7) & 7 relative to the 1024-aligned brick base
          p.base_off = bo_mode == 0 ? 0 : (((shifts[si] * row_bytes + kb) >> 7) & 7);
          probe<<<1, 128, 100 * 1024>>>(p);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("mode %d shift %d kb %d: CUDA error %s\n", mode, shifts[si], kb, cudaGetErrorString(e)); return 2; }
          std::vector<float> o(128 * 16);
          cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
          int bad = 0, first_bad = -1;
          for (int m = 0; m < 128; ++m) {
            const int line = m / 8, w = m % 8;
            const int row = line * WB + w + shifts[si];
            for (int n = 0; n < 16; ++n) {
              const int c = kb / 2 + n;
              const float want = (float)((row % 61) * 4 + (c / 16) % 4) * ((c % 16) == (row % 16) ? 1.f : 0.5f);
              if (o[m * 16 + n] != want) { if (first_bad < 0) first_bad = m * 16 + n; ++bad; }
            }
          }
          printf("SW%-3d base_off=%s shift=%2d kbyte=%3d : %s (%d bad%s)\n", row_bytes, bo_mode ? "addr" : "0   ", shifts[si], kb,
                 bad ? "MISMATCH" : "ok", bad, "");
          if (bad && si < 3 && kb == 0) {
            printf("    first bad at m=%d n=%d got %.1f; row0 got:", first_bad / 16, first_bad % 16, o[first_bad]);
            for (int n = 0; n < 16; ++n) printf(" %.1f", o[n]);
            printf("\n    m=1 got:");
            for (int n = 0; n < 16; ++n) printf(" %.1f", o[16 + n]);
            printf("\n");
          }
          n_fail_total += bad != 0;
        }
    cudaFree(dx); cudaFree(dout);
  }
  printf("probe done, %d failing configurations\n", n_fail_total);
  return 0;
}
