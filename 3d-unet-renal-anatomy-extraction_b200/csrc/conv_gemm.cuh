// Parameters of the tcgen05 "shifted GEMM" convolution kernel (conv_gemm.cu).
//
// One kernel covers every dense contraction of the U-Net forward / data-gradient path
// (network.py:394-403 Conv3d k3/k1, stride 1/2; network.py:312-314 ConvTranspose3d + pad; and the
// autograd data gradients of all of them) by expressing each as
//
//     out[v, n] = sum_{tap, c} A[v + shift(tap), c] * Wp[tap, c, n]
//
// on a "tile grid" (the output grid for ordinary convs, the coarse grid for strided / transposed
// ones).  The halo brick of A lives in shared memory once and every tap is just a different start
// address of the UMMA shared-memory descriptor (no-swizzle K-major layout), so an A byte is read
// from L2 ~1.4-1.9x instead of 27x.
#pragma once
#include "common.cuh"

namespace u3d {

constexpr int CG_HT = 16;            // output tile: 16 (h) x 8 (w) voxels = the 128 rows of one UMMA
constexpr int CG_WT = 8;
constexpr int CG_HB = CG_HT + 2;     // brick with a 1-voxel halo on both sides
constexpr int CG_WB = CG_WT + 2;
constexpr int CG_BOX_BYTES = CG_HB * CG_WB * 16;     // one (plane, 8-channel chunk) TMA box = 2880 B   (weight-gradient kernel)
constexpr int CG_CHUNK_PITCH = 2944;                 // padded to 128 B (TMA smem destination alignment) (weight-gradient kernel)
// conv_gemm A slabs: one TMA box per plane holds the whole channel group, (18 x 10 voxels) rows of 16 * G bytes
// (G = 2 / 4 / 8 chunks = 32 / 64 / 128 B), written with SWIZZLE_32B / 64B / 128B = the swizzled K-major UMMA layouts
// (row pitch = swizzle span, 8-row group pitch SBO = one brick line).  A 16-byte-row box costs one shared-memory write
// wavefront and one 32-byte L2 sector PER ROW: at level 0 the TMA writes then took as many shared-memory cycles as all
// tensor-core operand reads together (profiles/r01_notes.md); whole-group rows cut that 2-8x.
__host__ __device__ constexpr int cg_row_bytes(int G) { return 16 * G; }
__host__ __device__ constexpr int cg_plane_bytes(int G) { return CG_HB * CG_WB * 16 * G; }
__host__ __device__ constexpr int cg_plane_pitch(int G) { return (cg_plane_bytes(G) + 1023) / 1024 * 1024; }
constexpr int CG_W_STAGES = 16;
constexpr int CG_A_STAGES = 4;       // at most this many A slabs in flight
constexpr int CG_EPI_WARPS = 4;      // 4 or 8
constexpr int CG_THREADS = 96 + 32 * CG_EPI_WARPS;      // warp0 A-TMA, warp1 W-TMA, warp2 MMA (+TMEM alloc), then the epilogue warps
constexpr int CG_MAX_MAPS = 8;

struct ConvGemmParams {
  CUtensorMap amap[CG_MAX_MAPS];   // A sources (NDHWC bf16): concat halves or stride-2 parity views
  // device tables (int32), laid out by the host plan:
  //   [0, n_cg)                     map id of cgroup
  //   [n_cg, 2 n_cg)                first channel (element index within its source) of cgroup
  //   [2 n_cg, 2 n_cg + n_taps)     tap shift: sd | sh << 8 | sw << 16   (each in 0..2, brick coords)
  //   then n_nblk * n_cg            bitmask of active taps per (nblock, cgroup)
  //   then n_nblk                   first packed-weight tile index of nblock
  //   then n_nblk                   output channel offset of nblock | (output tensor index << 30)
  //   then n_nblk                   output coordinate offset: od | oh << 8 | ow << 16
  const int* tab;
  const bf16* w;          // packed weight tiles [G][fuse * nblk][8], in consumption order
  bf16* out;          // output tensor 0
  bf16* out2;         // output tensor 1 (data gradient of a channel concat), same strides / out_C
  const float* bias;      // [n_nblk * nblk] fp32 or null
  const bf16* addend;     // tensor with out's strides added in the epilogue, or null
  const bf16* addend2;    // addend for out2
  double* stats;          // [N][stats_C][2] (sum, sum of squares) accumulated with atomics, or null
  int* err;               // device error word
  int N, D, H, W;         // tile-grid extents
  int tiles_h, tiles_w, segs_d, Dt;
  int n_nblk, nblk;       // nblk in {32, 64, 96, 128}; Dt * nblk <= 256
  int G, n_cg, n_taps;    // G chunks (of 8 channels) per cgroup, G even
  int in_f16, out_f16;    // 16-bit storage of A / weights and of out / addend: 0 = bf16, 1 = fp16
  int wT, w_stages;       // weight ring: taps per stage (their tiles are contiguous) and ring depth (2..16)
  int a_stages;           // A slabs (channel groups) in flight: 2..4
  int nbuf;               // TMEM accumulator buffers: 2 (Dt*nblk <= 256, epilogue overlaps next item) or 1 (<= 512)
  int fuse;               // 1, or 3: a weight tile holds the d-taps 2,1,0 of one (kh,kw) and taps carry sd = 0
  long long out_sN, out_sD, out_sH, out_sW;   // element strides of out / addend
  int out_C;              // channels physically present in out (store mask)
  int stats_C;
  int omul;               // output coordinate = tile-grid coordinate * omul + offset(nblock)
  int zD, zH, zW;         // output planes forced to zero (ConvTranspose3d + ConstantPad3d), or -1
  int act;                // 0 none, 1 LeakyReLU(0.01) applied after bias / addend
  int dense;              // 1: fuse == 3, n_taps == 9 in (kh,kw) order, every tap of every (N block, group) active
  int n_work;
  long long* dbg_out;     // timing experiments only: CTA 0's MMA warp writes {total, acc_empty wait, a_full wait, w_full wait, items} cycles
  int dbg;                // timing experiments only (env U3D_DBG): 1 = load only chunk 0 of every A group, 2 = no statistics, 4 = no stores, 8 = no MMAs (dense path), 16 = no epilogue work
};

size_t conv_gemm_smem_bytes(int Dt, int G, int nblk, int fuse, int wT, int w_stages, int a_stages);
int conv_gemm_launch(const ConvGemmParams& p, int num_sms, cudaStream_t stream);

}  // namespace u3d
