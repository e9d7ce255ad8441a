// Weight-gradient kernel on tcgen05 tensor cores (wgrad_gemm.cu).
//
//   dW[tap][cin][cout] = sum over voxels v of  x[v + shift(tap)][cin] * dy[v][cout]
//
// Both operands are consumed "MN-major" straight from the NDHWC bricks TMA drops into shared memory:
// the contraction index is the voxel (K = 16 voxels per MMA = two 8-voxel w-lines), M = 128 rows =
// 16 consecutive 8-channel chunks of the x brick (which may run across planes, giving the kd taps
// for free when Cin is small), N = the dy channels.  Every tap is again just a different start
// address.  A CTA owns one "job" (an M block, an N block and up to 512 TMEM columns worth of taps)
// and a 1/split share of the voxel tiles; partial sums are added to dW with fp32 atomics.
#pragma once
#include "common.cuh"
#include "conv_gemm.cuh"

namespace u3d {

constexpr int WG_MAX_MAPS = 16;
constexpr int WG_THREADS = 192;        // warp0 TMA, warp1 MMA (+TMEM), warps 2-5 epilogue
constexpr int WG_DY_BOX_BYTES = CG_HT * CG_WT * 16;   // 2048
constexpr int WG_MAX_G = 32;
// job table layout (int32), job j at tab + j * job_stride
constexpr int WG_J_DT = 0, WG_J_PX = 1, WG_J_XD0 = 2, WG_J_GX = 3, WG_J_GY = 4, WG_J_NENT = 5, WG_J_LD = 6,
              WG_J_XLIST = 8,                          // Gx x (map, channel)
              WG_J_YLIST = WG_J_XLIST + 2 * WG_MAX_G,  // Gy x (map, channel)
              WG_J_ENT = WG_J_YLIST + 2 * WG_MAX_G;    // entries
// entry: [0] a_off (bytes into the x stage), [1] TMEM column, [2..18) row_off per 8-row group (-1 = discard),
//        [18..18+32) col_off per 8-column group (-1 = discard).  dW element = row_off + (r%8)*ld + col_off + c%8.
constexpr int WG_E_AOFF = 0, WG_E_COL = 1, WG_E_ROW = 2, WG_E_COLOFF = 18, WG_E_SIZE = 18 + WG_MAX_G;

struct WgradParams {
  CUtensorMap map[WG_MAX_MAPS];
  const int* tab;
  float* dw;
  int* err;
  int N, D, H, W;
  int tiles_h, tiles_w;
  int n_jobs, job_stride, split;
  int x_f16;        // storage of the x and dy views: 0 = bf16, 1 = fp16 (tcgen05.mma rejects mixed A/B formats)
  int dbg;                // timing experiments only (env U3D_DBG): 32 = no loads
};

int wgrad_gemm_launch(const WgradParams& p, cudaStream_t stream);

}  // namespace u3d
