// tcgen05 shifted-GEMM convolution kernel for sm_100a (see conv_gemm.cuh for the formulation).
//
// Replaces the cuDNN calls behind nn.Conv3d / nn.ConvTranspose3d forward and their data gradients
// (reference call sites: network.py:394-395,403,411,415,541,545 and 312-314).
//
// Roles (warp-specialised, persistent over work items = (output tile, N block)):
//   warp 0  : A producer  - TMA 5-D tiled loads of (18 x 10 voxel) x 8-channel boxes, OOB zero fill = padding
//   warp 1  : W producer  - cp.async.bulk of pre-packed weight tiles, ring of up to 16 stages sized by the host plan
//   warp 2  : MMA issuer  - tcgen05.mma kind::f16, M=128 (16h x 8w voxels), N=nblk, K=16; accumulators in TMEM,
//             double buffered (2 x 256 columns) so the epilogue of item i overlaps the MMAs of item i+1
//   warps 3-6: epilogue   - tcgen05.ld, + bias / + addend, zero planes, per-(n,c) sum / sum^2 for
//             InstanceNorm (warp butterfly, fp64 atomics once per sample change), bf16 NDHWC stores
#include "conv_gemm.cuh"

namespace u3d {

namespace {

struct SmemCtl {
  uint64_t a_full[CG_A_STAGES], a_empty[CG_A_STAGES];
  uint64_t w_full[CG_W_STAGES], w_empty[CG_W_STAGES];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  int abort_flag;
  uint32_t tap_off[32];     // byte offset of each tap's window inside an A slab
};

// column sums over the 32 lanes of a warp: in v[j] = value of column j for this lane's row;
// returns the sum of column `lane`.  31 shuffles (recursive halving).
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      float keep = up ? v[j + half] : v[j];
      float send = up ? v[j] : v[j + half];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// All MMAs of one weight tile, fully unrolled for a compile-time (planes per segment, K steps): straight-line
// UIADD3.64 + UTCHMMA pairs.  (A rolled loop costs ~130 cycles per MMA on the single issuing thread: every
// uniform-datapath instruction of the loop-carried address chain stalls ~13 cycles; profiles/r01_notes.md.)
//   FUSE == 1: tap = one (sd,sh,sw) shift, one MMA of N = nblk per (plane, K step)
//   FUSE == 3: tap = one (sh,sw); the tile holds the d-shifts 2,1,0; input plane pl feeds output planes
//              pl-2..pl in ONE MMA of N = cnt * nblk (first tap of an item runs unfused so that the first MMA
//              into every accumulator overwrites it)
template <int DT, int G2, int FUSE>
__device__ __forceinline__ void issue_tap(uint32_t acc0, uint64_t adesc0, uint64_t bdesc0, uint32_t idesc1,
                                          uint32_t idesc2, uint32_t idesc3, uint32_t nblk, uint32_t accum) {
  constexpr uint64_t A_DINC = (uint64_t)(cg_plane_pitch(2 * G2) >> 4);        // one plane
  constexpr uint64_t A_KINC = 2;                                              // one K=16 step = 32 B inside the swizzled row
  const uint64_t b_kinc = (uint64_t)(2u * FUSE * nblk);                        // (2 chunks * FUSE*nblk*16 B) / 16
  if (FUSE == 1) {
#pragma unroll
    for (int d = 0; d < DT; ++d) {
#pragma unroll
      for (int kk = 0; kk < G2; ++kk)
        tc_mma_bf16(acc0 + (uint32_t)d * nblk, adesc0 + d * A_DINC + kk * A_KINC, bdesc0 + kk * b_kinc, idesc1,
                    kk == 0 ? accum : 1u);
    }
  } else if (accum == 0) {
#pragma unroll 1
    for (int d = 0; d < DT; ++d) {
#pragma unroll 1
      for (int sd = 0; sd < 3; ++sd) {
        uint64_t a = adesc0 + (uint64_t)(d + sd) * A_DINC, b = bdesc0 + (uint64_t)(2 - sd) * nblk;
        for (int kk = 0; kk < G2; ++kk, a += A_KINC, b += b_kinc)
          tc_mma_bf16(acc0 + (uint32_t)d * nblk, a, b, idesc1, (sd | kk) ? 1u : 0u);
      }
    }
  } else {
#pragma unroll
    for (int pl = 0; pl < DT + 2; ++pl) {
      const int hi = pl < 2 ? pl : 2;
      const int lo = pl - DT + 1 > 0 ? pl - DT + 1 : 0;
      const int cnt = hi - lo + 1;
      const uint32_t id = cnt == 3 ? idesc3 : (cnt == 2 ? idesc2 : idesc1);
#pragma unroll
      for (int kk = 0; kk < G2; ++kk)
        tc_mma_bf16(acc0 + (uint32_t)(pl - hi) * nblk, adesc0 + pl * A_DINC + kk * A_KINC,
                    bdesc0 + (uint64_t)(2 - hi) * nblk + kk * b_kinc, id, 1u);
    }
  }
}

// One weight-ring stage = a batch of taps (their tiles are contiguous in the packed stream): one barrier wait, one
// dispatch on (planes, K steps) and one commit per batch instead of per tap.  `mask` holds the taps of the batch.
template <int DT, int G2, int FUSE>
__device__ __forceinline__ void issue_batch(uint32_t mask, int cnt, const uint32_t* tap_off, uint32_t slab, uint32_t wst_addr,
                                         uint32_t wtile_bytes, uint32_t b_lbo, uint32_t acc0, uint32_t idesc1,
                                         uint32_t idesc2, uint32_t idesc3, uint32_t nblk, uint32_t accum) {
  for (int i = 0; i < cnt; ++i) {
    const int tap = __ffs(mask) - 1;
    mask &= mask - 1;
    const uint64_t adesc0 = umma_desc_sw(slab + tap_off[tap], CG_WB * 32 * G2, 32 * G2);
    const uint64_t bdesc0 = umma_desc(wst_addr + (uint32_t)i * wtile_bytes, b_lbo, 128);
    issue_tap<DT, G2, FUSE>(acc0, adesc0, bdesc0, idesc1, idesc2, idesc3, nblk, accum);
    accum = 1;
  }
}

template <int FUSE>
__device__ __forceinline__ void issue_batch_dispatch(int Dt, int G2, uint32_t mask, int cnt, const uint32_t* tap_off,
                                                     uint32_t slab, uint32_t wst_addr, uint32_t wtile_bytes, uint32_t b_lbo,
                                                     uint32_t acc0, uint32_t i1, uint32_t i2, uint32_t i3, uint32_t nblk,
                                                     uint32_t accum) {
#define U3D_CASE(DT, GG)          \
  if (Dt == DT && G2 == GG)       \
    return issue_batch<DT, GG, FUSE>(mask, cnt, tap_off, slab, wst_addr, wtile_bytes, b_lbo, acc0, i1, i2, i3, nblk, accum);
  U3D_CASE(8, 1) U3D_CASE(8, 2) U3D_CASE(4, 1) U3D_CASE(4, 2) U3D_CASE(4, 4) U3D_CASE(2, 1) U3D_CASE(2, 2) U3D_CASE(2, 4)
  U3D_CASE(1, 1) U3D_CASE(1, 2) U3D_CASE(1, 4) U3D_CASE(8, 4)
#undef U3D_CASE
}


// ---- dense fast path --------------------------------------------------------------------------------------------
// Ordinary 3x3x3 stride-1 convolutions with d-tap fusion and every tap active (the layers that hold >90 % of the
// FLOPs): the nine (kh,kw) taps, the planes and the K steps are unrolled at compile time with N block width NBLK a
// template parameter, so every descriptor / TMEM address is `base + immediate` (2-3 uniform instructions per MMA).
// The generic path pays ~45 instructions of mask / table / descriptor bookkeeping per tap on the single issuing
// thread, which bounded the level-0/1 layers (ncu: MMA warp 62 % of its samples in issue code, tensor pipe 42 %).
template <int DT, int G2, int NBLK>
__device__ __forceinline__ void issue_tap_dense(uint32_t acc0, uint64_t aslab, uint64_t bdesc0, uint32_t idesc1,
                                                uint32_t idesc2, uint32_t idesc3, const int t, const bool first) {
  constexpr uint64_t A_DINC = (uint64_t)(cg_plane_pitch(2 * G2) >> 4);
  constexpr uint64_t A_KINC = 2;
  constexpr uint64_t B_KINC = (uint64_t)(2 * 3 * NBLK);
  const uint64_t adesc0 = aslab + (uint64_t)(((t / 3) * CG_WB + (t % 3)) * 2 * G2);   // (kh * 10 + kw) rows of 32 * G2 bytes, in 16-B units
  (void)first;      // the epilogue hands every accumulator buffer back ZEROED (tcgen05.st), so every tap accumulates
#pragma unroll
  for (int pl = 0; pl < DT + 2; ++pl) {
    const int hi = pl < 2 ? pl : 2;
    const int lo = pl - DT + 1 > 0 ? pl - DT + 1 : 0;
    const int cnt = hi - lo + 1;
    const uint32_t id = cnt == 3 ? idesc3 : (cnt == 2 ? idesc2 : idesc1);
#pragma unroll
    for (int kk = 0; kk < G2; ++kk)
      tc_mma_bf16(acc0 + (uint32_t)((pl - hi) * NBLK), adesc0 + (uint64_t)pl * A_DINC + kk * A_KINC,
                  bdesc0 + (uint64_t)((2 - hi) * NBLK) + kk * B_KINC, id, 1u);
  }
}

struct MmaRoleArgs {
  SmemCtl* ctl;
  uint32_t slab0, slab_bytes, wring0, wstage_bytes, wtile_bytes, tmem_base;
  int n_tiles;
};

template <int DT, int G2, int NBLK>
__device__ __noinline__ void mma_role_dense(const ConvGemmParams& p, const MmaRoleArgs r) {
  // The WHOLE role runs in one elected thread, barrier waits included: any per-tap warp-level step (elect.sync,
  // __syncwarp, a warp-wide barrier poll) between two taps lets the shallow tcgen05 queue drain -- measured +36 % on
  // the level-0 layers against the back-to-back rate of tools/probe/mma_rate.cu.
  if (!elect_one()) return;
  SmemCtl* ctl = r.ctl;
  volatile int* abort_flag = &ctl->abort_flag;
  const uint32_t idesc1 = umma_idesc_bf16(128, NBLK, 0, 0, p.in_f16, p.in_f16);
  const uint32_t idesc2 = umma_idesc_bf16(128, 2 * NBLK, 0, 0, p.in_f16, p.in_f16);
  const uint32_t idesc3 = umma_idesc_bf16(128, 3 * NBLK, 0, 0, p.in_f16, p.in_f16);
  const uint32_t a_stages = (uint32_t)p.a_stages, w_stages = (uint32_t)p.w_stages;
  const int wT = p.wT, n_cg = p.n_cg;
  const uint64_t wtile16 = (uint64_t)(r.wtile_bytes >> 4);
  uint32_t a_it = 0, w_it = 0, acc_it = 0;
  const bool timing = p.dbg_out != nullptr && blockIdx.x == 0;
  long long t_acc = 0, t_a = 0, t_w = 0, n_items = 0, t_begin = timing ? clock64() : 0;
  for (int item = blockIdx.x; item < p.n_work; item += gridDim.x) {
    const uint32_t buf = p.nbuf == 2 ? (acc_it & 1) : 0u, aph = p.nbuf == 2 ? ((acc_it >> 1) & 1) : (acc_it & 1);
    long long tq = timing ? clock64() : 0;
    if (!mbar_wait(smem_u32(&ctl->acc_empty[buf]), aph ^ 1, abort_flag, p.err, 103)) return;
    if (timing) t_acc += clock64() - tq, ++n_items;
    tc_fence_after();
    const uint32_t acc0 = r.tmem_base + buf * 256;
    for (int cg = 0; cg < n_cg; ++cg) {
      const uint32_t ast = a_it % a_stages, aphase = (a_it / a_stages) & 1;
      tq = timing ? clock64() : 0;
      if (!mbar_wait(smem_u32(&ctl->a_full[ast]), aphase, abort_flag, p.err, 104)) return;
      if (timing) t_a += clock64() - tq;
      tc_fence_after();
      const uint64_t aslab = umma_desc_sw(r.slab0 + ast * r.slab_bytes, CG_WB * 32 * G2, 32 * G2);
      uint64_t bdesc = 0;
      uint32_t wst = 0;
      int in_batch = 0;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (in_batch == 0) {
          wst = w_it % w_stages;
          tq = timing ? clock64() : 0;
          if (!mbar_wait(smem_u32(&ctl->w_full[wst]), (w_it / w_stages) & 1, abort_flag, p.err, 105)) return;
          if (timing) t_w += clock64() - tq;
          tc_fence_after();
          bdesc = umma_desc(r.wring0 + wst * r.wstage_bytes, 3 * NBLK * 16, 128);
        }
        ++in_batch;
        if (!(p.dbg & 8)) issue_tap_dense<DT, G2, NBLK>(acc0, aslab, bdesc, idesc1, idesc2, idesc3, t, cg == 0);
        bdesc += wtile16;
        if (in_batch == wT || t == 8) {
          tc_commit(smem_u32(&ctl->w_empty[wst]));
          ++w_it;
          in_batch = 0;
        }
      }
      tc_commit(smem_u32(&ctl->a_empty[ast]));
      ++a_it;
    }
    tc_commit(smem_u32(&ctl->acc_full[buf]));
    ++acc_it;
  }
  if (timing) {
    p.dbg_out[0] = clock64() - t_begin;
    p.dbg_out[1] = t_acc;
    p.dbg_out[2] = t_a;
    p.dbg_out[3] = t_w;
    p.dbg_out[4] = n_items;
  }
}

__device__ __forceinline__ bool mma_role_dense_dispatch(const ConvGemmParams& p, const MmaRoleArgs& r) {
  const int G2 = p.G / 2;
#define U3D_DENSE(DT, GG, NB)                              \
  if (p.Dt == DT && G2 == GG && p.nblk == NB) {            \
    mma_role_dense<DT, GG, NB>(p, r);                      \
    return true;                                           \
  }
#define U3D_DENSE_G(DT, NB) U3D_DENSE(DT, 1, NB) U3D_DENSE(DT, 2, NB) U3D_DENSE(DT, 4, NB)
  U3D_DENSE_G(8, 32) U3D_DENSE_G(4, 32) U3D_DENSE_G(2, 32) U3D_DENSE_G(1, 32)
  U3D_DENSE_G(4, 64) U3D_DENSE_G(2, 64) U3D_DENSE_G(1, 64)
  U3D_DENSE(8, 1, 64) U3D_DENSE(8, 2, 64)
#undef U3D_DENSE_G
#undef U3D_DENSE
  return false;
}

__global__ void __launch_bounds__(CG_THREADS, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int planes = p.Dt + 2;
  const uint32_t plane_pitch = (uint32_t)cg_plane_pitch(p.G);
  const uint32_t row_bytes = (uint32_t)cg_row_bytes(p.G);
  const uint32_t slab_bytes = (uint32_t)planes * plane_pitch;
  const uint32_t wtile_bytes = (uint32_t)p.G * p.fuse * p.nblk * 16;
  const int wT = p.wT;                                   // taps per weight-ring stage
  const uint32_t w_stages = (uint32_t)p.w_stages;
  const uint32_t wstage_bytes = (uint32_t)wT * wtile_bytes;
  const uint32_t slab0 = smem_u32(smem) + 1024;
  const uint32_t a_stages = (uint32_t)p.a_stages;
  const uint32_t wring0 = slab0 + a_stages * slab_bytes;

  const int* tab_map = p.tab;
  const int* tab_ch = p.tab + p.n_cg;
  const int* tab_shift = p.tab + 2 * p.n_cg;
  const int* tab_mask = tab_shift + p.n_taps;
  const int* tab_wbase = tab_mask + p.n_nblk * p.n_cg;
  const int* tab_coff = tab_wbase + p.n_nblk;
  const int* tab_ooff = tab_coff + p.n_nblk;

  if (threadIdx.x == 0) {
    for (int i = 0; i < CG_A_STAGES; ++i) {
      mbar_init(smem_u32(&ctl->a_full[i]), 1);
      mbar_init(smem_u32(&ctl->a_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&ctl->acc_full[i]), 1);
      mbar_init(smem_u32(&ctl->acc_empty[i]), CG_EPI_WARPS);
    }
    for (int i = 0; i < CG_W_STAGES; ++i) {
      mbar_init(smem_u32(&ctl->w_full[i]), 1);
      mbar_init(smem_u32(&ctl->w_empty[i]), 1);
    }
    ctl->abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&ctl->tmem_base), 512);
    tmem_relinquish();
  }
  if (threadIdx.x >= 96 && threadIdx.x < 96 + 32 && (int)threadIdx.x - 96 < p.n_taps) {
    const int sh = tab_shift[threadIdx.x - 96];
    ctl->tap_off[threadIdx.x - 96] =
        (uint32_t)(sh & 0xff) * plane_pitch + (uint32_t)(((sh >> 8) & 0xff) * CG_WB + ((sh >> 16) & 0xff)) * row_bytes;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  if (p.dense) {
    // dense path: accumulators start (and are handed back by the epilogue) zeroed, so that no tap needs the
    // overwrite form -- the first tap of an item then runs d-fused like all others (10 wide MMAs instead of 24 narrow)
    if (warp >= 3 && warp < 7) {
      for (int c = 0; c < 512; c += 32) tmem_st_zero_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + c);
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  volatile int* abort_flag = &ctl->abort_flag;
  const uint32_t* tap_off = ctl->tap_off;

  const int n_tiles = p.N * p.segs_d * p.tiles_h * p.tiles_w;

  // The three single-issuer roles keep their whole warp converged and guard only the issuing
  // instructions with elect.sync: descriptors / coordinates then stay in uniform registers
  // (a `lane == 0` region makes ptxas serialise every TMA / MMA operand through a waterfall loop).
  const int planes_i = planes, G = p.G, nblk = p.nblk, Dt = p.Dt, n_cg = p.n_cg;
  if (warp == 0) {
    // ================= A producer =================
    if (elect_one())
      for (int i = 0; i < CG_MAX_MAPS; ++i) tma_prefetch_desc(&p.amap[i]);
    uint32_t a_it = 0;
    bool ok = true;
    for (int item = blockIdx.x; item < p.n_work && ok; item += gridDim.x) {
      const int nb = item / n_tiles;
      int t = item - nb * n_tiles;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int seg = t % p.segs_d;
      const int n = t / p.segs_d;
      const int w0 = tw * CG_WT - 1, h0 = th * CG_HT - 1, d0 = seg * Dt - 1;
      for (int cg = 0; cg < n_cg; ++cg) {
        if (__ldg(&tab_mask[nb * n_cg + cg]) == 0) continue;
        const uint32_t st = a_it % a_stages, ph = (a_it / a_stages) & 1;
        if (!mbar_wait_relaxed(smem_u32(&ctl->a_empty[st]), ph ^ 1, abort_flag, p.err, 101)) { ok = false; break; }
        const uint32_t full = smem_u32(&ctl->a_full[st]);
        const CUtensorMap* m = &p.amap[__ldg(&tab_map[cg])];
        const int ch0 = __ldg(&tab_ch[cg]);
        if (elect_one()) {
          if (p.dbg & 32) {
            mbar_arrive(full);                      // timing experiment: no A loads
          } else {
            mbar_expect_tx(full, (uint32_t)planes_i * (uint32_t)cg_plane_bytes(G));
            uint32_t dst = slab0 + st * slab_bytes;
            for (int pl = 0; pl < planes_i; ++pl, dst += plane_pitch)
              tma_load_5d(dst, m, full, ch0, w0, h0, d0 + pl, n);
          }
        }
        __syncwarp();
        ++a_it;
      }
    }
  } else if (warp == 1) {
    // ================= W producer =================
    uint32_t w_it = 0;
    bool ok = true;
    for (int item = blockIdx.x; item < p.n_work && ok; item += gridDim.x) {
      const int nb = item / n_tiles;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w) + (size_t)__ldg(&tab_wbase[nb]) * wtile_bytes;
      for (int cg = 0; cg < n_cg && ok; ++cg) {
        int left = __popc((uint32_t)__ldg(&tab_mask[nb * n_cg + cg]));
        while (left > 0) {
          const int cnt = left < wT ? left : wT;
          left -= cnt;
          const uint32_t st = w_it % w_stages, ph = (w_it / w_stages) & 1;
          if (!mbar_wait_relaxed(smem_u32(&ctl->w_empty[st]), ph ^ 1, abort_flag, p.err, 102)) { ok = false; break; }
          if (elect_one()) {
            const uint32_t full = smem_u32(&ctl->w_full[st]);
            if (p.dbg & 64) {
              mbar_arrive(full);                    // timing experiment: no weight loads
            } else {
              mbar_expect_tx(full, (uint32_t)cnt * wtile_bytes);
              bulk_load(wring0 + st * wstage_bytes, src, (uint32_t)cnt * wtile_bytes, full);
            }
          }
          __syncwarp();
          src += (size_t)cnt * wtile_bytes;
          ++w_it;
        }
      }
    }
  } else if (warp == 2) {
    // ================= MMA issuer =================
    if (p.dense) {
      MmaRoleArgs r;
      r.ctl = ctl; r.slab0 = slab0; r.slab_bytes = slab_bytes; r.wring0 = wring0; r.wstage_bytes = wstage_bytes;
      r.wtile_bytes = wtile_bytes; r.tmem_base = tmem_base; r.n_tiles = n_tiles;
      if (!mma_role_dense_dispatch(p, r) && lane == 0) {       // the host never sets `dense` for other shapes
        ctl->abort_flag = 1;
        atomicCAS(p.err, 0, 107);
      }
    } else if (elect_one()) {
    // generic path (strided / transposed / 1x1x1 / wide-N layers, partial tap masks): like the dense path the WHOLE
    // role runs in one elected thread, barrier waits included
    const uint32_t idesc = umma_idesc_bf16(128, nblk, 0, 0, p.in_f16, p.in_f16);
    const uint32_t idesc2 = umma_idesc_bf16(128, 2 * nblk, 0, 0, p.in_f16, p.in_f16),
                   idesc3 = umma_idesc_bf16(128, 3 * nblk, 0, 0, p.in_f16, p.in_f16);
    const int G2 = G / 2;
    const int fuse = p.fuse;
    // descriptor increments (the start-address field counts 16-byte units)
    const uint32_t b_lbo = (uint32_t)fuse * nblk * 16;             // stride between the two K chunks of one MMA
    uint32_t a_it = 0, w_it = 0, acc_it = 0;
    bool ok = true;
    for (int item = blockIdx.x; item < p.n_work && ok; item += gridDim.x) {
      const int nb = item / n_tiles;
      const uint32_t buf = p.nbuf == 2 ? (acc_it & 1) : 0u, aph = p.nbuf == 2 ? ((acc_it >> 1) & 1) : (acc_it & 1);
      if (!mbar_wait(smem_u32(&ctl->acc_empty[buf]), aph ^ 1, abort_flag, p.err, 103)) break;
      tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * 256;
      uint32_t accum = 0;                       // 0 for the first tap of the item: overwrite the accumulators
      for (int cg = 0; cg < n_cg && ok; ++cg) {
        uint32_t mask = (uint32_t)__ldg(&tab_mask[nb * n_cg + cg]);
        if (mask == 0) continue;
        const uint32_t ast = a_it % a_stages, aphase = (a_it / a_stages) & 1;
        if (!mbar_wait(smem_u32(&ctl->a_full[ast]), aphase, abort_flag, p.err, 104)) { ok = false; break; }
        tc_fence_after();
        const uint32_t slab = slab0 + ast * slab_bytes;
        while (mask) {
          const int left = __popc(mask);
          const int cnt = left < wT ? left : wT;
          uint32_t batch = mask;                       // the `cnt` lowest set bits
          for (int i = 0; i < cnt; ++i) mask &= mask - 1;
          batch &= ~mask;
          const uint32_t wst = w_it % w_stages, wph = (w_it / w_stages) & 1;
          if (!mbar_wait(smem_u32(&ctl->w_full[wst]), wph, abort_flag, p.err, 105)) { ok = false; break; }
          tc_fence_after();
          if (fuse == 1)
            issue_batch_dispatch<1>(Dt, G2, batch, cnt, tap_off, slab, wring0 + wst * wstage_bytes, wtile_bytes, b_lbo, acc0,
                                    idesc, idesc2, idesc3, (uint32_t)nblk, accum);
          else
            issue_batch_dispatch<3>(Dt, G2, batch, cnt, tap_off, slab, wring0 + wst * wstage_bytes, wtile_bytes, b_lbo, acc0,
                                    idesc, idesc2, idesc3, (uint32_t)nblk, accum);
          tc_commit(smem_u32(&ctl->w_empty[wst]));
          ++w_it;
          accum = 1;
        }
        if (!ok) break;
        tc_commit(smem_u32(&ctl->a_empty[ast]));
        ++a_it;
      }
      if (!ok) break;
      tc_commit(smem_u32(&ctl->acc_full[buf]));
      ++acc_it;
    }
    }
  } else {
    // ================= epilogue (warps 3..6) =================
    // CG_EPI_WARPS / 4 warps per TMEM lane quarter; the (channel block, plane) units of an item alternate between them.
    // (Measured: 8 epilogue warps are slower than 4 -- they take issue slots from the MMA warp, the critical role.)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 3) >> 2;       // which of the quarter's warps (always 0 with 4 epilogue warps)
    const int row = q * 32 + lane;
    const int line = row >> 3, wi = row & 7;
    const int n_cc = p.nblk / 32;
    const int of16 = p.out_f16;
    const bool do_stats = p.stats != nullptr && !(p.dbg & 2);
    float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
    int cur_n = -1, cur_nb = -1;
    uint32_t acc_it = 0;
    auto flush_stats = [&]() {
      if (!do_stats || cur_n < 0) return;
      const int coff = __ldg(&tab_coff[cur_nb]) & 0x3fffffff;
      for (int cc = 0; cc < n_cc; ++cc) {
        const int c = coff + cc * 32 + lane;
        if (c < p.stats_C) {
          double* s = p.stats + ((size_t)cur_n * p.stats_C + c) * 2;
          atomicAdd(s, (double)ssum[cc]);
          atomicAdd(s + 1, (double)ssq[cc]);
        }
        ssum[cc] = 0.f;
        ssq[cc] = 0.f;
      }
    };
    for (int item = blockIdx.x; item < p.n_work; item += gridDim.x) {
      const int nb = item / n_tiles;
      int t = item - nb * n_tiles;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int seg = t % p.segs_d;
      const int n = t / p.segs_d;
      if (n != cur_n || nb != cur_nb) {
        flush_stats();
        cur_n = n;
        cur_nb = nb;
      }
      const uint32_t buf = p.nbuf == 2 ? (acc_it & 1) : 0u, aph = p.nbuf == 2 ? ((acc_it >> 1) & 1) : (acc_it & 1);
      if (!mbar_wait_relaxed(smem_u32(&ctl->acc_full[buf]), aph, abort_flag, p.err, 106)) break;
      tc_fence_after();
      const int coff_raw = __ldg(&tab_coff[nb]);
      const int coff = coff_raw & 0x3fffffff;
      bf16* const outp = (coff_raw >> 30) ? p.out2 : p.out;
      const bf16* const addp = (coff_raw >> 30) ? p.addend2 : p.addend;
      const int ooff = __ldg(&tab_ooff[nb]);
      const int gh = th * CG_HT + line, gw = tw * CG_WT + wi;
      const int oh = gh * p.omul + ((ooff >> 8) & 0xff), ow = gw * p.omul + ((ooff >> 16) & 0xff);
      const bool hw_ok = gh < p.H && gw < p.W;
      for (int cc = 0; cc < ((p.dbg & 16) ? 0 : n_cc); ++cc) {
        // per-thread partial sums of this warp's planes; ONE cross-lane reduction per (item, cc) instead of per plane
        float a1[32], a2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) a1[j] = a2[j] = 0.f;
        bool any = false;
        const int c0 = coff + cc * 32;
        float bias_v[32];
        if (p.bias != nullptr) {
          const float* b = p.bias + nb * p.nblk + cc * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) bias_v[j] = __ldg(&b[j]);
        }
        // residual-gradient addend: the 64 bytes of plane d + 1 are requested while plane d is converted and stored (the
        // load used to sit between tcgen05.ld and the store of the same plane: ~1 us of exposed global latency per
        // plane made the epilogue, not the MMAs, the critical path of every dgrad launch with an addend -- 430 vs 240 us
        // at level 0)
        constexpr int D_STEP = CG_EPI_WARPS / 4;
        const int d_first = (CG_EPI_WARPS == 8 ? ((cc * p.Dt + half) & 1) : 0);
        uint4 add_nxt[4];
        auto addend_fetch = [&](int d, uint4 (&dst)[4]) {
          const int gd = seg * p.Dt + d;
          const int od = gd * p.omul + (ooff & 0xff);
          if (addp == nullptr || !(hw_ok && gd < p.D) || d >= p.Dt) return;
          const long long off = (long long)n * p.out_sN + (long long)od * p.out_sD + (long long)oh * p.out_sH +
                                (long long)ow * p.out_sW + coff;
          const uint4* ap = reinterpret_cast<const uint4*>(addp + off + cc * 32);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            if (c0 + k4 * 8 < p.out_C) dst[k4] = __ldg(ap + k4);
        };
        addend_fetch(d_first, add_nxt);
        for (int d = d_first; d < p.Dt; d += D_STEP) {
          const int gd = seg * p.Dt + d;
          const int od = gd * p.omul + (ooff & 0xff);
          const bool valid = hw_ok && gd < p.D;
          const bool zero = (od == p.zD) || (oh == p.zH) || (ow == p.zW);
          const long long off = (long long)n * p.out_sN + (long long)od * p.out_sD + (long long)oh * p.out_sH +
                                (long long)ow * p.out_sW + coff;
          uint4 add_cur[4];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) add_cur[k4] = add_nxt[k4];
          addend_fetch(d + D_STEP, add_nxt);
          uint32_t raw[32];
          __syncwarp();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256 + (uint32_t)d * p.nblk + cc * 32;
          tmem_ld_32x32(taddr, raw);
          tmem_ld_wait();
          if (p.dense) tmem_st_zero_32x32(taddr);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bias_v[j];
          }
          if (addp != nullptr && valid) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (c0 + k4 * 8 < p.out_C) {
                const uint4 u = add_cur[k4];
                float2 f;
                f = unpack_2x16(u.x, of16); v[k4 * 8 + 0] += f.x; v[k4 * 8 + 1] += f.y;
                f = unpack_2x16(u.y, of16); v[k4 * 8 + 2] += f.x; v[k4 * 8 + 3] += f.y;
                f = unpack_2x16(u.z, of16); v[k4 * 8 + 4] += f.x; v[k4 * 8 + 5] += f.y;
                f = unpack_2x16(u.w, of16); v[k4 * 8 + 6] += f.x; v[k4 * 8 + 7] += f.y;
              }
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.01f * v[j];
          }
          if (!valid || zero) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (valid && !(p.dbg & 4)) {
            uint4* op = reinterpret_cast<uint4*>(outp + off + cc * 32);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (c0 + k4 * 8 < p.out_C) {
                uint4 u;
                u.x = pack_2x16(v[k4 * 8 + 0], v[k4 * 8 + 1], of16);
                u.y = pack_2x16(v[k4 * 8 + 2], v[k4 * 8 + 3], of16);
                u.z = pack_2x16(v[k4 * 8 + 4], v[k4 * 8 + 5], of16);
                u.w = pack_2x16(v[k4 * 8 + 6], v[k4 * 8 + 7], of16);
                op[k4] = u;
              }
            }
          }
          if (do_stats) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              a1[j] += v[j];
              a2[j] = fmaf(v[j], v[j], a2[j]);
            }
            any = true;
          }
        }
        if (do_stats && any) {              // `any` is warp-uniform
          ssum[cc] += warp_colsum32(a1, lane);
          ssq[cc] += warp_colsum32(a2, lane);
        }
      }
      if (p.dense) tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ctl->acc_empty[buf]));
      ++acc_it;
    }
    flush_stats();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

size_t conv_gemm_smem_bytes(int Dt, int G, int nblk, int fuse, int wT, int w_stages, int a_stages) {
  return 1024 /*align slack*/ + 1024 /*ctl*/ + (size_t)a_stages * (Dt + 2) * cg_plane_pitch(G) +
         (size_t)w_stages * wT * G * fuse * nblk * 16;
}

int conv_gemm_launch(const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  if (p.G < 2 || (p.G & 1) || p.nblk % 32 != 0 || p.nblk > 128 || p.Dt < 1 || (p.nbuf != 1 && p.nbuf != 2) ||
      p.Dt * p.nblk * p.nbuf > 512 ||
      p.n_taps < 1 || p.n_taps > 32 || p.n_work < 1 || (p.Dt != 1 && p.Dt != 2 && p.Dt != 4 && p.Dt != 8) || (p.G != 2 && p.G != 4 && p.G != 8) || (p.fuse != 1 && p.fuse != 3) || p.fuse * p.nblk > 256)
    return U3D_ERR_INVALID;
  if (p.wT < 1 || p.wT > 32 || p.w_stages < 2 || p.w_stages > CG_W_STAGES || p.a_stages < 2 || p.a_stages > CG_A_STAGES)
    return U3D_ERR_INVALID;
  const size_t smem = conv_gemm_smem_bytes(p.Dt, p.G, p.nblk, p.fuse, p.wT, p.w_stages, p.a_stages);
  if (smem > 227 * 1024) return U3D_ERR_INVALID;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    if (cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return U3D_ERR_CUDA;
  }
  const int grid = p.n_work < num_sms ? p.n_work : num_sms;
  conv_gemm_kernel<<<grid, CG_THREADS, smem, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA;
}

}  // namespace u3d
