// tcgen05 weight-gradient kernel for sm_100a (see wgrad_gemm.cuh).
// Replaces the cuDNN backward-filter dispatch behind nn.Conv3d / nn.ConvTranspose3d
// (reference layers: network.py:394-395,403,411,312-313).
#include "wgrad_gemm.cuh"

namespace u3d {

namespace {

struct WgCtl {
  uint64_t full[2], empty[2], acc_full;
  uint32_t tmem_base;
  int abort_flag;
  uint32_t ent_aoff[16], ent_col[16];
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_gemm_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;
  WgCtl* ctl = reinterpret_cast<WgCtl*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int job = blockIdx.x / p.split, sid = blockIdx.x - job * p.split;
  const int* jt = p.tab + (size_t)job * p.job_stride;
  const int Dt = jt[WG_J_DT], Px = jt[WG_J_PX], xd0 = jt[WG_J_XD0], Gx = jt[WG_J_GX], Gy = jt[WG_J_GY];
  const int n_ent = jt[WG_J_NENT], ld = jt[WG_J_LD];
  const int segs = (p.D + Dt - 1) / Dt;
  const int n_tiles = p.N * segs * p.tiles_h * p.tiles_w;
  if (sid >= n_tiles) return;                       // nothing to contribute (uniform for the CTA)

  // swizzled whole-row boxes (plan.py: wx / wy chunks per TMA box) or the 16-byte-row layout (wx == 0)
  const int sw_word = jt[7];
  const int wx = sw_word & 0xff, wy = (sw_word >> 8) & 0xff, nbx = (sw_word >> 16) & 0xff, nby = (sw_word >> 24) & 0xff;
  const bool sw = wx != 0;
  // pair mode (16-byte-row layout only): two consecutive dy planes side by side in N (plan.py), so an MMA is N = 2*Gy*8
  // wide and the plane loop advances by two; taps kd = row plane - column plane are sorted out by the epilogue offsets
  const int npl = (sw_word >> 30) & 1 ? 2 : 1;
  const uint32_t rbx = 16u * wx, rby = 16u * wy;                                   // row bytes
  const uint32_t x_box_bytes = (uint32_t)CG_HB * CG_WB * rbx, y_box_bytes = (uint32_t)CG_HT * CG_WT * rby;
  const uint32_t x_pitch = sw ? (x_box_bytes + 8 * rbx - 1) / (8 * rbx) * (8 * rbx) : 0u, y_pitch = y_box_bytes;
  const uint32_t xplane = sw ? (uint32_t)nbx * x_pitch : (uint32_t)Gx * CG_CHUNK_PITCH,
                 yplane = sw ? (uint32_t)nby * y_pitch : (uint32_t)Gy * WG_DY_BOX_BYTES;
  const uint32_t xstage = (uint32_t)Px * xplane, ystage = (uint32_t)Dt * yplane;
  const uint32_t stage_bytes = xstage + ystage;
  const uint32_t stage0 = smem_u32(smem) + 1024;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&ctl->full[i]), 1);
      mbar_init(smem_u32(&ctl->empty[i]), 1);
    }
    mbar_init(smem_u32(&ctl->acc_full), 1);
    ctl->abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&ctl->tmem_base), 512);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && (int)threadIdx.x - 64 < n_ent && threadIdx.x < 64 + 16) {
    const int* ent = jt + WG_J_ENT + (threadIdx.x - 64) * WG_E_SIZE;
    ctl->ent_aoff[threadIdx.x - 64] = sw ? ((uint32_t)ent[WG_E_AOFF] >> 4) * rbx : (uint32_t)ent[WG_E_AOFF];   // row shift
    ctl->ent_col[threadIdx.x - 64] = (uint32_t)ent[WG_E_COL];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  volatile int* abort_flag = &ctl->abort_flag;
  const uint32_t* ent_aoff = ctl->ent_aoff;
  const uint32_t* ent_col = ctl->ent_col;

  if (warp == 0) {
    if (elect_one())
      for (int i = 0; i < WG_MAX_MAPS; ++i) tma_prefetch_desc(&p.map[i]);
    uint32_t it = 0;
    for (int t = sid; t < n_tiles; t += p.split, ++it) {
      const uint32_t st = it & 1, ph = (it >> 1) & 1;
      if (!mbar_wait(smem_u32(&ctl->empty[st]), ph ^ 1, abort_flag, p.err, 201)) break;
      int r = t;
      const int tw = r % p.tiles_w; r /= p.tiles_w;
      const int th = r % p.tiles_h; r /= p.tiles_h;
      const int seg = r % segs;
      const int n = r / segs;
      if (elect_one()) {
        const uint32_t full = smem_u32(&ctl->full[st]);
        if (p.dbg & 32) {
          mbar_arrive(full);                            // timing experiment: no loads
        } else if (sw) {
          mbar_expect_tx(full, (uint32_t)Px * nbx * x_box_bytes + (uint32_t)Dt * nby * y_box_bytes);
          uint32_t dst = stage0 + st * stage_bytes;
          for (int pl = 0; pl < Px; ++pl)
            for (int b = 0; b < nbx; ++b, dst += x_pitch)
              tma_load_5d(dst, &p.map[__ldg(&jt[WG_J_XLIST + 2 * b])], full, __ldg(&jt[WG_J_XLIST + 2 * b + 1]),
                          tw * CG_WT - 1, th * CG_HT - 1, seg * Dt + xd0 + pl, n);
          for (int d = 0; d < Dt; ++d)
            for (int b = 0; b < nby; ++b, dst += y_pitch) {
              // bits 16-17 of the channel word: w shift of this dy box (plan.py "w-shift" mode: the three kw taps of a
              // 3x3x3 layer are three copies of the dy tile shifted by +1 / 0 / -1 voxel in w, side by side in N)
              const int cw = __ldg(&jt[WG_J_YLIST + 2 * b + 1]);
              const int code = (cw >> 16) & 3;
              tma_load_5d(dst, &p.map[__ldg(&jt[WG_J_YLIST + 2 * b])], full, cw & 0xffff,
                          tw * CG_WT + (code ? code - 2 : 0), th * CG_HT, seg * Dt + d, n);
            }
        } else {
          mbar_expect_tx(full, (uint32_t)Px * Gx * CG_BOX_BYTES + (uint32_t)Dt * Gy * WG_DY_BOX_BYTES);
          uint32_t dst = stage0 + st * stage_bytes;
          for (int pl = 0; pl < Px; ++pl)
            for (int g = 0; g < Gx; ++g, dst += CG_CHUNK_PITCH)
              tma_load_5d(dst, &p.map[__ldg(&jt[WG_J_XLIST + 2 * g])], full, __ldg(&jt[WG_J_XLIST + 2 * g + 1]),
                          tw * CG_WT - 1, th * CG_HT - 1, seg * Dt + xd0 + pl, n);
          for (int d = 0; d < Dt; ++d)
            for (int g = 0; g < Gy; ++g, dst += WG_DY_BOX_BYTES)
              tma_load_5d(dst, &p.map[__ldg(&jt[WG_J_YLIST + 2 * g])], full, __ldg(&jt[WG_J_YLIST + 2 * g + 1]),
                          tw * CG_WT, th * CG_HT, seg * Dt + d, n);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // The whole issuing role runs in ONE elected thread, barrier waits included (no elect.sync / __syncwarp between
    // entries): a warp-level step between MMAs lets the shallow tcgen05 queue drain (measured on the conv kernel).
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, Gy * 8 * npl, 1, 1, p.x_f16, p.x_f16);
      // one K step = 16 voxels = two 8-voxel lines of the brick / of the dy tile.  The two layouts get separate loops so
      // that the 16-byte-row one keeps compile-time descriptor increments (its layers are bound by the MMA issue rate).
      const uint64_t a_kinc_sw = (uint64_t)((2 * CG_WB * rbx) >> 4), b_kinc_sw = (uint64_t)((2 * CG_WT * rby) >> 4);
      const uint64_t a_dinc = (uint64_t)(xplane >> 4), b_dinc = (uint64_t)(yplane >> 4);
      uint32_t it = 0;
      bool ok = true;
      for (int t = sid; t < n_tiles && ok; t += p.split, ++it) {
        const uint32_t st = it & 1, ph = (it >> 1) & 1;
        if (!mbar_wait(smem_u32(&ctl->full[st]), ph, abort_flag, p.err, 202)) { ok = false; break; }
        tc_fence_after();
        const uint32_t xs = stage0 + st * stage_bytes, ys = xs + xstage;
        if (sw) {
          const uint64_t b0 = umma_desc_mn_sw(ys, y_pitch, CG_WT * rby, rby);
          for (int e = 0; e < n_ent; ++e) {
            uint64_t a = umma_desc_mn_sw(xs + ent_aoff[e], x_pitch, CG_WB * rbx, rbx);
            uint64_t b = b0;
            const uint32_t acc = tmem_base + ent_col[e];
            for (int d = 0; d < Dt; ++d, a += a_dinc, b += b_dinc) {
              uint64_t ak = a, bk = b;
              tc_mma_bf16(acc, ak, bk, idesc, (it == 0 && d == 0) ? 0u : 1u);
#pragma unroll
              for (int k = 1; k < 8; ++k) {
                ak += a_kinc_sw;
                bk += b_kinc_sw;
                tc_mma_bf16(acc, ak, bk, idesc, 1u);
              }
            }
          }
        } else {
          constexpr uint64_t a_kinc = (uint64_t)((2 * CG_WB * 16) >> 4), b_kinc = (uint64_t)((2 * 128) >> 4);
          const uint64_t b0 = umma_desc(ys, 128, WG_DY_BOX_BYTES);
          for (int e = 0; e < n_ent; ++e) {
            uint64_t a = umma_desc(xs + ent_aoff[e], CG_WB * 16, CG_CHUNK_PITCH);
            uint64_t b = b0;
            const uint32_t acc = tmem_base + ent_col[e];
            for (int d = 0; d < Dt; d += npl, a += npl * a_dinc, b += npl * b_dinc) {
              tc_mma_bf16(acc, a, b, idesc, (it == 0 && d == 0) ? 0u : 1u);
#pragma unroll
              for (int k = 1; k < 8; ++k) tc_mma_bf16(acc, a + k * a_kinc, b + k * b_kinc, idesc, 1u);
            }
          }
        }
        tc_commit(smem_u32(&ctl->empty[st]));
      }
      if (ok) tc_commit(smem_u32(&ctl->acc_full));
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (mbar_wait(smem_u32(&ctl->acc_full), 0, abort_flag, p.err, 203)) {
      tc_fence_after();
      const int n_cc = Gy * npl / 4;
      for (int e = 0; e < n_ent; ++e) {
        const int* ent = jt + WG_J_ENT + e * WG_E_SIZE;
        const int col = __ldg(&ent[WG_E_COL]);
        const int ro = __ldg(&ent[WG_E_ROW + (row >> 3)]);
        for (int cc = 0; cc < n_cc; ++cc) {
          uint32_t raw[32];
          __syncwarp();
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col + cc * 32, raw);
          tmem_ld_wait();
          if (ro < 0) continue;
          float* base = p.dw + (size_t)ro + (size_t)(row & 7) * ld;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int co = __ldg(&ent[WG_E_COLOFF + cc * 4 + h]);
            if (co < 0) continue;
            red_add_v4(base + co, __uint_as_float(raw[h * 8 + 0]), __uint_as_float(raw[h * 8 + 1]),
                       __uint_as_float(raw[h * 8 + 2]), __uint_as_float(raw[h * 8 + 3]));
            red_add_v4(base + co + 4, __uint_as_float(raw[h * 8 + 4]), __uint_as_float(raw[h * 8 + 5]),
                       __uint_as_float(raw[h * 8 + 6]), __uint_as_float(raw[h * 8 + 7]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int wgrad_gemm_launch(const WgradParams& p, cudaStream_t stream) {
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    if (cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return U3D_ERR_CUDA;
  }
  // shared memory is sized for the largest stage any job can ask for; the host plan keeps
  // 2 * (Px*Gx*2944 + Dt*Gy*2048) + 2 KiB under the 227 KiB limit.
  wgrad_gemm_kernel<<<p.n_jobs * p.split, WG_THREADS, 227 * 1024, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA;
}

}  // namespace u3d
