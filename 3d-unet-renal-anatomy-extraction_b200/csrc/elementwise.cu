// HBM-bound kernels of the U-Net hot path (sm_100a): InstanceNorm(+dropout)+LeakyReLU(+residual)
// forward / backward, the 1-channel stem conv and the 1x1x1 classifier head (both far below the
// tensor-core ridge), the fused softmax + Dice / focal loss and its gradient, and the sliding-window
// blend.  Activations are bf16 NDHWC with the channel count padded to a multiple of 8 (16-byte
// vectors); one thread owns one 8-channel vector, threads of a block tile (channel chunk, voxel) so
// every warp access is a run of consecutive 16-byte words.
#include "kernels.cuh"
#include <cooperative_groups.h>

namespace u3d {

namespace {

constexpr float LRELU = 0.01f;     // nn.LeakyReLU default slope (network.py:165,390)
constexpr int IN_U = 4;            // voxels per thread and loop trip in the InstanceNorm kernels (loads issued up front)

// f16 = 1: the tensor is stored as IEEE half (forward activations in "fp16" precision mode), else bfloat16
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8], int f16 = 0) {
  float2 t;
  t = unpack_2x16(u.x, f16); f[0] = t.x; f[1] = t.y;
  t = unpack_2x16(u.y, f16); f[2] = t.x; f[3] = t.y;
  t = unpack_2x16(u.z, f16); f[4] = t.x; f[5] = t.y;
  t = unpack_2x16(u.w, f16); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int f16 = 0) {
  uint4 u;
  u.x = pack_2x16(f[0], f[1], f16);
  u.y = pack_2x16(f[2], f[3], f16);
  u.z = pack_2x16(f[4], f[5], f16);
  u.w = pack_2x16(f[6], f[7], f16);
  return u;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm3d(affine=False, eps=1e-5) statistics -> (mean, scale) per (n, c).
// network.py:175,401,315.  With Dropout3d in front (network.py:159-160,412-413) the normalised
// tensor is z = m*y, m in {0, 1/(1-p)} per (n,c):  IN(z) = (y - mean) * m / sqrt(m^2 var + eps).
// ---------------------------------------------------------------------------------------------
__global__ void in_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ drop,
                                   float2* __restrict__ table, int NC, double inv_count, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NC) return;
  const double mean = stats[2 * i] * inv_count;
  double var = stats[2 * i + 1] * inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  const double m = drop ? (double)drop[i] : 1.0;
  const double scale = m / sqrt(m * m * var + (double)eps);
  table[i] = make_float2((float)mean, (float)scale);
}

// out = lrelu((y - mean) * scale [+ shift] [+ skip]);  shift (per (n, c), optional) carries BatchNorm's affine offset
template <bool HAS_SKIP>
__global__ void __launch_bounds__(256, 2) in_apply_kernel(const uint4* __restrict__ y, const uint4* __restrict__ skip,
                                uint4* __restrict__ out, const float2* __restrict__ table,
                                const float* __restrict__ shift, int chunks, long long V, int Cp, int af) {
  const int n = blockIdx.y;
  const int ch = threadIdx.x;                   // 8-channel chunk
  float mean[8], scale[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 t = table[(size_t)n * Cp + ch * 8 + j];
    mean[j] = t.x;
    scale[j] = t.y;
    sh[j] = shift ? shift[(size_t)n * Cp + ch * 8 + j] : 0.f;
  }
  const size_t base = (size_t)n * V * chunks;
  // IN_U voxels per thread and trip: all loads of a trip are issued before the first use (memory-level parallelism; with
  // one voxel per trip the small level-2..4 tensors ran at 1.7-2.7 TB/s, latency-bound)
  const long long stride = (long long)gridDim.x * blockDim.y;
  for (long long v = (long long)blockIdx.x * blockDim.y + threadIdx.y; v < V; v += stride * IN_U) {
    uint4 ry[IN_U], rs[IN_U];
#pragma unroll
    for (int u = 0; u < IN_U; ++u) {
      const long long vv = v + u * stride;
      if (vv < V) {
        const size_t idx = base + (size_t)vv * chunks + ch;
        ry[u] = ld_stream(y + idx);
        if (HAS_SKIP) rs[u] = ld_stream(skip + idx);
      }
    }
#pragma unroll
    for (int u = 0; u < IN_U; ++u) {
      const long long vv = v + u * stride;
      if (vv >= V) break;
      float f[8], s[8];
      unpack8(ry[u], f, af);
      if (HAS_SKIP) unpack8(rs[u], s, af);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float z = fmaf(f[j] - mean[j], scale[j], sh[j]);
        if (HAS_SKIP) z += s[j];
        f[j] = z > 0.f ? z : LRELU * z;
      }
      out[base + (size_t)vv * chunks + ch] = pack8(f, af);
    }
  }
}

// Backward, pass 1:  g = (dout [+ dout2]) * lrelu'(out)   (sign(out) == sign(pre-activation));
// accumulates sum(g), sum(g * yhat) per (n, c).  g is also d(skip) of a residual block.
template <bool HAS_D2>
__global__ void __launch_bounds__(256, 2) in_bwd_reduce_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ dout2,
                                     const uint4* __restrict__ out, const uint4* __restrict__ y,
                                     uint4* __restrict__ g, const float2* __restrict__ table,
                                     const float* __restrict__ shift, double* __restrict__ sums, int chunks,
                                     long long V, int Cp, int af) {
  extern __shared__ float red[];   // [blockDim.y][chunks*8][2]
  const int n = blockIdx.y;
  const int ch = threadIdx.x;
  float mean[8], scale[8], sh[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 t = table[(size_t)n * Cp + ch * 8 + j];
    mean[j] = t.x;
    scale[j] = t.y;
    sh[j] = shift ? shift[(size_t)n * Cp + ch * 8 + j] : 0.f;     // yhat = (y - mean) * scale + shift
    s1[j] = 0.f;
    s2[j] = 0.f;
  }
  constexpr int RU = HAS_D2 ? 2 : IN_U;          // four tensors in flight per voxel with dout2: keep the registers
  const size_t base = (size_t)n * V * chunks;
  const long long stride = (long long)gridDim.x * blockDim.y;
  for (long long v = (long long)blockIdx.x * blockDim.y + threadIdx.y; v < V; v += stride * RU) {
    uint4 rd[RU], rd2[RU], ro[RU], ry[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const long long vv = v + u * stride;
      if (vv < V) {
        const size_t idx = base + (size_t)vv * chunks + ch;
        rd[u] = ld_stream(dout + idx);
        if (HAS_D2) rd2[u] = ld_stream(dout2 + idx);
        ry[u] = ld_stream(y + idx);
        if (out != nullptr) ro[u] = ld_stream(out + idx);
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const long long vv = v + u * stride;
      if (vv >= V) break;
      float d[8], o[8], yy[8];
      unpack8(rd[u], d, af);
      if (HAS_D2) {
        float d2[8];
        unpack8(rd2[u], d2, af);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += d2[j];
      }
      unpack8(ry[u], yy, af);
      if (out != nullptr) {
        unpack8(ro[u], o, af);
      } else {
        // no residual input: the activation's sign is the sign of the normalised value, `out` need not be read
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(yy[j] - mean[j], scale[j], sh[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = o[j] > 0.f ? d[j] : LRELU * d[j];
      if (g != nullptr) {
        const uint4 gp = pack8(d, af);
        g[base + (size_t)vv * chunks + ch] = gp;
        unpack8(gp, d, af);    // reduce what pass 2 will read back (the rounded g)
      }                        // g == null: pass 2 recomputes the same fp32 g from dout (no 16-bit round trip)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float yh = fmaf(yy[j] - mean[j], scale[j], sh[j]);
        s1[j] += d[j];
        s2[j] += d[j] * yh;
      }
    }
  }
  const int C8 = chunks * 8;
  float* my = red + ((size_t)threadIdx.y * C8 + ch * 8) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    my[2 * j] = s1[j];
    my[2 * j + 1] = s2[j];
  }
  __syncthreads();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < C8 * 2; i += blockDim.x * blockDim.y) {
    float acc = 0.f;
    for (int r = 0; r < (int)blockDim.y; ++r) acc += red[(size_t)r * C8 * 2 + i];
    atomicAdd(&sums[((size_t)n * Cp) * 2 + i], (double)acc);
  }
}

// Backward, pass 2:  dy = scale * (g - mean(g) - yhat * mean(g * yhat)); optional zeroing of the
// ConstantPad3d planes of a ConvTrans3D output (network.py:314) and per-channel sum(dy) (its bias grad).
template <bool ZERO_LAST, bool HAS_DSUM>
__global__ void __launch_bounds__(256, 2) in_bwd_apply_kernel(const uint4* __restrict__ g, const uint4* __restrict__ y,
                                    uint4* __restrict__ dy, const float2* __restrict__ table,
                                    const double* __restrict__ sums, const float* __restrict__ coef,
                                    double* __restrict__ dsum, int chunks,
                                    long long V, int Cp, double inv_count, int zero_last, int D, int H, int W, int af,
                                    int g_is_dout) {
  extern __shared__ float red[];
  const int n = blockIdx.y;
  const int ch = threadIdx.x;
  // dy = scale * (g - mg - yhat * mgy) = g * A + y * B + C  with per-channel constants
  // g_is_dout: `g` holds the upstream gradient of a norm WITHOUT residual input; g = dout * lrelu'(yhat) is recomputed
  // here (sign of (y - mean) * scale, as pass 1 did) instead of being written and re-read in 16 bit
  float ca[8], cb[8], cc[8], acc[8], tm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const size_t c = (size_t)n * Cp + ch * 8 + j;
    const float2 t = table[c];
    if (coef != nullptr) {          // BatchNorm: the host supplies dy = g * A + y * B + C per (n, c)
      ca[j] = coef[3 * c];
      cb[j] = coef[3 * c + 1];
      cc[j] = coef[3 * c + 2];
    } else {
      const float mg = (float)(sums[2 * c] * inv_count), mgy = (float)(sums[2 * c + 1] * inv_count);
      ca[j] = t.y;
      cb[j] = -t.y * t.y * mgy;
      cc[j] = -t.y * mg + t.y * t.y * mgy * t.x;
    }
    tm[j] = t.x;
    acc[j] = 0.f;
  }
  const size_t base = (size_t)n * V * chunks;
  const long long stride = (long long)gridDim.x * blockDim.y;
  for (long long v = (long long)blockIdx.x * blockDim.y + threadIdx.y; v < V; v += stride * IN_U) {
    uint4 rg[IN_U], ry[IN_U];
#pragma unroll
    for (int u = 0; u < IN_U; ++u) {
      const long long vv = v + u * stride;
      if (vv < V) {
        const size_t idx = base + (size_t)vv * chunks + ch;
        rg[u] = ld_stream(g + idx);
        ry[u] = ld_stream(y + idx);
      }
    }
#pragma unroll
    for (int u = 0; u < IN_U; ++u) {
      const long long vv = v + u * stride;
      if (vv >= V) break;
      float gg[8], yy[8];
      unpack8(rg[u], gg, af);
      unpack8(ry[u], yy, af);
      bool z = false;
      if (ZERO_LAST) {
        // V < 2^31 per sample: 32-bit index arithmetic (the 64-bit divisions doubled this kernel's time)
        const unsigned v32 = (unsigned)vv, q1 = v32 / (unsigned)W, wq = v32 - q1 * (unsigned)W;
        const unsigned dq = q1 / (unsigned)H, hq = q1 - dq * (unsigned)H;
        z = (wq == (unsigned)(W - 1)) || (hq == (unsigned)(H - 1)) || (dq == (unsigned)(D - 1));
      }
      if (g_is_dout) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gg[j] = (yy[j] - tm[j]) * ca[j] > 0.f ? gg[j] : LRELU * gg[j];     // ca = scale (InstanceNorm)
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float r = fmaf(gg[j], ca[j], fmaf(yy[j], cb[j], cc[j]));
        if (ZERO_LAST && z) r = 0.f;
        gg[j] = r;
      }
      const uint4 o = pack8(gg, af);
      dy[base + (size_t)vv * chunks + ch] = o;
      if (HAS_DSUM) {
        unpack8(o, gg, af);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += gg[j];
      }
    }
  }
  if (HAS_DSUM) {
    const int C8 = chunks * 8;
    float* my = red + (size_t)threadIdx.y * C8 + ch * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) my[j] = acc[j];
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int i = tid; i < C8; i += blockDim.x * blockDim.y) {
      float a = 0.f;
      for (int r = 0; r < (int)blockDim.y; ++r) a += red[(size_t)r * C8 + i];
      atomicAdd(&dsum[i], (double)a);
    }
  }
}

// Both backward passes of an InstanceNorm in ONE launch, for tensors whose per-sample slice is small (levels 2-4 of the
// default net: 0.5-8 MB).  A thread-block CLUSTER owns one 8-channel chunk of one sample: its CTAs split the voxels,
// reduce locally, exchange the partial sums through distributed shared memory (added in rank order: reproducible, no
// atomics), then apply from L2-resident data.  The two-kernel path costs two launches of 10-20 us each there
// (latency-bound, 0.8-2.4 TB/s); a first single-CTA version of this kernel was no faster than the pair (21 us at level
// 3: 30-60 CTAs each walking 4096 voxels serially) -- the cluster is what shortens the per-thread chain to 1-4 trips.
// g == nullptr (only with out == nullptr, dout2 == nullptr): the activation gradient is recomputed in pass 2.
template <bool HAS_D2>
__global__ void __launch_bounds__(256) in_bwd_small_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ dout2,
                                   const uint4* __restrict__ out, const uint4* __restrict__ y, uint4* __restrict__ g,
                                   uint4* __restrict__ dy, const float2* __restrict__ table, double* __restrict__ sums,
                                   int chunks, int V, int Cp, double inv_count, int af) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int KC = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  extern __shared__ float red[];   // [blockDim.x][8][2]
  __shared__ float part[16];       // this CTA's partial sums: read by every CTA of the cluster
  __shared__ double tot[16];
  const int n = blockIdx.y;
  const int ch = blockIdx.x / KC;
  const int T = blockDim.x;
  float mean[8], scale[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 t = table[(size_t)n * Cp + ch * 8 + j];
    mean[j] = t.x;
    scale[j] = t.y;
    s1[j] = 0.f;
    s2[j] = 0.f;
  }
  constexpr int RU = HAS_D2 ? 2 : IN_U;
  const size_t base = (size_t)n * V * chunks;
  const int v0 = crank * T + threadIdx.x, vstep = KC * T;
  for (int v = v0; v < V; v += vstep * RU) {
    uint4 rd[RU], rd2[RU], ro[RU], ry[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int vv = v + u * vstep;
      if (vv < V) {
        const size_t idx = base + (size_t)vv * chunks + ch;
        rd[u] = dout[idx];
        if (HAS_D2) rd2[u] = dout2[idx];
        ry[u] = y[idx];
        if (out != nullptr) ro[u] = out[idx];
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int vv = v + u * vstep;
      if (vv >= V) break;
      float d[8], o[8], yy[8];
      unpack8(rd[u], d, af);
      if (HAS_D2) {
        float d2[8];
        unpack8(rd2[u], d2, af);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += d2[j];
      }
      unpack8(ry[u], yy, af);
      if (out != nullptr) {
        unpack8(ro[u], o, af);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (yy[j] - mean[j]) * scale[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = o[j] > 0.f ? d[j] : LRELU * d[j];
      if (g != nullptr) {
        const uint4 gp = pack8(d, af);
        g[base + (size_t)vv * chunks + ch] = gp;     // re-read by this same thread in pass 2
        unpack8(gp, d, af);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float yh = (yy[j] - mean[j]) * scale[j];
        s1[j] += d[j];
        s2[j] += d[j] * yh;
      }
    }
  }
  // warp tree, then the warps' partials in warp order, then the CTAs' partials in rank order
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
      s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
    }
  }
  const int warp = threadIdx.x >> 5, nwarps = T >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[warp * 16 + 2 * j] = s1[j];
      red[warp * 16 + 2 * j + 1] = s2[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float acc = 0.f;
    for (int w = 0; w < nwarps; ++w) acc += red[w * 16 + threadIdx.x];
    part[threadIdx.x] = acc;
  }
  cluster.sync();
  if (threadIdx.x < 16) {
    float acc = 0.f;
    for (int r = 0; r < KC; ++r) acc += *cluster.map_shared_rank(&part[threadIdx.x], r);
    tot[threadIdx.x] = (double)acc;
    if (crank == 0) sums[((size_t)n * Cp + (size_t)ch * 8) * 2 + threadIdx.x] = (double)acc;
  }
  cluster.sync();      // no CTA leaves (or its `part` dies) while a peer still reads it; publishes `tot` to the block
  // pass 2: dy = scale * (g - mean(g) - yhat * mean(g * yhat)) = g * A + y * B + C
  float ca[8], cb[8], cc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float mg = (float)(tot[2 * j] * inv_count), mgy = (float)(tot[2 * j + 1] * inv_count);
    ca[j] = scale[j];
    cb[j] = -scale[j] * scale[j] * mgy;
    cc[j] = -scale[j] * mg + scale[j] * scale[j] * mgy * mean[j];
  }
  for (int v = v0; v < V; v += vstep * IN_U) {
    uint4 rg[IN_U], ry[IN_U];
#pragma unroll
    for (int u = 0; u < IN_U; ++u) {
      const int vv = v + u * vstep;
      if (vv < V) {
        const size_t idx = base + (size_t)vv * chunks + ch;
        rg[u] = g != nullptr ? g[idx] : dout[idx];
        ry[u] = y[idx];
      }
    }
#pragma unroll
    for (int u = 0; u < IN_U; ++u) {
      const int vv = v + u * vstep;
      if (vv >= V) break;
      float gg[8], yy[8];
      unpack8(rg[u], gg, af);
      unpack8(ry[u], yy, af);
      if (g == nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gg[j] = (yy[j] - mean[j]) * ca[j] > 0.f ? gg[j] : LRELU * gg[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) gg[j] = fmaf(gg[j], ca[j], fmaf(yy[j], cb[j], cc[j]));
      dy[base + (size_t)vv * chunks + ch] = pack8(gg, af);
    }
  }
}

// per-channel sum over (n, voxels) of a bf16 NDHWC tensor (bias gradients)
__global__ void channel_sum_kernel(const uint4* __restrict__ x, double* __restrict__ dsum, int chunks, long long NV) {
  extern __shared__ float red[];
  const int ch = threadIdx.x;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.y + threadIdx.y; v < NV; v += (long long)gridDim.x * blockDim.y) {
    float f[8];
    unpack8(ld_stream(x + (size_t)v * chunks + ch), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  }
  const int C8 = chunks * 8;
  float* my = red + (size_t)threadIdx.y * C8 + ch * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) my[j] = acc[j];
  __syncthreads();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < C8; i += blockDim.x * blockDim.y) {
    float a = 0.f;
    for (int r = 0; r < (int)blockDim.y; ++r) a += red[(size_t)r * C8 + i];
    atomicAdd(&dsum[i], (double)a);
  }
}

// ---------------------------------------------------------------------------------------------
// Stem: Conv3d(Cin -> C, k3, p1) + bias on an fp32 NCDHW input (Cin = 1 in every reference script), bf16 NDHWC output
// (network.py:541,550 -- no norm / activation follows).  K = 27 Cin: HBM-bound, CUDA cores.
// ---------------------------------------------------------------------------------------------
template <int CP, bool ONE>
__global__ void stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w /*[Cin][27][CP]*/,
                                const float* __restrict__ b /*[CP]*/, bf16* __restrict__ out, int N, int Cin_rt, int D,
                                int H, int W, int af) {
  extern __shared__ float ws[];                  // Cin * 27 * CP weights, then CP biases
  const int Cin = ONE ? 1 : Cin_rt;              // the single-channel stem of every reference script: loop folded away
  const int nw = Cin * 27 * CP;
  for (int i = threadIdx.x; i < nw + CP; i += blockDim.x) ws[i] = i < nw ? w[i] : b[i - nw];
  __syncthreads();
  const long long V = (long long)D * H * W, total = (long long)N * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / V);
    const long long v = i - (long long)n * V;
    const int xw = (int)(v % W), xh = (int)((v / W) % H), xd = (int)(v / ((long long)W * H));
    float acc[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = ws[nw + c];
#pragma unroll 1
    for (int ci = 0; ci < Cin; ++ci) {
      const float* xn = x + ((size_t)n * Cin + ci) * V;
      const float* wc = ws + ci * 27 * CP;
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
        const int dd = xd + kd - 1;
        if (dd < 0 || dd >= D) continue;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int hh = xh + kh - 1;
          if (hh < 0 || hh >= H) continue;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int ww = xw + kw - 1;
            if (ww < 0 || ww >= W) continue;
            const float xv = __ldg(xn + ((size_t)dd * H + hh) * W + ww);
            const float* wt = wc + ((kd * 3 + kh) * 3 + kw) * CP;
#pragma unroll
            for (int c = 0; c < CP; ++c) acc[c] = fmaf(xv, wt[c], acc[c]);
          }
        }
      }
    }
    uint4* op = reinterpret_cast<uint4*>(out + (size_t)i * CP);
#pragma unroll
    for (int k = 0; k < CP / 8; ++k) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = acc[k * 8 + j];
      op[k] = pack8(f, af);
    }
  }
}

// Stem weight/bias gradient: dW[tap][c] = sum_v x[v + tap] * dy[v][c];  db[c] = sum_v dy[v][c].
// A thread owns (8-channel chunk, kd) and walks a run of SEG voxels along w: per voxel one 16-byte dy load, three
// new x values (the 3x3 (kh,kw) window slides along w in registers) and 72 FMAs into register accumulators
// [9 taps][8 channels] (+ 8 bias sums on the kd == 1 threads).  Block = (CP/8, 3, RUNS); the RUNS partials are
// reduced in shared memory, then one fp32 atomic per (tap, channel) and block into dw[28][CP].
// One input channel per launch: x points at that channel of sample 0, x_sN is the sample stride (Cin * D * H * W).
template <int CP>
__global__ void __launch_bounds__(CP / 8 * 3 * 16) stem_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dy,
                                                                    float* __restrict__ dw, int N, int D, int H, int W,
                                                                    long long x_sN, int af) {
  constexpr int SEG = 16, RUNS = 16;
  const int ch = threadIdx.x, kd = threadIdx.y, run = threadIdx.z;
  const int segs_w = (W + SEG - 1) / SEG;
  const long long n_runs = (long long)N * D * H * segs_w;
  float acc[9][8], bsum[8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
  for (long long r = (long long)blockIdx.x * RUNS + run; r < n_runs; r += (long long)gridDim.x * RUNS) {
    long long q = r;
    const int sw = (int)(q % segs_w); q /= segs_w;
    const int h = (int)(q % H); q /= H;
    const int d = (int)(q % D);
    const int n = (int)(q / D);
    const int w0 = sw * SEG;
    const int xd = d + kd - 1;
    const bool d_ok = xd >= 0 && xd < D;
    const float* xrow[3];
    bool row_ok[3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int xh = h + kh - 1;
      row_ok[kh] = d_ok && xh >= 0 && xh < H;
      xrow[kh] = x + (size_t)n * x_sN + ((size_t)(row_ok[kh] ? xd : 0) * H + (row_ok[kh] ? xh : 0)) * W;
    }
    float win[3][3];          // win[kh][kw] = x[.., w + kw - 1]
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      win[kh][1] = (row_ok[kh] && w0 - 1 >= 0) ? __ldg(xrow[kh] + w0 - 1) : 0.f;
      win[kh][2] = (row_ok[kh] && w0 < W) ? __ldg(xrow[kh] + w0) : 0.f;
    }
    const uint4* dyp = reinterpret_cast<const uint4*>(dy + ((size_t)((size_t)n * D + d) * H + h) * W * CP) + ch;
#pragma unroll 4
    for (int i = 0; i < SEG; ++i) {
      const int w = w0 + i;
      if (w >= W) break;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        win[kh][0] = win[kh][1];
        win[kh][1] = win[kh][2];
        win[kh][2] = (row_ok[kh] && w + 1 < W) ? __ldg(xrow[kh] + w + 1) : 0.f;
      }
      float f[8];
      unpack8(__ldg(dyp + (size_t)w * (CP / 8)), f, af);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[kh * 3 + kw][j] = fmaf(win[kh][kw], f[j], acc[kh * 3 + kw][j]);
      if (kd == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) bsum[j] += f[j];
      }
    }
  }
  __shared__ float red[RUNS][28 * CP / 4 + 1];      // reduced in 4 passes (one per 7-tap slice) to stay under 48 KB
  const int tid = (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y * blockDim.z;
#pragma unroll 1
  for (int pass = 0; pass < 4; ++pass) {      // taps [7 pass, 7 pass + 7); tap 27 = bias row
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int tap = kd * 9 + t;
      if (tap / 7 == pass) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[run][(tap - 7 * pass) * CP + ch * 8 + j] = acc[t][j];
      }
    }
    if (pass == 3 && kd == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[run][(27 - 21) * CP + ch * 8 + j] = bsum[j];
    }
    __syncthreads();
    for (int i = tid; i < 7 * CP; i += nthr) {
      float a = 0.f;
#pragma unroll
      for (int rr = 0; rr < RUNS; ++rr) a += red[rr][i];
      atomicAdd(&dw[pass * 7 * CP + i], a);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Stem weight/bias gradient on the (legacy, warp-level) tensor-core path: the contraction over voxels is a GEMM with
// M = 32 rows (27 taps, row 27 = ones -> the bias sum, rows 28..31 empty), N = CP channels, K = voxels:
//     dw[tap][c] = sum_v x[v + shift(tap)] * dy[v][c]
// A warp walks the 16-voxel runs of a w-row: the 3x3 window of fp32 x rows is staged in shared memory (gathering the A
// fragment straight from global cost ~12 L1 sector operations per voxel and bounded the kernel) and split into a 16-bit
// high and low part (two MMAs: the result keeps ~fp32 accuracy in x); the dy run (16 voxels x CP channels) goes
// through shared memory and ldmatrix.trans.  72 FMAs per voxel and thread of the CUDA-core
// kernel above become 2 x 2 x CP/8 mma.sync.m16n8k16 per 16 voxels and warp -- the kernel is left HBM-bound on dy.
// K = 27 is far too small for tcgen05 (M = 128 rows minimum): mma.sync is the right tool for this one layer.
// ---------------------------------------------------------------------------------------------
// Stage the 3x3 window of fp32 x rows around (d, h), columns [w0 - 4, w0 + 20), into shared memory with cp.async (row pitch 28
// floats; element (row rr, column jj) = x[d + rr/3 - 1][h + rr%3 - 1][w0 - 4 + jj], zero outside the volume).  16-byte pieces
// when W % 4 == 0 (54 per run instead of 9 x 18 four-byte ones: the staging was what bounded both stem kernels).
constexpr int STEM_XP = 28;
__device__ __forceinline__ void stem_stage_x(uint32_t dst, const float* __restrict__ xn, const float* __restrict__ any, int d, int h,
                                             int w0, int D, int H, int W, int lane) {
  if ((W & 3) == 0) {
    for (int j = lane; j < 9 * 6; j += 32) {
      const int rr = j / 6, pc = j - rr * 6;
      const int xd = d + rr / 3 - 1, xh = h + rr % 3 - 1, xw = w0 - 4 + 4 * pc;
      const bool ok = xd >= 0 && xd < D && xh >= 0 && xh < H && xw >= 0 && xw < W;
      const void* src = ok ? (const void*)(xn + ((size_t)xd * H + xh) * W + xw) : (const void*)any;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (rr * STEM_XP + 4 * pc) * 4), "l"(src), "r"(ok ? 16 : 0));
    }
  } else {
    for (int j = lane; j < 9 * 18; j += 32) {
      const int rr = j / 18, jj = j - rr * 18 + 3;
      const int xd = d + rr / 3 - 1, xh = h + rr % 3 - 1, xw = w0 - 4 + jj;
      const bool ok = xd >= 0 && xd < D && xh >= 0 && xh < H && xw >= 0 && xw < W;
      const void* src = ok ? (const void*)(xn + ((size_t)xd * H + xh) * W + xw) : (const void*)any;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + (rr * STEM_XP + jj) * 4), "l"(src), "r"(ok ? 4 : 0));
    }
  }
}

template <bool F16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CP, bool F16>
__global__ void __launch_bounds__(256) stem_wgrad_mma_kernel(const float* __restrict__ x, const bf16* __restrict__ dy,
                                                            float* __restrict__ dw, int N, int D, int H, int W,
                                                            long long x_sN) {
  constexpr int NT = CP / 8;                       // n-tiles of 8 channels
  constexpr int PITCH = CP * 2 + 16;               // bytes per voxel row of the staged dy run (+16: ldmatrix bank spread)
  extern __shared__ __align__(16) uint8_t stem_smem[];
  typedef uint8_t (*SdyT)[2][16 * PITCH];
  typedef float (*SxT)[2][10 * STEM_XP];
  SdyT sdy = reinterpret_cast<SdyT>(stem_smem);                                  // per warp, double buffered: the dy run
  SxT sx = reinterpret_cast<SxT>(stem_smem + 8 * 2 * 16 * PITCH);                // ... the 3x3 window of x rows + a zero row
  float* sred = reinterpret_cast<float*>(stem_smem + 8 * 2 * 16 * PITCH + 8 * 2 * 10 * STEM_XP * 4);     // [32][CP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;       // fragment coordinates
  for (int i = threadIdx.x; i < 32 * CP; i += blockDim.x) sred[i] = 0.f;
  for (int i = threadIdx.x; i < 8 * 2 * STEM_XP; i += blockDim.x)               // row 9 of every window buffer: zeros, never staged
    sx[i / (2 * STEM_XP)][(i / STEM_XP) & 1][9 * STEM_XP + i % STEM_XP] = 0.f;
  __syncthreads();
  float acc[2][NT][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[mt][nt][r] = 0.f;
  const long long n_rows = (long long)N * D * H;
  const int cpr = (W + 15) / 16;                   // 16-voxel runs per w-row
  // this thread's four taps as (row of the staged 3x3 x-window, kw): row 9 = zeros (taps 28..31), tap 27 = ones
  int trow[4], tkw[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int tap = gid + 8 * r;
    trow[r] = tap < 27 ? tap / 3 : 9;
    tkw[r] = tap < 27 ? tap % 3 : 0;
  }
  // the warp's runs form one sequence (row-major over its rows); run i+1 is fetched with cp.async into the other half of
  // the staging buffers while run i is multiplied -- without this every run paid two global-load latencies in series
  // 32-bit index arithmetic (the launcher refuses volumes with 2^31 runs or more): 64-bit divisions cost ~50 instructions each
  const int row0 = blockIdx.x * 8 + warp, row_step = gridDim.x * 8;
  const int my_rows = row0 < (int)n_rows ? ((int)n_rows - row0 + row_step - 1) / row_step : 0;
  const int my_runs = my_rows * cpr;
  auto fetch = [&](int i, int buf) {
    const int ri = i / cpr;
    const int row = row0 + ri * row_step;
    const int w0 = (i - ri * cpr) * 16;
    int q = row;
    const int h = q % H; q /= H;
    const int d = q % D;
    const int n = q / D;
    const float* xn = x + (size_t)n * x_sN;
    const bf16* drow = dy + ((((size_t)n * D + d) * H + h) * W + w0) * CP;
    const uint32_t dst_y = smem_u32(sdy[warp][buf]), dst_x = smem_u32(sx[warp][buf]);
    for (int j = lane; j < 16 * NT; j += 32) {                 // dy run: 16 voxels x CP channels, rows past W zero-filled
      const int v = j / NT, c8 = j - v * NT;
      const bool ok = w0 + v < W;
      const void* src = ok ? (const void*)(drow + (size_t)j * 8) : (const void*)dy;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_y + v * PITCH + c8 * 16), "l"(src), "r"(ok ? 16 : 0));
    }
    stem_stage_x(dst_x, xn, x, d, h, w0, D, H, W, lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (my_runs > 0) fetch(0, 0);
  for (int i = 0; i < my_runs; ++i) {
    const int buf = i & 1;
    if (i + 1 < my_runs) {
      fetch(i + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const float* myx = sx[warp][buf];
    const uint32_t my_s = smem_u32(sdy[warp][buf]);
    // ---- A fragments: x[v + shift(tap)] for this thread's 4 taps x 4 voxels (k = 2 tig, 2 tig + 1, + 8, + 9)
    uint32_t ahi[2][4], alo[2][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {                  // row block r: tap = gid + 8 r -> m-tile r / 2, fragment regs (r & 1) + {0, 2}
      const float* xr = myx + trow[r] * STEM_XP + tkw[r] + 3;
      float hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = 2 * tig + (j & 1) + 8 * (j >> 1);
        float t = xr[k];
        if (gid + 8 * r == 27) t = 1.f;            // the bias row: sum of dy (rows of dy past W are zero)
        hi[j] = F16 ? __half2float(__float2half_rn(t)) : __bfloat162float(__float2bfloat16_rn(t));
        lo[j] = t - hi[j];
      }
      const int mt = r >> 1, base = r & 1;         // a0/a2 belong to row gid, a1/a3 to row gid + 8 of the m-tile
      ahi[mt][base] = pack_2x16(hi[0], hi[1], F16);
      ahi[mt][base + 2] = pack_2x16(hi[2], hi[3], F16);
      alo[mt][base] = pack_2x16(lo[0], lo[1], F16);
      alo[mt][base + 2] = pack_2x16(lo[2], lo[3], F16);
    }
    // ---- B fragments from shared memory (ldmatrix.trans: thread gets (k = 2 tig + {0,1}, n = gid)) and the MMAs
#pragma unroll
    for (int nt = 0; nt < NT; nt += 2) {
      // four 8x8 matrices: (voxels 0-7, tile nt), (voxels 8-15, tile nt), (voxels 0-7, tile nt+1), (voxels 8-15, tile nt+1)
      const int mat = lane >> 3, rrow = lane & 7;
      const uint32_t addr = my_s + (uint32_t)((rrow + 8 * (mat & 1)) * PITCH + (nt + (mat >> 1)) * 16);
      uint32_t b[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(addr));
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma_16816<F16>(acc[mt][nt], ahi[mt], b[0], b[1]);
        mma_16816<F16>(acc[mt][nt], alo[mt], b[0], b[1]);
        mma_16816<F16>(acc[mt][nt + 1], ahi[mt], b[2], b[3]);
        mma_16816<F16>(acc[mt][nt + 1], alo[mt], b[2], b[3]);
      }
    }
    __syncwarp();                                  // everyone is done with this half before the run after next lands in it
  }
  // ---- CTA reduction in shared memory, then one atomic per (tap row < 28, channel)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int row = mt * 16 + gid, col = nt * 8 + 2 * tig;
      atomicAdd(&sred[row * CP + col], acc[mt][nt][0]);
      atomicAdd(&sred[row * CP + col + 1], acc[mt][nt][1]);
      atomicAdd(&sred[(row + 8) * CP + col], acc[mt][nt][2]);
      atomicAdd(&sred[(row + 8) * CP + col + 1], acc[mt][nt][3]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 28 * CP; i += blockDim.x) atomicAdd(&dw[i], sred[i]);
}

// ---------------------------------------------------------------------------------------------
// Single-channel stem forward on mma.sync: out[v][c] = b[c] + sum_tap x[v + shift(tap)] * w[tap][c] as a GEMM with
// M = 16 voxels of a w-run, K = 32 (27 taps, tap 27 = ones against the bias row, 28..31 empty), N = CP channels.
// x and w are split into 16-bit high and low parts (hi*hi + hi*lo + lo*hi: fp32-grade products, like the FMA kernel).
// The 3x3 window of x rows arrives by cp.async one run ahead; the 16 x CP output run is staged in shared memory and leaves as
// one contiguous 16-bit NDHWC block (16-byte stores).  HBM-bound on the output; also the first kernel of every window forward.
// ---------------------------------------------------------------------------------------------
template <int CP, bool F16>
__global__ void __launch_bounds__(256) stem_fwd_mma_kernel(const float* __restrict__ x, const float* __restrict__ w /*[27][CP]*/,
                                                          const float* __restrict__ b /*[CP]*/, bf16* __restrict__ out,
                                                          int N, int D, int H, int W) {
  constexpr int NT = CP / 8;
  constexpr int OPITCH = CP * 2 + 16;              // bytes per voxel row of the staged output run
  __shared__ __align__(16) float sx[8][2][10 * STEM_XP];
  __shared__ __align__(16) uint8_t so[8][16 * OPITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  // ---- B fragments (taps x channels), loop invariant: b0 = (k = 2 tig + {0,1}, n = gid), b1 = k + 8; two k-steps
  uint32_t bhi[2][NT][2], blo[2][NT][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float v[2], hi[2], lo[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int tap = ks * 16 + r * 8 + 2 * tig + j, c = nt * 8 + gid;
          v[j] = tap < 27 ? __ldg(&w[tap * CP + c]) : (tap == 27 ? __ldg(&b[c]) : 0.f);
          hi[j] = F16 ? __half2float(__float2half_rn(v[j])) : __bfloat162float(__float2bfloat16_rn(v[j]));
          lo[j] = v[j] - hi[j];
        }
        bhi[ks][nt][r] = pack_2x16(hi[0], hi[1], F16);
        blo[ks][nt][r] = pack_2x16(lo[0], lo[1], F16);
      }
  // this thread's A columns: taps ks*16 + {2 tig, 2 tig + 1, 2 tig + 8, 2 tig + 9} as (window row, kw); 27 = ones, > 27 zero row
  int trow[2][4], tkw[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tap = ks * 16 + 2 * tig + (j & 1) + 8 * (j >> 1);
      trow[ks][j] = tap < 27 ? tap / 3 : 9;
      tkw[ks][j] = tap < 27 ? tap % 3 : 0;
    }
  for (int i = threadIdx.x; i < 8 * 2 * STEM_XP; i += blockDim.x)               // row 9 of every window buffer: zeros, never staged
    sx[i / (2 * STEM_XP)][(i / STEM_XP) & 1][9 * STEM_XP + i % STEM_XP] = 0.f;
  __syncthreads();
  const long long n_rows = (long long)N * D * H;
  const int cpr = (W + 15) / 16;
  // 32-bit index arithmetic (the launcher refuses volumes with 2^31 runs or more): 64-bit divisions cost ~50 instructions each
  const int row0 = blockIdx.x * 8 + warp, row_step = gridDim.x * 8;
  const int my_rows = row0 < (int)n_rows ? ((int)n_rows - row0 + row_step - 1) / row_step : 0;
  const int my_runs = my_rows * cpr;
  auto fetch = [&](int i, int buf) {
    const int ri = i / cpr;
    const int row = row0 + ri * row_step;
    const int w0 = (i - ri * cpr) * 16;
    int q = row;
    const int h = q % H; q /= H;
    const int d = q % D;
    const int n = q / D;
    const float* xn = x + (size_t)n * D * H * W;
    const uint32_t dst_x = smem_u32(sx[warp][buf]);
    stem_stage_x(dst_x, xn, x, d, h, w0, D, H, W, lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (my_runs > 0) fetch(0, 0);
  for (int i = 0; i < my_runs; ++i) {
    const int buf = i & 1;
    if (i + 1 < my_runs) {
      fetch(i + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const float* myx = sx[warp][buf];
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[nt][r] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      // A fragment: a0 = (voxel gid, taps 2 tig + {0,1}), a1 = voxel gid + 8, a2 / a3 = taps + 8
      uint32_t ahi[4], alo[4];
#pragma unroll
      for (int vr = 0; vr < 2; ++vr) {
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = myx[trow[ks][j] * STEM_XP + tkw[ks][j] + 3 + gid + 8 * vr];
          if (ks == 1 && 2 * tig + (j & 1) + 8 * (j >> 1) == 11) t = 1.f;      // tap 27: the bias row of B
          hi[j] = F16 ? __half2float(__float2half_rn(t)) : __bfloat162float(__float2bfloat16_rn(t));
          lo[j] = t - hi[j];
        }
        ahi[vr] = pack_2x16(hi[0], hi[1], F16);
        ahi[vr + 2] = pack_2x16(hi[2], hi[3], F16);
        alo[vr] = pack_2x16(lo[0], lo[1], F16);
        alo[vr + 2] = pack_2x16(lo[2], lo[3], F16);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        mma_16816<F16>(acc[nt], ahi, bhi[ks][nt][0], bhi[ks][nt][1]);
        mma_16816<F16>(acc[nt], ahi, blo[ks][nt][0], blo[ks][nt][1]);
        mma_16816<F16>(acc[nt], alo, bhi[ks][nt][0], bhi[ks][nt][1]);
      }
    }
    // ---- stage the 16 x CP run (c0,c1 = voxel gid, channels nt*8 + 2 tig + {0,1}; c2,c3 = voxel gid + 8), store coalesced
    uint8_t* myo = so[warp];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      *reinterpret_cast<uint32_t*>(myo + gid * OPITCH + (nt * 8 + 2 * tig) * 2) = pack_2x16(acc[nt][0], acc[nt][1], F16);
      *reinterpret_cast<uint32_t*>(myo + (gid + 8) * OPITCH + (nt * 8 + 2 * tig) * 2) = pack_2x16(acc[nt][2], acc[nt][3], F16);
    }
    __syncwarp();
    {
      const int ri = i / cpr;
      const int row = row0 + ri * row_step;
      const int w0 = (i - ri * cpr) * 16;
      uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)row * W + w0) * CP);
      for (int j = lane; j < 16 * NT; j += 32) {
        const int v = j / NT, c8 = j - v * NT;
        if (w0 + v < W) dst[j] = *reinterpret_cast<const uint4*>(myo + v * OPITCH + c8 * 16);
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Head: Conv3d(C -> K, k1) + bias (network.py:545-547,563): bf16 NDHWC in, fp32 NCDHW logits out.
// ---------------------------------------------------------------------------------------------
template <int CP, int KMAX>
__global__ void head_fwd_kernel(const bf16* __restrict__ a, const float* __restrict__ w /*[K][CP]*/,
                                const float* __restrict__ b, float* __restrict__ logits, int K, int N, long long V,
                                int af) {
  __shared__ float ws[KMAX * CP + KMAX];
  for (int i = threadIdx.x; i < K * CP + K; i += blockDim.x) ws[i] = i < K * CP ? w[i] : b[i - K * CP];
  __syncthreads();
  // one thread = one voxel: its CP / 8 16-byte loads are issued together; samples in an outer loop (no 64-bit division)
  for (int n = 0; n < N; ++n) {
    const bf16* an = a + (size_t)n * V * CP;
    float* ln = logits + (size_t)n * K * V;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
      float acc[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) acc[k] = k < K ? ws[K * CP + k] : 0.f;
      const uint4* ap = reinterpret_cast<const uint4*>(an + (size_t)v * CP);
      uint4 ra[CP / 8];
#pragma unroll
      for (int c8 = 0; c8 < CP / 8; ++c8) ra[c8] = ld_stream(ap + c8);
#pragma unroll
      for (int c8 = 0; c8 < CP / 8; ++c8) {
        float f[8];
        unpack8(ra[c8], f, af);
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[k] = fmaf(f[j], ws[k * CP + c8 * 8 + j], acc[k]);
          }
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) ln[(size_t)k * V + v] = acc[k];
    }
  }
}

// da[v][c] = sum_k dl[k][v] w[k][c];  dW[k][c] += sum_v dl[k][v] a[v][c];  db[k] += sum_v dl[k][v]
// Each thread keeps its own K x CP partial of dW in registers over all its voxels; one shuffle + shared
// reduction per block at the end.
template <int CP, int KMAX>
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dl, const bf16* __restrict__ a,
                                                       const float* __restrict__ w, bf16* __restrict__ da,
                                                       float* __restrict__ dw /*[K][CP]+[K]*/,
                                                       const float* __restrict__ gscale, int K, int N, long long V,
                                                       int af) {
  // A thread owns (voxel, 8-channel chunk): 16-byte loads / stores that are consecutive across the lanes of a warp and
  // 4 x 8 weight-gradient accumulators per thread (the former one-voxel-per-thread layout held 4 x 32 of them in 168
  // registers: 15 % occupancy, 21 % of the HBM bandwidth).
  constexpr int CH = CP / 8;
  __shared__ float ws[KMAX * CP];
  __shared__ float red[KMAX * CP + KMAX];
  for (int i = threadIdx.x; i < KMAX * CP; i += blockDim.x) ws[i] = i < K * CP ? w[i] : 0.f;
  for (int i = threadIdx.x; i < KMAX * CP + KMAX; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const float gs = gscale ? __ldg(gscale) : 1.f;      // internal gradient scale (fp16 mode); dW / db stay unscaled
  const int ch = threadIdx.x % CH, vin = threadIdx.x / CH, vpb = blockDim.x / CH;
  float wk[KMAX][8], pw[KMAX][8], pb[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    pb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      pw[k][j] = 0.f;
      wk[k][j] = ws[k * CP + ch * 8 + j];
    }
  }
  // HB_U voxels per thread and trip, all loads issued before the first use (the one-voxel loop ran at 2.2 TB/s: two
  // dependent global loads and a 64-bit division per trip); samples are walked in an outer loop, so no division at all
  constexpr int HB_U = 4;
  const long long stride = (long long)gridDim.x * vpb;
  for (int n = 0; n < N; ++n) {
    const float* dln = dl + (size_t)n * K * V;
    const bf16* an = a + (size_t)n * V * CP;
    bf16* dan = da + (size_t)n * V * CP;
    for (long long v0 = (long long)blockIdx.x * vpb + vin; v0 < V; v0 += stride * HB_U) {
      float g[HB_U][KMAX];
      uint4 ra[HB_U];
#pragma unroll
      for (int u = 0; u < HB_U; ++u) {
        const long long v = v0 + u * stride;
        if (v < V) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k) g[u][k] = k < K ? __ldg(&dln[(size_t)k * V + v]) : 0.f;
          ra[u] = ld_stream(reinterpret_cast<const uint4*>(an + (size_t)v * CP) + ch);
        }
      }
#pragma unroll
      for (int u = 0; u < HB_U; ++u) {
        const long long v = v0 + u * stride;
        if (v >= V) break;
        float f[8], t[8];
        unpack8(ra[u], f, af);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float sacc = 0.f;
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            sacc = fmaf(g[u][k], wk[k][j], sacc);              // rows k >= K of the weights are zero
            pw[k][j] = fmaf(g[u][k], f[j], pw[k][j]);
          }
          t[j] = sacc * gs;
        }
        reinterpret_cast<uint4*>(dan + (size_t)v * CP)[ch] = pack8(t, af);
        if (ch == 0) {
#pragma unroll
          for (int k = 0; k < KMAX; ++k) pb[k] += g[u][k];
        }
      }
    }
  }
  // lanes with equal (lane % CH) hold the same chunk (CH divides 32): butterfly over the other lane bits
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sv = pw[k][j];
#pragma unroll
      for (int o = 16; o >= CH; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
      if (lane < CH) atomicAdd(&red[k * CP + lane * 8 + j], sv);
    }
    float sb = pb[k];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sb += __shfl_xor_sync(0xffffffffu, sb, o);
    if (lane == 0) atomicAdd(&red[KMAX * CP + k], sb);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * CP; i += blockDim.x) atomicAdd(&dw[i], red[i]);
  for (int i = threadIdx.x; i < K; i += blockDim.x) atomicAdd(&dw[K * CP + i], red[KMAX * CP + i]);
}

// ---------------------------------------------------------------------------------------------
// Loss: softmax over classes fused with the batch Tversky-Dice and focal/CE sums
// (loss.py:7-11 softmax, :32-48 dice, :70-80 focal; one-hot never materialised).
//   sums[c] = { TP = sum p_c g_c,  SP = sum p_c,  SG = sum g_c,  F = sum g_c * -(1-p_c)^gamma * log p_c }
// ---------------------------------------------------------------------------------------------
template <int KMAX>
__device__ __forceinline__ void softmax_k(const float (&z)[KMAX], int K, float (&p)[KMAX], float& lse) {
  float m = z[0];
#pragma unroll
  for (int k = 1; k < KMAX; ++k)
    if (k < K) m = fmaxf(m, z[k]);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    p[k] = k < K ? expf(z[k] - m) : 0.f;
    s += p[k];
  }
  const float inv = 1.f / s;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) p[k] *= inv;
  lse = m + logf(s);
}

template <int KMAX>
__global__ void loss_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                double* __restrict__ sums /*[K][4]*/, int K, int N, long long V, float gamma) {
  // Reproducible sums: fixed per-thread order, shuffle tree, then the warps' partials are added in warp order in fp64;
  // only the cross-block fp64 atomics are unordered (1e-16 relative, invisible after the fp32 rounding of coef).  A
  // float atomic here made dlogits differ in the last bit from run to run, and the 16-bit backward chain re-rounds
  // that into ~0.5 % of the level-0 gradients (profiles/r02_notes.md, "run-to-run reproducibility").
  __shared__ float red[32][KMAX * 4];
  float tp[KMAX], sp[KMAX], sg[KMAX], fo[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) tp[k] = sp[k] = sg[k] = fo[k] = 0.f;
  const long long total = (long long)N * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / V);
    const long long v = i - (long long)n * V;
    float z[KMAX], p[KMAX], lse;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) z[k] = k < K ? __ldg(&logits[((size_t)n * K + k) * V + v]) : 0.f;
    softmax_k<KMAX>(z, K, p, lse);
    const int t = (int)target[i];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k >= K) continue;
      sp[k] += p[k];
      if (k == t) {
        tp[k] += p[k];
        sg[k] += 1.f;
        const float logpt = z[k] - lse;
        const float om = 1.f - p[k];
        const float mod = gamma == 0.f ? 1.f : (gamma == 2.f ? om * om : powf(fmaxf(om, 0.f), gamma));
        fo[k] += -mod * logpt;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k >= K) continue;
    float a = tp[k], b = sp[k], c = sg[k], d = fo[k];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
      d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if ((threadIdx.x & 31) == 0) {
      float* r = red[threadIdx.x >> 5];
      r[k * 4 + 0] = a;
      r[k * 4 + 1] = b;
      r[k * 4 + 2] = c;
      r[k * 4 + 3] = d;
    }
  }
  __syncthreads();
  const int warps = (blockDim.x + 31) >> 5;
  for (int i = threadIdx.x; i < K * 4; i += blockDim.x) {
    double t = 0.0;
    for (int w = 0; w < warps; ++w) t += (double)red[w][i];
    atomicAdd(&sums[i], t);
  }
}

// coef[k] = { a_k = dL/dTP_k, b_k = dL/dSP_k, f_k = w_k * K / (N V) (focal weight), unused };
// dL/dz_j = p_j (q_j - sum_c p_c q_c),  q_c = a_c g_c + b_c + [c == t] f_t * fprime(p_t)
template <int KMAX>
__global__ void loss_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                const float* __restrict__ coef /*[K][4]*/, const float* __restrict__ gscale,
                                float* __restrict__ dlogits, int K, int N, long long V, float gamma, int use_focal) {
  __shared__ float cf[KMAX * 4];
  for (int i = threadIdx.x; i < K * 4; i += blockDim.x) cf[i] = coef[i];
  __syncthreads();
  const float gs = gscale ? gscale[0] : 1.f;
  const long long total = (long long)N * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / V);
    const long long v = i - (long long)n * V;
    float z[KMAX], p[KMAX], q[KMAX], lse;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) z[k] = k < K ? __ldg(&logits[((size_t)n * K + k) * V + v]) : 0.f;
    softmax_k<KMAX>(z, K, p, lse);
    const int t = (int)target[i];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k >= K) { q[k] = 0.f; continue; }
      float qq = cf[k * 4 + 1];
      if (k == t) {
        qq += cf[k * 4 + 0];
        if (use_focal) {
          const float pt = fmaxf(p[k], 1e-30f);
          const float logpt = z[k] - lse;
          const float om = 1.f - pt;
          float fp;   // d/dp [ -(1-p)^gamma log p ]
          if (gamma == 0.f) fp = -1.f / pt;
          else if (gamma == 2.f) fp = 2.f * om * logpt - om * om / pt;
          else fp = gamma * powf(fmaxf(om, 0.f), gamma - 1.f) * logpt - powf(fmaxf(om, 0.f), gamma) / pt;
          qq += cf[k * 4 + 2] * fp;
        }
      }
      q[k] = qq;
      dot += p[k] * qq;
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) dlogits[((size_t)n * K + k) * V + v] = gs * p[k] * (q[k] - dot);
  }
}

// ---------------------------------------------------------------------------------------------
// Sliding-window blend (trainer.py:72-96): result[:, tile] += softmax(logits) * w ; weight[tile] += w
// then result / weight -> argmax (uint8) or probabilities.  w == null is the reference's uniform blend.
//
// The sums are kept in 64-bit FIXED POINT (scale 2^54): every fp32 term p * w converts exactly, integer addition is
// associative, so the sums -- and with them the label map -- do not depend on the order in which windows are added:
// the windows of one volume can be dealt to any number of GPUs and reduced in any order (NCCL reduce-scatter of
// int64) and give bit-identical labels.  (The reference adds fp32 in window order; its label map differs from the
// exact one only where two class probabilities tie to ~1e-7.)  Up to 512 overlapping windows fit below 2^63.
//
// Layout of `acc`: [n_slab][K + 1][Xs][Y][Z] int64, x = slab * Xs + xs; channel K is the weight sum.  One slab
// (Xs = X) in a single-process run; with R ranks Xs = ceil(X / R) and slab r is what rank r owns after the
// reduce-scatter.
// ---------------------------------------------------------------------------------------------
constexpr float SW_FIXED_SCALE = 18014398509481984.f;        // 2^54

template <int KMAX>
__global__ void sw_accumulate_kernel(const float* __restrict__ logits /*[K][px][py][pz]*/,
                                     const float* __restrict__ window /*[px][py][pz] or null*/,
                                     long long* __restrict__ acc, int K, int px, int py, int pz, int x0, int y0, int z0,
                                     int Xs, int Y, int Z) {
  const long long P = (long long)px * py * pz;
  const size_t YZ = (size_t)Y * Z;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
    const int iz = (int)(i % pz), iy = (int)((i / pz) % py), ix = (int)(i / ((long long)pz * py));
    float z[KMAX], p[KMAX], lse;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) z[k] = k < K ? __ldg(&logits[(size_t)k * P + i]) : 0.f;
    if (K == 1) {
      p[0] = 1.f / (1.f + expf(-z[0]));      // trainer.py:76 sigmoid branch
    } else {
      softmax_k<KMAX>(z, K, p, lse);
    }
    const float wv = window ? __ldg(&window[i]) : 1.f;
    const int x = x0 + ix, slab = x / Xs, xs = x - slab * Xs;
    long long* a = acc + ((size_t)slab * (K + 1) * Xs + xs) * YZ + (size_t)(y0 + iy) * Z + (z0 + iz);
    const size_t cstride = (size_t)Xs * YZ;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) a[(size_t)k * cstride] += __float2ll_rn(p[k] * wv * SW_FIXED_SCALE);
    a[(size_t)K * cstride] += __float2ll_rn(wv * SW_FIXED_SCALE);
  }
}

template <int KMAX>
__global__ void sw_finalize_kernel(const long long* __restrict__ acc /*[K + 1][n]*/, uint8_t* __restrict__ labels,
                                   float* __restrict__ probs /*[n][K] or null*/, int K, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long wsum = acc[(size_t)K * n + i];
    long long s[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) s[k] = k < K ? acc[(size_t)k * n + i] : 0;
    if (probs) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) probs[(size_t)i * K + k] = wsum == 0 ? __int_as_float(0x7fc00000) /* 0/0 = NaN where uncovered */
                                                         : (float)((double)s[k] / (double)wsum);
    }
    if (labels) {
      // argmax(softmax(result / weight)) == argmax of the integer sums (same positive divisor): first maximum wins like
      // torch.argmax; uncovered voxels (NaN in the reference) -> index 0
      int best = 0;
      if (K == 1) {
        best = wsum == 0 ? 0 : (int)rintf((float)((double)s[0] / (double)wsum));      // trainer.py:91-96: squeeze + np.round
      } else if (wsum != 0) {
        long long bv = s[0];
#pragma unroll
        for (int k = 1; k < KMAX; ++k)
          if (k < K && s[k] > bv) { bv = s[k]; best = k; }
      }
      labels[i] = (uint8_t)best;
    }
  }
}

// Weight packing: out[i] = 16-bit( idx[i] < 0 ? 0 : w[idx[i]] ).  Turns the fp32 PyTorch-layout parameter into the
// tile stream the conv kernel's weight ring consumes (one pass: 4 B index + 4 B gather + 2 B store per element).
__global__ void weight_pack_kernel(const float* __restrict__ w, const int* __restrict__ idx, uint16_t* __restrict__ out,
                                   long long n, int f16) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += (long long)gridDim.x * blockDim.x * 8) {
    if (i + 8 <= n) {
      const int4 a = __ldg(reinterpret_cast<const int4*>(idx + i));
      const int4 b = __ldg(reinterpret_cast<const int4*>(idx + i + 4));
      const int id[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = id[j] < 0 ? 0.f : __ldg(w + id[j]);
      *reinterpret_cast<uint4*>(out + i) = pack8(f, f16);
    } else {
      for (long long k = i; k < n; ++k) {
        const int id = idx[k];
        const float v = id < 0 ? 0.f : w[id];
        out[k] = (uint16_t)(pack_2x16(v, 0.f, f16) & 0xffffu);
      }
    }
  }
}

// grid.x for the (voxel-row, sample) kernels: ~2 voxel rows per thread for small tensors (a thread that loops over
// 16-32 rows of a 1 MB tensor is DRAM-latency bound: ~20 us launches at levels 3-4), capped at `waves` CTA waves over
// the whole (gx, N) grid for large ones.
inline int grid_rows(long long V, int rows_per_block, int N, int num_sms, int waves) {
  // blocks per sample: at least one IN_U-voxel trip per thread, at most `waves` resident blocks per SM over all samples
  // (few long-lived blocks: the per-thread prologue -- tables, fp64 sums -- and the per-block atomics are not free)
  long long need = (V + (long long)IN_U * rows_per_block - 1) / ((long long)IN_U * rows_per_block);
  long long cap = ((long long)num_sms * waves + N - 1) / N;
  long long g = need < cap ? need : cap;
  return (int)(g < 1 ? 1 : g);
}
inline int grid_for(long long work_items, int per_block, int num_sms, int waves) {
  long long need = (work_items + per_block - 1) / per_block;
  long long cap = (long long)num_sms * waves;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}


// ---------------------------------------------------------------------------------------------
// MaxPoolBlock: nn.MaxPool3d(kernel_size=2, stride=2) (network.py:452-463) on 16-bit NDHWC.
// One thread = one output voxel x 8 channels.  The window is scanned in PyTorch's order (kd, kh, kw)
// with `v > best || isnan(v)`, so the first maximum wins and a NaN propagates; the winner is kept as
// a 3-bit code kd*4 + kh*2 + kw (PyTorch's flat index = ((2d+kd)*H + 2h+kh)*W + 2w+kw).
// ---------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, uint2* __restrict__ code,
                                   int chunks, int Do, int Ho, int Wo, long long total /* N*Do*Ho*Wo */, int af) {
  const int ch = threadIdx.x;
  const int H = 2 * Ho, W = 2 * Wo;
  for (long long v = (long long)blockIdx.x * blockDim.y + threadIdx.y; v < total; v += (long long)gridDim.x * blockDim.y) {
    const int w = (int)(v % Wo);
    long long t = v / Wo;
    const int h = (int)(t % Ho);
    t /= Ho;
    const int d = (int)(t % Do);
    const long long n = t / Do;
    const size_t in0 = ((((size_t)n * 2 * Do + 2 * d) * H + 2 * h) * W + 2 * w) * chunks + ch;
    float best[8];
    unsigned arg[8];
    uint4 raw[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      raw[k] = ld_stream(x + in0 + ((size_t)(k >> 2) * H * W + (size_t)((k >> 1) & 1) * W + (k & 1)) * chunks);
    unpack8(raw[0], best, af);
#pragma unroll
    for (int j = 0; j < 8; ++j) arg[j] = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      float f[8];
      unpack8(raw[k], f, af);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (f[j] > best[j] || f[j] != f[j]) {
          best[j] = f[j];
          arg[j] = k;
        }
    }
    out[(size_t)v * chunks + ch] = pack8(best, af);
    uint2 c;
    c.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
    c.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
    code[(size_t)v * chunks + ch] = c;
  }
}

// dx[window position == code] = dout, 0 elsewhere (windows do not overlap: every input voxel is written once).
__global__ void maxpool_bwd_kernel(const uint4* __restrict__ dout, const uint2* __restrict__ code, uint4* __restrict__ dx,
                                   int chunks, int Do, int Ho, int Wo, long long total) {
  const int ch = threadIdx.x;
  const int H = 2 * Ho, W = 2 * Wo;
  for (long long v = (long long)blockIdx.x * blockDim.y + threadIdx.y; v < total; v += (long long)gridDim.x * blockDim.y) {
    const int w = (int)(v % Wo);
    long long t = v / Wo;
    const int h = (int)(t % Ho);
    t /= Ho;
    const int d = (int)(t % Do);
    const long long n = t / Do;
    const size_t in0 = ((((size_t)n * 2 * Do + 2 * d) * H + 2 * h) * W + 2 * w) * chunks + ch;
    const uint4 g = ld_stream(dout + (size_t)v * chunks + ch);
    const uint2 c = code[(size_t)v * chunks + ch];
    const unsigned gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned cw = q < 2 ? c.x : c.y;
        const unsigned a0 = (cw >> ((q & 1) * 16)) & 0xffu, a1 = (cw >> ((q & 1) * 16 + 8)) & 0xffu;
        o[q] = (a0 == (unsigned)k ? (gw[q] & 0xffffu) : 0u) | (a1 == (unsigned)k ? (gw[q] & 0xffff0000u) : 0u);
      }
      dx[in0 + ((size_t)(k >> 2) * H * W + (size_t)((k >> 1) & 1) * W + (k & 1)) * chunks] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// AttBlock (network.py:353-371): out = xs * sigmoid(z) with xs = conv(skip), z = conv(lrelu(conv(skip)+conv(gate)))
// (the three 1x1x1 convs run on the tensor-core kernel; these are the gate's pointwise parts).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void att_gate_fwd_kernel(const uint4* __restrict__ xs, const uint4* __restrict__ z, uint4* __restrict__ out,
                                    long long n16, int af) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    unpack8(ld_stream(xs + i), a, af);
    unpack8(ld_stream(z + i), b, af);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= sigmoidf_(b[j]);
    out[i] = pack8(a, af);
  }
}

// dxs = dout * r, dz = dout * xs * r (1 - r), r = sigmoid(z); sums[c] = {sum dxs, sum dz} over (n, voxels)
__global__ void att_gate_bwd_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ xs, const uint4* __restrict__ z,
                                    uint4* __restrict__ dxs, uint4* __restrict__ dz, double* __restrict__ sums,
                                    int chunks, long long NV, int af) {
  extern __shared__ float red[];   // [blockDim.y][chunks*8][2]
  const int ch = threadIdx.x;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.y + threadIdx.y; v < NV; v += (long long)gridDim.x * blockDim.y) {
    const size_t idx = (size_t)v * chunks + ch;
    float d[8], a[8], b[8], e[8];
    unpack8(ld_stream(dout + idx), d, af);
    unpack8(ld_stream(xs + idx), a, af);
    unpack8(ld_stream(z + idx), b, af);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float r = sigmoidf_(b[j]);
      e[j] = d[j] * a[j] * r * (1.f - r);
      d[j] = d[j] * r;
    }
    const uint4 o1 = pack8(d, af), o2 = pack8(e, af);
    dxs[idx] = o1;
    dz[idx] = o2;
    unpack8(o1, d, af);
    unpack8(o2, e, af);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] += d[j];
      s2[j] += e[j];
    }
  }
  const int C8 = chunks * 8;
  float* my = red + ((size_t)threadIdx.y * C8 + ch * 8) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    my[2 * j] = s1[j];
    my[2 * j + 1] = s2[j];
  }
  __syncthreads();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < C8 * 2; i += blockDim.x * blockDim.y) {
    float acc = 0.f;
    for (int r = 0; r < (int)blockDim.y; ++r) acc += red[(size_t)r * C8 * 2 + i];
    atomicAdd(&sums[i], (double)acc);
  }
}

// dpre = df * lrelu'(f); t = dxs + dpre; sum[c] = sum dpre over (n, voxels)
__global__ void att_mid_bwd_kernel(const uint4* __restrict__ df, const uint4* __restrict__ f, const uint4* __restrict__ dxs,
                                   uint4* __restrict__ dpre, uint4* __restrict__ t, double* __restrict__ sum,
                                   int chunks, long long NV, int af) {
  extern __shared__ float red[];   // [blockDim.y][chunks*8]
  const int ch = threadIdx.x;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.y + threadIdx.y; v < NV; v += (long long)gridDim.x * blockDim.y) {
    const size_t idx = (size_t)v * chunks + ch;
    float d[8], a[8], x[8];
    unpack8(ld_stream(df + idx), d, af);
    unpack8(ld_stream(f + idx), a, af);
    unpack8(ld_stream(dxs + idx), x, af);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = a[j] > 0.f ? d[j] : LRELU * d[j];
    const uint4 o = pack8(d, af);
    dpre[idx] = o;
    unpack8(o, d, af);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] += d[j];
      x[j] += d[j];
    }
    t[idx] = pack8(x, af);
  }
  const int C8 = chunks * 8;
  float* my = red + (size_t)threadIdx.y * C8 + ch * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) my[j] = acc[j];
  __syncthreads();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < C8; i += blockDim.x * blockDim.y) {
    float a = 0.f;
    for (int r = 0; r < (int)blockDim.y; ++r) a += red[(size_t)r * C8 + i];
    atomicAdd(&sum[i], (double)a);
  }
}


// ---------------------------------------------------------------------------------------------
// Batched gather: every job is  out[i] = idx[i] < 0 ? 0 : src[idx[i]]  (src = src0 ++ src1 when src1 != NULL),
// written as 16-bit (weight packing: all conv layers' forward + data-gradient tile streams of a step in ONE launch
// instead of 92) or as fp32 times an optional device scalar (weight-gradient unpacking: all layers' dW accumulators
// back to the PyTorch parameter layout in ONE launch instead of ~50 index_selects).  One CTA = 2048 consecutive
// elements of one job; the job of a CTA is found by bisection in the prefix table `first_block`.
// ---------------------------------------------------------------------------------------------
__global__ void gather_multi_kernel(const unet3d_gather_job* __restrict__ jobs, const int* __restrict__ first_block,
                                    int n_jobs, const float* __restrict__ scale, char* out_base) {
  int lo = 0, hi = n_jobs;                    // first_block[lo] <= blockIdx.x < first_block[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((int)blockIdx.x >= __ldg(&first_block[mid])) lo = mid; else hi = mid;
  }
  const unet3d_gather_job j = jobs[lo];
  const long long base = ((long long)(blockIdx.x - __ldg(&first_block[lo])) * 256 + threadIdx.x) * 8;
  if (base >= j.n) return;
  const float* __restrict__ s0 = j.src0;
  const float* __restrict__ s1 = j.src1;
  const int n0 = j.n0;
  int id[8];
  const int cnt = (j.n - base >= 8) ? 8 : (int)(j.n - base);
  if (cnt == 8) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(j.idx + base));
    const int4 b = __ldg(reinterpret_cast<const int4*>(j.idx + base + 4));
    id[0] = a.x; id[1] = a.y; id[2] = a.z; id[3] = a.w; id[4] = b.x; id[5] = b.y; id[6] = b.z; id[7] = b.w;
  } else {
    for (int k = 0; k < 8; ++k) id[k] = k < cnt ? j.idx[base + k] : -1;
  }
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = id[k];
    f[k] = i < 0 ? 0.f : ((s1 != nullptr && i >= n0) ? __ldg(s1 + (i - n0)) : __ldg(s0 + i));
  }
  if (j.mode == 2) {
    const float sc = scale ? __ldg(scale) : 1.f;
    float* o = reinterpret_cast<float*>(out_base + reinterpret_cast<uintptr_t>(j.out)) + base;
    if (cnt == 8 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
      reinterpret_cast<float4*>(o)[0] = make_float4(f[0] * sc, f[1] * sc, f[2] * sc, f[3] * sc);
      reinterpret_cast<float4*>(o)[1] = make_float4(f[4] * sc, f[5] * sc, f[6] * sc, f[7] * sc);
    } else {
      for (int k = 0; k < cnt; ++k) o[k] = f[k] * sc;
    }
  } else {
    uint16_t* o = reinterpret_cast<uint16_t*>(out_base + reinterpret_cast<uintptr_t>(j.out)) + base;
    if (cnt == 8 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
      *reinterpret_cast<uint4*>(o) = pack8(f, j.mode);
    } else {
      for (int k = 0; k < cnt; ++k) o[k] = (uint16_t)(pack_2x16(f[k], 0.f, j.mode) & 0xffffu);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Weight-gradient accumulators dw[k3][Kp][Np] (fp32) -> PyTorch parameter layout, all layers of a step in one launch.
// The generic gather above reads one 4-byte word per 32-byte sector (the parameter layout runs tap-fastest, the
// accumulator channel-fastest).  Here one thread owns one whole sector = 8 consecutive dy channels of one (tap, row):
// two 16-byte loads, eight stores; consecutive threads are consecutive taps, then rows, i.e. consecutive destination
// addresses, so every one of the eight store instructions of a warp writes one contiguous run.
// destination element = rowmap[row] + tap + col * col_stride.
// ---------------------------------------------------------------------------------------------
__global__ void dw_unpack_kernel(const unet3d_unpack_job* __restrict__ jobs, const int* __restrict__ first_block,
                                 int n_jobs, const float* __restrict__ scale, char* out_base) {
  int lo = 0, hi = n_jobs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((int)blockIdx.x >= __ldg(&first_block[mid])) lo = mid; else hi = mid;
  }
  const unet3d_unpack_job j = jobs[lo];
  const long long item = (long long)(blockIdx.x - __ldg(&first_block[lo])) * 256 + threadIdx.x;
  const long long items = (long long)j.k3 * j.Kp * (j.Np >> 3);
  if (item >= items) return;
  const int tap = (int)(item % j.k3);
  const long long r = item / j.k3;
  const int row = (int)(r % j.Kp), cg = (int)(r / j.Kp);
  const int ro = __ldg(&j.rowmap[row]);
  if (ro < 0) return;
  const float4* src = reinterpret_cast<const float4*>(j.dw + ((size_t)tap * j.Kp + row) * j.Np + 8 * cg);
  const float4 a = __ldg(src), b = __ldg(src + 1);
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  const float sc = scale ? __ldg(scale) : 1.f;
  float* o = reinterpret_cast<float*>(out_base + reinterpret_cast<uintptr_t>(j.out)) + ro + tap;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int col = 8 * cg + k;
    if (col < j.Ncols) o[(long long)col * j.col_stride] = v[k] * sc;
  }
}

}  // namespace

// ----------------------------------- launchers ---------------------------------------------------
static inline dim3 cv_block(int chunks) {
  int ty = 256 / chunks;
  if (ty < 1) ty = 1;
  return dim3(chunks, ty, 1);
}
#define U3D_CHECK_LAUNCH() (cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA)

int weight_pack(const float* w, const int* idx, void* out, long long n, int f16, int num_sms, cudaStream_t s) {
  if (n <= 0) return U3D_ERR_INVALID;
  const int g = grid_for(n, 256 * 8, num_sms, 8);
  weight_pack_kernel<<<g, 256, 0, s>>>(w, idx, reinterpret_cast<uint16_t*>(out), n, f16);
  return U3D_CHECK_LAUNCH();
}

int in_finalize(const double* stats, const float* drop, float* table, int NC, double count, float eps, cudaStream_t s) {
  in_finalize_kernel<<<(NC + 127) / 128, 128, 0, s>>>(stats, drop, reinterpret_cast<float2*>(table), NC, 1.0 / count, eps);
  return U3D_CHECK_LAUNCH();
}

int in_apply(const bf16* y, const bf16* skip, bf16* out, const float* table, const float* shift, int N, long long V,
             int Cp, int af, int num_sms, cudaStream_t s) {
  if (Cp % 8 || Cp / 8 > 256) return U3D_ERR_INVALID;
  const int chunks = Cp / 8;
  dim3 blk = cv_block(chunks);
  const int gx = grid_rows(V, blk.y, N, num_sms, 8);
  dim3 grd(gx, N);
  if (skip)
    in_apply_kernel<true><<<grd, blk, 0, s>>>((const uint4*)y, (const uint4*)skip, (uint4*)out, (const float2*)table, shift, chunks, V, Cp, af);
  else
    in_apply_kernel<false><<<grd, blk, 0, s>>>((const uint4*)y, nullptr, (uint4*)out, (const float2*)table, shift, chunks, V, Cp, af);
  return U3D_CHECK_LAUNCH();
}

int in_bwd_reduce(const bf16* dout, const bf16* dout2, const bf16* out, const bf16* y, bf16* g, const float* table,
                  const float* shift, double* sums, int N, long long V, int Cp, int af, int num_sms, cudaStream_t s) {
  if (Cp % 8 || Cp / 8 > 256) return U3D_ERR_INVALID;
  const int chunks = Cp / 8;
  dim3 blk = cv_block(chunks);
  const int gx = grid_rows(V, blk.y, N, num_sms, 2);
  dim3 grd(gx, N);
  const size_t sm = (size_t)blk.y * Cp * 2 * sizeof(float);
  if (dout2)
    in_bwd_reduce_kernel<true><<<grd, blk, sm, s>>>((const uint4*)dout, (const uint4*)dout2, (const uint4*)out,
                                                    (const uint4*)y, (uint4*)g, (const float2*)table, shift, sums, chunks, V, Cp, af);
  else
    in_bwd_reduce_kernel<false><<<grd, blk, sm, s>>>((const uint4*)dout, nullptr, (const uint4*)out, (const uint4*)y,
                                                     (uint4*)g, (const float2*)table, shift, sums, chunks, V, Cp, af);
  return U3D_CHECK_LAUNCH();
}

int in_bwd_apply(const bf16* g, const bf16* y, bf16* dy, const float* table, const double* sums, const float* coef,
                 double* dsum, int N, int D, int H, int W, int Cp, int flags, int af, int num_sms, cudaStream_t s) {
  const int zero_last = flags & 1, g_is_dout = (flags >> 1) & 1;
  if (Cp % 8 || Cp / 8 > 256) return U3D_ERR_INVALID;
  const int chunks = Cp / 8;
  const long long V = (long long)D * H * W;
  dim3 blk = cv_block(chunks);
  const int gx = grid_rows(V, blk.y, N, num_sms, 2);
  dim3 grd(gx, N);
  const size_t sm = dsum ? (size_t)blk.y * Cp * sizeof(float) : 0;
#define U3D_BWD_APPLY(Z, S)                                                                                         \
  in_bwd_apply_kernel<Z, S><<<grd, blk, sm, s>>>((const uint4*)g, (const uint4*)y, (uint4*)dy, (const float2*)table, \
                                                 sums, coef, dsum, chunks, V, Cp, 1.0 / (double)V, zero_last, D, H, W, af, g_is_dout)
  if (g_is_dout && coef != nullptr) return U3D_ERR_INVALID;        // recomputation is for the InstanceNorm path only
  if (zero_last && dsum) U3D_BWD_APPLY(true, true);
  else if (zero_last) U3D_BWD_APPLY(true, false);
  else if (dsum) U3D_BWD_APPLY(false, true);
  else U3D_BWD_APPLY(false, false);
#undef U3D_BWD_APPLY
  return U3D_CHECK_LAUNCH();
}

int in_bwd_small(const bf16* dout, const bf16* dout2, const bf16* out, const bf16* y, bf16* g, bf16* dy, const float* table,
                 double* sums, int N, long long V, int Cp, int af, int num_sms, cudaStream_t s) {
  if (Cp % 8 || V < 1 || V > (1 << 24)) return U3D_ERR_INVALID;
  if (g == nullptr && (out != nullptr || dout2 != nullptr)) return U3D_ERR_INVALID;
  const int chunks = Cp / 8, T = 256;
  // cluster size: enough CTAs to cover the SMs, at most 8 (the portable limit), at least one voxel row per CTA
  int kc = 1;
  while (kc < 8 && (long long)chunks * N * kc < num_sms && (long long)2 * kc * T <= V) kc *= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(chunks * kc, N);
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = (T / 32) * 16 * sizeof(float);
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kc;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const uint4 *pd = (const uint4*)dout, *pd2 = (const uint4*)dout2, *po = (const uint4*)out, *py = (const uint4*)y;
  uint4 *pg = (uint4*)g, *pdy = (uint4*)dy;
  const float2* pt = (const float2*)table;
  const int Vi = (int)V;
  const double inv = 1.0 / (double)V;
  cudaError_t e = dout2 ? cudaLaunchKernelEx(&cfg, in_bwd_small_kernel<true>, pd, pd2, po, py, pg, pdy, pt, sums, chunks, Vi, Cp, inv, af)
                        : cudaLaunchKernelEx(&cfg, in_bwd_small_kernel<false>, pd, pd2, po, py, pg, pdy, pt, sums, chunks, Vi, Cp, inv, af);
  return e == cudaSuccess ? U3D_CHECK_LAUNCH() : U3D_ERR_CUDA;
}

int channel_sum(const bf16* x, double* dsum, long long NV, int Cp, int num_sms, cudaStream_t s) {
  if (Cp % 8 || Cp / 8 > 256) return U3D_ERR_INVALID;
  const int chunks = Cp / 8;
  dim3 blk = cv_block(chunks);
  const int gx = grid_for(NV, blk.y * 8, num_sms, 8);
  channel_sum_kernel<<<gx, blk, (size_t)blk.y * Cp * sizeof(float), s>>>((const uint4*)x, dsum, chunks, NV);
  return U3D_CHECK_LAUNCH();
}

int stem_fwd(const float* x, const float* w, const float* b, bf16* out, int N, int Cin, int D, int H, int W, int Cp,
             int af, int num_sms, cudaStream_t s) {
  const long long total = (long long)N * D * H * W;
  const int g = grid_for(total, 128, num_sms, 16);
  const size_t smem = ((size_t)Cin * 27 * Cp + Cp) * sizeof(float);
  if (Cin < 1 || smem > 48 * 1024) return U3D_ERR_UNSUPPORTED;
  // Single-channel stems run on the mma.sync kernel (0.21 vs 0.32 ms at 2 x 128^3).  Round 1 kept it opt-in because a
  // 2-GPU check showed 4e-3 of gradient noise with it; tools/diag/stem_partition.py (round 2) shows its output is
  // bit-identical for N = 4 and for the same samples as 2 + 2 and from run to run, i.e. it does not depend on the
  // partition; U3D_STEM_FWD_FMA=1 selects the CUDA-core kernel for comparisons.
  static const bool use_mma = getenv("U3D_STEM_FWD_FMA") == nullptr;
  if (Cin == 1 && use_mma && (Cp == 16 || Cp == 32) && total < 0x7fffffffLL) {
    const int gm = grid_for((long long)N * D * H, 8 * 4, num_sms, 4);
    if (Cp == 32) {
      if (af) stem_fwd_mma_kernel<32, true><<<gm, 256, 0, s>>>(x, w, b, out, N, D, H, W);
      else stem_fwd_mma_kernel<32, false><<<gm, 256, 0, s>>>(x, w, b, out, N, D, H, W);
    } else {
      if (af) stem_fwd_mma_kernel<16, true><<<gm, 256, 0, s>>>(x, w, b, out, N, D, H, W);
      else stem_fwd_mma_kernel<16, false><<<gm, 256, 0, s>>>(x, w, b, out, N, D, H, W);
    }
    return U3D_CHECK_LAUNCH();
  }
#define U3D_SF(CPV)                                                                                  \
  do {                                                                                               \
    if (Cin == 1) stem_fwd_kernel<CPV, true><<<g, 128, smem, s>>>(x, w, b, out, N, Cin, D, H, W, af); \
    else stem_fwd_kernel<CPV, false><<<g, 128, smem, s>>>(x, w, b, out, N, Cin, D, H, W, af);        \
  } while (0)
  if (Cp == 32) U3D_SF(32);
  else if (Cp == 16) U3D_SF(16);
  else if (Cp == 48) U3D_SF(48);
  else if (Cp == 64) U3D_SF(64);
  else return U3D_ERR_UNSUPPORTED;
#undef U3D_SF
  return U3D_CHECK_LAUNCH();
}

int stem_wgrad(const float* x, const bf16* dy, float* dw, int N, int D, int H, int W, long long x_sN, int Cp, int af,
               int num_sms, cudaStream_t s) {
  const long long n_runs = (long long)N * D * H * ((W + 15) / 16);
  static const bool legacy = getenv("U3D_STEM_WGRAD_FMA") != nullptr;        // the CUDA-core kernel, for comparisons
  if (!legacy && n_runs < 0x7fffffffLL) {
    const int g = grid_for((long long)N * D * H, 8 * 4, num_sms, 4);
#define U3D_SW(CPV)                                                                                          \
  do {                                                                                                       \
    constexpr int smem = 8 * 2 * 16 * (CPV * 2 + 16) + 8 * 2 * 10 * STEM_XP * 4 + 32 * CPV * 4;                   \
    static bool attr[64] = {};                                                                               \
    if (first_use_on_device(attr)) {                                                                         \
      cudaFuncSetAttribute(stem_wgrad_mma_kernel<CPV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  \
      cudaFuncSetAttribute(stem_wgrad_mma_kernel<CPV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    }                                                                                                        \
    if (af) stem_wgrad_mma_kernel<CPV, true><<<g, 256, smem, s>>>(x, dy, dw, N, D, H, W, x_sN);              \
    else stem_wgrad_mma_kernel<CPV, false><<<g, 256, smem, s>>>(x, dy, dw, N, D, H, W, x_sN);                \
  } while (0)
    if (Cp == 32) U3D_SW(32);
    else if (Cp == 16) U3D_SW(16);
    else if (Cp == 48) U3D_SW(48);
    else if (Cp == 64) U3D_SW(64);
    else return U3D_ERR_UNSUPPORTED;
#undef U3D_SW
    return U3D_CHECK_LAUNCH();
  }
  dim3 blk(Cp / 8, 3, 16);
  const int g = grid_for(n_runs, 16 * 4, num_sms, 4);
  if (Cp == 32) stem_wgrad_kernel<32><<<g, blk, 0, s>>>(x, dy, dw, N, D, H, W, x_sN, af);
  else if (Cp == 16) stem_wgrad_kernel<16><<<g, blk, 0, s>>>(x, dy, dw, N, D, H, W, x_sN, af);
  else if (Cp == 48) stem_wgrad_kernel<48><<<g, blk, 0, s>>>(x, dy, dw, N, D, H, W, x_sN, af);
  else if (Cp == 64) stem_wgrad_kernel<64><<<g, blk, 0, s>>>(x, dy, dw, N, D, H, W, x_sN, af);
  else return U3D_ERR_UNSUPPORTED;
  return U3D_CHECK_LAUNCH();
}

int head_fwd(const bf16* a, const float* w, const float* b, float* logits, int K, int N, long long V, int Cp, int af,
             int num_sms, cudaStream_t s) {
  if (K < 1 || K > 8) return U3D_ERR_UNSUPPORTED;
  const int g = grid_for((long long)N * V, 256, num_sms, 16);
  if (Cp == 32) head_fwd_kernel<32, 8><<<g, 256, 0, s>>>(a, w, b, logits, K, N, V, af);
  else if (Cp == 16) head_fwd_kernel<16, 8><<<g, 256, 0, s>>>(a, w, b, logits, K, N, V, af);
  else if (Cp == 48) head_fwd_kernel<48, 8><<<g, 256, 0, s>>>(a, w, b, logits, K, N, V, af);
  else if (Cp == 64) head_fwd_kernel<64, 8><<<g, 256, 0, s>>>(a, w, b, logits, K, N, V, af);
  else return U3D_ERR_UNSUPPORTED;
  return U3D_CHECK_LAUNCH();
}

int head_bwd(const float* dl, const bf16* a, const float* w, bf16* da, float* dw, const float* gscale, int K, int N,
             long long V, int Cp, int af, int num_sms, cudaStream_t s) {
  if (K < 1 || K > 4) return U3D_ERR_UNSUPPORTED;
  const int g = grid_for((long long)N * V, (256 / (Cp / 8)) * 8, num_sms, 8);
  if (Cp == 32) head_bwd_kernel<32, 4><<<g, 256, 0, s>>>(dl, a, w, da, dw, gscale, K, N, V, af);
  else if (Cp == 16) head_bwd_kernel<16, 4><<<g, 256, 0, s>>>(dl, a, w, da, dw, gscale, K, N, V, af);
  else return U3D_ERR_UNSUPPORTED;
  return U3D_CHECK_LAUNCH();
}

int loss_fwd(const float* logits, const long long* target, double* sums, int K, int N, long long V, float gamma,
             int num_sms, cudaStream_t s) {
  if (K < 2 || K > 8) return U3D_ERR_UNSUPPORTED;
  const int g = grid_for((long long)N * V, 256 * 4, num_sms, 8);
  loss_fwd_kernel<8><<<g, 256, 0, s>>>(logits, target, sums, K, N, V, gamma);
  return U3D_CHECK_LAUNCH();
}

int loss_bwd(const float* logits, const long long* target, const float* coef, const float* gscale, float* dlogits, int K,
             int N, long long V, float gamma, int use_focal, int num_sms, cudaStream_t s) {
  if (K < 2 || K > 8) return U3D_ERR_UNSUPPORTED;
  const int g = grid_for((long long)N * V, 256 * 4, num_sms, 8);
  loss_bwd_kernel<8><<<g, 256, 0, s>>>(logits, target, coef, gscale, dlogits, K, N, V, gamma, use_focal);
  return U3D_CHECK_LAUNCH();
}

int sw_accumulate(const float* logits, const float* window, long long* acc, int K, int px, int py, int pz,
                  int x0, int y0, int z0, int X, int Y, int Z, int Xs, int num_sms, cudaStream_t s) {
  if (K < 1 || K > 8) return U3D_ERR_UNSUPPORTED;
  if (x0 < 0 || y0 < 0 || z0 < 0 || x0 + px > X || y0 + py > Y || z0 + pz > Z || Xs < 1) return U3D_ERR_INVALID;
  const int g = grid_for((long long)px * py * pz, 256 * 2, num_sms, 8);
  sw_accumulate_kernel<8><<<g, 256, 0, s>>>(logits, window, acc, K, px, py, pz, x0, y0, z0, Xs, Y, Z);
  return U3D_CHECK_LAUNCH();
}

int sw_finalize(const long long* acc, uint8_t* labels, float* probs, int K, long long n, int num_sms, cudaStream_t s) {
  if (K < 1 || K > 8) return U3D_ERR_UNSUPPORTED;
  const int g = grid_for(n, 256 * 2, num_sms, 8);
  sw_finalize_kernel<8><<<g, 256, 0, s>>>(acc, labels, probs, K, n);
  return U3D_CHECK_LAUNCH();
}

int maxpool_fwd(const bf16* x, bf16* out, uint8_t* code, int N, int D, int H, int W, int Cp, int af, int num_sms,
                cudaStream_t s) {
  if (Cp % 8 || Cp / 8 > 256 || (D | H | W) & 1 || N < 1) return U3D_ERR_INVALID;
  const int chunks = Cp / 8;
  const long long total = (long long)N * (D / 2) * (H / 2) * (W / 2);
  dim3 blk = cv_block(chunks);
  const int g = grid_for(total, blk.y * 2, num_sms, 8);
  maxpool_fwd_kernel<<<g, blk, 0, s>>>((const uint4*)x, (uint4*)out, (uint2*)code, chunks, D / 2, H / 2, W / 2, total, af);
  return U3D_CHECK_LAUNCH();
}

int maxpool_bwd(const bf16* dout, const uint8_t* code, bf16* dx, int N, int D, int H, int W, int Cp, int num_sms,
                cudaStream_t s) {
  if (Cp % 8 || Cp / 8 > 256 || (D | H | W) & 1 || N < 1) return U3D_ERR_INVALID;
  const int chunks = Cp / 8;
  const long long total = (long long)N * (D / 2) * (H / 2) * (W / 2);
  dim3 blk = cv_block(chunks);
  const int g = grid_for(total, blk.y * 2, num_sms, 8);
  maxpool_bwd_kernel<<<g, blk, 0, s>>>((const uint4*)dout, (const uint2*)code, (uint4*)dx, chunks, D / 2, H / 2, W / 2, total);
  return U3D_CHECK_LAUNCH();
}

int att_gate_fwd(const bf16* xs, const bf16* z, bf16* out, long long n_elem, int af, int num_sms, cudaStream_t s) {
  if (n_elem <= 0 || n_elem % 8) return U3D_ERR_INVALID;
  const long long n16 = n_elem / 8;
  const int g = grid_for(n16, 256 * 2, num_sms, 16);
  att_gate_fwd_kernel<<<g, 256, 0, s>>>((const uint4*)xs, (const uint4*)z, (uint4*)out, n16, af);
  return U3D_CHECK_LAUNCH();
}

int att_gate_bwd(const bf16* dout, const bf16* xs, const bf16* z, bf16* dxs, bf16* dz, double* sums, long long NV, int Cp,
                 int af, int num_sms, cudaStream_t s) {
  if (Cp % 8 || Cp / 8 > 256 || NV <= 0) return U3D_ERR_INVALID;
  const int chunks = Cp / 8;
  dim3 blk = cv_block(chunks);
  const int g = grid_for(NV, blk.y * 8, num_sms, 8);
  att_gate_bwd_kernel<<<g, blk, (size_t)blk.y * Cp * 2 * sizeof(float), s>>>(
      (const uint4*)dout, (const uint4*)xs, (const uint4*)z, (uint4*)dxs, (uint4*)dz, sums, chunks, NV, af);
  return U3D_CHECK_LAUNCH();
}

int att_mid_bwd(const bf16* df, const bf16* f, const bf16* dxs, bf16* dpre, bf16* t, double* sum, long long NV, int Cp,
                int af, int num_sms, cudaStream_t s) {
  if (Cp % 8 || Cp / 8 > 256 || NV <= 0) return U3D_ERR_INVALID;
  const int chunks = Cp / 8;
  dim3 blk = cv_block(chunks);
  const int g = grid_for(NV, blk.y * 8, num_sms, 8);
  att_mid_bwd_kernel<<<g, blk, (size_t)blk.y * Cp * sizeof(float), s>>>(
      (const uint4*)df, (const uint4*)f, (const uint4*)dxs, (uint4*)dpre, (uint4*)t, sum, chunks, NV, af);
  return U3D_CHECK_LAUNCH();
}

int gather_multi(const unet3d_gather_job* jobs, const int* first_block, int n_jobs, int n_blocks, const float* scale,
                 void* out_base, cudaStream_t s) {
  if (n_jobs < 1 || n_blocks < 1) return U3D_ERR_INVALID;
  gather_multi_kernel<<<n_blocks, 256, 0, s>>>(jobs, first_block, n_jobs, scale, reinterpret_cast<char*>(out_base));
  return U3D_CHECK_LAUNCH();
}

int dw_unpack(const unet3d_unpack_job* jobs, const int* first_block, int n_jobs, int n_blocks, const float* scale,
              void* out_base, cudaStream_t s) {
  if (n_jobs < 1 || n_blocks < 1) return U3D_ERR_INVALID;
  dw_unpack_kernel<<<n_blocks, 256, 0, s>>>(jobs, first_block, n_jobs, scale, reinterpret_cast<char*>(out_base));
  return U3D_CHECK_LAUNCH();
}

}  // namespace u3d
