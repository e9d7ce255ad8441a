// Train-loader augmentation on the device (reference transform.py:176-301): mirror, contrast,
// brightness, gamma on a float32 (X, Y, Z, C) patch / uint8 (X, Y, Z) label that already sit in HBM (the random
// rescale-crop in front of them is csrc/resample.cu's zoom).  All are HBM-bound single passes over a ~8 MB patch; what
// makes them worth a file is BIT-EXACTNESS with the reference's numpy arithmetic under a fixed numpy seed:
//   * adjust_contrast needs `input.mean()`: numpy's float32 pairwise summation (blocks of <= 128 elements summed with
//     eight strided accumulators, combined over an uneven binary tree) is reproduced operation for operation --
//     aug_leafsum_kernel evaluates the leaves, aug_tree_kernel walks the recursion;
//   * every elementwise formula uses separately rounded fp32 operations in numpy's order (no FMA contraction);
//   * min / max are order independent; the statistics never leave the device (no host synchronisation).
// np.power (adjust_gamma) is a SIMD routine on the host (SVML / AVX-512 for float32 on x86: not correctly rounded, and
// different from CPU to CPU), so it has no bit pattern to reproduce; here it is the double-precision pow rounded to
// float, within 1 float32 ulp of numpy's value on the reference's golden vectors (tests/test_augment_gpu.py).
#include "kernels.cuh"

namespace u3d {

namespace {

__device__ __forceinline__ long long flip_index(long long i, int X, int Y, int Z, int C, int fx, int fy, int fz) {
  const int c = (int)(i % C);
  long long v = i / C;
  int z = (int)(v % Z); v /= Z;
  int y = (int)(v % Y);
  int x = (int)(v / Y);
  if (fx) x = X - 1 - x;
  if (fy) y = Y - 1 - y;
  if (fz) z = Z - 1 - z;
  return (((long long)x * Y + y) * Z + z) * C + c;
}

template <typename T>
__global__ void aug_flip_kernel(const T* __restrict__ in, T* __restrict__ out, long long n, int X, int Y, int Z, int C,
                                int fx, int fy, int fz) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = in[flip_index(i, X, Y, Z, C, fx, fy, fz)];
}

// order-preserving float <-> int encoding for atomicMin / atomicMax
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// stats_i: [0] = min (ordered int), [1] = max (ordered int); initialised by aug_stats_init_kernel
__global__ void aug_stats_init_kernel(int* stats_i) {
  stats_i[0] = 0x7fffffff;
  stats_i[1] = (int)0x80000000;
}

__global__ void aug_minmax_kernel(const float* __restrict__ x, long long n, int* __restrict__ stats_i) {
  float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&stats_i[0], f2ord(lo));
    atomicMax(&stats_i[1], f2ord(hi));
  }
}

// numpy's pairwise-sum leaf (n <= 128), one thread per leaf:
//   n < 8: res = 0; res += a[i] in order
//   else : r[0..7] = a[0..7]; r[j] += a[i + j] for i = 8, 16, ... < n - n % 8;
//          res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)); then the n % 8 tail in order
__global__ void aug_leafsum_kernel(const float* __restrict__ x, const long long* __restrict__ leaf_off, int n_leaves,
                                   float* __restrict__ leaf_sum) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_leaves) return;
  const float* a = x + leaf_off[l];
  const int n = (int)(leaf_off[l + 1] - leaf_off[l]);
  float res;
  if (n < 8) {
    res = 0.f;
    for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
  } else {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    const int m = n - (n % 8);
    for (int i = 8; i < m; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    }
    res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                    __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (int i = m; i < n; ++i) res = __fadd_rn(res, a[i]);
  }
  leaf_sum[l] = res;
}

// numpy's recursion above the leaves: sum(n) = sum(n2) + sum(n - n2), n2 = (n / 2) rounded down to a multiple of 8.
// One thread, explicit stack; consumes the leaf sums in order.  stats_f[2] = mean = sum / (float)n, stats_f[3] = sum.
__global__ void aug_tree_kernel(const float* __restrict__ leaf_sum, long long n_total, float* __restrict__ stats_f) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  long long fn[48];
  float fl[48];
  int fs[48];
  int sp = 0, k = 0;
  fn[0] = n_total; fs[0] = 0; fl[0] = 0.f;
  float ret = 0.f;
  while (sp >= 0) {
    const long long n = fn[sp];
    if (n <= 128) {
      ret = leaf_sum[k++];
      --sp;
      continue;
    }
    long long n2 = n / 2;
    n2 -= n2 % 8;
    if (fs[sp] == 0) {
      fs[sp] = 1;
      ++sp; fn[sp] = n2; fs[sp] = 0;
    } else if (fs[sp] == 1) {
      fl[sp] = ret;
      fs[sp] = 2;
      ++sp; fn[sp] = n - n2; fs[sp] = 0;
    } else {
      ret = __fadd_rn(fl[sp], ret);
      --sp;
    }
  }
  stats_f[3] = ret;
  stats_f[2] = __fdiv_rn(ret, (float)n_total);
}

// adjust_contrast / adjust_brightness (transform.py:176-185): out = (x - a) * factor + a, a = mean or min
__global__ void aug_affine_kernel(const float* __restrict__ x, float* __restrict__ out, long long n,
                                  const float* __restrict__ stats_f, const int* __restrict__ stats_i, int which, float factor) {
  const float a = which == 0 ? stats_f[2] : ord2f(stats_i[0]);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(__fsub_rn(x[i], a), factor), a);
}

// adjust_gamma (transform.py:188-193): arange = max - min + eps; out = power((x - min) / arange, gamma) * arange + min
__global__ void aug_gamma_kernel(const float* __restrict__ x, float* __restrict__ out, long long n,
                                 const int* __restrict__ stats_i, float gamma, float eps) {
  const float lo = ord2f(stats_i[0]), hi = ord2f(stats_i[1]);
  const float arange = __fadd_rn(__fsub_rn(hi, lo), eps);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float t = __fdiv_rn(__fsub_rn(x[i], lo), arange);
    const float pw = (float)pow((double)t, (double)gamma);
    out[i] = __fadd_rn(__fmul_rn(pw, arange), lo);
  }
}

inline int aug_grid(long long n, int num_sms) {
  long long need = (n + 255) / 256;
  long long cap = (long long)num_sms * 8;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace

#define U3D_CHECK_LAUNCH() (cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA)

int aug_flip(const void* in, void* out, int elem_bytes, int X, int Y, int Z, int C, int fx, int fy, int fz, int num_sms,
             cudaStream_t s) {
  if (X < 1 || Y < 1 || Z < 1 || C < 1 || in == out) return U3D_ERR_INVALID;
  const long long n = (long long)X * Y * Z * C;
  const int g = aug_grid(n, num_sms);
  if (elem_bytes == 4)
    aug_flip_kernel<float><<<g, 256, 0, s>>>((const float*)in, (float*)out, n, X, Y, Z, C, fx, fy, fz);
  else if (elem_bytes == 1)
    aug_flip_kernel<uint8_t><<<g, 256, 0, s>>>((const uint8_t*)in, (uint8_t*)out, n, X, Y, Z, C, fx, fy, fz);
  else
    return U3D_ERR_UNSUPPORTED;
  return U3D_CHECK_LAUNCH();
}

int aug_stats(const float* x, long long n, const long long* leaf_off, int n_leaves, float* leaf_scratch, float* stats,
              int num_sms, cudaStream_t s) {
  if (n < 1) return U3D_ERR_INVALID;
  int* stats_i = reinterpret_cast<int*>(stats);
  aug_stats_init_kernel<<<1, 1, 0, s>>>(stats_i);
  aug_minmax_kernel<<<aug_grid(n, num_sms), 256, 0, s>>>(x, n, stats_i);
  if (leaf_off != nullptr) {
    if (n_leaves < 1 || leaf_scratch == nullptr) return U3D_ERR_INVALID;
    aug_leafsum_kernel<<<(n_leaves + 127) / 128, 128, 0, s>>>(x, leaf_off, n_leaves, leaf_scratch);
    aug_tree_kernel<<<1, 32, 0, s>>>(leaf_scratch, n, stats);
  }
  return U3D_CHECK_LAUNCH();
}

int aug_affine(const float* x, float* out, long long n, const float* stats, int which, float factor, int num_sms,
               cudaStream_t s) {
  if (n < 1 || (which != 0 && which != 1)) return U3D_ERR_INVALID;
  aug_affine_kernel<<<aug_grid(n, num_sms), 256, 0, s>>>(x, out, n, stats, reinterpret_cast<const int*>(stats), which, factor);
  return U3D_CHECK_LAUNCH();
}

int aug_gamma(const float* x, float* out, long long n, const float* stats, float gamma, float eps, int num_sms,
              cudaStream_t s) {
  if (n < 1) return U3D_ERR_INVALID;
  aug_gamma_kernel<<<aug_grid(n, num_sms), 256, 0, s>>>(x, out, n, reinterpret_cast<const int*>(stats), gamma, eps);
  return U3D_CHECK_LAUNCH();
}

}  // namespace u3d
