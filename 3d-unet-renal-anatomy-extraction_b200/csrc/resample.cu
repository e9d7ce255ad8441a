// Case-level resampling either side of the window loop -- HBM-bound kernels.
//
//   transform.rescale / resize (transform.py:32-100) = scipy.ndimage.zoom(order=1, mode='reflect'), per channel;
//   labels with >= 3 classes: one float one-hot volume per class, zoomed, arg-maxed (first maximum wins);
//   data.resample_normalize_case (data.py:258-275): clip to the percentiles, z-score -- fused into the image zoom.
//
// Bit-exact with SciPy (NI_ZoomShift, order 1): the per-axis coordinate cc = o * (n_in-1)/(n_out-1), its floor and the
// weights (1-x, x) are float64; a voxel is the float64 sum over the 8 corners -- last axis fastest -- of
// ((v * w0) * w1) * w2, rounded once to float32.  Every float64 operation below is an explicit round-to-nearest
// intrinsic so that nothing is contracted into an FMA.
#include "kernels.cuh"

namespace u3d {
namespace {

// per-axis tables for one launch, in the caller's workspace: for every output index the two source indices and the
// two float64 weights (the second index of the last sample is reflected back onto the volume: d c b a | a b c d).
__global__ void zoom_tables_kernel(int* __restrict__ idx, double* __restrict__ wts, int n_in0, int n_in1, int n_in2,
                                   int n_out0, int n_out1, int n_out2, double st0, double st1, double st2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int o = i, n_in, base = 0;
  double st;
  if (o < n_out0) { n_in = n_in0; st = st0; }
  else if ((o -= n_out0) < n_out1) { n_in = n_in1; st = st1; base = n_out0; }
  else if ((o -= n_out1) < n_out2) { n_in = n_in2; st = st2; base = n_out0 + n_out1; }
  else return;
  const double cc = __dmul_rn((double)o, st);
  const double fl = floor(cc);
  const double x = __dsub_rn(cc, fl);
  int i0 = (int)fl, i1 = i0 + 1;
  if (i1 >= n_in) i1 = 2 * n_in - i1 - 1;
  i0 = min(max(i0, 0), n_in - 1);
  i1 = min(max(i1, 0), n_in - 1);
  idx[2 * (base + o)] = i0;
  idx[2 * (base + o) + 1] = i1;
  wts[2 * (base + o)] = __dsub_rn(1.0, x);
  wts[2 * (base + o) + 1] = x;
}

struct ZoomNorm {
  int on;
  float lo[4], hi[4], mean[4], den[4];
};

template <typename T> __device__ __forceinline__ double load_as_double(const T* p) { return (double)__ldg(p); }

// Keep a loop-invariant kernel parameter in a register: without this the compiler re-reads the parameters from the constant
// bank inside the voxel loop, and those loads (ADU pipe) -- not float64 arithmetic or HBM -- bounded the kernels (ncu: 91 %).
__device__ __forceinline__ long long pin(long long v) { asm volatile("" : "+l"(v)); return v; }
__device__ __forceinline__ int pin(int v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ float pin(float v) { asm volatile("" : "+f"(v)); return v; }
template <typename T> __device__ __forceinline__ T* pin(T* p) { asm volatile("" : "+l"(p)); return p; }

// one CTA = one or more (x, y) output rows, threads run along z (coalesced stores, contiguous corner loads); the x / y
// table entries are uniform per row, so a voxel costs its 8 loads, 24 + 8 float64 operations and a handful of integer ones
template <typename TIn, typename TOut>
__global__ void zoom_linear_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, const int* __restrict__ idx,
                                   const double* __restrict__ wts, int C, int OX, int OY, int OZ, long long isx,
                                   long long isy, long long isz, long long isc, long long osx, long long osy,
                                   long long osz, long long osc, ZoomNorm nm) {
  const int rows = OX * OY;
  in = pin(in); out = pin(out); idx = pin(idx); wts = pin(wts);
  C = pin(C); OZ = pin(OZ); isz = pin(isz); isc = pin(isc); osz = pin(osz); osc = pin(osc);
  const int on = pin(nm.on);
  // channel 0's constants live in registers; further channels (rare) read the parameter block
  const float lo0 = pin(nm.lo[0]), hi0 = pin(nm.hi[0]), mean0 = pin(nm.mean[0]), den0 = pin(nm.den[0]);
  const int2* idz = reinterpret_cast<const int2*>(idx) + OX + OY;
  const double2* wtz = reinterpret_cast<const double2*>(wts) + OX + OY;
  for (int row = blockIdx.x * blockDim.y + threadIdx.y; row < rows; row += gridDim.x * blockDim.y) {
    const int ox = row / OY, oy = row - ox * OY;
    const int2 ix = __ldg(reinterpret_cast<const int2*>(idx) + ox);
    const int2 iy = __ldg(reinterpret_cast<const int2*>(idx) + OX + oy);
    const double2 wx = __ldg(reinterpret_cast<const double2*>(wts) + ox);
    const double2 wy = __ldg(reinterpret_cast<const double2*>(wts) + OX + oy);
    const long long bxy[4] = {ix.x * isx + iy.x * isy, ix.x * isx + iy.y * isy, ix.y * isx + iy.x * isy,
                              ix.y * isx + iy.y * isy};
    const double ax[2] = {wx.x, wx.y}, ay[2] = {wy.x, wy.y};
    TOut* orow = out + ox * osx + oy * osy;
    for (int oz = threadIdx.x; oz < OZ; oz += blockDim.x) {
      const int2 iz = __ldg(idz + oz);
      const double2 wz = __ldg(wtz + oz);
      const long long bz[2] = {iz.x * isz, iz.y * isz};
      const double az[2] = {wz.x, wz.y};
      for (int c = 0; c < C; ++c) {
        const TIn* src = in + c * isc;
        double t = 0.0;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int d = 0; d < 2; ++d) {
              const double v = load_as_double(src + bxy[a * 2 + b] + bz[d]);
              t = __dadd_rn(t, __dmul_rn(__dmul_rn(__dmul_rn(v, ax[a]), ay[b]), az[d]));
            }
        float r = __double2float_rn(t);
        if (on) {
          float lo = lo0, hi = hi0, mean = mean0, den = den0;
          if (c > 0) {
            const int cn = c < 4 ? c : 3;
            lo = nm.lo[cn], hi = nm.hi[cn], mean = nm.mean[cn], den = nm.den[cn];
          }
          r = fminf(fmaxf(r, lo), hi);                                     // np.clip
          r = __fdiv_rn(__fsub_rn(r, mean), den);                          // (x - mean) / (std + 1e-8), float32
        }
        TOut* dst = orow + oz * osz + c * osc;
        if constexpr (sizeof(TOut) == 1) *dst = (TOut)r;                   // .astype(uint8): truncation
        else *dst = r;
      }
    }
  }
}

// labels with >= 3 classes: argmax_c float32( sum over corners with label c, in corner order, of (w0*w1)*w2 ).
// Classes absent from the 8 corners score 0 and can never win (the weights sum to 1), so only the corner labels compete;
// ties go to the smaller class index (np.argmax).
__global__ void zoom_label_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int* __restrict__ idx,
                                  const double* __restrict__ wts, int OX, int OY, int OZ, long long isx, long long isy,
                                  long long isz, long long osx, long long osy, long long osz) {
  const int rows = OX * OY;
  in = pin(in); out = pin(out); idx = pin(idx); wts = pin(wts);
  OZ = pin(OZ); isz = pin(isz); osz = pin(osz);
  const int2* idz = reinterpret_cast<const int2*>(idx) + OX + OY;
  const double2* wtz = reinterpret_cast<const double2*>(wts) + OX + OY;
  for (int row = blockIdx.x * blockDim.y + threadIdx.y; row < rows; row += gridDim.x * blockDim.y) {
    const int ox = row / OY, oy = row - ox * OY;
    const int2 ix = __ldg(reinterpret_cast<const int2*>(idx) + ox);
    const int2 iy = __ldg(reinterpret_cast<const int2*>(idx) + OX + oy);
    const long long bxy[4] = {ix.x * isx + iy.x * isy, ix.x * isx + iy.y * isy, ix.y * isx + iy.x * isy,
                              ix.y * isx + iy.y * isy};
    uint8_t* orow = out + ox * osx + oy * osy;
    for (int oz = threadIdx.x; oz < OZ; oz += blockDim.x) {
      const int2 iz = __ldg(idz + oz);
      const long long bz[2] = {iz.x * isz, iz.y * isz};
      int lab[8];
      bool same = true;
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        lab[h] = __ldg(in + bxy[h >> 1] + bz[h & 1]);
        same = same && lab[h] == lab[0];
      }
      int best = lab[0];
      if (!same) {
        const double2 wx = __ldg(reinterpret_cast<const double2*>(wts) + ox);
        const double2 wy = __ldg(reinterpret_cast<const double2*>(wts) + OX + oy);
        const double2 wz = __ldg(wtz + oz);
        const double ax[2] = {wx.x, wx.y}, ay[2] = {wy.x, wy.y}, az[2] = {wz.x, wz.y};
        double p[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) p[h] = __dmul_rn(__dmul_rn(ax[h >> 2], ay[(h >> 1) & 1]), az[h & 1]);
        float bestv = -1.f;
        best = 256;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const int c = lab[h];
          double t = 0.0;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (lab[j] == c) t = __dadd_rn(t, p[j]);
          const float tv = __double2float_rn(t);
          if (tv > bestv || (tv == bestv && c < best)) { bestv = tv; best = c; }
        }
      }
      orow[oz * osz] = (uint8_t)best;
    }
  }
}

inline double axis_step(int n_in, int n_out) { return n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 1.0; }

int build_tables(const int* ishape, const int* oshape, void* ws, size_t ws_bytes, int** idx, double** wts, cudaStream_t s) {
  for (int d = 0; d < 3; ++d)
    if (ishape[d] < 1 || oshape[d] < 1) return U3D_ERR_INVALID;
  const long long n = (long long)oshape[0] + oshape[1] + oshape[2];
  if (!ws || ws_bytes < zoom_workspace_bytes(oshape[0], oshape[1], oshape[2]) || ((uintptr_t)ws & 15)) return U3D_ERR_INVALID;
  *wts = reinterpret_cast<double*>(ws);
  *idx = reinterpret_cast<int*>(*wts + 2 * n);
  zoom_tables_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(*idx, *wts, ishape[0], ishape[1], ishape[2], oshape[0],
                                                                 oshape[1], oshape[2], axis_step(ishape[0], oshape[0]),
                                                                 axis_step(ishape[1], oshape[1]),
                                                                 axis_step(ishape[2], oshape[2]));
  return cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA;
}

// threads along z: 32..128; short rows share a CTA (blockDim.y rows) so that a CTA always has 128 threads
inline void row_geometry(const int* oshape, int num_sms, dim3* grid, dim3* block) {
  const int tz = oshape[2] > 64 ? 128 : (oshape[2] > 32 ? 64 : 32);
  const int ty = 128 / tz;
  const long long rows = (long long)oshape[0] * oshape[1];
  const long long want = (rows + ty - 1) / ty, cap = (long long)num_sms * 64;
  *block = dim3(tz, ty, 1);
  *grid = dim3((unsigned)(want < 1 ? 1 : (want < cap ? want : cap)), 1, 1);
}

}  // namespace

size_t zoom_workspace_bytes(int ox, int oy, int oz) {
  const size_t n = (size_t)ox + oy + oz;
  return n * 2 * sizeof(double) + n * 2 * sizeof(int);
}

int zoom_linear(const void* in, int in_u8, void* out, int out_u8, int C, const int* ishape, const long long* istride,
                const int* oshape, const long long* ostride, const float* norm_host, void* ws, size_t ws_bytes,
                int num_sms, cudaStream_t s) {
  if (C < 1 || (norm_host && (C > 4 || in_u8 || out_u8))) return U3D_ERR_UNSUPPORTED;
  int* idx;
  double* wts;
  if (int rc = build_tables(ishape, oshape, ws, ws_bytes, &idx, &wts, s)) return rc;
  ZoomNorm nm{};
  if (norm_host) {
    nm.on = 1;
    for (int c = 0; c < C; ++c) {
      nm.lo[c] = norm_host[4 * c], nm.hi[c] = norm_host[4 * c + 1];
      nm.mean[c] = norm_host[4 * c + 2], nm.den[c] = norm_host[4 * c + 3];
    }
  }
  if ((long long)oshape[0] * oshape[1] > 0x7fffffffLL) return U3D_ERR_UNSUPPORTED;
  dim3 g, b;
  row_geometry(oshape, num_sms, &g, &b);
#define U3D_ZOOM(TI, TO)                                                                                              \
  zoom_linear_kernel<TI, TO><<<g, b, 0, s>>>((const TI*)in, (TO*)out, idx, wts, C, oshape[0], oshape[1], oshape[2], \
                                               istride[0], istride[1], istride[2], istride[3], ostride[0], ostride[1], \
                                               ostride[2], ostride[3], nm)
  if (!in_u8 && !out_u8) U3D_ZOOM(float, float);
  else if (in_u8 && out_u8) U3D_ZOOM(uint8_t, uint8_t);
  else if (in_u8) U3D_ZOOM(uint8_t, float);
  else return U3D_ERR_UNSUPPORTED;
#undef U3D_ZOOM
  return cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA;
}

int zoom_label(const uint8_t* in, uint8_t* out, const int* ishape, const long long* istride, const int* oshape,
               const long long* ostride, void* ws, size_t ws_bytes, int num_sms, cudaStream_t s) {
  int* idx;
  double* wts;
  if (int rc = build_tables(ishape, oshape, ws, ws_bytes, &idx, &wts, s)) return rc;
  if ((long long)oshape[0] * oshape[1] > 0x7fffffffLL) return U3D_ERR_UNSUPPORTED;
  dim3 g, b;
  row_geometry(oshape, num_sms, &g, &b);
  zoom_label_kernel<<<g, b, 0, s>>>(in, out, idx, wts, oshape[0], oshape[1], oshape[2],
                                                                 istride[0], istride[1], istride[2], ostride[0],
                                                                 ostride[1], ostride[2]);
  return cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA;
}

}  // namespace u3d
