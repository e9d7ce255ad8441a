// Cascade inference glue: the steps between the coarse and the detail model.
//
//   data.regions_crop_case (data.py:464-492) / transform.remove_small_region (transform.py:5-11):
//       scipy.ndimage.label (6-connectivity), component sizes, bounding boxes  -> ccl_* kernels (union-find in HBM)
//   trainer.cascade_predict_case merge (trainer.py:189-241):
//       result[bbox] += region probabilities; result_n[bbox] += 1; mean; softmax + argmax -> region_accumulate / merge_finalize
//
// Integer work (component numbering, sizes, boxes, label maps) is bit-exact with SciPy / numpy: a component's id is the
// linear index of its raster-first voxel, so sorting the roots reproduces scipy.ndimage.label's numbering.
#include "kernels.cuh"

namespace u3d {
namespace {

// loads bypass the (non-coherent) L1: a stale parent only costs extra iterations -- every link is validated by the
// atomicMin that installs it -- but fresh values converge faster
__device__ __forceinline__ int uf_find(const int* L, int i) {
  const volatile int* V = L;
  int p = V[i];
  while (p != i) {
    i = p;
    p = V[i];
  }
  return i;
}

// union by smaller index: the root of a tree is always its smallest member (L[i] <= i everywhere)
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  bool done = false;
  while (!done) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a < b) {
      const int old = atomicMin(&L[b], a);
      done = old == b;
      b = old;
    } else if (b < a) {
      const int old = atomicMin(&L[a], b);
      done = old == a;
      a = old;
    } else {
      done = true;
    }
  }
}

__global__ void ccl_init_kernel(const uint8_t* __restrict__ mask, int* __restrict__ L, int Y, int Z, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int l = -1;
    if (mask[i]) {
      // point straight at the start of the z-run this voxel belongs to (first merge along the contiguous axis without a
      // single atomic; the union pass links runs across y and x)
      long long r = i;
      int z = (int)(i % Z);
      while (z > 0 && mask[r - 1]) { --r; --z; }
      l = (int)r;
    }
    L[i] = l;
  }
}

__global__ void ccl_merge_kernel(const uint8_t* __restrict__ mask, int* __restrict__ L, int Y, int Z, long long n) {
  const long long YZ = (long long)Y * Z;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (!mask[i]) continue;
    const int z = (int)(i % Z), y = (int)((i / Z) % Y);
    // every voxel of a z-run already points at the run's first voxel, so two runs need ONE link: where their overlap
    // starts, i.e. unless the pair one step back along z is foreground too (it has made, or will make, the same link)
    const bool back = z > 0 && mask[i - 1];
    if (y > 0 && mask[i - Z] && !(back && mask[i - Z - 1])) uf_union(L, (int)i, (int)(i - Z));
    if (i >= YZ && mask[i - YZ] && !(back && mask[i - YZ - 1])) uf_union(L, (int)i, (int)(i - YZ));
  }
}

__global__ void ccl_compress_kernel(int* __restrict__ L, uint8_t* __restrict__ is_root, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int l = L[i];
    uint8_t r = 0;
    if (l >= 0) {
      const int root = uf_find(L, (int)i);
      L[i] = root;             // roots never change after the merge pass, so concurrent compression is safe
      r = root == (int)i;
    }
    if (is_root) is_root[i] = r;
  }
}

// per-component voxel count and bounding box; roots[] is sorted (raster order), a voxel finds its component by bisection.
// stats[c] = {count, xmin, xmax, ymin, ymax, zmin, zmax, 0}; the caller initialises min to INT_MAX and max / count to 0/-1.
__global__ void ccl_stats_kernel(const int* __restrict__ L, const int* __restrict__ roots, int n_roots,
                                 int* __restrict__ stats, int Y, int Z, long long n) {
  // a CTA walks consecutive voxels, so nearly all of its foreground belongs to one component: that component's seven
  // numbers are aggregated in shared memory and flushed once (millions of same-address L2 atomics otherwise)
  __shared__ int s_comp, s_stat[7];
  if (threadIdx.x == 0) {
    s_comp = -1;
    s_stat[0] = 0;
    s_stat[1] = s_stat[3] = s_stat[5] = 0x7fffffff;
    s_stat[2] = s_stat[4] = s_stat[6] = -1;
  }
  __syncthreads();
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 - threadIdx.x % 32 < n;
       i0 += (long long)gridDim.x * blockDim.x) {
    const int l = i0 < n ? L[i0] : -1;
    int comp = -1;
    if (l >= 0) {
      int lo = 0, hi = n_roots;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&roots[mid]) <= l) lo = mid; else hi = mid;
      }
      comp = lo;
    }
    const int z = (int)(i0 % Z), y = (int)((i0 / Z) % Y), x = (int)(i0 / ((long long)Y * Z));
    // warp aggregation: lanes of one component elect a leader
    unsigned todo = __ballot_sync(0xffffffffu, comp >= 0);
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int lc = __shfl_sync(0xffffffffu, comp, leader);
      const unsigned grp = __ballot_sync(0xffffffffu, comp == lc);
      const bool mine = comp == lc;
      const int cnt = __popc(grp);
      const int xmin = __reduce_min_sync(0xffffffffu, mine ? x : 0x7fffffff);
      const int xmax = __reduce_max_sync(0xffffffffu, mine ? x : -1);
      const int ymin = __reduce_min_sync(0xffffffffu, mine ? y : 0x7fffffff);
      const int ymax = __reduce_max_sync(0xffffffffu, mine ? y : -1);
      const int zmin = __reduce_min_sync(0xffffffffu, mine ? z : 0x7fffffff);
      const int zmax = __reduce_max_sync(0xffffffffu, mine ? z : -1);
      if ((int)(threadIdx.x & 31) == leader) {
        int owner = s_comp;
        if (owner < 0) {
          owner = atomicCAS(&s_comp, -1, lc);
          if (owner < 0) owner = lc;
        }
        int* s = owner == lc ? s_stat : stats + 8 * lc;
        atomicAdd(s + 0, cnt);
        atomicMin(s + 1, xmin); atomicMax(s + 2, xmax);
        atomicMin(s + 3, ymin); atomicMax(s + 4, ymax);
        atomicMin(s + 5, zmin); atomicMax(s + 6, zmax);
      }
      todo &= ~grp;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_comp >= 0 && s_stat[0] > 0) {
    int* s = stats + 8 * s_comp;
    atomicAdd(s + 0, s_stat[0]);
    atomicMin(s + 1, s_stat[1]); atomicMax(s + 2, s_stat[2]);
    atomicMin(s + 3, s_stat[3]); atomicMax(s + 4, s_stat[4]);
    atomicMin(s + 5, s_stat[5]); atomicMax(s + 6, s_stat[6]);
  }
}

// result[x0+i, y0+j, z0+k, :] += pred[i0+i, j0+j, k0+k, :] (float64 running sums, like the reference's numpy arrays);
// count[...] += 1.  Channel-last on both sides; one thread per voxel of the overlap box.
__global__ void region_accumulate_kernel(const float* __restrict__ pred, double* __restrict__ result, int* __restrict__ count,
                                         int K, int nx, int ny, int nz, long long psx, long long psy, long long psz,
                                         int Y, int Z, int x0, int y0, int z0) {
  const long long n = (long long)nx * ny * nz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % nz), j = (int)((i / nz) % ny), ii = (int)(i / ((long long)nz * ny));
    const float* p = pred + ii * psx + j * psy + k * psz;
    const long long o = ((long long)(x0 + ii) * Y + (y0 + j)) * Z + (z0 + k);
    for (int c = 0; c < K; ++c) result[o * K + c] += (double)__ldg(p + c);
    count[o] += 1;
  }
}

// mean where covered, then argmax (softmax is monotonic; NaN -- the uncovered border of a region's tile grid -- counts as
// the maximum, like np.argmax) or, for one class, round half to even.
__global__ void merge_finalize_kernel(const double* __restrict__ result, const int* __restrict__ count,
                                      uint8_t* __restrict__ labels, int K, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cnt = count[i];
    const double inv = cnt > 0 ? (double)cnt : 1.0;
    if (K == 1) {
      const double v = __ddiv_rn(result[i], inv);
      labels[i] = (uint8_t)(int)rint(v);                     // np.around; NaN -> 0 like numpy's uint8 cast of NaN on x86?  see host
      continue;
    }
    int best = 0;
    double bv = __ddiv_rn(result[i * K], inv);
    if (!(bv != bv)) {
      for (int c = 1; c < K; ++c) {
        const double v = __ddiv_rn(result[i * K + c], inv);
        if (v != v) { best = c; break; }                     // first NaN wins
        if (v > bv) { bv = v; best = c; }
      }
    }
    labels[i] = (uint8_t)best;
  }
}

inline int grid_for_n(long long n, int per_block, int num_sms) {
  const long long want = (n + per_block - 1) / per_block, cap = (long long)num_sms * 32;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

// trainer.evaluate_case (trainer.py:348-356): per label c the three counts its Dice needs, |pred == c AND label == c|,
// |pred == c|, |label == c|, in one pass over the two uint8 volumes (the reference makes 2 float32 masks and 3 masked
// sums per class on the CPU).  Exact integer counts; per-block shared-memory histogram, one atomic per (block, bin).
__global__ void overlap_counts_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ label, long long n,
                                      unsigned long long* __restrict__ counts /*[3][256]*/) {
  __shared__ unsigned int h[3][256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) (&h[0][0])[i] = 0u;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned p = pred[i], l = label[i];
    if (p == l) atomicAdd(&h[0][p], 1u);
    atomicAdd(&h[1][p], 1u);
    atomicAdd(&h[2][l], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
    const unsigned v = (&h[0][0])[i];
    if (v) atomicAdd(&counts[i], (unsigned long long)v);
  }
}

}  // namespace

int overlap_counts(const uint8_t* pred, const uint8_t* label, long long n, unsigned long long* counts, int num_sms,
                   cudaStream_t s) {
  if (n < 1) return U3D_ERR_INVALID;
  long long need = (n + 256 * 16 - 1) / (256 * 16);
  const int g = (int)(need < (long long)num_sms * 4 ? need : (long long)num_sms * 4);
  overlap_counts_kernel<<<g, 256, 0, s>>>(pred, label, n, counts);
  return cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA;
}

int ccl_label(const uint8_t* mask, int* labels, uint8_t* is_root, int X, int Y, int Z, int num_sms, cudaStream_t s) {
  const long long n = (long long)X * Y * Z;
  if (X < 1 || Y < 1 || Z < 1 || n > 0x7fffffffLL) return U3D_ERR_UNSUPPORTED;
  const int g = grid_for_n(n, 256, num_sms);
  ccl_init_kernel<<<g, 256, 0, s>>>(mask, labels, Y, Z, n);
  ccl_merge_kernel<<<g, 256, 0, s>>>(mask, labels, Y, Z, n);
  ccl_compress_kernel<<<g, 256, 0, s>>>(labels, is_root, n);
  return cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA;
}

int ccl_stats(const int* labels, const int* roots, int n_roots, int* stats, int X, int Y, int Z, int num_sms,
              cudaStream_t s) {
  const long long n = (long long)X * Y * Z;
  if (n_roots < 1 || n > 0x7fffffffLL) return U3D_ERR_INVALID;
  ccl_stats_kernel<<<grid_for_n(n, 256, num_sms), 256, 0, s>>>(labels, roots, n_roots, stats, Y, Z, n);
  return cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA;
}

int region_accumulate(const float* pred, double* result, int* count, int K, const int* box_n, const long long* pstride,
                      const int* dst0, int Y, int Z, int num_sms, cudaStream_t s) {
  if (K < 1 || box_n[0] < 1 || box_n[1] < 1 || box_n[2] < 1) return U3D_ERR_INVALID;
  const long long n = (long long)box_n[0] * box_n[1] * box_n[2];
  region_accumulate_kernel<<<grid_for_n(n, 256, num_sms), 256, 0, s>>>(pred, result, count, K, box_n[0], box_n[1], box_n[2],
                                                                       pstride[0], pstride[1], pstride[2], Y, Z, dst0[0],
                                                                       dst0[1], dst0[2]);
  return cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA;
}

int merge_finalize(const double* result, const int* count, uint8_t* labels, int K, long long n, int num_sms, cudaStream_t s) {
  if (K < 1 || n < 1) return U3D_ERR_INVALID;
  merge_finalize_kernel<<<grid_for_n(n, 256, num_sms), 256, 0, s>>>(result, count, labels, K, n);
  return cudaGetLastError() == cudaSuccess ? U3D_OK : U3D_ERR_CUDA;
}

}  // namespace u3d
