/* unet3d_b200 -- C ABI of the B200 (sm_100a) kernels behind the 3D U-Net train / infer hot path.
 *
 * The reference (icrdr/3D-UNet-Renal-Anatomy-Extraction) is pure Python and has no FFI; the calls
 * it makes on this path are torch.nn layer calls that dispatch to cuDNN/ATen.  Each entry point
 * below replaces one such dispatch (reference file:line given per function) and is what a ctypes
 * binding on the reference side would bind.
 *
 * Conventions
 *   - return 0 on success, a U3D_ERR_* code otherwise; unet3d_last_error_string() explains.
 *   - the caller owns every buffer (device pointers unless stated); the library never allocates
 *     or frees device memory and keeps no references after the call returns.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation.
 *   - activations are 16-bit NDHWC with the channel count padded to a multiple of 16 ("Cp"): bf16 by
 *     default; `act_f16` / `in_f16` / `out_f16` / `x_f16` = 1 selects IEEE fp16 storage for activations,
 *     packed weights AND gradient tensors alike (tcgen05.mma cannot mix operand formats); accumulation is
 *     always fp32.  fp16 gradients are kept in range by one scalar `grad_scale` applied in unet3d_head_bwd;
 *     logits and the stem input are fp32 NCDHW exactly as PyTorch lays them out.
 */
#ifndef UNET3D_B200_H
#define UNET3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define U3D_OK 0
#define U3D_ERR_INVALID 1
#define U3D_ERR_CUDA 2
#define U3D_ERR_TIMEOUT 3
#define U3D_ERR_UNSUPPORTED 4

#define U3D_MAX_SRC 8

const char* unet3d_version(void);
const char* unet3d_last_error_string(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int unet3d_num_sms(void);
/* Size the grids of all later launches on the current device for `limit` SMs (0 = all of them).  The data-parallel step
 * sets it to (SMs - NCCL_MAX_CTAS) for the launches that run while a gradient all-reduce is in flight: the persistent
 * GEMM kernels assign tiles to CTAs statically, so a CTA that has to wait for an SM NCCL occupies doubles its launch. */
int unet3d_set_sm_limit(int limit);

/* One bf16 NDHWC view feeding the A operand of a contraction: a whole tensor, one half of a
 * channel concat (network.py:350 torch.cat -- never materialised), or one stride-2 parity
 * sub-lattice of a tensor (network.py:126 stride-2 convs).  Strides are in BYTES. */
typedef struct unet3d_src {
  const void* ptr;
  int C, W, H, D, N;
  long long sW, sH, sD, sN;
} unet3d_src;

/* Shifted-GEMM convolution on tcgen05 tensor cores (conv_gemm.cu).  Replaces the cuDNN dispatch of
 *   nn.Conv3d forward            network.py:394-395 (k3, stride 1/2), :403 (k1 skip), :411, :541
 *   nn.ConvTranspose3d forward   network.py:312-313 followed by ConstantPad3d :314 (zD/zH/zW)
 *   and the autograd data gradients of all of the above.
 * `tab` (device int32) and `w` (device bf16, packed tiles) are produced by the host plan
 * (3d-unet-renal-anatomy-extraction_b200/plan.py); their layout is documented in csrc/conv_gemm.cuh. */
typedef struct unet3d_conv_args {
  int n_src;
  unet3d_src src[U3D_MAX_SRC];
  const int* tab;
  const void* w;
  void* out;
  void* out2;            /* second output (data gradient of a concat input) or NULL */
  const float* bias;     /* fp32 [n_nblk*nblk] or NULL */
  const void* addend;    /* bf16, same strides as out, added in the epilogue, or NULL */
  const void* addend2;   /* addend for out2 */
  double* stats;         /* fp64 [N][stats_C][2] running (sum, sum^2) for InstanceNorm, or NULL */
  int* err;              /* device int32 error word (0 = ok) */
  int N, D, H, W;        /* tile-grid extents */
  int Dt, n_nblk, nblk, G, n_cg, n_taps, fuse, nbuf, wT, w_stages;
  int in_f16, out_f16;   /* 16-bit format of A + weights / of out + addend: 0 = bf16, 1 = fp16 */
  long long out_sN, out_sD, out_sH, out_sW;   /* ELEMENT strides of out/addend */
  int out_C, stats_C, omul, zD, zH, zW;
  int act;                /* epilogue activation after bias/addend: 0 none, 1 LeakyReLU(0.01) */
  int a_stages;           /* A slabs in flight (2..4) */
  int dense;              /* 1 = every one of the 9 fused (kh,kw) taps is active for every (N block, channel group) */
  long long* dbg_out;     /* NULL, or 8 int64 slots for the MMA warp's cycle counters (tuning experiments only) */
} unet3d_conv_args;
int unet3d_conv_gemm(const unet3d_conv_args* a, void* stream);
size_t unet3d_conv_gemm_smem_bytes(int Dt, int G, int nblk, int fuse, int wT, int w_stages, int a_stages);

/* Pack an fp32 parameter (PyTorch layout) into the 16-bit tile stream unet3d_conv_gemm consumes:
 * out[i] = idx[i] < 0 ? 0 : w[idx[i]]; idx (device int32, n elements, n % 8 == 0 preferred) comes from the host plan. */
int unet3d_weight_pack(const float* w, const int* idx, void* out, long long n, int out_f16, void* stream);

/* Batched form of the gather above, for all layers of a step in one launch.  Job i covers the blocks
 * [first_block[i], first_block[i+1]) of 2048 elements:  out[k] = idx[k] < 0 ? 0 : src[idx[k]],  src = src0 (n0 elements)
 * followed by src1 when src1 != NULL.  mode 0 / 1: out is bf16 / fp16 (weight packing); mode 2: out is fp32 and is
 * multiplied by *scale when scale != NULL (weight-gradient accumulators -> PyTorch parameter layout).
 * Every job's `out` is taken relative to out_base (pass NULL for absolute pointers). */
typedef struct unet3d_gather_job {
  const float* src0;
  const float* src1;
  const int* idx;
  void* out;
  long long n;
  int n0;
  int mode;
} unet3d_gather_job;
int unet3d_gather_multi(const unet3d_gather_job* jobs_dev, const int* first_block_dev, int n_jobs, int n_blocks,
                        const float* scale, void* out_base, void* stream);

/* Weight-gradient accumulators (unet3d_wgrad_gemm's dw[k3][Kp][Np], fp32) -> PyTorch parameter layout, all layers in one
 * launch: out[rowmap[row] + tap + col * col_stride] = dw[tap][row][col] * (*scale), rows with rowmap < 0 (channel padding)
 * and cols >= Ncols skipped.  nn.Conv3d weight (Cout, Cin, k, k, k): rowmap[row] = cin * k3, col_stride = Cin * k3;
 * nn.ConvTranspose3d weight (Cin, Cout, k, k, k): rowmap[row] = cin * Cout * k3, col_stride = k3.
 * Job i covers the blocks [first_block[i], first_block[i+1]) of 256 (tap, row, 8-column group) items. */
typedef struct unet3d_unpack_job {
  const float* dw;
  const int* rowmap;     /* device int32 [Kp] */
  void* out;             /* relative to out_base (pass NULL for absolute pointers) */
  int k3, Kp, Np, Ncols;
  long long col_stride;
} unet3d_unpack_job;
int unet3d_dw_unpack(const unet3d_unpack_job* jobs_dev, const int* first_block_dev, int n_jobs, int n_blocks,
                     const float* scale, void* out_base, void* stream);

/* Weight gradient on tcgen05 tensor cores (wgrad_gemm.cu): dW[tap][cin][cout] = sum_v x[v+tap][cin] dy[v][cout].
 * Replaces the cuDNN backward-filter dispatch of the layers listed above. */
typedef struct unet3d_wgrad_args {
  int n_src;                     /* x views (box 10 x 18, halo) and dy views (box 8 x 16) share one array */
  unet3d_src src[16];
  int box_w[16], box_h[16];
  int box_c[16];                 /* channels per TMA box of each view: 0 / 8 = 16-byte rows without swizzle; 16 / 32 / 64 =
                                    whole rows with SWIZZLE_32B / 64B / 128B (swizzled MN-major operands) */
  const int* tab;                /* device int32 job table, layout in csrc/wgrad_gemm.cuh */
  float* dw;                     /* fp32 accumulator the partial sums are atomically added to */
  int* err;
  int N, D, H, W;                /* tile-grid extents */
  int n_jobs, job_stride;
  int split;                     /* CTAs per job (split over voxel tiles) */
  int x_f16;                     /* x AND dy views stored as fp16 (operands of one MMA must share the format) */
} unet3d_wgrad_args;
int unet3d_wgrad_gemm(const unet3d_wgrad_args* a, void* stream);

/* InstanceNorm3d(affine=False) [+ Dropout3d channel mask] + LeakyReLU(0.01) [+ residual add]
 * network.py:159-160,175-176,401-402,412-416,315-316.  stats come from the conv epilogue. */
int unet3d_in_finalize(const double* stats, const float* drop_scale, float* table, int NC, double count, float eps,
                       void* stream);
/* out = lrelu((y - mean) * scale [+ shift] [+ skip]); table = (mean, scale) per (n, c); shift: NULL, or fp32 per (n, c)
 * (BatchNorm3d's affine offset / dropped-channel constant, network.py:38-69). */
int unet3d_in_apply(const void* y, const void* skip, void* out, const float* table, const float* shift, int N,
                    long long V, int Cp, int act_f16, void* stream);
/* in_bwd_reduce: `out` may be NULL when the norm had no residual input (the activation's sign is then taken from the
 * normalised value and the activation output is not read); dout2 may be NULL; g may be NULL (only the sums are wanted:
 * unet3d_in_bwd_apply then recomputes g from dout, flag bit 1). */
int unet3d_in_bwd_reduce(const void* dout, const void* dout2, const void* out, const void* y, void* g,
                         const float* table, const float* shift, double* sums, int N, long long V, int Cp, int act_f16,
                         void* stream);
/* dy = g * A + y * B + C per (n, c): from (table, sums) for InstanceNorm, or from coef (fp32 [N][Cp][3], BatchNorm).
 * The `zero_last` argument is a flag word: bit 0 = zero the ConstantPad3d planes (network.py:314), bit 1 = `g` holds the
 * UPSTREAM gradient of a norm without residual input and g = dout * lrelu'((y - mean) * scale) is recomputed (InstanceNorm
 * only, coef must be NULL). */
int unet3d_in_bwd_apply(const void* g, const void* y, void* dy, const float* table, const double* sums,
                        const float* coef, double* dsum, int N, int D, int H, int W, int Cp, int zero_last, int act_f16,
                        void* stream);
/* in_bwd_reduce + in_bwd_apply of an InstanceNorm (no pad planes, no bias sum) in ONE launch, for small per-sample
 * slices (V * Cp of a few MB: levels 3-4 of the default net): `sums` ([N][Cp][2] fp64) is WRITTEN, not accumulated.
 * g may be NULL only when out and dout2 are NULL (the activation gradient is then recomputed, as with flag bit 1 above).
 * Same arithmetic as the two-kernel path up to the summation order of the fp32 partial sums. */
int unet3d_in_bwd_small(const void* dout, const void* dout2, const void* out, const void* y, void* g, void* dy,
                        const float* table, double* sums, int N, long long V, int Cp, int act_f16, void* stream);
int unet3d_channel_sum(const void* x, double* dsum, long long NV, int Cp, void* stream);

/* Stem Conv3d(Cin->C,k3,p1)+bias, fp32 NCDHW in -> bf16 NDHWC out (network.py:541,550); w fp32 [Cin][27][Cp].
 * Weight/bias gradient ONE INPUT CHANNEL per call: x points at that channel of sample 0, x_sample_stride = Cin*D*H*W
 * elements; dw fp32 [28][Cp]: that channel's 27 taps, then the bias row (the same for every channel); accumulated
 * atomically. */
int unet3d_stem_fwd(const float* x, const float* w, const float* b, void* out, int N, int Cin, int D, int H, int W,
                    int Cp, int act_f16, void* stream);
int unet3d_stem_wgrad(const float* x, const void* dy, float* dw, int N, int D, int H, int W, long long x_sample_stride,
                      int Cp, int act_f16, void* stream);

/* Head Conv3d(C->K,k1)+bias, bf16 NDHWC in -> fp32 NCDHW logits (network.py:545-547,563), and backward
 * (da bf16 NDHWC; dw fp32 [K][Cp] followed by [K] bias grads, accumulated atomically). */
int unet3d_head_fwd(const void* a, const float* w, const float* b, float* logits, int K, int N, long long V, int Cp,
                    int act_f16, void* stream);
/* grad_scale: optional device scalar multiplied into da (the internal loss scale of fp16 mode); dw is unscaled */
int unet3d_head_bwd(const float* dlogits, const void* a, const float* w, void* da, float* dw, const float* grad_scale,
                    int K, int N, long long V, int Cp, int act_f16, void* stream);

/* Fused softmax + batch Tversky-Dice / focal sums (loss.py:7-11,32-48,70-80) and the logits gradient.
 * sums: fp64 [K][4] = {TP, sum p, sum g, focal}; coef: fp32 [K][4] = {dL/dTP, dL/dSP, focal weight, 0}. */
int unet3d_loss_fwd(const float* logits, const long long* target, double* sums, int K, int N, long long V,
                    float gamma, void* stream);
int unet3d_loss_bwd(const float* logits, const long long* target, const float* coef, const float* grad_scale,
                    float* dlogits, int K, int N, long long V, float gamma, int use_focal, void* stream);

/* Sliding-window blending (trainer.py:72-96: result[:, tile] += softmax(out); result_n[tile] += 1; result / result_n;
 * argmax).  acc: int64 [n_slab][K + 1][Xs][Y][Z], x = slab * Xs + xs, channel K = the weight sum; sums are 2^54 fixed
 * point, so they (and the label map) do not depend on the order in which windows -- or the partial buffers of several
 * GPUs -- are added.  window == NULL: uniform blending (the reference), else a (px, py, pz) fp32 weight map.
 * unet3d_sw_finalize works on ONE slab ([K + 1][n] with n = Xs * Y * Z): labels uint8 [n] (first maximum wins; uncovered
 * voxels -> 0) and / or probs fp32 [n][K] (NaN where uncovered, like the reference's 0 / 0). */
int unet3d_sw_accumulate(const float* logits, const float* window, long long* acc, int K, int px, int py, int pz,
                         int x0, int y0, int z0, int X, int Y, int Z, int Xs, void* stream);
int unet3d_sw_finalize(const long long* acc, uint8_t* labels, float* probs, int K, long long n, void* stream);

/* Train-loader augmentation on a patch that is already in HBM (transform.py:176-301), bit-exact with the reference's
 * numpy arithmetic (csrc/augment.cu).
 *   unet3d_aug_flip   : RandomMirror (transform.py:279-301): out = np.flip(in, axes) for a (X, Y, Z, C) array of 4-byte or
 *                       1-byte elements; in != out.
 *   unet3d_aug_stats  : stats[0] = min, stats[1] = max (order-preserving int encodings, read only by the two kernels
 *                       below), and -- when leaf_off != NULL -- stats[2] = numpy's float32 input.mean(), stats[3] = sum:
 *                       leaf_off[n_leaves + 1] = the leaf boundaries of numpy's pairwise summation of n elements
 *                       (augment.py computes them once per n), leaf_scratch = n_leaves floats.
 *   unet3d_aug_affine : adjust_contrast (which = 0: a = mean) / adjust_brightness (which = 1: a = min), transform.py:176-185:
 *                       out = (x - a) * factor + a in separately rounded fp32 operations.  in-place allowed.
 *   unet3d_aug_gamma  : adjust_gamma (transform.py:188-193): arange = max - min + eps;
 *                       out = power((x - min) / arange, gamma) * arange + min.  in-place allowed. */
int unet3d_aug_flip(const void* in, void* out, int elem_bytes, int X, int Y, int Z, int C, int fx, int fy, int fz,
                    void* stream);
int unet3d_aug_stats(const float* x, long long n, const long long* leaf_off, int n_leaves, float* leaf_scratch,
                     float* stats, void* stream);
int unet3d_aug_affine(const float* x, float* out, long long n, const float* stats, int which, float factor, void* stream);
int unet3d_aug_gamma(const float* x, float* out, long long n, const float* stats, float gamma, float eps, void* stream);

/* MaxPoolBlock = nn.MaxPool3d(kernel_size=2, stride=2) (network.py:452-463) on 16-bit NDHWC (N, D, H, W, Cp), D/H/W even.
 * code: uint8 (N, D/2, H/2, W/2, Cp) = kd*4 + kh*2 + kw of the winner, PyTorch's scan order and tie/NaN rule, i.e.
 * torch's return_indices value is ((2d+kd)*H + 2h+kh)*W + 2w+kw.  bwd writes every element of dx (no pre-zeroing). */
int unet3d_maxpool3d_fwd(const void* x, void* out, uint8_t* code, int N, int D, int H, int W, int Cp, int act_f16,
                         void* stream);
int unet3d_maxpool3d_bwd(const void* dout, const uint8_t* code, void* dx, int N, int D, int H, int W, int Cp,
                         void* stream);

/* Pointwise parts of the attention gate AttBlock (network.py:353-371); its three 1x1x1 convs are unet3d_conv_gemm calls
 * (the middle one with act = 1).  fwd: out = xs * sigmoid(z).  bwd: dxs = dout * r, dz = dout * xs * r * (1 - r),
 * sums fp64 [Cp][2] += {sum dxs, sum dz}.  mid_bwd: dpre = df * lrelu'(f), t = dxs + dpre, sum fp64 [Cp] += sum dpre. */
int unet3d_att_gate_fwd(const void* xs, const void* z, void* out, long long n_elem, int act_f16, void* stream);
int unet3d_att_gate_bwd(const void* dout, const void* xs, const void* z, void* dxs, void* dz, double* sums, long long NV,
                        int Cp, int act_f16, void* stream);
int unet3d_att_mid_bwd(const void* df, const void* f, const void* dxs, void* dpre, void* t, double* sum, long long NV,
                       int Cp, int act_f16, void* stream);

/* Case-level resampling either side of the window loop: transform.rescale / transform.resize (transform.py:32-100), i.e.
 * scipy.ndimage.zoom(order=1, mode='reflect') per channel, bit-exact (float64 corner sum in SciPy's order, one rounding
 * to float32).  Shapes are (x, y, z); strides are in ELEMENTS: (x, y, z, channel) -- any layout (the reference's
 * channel-last numpy volumes, the model's NCDHW input, a padded destination) is a choice of strides and base pointer.
 * norm_host: NULL, or a HOST array [C][4] = {pct_00_5, pct_99_5, mean, std + 1e-8} as float32: the clip + z-score of
 * data.resample_normalize_case (data.py:266-272) applied to the zoomed value (float -> float only, C <= 4).
 * in_u8 / out_u8: uint8 volumes; uint8 -> uint8 is the reference's route for labels with < 3 classes (zoom as float32,
 * truncate back, transform.py:54-71).  unet3d_zoom_label is the >= 3 classes route (transform.py:72-78): one float
 * one-hot volume per class, zoomed, argmax with the first maximum winning.
 * workspace: device memory, 16-byte aligned, >= unet3d_zoom_workspace_bytes(out shape) (per-axis index / weight tables,
 * rebuilt by every call on `stream`). */
size_t unet3d_zoom_workspace_bytes(int out_x, int out_y, int out_z);
int unet3d_zoom_linear(const void* in, int in_u8, void* out, int out_u8, int C, const int in_shape[3],
                       const long long in_stride[4], const int out_shape[3], const long long out_stride[4],
                       const float* norm_host, void* workspace, size_t workspace_bytes, void* stream);
int unet3d_zoom_label(const uint8_t* in, uint8_t* out, const int in_shape[3], const long long in_stride[3],
                      const int out_shape[3], const long long out_stride[3], void* workspace, size_t workspace_bytes,
                      void* stream);

/* Cascade glue (trainer.cascade_predict_case, trainer.py:164-245; data.regions_crop_case, data.py:464-492;
 * transform.remove_small_region, transform.py:5-11).
 * unet3d_ccl_label: connected components of a binary uint8 volume (X, Y, Z), 6-connectivity = scipy.ndimage.label's default
 *   structure.  labels (int32, X*Y*Z): linear index of the component's raster-first voxel, -1 for background; is_root
 *   (uint8, X*Y*Z, may be NULL): 1 at those first voxels -- their sorted positions number the components like SciPy does.
 * unet3d_ccl_stats: roots = the sorted root indices (device int32 [n_roots]); stats (device int32 [n_roots][8]) must be
 *   pre-set to {0, INT_MAX, -1, INT_MAX, -1, INT_MAX, -1, 0} and receives {voxels, xmin, xmax, ymin, ymax, zmin, zmax, 0}
 *   (np.bincount / scipy.ndimage.find_objects).
 * unet3d_region_accumulate: result[(dst0 + i)][:] += pred[i][:], count[dst0 + i] += 1 over a box of box_n voxels; result is
 *   float64 (X, Y, Z, K) channel-last, pred float32 channel-last with ELEMENT strides pstride (x, y, z) (trainer.py:221-222).
 * unet3d_merge_finalize: mean where count > 0, then argmax (K > 1; NaN counts as the maximum like np.argmax) or round half
 *   to even (K == 1) -> uint8 (trainer.py:227-239). */
int unet3d_ccl_label(const uint8_t* mask, int* labels, uint8_t* is_root, int X, int Y, int Z, void* stream);
int unet3d_ccl_stats(const int* labels, const int* roots, int n_roots, int* stats, int X, int Y, int Z, void* stream);
int unet3d_region_accumulate(const float* pred, double* result, int* count, int K, const int box_n[3],
                             const long long pstride[3], const int dst0[3], int Y, int Z, void* stream);
int unet3d_merge_finalize(const double* result, const int* count, uint8_t* labels, int K, long long n_voxels, void* stream);

/* trainer.evaluate_case (trainer.py:348-356): counts [3][256] uint64 (zeroed by the caller) += per label value c
 * { |pred == c AND label == c|, |pred == c|, |label == c| } over two uint8 volumes of n voxels -- what the per-label Dice
 * (loss.dice with alpha = beta = 0.5) needs: TP, FP = |pred| - TP, FN = |label| - TP. */
int unet3d_overlap_counts(const uint8_t* pred, const uint8_t* label, long long n, unsigned long long* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNET3D_B200_H */
