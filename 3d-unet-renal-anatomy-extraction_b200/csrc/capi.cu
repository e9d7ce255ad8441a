// C-ABI of libunet3d_b200.so (declared in include/unet3d_b200.h).  Thin: argument checks,
// TMA tensor-map encoding, launcher calls.  No device allocation, no synchronisation.
#include "../../include/unet3d_b200.h"
#include "conv_gemm.cuh"
#include "wgrad_gemm.cuh"
#include "kernels.cuh"

#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* a = "", long long b = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

int check(int rc, const char* what) {
  if (rc == U3D_OK) return rc;
  if (rc == U3D_ERR_CUDA) {
    cudaError_t e = cudaGetLastError();
    snprintf(g_err, sizeof(g_err), "%s: CUDA error: %s", what, cudaGetErrorString(e));
  } else {
    snprintf(g_err, sizeof(g_err), "%s: error code %d (invalid / unsupported arguments)", what, rc);
  }
  return rc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 5-D (C, W, H, D, N) 16-bit view; box = (chans channels, bw, bh, 1, 1); OOB reads give zeros (conv padding).
// chans = 8: 16-byte rows, no swizzle (weight-gradient bricks); 16 / 32 / 64: SWIZZLE_32B / 64B / 128B (conv_gemm slabs).
int encode_src(CUtensorMap* m, const unet3d_src& s, int bw, int bh, int chans = 8) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(U3D_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found%s");
  if (s.ptr == nullptr || (reinterpret_cast<uintptr_t>(s.ptr) & 15) || s.C % 8 || (s.sW & 15) || (s.sH & 15) ||
      (s.sD & 15) || (s.sN & 15))
    return fail(U3D_ERR_INVALID, "source view must be 16-byte aligned with C %% 8 == 0%s");
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (chans == 16) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (chans == 32) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (chans == 64) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else if (chans != 8) return fail(U3D_ERR_INVALID, "box channel count must be 8, 16, 32 or 64%s");
  cuuint64_t dims[5] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.D, (cuuint64_t)s.N};
  cuuint64_t strides[4] = {(cuuint64_t)s.sW, (cuuint64_t)s.sH, (cuuint64_t)s.sD, (cuuint64_t)s.sN};
  cuuint32_t box[5] = {(cuuint32_t)chans, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(s.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(U3D_ERR_CUDA, "cuTensorMapEncodeTiled failed%s (CUresult %lld)", "", (long long)r);
  return U3D_OK;
}

int g_num_sms[64] = {};      // per device ordinal (one process may drive several GPUs)
int g_sm_limit[64] = {};     // unet3d_set_sm_limit: grids are sized for this many SMs while it is > 0
int num_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  const bool in_range = dev >= 0 && dev < 64;
  if (in_range && g_num_sms[dev] > 0) {
    n = g_num_sms[dev];
  } else {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (in_range) g_num_sms[dev] = n;
  }
  if (in_range && g_sm_limit[dev] > 0 && g_sm_limit[dev] < n) n = g_sm_limit[dev];
  return n;
}

}  // namespace

using namespace u3d;

extern "C" {

const char* unet3d_version(void) { return "unet3d_b200 0.1 (sm_100a)"; }
const char* unet3d_last_error_string(void) { return g_err; }
int unet3d_num_sms(void) { return num_sms(); }
int unet3d_set_sm_limit(int limit) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || limit < 0) return fail(U3D_ERR_INVALID, "set_sm_limit%s");
  g_sm_limit[dev] = limit;
  return U3D_OK;
}

size_t unet3d_conv_gemm_smem_bytes(int Dt, int G, int nblk, int fuse, int wT, int w_stages, int a_stages) {
  return conv_gemm_smem_bytes(Dt, G, nblk, fuse, wT, w_stages, a_stages);
}

int unet3d_conv_gemm(const unet3d_conv_args* a, void* stream) {
  if (!a || a->n_src < 1 || a->n_src > CG_MAX_MAPS || !a->tab || !a->w || !a->out || !a->err)
    return fail(U3D_ERR_INVALID, "conv_gemm: null / out-of-range argument%s");
  const int sms = num_sms();
  if (sms <= 0) return fail(U3D_ERR_CUDA, "conv_gemm: no CUDA device%s");
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < CG_MAX_MAPS; ++i) {
    const unet3d_src& s = a->src[i < a->n_src ? i : 0];
    int rc = encode_src(&p.amap[i], s, CG_WB, CG_HB, 8 * a->G);
    if (rc != U3D_OK) return rc;
  }
  p.tab = a->tab;
  p.w = reinterpret_cast<const bf16*>(a->w);
  p.out = reinterpret_cast<bf16*>(a->out);
  p.out2 = reinterpret_cast<bf16*>(a->out2);
  p.addend2 = reinterpret_cast<const bf16*>(a->addend2);
  p.bias = a->bias;
  p.addend = reinterpret_cast<const bf16*>(a->addend);
  p.stats = a->stats;
  p.err = a->err;
  p.N = a->N; p.D = a->D; p.H = a->H; p.W = a->W;
  p.Dt = a->Dt;
  p.tiles_h = (a->H + CG_HT - 1) / CG_HT;
  p.tiles_w = (a->W + CG_WT - 1) / CG_WT;
  p.segs_d = (a->D + a->Dt - 1) / (a->Dt > 0 ? a->Dt : 1);
  p.n_nblk = a->n_nblk; p.nblk = a->nblk; p.G = a->G; p.n_cg = a->n_cg; p.n_taps = a->n_taps;
  p.fuse = a->fuse;
  p.nbuf = a->nbuf;
  p.wT = a->wT;
  p.w_stages = a->w_stages;
  p.a_stages = a->a_stages;
  p.dbg_out = a->dbg_out;
  p.dense = a->dense && !(getenv("U3D_NO_DENSE") != nullptr);
  p.in_f16 = a->in_f16;
  p.out_f16 = a->out_f16;
  p.out_sN = a->out_sN; p.out_sD = a->out_sD; p.out_sH = a->out_sH; p.out_sW = a->out_sW;
  p.out_C = a->out_C; p.stats_C = a->stats_C; p.omul = a->omul;
  p.zD = a->zD; p.zH = a->zH; p.zW = a->zW;
  p.act = a->act;
  { const char* e = getenv("U3D_DBG"); p.dbg = e ? atoi(e) : 0; }
  p.n_work = a->n_nblk * a->N * p.segs_d * p.tiles_h * p.tiles_w;
  return check(conv_gemm_launch(p, sms, reinterpret_cast<cudaStream_t>(stream)), "conv_gemm");
}

int unet3d_wgrad_gemm(const unet3d_wgrad_args* a, void* stream) {
  if (!a || a->n_src < 1 || a->n_src > WG_MAX_MAPS || !a->tab || !a->dw || !a->err || a->n_jobs < 1 || a->split < 1)
    return fail(U3D_ERR_INVALID, "wgrad_gemm: null / out-of-range argument%s");
  const int sms = num_sms();
  if (sms <= 0) return fail(U3D_ERR_CUDA, "wgrad_gemm: no CUDA device%s");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < WG_MAX_MAPS; ++i) {
    const unet3d_src& s = a->src[i < a->n_src ? i : 0];
    const int j = i < a->n_src ? i : 0;
    int rc = encode_src(&p.map[i], s, a->box_w[j], a->box_h[j], a->box_c[j] > 0 ? a->box_c[j] : 8);
    if (rc != U3D_OK) return rc;
  }
  p.tab = a->tab;
  p.dw = a->dw;
  p.err = a->err;
  p.N = a->N; p.D = a->D; p.H = a->H; p.W = a->W;
  p.tiles_h = (a->H + CG_HT - 1) / CG_HT;
  p.tiles_w = (a->W + CG_WT - 1) / CG_WT;
  p.n_jobs = a->n_jobs;
  p.job_stride = a->job_stride;
  p.split = a->split;
  p.x_f16 = a->x_f16;
  { const char* e = getenv("U3D_DBG"); p.dbg = e ? atoi(e) : 0; }
  return check(wgrad_gemm_launch(p, reinterpret_cast<cudaStream_t>(stream)), "wgrad_gemm");
}

int unet3d_weight_pack(const float* w, const int* idx, void* out, long long n, int out_f16, void* stream) {
  return check(weight_pack(w, idx, out, n, out_f16, num_sms(), (cudaStream_t)stream), "weight_pack");
}
int unet3d_in_finalize(const double* stats, const float* drop_scale, float* table, int NC, double count, float eps,
                       void* stream) {
  return check(in_finalize(stats, drop_scale, table, NC, count, eps, (cudaStream_t)stream), "in_finalize");
}
int unet3d_in_apply(const void* y, const void* skip, void* out, const float* table, const float* shift, int N,
                    long long V, int Cp, int act_f16, void* stream) {
  return check(in_apply((const bf16*)y, (const bf16*)skip, (bf16*)out, table, shift, N, V, Cp, act_f16, num_sms(), (cudaStream_t)stream),
               "in_apply");
}
int unet3d_in_bwd_reduce(const void* dout, const void* dout2, const void* out, const void* y, void* g,
                         const float* table, const float* shift, double* sums, int N, long long V, int Cp, int act_f16,
                         void* stream) {
  return check(in_bwd_reduce((const bf16*)dout, (const bf16*)dout2, (const bf16*)out, (const bf16*)y, (bf16*)g, table,
                             shift, sums, N, V, Cp, act_f16, num_sms(), (cudaStream_t)stream),
               "in_bwd_reduce");
}
int unet3d_in_bwd_apply(const void* g, const void* y, void* dy, const float* table, const double* sums,
                        const float* coef, double* dsum, int N, int D, int H, int W, int Cp, int zero_last, int act_f16,
                        void* stream) {
  return check(in_bwd_apply((const bf16*)g, (const bf16*)y, (bf16*)dy, table, sums, coef, dsum, N, D, H, W, Cp, zero_last,
                            act_f16, num_sms(), (cudaStream_t)stream),
               "in_bwd_apply");
}
int unet3d_in_bwd_small(const void* dout, const void* dout2, const void* out, const void* y, void* g, void* dy,
                        const float* table, double* sums, int N, long long V, int Cp, int act_f16, void* stream) {
  return check(in_bwd_small((const bf16*)dout, (const bf16*)dout2, (const bf16*)out, (const bf16*)y, (bf16*)g, (bf16*)dy, table,
                            sums, N, V, Cp, act_f16, num_sms(), (cudaStream_t)stream),
               "in_bwd_small");
}
int unet3d_channel_sum(const void* x, double* dsum, long long NV, int Cp, void* stream) {
  return check(channel_sum((const bf16*)x, dsum, NV, Cp, num_sms(), (cudaStream_t)stream), "channel_sum");
}
int unet3d_stem_fwd(const float* x, const float* w, const float* b, void* out, int N, int Cin, int D, int H, int W,
                    int Cp, int act_f16, void* stream) {
  return check(stem_fwd(x, w, b, (bf16*)out, N, Cin, D, H, W, Cp, act_f16, num_sms(), (cudaStream_t)stream), "stem_fwd");
}
int unet3d_stem_wgrad(const float* x, const void* dy, float* dw, int N, int D, int H, int W, long long x_sample_stride,
                      int Cp, int act_f16, void* stream) {
  return check(stem_wgrad(x, (const bf16*)dy, dw, N, D, H, W, x_sample_stride, Cp, act_f16, num_sms(),
                          (cudaStream_t)stream), "stem_wgrad");
}
int unet3d_head_fwd(const void* a, const float* w, const float* b, float* logits, int K, int N, long long V, int Cp,
                    int act_f16, void* stream) {
  return check(head_fwd((const bf16*)a, w, b, logits, K, N, V, Cp, act_f16, num_sms(), (cudaStream_t)stream), "head_fwd");
}
int unet3d_head_bwd(const float* dlogits, const void* a, const float* w, void* da, float* dw, const float* grad_scale,
                    int K, int N, long long V, int Cp, int act_f16, void* stream) {
  return check(head_bwd(dlogits, (const bf16*)a, w, (bf16*)da, dw, grad_scale, K, N, V, Cp, act_f16, num_sms(),
                        (cudaStream_t)stream),
               "head_bwd");
}
int unet3d_loss_fwd(const float* logits, const long long* target, double* sums, int K, int N, long long V,
                    float gamma, void* stream) {
  return check(loss_fwd(logits, target, sums, K, N, V, gamma, num_sms(), (cudaStream_t)stream), "loss_fwd");
}
int unet3d_loss_bwd(const float* logits, const long long* target, const float* coef, const float* grad_scale,
                    float* dlogits, int K, int N, long long V, float gamma, int use_focal, void* stream) {
  return check(loss_bwd(logits, target, coef, grad_scale, dlogits, K, N, V, gamma, use_focal, num_sms(),
                        (cudaStream_t)stream),
               "loss_bwd");
}
int unet3d_sw_accumulate(const float* logits, const float* window, long long* acc, int K, int px, int py, int pz,
                         int x0, int y0, int z0, int X, int Y, int Z, int Xs, void* stream) {
  return check(sw_accumulate(logits, window, acc, K, px, py, pz, x0, y0, z0, X, Y, Z, Xs, num_sms(), (cudaStream_t)stream),
               "sw_accumulate");
}
int unet3d_sw_finalize(const long long* acc, uint8_t* labels, float* probs, int K, long long n, void* stream) {
  return check(sw_finalize(acc, labels, probs, K, n, num_sms(), (cudaStream_t)stream), "sw_finalize");
}

int unet3d_aug_flip(const void* in, void* out, int elem_bytes, int X, int Y, int Z, int C, int fx, int fy, int fz,
                    void* stream) {
  return check(aug_flip(in, out, elem_bytes, X, Y, Z, C, fx, fy, fz, num_sms(), (cudaStream_t)stream), "aug_flip");
}
int unet3d_aug_stats(const float* x, long long n, const long long* leaf_off, int n_leaves, float* leaf_scratch,
                     float* stats, void* stream) {
  return check(aug_stats(x, n, leaf_off, n_leaves, leaf_scratch, stats, num_sms(), (cudaStream_t)stream), "aug_stats");
}
int unet3d_aug_affine(const float* x, float* out, long long n, const float* stats, int which, float factor, void* stream) {
  return check(aug_affine(x, out, n, stats, which, factor, num_sms(), (cudaStream_t)stream), "aug_affine");
}
int unet3d_aug_gamma(const float* x, float* out, long long n, const float* stats, float gamma, float eps, void* stream) {
  return check(aug_gamma(x, out, n, stats, gamma, eps, num_sms(), (cudaStream_t)stream), "aug_gamma");
}
int unet3d_maxpool3d_fwd(const void* x, void* out, uint8_t* code, int N, int D, int H, int W, int Cp, int act_f16,
                         void* stream) {
  return check(maxpool_fwd((const bf16*)x, (bf16*)out, code, N, D, H, W, Cp, act_f16, num_sms(), (cudaStream_t)stream),
               "maxpool3d_fwd");
}
int unet3d_maxpool3d_bwd(const void* dout, const uint8_t* code, void* dx, int N, int D, int H, int W, int Cp,
                         void* stream) {
  return check(maxpool_bwd((const bf16*)dout, code, (bf16*)dx, N, D, H, W, Cp, num_sms(), (cudaStream_t)stream),
               "maxpool3d_bwd");
}
int unet3d_att_gate_fwd(const void* xs, const void* z, void* out, long long n_elem, int act_f16, void* stream) {
  return check(att_gate_fwd((const bf16*)xs, (const bf16*)z, (bf16*)out, n_elem, act_f16, num_sms(), (cudaStream_t)stream),
               "att_gate_fwd");
}
int unet3d_att_gate_bwd(const void* dout, const void* xs, const void* z, void* dxs, void* dz, double* sums, long long NV,
                        int Cp, int act_f16, void* stream) {
  return check(att_gate_bwd((const bf16*)dout, (const bf16*)xs, (const bf16*)z, (bf16*)dxs, (bf16*)dz, sums, NV, Cp, act_f16,
                            num_sms(), (cudaStream_t)stream),
               "att_gate_bwd");
}
int unet3d_att_mid_bwd(const void* df, const void* f, const void* dxs, void* dpre, void* t, double* sum, long long NV,
                       int Cp, int act_f16, void* stream) {
  return check(att_mid_bwd((const bf16*)df, (const bf16*)f, (const bf16*)dxs, (bf16*)dpre, (bf16*)t, sum, NV, Cp, act_f16,
                           num_sms(), (cudaStream_t)stream),
               "att_mid_bwd");
}
int unet3d_gather_multi(const unet3d_gather_job* jobs_dev, const int* first_block_dev, int n_jobs, int n_blocks,
                        const float* scale, void* out_base, void* stream) {
  return check(gather_multi(jobs_dev, first_block_dev, n_jobs, n_blocks, scale, out_base, (cudaStream_t)stream), "gather_multi");
}
int unet3d_dw_unpack(const unet3d_unpack_job* jobs_dev, const int* first_block_dev, int n_jobs, int n_blocks,
                     const float* scale, void* out_base, void* stream) {
  return check(dw_unpack(jobs_dev, first_block_dev, n_jobs, n_blocks, scale, out_base, (cudaStream_t)stream), "dw_unpack");
}
size_t unet3d_zoom_workspace_bytes(int out_x, int out_y, int out_z) { return zoom_workspace_bytes(out_x, out_y, out_z); }
int unet3d_zoom_linear(const void* in, int in_u8, void* out, int out_u8, int C, const int in_shape[3],
                       const long long in_stride[4], const int out_shape[3], const long long out_stride[4],
                       const float* norm_host, void* workspace, size_t workspace_bytes, void* stream) {
  if (!in || !out || !in_shape || !in_stride || !out_shape || !out_stride) return check(U3D_ERR_INVALID, "zoom_linear");
  return check(zoom_linear(in, in_u8, out, out_u8, C, in_shape, in_stride, out_shape, out_stride, norm_host, workspace,
                           workspace_bytes, num_sms(), (cudaStream_t)stream),
               "zoom_linear");
}
int unet3d_zoom_label(const uint8_t* in, uint8_t* out, const int in_shape[3], const long long in_stride[3],
                      const int out_shape[3], const long long out_stride[3], void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (!in || !out || !in_shape || !in_stride || !out_shape || !out_stride) return check(U3D_ERR_INVALID, "zoom_label");
  return check(zoom_label(in, out, in_shape, in_stride, out_shape, out_stride, workspace, workspace_bytes, num_sms(),
                          (cudaStream_t)stream),
               "zoom_label");
}
int unet3d_ccl_label(const uint8_t* mask, int* labels, uint8_t* is_root, int X, int Y, int Z, void* stream) {
  if (!mask || !labels) return check(U3D_ERR_INVALID, "ccl_label");
  return check(ccl_label(mask, labels, is_root, X, Y, Z, num_sms(), (cudaStream_t)stream), "ccl_label");
}
int unet3d_ccl_stats(const int* labels, const int* roots, int n_roots, int* stats, int X, int Y, int Z, void* stream) {
  if (!labels || !roots || !stats) return check(U3D_ERR_INVALID, "ccl_stats");
  return check(ccl_stats(labels, roots, n_roots, stats, X, Y, Z, num_sms(), (cudaStream_t)stream), "ccl_stats");
}
int unet3d_region_accumulate(const float* pred, double* result, int* count, int K, const int box_n[3],
                             const long long pstride[3], const int dst0[3], int Y, int Z, void* stream) {
  if (!pred || !result || !count || !box_n || !pstride || !dst0) return check(U3D_ERR_INVALID, "region_accumulate");
  return check(region_accumulate(pred, result, count, K, box_n, pstride, dst0, Y, Z, num_sms(), (cudaStream_t)stream),
               "region_accumulate");
}
int unet3d_merge_finalize(const double* result, const int* count, uint8_t* labels, int K, long long n_voxels, void* stream) {
  if (!result || !count || !labels) return check(U3D_ERR_INVALID, "merge_finalize");
  return check(merge_finalize(result, count, labels, K, n_voxels, num_sms(), (cudaStream_t)stream), "merge_finalize");
}
int unet3d_overlap_counts(const uint8_t* pred, const uint8_t* label, long long n, unsigned long long* counts, void* stream) {
  if (!pred || !label || !counts) return check(U3D_ERR_INVALID, "overlap_counts");
  return check(overlap_counts(pred, label, n, counts, num_sms(), (cudaStream_t)stream), "overlap_counts");
}

}  // extern "C"
