// Internal launcher declarations (host side); the C-ABI in capi.cu forwards to these.
#pragma once
#include "common.cuh"

namespace u3d {

int weight_pack(const float* w, const int* idx, void* out, long long n, int f16, int num_sms, cudaStream_t s);
int in_finalize(const double* stats, const float* drop, float* table, int NC, double count, float eps, cudaStream_t s);
int in_apply(const bf16* y, const bf16* skip, bf16* out, const float* table, const float* shift, int N, long long V,
             int Cp, int af, int num_sms, cudaStream_t s);
int in_bwd_reduce(const bf16* dout, const bf16* dout2, const bf16* out, const bf16* y, bf16* g, const float* table,
                  const float* shift, double* sums, int N, long long V, int Cp, int af, int num_sms, cudaStream_t s);
int in_bwd_apply(const bf16* g, const bf16* y, bf16* dy, const float* table, const double* sums, const float* coef,
                 double* dsum, int N, int D, int H, int W, int Cp, int zero_last, int af, int num_sms, cudaStream_t s);
int in_bwd_small(const bf16* dout, const bf16* dout2, const bf16* out, const bf16* y, bf16* g, bf16* dy, const float* table,
                 double* sums, int N, long long V, int Cp, int af, int num_sms, cudaStream_t s);
int channel_sum(const bf16* x, double* dsum, long long NV, int Cp, int num_sms, cudaStream_t s);
int stem_fwd(const float* x, const float* w, const float* b, bf16* out, int N, int Cin, int D, int H, int W, int Cp,
             int af, int num_sms, cudaStream_t s);
int stem_wgrad(const float* x, const bf16* dy, float* dw, int N, int D, int H, int W, long long x_sN, int Cp, int af,
               int num_sms, cudaStream_t s);
int head_fwd(const bf16* a, const float* w, const float* b, float* logits, int K, int N, long long V, int Cp, int af,
             int num_sms, cudaStream_t s);
int head_bwd(const float* dl, const bf16* a, const float* w, bf16* da, float* dw, const float* gscale, int K, int N,
             long long V, int Cp, int af, int num_sms, cudaStream_t s);
int loss_fwd(const float* logits, const long long* target, double* sums, int K, int N, long long V, float gamma,
             int num_sms, cudaStream_t s);
int loss_bwd(const float* logits, const long long* target, const float* coef, const float* gscale, float* dlogits, int K,
             int N, long long V, float gamma, int use_focal, int num_sms, cudaStream_t s);
int sw_accumulate(const float* logits, const float* window, long long* acc, int K, int px, int py, int pz,
                  int x0, int y0, int z0, int X, int Y, int Z, int Xs, int num_sms, cudaStream_t s);
int sw_finalize(const long long* acc, uint8_t* labels, float* probs, int K, long long n, int num_sms, cudaStream_t s);
// csrc/augment.cu
int aug_flip(const void* in, void* out, int elem_bytes, int X, int Y, int Z, int C, int fx, int fy, int fz, int num_sms,
             cudaStream_t s);
int aug_stats(const float* x, long long n, const long long* leaf_off, int n_leaves, float* leaf_scratch, float* stats,
              int num_sms, cudaStream_t s);
int aug_affine(const float* x, float* out, long long n, const float* stats, int which, float factor, int num_sms,
               cudaStream_t s);
int aug_gamma(const float* x, float* out, long long n, const float* stats, float gamma, float eps, int num_sms,
              cudaStream_t s);
int maxpool_fwd(const bf16* x, bf16* out, uint8_t* code, int N, int D, int H, int W, int Cp, int af, int num_sms,
                cudaStream_t s);
int maxpool_bwd(const bf16* dout, const uint8_t* code, bf16* dx, int N, int D, int H, int W, int Cp, int num_sms,
                cudaStream_t s);
int att_gate_fwd(const bf16* xs, const bf16* z, bf16* out, long long n_elem, int af, int num_sms, cudaStream_t s);
int att_gate_bwd(const bf16* dout, const bf16* xs, const bf16* z, bf16* dxs, bf16* dz, double* sums, long long NV, int Cp,
                 int af, int num_sms, cudaStream_t s);
int att_mid_bwd(const bf16* df, const bf16* f, const bf16* dxs, bf16* dpre, bf16* t, double* sum, long long NV, int Cp,
                int af, int num_sms, cudaStream_t s);
int gather_multi(const unet3d_gather_job* jobs, const int* first_block, int n_jobs, int n_blocks, const float* scale,
                 void* out_base, cudaStream_t s);
int dw_unpack(const unet3d_unpack_job* jobs, const int* first_block, int n_jobs, int n_blocks, const float* scale,
              void* out_base, cudaStream_t s);
size_t zoom_workspace_bytes(int ox, int oy, int oz);
int zoom_linear(const void* in, int in_u8, void* out, int out_u8, int C, const int* ishape, const long long* istride,
                const int* oshape, const long long* ostride, const float* norm_host, void* ws, size_t ws_bytes,
                int num_sms, cudaStream_t s);
int zoom_label(const uint8_t* in, uint8_t* out, const int* ishape, const long long* istride, const int* oshape,
               const long long* ostride, void* ws, size_t ws_bytes, int num_sms, cudaStream_t s);
int ccl_label(const uint8_t* mask, int* labels, uint8_t* is_root, int X, int Y, int Z, int num_sms, cudaStream_t s);
int ccl_stats(const int* labels, const int* roots, int n_roots, int* stats, int X, int Y, int Z, int num_sms,
              cudaStream_t s);
int region_accumulate(const float* pred, double* result, int* count, int K, const int* box_n, const long long* pstride,
                      const int* dst0, int Y, int Z, int num_sms, cudaStream_t s);
int merge_finalize(const double* result, const int* count, uint8_t* labels, int K, long long n, int num_sms, cudaStream_t s);
int overlap_counts(const uint8_t* pred, const uint8_t* label, long long n, unsigned long long* counts, int num_sms,
                   cudaStream_t s);

}  // namespace u3d
