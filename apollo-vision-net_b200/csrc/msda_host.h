// Host-side glue shared by the translation units of libmsda_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/msda_b200.h"

namespace msda {

// Arguments of the op-boundary calls, validated by abi.cu before any launch.
struct Problem {
  const void* value = nullptr;
  const int64_t* shapes = nullptr;
  const int64_t* starts = nullptr;
  const void* loc = nullptr;
  const void* attn = nullptr;
  void* out = nullptr;
  const void* grad_out = nullptr;
  float* g_value = nullptr;
  float* g_loc = nullptr;
  float* g_attn = nullptr;
  int B = 0, Nk = 0, M = 0, Dh = 0, L = 0, Nq = 0, P = 0;
  int value_dtype = MSDA_F32, coord_dtype = MSDA_F32;
  // > 0: `loc` holds PIXEL coordinates of the (single) level and the location gradient is multiplied by this
  // factor instead of the level size (the DCNv3 entry points: factor = offset_scale)
  float pixel_scale = 0.f;
};

// Arguments of the fused attention cores (sca_* / tsa_*).
struct FusedProblem {
  const void* value = nullptr;
  const int64_t* shapes = nullptr;
  const int64_t* starts = nullptr;
  const void* offsets = nullptr;       // coord_dtype
  const void* logits = nullptr;
  const float* ref = nullptr;          // ref_cam (SCA) or ref points (TSA / decoder)
  const uint8_t* bev_mask = nullptr;   // SCA only
  const uint32_t* hit_bits = nullptr;  // SCA only
  void* out = nullptr;
  float* attn_out = nullptr;
  const void* g_out = nullptr;
  void* g_value = nullptr;             // fp32, or fp16 scaled by *acc_scale when acc_half
  const float* acc_scale = nullptr;
  int acc_half = 0;
  void* g_tail = nullptr;              // replicas of the fp16 accumulator's coarse tail (see fused.cu)
  int tail_copies = 0, tail_px = 0;
  void* g_offsets = nullptr;           // coord_dtype
  void* g_logits = nullptr;
  int bs = 0, groups = 0;              // groups = num_cam (SCA) or Q (TSA)
  int Nk = 0, M = 0, Dh = 0, L = 0, P = 0, D = 1, Nq = 0, bev_w = 0;
  float clamp = -1.f;
  int value_dtype = MSDA_F32;
  int coord_dtype = MSDA_F32;
  // elements between the offset / logit rows of consecutive queries (0 = dense); the gradient
  // buffers have the layout of their inputs
  long long off_stride = 0, log_stride = 0;
  // SCA backward: tensor-core pass for the coarse pyramid levels (coarse_scatter.cu); all three or none
  void* coarse_rec = nullptr;          // records, sca_coarse_workspace_bytes() bytes
  const int32_t* hit_index = nullptr;  // (groups, Nq) per-camera hit lists of batch element 0
  const int32_t* hit_count = nullptr;  // (groups,)
};

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
void count_launch();

int launch_msda_fwd(const Problem& pr, cudaStream_t st);
int launch_msda_bwd(const Problem& pr, cudaStream_t st);
int launch_point_sampling(const float* ref_3d, const float* lidar2img, const double* pc,
                          float img_h, float img_w, int bs, int num_cam, int HW, int D,
                          float* ref_cam, uint8_t* bev_mask, uint32_t* hit_bits, int32_t* hit_index,
                          int32_t* hit_count, cudaStream_t st);
int launch_sca_fwd(const FusedProblem& fp, cudaStream_t st);
int launch_sca_bwd(const FusedProblem& fp, cudaStream_t st);
bool sca_coarse_active(const FusedProblem& fp);
bool coarse_supported(int Dh, int P, int value_dtype);
long long coarse_record_bytes(int bs, int cams, int Nq, int M, int P);
int launch_coarse_scatter(const void* rec, const int32_t* hit_index, const int32_t* hit_count, const void* g_out,
                          void* g_value, const float* acc_scale, int acc_half, const int64_t* shapes,
                          const int64_t* starts, int bs, int cams, int Nq, int Nk, int M, int Dh, int L, int P,
                          int value_dtype, cudaStream_t st);
int launch_hit_lists(const uint32_t* hit_bits, int num_cam, int HW, int32_t* hit_index, int32_t* hit_count,
                     cudaStream_t st);
int launch_tsa_fwd(const FusedProblem& fp, cudaStream_t st);
int launch_tsa_bwd(const FusedProblem& fp, cudaStream_t st);
int rowops_partial_rows();
// Dropout fused into the row kernels (rowops.cu): `key` = (seed, step) uint64 pair on the device, NULL = off.
struct RowDropout {
  const void* key = nullptr;
  void* key_save = nullptr;            // forward: receives the pair used (for the backward)
  uint32_t site = 0;
  float p = 0.f;
};
int launch_ln(bool bwd, const void* x, const void* dy, const void* gamma, const void* beta, void* y,
              float* mean, float* rstd, void* dx, void* dgamma_dbeta, float* partial,
              long long rows, int C, float eps, int dtype, const void* residual, void* sum_out,
              bool dxsum, const RowDropout& rd, void* dx_masked, cudaStream_t st);
int launch_colsum(const void* x, const void* y, void* dx, void* out, float* partial, long long rows, int C,
                  int dtype, int out_dtype, float relu_scale, cudaStream_t st);
int launch_relu_dropout(void* x, void* mask_out, long long n, int dtype, const RowDropout& rd, cudaStream_t st);
int launch_flatten_level(const void* feat, const void* cams, const void* lvl, void* out, int bs, int num_cam,
                         int C, int hw, long long Nk, long long start, int dtype, cudaStream_t st);
int launch_rotate_nearest(const void* prev, void* out, const float* theta, const float* xs, const float* ys,
                          int bs, int H, int W, int C, int dtype, cudaStream_t st);
constexpr int kWgHeaderFloats = 64;      // ticket counter in front of the linear_wgrad scratch
long long wgrad_workspace_floats(int O, int I);
int launch_wgrad(const void* dy, const void* x, void* dW, void* db, float* ws, long long N, int O, int I,
                 int dtype, cudaStream_t st);
int launch_grad_scale(const void* g, long long n, int dtype, float limit, float* ws, void* zero, long long zero_bytes,
                      cudaStream_t st);
int launch_unscale_cast(const void* acc16, void* out, const float* scale, long long n, int out_dtype,
                        const void* tail, int copies, long long map_elems, long long tail_elems,
                        float* ws, void* colsum_out, int C, int* overflow, long long out_ld, cudaStream_t st);

}  // namespace msda
