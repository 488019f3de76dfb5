// DCNv3 (InternImage's deformable convolution v3) on the sm_100a gather / scatter kernels of the MSDA
// operator.
//
// Reference: projects/mmdet3d_plugin/bevformer/backbones/ops_dcnv3/src/cuda/dcnv3_cuda.cu:28-173
// (dcnv3_cuda_forward / dcnv3_cuda_backward) and dcnv3_im2col_cuda.cuh:216-275 (forward kernel), :278-839
// (the col2im family).  DCNv3 is multi-scale deformable attention with ONE level whose sampling positions
// are stated in pixels of the input map: output pixel (ho, wo), group g, kernel point k = i * kernel_h + j
// samples at
//     w = p0_w - c_w * s + (i * dilation_w + offset_w) * s,     p0_w = c_w - pad_w + wo * stride_w,
//     h = p0_h - c_h * s + (j * dilation_h + offset_h) * s,     c = (dilation * (kernel - 1)) >> 1, s = offset_scale
// bilinearly with zero padding (a sample takes part when -1 < h < H_in and -1 < w < W_in), weighted by
// `mask`.  So the op is: one small kernel that turns (offset, geometry) into pixel positions, then
// msda_fwd / msda_bwd in their pixel-coordinate mode (value = input viewed as (N, H_in * W_in, group,
// group_channels), P = kernel_h * kernel_w, attention weights = mask).  The position kernel replaces the
// reference-era design's per-thread recomputation (one thread per output CHANNEL recomputes every position;
// here a position is computed once and shared by the 16-byte channel lanes of its group), and the backward
// inherits the warp-shuffle reductions and 16-byte vector reductions of msda_bwd instead of the
// shared-memory reduce-then-atomicAdd variants selected by channel count (dcnv3_im2col_cuda.cuh:106-146).
#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {
namespace {

constexpr int kDcnHeaderFloats = 8;       // int64 (H_in, W_in) + int64 start, padded to 32 bytes

struct DcnGeom {
  int N, H_in, W_in, H_out, W_out, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w;
  int group, group_channels;
  float offset_scale;
};

template <typename T>
__global__ void __launch_bounds__(256)
dcnv3_positions_kernel(const T* __restrict__ offset, const T* __restrict__ mask, float* __restrict__ scratch,
                       float* __restrict__ loc, float* __restrict__ mask32, const DcnGeom g, long long total) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int64_t* tab = reinterpret_cast<int64_t*>(scratch);
    tab[0] = g.H_in;
    tab[1] = g.W_in;
    tab[2] = 0;
  }
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int K = g.kernel_h * g.kernel_w;
  const int k = (int)(idx % K);
  long long t = idx / K / g.group;
  const int wo = (int)(t % g.W_out);
  t /= g.W_out;
  const int ho = (int)(t % g.H_out);
  const int i = k / g.kernel_h, j = k % g.kernel_h;           // the reference loops kernel_w outside kernel_h
  const int cw = (g.dilation_w * (g.kernel_w - 1)) >> 1, ch = (g.dilation_h * (g.kernel_h - 1)) >> 1;
  const float p0w = (float)(cw - g.pad_w + wo * g.stride_w), p0h = (float)(ch - g.pad_h + ho * g.stride_h);
  const float p0w_ = fmaf(-(float)cw, g.offset_scale, p0w), p0h_ = fmaf(-(float)ch, g.offset_scale, p0h);
  const float ow = to_f32<T>(offset[2 * idx]), oh = to_f32<T>(offset[2 * idx + 1]);
  loc[2 * idx] = fmaf((float)(i * g.dilation_w) + ow, g.offset_scale, p0w_);
  loc[2 * idx + 1] = fmaf((float)(j * g.dilation_h) + oh, g.offset_scale, p0h_);
  if (mask32 != nullptr) mask32[idx] = to_f32<T>(mask[idx]);
}

int check_geom(const char* what, const DcnGeom& g, int dtype) {
  if (g.N <= 0 || g.H_in <= 0 || g.W_in <= 0 || g.H_out <= 0 || g.W_out <= 0 || g.kernel_h <= 0 || g.kernel_w <= 0 ||
      g.stride_h <= 0 || g.stride_w <= 0 || g.pad_h < 0 || g.pad_w < 0 || g.dilation_h <= 0 || g.dilation_w <= 0 ||
      g.group <= 0 || g.group_channels <= 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: invalid sizes", what);
  if (dtype != MSDA_F32 && dtype != MSDA_F16 && dtype != MSDA_BF16)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: unknown dtype %d", what, dtype);
  const int ho = (g.H_in + 2 * g.pad_h - (g.dilation_h * (g.kernel_h - 1) + 1)) / g.stride_h + 1;
  const int wo = (g.W_in + 2 * g.pad_w - (g.dilation_w * (g.kernel_w - 1) + 1)) / g.stride_w + 1;
  if (ho != g.H_out || wo != g.W_out)               // dcnv3_cuda.cu:46-51
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: output size %d x %d does not follow from the geometry (%d x %d)", what,
                     g.H_out, g.W_out, ho, wo);
  if (!(g.offset_scale > 0.f))
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: offset_scale must be positive", what);
  if ((long long)g.H_in * g.W_in > 0x7fffffffLL || (long long)g.H_out * g.W_out > 0x7fffffffLL)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: map too large", what);
  return MSDA_OK;
}

long long samples(const DcnGeom& g) {
  return (long long)g.N * g.H_out * g.W_out * g.group * g.kernel_h * g.kernel_w;
}

// positions (+ fp32 copy of the mask for 16-bit inputs) into the scratch; fills the msda Problem
int prepare(const char* what, const DcnGeom& g, int dtype, const void* input, const void* offset, const void* mask,
            float* scratch, Problem& pr, cudaStream_t st) {
  if (!input || !offset || !mask || !scratch) return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: NULL pointer", what);
  const long long n = samples(g);
  float* loc = scratch + kDcnHeaderFloats;
  float* mask32 = dtype == MSDA_F32 ? nullptr : loc + 2 * n;
  const long long grid = (n + 255) / 256;
  if (grid > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "%s: too many samples", what);
  if (dtype == MSDA_F32)
    dcnv3_positions_kernel<float><<<(unsigned)grid, 256, 0, st>>>(static_cast<const float*>(offset),
                                                                   static_cast<const float*>(mask), scratch, loc, mask32, g, n);
  else if (dtype == MSDA_BF16)
    dcnv3_positions_kernel<__nv_bfloat16><<<(unsigned)grid, 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(offset), static_cast<const __nv_bfloat16*>(mask), scratch, loc, mask32, g, n);
  else
    dcnv3_positions_kernel<__half><<<(unsigned)grid, 256, 0, st>>>(static_cast<const __half*>(offset),
                                                                    static_cast<const __half*>(mask), scratch, loc, mask32, g, n);
  count_launch();
  if (int rc = check_launch(what)) return rc;
  pr.value = input;
  pr.shapes = reinterpret_cast<const int64_t*>(scratch);
  pr.starts = reinterpret_cast<const int64_t*>(scratch) + 2;
  pr.loc = loc;
  pr.attn = dtype == MSDA_F32 ? mask : static_cast<const void*>(mask32);
  pr.B = g.N; pr.Nk = g.H_in * g.W_in; pr.M = g.group; pr.Dh = g.group_channels; pr.L = 1;
  pr.Nq = g.H_out * g.W_out; pr.P = g.kernel_h * g.kernel_w;
  pr.value_dtype = dtype; pr.coord_dtype = MSDA_F32;
  pr.pixel_scale = g.offset_scale;
  return MSDA_OK;
}

}  // namespace
}  // namespace msda

using namespace msda;

extern "C" {

int64_t dcnv3_scratch_floats(int N, int H_out, int W_out, int group, int kernel_h, int kernel_w, int dtype) {
  if (N <= 0 || H_out <= 0 || W_out <= 0 || group <= 0 || kernel_h <= 0 || kernel_w <= 0) return -1;
  const long long n = (long long)N * H_out * W_out * group * kernel_h * kernel_w;
  return kDcnHeaderFloats + 2 * n + (dtype == MSDA_F32 ? 0 : n);
}

int dcnv3_fwd(const void* input, const void* offset, const void* mask, void* output, float* scratch, int N, int H_in,
              int W_in, int H_out, int W_out, int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h,
              int pad_w, int dilation_h, int dilation_w, int group, int group_channels, float offset_scale, int dtype,
              void* stream) {
  const DcnGeom g{N, H_in, W_in, H_out, W_out, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                  dilation_w, group, group_channels, offset_scale};
  if (int rc = check_geom("dcnv3_fwd", g, dtype)) return rc;
  if (!output) return set_error(MSDA_ERR_BAD_ARGUMENT, "dcnv3_fwd: NULL output");
  Problem pr;
  if (int rc = prepare("dcnv3_fwd", g, dtype, input, offset, mask, scratch, pr, static_cast<cudaStream_t>(stream)))
    return rc;
  pr.out = output;
  return launch_msda_fwd(pr, static_cast<cudaStream_t>(stream));
}

int dcnv3_bwd(const void* input, const void* offset, const void* mask, const void* grad_output, float* grad_input,
              float* grad_offset, float* grad_mask, float* scratch, int N, int H_in, int W_in, int H_out, int W_out,
              int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w, int dilation_h,
              int dilation_w, int group, int group_channels, float offset_scale, int dtype, void* stream) {
  const DcnGeom g{N, H_in, W_in, H_out, W_out, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                  dilation_w, group, group_channels, offset_scale};
  if (int rc = check_geom("dcnv3_bwd", g, dtype)) return rc;
  if (!grad_output || !grad_input || !grad_offset || !grad_mask)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "dcnv3_bwd: NULL pointer");
  Problem pr;
  if (int rc = prepare("dcnv3_bwd", g, dtype, input, offset, mask, scratch, pr, static_cast<cudaStream_t>(stream)))
    return rc;
  pr.grad_out = grad_output;
  pr.g_value = grad_input;
  pr.g_loc = grad_offset;
  pr.g_attn = grad_mask;
  return launch_msda_bwd(pr, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
