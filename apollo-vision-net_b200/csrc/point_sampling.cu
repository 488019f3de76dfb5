// BEV geometry for spatial cross-attention on the device, with no host synchronisation.
//
// Replaces BEVFormerEncoder.point_sampling (encoder.py:147-239) and the per-camera
// `mask.sum(-1).nonzero()` compaction + `max(len(...))` host sync of
// SpatialCrossAttention.forward (spatial_cross_attention.py:135-139).
//
// bev_mask must be bit-exact w.r.t. the reference's CPU path, so the arithmetic is spelled
// out with round-to-nearest intrinsics in the order torch's fp32 CPU kernels use (verified in
// the build container, see DESIGN.md "bit-exact mask"): separate multiply and add for the
// pc_range scaling, a left-to-right un-fused 4-term dot product for the lidar2img matmul,
// IEEE division for the perspective divide and for the image-size normalisation.
#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {

__global__ void __launch_bounds__(256)
point_sampling_kernel(const float* __restrict__ ref_3d, const float* __restrict__ lidar2img,
                      float sx, float sy, float sz, float ox, float oy, float oz,
                      float img_h, float img_w, int bs, int num_cam, int HW, int D,
                      float* __restrict__ ref_cam, uint8_t* __restrict__ bev_mask,
                      uint32_t* __restrict__ hit_bits) {
  extern __shared__ float s_l2i[];                       // (bs*num_cam, 12): rows 0..2 of each matrix
  const int nmat = bs * num_cam;
  for (int i = threadIdx.x; i < nmat * 12; i += blockDim.x)
    s_l2i[i] = lidar2img[(i / 12) * 16 + (i % 12)];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)bs * HW) return;
  const int b = (int)(idx / HW);
  const int q = (int)(idx % HW);
  const float eps = 1e-5f;
  uint32_t bits = 0;
  for (int d = 0; d < D; ++d) {
    const float* r = ref_3d + (((size_t)b * D + d) * HW + q) * 3;
    const float X = __fadd_rn(__fmul_rn(r[0], sx), ox);
    const float Y = __fadd_rn(__fmul_rn(r[1], sy), oy);
    const float Z = __fadd_rn(__fmul_rn(r[2], sz), oz);
    for (int c = 0; c < num_cam; ++c) {
      const float* a = s_l2i + (b * num_cam + c) * 12;
      float cam[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float acc = __fmul_rn(a[4 * i], X);
        acc = __fadd_rn(acc, __fmul_rn(a[4 * i + 1], Y));
        acc = __fadd_rn(acc, __fmul_rn(a[4 * i + 2], Z));
        acc = __fadd_rn(acc, a[4 * i + 3]);              // homogeneous coordinate is exactly 1
        cam[i] = acc;
      }
      bool ok = cam[2] > eps;
      const float zc = fmaxf(cam[2], eps);
      const float u = __fdiv_rn(__fdiv_rn(cam[0], zc), img_w);
      const float v = __fdiv_rn(__fdiv_rn(cam[1], zc), img_h);
      ok = ok && (v > 0.0f) && (v < 1.0f) && (u < 1.0f) && (u > 0.0f);
      const size_t o = (((size_t)c * bs + b) * HW + q) * D + d;
      ref_cam[2 * o] = u;
      ref_cam[2 * o + 1] = v;
      bev_mask[o] = ok ? 1 : 0;
      if (ok) bits |= (1u << c);
    }
  }
  hit_bits[idx] = bits;
}

// One CTA per camera: ordered compaction of the batch-0 hit list (ascending query index, what
// nonzero() returns), via ballot + block prefix sums.
__global__ void __launch_bounds__(1024)
hit_compaction_kernel(const uint32_t* __restrict__ hit_bits, int HW, int32_t* __restrict__ hit_index,
                      int32_t* __restrict__ hit_count) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int cam = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int32_t* out = hit_index + (size_t)cam * HW;
  if (tid == 0) base = 0;
  __syncthreads();
  for (int q0 = 0; q0 < HW; q0 += 1024) {
    const int q = q0 + tid;
    const bool hit = (q < HW) && ((hit_bits[q] >> cam) & 1u);
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_tot[w];
    const int b0 = base;
    if (hit) out[b0 + woff + before] = q;
    __syncthreads();
    if (tid == 1023) base = b0 + woff + __popc(bal);
    __syncthreads();
  }
  const int n = base;
  for (int i = n + tid; i < HW; i += 1024) out[i] = -1;
  if (tid == 0) hit_count[cam] = n;
}

int launch_point_sampling(const float* ref_3d, const float* lidar2img, const double* pc,
                          float img_h, float img_w, int bs, int num_cam, int HW, int D,
                          float* ref_cam, uint8_t* bev_mask, uint32_t* hit_bits, int32_t* hit_index,
                          int32_t* hit_count, cudaStream_t st) {
  const float sx = (float)(pc[3] - pc[0]), sy = (float)(pc[4] - pc[1]), sz = (float)(pc[5] - pc[2]);
  const long long n = (long long)bs * HW;
  const size_t smem = (size_t)bs * num_cam * 12 * sizeof(float);
  if (smem > 40 * 1024) return set_error(MSDA_ERR_UNSUPPORTED, "bev_point_sampling: bs*num_cam too large");
  point_sampling_kernel<<<(unsigned)((n + 255) / 256), 256, smem, st>>>(
      ref_3d, lidar2img, sx, sy, sz, (float)pc[0], (float)pc[1], (float)pc[2], img_h, img_w, bs,
      num_cam, HW, D, ref_cam, bev_mask, hit_bits);
  count_launch();
  if (int rc = check_launch("bev_point_sampling")) return rc;
  if (!hit_index || !hit_count) return MSDA_OK;      // the fused forward only needs the bit field
  return launch_hit_lists(hit_bits, num_cam, HW, hit_index, hit_count, st);
}

int launch_hit_lists(const uint32_t* hit_bits, int num_cam, int HW, int32_t* hit_index, int32_t* hit_count,
                     cudaStream_t st) {
  hit_compaction_kernel<<<num_cam, 1024, 0, st>>>(hit_bits, HW, hit_index, hit_count);
  count_launch();
  return check_launch("bev_hit_lists");
}

}  // namespace msda
