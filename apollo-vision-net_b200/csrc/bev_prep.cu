// Pre-processing in front of the BEV encoder (reference
// PerceptionTransformer.get_bev_features, transformer.py:119-298), without the reference's host
// synchronisations (its `isfinite().all()` guards) and per-sample Python loops:
//
//  * bev_flatten_level: one pyramid level of the multi-camera image features
//    (bs, num_cam, C, h*w) -> rows [start, start + h*w) of feat_flatten (num_cam, Nk, bs, C), with
//    non-finite values zeroed (transformer.py:246-247, 282-283 without the device->host check) and
//    the camera and level embeddings added (:249-253).  A tiled shared-memory transpose: both
//    sides are accessed in full 128-byte lines.
//  * bev_rotate_nearest: the ego-motion rotation of the previous BEV (:182-203), i.e.
//    torchvision's rotate(nearest, no expand, zero fill) for all samples in one launch.  The
//    source coordinate follows torchvision's fp32 arithmetic step by step (affine grid from the
//    pixel-centre linspace, theta pre-divided by the half sizes, grid_sample's un-normalisation,
//    round-half-to-even), so the picked pixels are the reference's.
#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {

template <typename T> __device__ __forceinline__ float sanitize(T v) {
  const float f = to_f32<T>(v);
  return isfinite(f) ? f : 0.f;
}

// grid: (ceil(hw / 32), ceil(C / 32), bs * num_cam); block (32, 8)
template <typename T>
__global__ void __launch_bounds__(256)
flatten_level_kernel(const T* __restrict__ feat, const T* __restrict__ cams_embeds,
                     const T* __restrict__ level_embed, T* __restrict__ out, int bs, int num_cam, int C,
                     int hw, long long Nk, long long start) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z / num_cam, cam = blockIdx.z % num_cam;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const T* src = feat + ((size_t)b * num_cam + cam) * C * hw;
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {                   // rows = channels, columns = pixels
    const int c = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < hw) ? sanitize<T>(src[(size_t)c * hw + p]) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {                   // rows = pixels, columns = channels
    const int p = p0 + j, c = c0 + threadIdx.x;
    if (p < hw && c < C) {
      float v = tile[threadIdx.x][j];
      // the reference adds the camera embedding, rounds, then adds the level embedding
      if (cams_embeds != nullptr) v = to_f32<T>(from_f32<T>(v + to_f32<T>(cams_embeds[(size_t)cam * C + c])));
      v = v + to_f32<T>(level_embed[c]);
      out[(((size_t)cam * Nk + start + p) * bs + b) * C + c] = from_f32<T>(v);
    }
  }
}

// One thread per (destination pixel, sample, 16-byte channel chunk).
template <typename T>
__global__ void __launch_bounds__(256)
rotate_nearest_kernel(const T* __restrict__ prev, T* __restrict__ out, const float* __restrict__ theta,
                      const float* __restrict__ xs, const float* __restrict__ ys, int bs, int H, int W,
                      int C) {
  constexpr int VEC = Vec16<T>::N;
  const int chunks = C / VEC;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)H * W * bs * chunks;
  if (idx >= total) return;
  const int ch = (int)(idx % chunks);
  const int b = (int)((idx / chunks) % bs);
  const int q = (int)(idx / ((long long)chunks * bs));
  const int y = q / W, x = q % W;
  const float* t = theta + b * 6;            // rescaled: [a/(W/2), b/(W/2), c/(W/2), d/(H/2), e/(H/2), f/(H/2)]
  const float xb = xs[x], yb = ys[y];
  // base_grid (x, y, 1) x rescaled_theta: three-term dot product, left to right, no contraction
  const float gx = __fadd_rn(__fadd_rn(__fmul_rn(xb, t[0]), __fmul_rn(yb, t[1])), t[2]);
  const float gy = __fadd_rn(__fadd_rn(__fmul_rn(xb, t[3]), __fmul_rn(yb, t[4])), t[5]);
  // grid_sample, align_corners = False: ((g + 1) * size - 1) / 2, nearest = round half to even
  const float ix = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), -1.f), 2.f);
  const float iy = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), -1.f), 2.f);
  const float rx = nearbyintf(ix), ry = nearbyintf(iy);
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (rx >= 0.f && rx <= (float)(W - 1) && ry >= 0.f && ry <= (float)(H - 1)) {
    const long long sq = (long long)ry * W + (long long)rx;
    v = ldg128(prev + ((size_t)sq * bs + b) * C + ch * VEC);
  }
  *reinterpret_cast<uint4*>(out + ((size_t)q * bs + b) * C + ch * VEC) = v;
}

template <typename T>
static int flatten_t(const void* feat, const void* cams, const void* lvl, void* out, int bs, int num_cam, int C,
                     int hw, long long Nk, long long start, cudaStream_t st) {
  const dim3 grid((hw + 31) / 32, (C + 31) / 32, bs * num_cam), block(32, 8);
  flatten_level_kernel<T><<<grid, block, 0, st>>>(static_cast<const T*>(feat), static_cast<const T*>(cams),
                                                  static_cast<const T*>(lvl), static_cast<T*>(out), bs,
                                                  num_cam, C, hw, Nk, start);
  count_launch();
  return check_launch("bev_flatten_level");
}

int launch_flatten_level(const void* feat, const void* cams, const void* lvl, void* out, int bs, int num_cam,
                         int C, int hw, long long Nk, long long start, int dtype, cudaStream_t st) {
  if ((long long)bs * num_cam > 65535)
    return set_error(MSDA_ERR_UNSUPPORTED, "bev_flatten_level: bs * num_cam must be <= 65535");
  if (dtype == MSDA_F32) return flatten_t<float>(feat, cams, lvl, out, bs, num_cam, C, hw, Nk, start, st);
  if (dtype == MSDA_BF16) return flatten_t<__nv_bfloat16>(feat, cams, lvl, out, bs, num_cam, C, hw, Nk, start, st);
  return flatten_t<__half>(feat, cams, lvl, out, bs, num_cam, C, hw, Nk, start, st);
}

template <typename T>
static int rotate_t(const void* prev, void* out, const float* theta, const float* xs, const float* ys, int bs,
                    int H, int W, int C, cudaStream_t st) {
  constexpr int VEC = Vec16<T>::N;
  if (C % VEC != 0)
    return set_error(MSDA_ERR_UNSUPPORTED, "bev_rotate_nearest: C=%d must be a multiple of %d", C, VEC);
  const long long total = (long long)H * W * bs * (C / VEC);
  const long long grid = (total + 255) / 256;
  if (grid > 0x7fffffffLL) return set_error(MSDA_ERR_UNSUPPORTED, "bev_rotate_nearest: too many elements");
  rotate_nearest_kernel<T><<<(unsigned)grid, 256, 0, st>>>(static_cast<const T*>(prev), static_cast<T*>(out),
                                                            theta, xs, ys, bs, H, W, C);
  count_launch();
  return check_launch("bev_rotate_nearest");
}

int launch_rotate_nearest(const void* prev, void* out, const float* theta, const float* xs, const float* ys,
                          int bs, int H, int W, int C, int dtype, cudaStream_t st) {
  if (dtype == MSDA_F32) return rotate_t<float>(prev, out, theta, xs, ys, bs, H, W, C, st);
  if (dtype == MSDA_BF16) return rotate_t<__nv_bfloat16>(prev, out, theta, xs, ys, bs, H, W, C, st);
  return rotate_t<__half>(prev, out, theta, xs, ys, bs, H, W, C, st);
}

}  // namespace msda
