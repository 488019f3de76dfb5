// Row-wise companions of the attention kernels in a BEVFormer layer (SURVEY.md section 8f rank 3:
// "encoder remainder"): LayerNorm forward / backward and the column sums that give the bias
// gradients of the Linear layers.  All three are pure HBM streams over (rows, C) activations
// (40 000 x 256 at the base config); they exist because the stock PyTorch kernels for exactly
// these shapes (bf16, many short rows) were the largest non-attention items of the measured step
// (profiles/r01_launches.md).
//
// LayerNorm follows torch.nn.LayerNorm (the reference builds its norms through mmcv's
// build_norm_layer(dict(type='LN')), custom_base_transformer_layer.py:156-161): biased variance,
// eps inside the square root, affine weight / bias, statistics in fp32.
#include "msda_common.cuh"
#include "msda_host.h"

namespace msda {

constexpr int kRowThreads = 256;

template <typename T> struct Vec16IO {
  static constexpr int N = Vec16<T>::N;
  __device__ __forceinline__ static void load(const T* p, float (&f)[N]) { Vec16<T>::unpack(ldg128(p), f); }
  __device__ __forceinline__ static void store(T* p, const float (&f)[N]) {
    *reinterpret_cast<uint4*>(p) = Vec16<T>::pack(f);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per row; every lane keeps its slice of the row (C / 32 elements, <= 32) in registers.
template <typename T, int PER_LANE>
__global__ void __launch_bounds__(kRowThreads)
ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma, const T* __restrict__ beta,
              T* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd,
              long long rows, int C, float eps) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int NV = PER_LANE / VEC;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  const long long nwarp = (long long)gridDim.x * (kRowThreads / 32);
  float g[PER_LANE], bb[PER_LANE];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float t[VEC];
    Vec16IO<T>::load(gamma + (v * 32 + lane) * VEC, t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[v * VEC + i] = t[i];
    Vec16IO<T>::load(beta + (v * 32 + lane) * VEC, t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) bb[v * VEC + i] = t[i];
  }
  for (long long r = warp; r < rows; r += nwarp) {
    const T* xr = x + r * C;
    float f[PER_LANE];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float t[VEC];
      Vec16IO<T>::load(xr + (v * 32 + lane) * VEC, t);
#pragma unroll
      for (int i = 0; i < VEC; ++i) { f[v * VEC + i] = t[i]; s += t[i]; }
    }
    const float mu = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) { const float d = f[i] - mu; q = fmaf(d, d, q); }
    const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
    T* yr = y + r * C;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float t[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) t[i] = (f[v * VEC + i] - mu) * rs * g[v * VEC + i] + bb[v * VEC + i];
      Vec16IO<T>::store(yr + (v * 32 + lane) * VEC, t);
    }
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

// dx per row; per-CTA partial sums of dgamma / dbeta into `partial` (gridDim.x, 2, C) fp32.
template <typename T, int PER_LANE>
__global__ void __launch_bounds__(kRowThreads)
ln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dx,
              float* __restrict__ partial, long long rows, int C) {
  constexpr int VEC = Vec16<T>::N;
  constexpr int NV = PER_LANE / VEC;
  extern __shared__ float sm[];                       // [warps][2][C]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * (kRowThreads / 32) + wid;
  const long long nwarp = (long long)gridDim.x * (kRowThreads / 32);
  float g[PER_LANE], dg[PER_LANE], db[PER_LANE];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float t[VEC];
    Vec16IO<T>::load(gamma + (v * 32 + lane) * VEC, t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) g[v * VEC + i] = t[i];
  }
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) { dg[i] = 0.f; db[i] = 0.f; }
  for (long long r = warp; r < rows; r += nwarp) {
    const float mu = mean[r], rs = rstd[r];
    float xh[PER_LANE], gd[PER_LANE];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float tx[VEC], td[VEC];
      Vec16IO<T>::load(x + r * C + (v * 32 + lane) * VEC, tx);
      Vec16IO<T>::load(dy + r * C + (v * 32 + lane) * VEC, td);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int k = v * VEC + i;
        xh[k] = (tx[i] - mu) * rs;
        gd[k] = td[i] * g[k];
        dg[k] = fmaf(td[i], xh[k], dg[k]);
        db[k] += td[i];
        s1 += gd[k];
        s2 = fmaf(gd[k], xh[k], s2);
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float t[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int k = v * VEC + i;
        t[i] = rs * (gd[k] - s1 - xh[k] * s2);
      }
      Vec16IO<T>::store(dx + r * C + (v * 32 + lane) * VEC, t);
    }
  }
  // CTA reduction of the per-warp partials, then one row of partials per CTA
  float* mine = sm + (size_t)wid * 2 * C;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = (v * 32 + lane) * VEC + i;
      mine[c] = dg[v * VEC + i];
      mine[C + c] = db[v * VEC + i];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += kRowThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kRowThreads / 32; ++w) s += sm[(size_t)w * 2 * C + c];
    partial[(size_t)blockIdx.x * 2 * C + c] = s;
  }
}

// out[c] = sum over `n` rows of partial[r][c]  (final stage of the two column reductions)
template <typename TO>
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, TO* __restrict__ out, int n, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < n; ++r) s += partial[(size_t)r * C + c];
  out[c] = from_f32<TO>(s);
}

// Column sums of a (rows, C) matrix: stage 1, per-CTA partials (gridDim.x, C) fp32.  A thread
// owns 16 bytes of columns; a CTA's threads cover C/VEC column groups x (256 / (C/VEC)) row lanes.
template <typename T>
__global__ void __launch_bounds__(kRowThreads)
colsum_partial_kernel(const T* __restrict__ x, float* __restrict__ partial, long long rows, int C) {
  constexpr int VEC = Vec16<T>::N;
  extern __shared__ float sm[];                       // [row_lanes][C]
  const int groups = C / VEC;                         // <= 256
  const int row_lanes = kRowThreads / groups;
  const int gidx = threadIdx.x % groups, rl = threadIdx.x / groups;
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  if (rl < row_lanes) {
    for (long long r = (long long)blockIdx.x * row_lanes + rl; r < rows; r += (long long)gridDim.x * row_lanes) {
      float t[VEC];
      Vec16IO<T>::load(x + r * C + gidx * VEC, t);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] += t[i];
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) sm[(size_t)rl * C + gidx * VEC + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kRowThreads) {
    float s = 0.f;
    for (int w = 0; w < row_lanes; ++w) s += sm[(size_t)w * C + c];
    partial[(size_t)blockIdx.x * C + c] = s;
  }
}

static int row_grid() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n * 4;
}

int rowops_partial_rows() { return row_grid(); }

template <typename T, int PER_LANE>
static int ln_launch(bool bwd, const void* x, const void* dy, const void* gamma, const void* beta, void* y,
                     float* mean, float* rstd, void* dx, float* partial, long long rows, int C, float eps,
                     cudaStream_t st) {
  const long long need = (rows + kRowThreads / 32 - 1) / (kRowThreads / 32);
  const int grid = (int)(need < row_grid() ? need : row_grid());
  if (grid <= 0) return MSDA_OK;
  if (!bwd) {
    ln_fwd_kernel<T, PER_LANE><<<grid, kRowThreads, 0, st>>>(
        static_cast<const T*>(x), static_cast<const T*>(gamma), static_cast<const T*>(beta),
        static_cast<T*>(y), mean, rstd, rows, C, eps);
    count_launch();
    return check_launch("ln_fwd");
  }
  const size_t smem = (size_t)(kRowThreads / 32) * 2 * C * sizeof(float);
  ln_bwd_kernel<T, PER_LANE><<<row_grid(), kRowThreads, smem, st>>>(
      static_cast<const T*>(x), static_cast<const T*>(dy), static_cast<const T*>(gamma), mean, rstd,
      static_cast<T*>(dx), partial, rows, C);
  count_launch();
  return check_launch("ln_bwd");
}

template <typename T>
static int ln_dispatch(bool bwd, const void* x, const void* dy, const void* gamma, const void* beta, void* y,
                       float* mean, float* rstd, void* dx, float* partial, long long rows, int C, float eps,
                       cudaStream_t st) {
  constexpr int VEC = Vec16<T>::N;
  if (C % (32 * VEC) != 0 || C / 32 > 32)
    return set_error(MSDA_ERR_UNSUPPORTED, "layer_norm: C=%d must be a multiple of %d and <= 1024", C, 32 * VEC);
  switch (C / 32) {
    case 4: return ln_launch<T, 4>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, rows, C, eps, st);
    case 8: return ln_launch<T, 8>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, rows, C, eps, st);
    case 16: return ln_launch<T, 16>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, rows, C, eps, st);
    case 32: return ln_launch<T, 32>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, rows, C, eps, st);
    default: break;
  }
  return set_error(MSDA_ERR_UNSUPPORTED, "layer_norm: C=%d not supported (128, 256, 512 or 1024)", C);
}

int launch_ln(bool bwd, const void* x, const void* dy, const void* gamma, const void* beta, void* y,
              float* mean, float* rstd, void* dx, void* dgamma_dbeta, float* partial,
              long long rows, int C, float eps, int dtype, cudaStream_t st) {
  int rc;
  if (dtype == MSDA_F32) rc = ln_dispatch<float>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, rows, C, eps, st);
  else if (dtype == MSDA_BF16) rc = ln_dispatch<__nv_bfloat16>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, rows, C, eps, st);
  else rc = ln_dispatch<__half>(bwd, x, dy, gamma, beta, y, mean, rstd, dx, partial, rows, C, eps, st);
  if (rc || !bwd) return rc;
  // final reduction of the per-CTA partial rows [dgamma | dbeta] into the (2, C) output strip
  const int n = row_grid();
  const int blocks = (2 * C + 255) / 256;
  if (dtype == MSDA_F32)
    colsum_final_kernel<float><<<blocks, 256, 0, st>>>(partial, static_cast<float*>(dgamma_dbeta), n, 2 * C);
  else if (dtype == MSDA_BF16)
    colsum_final_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(partial, static_cast<__nv_bfloat16*>(dgamma_dbeta), n, 2 * C);
  else
    colsum_final_kernel<__half><<<blocks, 256, 0, st>>>(partial, static_cast<__half*>(dgamma_dbeta), n, 2 * C);
  count_launch();
  return check_launch("ln_bwd(final)");
}

int launch_colsum(const void* x, void* out, float* partial, long long rows, int C, int dtype, int out_dtype,
                  cudaStream_t st) {
  const int vec = dtype == MSDA_F32 ? 4 : 8;
  if (C % vec != 0 || C / vec > kRowThreads)
    return set_error(MSDA_ERR_UNSUPPORTED, "colsum: C=%d must be a multiple of %d and <= %d", C, vec, vec * kRowThreads);
  const int groups = C / vec, row_lanes = kRowThreads / groups;
  const long long need = (rows + row_lanes - 1) / row_lanes;
  const int grid = (int)(need < row_grid() ? (need > 0 ? need : 1) : row_grid());
  const size_t smem = (size_t)row_lanes * C * sizeof(float);
  if (dtype == MSDA_F32) colsum_partial_kernel<float><<<grid, kRowThreads, smem, st>>>(static_cast<const float*>(x), partial, rows, C);
  else if (dtype == MSDA_BF16) colsum_partial_kernel<__nv_bfloat16><<<grid, kRowThreads, smem, st>>>(static_cast<const __nv_bfloat16*>(x), partial, rows, C);
  else colsum_partial_kernel<__half><<<grid, kRowThreads, smem, st>>>(static_cast<const __half*>(x), partial, rows, C);
  count_launch();
  if (int rc = check_launch("colsum")) return rc;
  if (out_dtype == MSDA_F32) colsum_final_kernel<float><<<(C + 255) / 256, 256, 0, st>>>(partial, static_cast<float*>(out), grid, C);
  else if (out_dtype == MSDA_BF16) colsum_final_kernel<__nv_bfloat16><<<(C + 255) / 256, 256, 0, st>>>(partial, static_cast<__nv_bfloat16*>(out), grid, C);
  else colsum_final_kernel<__half><<<(C + 255) / 256, 256, 0, st>>>(partial, static_cast<__half*>(out), grid, C);
  count_launch();
  return check_launch("colsum(final)");
}

}  // namespace msda
