// extern "C" surface of libmsda_b200.so: argument validation, error reporting, host-buffer
// variants.  See include/msda_b200.h for the contract of every entry point.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "msda_host.h"

namespace msda {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MSDA_ERR_CUDA, "%s: CUDA error: %s", what, cudaGetErrorString(e));
  return MSDA_OK;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static size_t dtype_size(int dt) { return dt == MSDA_F32 ? 4 : 2; }

static int validate(const Problem& p, int im2col_step, bool bwd, const char* fn) {
  if (p.B < 0 || p.Nq < 0 || p.Nk < 0 || p.M <= 0 || p.Dh <= 0 || p.L <= 0 || p.P <= 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: invalid sizes B=%d Nk=%d M=%d Dh=%d L=%d Nq=%d P=%d",
                     fn, p.B, p.Nk, p.M, p.Dh, p.L, p.Nq, p.P);
  if (p.L > MSDA_MAX_LEVELS)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: num_levels %d > %d", fn, p.L, MSDA_MAX_LEVELS);
  if (p.value_dtype != MSDA_F32 && p.value_dtype != MSDA_F16 && p.value_dtype != MSDA_BF16)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: unknown value dtype %d", fn, p.value_dtype);
  if (p.coord_dtype != MSDA_F32 && p.coord_dtype != p.value_dtype)
    return set_error(MSDA_ERR_UNSUPPORTED,
                     "%s: locations/weights must be fp32 or the value dtype (got %d vs %d)", fn,
                     p.coord_dtype, p.value_dtype);
  // mmcv: im2col_step_ = min(batch, im2col_step); AT_ASSERTM(batch % im2col_step_ == 0)
  if (im2col_step <= 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: im2col_step must be positive (got %d)", fn, im2col_step);
  if (p.B > 0) {
    const int step = p.B < im2col_step ? p.B : im2col_step;
    if (p.B % step != 0)
      return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: batch(%d) must divide im2col_step(%d)", fn, p.B, step);
  }
  if ((long long)p.Nk * p.M * p.Dh >= 0x7fffffffLL)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: Nk*M*Dh must be < 2^31", fn);
  const bool empty = (long long)p.B * p.Nq == 0;
  if (!empty) {
    if (!p.value && p.Nk > 0) return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: value is NULL", fn);
    if (!p.shapes || !p.starts || !p.loc || !p.attn)
      return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: NULL input pointer", fn);
    if (!bwd && !p.out) return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: out is NULL", fn);
    if (bwd && (!p.grad_out || !p.g_value || !p.g_loc || !p.g_attn))
      return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: NULL gradient pointer", fn);
  }
  return MSDA_OK;
}

static int validate_fused(const FusedProblem& f, bool bwd, bool sca, const char* fn) {
  if (f.bs < 0 || f.groups <= 0 || f.Nq < 0 || f.Nk < 0 || f.M <= 0 || f.Dh <= 0 || f.L <= 0 ||
      f.P <= 0 || f.D <= 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: invalid sizes", fn);
  if (f.L > MSDA_MAX_LEVELS)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: num_levels %d > %d", fn, f.L, MSDA_MAX_LEVELS);
  if (f.value_dtype != MSDA_F32 && f.value_dtype != MSDA_F16 && f.value_dtype != MSDA_BF16)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: unknown value dtype %d", fn, f.value_dtype);
  if (sca && f.P % f.D != 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: num_points (%d) must be a multiple of the Z anchors (%d)",
                     fn, f.P, f.D);
  if (sca && f.groups > 32)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: at most 32 cameras (got %d)", fn, f.groups);
  if ((long long)f.Nk * f.M * f.Dh >= 0x7fffffffLL)
    return set_error(MSDA_ERR_UNSUPPORTED, "%s: Nk*M*Dh must be < 2^31", fn);
  {
    // per-query row lengths of the offsets / logits tensors; a stride of 0 means dense
    const long long per_q = (long long)f.M * (sca ? 1 : f.groups) * f.L * f.P;
    if ((f.off_stride != 0 && f.off_stride < 2 * per_q) || (f.log_stride != 0 && f.log_stride < per_q))
      return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: offsets / logits stride shorter than a row", fn);
  }
  if ((long long)f.bs * f.Nq == 0) return MSDA_OK;
  if (!f.value || !f.shapes || !f.starts || !f.offsets || !f.logits || !f.ref)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: NULL input pointer", fn);
  if (sca && (!f.bev_mask || !f.hit_bits))
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: NULL mask pointer", fn);
  if (!bwd && !f.out) return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: out is NULL", fn);
  if (bwd && (!f.g_out || !f.g_value || !f.g_offsets || !f.g_logits))
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: NULL gradient pointer", fn);
  if (bwd && f.acc_half && (f.value_dtype == MSDA_F32 || !f.acc_scale))
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: the fp16 accumulator needs a 16-bit value dtype and a scale", fn);
  return MSDA_OK;
}

struct HostLayout {
  size_t value, shapes, starts, loc, attn, out, grad_out, g_value, g_loc, g_attn, total;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static HostLayout host_layout(int B, int Nk, int M, int Dh, int L, int Nq, int P, int vdt, int cdt,
                              bool bwd) {
  HostLayout h{};
  const size_t ev = dtype_size(vdt), ec = dtype_size(cdt);
  const size_t n_val = (size_t)B * Nk * M * Dh, n_s = (size_t)B * Nq * M * L * P;
  const size_t n_out = (size_t)B * Nq * M * Dh;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  h.value = take(n_val * ev);
  h.shapes = take((size_t)L * 2 * 8);
  h.starts = take((size_t)L * 8);
  h.loc = take(n_s * 2 * ec);
  h.attn = take(n_s * ec);
  h.out = take(n_out * ev);
  if (bwd) {
    h.grad_out = take(n_out * ev);
    h.g_value = take(n_val * 4);
    h.g_loc = take(n_s * 2 * 4);
    h.g_attn = take(n_s * 4);
  }
  h.total = off;
  return h;
}

}  // namespace msda

using namespace msda;

#define CUDA_TRY(expr, what)                                                              \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return set_error(MSDA_ERR_CUDA, "%s: CUDA error: %s", what, cudaGetErrorString(e_)); \
  } while (0)

extern "C" {

int msda_abi_version(void) { return MSDA_ABI_VERSION; }
const char* msda_last_error(void) { return g_err; }
int64_t msda_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int msda_fwd(const void* value, const int64_t* shapes, const int64_t* starts, const void* loc,
             const void* attn, void* out, int B, int Nk, int M, int Dh, int L, int Nq, int P,
             int value_dtype, int coord_dtype, int im2col_step, void* stream) {
  Problem p;
  p.value = value; p.shapes = shapes; p.starts = starts; p.loc = loc; p.attn = attn; p.out = out;
  p.B = B; p.Nk = Nk; p.M = M; p.Dh = Dh; p.L = L; p.Nq = Nq; p.P = P;
  p.value_dtype = value_dtype; p.coord_dtype = coord_dtype;
  if (int rc = validate(p, im2col_step, false, "msda_fwd")) return rc;
  if ((long long)B * Nq == 0) return MSDA_OK;
  return launch_msda_fwd(p, static_cast<cudaStream_t>(stream));
}

int msda_bwd(const void* value, const int64_t* shapes, const int64_t* starts, const void* loc,
             const void* attn, const void* grad_out, float* g_value, float* g_loc, float* g_attn,
             int B, int Nk, int M, int Dh, int L, int Nq, int P, int value_dtype, int coord_dtype,
             int im2col_step, void* stream) {
  Problem p;
  p.value = value; p.shapes = shapes; p.starts = starts; p.loc = loc; p.attn = attn;
  p.grad_out = grad_out; p.g_value = g_value; p.g_loc = g_loc; p.g_attn = g_attn;
  p.B = B; p.Nk = Nk; p.M = M; p.Dh = Dh; p.L = L; p.Nq = Nq; p.P = P;
  p.value_dtype = value_dtype; p.coord_dtype = coord_dtype;
  if (int rc = validate(p, im2col_step, true, "msda_bwd")) return rc;
  if ((long long)B * Nq == 0) return MSDA_OK;
  return launch_msda_bwd(p, static_cast<cudaStream_t>(stream));
}

int64_t msda_host_scratch_bytes(int B, int Nk, int M, int Dh, int L, int Nq, int P, int value_dtype,
                                int coord_dtype, int with_backward) {
  if (B < 0 || Nk < 0 || M <= 0 || Dh <= 0 || L <= 0 || Nq < 0 || P <= 0) return -1;
  return (int64_t)host_layout(B, Nk, M, Dh, L, Nq, P, value_dtype, coord_dtype, with_backward != 0).total;
}

static int host_call(const void* value_h, const int64_t* shapes_h, const int64_t* starts_h,
                     const void* loc_h, const void* attn_h, const void* grad_out_h, void* out_h,
                     float* g_value_h, float* g_loc_h, float* g_attn_h, int B, int Nk, int M, int Dh,
                     int L, int Nq, int P, int vdt, int cdt, void* scratch, int64_t scratch_bytes,
                     void* stream, bool bwd) {
  const char* fn = bwd ? "msda_fwd_bwd_host" : "msda_fwd_host";
  if (B < 0 || Nk < 0 || M <= 0 || Dh <= 0 || L <= 0 || Nq < 0 || P <= 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: invalid sizes", fn);
  const HostLayout h = host_layout(B, Nk, M, Dh, L, Nq, P, vdt, cdt, bwd);
  if (!scratch || scratch_bytes < (int64_t)h.total)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "%s: scratch too small (%lld < %lld bytes)", fn,
                     (long long)scratch_bytes, (long long)h.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* d = static_cast<char*>(scratch);
  const size_t ev = dtype_size(vdt), ec = dtype_size(cdt);
  const size_t n_val = (size_t)B * Nk * M * Dh, n_s = (size_t)B * Nq * M * L * P;
  const size_t n_out = (size_t)B * Nq * M * Dh;
  CUDA_TRY(cudaMemcpyAsync(d + h.value, value_h, n_val * ev, cudaMemcpyHostToDevice, st), fn);
  CUDA_TRY(cudaMemcpyAsync(d + h.shapes, shapes_h, (size_t)L * 16, cudaMemcpyHostToDevice, st), fn);
  CUDA_TRY(cudaMemcpyAsync(d + h.starts, starts_h, (size_t)L * 8, cudaMemcpyHostToDevice, st), fn);
  CUDA_TRY(cudaMemcpyAsync(d + h.loc, loc_h, n_s * 2 * ec, cudaMemcpyHostToDevice, st), fn);
  CUDA_TRY(cudaMemcpyAsync(d + h.attn, attn_h, n_s * ec, cudaMemcpyHostToDevice, st), fn);
  int rc = msda_fwd(d + h.value, (const int64_t*)(d + h.shapes), (const int64_t*)(d + h.starts),
                    d + h.loc, d + h.attn, d + h.out, B, Nk, M, Dh, L, Nq, P, vdt, cdt, B > 0 ? B : 1, st);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(out_h, d + h.out, n_out * ev, cudaMemcpyDeviceToHost, st), fn);
  if (bwd) {
    CUDA_TRY(cudaMemcpyAsync(d + h.grad_out, grad_out_h, n_out * ev, cudaMemcpyHostToDevice, st), fn);
    CUDA_TRY(cudaMemsetAsync(d + h.g_value, 0, n_val * 4, st), fn);
    rc = msda_bwd(d + h.value, (const int64_t*)(d + h.shapes), (const int64_t*)(d + h.starts),
                  d + h.loc, d + h.attn, d + h.grad_out, (float*)(d + h.g_value), (float*)(d + h.g_loc),
                  (float*)(d + h.g_attn), B, Nk, M, Dh, L, Nq, P, vdt, cdt, B > 0 ? B : 1, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(g_value_h, d + h.g_value, n_val * 4, cudaMemcpyDeviceToHost, st), fn);
    CUDA_TRY(cudaMemcpyAsync(g_loc_h, d + h.g_loc, n_s * 8, cudaMemcpyDeviceToHost, st), fn);
    CUDA_TRY(cudaMemcpyAsync(g_attn_h, d + h.g_attn, n_s * 4, cudaMemcpyDeviceToHost, st), fn);
  }
  CUDA_TRY(cudaStreamSynchronize(st), fn);
  return MSDA_OK;
}

int msda_fwd_host(const void* value_host, const int64_t* shapes_host, const int64_t* starts_host,
                  const void* loc_host, const void* attn_host, void* out_host, int B, int Nk, int M,
                  int Dh, int L, int Nq, int P, int value_dtype, int coord_dtype, void* scratch,
                  int64_t scratch_bytes, void* stream) {
  return host_call(value_host, shapes_host, starts_host, loc_host, attn_host, nullptr, out_host,
                   nullptr, nullptr, nullptr, B, Nk, M, Dh, L, Nq, P, value_dtype, coord_dtype,
                   scratch, scratch_bytes, stream, false);
}

int msda_fwd_bwd_host(const void* value_host, const int64_t* shapes_host, const int64_t* starts_host,
                      const void* loc_host, const void* attn_host, const void* grad_out_host,
                      void* out_host, float* g_value_host, float* g_loc_host, float* g_attn_host,
                      int B, int Nk, int M, int Dh, int L, int Nq, int P, int value_dtype,
                      int coord_dtype, void* scratch, int64_t scratch_bytes, void* stream) {
  return host_call(value_host, shapes_host, starts_host, loc_host, attn_host, grad_out_host, out_host,
                   g_value_host, g_loc_host, g_attn_host, B, Nk, M, Dh, L, Nq, P, value_dtype,
                   coord_dtype, scratch, scratch_bytes, stream, true);
}

int bev_point_sampling(const float* ref_3d, const float* lidar2img, const double* pc_range_host,
                       float img_h, float img_w, int bs, int num_cam, int HW, int D, float* ref_cam,
                       uint8_t* bev_mask, uint32_t* hit_bits, int32_t* hit_index, int32_t* hit_count,
                       void* stream) {
  if (bs <= 0 || num_cam <= 0 || num_cam > 32 || HW <= 0 || D <= 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "bev_point_sampling: invalid sizes bs=%d cams=%d HW=%d D=%d",
                     bs, num_cam, HW, D);
  if (!ref_3d || !lidar2img || !pc_range_host || !ref_cam || !bev_mask || !hit_bits ||
      ((hit_index == nullptr) != (hit_count == nullptr)))
    return set_error(MSDA_ERR_BAD_ARGUMENT, "bev_point_sampling: NULL pointer");
  return launch_point_sampling(ref_3d, lidar2img, pc_range_host, img_h, img_w, bs, num_cam, HW, D,
                               ref_cam, bev_mask, hit_bits, hit_index, hit_count,
                               static_cast<cudaStream_t>(stream));
}

int sca_fwd(const void* value, const int64_t* shapes, const int64_t* starts, const void* offsets,
            const void* logits, const float* ref_cam, const uint8_t* bev_mask,
            const uint32_t* hit_bits, void* slots, float* attn_out, int bs, int num_cam, int Nk,
            int M, int Dh, int L, int P, int D, int HW, int bev_w, int value_dtype, int coord_dtype,
            int64_t offsets_stride, int64_t logits_stride, void* stream) {
  FusedProblem f;
  f.off_stride = offsets_stride; f.log_stride = logits_stride;
  f.value = value; f.shapes = shapes; f.starts = starts; f.offsets = offsets; f.logits = logits;
  f.ref = ref_cam; f.bev_mask = bev_mask; f.hit_bits = hit_bits; f.out = slots; f.attn_out = attn_out;
  f.bs = bs; f.groups = num_cam; f.Nk = Nk; f.M = M; f.Dh = Dh; f.L = L; f.P = P; f.D = D;
  f.Nq = HW; f.bev_w = bev_w; f.value_dtype = value_dtype; f.coord_dtype = coord_dtype;
  if (int rc = validate_fused(f, false, true, "sca_fwd")) return rc;
  if ((long long)bs * HW == 0) return MSDA_OK;
  return launch_sca_fwd(f, static_cast<cudaStream_t>(stream));
}

int sca_bwd(const void* value, const int64_t* shapes, const int64_t* starts, const void* offsets,
            const void* logits, const float* ref_cam, const uint8_t* bev_mask,
            const uint32_t* hit_bits, const void* g_slots, void* g_value, void* g_offsets,
            void* g_logits, int bs, int num_cam, int Nk, int M, int Dh, int L, int P, int D, int HW,
            int bev_w, int value_dtype, int coord_dtype, int64_t offsets_stride, int64_t logits_stride,
            int accum_dtype, const float* accum_scale, void* g_value_tail, int tail_copies, int tail_pixels,
            void* coarse_records, const int32_t* hit_index, const int32_t* hit_count, void* stream) {
  FusedProblem f;
  if (coarse_records || hit_index || hit_count) {
    if (!coarse_records || !hit_index || !hit_count)
      return set_error(MSDA_ERR_BAD_ARGUMENT, "sca_bwd: coarse_records, hit_index and hit_count go together");
    if (!coarse_supported(Dh, P, value_dtype))
      return set_error(MSDA_ERR_UNSUPPORTED, "sca_bwd: the tensor-core coarse pass needs a 16-bit value dtype, "
                       "head_dim 32 and <= 8 points (sca_coarse_workspace_bytes() returns 0 otherwise)");
    if (reinterpret_cast<uintptr_t>(coarse_records) % 16 != 0)
      return set_error(MSDA_ERR_BAD_ARGUMENT, "sca_bwd: coarse_records must be 16-byte aligned");
    f.coarse_rec = coarse_records; f.hit_index = hit_index; f.hit_count = hit_count;
  }
  f.off_stride = offsets_stride; f.log_stride = logits_stride;
  f.acc_half = accum_dtype == MSDA_F16; f.acc_scale = accum_scale;
  if (g_value_tail && tail_copies > 0) {
    if (!f.acc_half || tail_pixels <= 0 || tail_pixels > Nk || tail_copies > 64)
      return set_error(MSDA_ERR_BAD_ARGUMENT, "sca_bwd: tail replicas need the fp16 accumulator and 0 < tail_pixels <= Nk");
    f.g_tail = g_value_tail; f.tail_copies = tail_copies; f.tail_px = tail_pixels;
  }
  f.value = value; f.shapes = shapes; f.starts = starts; f.offsets = offsets; f.logits = logits;
  f.ref = ref_cam; f.bev_mask = bev_mask; f.hit_bits = hit_bits; f.g_out = g_slots;
  f.g_value = g_value; f.g_offsets = g_offsets; f.g_logits = g_logits;
  f.bs = bs; f.groups = num_cam; f.Nk = Nk; f.M = M; f.Dh = Dh; f.L = L; f.P = P; f.D = D;
  f.Nq = HW; f.bev_w = bev_w; f.value_dtype = value_dtype; f.coord_dtype = coord_dtype;
  if (int rc = validate_fused(f, true, true, "sca_bwd")) return rc;
  if ((long long)bs * HW == 0) return MSDA_OK;
  return launch_sca_bwd(f, static_cast<cudaStream_t>(stream));
}

int64_t sca_coarse_workspace_bytes(int bs, int num_cam, int HW, int M, int Dh, int P, int value_dtype) {
  if (bs <= 0 || num_cam <= 0 || HW <= 0 || M <= 0 || !coarse_supported(Dh, P, value_dtype)) return 0;
  return coarse_record_bytes(bs, num_cam, HW, M, P);
}

int bev_hit_lists(const uint32_t* hit_bits, int num_cam, int HW, int32_t* hit_index, int32_t* hit_count,
                  void* stream) {
  if (!hit_bits || !hit_index || !hit_count || num_cam <= 0 || num_cam > 32 || HW <= 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "bev_hit_lists: invalid arguments");
  return launch_hit_lists(hit_bits, num_cam, HW, hit_index, hit_count, static_cast<cudaStream_t>(stream));
}

int tsa_fwd(const void* value, const int64_t* shapes, const int64_t* starts, const void* offsets,
            const void* logits, const float* ref, void* out, int bs, int Q, int Nk, int M, int Dh,
            int L, int P, int Nq, int bev_w, float clamp, int value_dtype, int coord_dtype,
            int64_t offsets_stride, int64_t logits_stride, void* stream) {
  FusedProblem f;
  f.off_stride = offsets_stride; f.log_stride = logits_stride;
  f.value = value; f.shapes = shapes; f.starts = starts; f.offsets = offsets; f.logits = logits;
  f.ref = ref; f.out = out;
  f.bs = bs; f.groups = Q; f.Nk = Nk; f.M = M; f.Dh = Dh; f.L = L; f.P = P; f.Nq = Nq;
  f.bev_w = bev_w; f.clamp = clamp; f.value_dtype = value_dtype; f.coord_dtype = coord_dtype;
  if (int rc = validate_fused(f, false, false, "tsa_fwd")) return rc;
  if ((long long)bs * Nq == 0) return MSDA_OK;
  return launch_tsa_fwd(f, static_cast<cudaStream_t>(stream));
}

int tsa_bwd(const void* value, const int64_t* shapes, const int64_t* starts, const void* offsets,
            const void* logits, const float* ref, const void* g_out, void* g_value,
            void* g_offsets, void* g_logits, int bs, int Q, int Nk, int M, int Dh, int L, int P,
            int Nq, int bev_w, float clamp, int value_dtype, int coord_dtype, int64_t offsets_stride,
            int64_t logits_stride, int accum_dtype, const float* accum_scale, void* stream) {
  FusedProblem f;
  f.off_stride = offsets_stride; f.log_stride = logits_stride;
  f.acc_half = accum_dtype == MSDA_F16; f.acc_scale = accum_scale;
  f.value = value; f.shapes = shapes; f.starts = starts; f.offsets = offsets; f.logits = logits;
  f.ref = ref; f.g_out = g_out; f.g_value = g_value; f.g_offsets = g_offsets; f.g_logits = g_logits;
  f.bs = bs; f.groups = Q; f.Nk = Nk; f.M = M; f.Dh = Dh; f.L = L; f.P = P; f.Nq = Nq;
  f.bev_w = bev_w; f.clamp = clamp; f.value_dtype = value_dtype; f.coord_dtype = coord_dtype;
  if (int rc = validate_fused(f, true, false, "tsa_bwd")) return rc;
  if ((long long)bs * Nq == 0) return MSDA_OK;
  return launch_tsa_bwd(f, static_cast<cudaStream_t>(stream));
}

int rowops_workspace_rows(void) { return rowops_partial_rows(); }

int ln_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
           int64_t rows, int C, float eps, int dtype, void* stream) {
  if (rows < 0 || C <= 0) return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_fwd: invalid sizes");
  if (rows == 0) return MSDA_OK;
  if (!x || !gamma || !beta || !y || !mean || !rstd) return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_fwd: NULL pointer");
  return launch_ln(false, x, nullptr, gamma, beta, y, mean, rstd, nullptr, nullptr, nullptr, rows, C, eps,
                   dtype, nullptr, nullptr, false, RowDropout{}, nullptr, static_cast<cudaStream_t>(stream));
}

int ln_residual_fwd(const void* x, const void* residual, const void* gamma, const void* beta, void* sum_out,
                    void* y, float* mean, float* rstd, int64_t rows, int C, float eps, int dtype,
                    void* stream) {
  if (rows < 0 || C <= 0) return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_residual_fwd: invalid sizes");
  if (rows == 0) return MSDA_OK;
  if (!x || !residual || !gamma || !beta || !sum_out || !y || !mean || !rstd)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_residual_fwd: NULL pointer");
  return launch_ln(false, x, nullptr, gamma, beta, y, mean, rstd, nullptr, nullptr, nullptr, rows, C, eps,
                   dtype, residual, sum_out, false, RowDropout{}, nullptr, static_cast<cudaStream_t>(stream));
}

int ln_residual_dropout_fwd(const void* x, const void* residual, const void* gamma, const void* beta, void* sum_out,
                            void* y, float* mean, float* rstd, int64_t rows, int C, float eps, int dtype,
                            const void* rng_state, void* key_save, uint32_t site, float p, void* stream) {
  if (rows < 0 || C <= 0) return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_residual_dropout_fwd: invalid sizes");
  if (rows == 0) return MSDA_OK;
  if (!x || !residual || !gamma || !beta || !sum_out || !y || !mean || !rstd || !rng_state)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_residual_dropout_fwd: NULL pointer");
  RowDropout rd;
  rd.key = rng_state; rd.key_save = key_save; rd.site = site; rd.p = p;
  return launch_ln(false, x, nullptr, gamma, beta, y, mean, rstd, nullptr, nullptr, nullptr, rows, C, eps,
                   dtype, residual, sum_out, false, rd, nullptr, static_cast<cudaStream_t>(stream));
}

int ln_bwd(const void* x, const void* dy, const void* gamma, const float* mean, const float* rstd,
           void* dx, void* dgamma_dbeta, float* partial, int64_t rows, int C, int dtype, void* stream) {
  if (rows < 0 || C <= 0) return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_bwd: invalid sizes");
  if (!x || !dy || !gamma || !mean || !rstd || !dx || !dgamma_dbeta || !partial)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_bwd: NULL pointer");
  return launch_ln(true, x, dy, gamma, nullptr, nullptr, const_cast<float*>(mean), const_cast<float*>(rstd),
                   dx, dgamma_dbeta, partial, rows, C, 0.f, dtype, nullptr, nullptr, false, RowDropout{}, nullptr,
                   static_cast<cudaStream_t>(stream));
}

int ln_bwd_dxsum(const void* x, const void* dy, const void* gamma, const float* mean, const float* rstd,
                 void* dx, void* dgamma_dbeta_dxsum, float* partial, int64_t rows, int C, int dtype,
                 void* stream) {
  if (rows < 0 || C <= 0) return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_bwd_dxsum: invalid sizes");
  if (!x || !dy || !gamma || !mean || !rstd || !dx || !dgamma_dbeta_dxsum || !partial)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_bwd_dxsum: NULL pointer");
  return launch_ln(true, x, dy, gamma, nullptr, nullptr, const_cast<float*>(mean), const_cast<float*>(rstd),
                   dx, dgamma_dbeta_dxsum, partial, rows, C, 0.f, dtype, nullptr, nullptr, true, RowDropout{}, nullptr,
                   static_cast<cudaStream_t>(stream));
}

int ln_bwd_dxsum_dropout(const void* x, const void* dy, const void* gamma, const float* mean, const float* rstd,
                         void* dx, void* dx_masked, void* dgamma_dbeta_dxsum, float* partial, int64_t rows, int C,
                         int dtype, const void* key, uint32_t site, float p, void* stream) {
  if (rows < 0 || C <= 0) return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_bwd_dxsum_dropout: invalid sizes");
  if (!x || !dy || !gamma || !mean || !rstd || !dx || !dx_masked || !dgamma_dbeta_dxsum || !partial || !key)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "ln_bwd_dxsum_dropout: NULL pointer");
  RowDropout rd;
  rd.key = key; rd.site = site; rd.p = p;
  return launch_ln(true, x, dy, gamma, nullptr, nullptr, const_cast<float*>(mean), const_cast<float*>(rstd),
                   dx, dgamma_dbeta_dxsum, partial, rows, C, 0.f, dtype, nullptr, nullptr, true, rd, dx_masked,
                   static_cast<cudaStream_t>(stream));
}

int relu_dropout_fwd(void* x, int64_t n, int dtype, const void* rng_state, void* key_save, uint32_t site, float p,
                     void* stream) {
  if (n < 0 || (n > 0 && !x) || !rng_state) return set_error(MSDA_ERR_BAD_ARGUMENT, "relu_dropout_fwd: bad argument");
  RowDropout rd;
  rd.key = rng_state; rd.key_save = key_save; rd.site = site; rd.p = p;
  return launch_relu_dropout(x, nullptr, n, dtype, rd, static_cast<cudaStream_t>(stream));
}

int dropout_keep_mask(uint8_t* mask, int64_t n, int dtype, const void* key, uint32_t site, float p, void* stream) {
  if (n < 0 || (n > 0 && !mask) || !key) return set_error(MSDA_ERR_BAD_ARGUMENT, "dropout_keep_mask: bad argument");
  RowDropout rd;
  rd.key = key; rd.site = site; rd.p = p;
  return launch_relu_dropout(nullptr, mask, n, dtype, rd, static_cast<cudaStream_t>(stream));
}

int colsum(const void* x, void* out, float* partial, int64_t rows, int C, int dtype, int out_dtype,
           void* stream) {
  if (rows < 0 || C <= 0) return set_error(MSDA_ERR_BAD_ARGUMENT, "colsum: invalid sizes");
  if (!x || !out || !partial) return set_error(MSDA_ERR_BAD_ARGUMENT, "colsum: NULL pointer");
  return launch_colsum(x, nullptr, nullptr, out, partial, rows, C, dtype, out_dtype, 1.f, static_cast<cudaStream_t>(stream));
}

int relu_bwd_colsum(const void* dy, const void* y, void* dx, void* colsum_out, float* partial, int64_t rows, int C,
                    int dtype, int out_dtype, float scale, void* stream) {
  if (rows < 0 || C <= 0) return set_error(MSDA_ERR_BAD_ARGUMENT, "relu_bwd_colsum: invalid sizes");
  if (!dy || !y || !dx || !colsum_out || !partial) return set_error(MSDA_ERR_BAD_ARGUMENT, "relu_bwd_colsum: NULL pointer");
  return launch_colsum(dy, y, dx, colsum_out, partial, rows, C, dtype, out_dtype, scale > 0.f ? scale : 1.f,
                       static_cast<cudaStream_t>(stream));
}

int grad_amax_scale(const void* g, int64_t n, int dtype, float limit, float* ws, void* stream) {
  if (n < 0 || !ws || (n > 0 && !g) || !(limit > 0.f))
    return set_error(MSDA_ERR_BAD_ARGUMENT, "grad_amax_scale: bad argument");
  return launch_grad_scale(g, n, dtype, limit, ws, nullptr, 0, static_cast<cudaStream_t>(stream));
}

int grad_amax_scale_zero(const void* g, int64_t n, int dtype, float limit, float* ws, void* zero, int64_t zero_bytes,
                         void* stream) {
  if (n < 0 || !ws || (n > 0 && !g) || !(limit > 0.f) || zero_bytes < 0 || (zero_bytes > 0 && !zero) ||
      zero_bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(zero) & 15) != 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "grad_amax_scale_zero: bad argument (zero: 16-byte aligned, a multiple of 16 bytes)");
  return launch_grad_scale(g, n, dtype, limit, ws, zero, zero_bytes, static_cast<cudaStream_t>(stream));
}

int unscale_cast(const void* acc_f16, void* out, const float* scale, int64_t n, int out_dtype,
                 const void* tail, int tail_copies, int64_t map_elems, int64_t tail_elems,
                 void* colsum_out, float* partial, int C, int32_t* overflow_flag, void* stream) {
  return unscale_cast_strided(acc_f16, out, scale, n, out_dtype, tail, tail_copies, map_elems, tail_elems, colsum_out,
                              partial, C, overflow_flag, 0, stream);
}

int unscale_cast_strided(const void* acc_f16, void* out, const float* scale, int64_t n, int out_dtype,
                         const void* tail, int tail_copies, int64_t map_elems, int64_t tail_elems,
                         void* colsum_out, float* partial, int C, int32_t* overflow_flag, int64_t out_row_stride,
                         void* stream) {
  if (out_row_stride != 0 &&
      (C <= 0 || C % 8 != 0 || n % C != 0 || out_row_stride < C || out_row_stride % 8 != 0 ||
       (reinterpret_cast<uintptr_t>(out) & 15) != 0 || (tail && tail_copies > 0)))
    return set_error(MSDA_ERR_BAD_ARGUMENT, "unscale_cast_strided: rows of C = 8k columns, a row stride >= C that is a "
                     "multiple of 8, a 16-byte aligned output and no tail replicas");
  if (colsum_out) {
    const int64_t m = (tail && tail_copies > 0) ? map_elems : n;
    if (!partial || C <= 0 || C % 8 != 0 || 256 % (C / 8) != 0 || m % C != 0 || n % 8 != 0)
      return set_error(MSDA_ERR_BAD_ARGUMENT, "unscale_cast: column sums need C = 8 * (a divisor of 256) dividing the map size");
  }
  if (n < 0 || (n > 0 && (!acc_f16 || !out || !scale)))
    return set_error(MSDA_ERR_BAD_ARGUMENT, "unscale_cast: bad argument");
  if (tail && tail_copies > 0 &&
      (map_elems <= 0 || tail_elems <= 0 || tail_elems > map_elems || n % map_elems != 0 || map_elems % 8 != 0 ||
       tail_elems % 8 != 0))
    return set_error(MSDA_ERR_BAD_ARGUMENT, "unscale_cast: inconsistent tail replica sizes");
  if (n == 0) return MSDA_OK;
  return launch_unscale_cast(acc_f16, out, scale, n, out_dtype, tail, tail_copies, map_elems, tail_elems,
                             partial, colsum_out, C, overflow_flag, out_row_stride, static_cast<cudaStream_t>(stream));
}

int bev_flatten_level(const void* feat, const void* cams_embeds, const void* level_embed, void* feat_flatten,
                      int bs, int num_cam, int C, int hw, int64_t Nk, int64_t start, int dtype, void* stream) {
  if (bs < 0 || num_cam <= 0 || C <= 0 || hw < 0 || Nk < 0 || start < 0 || start + hw > Nk)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "bev_flatten_level: invalid sizes");
  if (dtype != MSDA_F32 && dtype != MSDA_F16 && dtype != MSDA_BF16)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "bev_flatten_level: unknown dtype %d", dtype);
  if ((long long)bs * hw == 0) return MSDA_OK;
  if (!feat || !level_embed || !feat_flatten) return set_error(MSDA_ERR_BAD_ARGUMENT, "bev_flatten_level: NULL pointer");
  return launch_flatten_level(feat, cams_embeds, level_embed, feat_flatten, bs, num_cam, C, hw, Nk, start, dtype,
                              static_cast<cudaStream_t>(stream));
}

int bev_rotate_nearest(const void* prev_bev, void* out, const float* theta, const float* xs, const float* ys,
                       int bs, int bev_h, int bev_w, int C, int dtype, void* stream) {
  if (bs < 0 || bev_h <= 0 || bev_w <= 0 || C <= 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "bev_rotate_nearest: invalid sizes");
  if (dtype != MSDA_F32 && dtype != MSDA_F16 && dtype != MSDA_BF16)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "bev_rotate_nearest: unknown dtype %d", dtype);
  if (bs == 0) return MSDA_OK;
  if (!prev_bev || !out || !theta || !xs || !ys || prev_bev == out)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "bev_rotate_nearest: NULL or aliased pointer");
  if ((reinterpret_cast<uintptr_t>(prev_bev) | reinterpret_cast<uintptr_t>(out)) & 15u)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "bev_rotate_nearest: tensors must be 16-byte aligned");
  return launch_rotate_nearest(prev_bev, out, theta, xs, ys, bs, bev_h, bev_w, C, dtype,
                               static_cast<cudaStream_t>(stream));
}

int64_t linear_wgrad_workspace_floats(int O, int I) { return wgrad_workspace_floats(O, I); }

int linear_wgrad(const void* dy, const void* x, void* dW, void* db, float* workspace, int64_t N, int O, int I,
                 int dtype, void* stream) {
  if (N < 0 || O <= 0 || I <= 0 || O % 8 != 0 || I % 8 != 0)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "linear_wgrad: N >= 0 and O, I multiples of 8 required (got %lld, %d, %d)",
                     (long long)N, O, I);
  if (!dy || !x || !dW || !workspace) return set_error(MSDA_ERR_BAD_ARGUMENT, "linear_wgrad: NULL pointer");
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dW) |
       reinterpret_cast<uintptr_t>(workspace)) & 15u)
    return set_error(MSDA_ERR_BAD_ARGUMENT, "linear_wgrad: tensors must be 16-byte aligned");
  return launch_wgrad(dy, x, dW, db, workspace, N, O, I, dtype, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
