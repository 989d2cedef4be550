// fused bias + leaky-relu (+ its one-pass backward with the bias-gradient reduction fused in).
// Semantics: utils/op/fused_bias_act_kernel.cu:18-49 of the reference (restated, not copied):
// HBM-streaming kernels, 16-byte accesses, bias index computed once per 16-byte vector.
#include "common.cuh"

namespace te {

template <typename A>
__device__ __forceinline__ A bias_act_apply(A x, A ref, int code, A alpha, A scale) {
  A y;
  switch (code) {
    case 12:
    case 32: y = A(0); break;
    case 30: y = x > A(0) ? x : x * alpha; break;
    case 31: y = ref > A(0) ? x : x * alpha; break;
    default: y = x; break;  // 10, 11
  }
  return y * scale;
}

// BIAS_MODE 0: none, 1: one bias per vector (step_b % VEC == 0), 2: consecutive biases
// (step_b == 1 && size_b % VEC == 0).  IDX is uint32_t when n < 2^31 (cheap division).
template <typename T, int VEC, int BIAS_MODE, typename IDX, typename BT>
__global__ void __launch_bounds__(256)
fused_bias_act_kernel(T* __restrict__ out, const T* __restrict__ x, const BT* __restrict__ bias,
                      const T* __restrict__ ref, int code, float alpha_f, float scale_f, IDX n_vec,
                      IDX step_b, IDX size_b) {
  using A = typename Acc<T>::type;
  const A alpha = A(alpha_f), scale = A(scale_f);
  struct alignas(sizeof(T) * VEC) V { T v[VEC]; };
  const IDX stride = IDX(gridDim.x) * blockDim.x;
  for (IDX iv = IDX(blockIdx.x) * blockDim.x + threadIdx.x; iv < n_vec; iv += stride) {
    V xin = reinterpret_cast<const V*>(x)[iv];
    V rin;
    if (code == 31 && ref != nullptr) rin = reinterpret_cast<const V*>(ref)[iv];
    A b[VEC];
    if (BIAS_MODE == 1) {
      A bb = A(to_acc(bias[(iv * VEC / step_b) % size_b]));
#pragma unroll
      for (int j = 0; j < VEC; ++j) b[j] = bb;
    } else if (BIAS_MODE == 2) {
      IDX c0 = (iv * VEC) % size_b;
#pragma unroll
      for (int j = 0; j < VEC; ++j) b[j] = A(to_acc(bias[c0 + j]));
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) b[j] = A(0);
    }
    V o;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      A r = (code == 31 && ref != nullptr) ? to_acc(rin.v[j]) : A(0);
      o.v[j] = from_acc<T, A>(bias_act_apply<A>(to_acc(xin.v[j]) + b[j], r, code, alpha, scale));
    }
    reinterpret_cast<V*>(out)[iv] = o;
  }
}

// Backward, NCHW-like (step_b > 1): one WARP per (plane, chunk); the plane's channel is uniform,
// so the bias gradient is a warp reduction + one atomic.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
fused_bias_act_bwd_planes(T* __restrict__ gin, typename Acc<T>::type* __restrict__ gbias,
                          const T* __restrict__ g, const T* __restrict__ ref, float alpha_f,
                          float scale_f, int64_t planes, int64_t step_b, int64_t size_b,
                          int chunk, int chunks_per_plane) {
  using A = typename Acc<T>::type;
  struct alignas(sizeof(T) * VEC) V { T v[VEC]; };
  const A alpha = A(alpha_f), scale = A(scale_f);
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = int64_t(gridDim.x) * (blockDim.x >> 5);
  const int64_t n_items = planes * chunks_per_plane;
  for (int64_t item = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); item < n_items;
       item += warps_total) {
    const int64_t plane = item / chunks_per_plane;
    const int cid = int(item - plane * chunks_per_plane);
    const int64_t lo = int64_t(cid) * chunk;
    const int64_t hi = (lo + chunk < step_b) ? lo + chunk : step_b;
    const int64_t base = plane * step_b;
    A acc = A(0);
    for (int64_t e = lo + int64_t(lane) * VEC; e < hi; e += 32 * VEC) {
      V gv = *reinterpret_cast<const V*>(g + base + e);
      V rv = *reinterpret_cast<const V*>(ref + base + e);
      V o;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        A gg = to_acc(gv.v[j]);
        A y = (to_acc(rv.v[j]) > A(0) ? gg : gg * alpha) * scale;
        o.v[j] = from_acc<T, A>(y);
        acc += y;
      }
      *reinterpret_cast<V*>(gin + base + e) = o;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) atomicAdd(gbias + (plane % size_b), acc);
  }
}

// Backward, channels-last / [rows, C] (step_b == 1): a 256-thread CTA covers `ctpb` column vectors x
// (256 / ctpb) rows at a time; each thread keeps VEC fixed columns and walks down its rows of the
// chunk; partial sums meet in shared memory and leave as one atomic per column per CTA.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
fused_bias_act_bwd_cols(T* __restrict__ gin, typename Acc<T>::type* __restrict__ gbias,
                        const T* __restrict__ g, const T* __restrict__ ref, float alpha_f,
                        float scale_f, int64_t rows, int64_t cols, int rows_per_block, int ctpb) {
  using A = typename Acc<T>::type;
  struct alignas(sizeof(T) * VEC) V { T v[VEC]; };
  __shared__ A red[256 * VEC];
  const A alpha = A(alpha_f), scale = A(scale_f);
  const int tx = threadIdx.x % ctpb, ty = threadIdx.x / ctpb;
  const int tys = blockDim.x / ctpb;
  const int64_t c0 = (int64_t(blockIdx.x) * ctpb + tx) * VEC;
  const bool col_ok = c0 < cols && ty < tys;
  const int64_t r_lo = int64_t(blockIdx.y) * rows_per_block;
  const int64_t r_hi = (r_lo + rows_per_block < rows) ? r_lo + rows_per_block : rows;
  A acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = A(0);
  if (col_ok) {
    // 4 rows per trip: 8 independent 16-byte loads in flight per thread
    int64_t r = r_lo + ty;
    for (; r + 3 * int64_t(tys) < r_hi; r += 4 * int64_t(tys)) {
      V gv[4], rv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t off = (r + u * int64_t(tys)) * cols + c0;
        gv[u] = *reinterpret_cast<const V*>(g + off);
        rv[u] = *reinterpret_cast<const V*>(ref + off);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        V o;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          A gg = to_acc(gv[u].v[j]);
          A y = (to_acc(rv[u].v[j]) > A(0) ? gg : gg * alpha) * scale;
          o.v[j] = from_acc<T, A>(y);
          acc[j] += y;
        }
        *reinterpret_cast<V*>(gin + (r + u * int64_t(tys)) * cols + c0) = o;
      }
    }
    for (; r < r_hi; r += tys) {
      const int64_t off = r * cols + c0;
      V gv = *reinterpret_cast<const V*>(g + off);
      V rv = *reinterpret_cast<const V*>(ref + off);
      V o;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        A gg = to_acc(gv.v[j]);
        A y = (to_acc(rv.v[j]) > A(0) ? gg : gg * alpha) * scale;
        o.v[j] = from_acc<T, A>(y);
        acc[j] += y;
      }
      *reinterpret_cast<V*>(gin + off) = o;
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) red[threadIdx.x * VEC + j] = acc[j];
  __syncthreads();
  if (ty == 0 && c0 < cols) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      A sum = A(0);
      for (int r = 0; r < tys; ++r) sum += red[(r * ctpb + tx) * VEC + j];
      atomicAdd(gbias + c0 + j, sum);
    }
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T, int VEC, typename IDX, typename BT>
static int launch_fwd_mode(T* out, const T* x, const BT* bias, const T* ref, int code, float alpha,
                           float scale, int64_t n, int64_t step_b, int64_t size_b, int mode,
                           cudaStream_t st) {
  const int64_t n_vec = n / VEC;
  const int threads = 256;
  const int blocks = grid_for(n_vec, threads, 16);
  if (mode == 0)
    fused_bias_act_kernel<T, VEC, 0, IDX, BT><<<blocks, threads, 0, st>>>(
        out, x, bias, ref, code, alpha, scale, IDX(n_vec), IDX(1), IDX(1));
  else if (mode == 1)
    fused_bias_act_kernel<T, VEC, 1, IDX, BT><<<blocks, threads, 0, st>>>(
        out, x, bias, ref, code, alpha, scale, IDX(n_vec), IDX(step_b), IDX(size_b));
  else
    fused_bias_act_kernel<T, VEC, 2, IDX, BT><<<blocks, threads, 0, st>>>(
        out, x, bias, ref, code, alpha, scale, IDX(n_vec), IDX(step_b), IDX(size_b));
  TE_CHECK_LAUNCH();
  return TE_OK;
}

template <typename T, typename BT>
static int fused_bias_act_typed(void* out_, const void* x_, const void* bias_, const void* ref_,
                                int code, float alpha, float scale, int64_t n, int64_t step_b,
                                int64_t size_b, cudaStream_t st) {
  T* out = static_cast<T*>(out_);
  const T* x = static_cast<const T*>(x_);
  const BT* bias = static_cast<const BT*>(bias_);
  const T* ref = static_cast<const T*>(ref_);
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (n % VEC == 0) && aligned16(out) && aligned16(x) && (!ref || aligned16(ref));
  int mode = 0, vec = 1;
  if (!bias) {
    mode = 0;
    vec = vec_ok ? VEC : 1;
  } else if (vec_ok && step_b % VEC == 0) {
    mode = 1;
    vec = VEC;
  } else if (vec_ok && step_b == 1 && size_b % VEC == 0) {
    mode = 2;
    vec = VEC;
  } else {
    mode = 1;  // scalar: one bias per element
    vec = 1;
  }
  const bool small = n < (int64_t(1) << 31);
  if (vec == VEC) {
    return small ? launch_fwd_mode<T, VEC, uint32_t, BT>(out, x, bias, ref, code, alpha, scale, n,
                                                         step_b, size_b, mode, st)
                 : launch_fwd_mode<T, VEC, uint64_t, BT>(out, x, bias, ref, code, alpha, scale, n,
                                                         step_b, size_b, mode, st);
  }
  return small ? launch_fwd_mode<T, 1, uint32_t, BT>(out, x, bias, ref, code, alpha, scale, n, step_b,
                                                     size_b, mode, st)
               : launch_fwd_mode<T, 1, uint64_t, BT>(out, x, bias, ref, code, alpha, scale, n, step_b,
                                                     size_b, mode, st);
}

template <typename T>
static int fused_bias_act_bwd_typed(void* gin_, void* gbias_, const void* g_, const void* ref_,
                                    float alpha, float scale, int64_t n, int64_t step_b,
                                    int64_t size_b, cudaStream_t st) {
  using A = typename Acc<T>::type;
  T* gin = static_cast<T*>(gin_);
  A* gbias = static_cast<A*>(gbias_);
  const T* g = static_cast<const T*>(g_);
  const T* ref = static_cast<const T*>(ref_);
  constexpr int VEC = 16 / sizeof(T);
  const bool al = aligned16(gin) && aligned16(g) && aligned16(ref);
  if (step_b == 1) {
    const int64_t cols = size_b, rows = n / size_b;
    TE_CHECK_ARG(rows * cols == n, "fused_bias_act_bwd: n not a multiple of size_b");
    const bool v = al && cols % VEC == 0;
    const int vec = v ? VEC : 1;
    const int threads = 256;
    const int64_t col_threads = cols / vec;
    int ctpb = 1;
    while (ctpb < 256 && ctpb < col_threads) ctpb <<= 1;  // power of two <= 256 covering the columns
    const unsigned gx = unsigned((col_threads + ctpb - 1) / ctpb);
    const int tys = threads / ctpb;
    // one full wave: 4 CTAs per SM (5 fit by registers; 1184 CTAs used to leave a 0.6-wave tail)
    int64_t want = (int64_t(kNumSMs) * 4) / gx;
    if (want < 1) want = 1;
    int64_t rpb = (rows + want - 1) / want;
    if (rpb < 16 * tys) rpb = 16 * tys;
    const unsigned gy = unsigned((rows + rpb - 1) / rpb);
    dim3 grid(gx, gy);
    if (v)
      fused_bias_act_bwd_cols<T, VEC><<<grid, threads, 0, st>>>(gin, gbias, g, ref, alpha, scale,
                                                                rows, cols, int(rpb), ctpb);
    else
      fused_bias_act_bwd_cols<T, 1><<<grid, threads, 0, st>>>(gin, gbias, g, ref, alpha, scale,
                                                              rows, cols, int(rpb), ctpb);
    TE_CHECK_LAUNCH();
    return TE_OK;
  }
  TE_CHECK_ARG(n % step_b == 0, "fused_bias_act_bwd: n not a multiple of step_b");
  const int64_t planes = n / step_b;
  const bool v = al && step_b % VEC == 0;
  const int chunk = 2048;
  const int cpp = int((step_b + chunk - 1) / chunk);
  const int threads = 256;
  const int blocks = grid_for(planes * cpp * 32, threads, 32);
  if (v)
    fused_bias_act_bwd_planes<T, VEC><<<blocks, threads, 0, st>>>(gin, gbias, g, ref, alpha, scale,
                                                                  planes, step_b, size_b, chunk, cpp);
  else
    fused_bias_act_bwd_planes<T, 1><<<blocks, threads, 0, st>>>(gin, gbias, g, ref, alpha, scale,
                                                                planes, step_b, size_b, chunk, cpp);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

}  // namespace te

extern "C" int te_fused_bias_act(void* out, const void* x, const void* bias, const void* ref,
                                 int act, int grad, float alpha, float scale, int64_t n,
                                 int64_t step_b, int64_t size_b, int dtype, void* stream) {
  using namespace te;
  TE_CHECK_ARG(n >= 0, "fused_bias_act: negative n");
  if (n == 0) return TE_OK;
  TE_CHECK_ARG(out && x, "fused_bias_act: null tensor pointer");
  TE_CHECK_ARG((act == 1 || act == 3) && grad >= 0 && grad <= 2,
               "fused_bias_act: act must be 1 or 3 and grad 0..2 (got act=%d grad=%d)", act, grad);
  TE_CHECK_ARG(!bias || (step_b >= 1 && size_b >= 1), "fused_bias_act: bad bias geometry");
  const int code = act * 10 + grad;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool bias_f32 = (dtype & TE_BIAS_F32) != 0;  // 16-bit activations with an f32 bias (master parameter)
  dtype &= 0xff;
  if (bias_f32 && dtype == TE_BF16)
    return fused_bias_act_typed<__nv_bfloat16, float>(out, x, bias, ref, code, alpha, scale, n, step_b, size_b, st);
  if (bias_f32 && dtype == TE_F16)
    return fused_bias_act_typed<__half, float>(out, x, bias, ref, code, alpha, scale, n, step_b, size_b, st);
  switch (dtype) {
    case TE_F32: return fused_bias_act_typed<float, float>(out, x, bias, ref, code, alpha, scale, n, step_b, size_b, st);
    case TE_BF16: return fused_bias_act_typed<__nv_bfloat16, __nv_bfloat16>(out, x, bias, ref, code, alpha, scale, n, step_b, size_b, st);
    case TE_F16: return fused_bias_act_typed<__half, __half>(out, x, bias, ref, code, alpha, scale, n, step_b, size_b, st);
    case TE_F64: return fused_bias_act_typed<double, double>(out, x, bias, ref, code, alpha, scale, n, step_b, size_b, st);
  }
  set_error("fused_bias_act: unknown dtype %d", dtype);
  return TE_ERR_INVALID;
}

extern "C" int te_fused_bias_act_bwd(void* grad_in, void* grad_bias, const void* g,
                                     const void* ref, float alpha, float scale, int64_t n,
                                     int64_t step_b, int64_t size_b, int dtype, void* stream) {
  using namespace te;
  if (n == 0) return TE_OK;
  TE_CHECK_ARG(grad_in && g && ref, "fused_bias_act_bwd: null tensor pointer");
  if (!grad_bias)
    return te_fused_bias_act(grad_in, g, nullptr, ref, 3, 1, alpha, scale, n, 1, 1, dtype, stream);
  TE_CHECK_ARG(step_b >= 1 && size_b >= 1, "fused_bias_act_bwd: bad bias geometry");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case TE_F32: return fused_bias_act_bwd_typed<float>(grad_in, grad_bias, g, ref, alpha, scale, n, step_b, size_b, st);
    case TE_BF16: return fused_bias_act_bwd_typed<__nv_bfloat16>(grad_in, grad_bias, g, ref, alpha, scale, n, step_b, size_b, st);
    case TE_F16: return fused_bias_act_bwd_typed<__half>(grad_in, grad_bias, g, ref, alpha, scale, n, step_b, size_b, st);
    case TE_F64: return fused_bias_act_bwd_typed<double>(grad_in, grad_bias, g, ref, alpha, scale, n, step_b, size_b, st);
  }
  set_error("fused_bias_act_bwd: unknown dtype %d", dtype);
  return TE_ERR_INVALID;
}
