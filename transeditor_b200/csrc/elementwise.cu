// Channels-last per-(sample, channel) scaling and its reduction: the modulation / demodulation
// multiplies of ModulatedConv2d (model_spatial_query.py:299-304 folds them into per-sample weights;
// here they act on activations) and the matching gradient reductions
//     y[b,p,c] = x[b,p,c] * s[b,c]                 (te_scale_bc)
//     out[b,c] = sum_p a[b,p,c] * b_[b,p,c]        (te_dot_bc)
// HBM-streaming, 16-byte accesses, f32 math; tensors are [B, P, C] with C contiguous (P = H*W).
#include "common.cuh"

namespace te {

// grid.y = sample; a CTA's threads walk the sample's P*C/VEC vectors; the channel index is one
// 32-bit modulo per vector and the per-sample scale row stays in L1.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
scale_bc_kernel(T* __restrict__ y, const T* __restrict__ x, const float* __restrict__ s, uint32_t pc_vec,
                uint32_t c_vec) {
  struct alignas(sizeof(T) * VEC) V { T v[VEC]; };
  const int64_t base = int64_t(blockIdx.y) * pc_vec;
  const float* srow = s + int64_t(blockIdx.y) * c_vec * VEC;
  const V* xin = reinterpret_cast<const V*>(x) + base;
  V* yout = reinterpret_cast<V*>(y) + base;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t iv = blockIdx.x * blockDim.x + threadIdx.x; iv < pc_vec; iv += stride) {
    const float* sp = srow + (iv % c_vec) * VEC;
    V in = xin[iv];
    V o;
#pragma unroll
    for (int j = 0; j < VEC; ++j) o.v[j] = from_acc<T, float>(float(to_acc(in.v[j])) * __ldg(sp + j));
    yout[iv] = o;
  }
}

// one CTA = (sample b, pixel chunk); thread = VEC channels x strided pixels; smem reduce over the
// pixel dimension of the CTA, then one atomic per channel.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
dot_bc_kernel(float* __restrict__ out, const T* __restrict__ a, const T* __restrict__ b_, int p_total, int c,
              int pix_per_cta) {
  struct alignas(sizeof(T) * VEC) V { T v[VEC]; };
  extern __shared__ float red[];  // [rows][c]
  const int c_vec = c / VEC;
  const int b = blockIdx.y;
  const int p_lo = blockIdx.x * pix_per_cta;
  const int p_hi = min(p_total, p_lo + pix_per_cta);
  const int rows = blockDim.x / c_vec > 0 ? blockDim.x / c_vec : 1;
  const int cv = threadIdx.x % c_vec, row = threadIdx.x / c_vec;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (row < rows) {
    const int64_t base = int64_t(b) * p_total * c;
    for (int p = p_lo + row; p < p_hi; p += rows) {
      const int64_t off = (base + int64_t(p) * c) / VEC + cv;
      V va = reinterpret_cast<const V*>(a)[off];
      V vb = reinterpret_cast<const V*>(b_)[off];
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += float(to_acc(va.v[j])) * float(to_acc(vb.v[j]));
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) red[row * c + cv * VEC + j] = acc[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float sum = 0.f;
    for (int r = 0; r < rows; ++r) sum += red[r * c + ch];
    atomicAdd(out + int64_t(b) * c + ch, sum);
  }
}

template <typename T>
static int scale_bc_typed(void* y, const void* x, const float* s, int64_t batch, int64_t pixels, int c,
                          cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  TE_CHECK_ARG(c % VEC == 0, "scale_bc: channel count %d must be a multiple of %d", c, VEC);
  const int64_t pc_vec = pixels * c / VEC;
  if (batch * pc_vec == 0) return TE_OK;
  TE_CHECK_ARG(pc_vec < (int64_t(1) << 31) && batch <= 65535, "scale_bc: tensor too large");
  int64_t want = (int64_t(kNumSMs) * 16 + batch - 1) / batch;       // CTAs per sample for ~16 waves
  int64_t need = (pc_vec + 255) / 256;
  const unsigned gx = unsigned(need < want ? need : want);
  dim3 grid(gx > 0 ? gx : 1, unsigned(batch));
  scale_bc_kernel<T, VEC><<<grid, 256, 0, st>>>(static_cast<T*>(y), static_cast<const T*>(x), s,
                                                uint32_t(pc_vec), uint32_t(c / VEC));
  TE_CHECK_LAUNCH();
  return TE_OK;
}

template <typename T>
static int dot_bc_typed(float* out, const void* a, const void* b, int64_t batch, int64_t pixels, int c,
                        cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  TE_CHECK_ARG(c % VEC == 0, "dot_bc: channel count %d must be a multiple of %d", c, VEC);
  TE_CHECK_ARG(batch <= 65535 && pixels < (int64_t(1) << 31), "dot_bc: tensor too large");
  if (batch * pixels == 0) return TE_OK;
  const int c_vec = c / VEC;
  int threads = 256;
  if (c_vec > threads) threads = ((c_vec + 31) / 32) * 32;
  TE_CHECK_ARG(threads <= 1024, "dot_bc: too many channels (%d)", c);
  const int rows = threads / c_vec > 0 ? threads / c_vec : 1;
  // enough CTAs to fill the chip, at least 4 pixels per thread row
  int64_t want = (int64_t(kNumSMs) * 8 + batch - 1) / batch;
  int64_t ppc = (pixels + want - 1) / want;
  if (ppc < 4 * rows) ppc = 4 * rows;
  const unsigned gx = unsigned((pixels + ppc - 1) / ppc);
  dim3 grid(gx, unsigned(batch));
  const size_t smem = size_t(rows) * c * sizeof(float);
  TE_CHECK_ARG(smem <= 48 * 1024, "dot_bc: shared-memory footprint too large");
  dot_bc_kernel<T, VEC><<<grid, threads, smem, st>>>(out, static_cast<const T*>(a), static_cast<const T*>(b),
                                                     int(pixels), c, int(ppc));
  TE_CHECK_LAUNCH();
  return TE_OK;
}


// Split-operand planes of an f32 tensor (see conv_tc.cu): 8 channels per thread = two 16-byte loads and one
// 16-byte store per plane.  Optional per-(sample, channel) modulation fused in (saves the scale_bc pass).
template <int NSEG>
__global__ void __launch_bounds__(256)
split_bf16_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ x, const float* __restrict__ s,
                  uint32_t pc_vec, uint32_t c_vec, int64_t plane_vecs) {
  const int64_t base = int64_t(blockIdx.y) * pc_vec;
  const float4* xin = reinterpret_cast<const float4*>(x) + 2 * base;
  uint4* out = reinterpret_cast<uint4*>(dst) + base;
  const float* srow = s ? s + int64_t(blockIdx.y) * c_vec * 8 : nullptr;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t iv = blockIdx.x * blockDim.x + threadIdx.x; iv < pc_vec; iv += stride) {
    const float4 a = xin[2 * iv], b = xin[2 * iv + 1];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (srow) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(srow + (iv % c_vec) * 8));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(srow + (iv % c_vec) * 8) + 1);
      v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w;
      v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
    }
#pragma unroll
    for (int sg = 0; sg < NSEG; ++sg) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
        v[2 * j] -= __bfloat162float(h0);       // exact in f32: the remainder has at most 16 significant bits
        v[2 * j + 1] -= __bfloat162float(h1);
        w[j] = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
      }
      out[sg * plane_vecs + iv] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}
}  // namespace te

extern "C" int te_scale_bc(void* y, const void* x, const float* s, int64_t batch, int64_t pixels, int channels,
                           int dtype, void* stream) {
  using namespace te;
  TE_CHECK_ARG(y && x && s, "scale_bc: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == TE_BF16) return scale_bc_typed<__nv_bfloat16>(y, x, s, batch, pixels, channels, st);
  if (dtype == TE_F32) return scale_bc_typed<float>(y, x, s, batch, pixels, channels, st);
  if (dtype == TE_F16) return scale_bc_typed<__half>(y, x, s, batch, pixels, channels, st);
  set_error("scale_bc: unsupported dtype %d", dtype);
  return TE_ERR_UNSUPPORTED;
}

extern "C" int te_split_bf16(void* dst, const float* x, const float* s, int64_t batch, int64_t pixels, int channels,
                             int nseg, void* stream) {
  using namespace te;
  TE_CHECK_ARG(dst && x, "split_bf16: null pointer");
  TE_CHECK_ARG(nseg >= 1 && nseg <= 3, "split_bf16: 1, 2 or 3 planes");
  TE_CHECK_ARG(channels % 8 == 0, "split_bf16: channel count %d must be a multiple of 8", channels);
  TE_CHECK_ARG(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(s)) & 15) == 0,
               "split_bf16: pointers must be 16-byte aligned");
  const int64_t pc_vec = pixels * channels / 8;
  if (batch * pc_vec == 0) return TE_OK;
  TE_CHECK_ARG(pc_vec < (int64_t(1) << 31) && batch <= 65535, "split_bf16: tensor too large");
  int64_t want = (int64_t(kNumSMs) * 16 + batch - 1) / batch;
  int64_t need = (pc_vec + 255) / 256;
  const unsigned gx = unsigned(need < want ? need : want);
  dim3 grid(gx > 0 ? gx : 1, unsigned(batch));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst);
  const int64_t plane = batch * pc_vec;
  if (nseg == 1) split_bf16_kernel<1><<<grid, 256, 0, st>>>(d, x, s, uint32_t(pc_vec), uint32_t(channels / 8), plane);
  else if (nseg == 2) split_bf16_kernel<2><<<grid, 256, 0, st>>>(d, x, s, uint32_t(pc_vec), uint32_t(channels / 8), plane);
  else split_bf16_kernel<3><<<grid, 256, 0, st>>>(d, x, s, uint32_t(pc_vec), uint32_t(channels / 8), plane);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

extern "C" int te_dot_bc(float* out, const void* a, const void* b, int64_t batch, int64_t pixels, int channels,
                         int dtype, void* stream) {
  using namespace te;
  TE_CHECK_ARG(out && a && b, "dot_bc: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == TE_BF16) return dot_bc_typed<__nv_bfloat16>(out, a, b, batch, pixels, channels, st);
  if (dtype == TE_F32) return dot_bc_typed<float>(out, a, b, batch, pixels, channels, st);
  if (dtype == TE_F16) return dot_bc_typed<__half>(out, a, b, batch, pixels, channels, st);
  set_error("dot_bc: unsupported dtype %d", dtype);
  return TE_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------------------------
// f32 master weights [O, I, K, K] -> bf16 tap-major [K*K, Opad, Ipad] (dst_n) and/or its transpose
// [K*K, Ipad, Opad] (dst_t), times `scale`, for a whole TABLE of weights in one launch: the operand layouts of
// te_conv_tc for the forward convolution and for its data gradient.  One CTA moves a 32 x 32 (o, i) tile with all its
// taps through shared memory (coalesced reads of 32*K*K consecutive floats per output channel, 64-byte writes).
namespace te {

struct PackTask {
  const float* src;
  __nv_bfloat16* dst_n;
  __nv_bfloat16* dst_t;
  int O, I, kk, opad, ipad;
  float scale;
  int tile0;
  int nseg;   // split-operand planes written (plane stride = kk * opad * ipad)
};
constexpr int PACK_MAX_TASKS = 64;
struct PackParams {
  PackTask t[PACK_MAX_TASKS];
  int n_tasks, total_tiles;
};

__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ PackParams P) {
  __shared__ float tile[32][32 * 9 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int j = 0;
  for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
    while (j + 1 < P.n_tasks && t >= P.t[j + 1].tile0) ++j;
    const PackTask& T = P.t[j];
    const int tiles_i = (T.I + 31) >> 5;
    const int local = t - T.tile0;
    const int o0 = (local / tiles_i) * 32, i0 = (local % tiles_i) * 32;
    const int no = T.O - o0 < 32 ? T.O - o0 : 32, ni = T.I - i0 < 32 ? T.I - i0 : 32;
    const int kk = T.kk, row_len = ni * kk;
    for (int r = warp; r < no; r += 8) {
      const float* s = T.src + (int64_t(o0 + r) * T.I + i0) * kk;
      for (int e = lane; e < row_len; e += 32) tile[r][e] = s[e];
    }
    __syncthreads();
    const int64_t plane = int64_t(kk) * T.opad * T.ipad;
    if (T.dst_n) {
      for (int q = warp; q < kk * no; q += 8) {
        const int tap = q / no, r = q - tap * no;
        if (lane < ni) {
          float v = tile[r][lane * kk + tap] * T.scale;
          __nv_bfloat16* d = T.dst_n + (int64_t(tap) * T.opad + o0 + r) * T.ipad + i0 + lane;
          for (int sg = 0; sg < T.nseg; ++sg) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            d[sg * plane] = h;
            v -= __bfloat162float(h);
          }
        }
      }
    }
    if (T.dst_t) {
      for (int q = warp; q < kk * ni; q += 8) {
        const int tap = q / ni, c = q - tap * ni;
        if (lane < no) {
          float v = tile[lane][c * kk + tap] * T.scale;
          __nv_bfloat16* d = T.dst_t + (int64_t(tap) * T.ipad + i0 + c) * T.opad + o0 + lane;
          for (int sg = 0; sg < T.nseg; ++sg) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            d[sg * plane] = h;
            v -= __bfloat162float(h);
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace te

extern "C" int te_pack_weights_tc(const te_pack_task* tasks, int n_tasks, void* stream) {
  using namespace te;
  TE_CHECK_ARG(tasks || n_tasks == 0, "pack_weights_tc: null task table");
  TE_CHECK_ARG(n_tasks >= 0, "pack_weights_tc: negative task count");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int base = 0; base < n_tasks; base += PACK_MAX_TASKS) {
    PackParams P;
    const int n = n_tasks - base < PACK_MAX_TASKS ? n_tasks - base : PACK_MAX_TASKS;
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
      const te_pack_task& s = tasks[base + i];
      TE_CHECK_ARG(s.src && (s.dst_n || s.dst_t), "pack_weights_tc: task %d has a null pointer", base + i);
      TE_CHECK_ARG(s.out_ch > 0 && s.in_ch > 0 && (s.taps == 1 || s.taps == 9),
                   "pack_weights_tc: task %d: need positive channel counts and 1 or 9 taps", base + i);
      PackTask& t = P.t[i];
      t.src = s.src;
      t.dst_n = static_cast<__nv_bfloat16*>(s.dst_n);
      t.dst_t = static_cast<__nv_bfloat16*>(s.dst_t);
      t.O = s.out_ch;
      t.I = s.in_ch;
      t.kk = s.taps;
      t.opad = (s.out_ch + 7) / 8 * 8;
      t.ipad = (s.in_ch + 7) / 8 * 8;
      t.scale = s.scale;
      t.nseg = s.split < 2 ? 1 : s.split;
      TE_CHECK_ARG(s.split >= 0 && s.split <= 3, "pack_weights_tc: task %d: split must be 0..3", base + i);
      t.tile0 = tiles;
      tiles += ((s.out_ch + 31) / 32) * ((s.in_ch + 31) / 32);
    }
    P.n_tasks = n;
    P.total_tiles = tiles;
    if (tiles == 0) continue;
    const int grid = tiles < 8 * kNumSMs ? tiles : 8 * kNumSMs;
    pack_weights_kernel<<<grid, 256, 0, st>>>(P);
    TE_CHECK_LAUNCH();
  }
  return TE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Weight-gradient unpack: the tcgen05 weight-gradient kernel accumulates tap-major [T][A][Bd] f32 (the layout whose
// rows are UMMA output tiles); the master weight and its gradient are [O][I][T] (taps innermost).  ATen's generic
// strided copy did this permute at ~1 TB/s, 117 times per training iteration; here a 32 x 32 (o, i) tile of all T
// taps goes through shared memory so that both the reads (along the fastest dimension of the workspace) and the
// writes (T * 32 contiguous floats per output row) are coalesced.
//   trans == 0:  out[b][o][i][t] = ws[b][t][o * ld + i]      (ws rows = output channels)
//   trans == 1:  out[b][o][i][t] = ws[b][t][i * ld + o]      (transposed convolutions keep the weight as [O, I]:
//                                                              the workspace's logical (cout, cin) is (I, O))
namespace te {

template <int TAPS>
__global__ void __launch_bounds__(256)
wgrad_unpack_kernel(float* __restrict__ out, float* __restrict__ ws, int o_dim, int i_dim, int rows, int ld, int trans,
                    int clear) {
  extern __shared__ float wu_tile[];   // [TAPS][32][33]
  const int o0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  float* wsb = ws + int64_t(blockIdx.z) * TAPS * rows * ld;
  float* outb = out + int64_t(blockIdx.z) * o_dim * i_dim * TAPS;
  const int cx = threadIdx.x & 31, cy = threadIdx.x >> 5;   // 32 x 8
  // all of this thread's loads first (TAPS x 4 independent 4-byte loads in flight), then the shared-memory transpose
  float v[TAPS][4];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = cy + 8 * q;
      const int o = trans ? o0 + cx : o0 + r, i = trans ? i0 + r : i0 + cx;
      const int64_t off = int64_t(t) * rows * ld + (trans ? int64_t(i) * ld + o : int64_t(o) * ld + i);
      const bool ok = o < o_dim && i < i_dim;
      v[t][q] = ok ? wsb[off] : 0.f;
      if (clear && ok) wsb[off] = 0.f;     // hand the accumulator back zeroed: the next launch needs no memset
    }
  }
#pragma unroll
  for (int t = 0; t < TAPS; ++t)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = cy + 8 * q;
      if (!trans) wu_tile[(t * 32 + r) * 33 + cx] = v[t][q];    // tile[t][o_l][i_l]
      else wu_tile[(t * 32 + cx) * 33 + r] = v[t][q];
    }
  __syncthreads();
  constexpr int row_len = 32 * TAPS;                   // floats of one output row segment: (i_l, t) contiguous
  for (int idx = threadIdx.x; idx < 32 * row_len; idx += 256) {
    const int ol = idx / row_len, rem = idx - ol * row_len;
    const int il = rem / TAPS, t = rem - il * TAPS;
    const int o = o0 + ol, i = i0 + il;
    if (o < o_dim && i < i_dim) outb[(int64_t(o) * i_dim + i) * TAPS + t] = wu_tile[(t * 32 + ol) * 33 + il];
  }
}

template <int TAPS>
static int wgrad_unpack_launch(float* out, float* ws, int batch, int o_dim, int i_dim, int rows, int ld, int trans,
                               int clear, cudaStream_t st) {
  dim3 grid((i_dim + 31) / 32, (o_dim + 31) / 32, batch);
  const size_t smem = size_t(TAPS) * 32 * 33 * sizeof(float);
  wgrad_unpack_kernel<TAPS><<<grid, 256, smem, st>>>(out, ws, o_dim, i_dim, rows, ld, trans, clear);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

}  // namespace te

extern "C" int te_wgrad_unpack(float* out, float* ws, int batch, int o_dim, int i_dim, int taps, int rows, int ld,
                               int trans, int clear, void* stream) {
  using namespace te;
  TE_CHECK_ARG(batch >= 0 && o_dim >= 0 && i_dim >= 0, "wgrad_unpack: negative size");
  if (int64_t(batch) * o_dim * i_dim == 0) return TE_OK;
  TE_CHECK_ARG(out && ws, "wgrad_unpack: null pointer");
  TE_CHECK_ARG(trans ? (rows >= i_dim && ld >= o_dim) : (rows >= o_dim && ld >= i_dim),
               "wgrad_unpack: workspace [%d, %d] smaller than the weight [%d, %d]", rows, ld, o_dim, i_dim);
  TE_CHECK_ARG(batch <= 65535, "wgrad_unpack: batch too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (taps) {
    case 1: return wgrad_unpack_launch<1>(out, ws, batch, o_dim, i_dim, rows, ld, trans, clear, st);
    case 4: return wgrad_unpack_launch<4>(out, ws, batch, o_dim, i_dim, rows, ld, trans, clear, st);
    case 9: return wgrad_unpack_launch<9>(out, ws, batch, o_dim, i_dim, rows, ld, trans, clear, st);
    default: break;
  }
  set_error("wgrad_unpack: kernels of 1, 4 (2x2) or 9 (3x3) taps only, got %d", taps);
  return TE_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------------------------
// Weight energy of the demodulation coefficient (model_spatial_query.py:301-303 with the style factored out):
//   energy[r] = coef * SUM_t w[r*taps + t]^2        (r = (o, i) row of the [O, I, k, k] weight, coef = scale^2)
// and its gradient  gw[r*taps + t] = w[r*taps + t] * g[r] * coef2   (coef2 = 2 scale^2).  ATen's reduce kernel ran
// the 9-element inner reduction at 0.7 TB/s (13 us per 512 x 512 layer, 18 layers per forward) plus two tiny
// elementwise launches; here a thread owns one row: 36 contiguous bytes in, one float out.
namespace te {

__global__ void __launch_bounds__(256)
weight_energy_kernel(float* __restrict__ energy, const float* __restrict__ w, int64_t rows, int taps, float coef) {
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < rows; r += int64_t(gridDim.x) * blockDim.x) {
    const float* p = w + r * taps;
    float s = 0.f;
    for (int t = 0; t < taps; ++t) {
      const float v = __ldg(p + t);
      s = fmaf(v, v, s);
    }
    energy[r] = s * coef;
  }
}

__global__ void __launch_bounds__(256)
weight_energy_bwd_kernel(float* __restrict__ gw, const float* __restrict__ w, const float* __restrict__ g, int64_t n,
                         int taps, float coef2) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    gw[i] = __ldg(w + i) * __ldg(g + i / taps) * coef2;
}

}  // namespace te

extern "C" int te_weight_energy(float* energy, const float* w, int64_t rows, int taps, float coef, void* stream) {
  using namespace te;
  TE_CHECK_ARG(rows >= 0 && taps >= 1, "weight_energy: bad shape");
  if (rows == 0) return TE_OK;
  TE_CHECK_ARG(energy && w, "weight_energy: null pointer");
  weight_energy_kernel<<<grid_for(rows, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(energy, w, rows, taps, coef);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

extern "C" int te_weight_energy_bwd(float* gw, const float* w, const float* g, int64_t rows, int taps, float coef2,
                                    void* stream) {
  using namespace te;
  TE_CHECK_ARG(rows >= 0 && taps >= 1, "weight_energy_bwd: bad shape");
  if (rows == 0) return TE_OK;
  TE_CHECK_ARG(gw && w && g, "weight_energy_bwd: null pointer");
  weight_energy_bwd_kernel<<<grid_for(rows * taps, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gw, w, g, rows * taps, taps, coef2);
  TE_CHECK_LAUNCH();
  return TE_OK;
}
