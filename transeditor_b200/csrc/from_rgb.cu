// The discriminator's first layer (ConvLayer(3, C, 1): EqualConv2d 1x1 + FusedLeakyReLU, model_spatial_query.py:
// 806-808 / 731-777) as dedicated HBM-streaming kernels.
//
// Why not the tensor-core engine: K = 3 input channels.  Padded to a 64-channel TMA box the layer read 21x its input,
// ran at 100 TFLOP/s and needed a 403 MB zero-fill + strided copy to build the padded operand; its gradients were two
// more padded launches plus a stand-alone LeakyReLU-backward pass.  Arithmetic here is 3 FMAs per output element: the
// layer is a pure stream — forward reads the f32 NCHW image (12 B per pixel) and writes the channels-last activation;
// backward reads the activation and its gradient ONCE and produces the weight, bias and (optionally) image gradients
// in the same pass, the LeakyReLU mask applied on the fly (the masked gradient is never written).
#include "common.cuh"

namespace te {

template <typename T> struct FrVec;
template <> struct FrVec<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      f[2 * q] = __uint_as_float(w[q] << 16);
      f[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
      w[q] = *reinterpret_cast<uint32_t*>(&t);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct FrVec<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// thread = (pixel row of the CTA, group of 8 output channels); a CTA walks pixels with a grid stride
template <typename T>
__global__ void __launch_bounds__(256)
from_rgb_fwd_kernel(T* __restrict__ y, const float* __restrict__ x, const float* __restrict__ w,
                    const float* __restrict__ bias, int plane, int cout, float wscale, float slope, float gain) {
  const int ovec = cout >> 3, rows = blockDim.x / ovec;
  const int cg = threadIdx.x % ovec, prow = threadIdx.x / ovec;
  float wr[8][3], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int o = cg * 8 + j;
#pragma unroll
    for (int c = 0; c < 3; ++c) wr[j][c] = __ldg(w + o * 3 + c) * wscale;
    br[j] = bias ? __ldg(bias + o) : 0.f;
  }
  // grid.y = sample: no division in the pixel loop
  const int64_t b = blockIdx.y;
  const float* xb = x + b * 3 * plane;
  T* yb = y + b * int64_t(plane) * cout + cg * 8;
#pragma unroll 4
  for (int p = blockIdx.x * rows + prow; p < plane; p += gridDim.x * rows) {
    const float* xp = xb + p;
    const float x0 = __ldg(xp), x1 = __ldg(xp + plane), x2 = __ldg(xp + 2 * plane);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = fmaf(x2, wr[j][2], fmaf(x1, wr[j][1], fmaf(x0, wr[j][0], br[j])));
      f[j] = (u > 0.f ? u : u * slope) * gain;
    }
    FrVec<T>::store(yb + int64_t(p) * cout, f);
  }
}

template <typename T, bool GX>
__global__ void __launch_bounds__(256, GX ? 2 : 3)
from_rgb_bwd_kernel(float* __restrict__ gw, float* __restrict__ gb, float* __restrict__ gx, const T* __restrict__ g,
                    const T* __restrict__ out, const float* __restrict__ x, const float* __restrict__ w,
                    int plane, int cout, float wscale, float slope, float gain) {
  extern __shared__ float fr_red[];   // [rows][cout * 4]
  const int ovec = cout >> 3, rows = blockDim.x / ovec;
  const int cg = threadIdx.x % ovec, prow = threadIdx.x / ovec;
  float wr[8][3];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < 3; ++c) wr[j][c] = GX ? __ldg(w + (cg * 8 + j) * 3 + c) * wscale : 0.f;
  float aw[8][3] = {}, ab[8] = {};
  const float m_pos = gain, m_neg = gain * slope;
  // all threads of a pixel group stay in the loop together (the shuffles below need the full group)
  const int64_t b = blockIdx.y;   // grid.y = sample
  const int step = gridDim.x * rows;
  const int iters = (plane + step - 1) / step;
  const T* gb_ = g + b * int64_t(plane) * cout + cg * 8;
  const T* ob_ = out + b * int64_t(plane) * cout + cg * 8;
#pragma unroll(GX ? 2 : 4)
  for (int it = 0; it < iters; ++it) {
    const int p = (it * gridDim.x + blockIdx.x) * rows + prow;
    const bool live = p < plane;
    float gpv[8] = {}, x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (live) {
      float gv[8], ov[8];
      FrVec<T>::load(gb_ + int64_t(p) * cout, gv);
      FrVec<T>::load(ob_ + int64_t(p) * cout, ov);
      const float* xp = x + b * 3 * plane + p;
      x0 = __ldg(xp); x1 = __ldg(xp + plane); x2 = __ldg(xp + 2 * plane);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        gpv[j] = gv[j] * (ov[j] > 0.f ? m_pos : m_neg);
        ab[j] += gpv[j];
        aw[j][0] = fmaf(gpv[j], x0, aw[j][0]);
        aw[j][1] = fmaf(gpv[j], x1, aw[j][1]);
        aw[j][2] = fmaf(gpv[j], x2, aw[j][2]);
      }
    }
    if (GX) {  // image gradient: sum over all output channels = over the ovec threads of this pixel (consecutive lanes)
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s0 = fmaf(gpv[j], wr[j][0], s0);
        s1 = fmaf(gpv[j], wr[j][1], s1);
        s2 = fmaf(gpv[j], wr[j][2], s2);
      }
      for (int o = ovec >> 1; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (live && cg == 0) {
        float* gxp = gx + b * 3 * plane + p;
        gxp[0] = s0; gxp[plane] = s1; gxp[2 * plane] = s2;
      }
    }
  }
  // weight / bias gradients: reduce the CTA's pixel rows in shared memory, then one atomic per value
  float* mine = fr_red + prow * (cout * 4) + cg * 32;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mine[j * 4 + 0] = aw[j][0]; mine[j * 4 + 1] = aw[j][1]; mine[j * 4 + 2] = aw[j][2]; mine[j * 4 + 3] = ab[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cout * 4; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += fr_red[r * (cout * 4) + i];
    const int o = i >> 2, c = i & 3;
    if (c < 3) atomicAdd(gw + o * 3 + c, s * wscale);
    else if (gb) atomicAdd(gb + o, s);
  }
}

static int from_rgb_check(int cout, int dtype, const char* what) {
  const int ovec = cout / 8;
  TE_CHECK_ARG(cout % 8 == 0 && ovec >= 1 && ovec <= 32 && (ovec & (ovec - 1)) == 0,
               "%s: the output channel count must be 8, 16, ..., 256 (a power of two), got %d", what, cout);
  TE_CHECK_ARG(dtype == TE_F32 || dtype == TE_BF16, "%s: activations must be f32 or bf16", what);
  return TE_OK;
}

}  // namespace te

extern "C" int te_from_rgb_fwd(void* y, const float* x, const float* w, const float* bias, int batch, int h, int wd,
                               int cout, float wscale, float slope, float gain, int dtype, void* stream) {
  using namespace te;
  if (int rc = from_rgb_check(cout, dtype, "from_rgb_fwd")) return rc;
  const int64_t total = int64_t(batch) * h * wd;
  if (total == 0) return TE_OK;
  TE_CHECK_ARG(y && x && w, "from_rgb_fwd: null pointer");
  TE_CHECK_ARG(batch <= 65535 && int64_t(h) * wd < (int64_t(1) << 31), "from_rgb_fwd: tensor too large");
  const int rows = 256 / (cout / 8), plane = h * wd;
  int gx_ = (kNumSMs * 8 + batch - 1) / batch;                 // ~8 waves of CTAs over the whole batch
  const int need = (plane + rows - 1) / rows;
  if (gx_ > need) gx_ = need;
  const dim3 grid(gx_ > 0 ? gx_ : 1, batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == TE_BF16)
    from_rgb_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<__nv_bfloat16*>(y), x, w, bias, plane, cout,
                                                              wscale, slope, gain);
  else
    from_rgb_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<float*>(y), x, w, bias, plane, cout, wscale, slope,
                                                      gain);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

extern "C" int te_from_rgb_bwd(float* gw, float* gbias, float* gx, const void* g, const void* out, const float* x,
                               const float* w, int batch, int h, int wd, int cout, float wscale, float slope,
                               float gain, int dtype, void* stream) {
  using namespace te;
  if (int rc = from_rgb_check(cout, dtype, "from_rgb_bwd")) return rc;
  const int64_t total = int64_t(batch) * h * wd;
  if (total == 0) return TE_OK;
  TE_CHECK_ARG(gw && g && out && x && w, "from_rgb_bwd: null pointer");
  TE_CHECK_ARG(batch <= 65535 && int64_t(h) * wd < (int64_t(1) << 31), "from_rgb_bwd: tensor too large");
  const int rows = 256 / (cout / 8), plane = h * wd;
  int gx_ = (kNumSMs * 4 + batch - 1) / batch;
  const int need = (plane + rows - 1) / rows;
  if (gx_ > need) gx_ = need;
  const dim3 grid(gx_ > 0 ? gx_ : 1, batch);
  const size_t smem = size_t(rows) * cout * 4 * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool cfg = false;
  if (!cfg) {
    TE_CHECK_CUDA(cudaFuncSetAttribute(from_rgb_bwd_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    TE_CHECK_CUDA(cudaFuncSetAttribute(from_rgb_bwd_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    TE_CHECK_CUDA(cudaFuncSetAttribute(from_rgb_bwd_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    TE_CHECK_CUDA(cudaFuncSetAttribute(from_rgb_bwd_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cfg = true;
  }
  if (dtype == TE_BF16) {
    const __nv_bfloat16 *gq = static_cast<const __nv_bfloat16*>(g), *oq = static_cast<const __nv_bfloat16*>(out);
    if (gx) from_rgb_bwd_kernel<__nv_bfloat16, true><<<grid, 256, smem, st>>>(gw, gbias, gx, gq, oq, x, w, plane, cout, wscale, slope, gain);
    else from_rgb_bwd_kernel<__nv_bfloat16, false><<<grid, 256, smem, st>>>(gw, gbias, gx, gq, oq, x, w, plane, cout, wscale, slope, gain);
  } else {
    const float *gq = static_cast<const float*>(g), *oq = static_cast<const float*>(out);
    if (gx) from_rgb_bwd_kernel<float, true><<<grid, 256, smem, st>>>(gw, gbias, gx, gq, oq, x, w, plane, cout, wscale, slope, gain);
    else from_rgb_bwd_kernel<float, false><<<grid, 256, smem, st>>>(gw, gbias, gx, gq, oq, x, w, plane, cout, wscale, slope, gain);
  }
  TE_CHECK_LAUNCH();
  return TE_OK;
}
