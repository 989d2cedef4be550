// Data formats either side of the path: decoded uint8 HWC pixels -> normalised network input (the tail of the
// reference's loader transform) and network output -> uint8 HWC pixels (torchvision's save_image arithmetic).
// Byte work, HBM-bound: 4 pixels per thread, 12-byte packed reads / 16-byte stores.
#include "common.cuh"

namespace te {

// ToTensor (x / 255) then Normalize(0.5, 0.5) ((t - 0.5) / 0.5), each step rounded to f32 like the torch ops.
__device__ __forceinline__ float prep_value(unsigned v) {
  const float t = __fdiv_rn(float(v), 255.0f);
  return __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);
}

template <typename T8>
__global__ void __launch_bounds__(256)
image_prep_kernel(float* __restrict__ nchw, T8* __restrict__ nhwc8, const uint8_t* __restrict__ src,
                  const uint8_t* __restrict__ flip, int batch, int h, int w) {
  const int64_t plane = int64_t(h) * w;
  if ((w & 3) == 0) {
    const int wq = w >> 2;
    const int64_t total = int64_t(batch) * h * wq;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += int64_t(gridDim.x) * blockDim.x) {
      const int xq = int(i % wq);
      const int64_t row = i / wq;  // b * h + y
      const int b = int(row / h);
      const bool fl = flip != nullptr && flip[b] != 0;
      const int x0 = xq * 4;                      // first OUTPUT pixel of this thread
      const int sx0 = fl ? (w - 4 - x0) : x0;     // first SOURCE pixel of the same 4-pixel group
      const uint32_t* sp = reinterpret_cast<const uint32_t*>(src + (row * w + sx0) * 3);
      const uint32_t w0 = __ldg(sp), w1 = __ldg(sp + 1), w2 = __ldg(sp + 2);
      unsigned px[4][3];
      px[0][0] = w0 & 255u; px[0][1] = (w0 >> 8) & 255u; px[0][2] = (w0 >> 16) & 255u;
      px[1][0] = w0 >> 24;  px[1][1] = w1 & 255u;        px[1][2] = (w1 >> 8) & 255u;
      px[2][0] = (w1 >> 16) & 255u; px[2][1] = w1 >> 24; px[2][2] = w2 & 255u;
      px[3][0] = (w2 >> 8) & 255u;  px[3][1] = (w2 >> 16) & 255u; px[3][2] = w2 >> 24;
      float v[4][3];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int sp_i = fl ? 3 - p : p;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[p][c] = prep_value(px[sp_i][c]);
      }
      const int y = int(row - int64_t(b) * h);
      if (nchw) {
        float* o = nchw + (int64_t(b) * 3) * plane + int64_t(y) * w + x0;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          *reinterpret_cast<float4*>(o + c * plane) = make_float4(v[0][c], v[1][c], v[2][c], v[3][c]);
      }
      if (nhwc8) {
        T8* o = nhwc8 + (row * w + x0) * 8;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          if constexpr (sizeof(T8) == 2) {
            Vec16<T8> q;
#pragma unroll
            for (int c = 0; c < 8; ++c) q.v[c] = from_acc<T8, float>(c < 3 ? v[p][c] : 0.f);
            *reinterpret_cast<uint4*>(o + p * 8) = q.raw;
          } else {
            *reinterpret_cast<float4*>(o + p * 8) = make_float4(v[p][0], v[p][1], v[p][2], 0.f);
            *reinterpret_cast<float4*>(o + p * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    }
    return;
  }
  // any width: one pixel per thread
  const int64_t total = int64_t(batch) * plane;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int x = int(i % w);
    const int64_t row = i / w;
    const int b = int(row / h);
    const int y = int(row - int64_t(b) * h);
    const bool fl = flip != nullptr && flip[b] != 0;
    const uint8_t* sp = src + (row * w + (fl ? w - 1 - x : x)) * 3;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = prep_value(sp[c]);
    if (nchw) {
#pragma unroll
      for (int c = 0; c < 3; ++c) nchw[(int64_t(b) * 3 + c) * plane + int64_t(y) * w + x] = v[c];
    }
    if (nhwc8) {
#pragma unroll
      for (int c = 0; c < 8; ++c) nhwc8[i * 8 + c] = from_acc<T8, float>(c < 3 ? v[c] : 0.f);
    }
  }
}

// torchvision.utils.save_image(normalize=True, range=(low, high)) per pixel:
//   clamp(x, low, high); (x - low) / max(high - low, 1e-5); * 255; + 0.5; clamp(0, 255); truncate to uint8
__device__ __forceinline__ uint8_t quant_value(float x, float low, float high, float den) {
  x = fminf(fmaxf(x, low), high);
  float t = __fdiv_rn(__fsub_rn(x, low), den);
  t = __fadd_rn(__fmul_rn(t, 255.0f), 0.5f);
  t = fminf(fmaxf(t, 0.f), 255.f);
  return uint8_t(int(t));
}

template <typename T>
__global__ void __launch_bounds__(256)
image_quantize_kernel(uint8_t* __restrict__ dst, const T* __restrict__ src, int batch, int h, int w, int64_t sb,
                      int64_t sc, int64_t sy, int64_t sx, float low, float high) {
  const float den = fmaxf(__fsub_rn(high, low), 1e-5f);
  const int64_t plane = int64_t(h) * w;
  const int64_t total = int64_t(batch) * plane;
  if ((w & 3) == 0 && sx == 1) {
    const int64_t total4 = total >> 2;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total4;
         i += int64_t(gridDim.x) * blockDim.x) {
      const int64_t pix = i * 4;
      const int x0 = int(pix % w);
      const int64_t row = pix / w;
      const int b = int(row / h);
      const int y = int(row - int64_t(b) * h);
      const T* sp = src + b * sb + y * sy + x0;
      uint8_t q[12];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int p = 0; p < 4; ++p) q[p * 3 + c] = quant_value(to_acc(sp[c * sc + p]), low, high, den);
      uint32_t* o = reinterpret_cast<uint32_t*>(dst + pix * 3);
      o[0] = q[0] | (q[1] << 8) | (q[2] << 16) | (uint32_t(q[3]) << 24);
      o[1] = q[4] | (q[5] << 8) | (q[6] << 16) | (uint32_t(q[7]) << 24);
      o[2] = q[8] | (q[9] << 8) | (q[10] << 16) | (uint32_t(q[11]) << 24);
    }
    return;
  }
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int x = int(i % w);
    const int64_t row = i / w;
    const int b = int(row / h);
    const int y = int(row - int64_t(b) * h);
    const T* sp = src + b * sb + y * sy + x * sx;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[i * 3 + c] = quant_value(to_acc(sp[c * sc]), low, high, den);
  }
}

}  // namespace te

extern "C" int te_image_prep(float* dst_nchw, void* dst_nhwc8, const uint8_t* src_hwc, const uint8_t* flip,
                             int batch, int h, int w, int nhwc_dtype, void* stream) {
  using namespace te;
  TE_CHECK_ARG(batch >= 0 && h >= 0 && w >= 0, "te_image_prep: negative size");
  TE_CHECK_ARG(nhwc_dtype == TE_F32 || nhwc_dtype == TE_BF16, "te_image_prep: nhwc dtype must be f32 or bf16");
  const int64_t n = int64_t(batch) * h * w;
  if (n == 0) return TE_OK;   // empty batch: nothing to do (the pointers of empty tensors are NULL)
  TE_CHECK_ARG(src_hwc != nullptr, "te_image_prep: src is NULL");
  TE_CHECK_ARG(dst_nchw != nullptr || dst_nhwc8 != nullptr, "te_image_prep: no destination");
  const int64_t items = (w & 3) == 0 ? n / 4 : n;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(items, 256, 8);
  if (nhwc_dtype == TE_BF16)
    image_prep_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(dst_nchw, static_cast<__nv_bfloat16*>(dst_nhwc8), src_hwc,
                                                            flip, batch, h, w);
  else
    image_prep_kernel<float><<<grid, 256, 0, st>>>(dst_nchw, static_cast<float*>(dst_nhwc8), src_hwc, flip, batch, h,
                                                    w);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

extern "C" int te_image_quantize(uint8_t* dst_hwc, const void* src, int batch, int h, int w, int64_t stride_b,
                                 int64_t stride_c, int64_t stride_y, int64_t stride_x, float low, float high,
                                 int dtype, void* stream) {
  using namespace te;
  TE_CHECK_ARG(batch >= 0 && h >= 0 && w >= 0, "te_image_quantize: negative size");
  TE_CHECK_ARG(dtype == TE_F32 || dtype == TE_BF16, "te_image_quantize: dtype must be f32 or bf16");
  const int64_t n = int64_t(batch) * h * w;
  if (n == 0) return TE_OK;
  TE_CHECK_ARG(dst_hwc != nullptr && src != nullptr, "te_image_quantize: NULL pointer");
  const bool vec = (w & 3) == 0 && stride_x == 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(vec ? n / 4 : n, 256, 8);
  if (dtype == TE_BF16)
    image_quantize_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(dst_hwc, static_cast<const __nv_bfloat16*>(src), batch,
                                                                h, w, stride_b, stride_c, stride_y, stride_x, low, high);
  else
    image_quantize_kernel<float><<<grid, 256, 0, st>>>(dst_hwc, static_cast<const float*>(src), batch, h, w, stride_b,
                                                        stride_c, stride_y, stride_x, low, high);
  TE_CHECK_LAUNCH();
  return TE_OK;
}
