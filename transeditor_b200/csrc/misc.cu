// Error plumbing, fused Adam+EMA, cross-attention core.
#include <stdarg.h>

#include "common.cuh"

namespace te {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam semantics, no weight decay / amsgrad) + EMA of the updated parameter.
__global__ void __launch_bounds__(256)
adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                float* __restrict__ v, float* __restrict__ ema, int64_t n4, int64_t n, float lr,
                float b1, float b2, float eps, float bc1, float bc2_sqrt, float ema_decay,
                float gscale, const int* __restrict__ step_dev) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  if (step_dev) {  // CUDA-graph friendly: the step count lives on the device
    const float t = float(*step_dev);
    bc1 = 1.f - ((b1 == 0.f) ? 0.f : powf(b1, t));
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  const float step_size = lr / bc1;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; float* gg = &gv.x; float* mm = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gg[j] * gscale;
      mm[j] = b1 * mm[j] + (1.f - b1) * gr;
      vp[j] = b2 * vp[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(vp[j]) / bc2_sqrt + eps;
      pp[j] -= step_size * (mm[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (ema) {
      float4 ev = reinterpret_cast<float4*>(ema)[i];
      float* ee = &ev.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) ee[j] = ema_decay * ee[j] + (1.f - ema_decay) * pp[j];
      reinterpret_cast<float4*>(ema)[i] = ev;
    }
  }
  // scalar tail (n not a multiple of 4)
  for (int64_t i = n4 * 4 + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gr = g[i] * gscale;
    const float mn = b1 * m[i] + (1.f - b1) * gr;
    const float vn = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mn;
    v[i] = vn;
    const float pn = p[i] - step_size * (mn / (sqrtf(vn) / bc2_sqrt + eps));
    p[i] = pn;
    if (ema) ema[i] = ema_decay * ema[i] + (1.f - ema_decay) * pn;
  }
}

// ---------------------------------------------------------------------------------------------
// Cross-attention core, one CTA per sample: 16 query tokens (from P) x 16 key tokens (from Z),
// 4 heads x 32 channels.  Everything lives in shared memory; 128 threads.
constexpr int AT_T = 16, AT_C = 128, AT_H = 4, AT_D = 32;

__global__ void __launch_bounds__(128)
attn_core_kernel(float* __restrict__ out, float* __restrict__ sim_out, const float* __restrict__ q,
                 const float* __restrict__ k, const float* __restrict__ v, float scale) {
  __shared__ float sq[AT_T][AT_C + 1], sk[AT_T][AT_C + 1], sv[AT_T][AT_C + 1];
  __shared__ float ss[AT_H][AT_T][AT_T + 1];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t base = int64_t(b) * AT_T * AT_C;
  for (int e = tid; e < AT_T * AT_C; e += 128) {
    const int t = e / AT_C, c = e - t * AT_C;
    sq[t][c] = q[base + e];
    sk[t][c] = k[base + e];
    sv[t][c] = v[base + e];
  }
  __syncthreads();
  // logits: 4*16*16 = 1024 entries, 8 per thread
  for (int e = tid; e < AT_H * AT_T * AT_T; e += 128) {
    const int l = e % AT_T, m = (e / AT_T) % AT_T, h = e / (AT_T * AT_T);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < AT_D; ++c) acc += sq[m][h * AT_D + c] * sk[l][h * AT_D + c];
    ss[h][m][l] = acc * scale;
  }
  __syncthreads();
  // softmax over l: 64 rows, threads 0..63 one row each
  if (tid < AT_H * AT_T) {
    const int h = tid / AT_T, m = tid - h * AT_T;
    float mx = -INFINITY;
#pragma unroll
    for (int l = 0; l < AT_T; ++l) mx = fmaxf(mx, ss[h][m][l]);
    float e[AT_T], sum = 0.f;
#pragma unroll
    for (int l = 0; l < AT_T; ++l) { e[l] = expf(ss[h][m][l] - mx); sum += e[l]; }
    const float inv = 1.f / sum;
#pragma unroll
    for (int l = 0; l < AT_T; ++l) {
      const float pv = e[l] * inv;
      ss[h][m][l] = pv;
      if (sim_out) sim_out[((int64_t(b) * AT_H + h) * AT_T + m) * AT_T + l] = pv;
    }
  }
  __syncthreads();
  // out[b, m, h*32+c] = sum_l sim[h][m][l] * v[l][h*32+c]   (the reference's reshape/permute, :894)
  for (int e = tid; e < AT_T * AT_C; e += 128) {
    const int m = e / AT_C, hc = e - m * AT_C, h = hc / AT_D;
    float acc = 0.f;
#pragma unroll
    for (int l = 0; l < AT_T; ++l) acc += ss[h][m][l] * sv[l][hc];
    out[base + e] = acc;
  }
}

}  // namespace te

extern "C" int te_version(void) { return 1000; }
extern "C" const char* te_last_error(void) { return te::g_err; }

extern "C" int te_adam_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n,
                           float lr, float beta1, float beta2, float eps, int step, float ema_decay,
                           float grad_scale, void* stream) {
  using namespace te;
  if (n == 0) return TE_OK;
  TE_CHECK_ARG(p && g && m && v, "adam_ema: null pointer");
  TE_CHECK_ARG(step >= 1, "adam_ema: step must be >= 1");
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                       reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                       reinterpret_cast<uintptr_t>(ema);
  const int64_t n4 = (al & 15) == 0 ? n / 4 : 0;
  const double bc1 = 1.0 - pow(double(beta1), double(step));
  const double bc2 = 1.0 - pow(double(beta2), double(step));
  adam_ema_kernel<<<grid_for(n4 > 0 ? n4 : n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, ema, n4, n, lr, beta1, beta2, eps, float(bc1), float(sqrt(bc2)), ema_decay,
      grad_scale, nullptr);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

extern "C" int te_adam_ema_devstep(float* p, const float* g, float* m, float* v, float* ema, int64_t n,
                                   float lr, float beta1, float beta2, float eps, const int* step_dev,
                                   float ema_decay, float grad_scale, void* stream) {
  using namespace te;
  if (n == 0) return TE_OK;
  TE_CHECK_ARG(p && g && m && v && step_dev, "adam_ema_devstep: null pointer");
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                       reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                       reinterpret_cast<uintptr_t>(ema);
  const int64_t n4 = (al & 15) == 0 ? n / 4 : 0;
  adam_ema_kernel<<<grid_for(n4 > 0 ? n4 : n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, ema, n4, n, lr, beta1, beta2, eps, 1.f, 1.f, ema_decay, grad_scale, step_dev);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

extern "C" int te_attn_core(float* out, float* sim_out, const float* q, const float* k,
                            const float* v, int batch, int tokens, void* stream) {
  using namespace te;
  TE_CHECK_ARG(out && q && k && v, "attn_core: null pointer");
  TE_CHECK_ARG(tokens == AT_T, "attn_core: built for 16 x 16 tokens (got %d)", tokens);
  if (batch == 0) return TE_OK;
  attn_core_kernel<<<batch, 128, 0, static_cast<cudaStream_t>(stream)>>>(out, sim_out, q, k, v,
                                                                         rsqrtf(float(AT_C)));
  TE_CHECK_LAUNCH();
  return TE_OK;
}
