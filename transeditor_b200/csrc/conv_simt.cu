// Full-precision (f32 / f64) gather convolution and its weight gradient, NCHW, SIMT.
// This is the PARITY engine: it reproduces the reference's fp32 arithmetic to ~1e-6 without ever
// materialising per-sample weights (conv(x, W*s*d) == d * conv(x*s, W)).
// The throughput engine is conv_tc.cu (tcgen05, bf16).
//
// One geometry struct (te_conv_geom) covers conv2d, strided conv2d, conv_transpose2d and all of
// their data gradients; see include/te_b200.h.
#include "common.cuh"

namespace te {

constexpr int CS_BM = 64;    // output channels per CTA
constexpr int CS_BN = 128;   // output pixels per CTA (within one sample)
constexpr int CS_BK = 16;    // reduction chunk over flattened (cin, ky, kx)
constexpr int CS_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(CS_THREADS)
conv2d_gather_kernel(T* __restrict__ y, const T* __restrict__ x, const T* __restrict__ w,
                     const T* __restrict__ in_scale, const T* __restrict__ out_scale,
                     const T* __restrict__ bias, const T* __restrict__ noise,
                     const T* __restrict__ noise_w, te_conv_geom g) {
  constexpr int NBUF = sizeof(T) == 8 ? 1 : 2;  // f64 tiles would exceed 48 KB double-buffered
  __shared__ __align__(16) T As[NBUF][CS_BK][CS_BM + 4];
  __shared__ __align__(16) T Bs[NBUF][CS_BK][CS_BN + 4];

  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * CS_BM;
  const int n0 = blockIdx.x * CS_BN;
  const int npix = g.hout * g.wout;
  const int kk = g.kh * g.kw;
  const int ktot = g.cin * kk;

  // B-tile loader: thread owns pixel column `bn` and rows bk0 + 2*j
  const int bn = tid & (CS_BN - 1);
  const int bk0 = tid >> 7;  // 0..1
  const int pix = n0 + bn;
  const bool pix_ok = pix < npix;
  const int oy = pix_ok ? pix / g.wout : 0;
  const int ox = pix_ok ? pix - oy * g.wout : 0;
  const int base_y = oy * g.down - g.pad_y;
  const int base_x = ox * g.down - g.pad_x;
  const T* xb = x + int64_t(b) * g.cin * g.hin * g.win;

  // A-tile loader: thread owns k column `ak` and rows am0 + 16*j
  const int ak = tid & (CS_BK - 1);
  const int am0 = tid >> 4;  // 0..15

  T a_reg[4], b_reg[8];

  auto load_tiles = [&](int k0) {
    // weights
    {
      const int kf = k0 + ak;
      int ci = 0, tap = 0;
      const bool k_ok = kf < ktot;
      if (k_ok) {
        ci = kf / kk;
        tap = kf - ci * kk;
        if (g.flip) tap = kk - 1 - tap;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + am0 + 16 * j;
        T v = T(0);
        if (k_ok && m < g.cout) v = w[int64_t(m) * g.w_so + int64_t(ci) * g.w_si + tap];
        a_reg[j] = v;
      }
    }
    // gathered inputs
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kf = k0 + bk0 + 2 * j;
      T v = T(0);
      if (pix_ok && kf < ktot) {
        const int ci = kf / kk;
        const int tap = kf - ci * kk;
        const int ky = tap / g.kw;
        const int kx = tap - ky * g.kw;
        int ty = base_y + ky, tx = base_x + kx;
        bool ok = ty >= 0 && tx >= 0;
        if (g.up > 1) {
          ok = ok && (ty % g.up == 0) && (tx % g.up == 0);
          ty /= g.up;
          tx /= g.up;
        }
        if (ok && ty < g.hin && tx < g.win) {
          v = xb[(int64_t(ci) * g.hin + ty) * g.win + tx];
          if (in_scale) v *= in_scale[int64_t(b) * g.cin + ci];
        }
      }
      b_reg[j] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) As[buf][ak][am0 + 16 * j] = a_reg[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) Bs[buf][bk0 + 2 * j][bn] = b_reg[j];
  };

  // compute mapping: 16 x 16 threads, each 4 (M) x 8 (N)
  const int tx = tid & 15, ty = tid >> 4;
  T acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = T(0);

  const int nk = (ktot + CS_BK - 1) / CS_BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = (NBUF == 2) ? (kc & 1) : 0;
    if (kc + 1 < nk) load_tiles((kc + 1) * CS_BK);
#pragma unroll
    for (int k = 0; k < CS_BK; ++k) {
      T a[4], bb[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[buf][k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bb[j] = Bs[buf][k][tx * 4 + j];
        bb[4 + j] = Bs[buf][k][64 + tx * 4 + j];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += a[i] * bb[j];
    }
    if (kc + 1 < nk) {
      if (NBUF == 1) __syncthreads();
      store_tiles((NBUF == 2) ? (buf ^ 1) : 0);
      __syncthreads();
    }
  }

  // epilogue
  const T nw = (noise && noise_w) ? noise_w[0] : T(0);
  const T slope = T(0.2), gain = T(1.4142135623730951);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.cout) continue;
    const T osc = out_scale ? out_scale[int64_t(b) * g.cout + m] : T(1);
    const T bv = bias ? bias[m] : T(0);
    T* yrow = y + (int64_t(b) * g.cout + m) * npix;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= npix) continue;
      T v = acc[i][j] * osc;
      if (noise) v += nw * noise[int64_t(b) * g.noise_bstride + n];
      v += bv;
      if (g.act == 1) v = (v > T(0) ? v : v * slope) * gain;
      yrow[n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: GEMM  M = cout (64), N = flattened (cin,ky,kx) (64), K = pixels of all samples
constexpr int WG_BM = 64, WG_BN = 64, WG_BK = 32;

template <typename T>
__global__ void __launch_bounds__(256)
conv2d_wgrad_kernel(T* __restrict__ gw, const T* __restrict__ x, const T* __restrict__ gy,
                    const T* __restrict__ in_scale, const T* __restrict__ out_scale,
                    te_conv_geom g, int pix_chunks, int chunk_len) {
  __shared__ __align__(16) T As[WG_BK][WG_BM + 4];  // gy tile  [pixel][cout]
  __shared__ __align__(16) T Bs[WG_BK][WG_BN + 4];  // x gather [pixel][kflat]
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * WG_BM;
  const int n0 = blockIdx.x * WG_BN;
  const int b = blockIdx.z / pix_chunks;
  const int chunk = blockIdx.z - b * pix_chunks;
  const int npix = g.hout * g.wout;
  const int kk = g.kh * g.kw;
  const int ntot = g.cin * kk;
  const int p_lo = chunk * chunk_len;
  const int p_hi = min(npix, p_lo + chunk_len);

  const int lp = tid & 31;   // pixel within the K chunk
  const int l8 = tid >> 5;   // 0..7
  const T* xb = x + int64_t(b) * g.cin * g.hin * g.win;
  const T* gyb = gy + int64_t(b) * g.cout * npix;

  // decode the 8 kflat columns this thread gathers (fixed for the whole kernel)
  int col_ci[8], col_ky[8], col_kx[8];
  bool col_ok[8];
  T col_scale[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = n0 + l8 + 8 * j;
    col_ok[j] = n < ntot;
    const int ci = col_ok[j] ? n / kk : 0;
    const int tap = col_ok[j] ? n - ci * kk : 0;
    col_ci[j] = ci;
    col_ky[j] = tap / g.kw;
    col_kx[j] = tap - col_ky[j] * g.kw;
    col_scale[j] = in_scale ? in_scale[int64_t(b) * g.cin + ci] : T(1);
  }
  T row_scale[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int m = m0 + l8 + 8 * j;
    row_scale[j] = (out_scale && m < g.cout) ? out_scale[int64_t(b) * g.cout + m] : T(1);
  }

  const int tx = tid & 15, ty = tid >> 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

  for (int p0 = p_lo; p0 < p_hi; p0 += WG_BK) {
    const int pix = p0 + lp;
    const bool pix_ok = pix < p_hi;
    const int oy = pix_ok ? pix / g.wout : 0;
    const int ox = pix_ok ? pix - oy * g.wout : 0;
    const int base_y = oy * g.down - g.pad_y;
    const int base_x = ox * g.down - g.pad_x;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = m0 + l8 + 8 * j;
      T v = T(0);
      if (pix_ok && m < g.cout) v = gyb[int64_t(m) * npix + pix] * row_scale[j];
      As[lp][l8 + 8 * j] = v;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      T v = T(0);
      if (pix_ok && col_ok[j]) {
        int ty_ = base_y + col_ky[j], tx_ = base_x + col_kx[j];
        bool ok = ty_ >= 0 && tx_ >= 0;
        if (g.up > 1) {
          ok = ok && (ty_ % g.up == 0) && (tx_ % g.up == 0);
          ty_ /= g.up;
          tx_ /= g.up;
        }
        if (ok && ty_ < g.hin && tx_ < g.win)
          v = xb[(int64_t(col_ci[j]) * g.hin + ty_) * g.win + tx_] * col_scale[j];
      }
      Bs[lp][l8 + 8 * j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WG_BK; ++k) {
      T a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bb[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= ntot) continue;
      const int ci = n / kk;
      int tap = n - ci * kk;
      if (g.flip) tap = kk - 1 - tap;
      atomicAdd(gw + int64_t(m) * g.w_so + int64_t(ci) * g.w_si + tap, acc[i][j]);
    }
  }
}

static int check_geom(const te_conv_geom* g, const char* who) {
  TE_CHECK_ARG(g != nullptr, "%s: null geometry", who);
  TE_CHECK_ARG(g->batch > 0 && g->cin > 0 && g->cout > 0 && g->hin > 0 && g->win > 0 &&
                   g->hout > 0 && g->wout > 0,
               "%s: non-positive dimension", who);
  TE_CHECK_ARG(g->kh >= 1 && g->kw >= 1 && g->kh <= 7 && g->kw <= 7, "%s: kernel size must be 1..7", who);
  TE_CHECK_ARG(g->up >= 1 && g->down >= 1, "%s: up/down must be >= 1", who);
  TE_CHECK_ARG(g->batch <= 65535, "%s: batch > 65535", who);
  return TE_OK;
}

template <typename T>
static int conv2d_simt_typed(void* y, const void* x, const void* w, const void* isc, const void* osc,
                             const void* bias, const void* noise, const void* noise_w,
                             const te_conv_geom& g, cudaStream_t st) {
  const int npix = g.hout * g.wout;
  dim3 grid((npix + CS_BN - 1) / CS_BN, (g.cout + CS_BM - 1) / CS_BM, g.batch);
  conv2d_gather_kernel<T><<<grid, CS_THREADS, 0, st>>>(
      static_cast<T*>(y), static_cast<const T*>(x), static_cast<const T*>(w),
      static_cast<const T*>(isc), static_cast<const T*>(osc), static_cast<const T*>(bias),
      static_cast<const T*>(noise), static_cast<const T*>(noise_w), g);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

template <typename T>
static int conv2d_wgrad_typed(void* gw, const void* x, const void* gy, const void* isc,
                              const void* osc, const te_conv_geom& g, cudaStream_t st) {
  const int npix = g.hout * g.wout;
  const int ntot = g.cin * g.kh * g.kw;
  const int gx = (ntot + WG_BN - 1) / WG_BN, gyb = (g.cout + WG_BM - 1) / WG_BM;
  // split the pixel reduction so that the grid fills the chip (~4 waves), >= 256 pixels per chunk
  int64_t base_ctas = int64_t(gx) * gyb * g.batch;
  int chunks = int((int64_t(kNumSMs) * 4 + base_ctas - 1) / base_ctas);
  int max_chunks = (npix + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int chunk_len = (npix + chunks - 1) / chunks;
  chunk_len = (chunk_len + WG_BK - 1) / WG_BK * WG_BK;
  chunks = (npix + chunk_len - 1) / chunk_len;
  TE_CHECK_ARG(int64_t(g.batch) * chunks <= 65535, "conv2d_wgrad: grid.z too large");
  dim3 grid(gx, gyb, g.batch * chunks);
  conv2d_wgrad_kernel<T><<<grid, 256, 0, st>>>(static_cast<T*>(gw), static_cast<const T*>(x),
                                               static_cast<const T*>(gy), static_cast<const T*>(isc),
                                               static_cast<const T*>(osc), g, chunks, chunk_len);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

}  // namespace te

extern "C" int te_conv2d_simt(void* y, const void* x, const void* w, const void* in_scale,
                              const void* out_scale, const void* bias, const void* noise,
                              const void* noise_w, const te_conv_geom* g, int dtype, void* stream) {
  using namespace te;
  int rc = check_geom(g, "conv2d_simt");
  if (rc) return rc;
  TE_CHECK_ARG(y && x && w, "conv2d_simt: null tensor pointer");
  TE_CHECK_ARG(g->act == 0 || g->act == 1, "conv2d_simt: act must be 0 or 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == TE_F32) return conv2d_simt_typed<float>(y, x, w, in_scale, out_scale, bias, noise, noise_w, *g, st);
  if (dtype == TE_F64) return conv2d_simt_typed<double>(y, x, w, in_scale, out_scale, bias, noise, noise_w, *g, st);
  set_error("conv2d_simt: dtype must be f32 or f64 (got %d); bf16 runs on te_conv2d_tc", dtype);
  return TE_ERR_UNSUPPORTED;
}

extern "C" int te_conv2d_wgrad_simt(void* gw, const void* x, const void* gy, const void* in_scale,
                                    const void* out_scale, const te_conv_geom* g, int dtype,
                                    void* stream) {
  using namespace te;
  int rc = check_geom(g, "conv2d_wgrad_simt");
  if (rc) return rc;
  TE_CHECK_ARG(gw && x && gy, "conv2d_wgrad_simt: null tensor pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == TE_F32) return conv2d_wgrad_typed<float>(gw, x, gy, in_scale, out_scale, *g, st);
  if (dtype == TE_F64) return conv2d_wgrad_typed<double>(gw, x, gy, in_scale, out_scale, *g, st);
  set_error("conv2d_wgrad_simt: dtype must be f32 or f64 (got %d)", dtype);
  return TE_ERR_UNSUPPORTED;
}
