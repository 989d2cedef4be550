// Grouped EqualLinear: every small-M linear layer of the path — the 2 x 16 per-column mapping linears with their
// PixelNorm, the 20 style modulations, adjust_style, the discriminator's final linears — and their data / weight
// gradients, many layers per launch.
//
// Shape of the problem: M = batch (1..64 rows), K, N up to 512 (8192 for D's first final linear): every weight is
// used by M rows only, so the layers are bound by STREAMING THE WEIGHTS from HBM/L2 (arithmetic intensity M/2
// FLOP/byte), not by math.  Four warps share one 8-column output tile, a quarter of the reduction each (enough 16-byte
// loads in flight: 27 -> 13 us per launch, the 16.8 MB mapping launch at ~1.3 TB/s): the B operand is
// read straight from global memory as 16-byte vectors (a lane's four k values of one weight row: every sector fully
// used), the 16-row A tile sits in shared memory, products on mma.sync.m16n8k8 TF32 — 3 x TF32 (hi*hi + hi*lo + lo*hi,
// ~1e-6 relative) in the fp32 parity mode, single-pass TF32 otherwise — with the k index permuted inside each 16-wide
// slice so that fragments are contiguous float4s.  No cuBLAS, no per-layer launches, no stacked weight copies.
#include "common.cuh"

namespace te {

constexpr int LIN_MAX_TASKS = 32;
constexpr int LIN_KC = 512;              // reduction chunk staged in shared memory
constexpr int LIN_LDA = LIN_KC + 16;     // padded row: conflict-free float4 fragment reads
constexpr int LIN_KS = 4;                // warps sharing one 8-column tile: each walks a quarter of the reduction
constexpr int LIN_TILES = 4;             // 8-column tiles per CTA (32 output columns)
constexpr int LIN_WARPS = LIN_KS * LIN_TILES;   // 16 warps: enough 16-byte weight loads in flight to stream from HBM
constexpr int LIN_COLS = 8 * LIN_TILES;

struct LinTaskDev {
  te_linear_task t;
  int first_block;   // prefix sum of work items
  int n_chunks, m_groups, k_splits;
};
struct LinParams {
  LinTaskDev task[LIN_MAX_TASKS];
  int n_tasks;
  int x3;
};

__device__ __forceinline__ uint32_t lin_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void lin_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool X3>
__device__ __forceinline__ void lin_mma_step(float (&acc)[4], const float (&a)[4], float b0, float b1) {
  uint32_t ah[4], b0h = lin_tf32(b0), b1h = lin_tf32(b1);
#pragma unroll
  for (int i = 0; i < 4; ++i) ah[i] = lin_tf32(a[i]);
  if constexpr (X3) {
    uint32_t al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) al[i] = lin_tf32(a[i] - __uint_as_float(ah[i]));
    const uint32_t b0l = lin_tf32(b0 - __uint_as_float(b0h)), b1l = lin_tf32(b1 - __uint_as_float(b1h));
    lin_mma(acc, al, b0h, b1h);
    lin_mma(acc, ah, b0l, b1l);
  }
  lin_mma(acc, ah, b0h, b1h);
}

// y[m, n] = act( alpha * rnorm[m] * SUM_k x[m, k] * B(k, n) + bias[n] * bias_mul )
//   w_trans == 0:  B(k, n) = w[n * w_ld + k]   (forward: the reduction runs along a weight row)
//   w_trans == 1:  B(k, n) = w[k * w_ld + n]   (data gradient: the reduction runs down a weight column)
// CTA = 16 warps = 4 column tiles x 4 reduction quarters; the quarters' partial fragments meet in shared memory.
template <bool X3>
__global__ void __launch_bounds__(LIN_WARPS * 32) linear_grouped_kernel(const __grid_constant__ LinParams P) {
  extern __shared__ float lin_smem[];
  float* As = lin_smem;                       // [16][LIN_LDA]
  float* rnorm = As + 16 * LIN_LDA;           // [16]
  float* part = rnorm + 16;                   // [LIN_WARPS][32 lanes][4]
  int ti = 0;
  while (ti + 1 < P.n_tasks && int(blockIdx.x) >= P.task[ti + 1].first_block) ++ti;
  const LinTaskDev& T = P.task[ti];
  const te_linear_task& t = T.t;
  int item = blockIdx.x - T.first_block;
  const int chunk = item % T.n_chunks; item /= T.n_chunks;
  const int mg = item % T.m_groups;
  const int ks = item / T.m_groups;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
  const int tile = warp / LIN_KS, kq = warp % LIN_KS;
  const int m0 = mg * 16;
  const int n0 = chunk * LIN_COLS + tile * 8;
  // this CTA's share of the reduction (whole 16-slices)
  const int k16 = (t.k + 15) >> 4;
  const int per = (k16 + T.k_splits - 1) / T.k_splits;
  const int kb = ks * per * 16, ke = min(t.k, (ks + 1) * per * 16);
  const bool vec_b = !t.w_trans && (t.w_ld & 3) == 0 && (t.k & 3) == 0 && (reinterpret_cast<uintptr_t>(t.w) & 15) == 0;
  const bool tile_live = n0 < t.n;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float ssq = 0.f;  // pixel norm: this thread's share of SUM_k x^2 for row `warp`

  for (int kc = kb; kc < ke; kc += LIN_KC) {
    const int klen = min(LIN_KC, ke - kc);
    const int kpad = (klen + 15) & ~15;
    __syncthreads();
    // stage x[m0 .. m0+16, kc .. kc+klen) (zero rows / columns beyond the task's extent): one warp per row
    {
      const bool row_ok = m0 + warp < t.m;
      const float* src = t.x + int64_t(m0 + warp) * t.x_rs;
      for (int k = lane; k < kpad; k += 32) {
        const float v = (row_ok && k < klen) ? __ldg(src + int64_t(kc + k) * t.x_cs) : 0.f;
        As[warp * LIN_LDA + k] = v;
        ssq += v * v;
      }
    }
    __syncthreads();
    if (tile_live) {
      const float* a0p = As + g * LIN_LDA + 4 * tq;
      const float* a1p = As + (g + 8) * LIN_LDA + 4 * tq;
      const int n = n0 + g;
      const bool col_ok = n < t.n;
      // this warp's quarter of the chunk's 16-wide slices
      const int nsl = kpad >> 4, qsl = (nsl + LIN_KS - 1) / LIN_KS;
      const int s_lo = kq * qsl * 16, s_hi = min(kpad, (kq + 1) * qsl * 16);
#pragma unroll 8
      for (int s = s_lo; s < s_hi; s += 16) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok) {
          const int k = kc + s + 4 * tq;
          if (vec_b) {
            if (s + 4 * tq < klen) bv = __ldg(reinterpret_cast<const float4*>(t.w + int64_t(n) * t.w_ld + k));
          } else if (!t.w_trans) {
            const float* wp = t.w + int64_t(n) * t.w_ld + k;
            const int left = klen - (s + 4 * tq);
            if (left > 0) bv.x = __ldg(wp);
            if (left > 1) bv.y = __ldg(wp + 1);
            if (left > 2) bv.z = __ldg(wp + 2);
            if (left > 3) bv.w = __ldg(wp + 3);
          } else {
            const float* wp = t.w + int64_t(k) * t.w_ld + n;
            const int left = klen - (s + 4 * tq);
            if (left > 0) bv.x = __ldg(wp);
            if (left > 1) bv.y = __ldg(wp + t.w_ld);
            if (left > 2) bv.z = __ldg(wp + 2 * t.w_ld);
            if (left > 3) bv.w = __ldg(wp + 3 * t.w_ld);
          }
        }
        const float4 r0 = *reinterpret_cast<const float4*>(a0p + s);
        const float4 r1 = *reinterpret_cast<const float4*>(a1p + s);
        // hardware k = tq, tq+4  <->  columns 4tq, 4tq+1 (first step) and 4tq+2, 4tq+3 (second step)
        const float a_lo[4] = {r0.x, r1.x, r0.y, r1.y};
        const float a_hi[4] = {r0.z, r1.z, r0.w, r1.w};
        lin_mma_step<X3>(acc, a_lo, bv.x, bv.y);
        lin_mma_step<X3>(acc, a_hi, bv.z, bv.w);
      }
    }
  }
  if (t.pixel_norm) {   // row `warp` was staged by this warp: reduce its squares across the lanes
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
    if (lane == 0) {
      const float rn = rsqrtf(ssq / float(t.k) + 1e-8f);
      rnorm[warp] = rn;
      const int m = m0 + warp;
      if (t.rnorm_out && chunk == 0 && m < t.m) t.rnorm_out[m] = rn;
    }
  }
  // the reduction quarters of a tile meet in shared memory
  *reinterpret_cast<float4*>(part + (warp * 32 + lane) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  __syncthreads();
  if (!tile_live || kq != 0) return;
#pragma unroll
  for (int q = 1; q < LIN_KS; ++q) {
    const float4 o = *reinterpret_cast<const float4*>(part + ((warp + q) * 32 + lane) * 4);
    acc[0] += o.x; acc[1] += o.y; acc[2] += o.z; acc[3] += o.w;
  }
  // accumulator fragment: acc[0], acc[1] = row g, columns 2tq, 2tq+1; acc[2], acc[3] = row g+8
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = m0 + g + 8 * h;
    if (m >= t.m) continue;
    const float rn = t.pixel_norm ? rnorm[g + 8 * h] : 1.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + 2 * tq + j;
      if (n >= t.n) continue;
      float v = acc[2 * h + j] * t.alpha * rn;
      float* dst = t.y + int64_t(m) * t.y_rs + int64_t(n) * t.y_cs;
      if (T.k_splits > 1) {
        if (ks == 0 && t.bias) v += __ldg(t.bias + n) * t.bias_mul;
        atomicAdd(dst, v);
      } else {
        if (t.bias) v += __ldg(t.bias + n) * t.bias_mul;
        if (t.act == 1) v = (v > 0.f ? v : 0.2f * v) * 1.4142135623730951f;
        *dst = v;
      }
    }
  }
}

// gw[n, k] = alpha * SUM_m g[m, n] * x[m, k]   (+ gbias[n] = bias_mul * SUM_m g[m, n]);  rows of x optionally
// multiplied by the PixelNorm factor.  Output-bound (one store per weight): 32 n x 128 k per CTA, float4 stores.
struct LinWgTaskDev {
  te_linear_wgrad_task t;
  int first_block;
  int n_tiles, k_tiles;
};
struct LinWgParams {
  LinWgTaskDev task[LIN_MAX_TASKS];
  int n_tasks;
};
constexpr int LWG_TN = 32, LWG_TK = 128, LWG_MAXM = 64;

__global__ void __launch_bounds__(256) linear_wgrad_grouped_kernel(const __grid_constant__ LinWgParams P) {
  __shared__ float Gs[LWG_MAXM][LWG_TN + 1];
  __shared__ __align__(16) float Xs[LWG_MAXM][LWG_TK];
  int ti = 0;
  while (ti + 1 < P.n_tasks && int(blockIdx.x) >= P.task[ti + 1].first_block) ++ti;
  const LinWgTaskDev& T = P.task[ti];
  const te_linear_wgrad_task& t = T.t;
  const int item = blockIdx.x - T.first_block;
  const int nt = item % T.n_tiles, kt = item / T.n_tiles;
  const int n0 = nt * LWG_TN, k0 = kt * LWG_TK;
  const int tn = threadIdx.x >> 5, tk = threadIdx.x & 31;   // thread: n = n0 + tn + 8 i (i < 4), k = k0 + 4 tk .. +3
  float acc[4][4] = {};
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int mb = 0; mb < t.m; mb += LWG_MAXM) {
    const int mlen = min(LWG_MAXM, t.m - mb);
    __syncthreads();
    for (int i = threadIdx.x; i < mlen * LWG_TN; i += 256) {
      const int m = i / LWG_TN, n = i % LWG_TN;
      Gs[m][n] = n0 + n < t.n ? __ldg(t.g + int64_t(mb + m) * t.g_rs + int64_t(n0 + n) * t.g_cs) : 0.f;
    }
    for (int i = threadIdx.x; i < mlen * LWG_TK; i += 256) {
      const int m = i / LWG_TK, k = i % LWG_TK;
      const float sc = t.x_scale ? __ldg(t.x_scale + mb + m) : 1.f;
      Xs[m][k] = k0 + k < t.k ? sc * __ldg(t.x + int64_t(mb + m) * t.x_rs + int64_t(k0 + k) * t.x_cs) : 0.f;
    }
    __syncthreads();
    for (int m = 0; m < mlen; ++m) {
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[m][4 * tk]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float gv = Gs[m][tn + 8 * i];
        acc[i][0] += gv * xv.x; acc[i][1] += gv * xv.y; acc[i][2] += gv * xv.z; acc[i][3] += gv * xv.w;
        bsum[i] += gv;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tn + 8 * i;
    if (n >= t.n) continue;
    const int k = k0 + 4 * tk;
    float* dst = t.gw + int64_t(n) * t.k + k;
    if (k + 3 < t.k && (t.k & 3) == 0 && (reinterpret_cast<uintptr_t>(t.gw) & 15) == 0) {
      *reinterpret_cast<float4*>(dst) =
          make_float4(acc[i][0] * t.alpha, acc[i][1] * t.alpha, acc[i][2] * t.alpha, acc[i][3] * t.alpha);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k + j < t.k) dst[j] = acc[i][j] * t.alpha;
    }
    if (kt == 0 && tk == 0 && t.gbias) t.gbias[n] = bsum[i] * t.bias_mul;
  }
}

}  // namespace te

extern "C" int te_linear_grouped(const te_linear_task* tasks, int n_tasks, int precision, void* stream) {
  using namespace te;
  TE_CHECK_ARG(n_tasks >= 0 && (tasks != nullptr || n_tasks == 0), "linear_grouped: null task table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool configured = false;
  const int smem = (16 * LIN_LDA + 16 + LIN_WARPS * 32 * 4) * 4;
  if (!configured) {
    TE_CHECK_CUDA(cudaFuncSetAttribute(linear_grouped_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TE_CHECK_CUDA(cudaFuncSetAttribute(linear_grouped_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  for (int base = 0; base < n_tasks; base += LIN_MAX_TASKS) {
    LinParams P;
    P.n_tasks = 0;
    P.x3 = precision == 0;
    int blocks = 0;
    for (int i = base; i < n_tasks && i < base + LIN_MAX_TASKS; ++i) {
      const te_linear_task& t = tasks[i];
      TE_CHECK_ARG(t.m >= 0 && t.n >= 0 && t.k >= 0, "linear_grouped: negative extent in task %d", i);
      if (t.m == 0 || t.n == 0) continue;
      TE_CHECK_ARG(t.x && t.w && t.y, "linear_grouped: null pointer in task %d", i);
      TE_CHECK_ARG(t.act == 0 || t.act == 1, "linear_grouped: act must be 0 or 1 (task %d)", i);
      TE_CHECK_ARG(t.k_splits >= 1, "linear_grouped: k_splits must be >= 1 (task %d)", i);
      TE_CHECK_ARG(t.k_splits == 1 || (t.act == 0 && !t.pixel_norm),
                   "linear_grouped: split-K accumulates into y (zeroed by the caller): no activation / pixel norm");
      TE_CHECK_ARG(!t.pixel_norm || t.k <= LIN_KC, "linear_grouped: pixel norm needs K <= %d", LIN_KC);
      LinTaskDev& d = P.task[P.n_tasks++];
      d.t = t;
      d.first_block = blocks;
      d.n_chunks = (t.n + LIN_COLS - 1) / LIN_COLS;
      d.m_groups = (t.m + 15) / 16;
      d.k_splits = t.k_splits;
      blocks += d.n_chunks * d.m_groups * d.k_splits;
    }
    if (blocks == 0) continue;
    if (P.x3)
      linear_grouped_kernel<true><<<blocks, LIN_WARPS * 32, smem, st>>>(P);
    else
      linear_grouped_kernel<false><<<blocks, LIN_WARPS * 32, smem, st>>>(P);
    TE_CHECK_LAUNCH();
  }
  return TE_OK;
}

extern "C" int te_linear_wgrad_grouped(const te_linear_wgrad_task* tasks, int n_tasks, void* stream) {
  using namespace te;
  TE_CHECK_ARG(n_tasks >= 0 && (tasks != nullptr || n_tasks == 0), "linear_wgrad_grouped: null task table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int base = 0; base < n_tasks; base += LIN_MAX_TASKS) {
    LinWgParams P;
    P.n_tasks = 0;
    int blocks = 0;
    for (int i = base; i < n_tasks && i < base + LIN_MAX_TASKS; ++i) {
      const te_linear_wgrad_task& t = tasks[i];
      TE_CHECK_ARG(t.m >= 0 && t.n >= 0 && t.k >= 0, "linear_wgrad_grouped: negative extent in task %d", i);
      if (t.n == 0 || t.k == 0) continue;
      TE_CHECK_ARG(t.gw && (t.m == 0 || (t.g && t.x)), "linear_wgrad_grouped: null pointer in task %d", i);
      LinWgTaskDev& d = P.task[P.n_tasks++];
      d.t = t;
      d.first_block = blocks;
      d.n_tiles = (t.n + LWG_TN - 1) / LWG_TN;
      d.k_tiles = (t.k + LWG_TK - 1) / LWG_TK;
      blocks += d.n_tiles * d.k_tiles;
    }
    if (blocks == 0) continue;
    linear_wgrad_grouped_kernel<<<blocks, 256, 0, st>>>(P);
    TE_CHECK_LAUNCH();
  }
  return TE_OK;
}
