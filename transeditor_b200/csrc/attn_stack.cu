// The dual-space interaction network (8 cross-attention blocks) as ONE forward launch and TWO backward
// launches.  Math: model_spatial_query.py:883-901 (Attention), :920-936 (AttentionBlock), :194-221 (EqualLinear).
//
// Shape of the problem: 16 tokens x 512 channels per sample, samples independent, 3 MB of f32 weights per
// block.  Every layer is a [16 x K] x [K x N] product — far too small for tcgen05 (M >= 64) and, launched as
// ~35 library kernels per block, pure launch latency.  Here one thread-block CLUSTER of 8 CTAs
// owns a sample for the whole stack:
//   * a layer's N output columns are dealt to the 8 CTAs in 8-column tiles (tile = rank + 8 i), so each CTA
//     streams 1/8 of the weights, straight from L2 into mma.sync B fragments (32-byte sectors fully used);
//   * inside a CTA the 8 warps split K; partial tiles are summed through shared memory;
//   * the epilogue (scale, bias, residual, GELU, ...) writes its columns into the SAME activation buffer of
//     all 8 CTAs through distributed shared memory, one cluster barrier per layer (4 per block);
//   * the M = 16 rows are exactly one m16n8k8 MMA tile; products are 3xTF32 (hi*hi + hi*lo + lo*hi with f32
//     accumulation), i.e. fp32-equivalent (~1e-6), so the same kernel serves the parity mode and the bench mode;
//   * layer norms (over a whole 16 x D sample), the 16x16x4-head softmax attention and their backward forms
//     are computed redundantly by every CTA from its full copy: cheaper than another exchange.
// Backward: the same structure walks the blocks in reverse for the data gradients and parks the per-layer
// output gradients in a workspace; the weight gradients of ALL layers and blocks (reduction over the B*16
// rows) are one grouped f32 GEMM launch afterwards.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace te {

constexpr int AS_T = 16;        // tokens (queries from P, keys/values from Z)
constexpr int AS_C = 512;       // model width
constexpr int AS_PL = 128;      // attention planes = 4 heads x 32
constexpr int AS_H = 4, AS_HD = 32;
constexpr int AS_NT = 512;      // threads per CTA
constexpr int AS_NW = AS_NT / 32;
constexpr int AS_MAXD = 528;    // widest input (512 + 16 one-hot token columns)
constexpr int AS_LD = 532;      // smem row stride of the wide buffers (16-byte rows, 532 % 32 == 20)
constexpr int AS_LDS = 132;     // smem row stride of the 128-wide buffers
constexpr int AS_MAXB = 8;      // most blocks per launch
constexpr int AS_WIDE = AS_T * AS_LD;     // floats of one wide buffer
constexpr int AS_NARROW = AS_T * AS_LDS;  // floats of one 128-wide buffer
constexpr int AS_PART = 64 * 128;         // partial C fragments of one chunk: (k-groups x tiles) = 64
constexpr int AS_GPACC = (AS_C / 8 / 4) * 128;        // a CTA's columns of grad p at cluster size 4
constexpr float AS_ATT_SCALE = 0.08838834764831845f;  // 128^-0.5 (:873)
constexpr float AS_LN_EPS = 1e-5f;

struct AsParams {
  te_attn_block blk[AS_MAXB];
  int n_blocks, batch;
  float lr_mul;
  const float *x0, *p0, *p;
  float* y;
  float* save;
};

struct AsBwdParams {
  te_attn_block blk[AS_MAXB];
  int n_blocks, batch;
  float lr_mul;
  const float* gy;
  const float* save;
  float* gws;
  float *g_x0, *g_p0, *g_p;
};

// What the forward keeps per block, each a contiguous [B*16, width] matrix (so the weight-gradient GEMM can
// read it as an operand), followed by the two layer-norm reciprocal standard deviations per sample.
struct AsSaveOff {
  int64_t xn, q, k, v, att, ln1, u, h, rstd, total;
};
__host__ __device__ inline AsSaveOff as_save_off(int batch, int din) {
  const int64_t r = int64_t(batch) * AS_T;
  AsSaveOff o;
  int64_t c = 0;
  o.xn = c;   c += r * din;
  o.q = c;    c += r * AS_PL;
  o.k = c;    c += r * AS_PL;
  o.v = c;    c += r * AS_PL;
  o.att = c;  c += r * AS_PL;
  o.ln1 = c;  c += r * AS_C;
  o.u = c;    c += r * AS_C;
  o.h = c;    c += r * AS_C;
  o.rstd = c; c += (2 * int64_t(batch) + 3) / 4 * 4;
  o.total = c;
  return o;
}
// Gradient workspace per block: gout | gu | gx1 | gq | gk | gv, each [B*16, width].
struct AsGwsOff {
  int64_t gout, gu, gx1, gq, gk, gv, total;
};
__host__ __device__ inline AsGwsOff as_gws_off(int batch) {
  const int64_t r = int64_t(batch) * AS_T;
  AsGwsOff o;
  o.gout = 0;
  o.gu = r * AS_C;
  o.gx1 = 2 * r * AS_C;
  o.gq = 3 * r * AS_C;
  o.gk = o.gq + r * AS_PL;
  o.gv = o.gk + r * AS_PL;
  o.total = o.gv + r * AS_PL;
  return o;
}

// ------------------------------------------------------------------------------------------------ MMA core
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragments of one 16-wide slice of K = two m16n8k8 steps.  X3: value = hi + lo, both TF32.
template <bool X3>
struct AFrag {
  uint32_t hi[2][4];
  uint32_t lo[X3 ? 2 : 1][4];
};

template <bool X3>
__device__ __forceinline__ void set_a(AFrag<X3>& f, int step, int e, float v) {
  const uint32_t h = to_tf32(v);
  f.hi[step][e] = h;
  if constexpr (X3) f.lo[step][e] = to_tf32(v - __uint_as_float(h));
}

template <bool X3>
__device__ __forceinline__ void mma_step(float (&acc)[4], const AFrag<X3>& f, int step, float b0, float b1) {
  const uint32_t b0h = to_tf32(b0), b1h = to_tf32(b1);
  if constexpr (X3) {
    const uint32_t b0l = to_tf32(b0 - __uint_as_float(b0h)), b1l = to_tf32(b1 - __uint_as_float(b1h));
    mma_tf32(acc, f.lo[step], b0h, b1h);
    mma_tf32(acc, f.hi[step], b0l, b1l);
  }
  mma_tf32(acc, f.hi[step], b0h, b1h);
}

// One layer for this CTA:  D[16, N] = A[16, K] * B[K, N]  restricted to the 8-column tiles rank, rank+CL, ...
//   KCONTIG:  B(k, n) = W[n*ld + k]   (forward: W is [out, in], the reduction runs along a weight row)
//   else:     B(k, n) = W[k*ld + n]   (data gradients: the reduction runs down a weight column)
// K is cut into 16-wide slices dealt to KW = min(16, K/16) warp groups; the 16 / KW remaining warp groups take
// different tiles.  Inside a slice the MMA's k index is PERMUTED (hardware k = t, t+4 <-> columns 4t, 4t+1 in the
// first step and 4t+2, 4t+3 in the second), so a lane's A values are one float4 per row and — KCONTIG — its B
// values one float4 of a weight row: 16-byte loads, every 32-byte sector fully used in both forms.
// Tiles are processed in chunks of 64 / KW; a chunk's partial fragments are summed through `part`, then
// epi(row, col, v0, v1, slot) runs for every pair of adjacent columns; slot enumerates the CTA's pairs.
template <int CL, bool X3, bool KCONTIG, int KS, typename Epi>
__device__ __forceinline__ void gemm16_impl(const float* __restrict__ A, int lda, const float* __restrict__ W, int ld,
                                            int K, int N, int rank, float* part, Epi epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int k16 = K >> 4;
  const int KW = k16 >= 16 ? 16 : (k16 >= 8 ? 8 : 4);
  const int TW = AS_NW / KW, CH = 64 / KW;
  const int kw = warp & (KW - 1), tw = warp / KW;

  AFrag<X3> af[KS];
#pragma unroll
  for (int j = 0; j < KS; ++j) {
    const int s = kw + KW * j;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
    if (s < k16) {
      r0 = *reinterpret_cast<const float4*>(A + g * lda + 16 * s + 4 * t);
      r1 = *reinterpret_cast<const float4*>(A + (g + 8) * lda + 16 * s + 4 * t);
    }
    set_a<X3>(af[j], 0, 0, r0.x); set_a<X3>(af[j], 0, 1, r1.x); set_a<X3>(af[j], 0, 2, r0.y); set_a<X3>(af[j], 0, 3, r1.y);
    set_a<X3>(af[j], 1, 0, r0.z); set_a<X3>(af[j], 1, 1, r1.z); set_a<X3>(af[j], 1, 2, r0.w); set_a<X3>(af[j], 1, 3, r1.w);
  }

  const int ntiles = N >> 3;
  const int mine = (ntiles - rank + CL - 1) / CL;
  constexpr int TB = KS == 3 ? 2 : 4;  // tiles whose weights are in flight together (register budget)
  float4 bw[TB][KS];
  // weights of tiles i0 .. i0+TB-1 of the chunk at c0 (this warp's share) -> registers
  auto load_b = [&](int c0, int i0) {
#pragma unroll
    for (int i = 0; i < TB; ++i) {
      const int li = c0 + tw + TW * (i0 + i);
      const int n = (rank + CL * li) * 8 + g;
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const int s = kw + KW * j;
        bw[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (li < mine && s < k16) {
          if constexpr (KCONTIG) {
            bw[i][j] = __ldg(reinterpret_cast<const float4*>(W + n * ld + 16 * s + 4 * t));
          } else {
            const float* wp = W + (16 * s + 4 * t) * ld + n;
            bw[i][j] = make_float4(__ldg(wp), __ldg(wp + ld), __ldg(wp + 2 * ld), __ldg(wp + 3 * ld));
          }
        }
      }
    }
  };
  auto mma_b = [&](int c0, int i0) {
#pragma unroll
    for (int i = 0; i < TB; ++i) {
      const int ti = tw + TW * (i0 + i);
      if (c0 + ti < mine) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < KS; ++j) {
          if (kw + KW * j < k16) {
            mma_step<X3>(acc, af[j], 0, bw[i][j].x, bw[i][j].y);
            mma_step<X3>(acc, af[j], 1, bw[i][j].z, bw[i][j].w);
          }
        }
        *reinterpret_cast<float4*>(part + (kw * CH + ti) * 128 + lane * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      }
    }
  };
  if (mine > 0) load_b(0, 0);
  for (int c0 = 0; c0 < mine; c0 += CH) {
#pragma unroll
    for (int i0 = 0; i0 < 4; i0 += TB) {
      if (i0 > 0) load_b(c0, i0);
      mma_b(c0, i0);
    }
    // the NEXT chunk's weights travel while this chunk is reduced and its epilogue runs (its registers are free now)
    if (c0 + CH < mine) load_b(c0 + CH, 0);
    __syncthreads();
    const int nvalid = mine - c0 < CH ? mine - c0 : CH;
    for (int idx = threadIdx.x; idx < nvalid * 64; idx += AS_NT) {
      const int ti = idx >> 6, r = idx & 63, lp = r >> 1, half = r & 1;
      float v0 = 0.f, v1 = 0.f;
      for (int q = 0; q < KW; ++q) {
        const float2 pv = *reinterpret_cast<const float2*>(part + (q * CH + ti) * 128 + lp * 4 + half * 2);
        v0 += pv.x;
        v1 += pv.y;
      }
      const int li = c0 + ti;
      epi((lp >> 2) + 8 * half, (rank + CL * li) * 8 + (lp & 3) * 2, v0, v1, li * 64 + r);
    }
    __syncthreads();
  }
}

template <int CL, bool X3, bool KCONTIG, typename Epi>
__device__ __forceinline__ void gemm16(const float* A, int lda, const float* W, int ld, int K, int N, int rank,
                                       float* part, Epi epi) {
  if (K > 512) {  // only block 0's 528-wide inputs need a third slice per warp group
    gemm16_impl<CL, X3, KCONTIG, 3>(A, lda, W, ld, K, N, rank, part, epi);
  } else {
    gemm16_impl<CL, X3, KCONTIG, 2>(A, lda, W, ld, K, N, rank, part, epi);
  }
}

// Same two values into the same place of all CTAs' shared memory.
template <int CL>
__device__ __forceinline__ void bcast2(cg::cluster_group& cl, float* buf, int off, float v0, float v1) {
  const float2 v = make_float2(v0, v1);
#pragma unroll
  for (int c = 0; c < CL; ++c) *reinterpret_cast<float2*>(cl.map_shared_rank(buf, c) + off) = v;
}

// Sum over the CTA, same value (same summation order) in every thread.
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < AS_NW; ++w) t += red[w];
  return t;
}

constexpr int AS_PER = (AS_MAXD + AS_NT - 1) / AS_NT;  // columns per thread and row (2)

// F.layer_norm(x, x.shape[1:]) of one sample held in smem [16][AS_LD]; returns 1/std.
__device__ __forceinline__ float layer_norm_sample(const float* X, float* XN, int d, float* red) {
  float x[AS_T][AS_PER];
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < AS_T; ++r)
#pragma unroll
    for (int j = 0; j < AS_PER; ++j) {
      const int c = threadIdx.x + j * AS_NT;
      x[r][j] = c < d ? X[r * AS_LD + c] : 0.f;
      s += x[r][j];
    }
  const float inv_n = 1.f / float(AS_T * d);
  const float mean = block_sum(s, red) * inv_n;
  float q = 0.f;
#pragma unroll
  for (int r = 0; r < AS_T; ++r)
#pragma unroll
    for (int j = 0; j < AS_PER; ++j) {
      const float dx = threadIdx.x + j * AS_NT < d ? x[r][j] - mean : 0.f;
      q += dx * dx;
    }
  const float rstd = rsqrtf(block_sum(q, red) * inv_n + AS_LN_EPS);
#pragma unroll
  for (int r = 0; r < AS_T; ++r)
#pragma unroll
    for (int j = 0; j < AS_PER; ++j) {
      const int c = threadIdx.x + j * AS_NT;
      if (c < d) XN[r * AS_LD + c] = (x[r][j] - mean) * rstd;
    }
  return rstd;
}

// Layer-norm backward of one sample:  out = res + rstd (G - mean(G) - xhat mean(G xhat)).  xhat ([16][d], global
// memory) is fetched ONCE into registers with every load in flight together — the cluster barrier in front of
// this step has just invalidated L1, so each access is an L2 round trip.
__device__ __forceinline__ void ln_bwd_sample(const float* G, const float* __restrict__ xhat, const float* res,
                                              float* out, int d, float rstd, float* red) {
  float xh[AS_T][AS_PER];
#pragma unroll
  for (int r = 0; r < AS_T; ++r)
#pragma unroll
    for (int j = 0; j < AS_PER; ++j) {
      const int c = threadIdx.x + j * AS_NT;
      xh[r][j] = c < d ? __ldg(xhat + r * d + c) : 0.f;
    }
  float a = 0.f, c2 = 0.f;
#pragma unroll
  for (int r = 0; r < AS_T; ++r)
#pragma unroll
    for (int j = 0; j < AS_PER; ++j) {
      const int c = threadIdx.x + j * AS_NT;
      if (c < d) {
        const float gv = G[r * AS_LD + c];
        a += gv;
        c2 += gv * xh[r][j];
      }
    }
  const float inv_n = 1.f / float(AS_T * d);
  const float m1 = block_sum(a, red) * inv_n;
  const float m2 = block_sum(c2, red) * inv_n;
#pragma unroll
  for (int r = 0; r < AS_T; ++r)
#pragma unroll
    for (int j = 0; j < AS_PER; ++j) {
      const int c = threadIdx.x + j * AS_NT;
      if (c < d) out[r * AS_LD + c] = res[r * AS_LD + c] + rstd * (G[r * AS_LD + c] - m1 - xh[r][j] * m2);
    }
}

// This CTA's share of the 16 rows (16 / CL consecutive rows) of a smem buffer -> the sample's [16][d] matrix in
// global memory (coalesced); the CTAs of the cluster hold identical copies.
template <int CL>
__device__ __forceinline__ void store_rows(float* dst, const float* src, int ld, int d, int rank) {
  constexpr int ROWS = AS_T / CL;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int row = ROWS * rank + r;
    for (int c = threadIdx.x; c < d; c += AS_NT) dst[row * d + c] = src[row * ld + c];
  }
}
__device__ __forceinline__ void load_sample(float* dst, int ld, const float* __restrict__ src, int d) {
  for (int c = threadIdx.x; c < d; c += AS_NT) {
    float v[AS_T];
#pragma unroll
    for (int r = 0; r < AS_T; ++r) v[r] = __ldg(src + r * d + c);  // 16 loads in flight
#pragma unroll
    for (int r = 0; r < AS_T; ++r) dst[r * ld + c] = v[r];
  }
}

// sim[h][m][l] = softmax_l(q[m].k[l] * scale) into S ([4][16][17]); all buffers in smem, ld = AS_LDS.
__device__ __forceinline__ void attn_similarity(const float* Q, const float* K, float* S) {
  for (int e = threadIdx.x; e < AS_H * AS_T * AS_T; e += AS_NT) {
    const int l = e & 15, m = (e >> 4) & 15, h = e >> 8;
    const float* q = Q + m * AS_LDS + h * AS_HD;
    const float* k = K + l * AS_LDS + h * AS_HD;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < AS_HD; ++c) acc += q[c] * k[c];
    S[(h * AS_T + m) * 17 + l] = acc * AS_ATT_SCALE;
  }
  __syncthreads();
  if (threadIdx.x < AS_H * AS_T) {
    float* row = S + threadIdx.x * 17;
    float mx = -INFINITY;
#pragma unroll
    for (int l = 0; l < AS_T; ++l) mx = fmaxf(mx, row[l]);
    float e[AS_T], sum = 0.f;
#pragma unroll
    for (int l = 0; l < AS_T; ++l) {
      e[l] = expf(row[l] - mx);
      sum += e[l];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int l = 0; l < AS_T; ++l) row[l] = e[l] * inv;
  }
  __syncthreads();
}

__device__ __forceinline__ float gelu_erf(float u) { return 0.5f * u * (1.f + erff(u * 0.7071067811865476f)); }
__device__ __forceinline__ float gelu_erf_grad(float u) {
  return 0.5f * (1.f + erff(u * 0.7071067811865476f)) + u * 0.3989422804014327f * expf(-0.5f * u * u);
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int AS_NBIAS = 3 * AS_PL + 4 * AS_C;  // b_q | b_k | b_v | b_o | b_m1 | b_m2 | b_proj of one block
constexpr int AS_FWD_SMEM = (4 * AS_WIDE + AS_PART + 64 + AS_NBIAS) * 4;

template <int CL, bool X3>
__global__ void __launch_bounds__(AS_NT, 1) attn_stack_fwd_kernel(const __grid_constant__ AsParams P) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = int(cl.block_rank());
  const int b = blockIdx.x / CL;
  extern __shared__ __align__(16) float sm[];
  float* X = sm;                 // block input (full copy)
  float* XN = X + AS_WIDE;       // layer_norm(x), later layer_norm(x1)
  float* X1 = XN + AS_WIDE;      // x after the attention residual
  float* HB = X1 + AS_WIDE;      // gelu(mlp[0]); q | k | v | att live here before it
  float* PART = HB + AS_WIDE;
  float* RED = PART + AS_PART;
  float* BIAS = RED + 64;        // this block's biases times lr_mul (fetched while the first layer norm runs)
  float* Qs = HB;
  float* Ks = Qs + AS_NARROW;
  float* Vs = Ks + AS_NARROW;
  float* ATT = Vs + AS_NARROW;
  const float lr = P.lr_mul;
  const float* Bq = BIAS;
  const float* Bk = Bq + AS_PL;
  const float* Bv = Bk + AS_PL;
  const float* Bo = Bv + AS_PL;
  const float* Bm1 = Bo + AS_C;
  const float* Bm2 = Bm1 + AS_C;
  const float* Bp = Bm2 + AS_C;

  load_sample(X, AS_LD, P.x0 + int64_t(b) * AS_T * P.blk[0].in_dim, P.blk[0].in_dim);
  __syncthreads();

  int64_t save_base = 0;
  for (int blk = 0; blk < P.n_blocks; ++blk) {
    const te_attn_block& W = P.blk[blk];
    const int din = W.in_dim, dp = W.param_dim;
    const bool has_proj = W.w_proj != nullptr;
    const bool last = blk + 1 == P.n_blocks;
    const float* Pm = (blk == 0 ? P.p0 : P.p) + int64_t(b) * AS_T * dp;
    const AsSaveOff so = as_save_off(P.batch, din);
    float* sv = P.save ? P.save + save_base : nullptr;
    const int64_t row0 = int64_t(b) * AS_T;
    const float s_in = lr * rsqrtf(float(din)), s_p = lr * rsqrtf(float(dp));
    const float s_pl = lr * rsqrtf(float(AS_PL)), s_c = lr * rsqrtf(float(AS_C));

    // ---- biases: loads issued now, parked in shared memory after the layer norm (their L2 round trip is hidden)
    constexpr int NB_PER = (AS_NBIAS + AS_NT - 1) / AS_NT;
    float bias_reg[NB_PER];
#pragma unroll
    for (int j = 0; j < NB_PER; ++j) {
      const int e = threadIdx.x + j * AS_NT;
      const float* src = nullptr;
      if (e < AS_PL) src = W.b_q + e;
      else if (e < 2 * AS_PL) src = W.b_k + (e - AS_PL);
      else if (e < 3 * AS_PL) src = W.b_v + (e - 2 * AS_PL);
      else if (e < 3 * AS_PL + AS_C) src = W.b_o + (e - 3 * AS_PL);
      else if (e < 3 * AS_PL + 2 * AS_C) src = W.b_m1 + (e - 3 * AS_PL - AS_C);
      else if (e < 3 * AS_PL + 3 * AS_C) src = W.b_m2 + (e - 3 * AS_PL - 2 * AS_C);
      else if (e < AS_NBIAS && has_proj) src = W.b_proj + (e - 3 * AS_PL - 3 * AS_C);
      bias_reg[j] = src ? __ldg(src) * lr : 0.f;
    }

    // ---- layer norm of the block input
    const float rstd0 = layer_norm_sample(X, XN, din, RED);
#pragma unroll
    for (int j = 0; j < NB_PER; ++j) {
      const int e = threadIdx.x + j * AS_NT;
      if (e < AS_NBIAS) BIAS[e] = bias_reg[j];
    }
    __syncthreads();
    if (sv) {
      store_rows<CL>(sv + so.xn + row0 * din, XN, AS_LD, din, rank);
      if (rank == 0 && threadIdx.x == 0) sv[so.rstd + 2 * b] = rstd0;
    }

    // ---- q from the P tokens, k and v from the normalised Z tokens -> all CTAs
    gemm16<CL, X3, true>(Pm, dp, W.w_q, dp, dp, AS_PL, rank, PART, [&](int row, int col, float v0, float v1, int) {
      bcast2<CL>(cl, Qs, row * AS_LDS + col, v0 * s_p + Bq[col], v1 * s_p + Bq[col + 1]);
    });
    gemm16<CL, X3, true>(XN, AS_LD, W.w_k, din, din, AS_PL, rank, PART, [&](int row, int col, float v0, float v1, int) {
      bcast2<CL>(cl, Ks, row * AS_LDS + col, v0 * s_in + Bk[col], v1 * s_in + Bk[col + 1]);
    });
    gemm16<CL, X3, true>(XN, AS_LD, W.w_v, din, din, AS_PL, rank, PART, [&](int row, int col, float v0, float v1, int) {
      bcast2<CL>(cl, Vs, row * AS_LDS + col, v0 * s_in + Bv[col], v1 * s_in + Bv[col + 1]);
    });
    cl.sync();

    // ---- attention (every CTA, from its full copy)
    if (sv) {
      store_rows<CL>(sv + so.q + row0 * AS_PL, Qs, AS_LDS, AS_PL, rank);
      store_rows<CL>(sv + so.k + row0 * AS_PL, Ks, AS_LDS, AS_PL, rank);
      store_rows<CL>(sv + so.v + row0 * AS_PL, Vs, AS_LDS, AS_PL, rank);
    }
    attn_similarity(Qs, Ks, PART);
    for (int e = threadIdx.x; e < AS_T * AS_PL; e += AS_NT) {
      const int m = e >> 7, hc = e & 127, h = hc >> 5;
      const float* srow = PART + (h * AS_T + m) * 17;
      float acc = 0.f;
#pragma unroll
      for (int l = 0; l < AS_T; ++l) acc += srow[l] * Vs[l * AS_LDS + hc];
      ATT[m * AS_LDS + hc] = acc;
    }
    __syncthreads();
    if (sv) store_rows<CL>(sv + so.att + row0 * AS_PL, ATT, AS_LDS, AS_PL, rank);

    // ---- x1 = (proj(x) | x) + attention output projection -> all CTAs
    if (has_proj) {
      gemm16<CL, X3, true>(ATT, AS_LDS, W.w_o, AS_PL, AS_PL, AS_C, rank, PART, [&](int row, int col, float v0, float v1, int) {
        *reinterpret_cast<float2*>(X1 + row * AS_LD + col) =
            make_float2(v0 * s_pl + Bo[col], v1 * s_pl + Bo[col + 1]);
      });
      gemm16<CL, X3, true>(X, AS_LD, W.w_proj, din, din, AS_C, rank, PART, [&](int row, int col, float v0, float v1, int) {
        const float2 a = *reinterpret_cast<const float2*>(X1 + row * AS_LD + col);
        bcast2<CL>(cl, X1, row * AS_LD + col, v0 * s_in + Bp[col] + a.x,
                   v1 * s_in + Bp[col + 1] + a.y);
      });
    } else {
      gemm16<CL, X3, true>(ATT, AS_LDS, W.w_o, AS_PL, AS_PL, AS_C, rank, PART, [&](int row, int col, float v0, float v1, int) {
        const float2 x = *reinterpret_cast<const float2*>(X + row * AS_LD + col);
        bcast2<CL>(cl, X1, row * AS_LD + col, v0 * s_pl + Bo[col] + x.x, v1 * s_pl + Bo[col + 1] + x.y);
      });
    }
    cl.sync();

    // ---- layer norm of x1
    const float rstd1 = layer_norm_sample(X1, XN, AS_C, RED);
    __syncthreads();
    if (sv) {
      store_rows<CL>(sv + so.ln1 + row0 * AS_C, XN, AS_LD, AS_C, rank);
      if (rank == 0 && threadIdx.x == 0) sv[so.rstd + 2 * b + 1] = rstd1;
    }

    // ---- h = gelu(mlp[0](ln1)) -> all CTAs
    gemm16<CL, X3, true>(XN, AS_LD, W.w_m1, AS_C, AS_C, AS_C, rank, PART, [&](int row, int col, float v0, float v1, int) {
      const float u0 = v0 * s_c + Bm1[col], u1 = v1 * s_c + Bm1[col + 1];
      const float h0 = gelu_erf(u0), h1 = gelu_erf(u1);
      if (sv) {
        *reinterpret_cast<float2*>(sv + so.u + (row0 + row) * AS_C + col) = make_float2(u0, u1);
        *reinterpret_cast<float2*>(sv + so.h + (row0 + row) * AS_C + col) = make_float2(h0, h1);
      }
      bcast2<CL>(cl, HB, row * AS_LD + col, h0, h1);
    });
    cl.sync();

    // ---- x2 = x1 + mlp[2](h) -> next block's input in all CTAs
    gemm16<CL, X3, true>(HB, AS_LD, W.w_m2, AS_C, AS_C, AS_C, rank, PART, [&](int row, int col, float v0, float v1, int) {
      const float2 x = *reinterpret_cast<const float2*>(X1 + row * AS_LD + col);
      const float y0 = v0 * s_c + Bm2[col] + x.x, y1 = v1 * s_c + Bm2[col + 1] + x.y;
      if (last) *reinterpret_cast<float2*>(P.y + (row0 + row) * AS_C + col) = make_float2(y0, y1);
      bcast2<CL>(cl, X, row * AS_LD + col, y0, y1);
    });
    cl.sync();
    save_base += so.total;
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int AS_BWD_SMEM = (4 * AS_WIDE + 4 * AS_NARROW + AS_PART + AS_GPACC + 64) * 4;

template <int CL, bool X3>
__global__ void __launch_bounds__(AS_NT, 1) attn_stack_bwd_kernel(const __grid_constant__ AsBwdParams P) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = int(cl.block_rank());
  const int b = blockIdx.x / CL;
  extern __shared__ __align__(16) float sm[];
  float* G2 = sm;                  // gradient of the block output (full copy); proj gradient in block 0
  float* GU = G2 + AS_WIDE;        // gradient of the pre-GELU activations; later q | k | v
  float* GL = GU + AS_WIDE;        // gradient of layer_norm(x1); later of layer_norm(x)
  float* GX1 = GL + AS_WIDE;       // gradient of x1
  float* GATT = GX1 + AS_WIDE;
  float* GQ = GATT + AS_NARROW;
  float* GK = GQ + AS_NARROW;
  float* GV = GK + AS_NARROW;
  float* PART = GV + AS_NARROW;
  float* GPACC = PART + AS_PART;   // this CTA's columns of the gradient of p, summed over blocks >= 1
  float* RED = GPACC + AS_GPACC;
  float* Qs = GU;
  float* Ks = Qs + AS_NARROW;
  float* Vs = Ks + AS_NARROW;
  float* S = PART;                 // similarity, then logit gradients
  const float lr = P.lr_mul;
  const int64_t row0 = int64_t(b) * AS_T;
  const AsGwsOff go = as_gws_off(P.batch);

  load_sample(G2, AS_LD, P.gy + row0 * AS_C, AS_C);
  for (int e = threadIdx.x; e < AS_GPACC; e += AS_NT) GPACC[e] = 0.f;
  __syncthreads();

  int64_t save_base = 0;
  for (int blk = 0; blk < P.n_blocks; ++blk) save_base += as_save_off(P.batch, P.blk[blk].in_dim).total;

  for (int blk = P.n_blocks - 1; blk >= 0; --blk) {
    const te_attn_block& W = P.blk[blk];
    const int din = W.in_dim, dp = W.param_dim;
    const bool has_proj = W.w_proj != nullptr;
    const AsSaveOff so = as_save_off(P.batch, din);
    save_base -= so.total;
    const float* sv = P.save + save_base;
    float* gw = P.gws + int64_t(blk) * go.total;
    const float s_in = lr * rsqrtf(float(din)), s_p = lr * rsqrtf(float(dp));
    const float s_pl = lr * rsqrtf(float(AS_PL)), s_c = lr * rsqrtf(float(AS_C));

    // ---- through mlp[2] and the GELU: gu = (g2 W2 s) * gelu'(u)
    gemm16<CL, X3, false>(G2, AS_LD, W.w_m2, AS_C, AS_C, AS_C, rank, PART, [&](int row, int col, float v0, float v1, int) {
      const float2 u = *reinterpret_cast<const float2*>(sv + so.u + (row0 + row) * AS_C + col);
      const float g0 = v0 * s_c * gelu_erf_grad(u.x), g1 = v1 * s_c * gelu_erf_grad(u.y);
      *reinterpret_cast<float2*>(gw + go.gu + (row0 + row) * AS_C + col) = make_float2(g0, g1);
      bcast2<CL>(cl, GU, row * AS_LD + col, g0, g1);
    });
    cl.sync();

    // ---- through mlp[0]
    gemm16<CL, X3, false>(GU, AS_LD, W.w_m1, AS_C, AS_C, AS_C, rank, PART, [&](int row, int col, float v0, float v1, int) {
      bcast2<CL>(cl, GL, row * AS_LD + col, v0 * s_c, v1 * s_c);
    });
    cl.sync();

    // ---- layer-norm backward at x1, plus the residual
    ln_bwd_sample(GL, sv + so.ln1 + row0 * AS_C, G2, GX1, AS_C, sv[so.rstd + 2 * b + 1], RED);
    __syncthreads();
    store_rows<CL>(gw + go.gx1 + row0 * AS_C, GX1, AS_LD, AS_C, rank);

    // ---- through the attention output projection
    gemm16<CL, X3, false>(GX1, AS_LD, W.w_o, AS_PL, AS_C, AS_PL, rank, PART, [&](int row, int col, float v0, float v1, int) {
      bcast2<CL>(cl, GATT, row * AS_LDS + col, v0 * s_pl, v1 * s_pl);
    });
    cl.sync();

    // ---- attention backward (every CTA, full copies)
    load_sample(Qs, AS_LDS, sv + so.q + row0 * AS_PL, AS_PL);
    load_sample(Ks, AS_LDS, sv + so.k + row0 * AS_PL, AS_PL);
    load_sample(Vs, AS_LDS, sv + so.v + row0 * AS_PL, AS_PL);
    __syncthreads();
    attn_similarity(Qs, Ks, S);
    // gv[l][hc] = sum_m sim[h][m][l] gatt[m][hc]
    for (int e = threadIdx.x; e < AS_T * AS_PL; e += AS_NT) {
      const int l = e >> 7, hc = e & 127, h = hc >> 5;
      float acc = 0.f;
#pragma unroll
      for (int m = 0; m < AS_T; ++m) acc += S[(h * AS_T + m) * 17 + l] * GATT[m * AS_LDS + hc];
      GV[l * AS_LDS + hc] = acc;
    }
    // gsim[h][m][l] = sum_c gatt[m][h32+c] v[l][h32+c]; glogit = sim (gsim - sum_l gsim sim) scale.  The 16 l's
    // of a row are 16 adjacent lanes.
    constexpr int NE = AS_H * AS_T * AS_T / AS_NT;
    float gl[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int e = threadIdx.x + i * AS_NT;
      const int l = e & 15, m = (e >> 4) & 15, h = e >> 8;
      const float* ga = GATT + m * AS_LDS + h * AS_HD;
      const float* vv = Vs + l * AS_LDS + h * AS_HD;
      float gs = 0.f;
#pragma unroll
      for (int c = 0; c < AS_HD; ++c) gs += ga[c] * vv[c];
      const float sim = S[(h * AS_T + m) * 17 + l];
      float dot = gs * sim;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      gl[i] = sim * (gs - dot) * AS_ATT_SCALE;
    }
    __syncthreads();  // every thread has read its similarities; overwrite them with the logit gradients
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int e = threadIdx.x + i * AS_NT;
      const int l = e & 15, m = (e >> 4) & 15, h = e >> 8;
      S[(h * AS_T + m) * 17 + l] = gl[i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < AS_T * AS_PL; e += AS_NT) {
      const int r = e >> 7, hc = e & 127, h = hc >> 5;
      float aq = 0.f, ak = 0.f;
#pragma unroll
      for (int j = 0; j < AS_T; ++j) {
        aq += S[(h * AS_T + r) * 17 + j] * Ks[j * AS_LDS + hc];  // gq[m=r] = sum_l glogit[m][l] k[l]
        ak += S[(h * AS_T + j) * 17 + r] * Qs[j * AS_LDS + hc];  // gk[l=r] = sum_m glogit[m][l] q[m]
      }
      GQ[r * AS_LDS + hc] = aq;
      GK[r * AS_LDS + hc] = ak;
    }
    __syncthreads();
    store_rows<CL>(gw + go.gq + row0 * AS_PL, GQ, AS_LDS, AS_PL, rank);
    store_rows<CL>(gw + go.gk + row0 * AS_PL, GK, AS_LDS, AS_PL, rank);
    store_rows<CL>(gw + go.gv + row0 * AS_PL, GV, AS_LDS, AS_PL, rank);

    // ---- through the k / v transforms (into layer_norm(x)), the q transform (into p) and block 0's proj
    gemm16<CL, X3, false>(GK, AS_LDS, W.w_k, din, AS_PL, din, rank, PART, [&](int row, int col, float v0, float v1, int) {
      *reinterpret_cast<float2*>(GL + row * AS_LD + col) = make_float2(v0 * s_in, v1 * s_in);
    });
    gemm16<CL, X3, false>(GV, AS_LDS, W.w_v, din, AS_PL, din, rank, PART, [&](int row, int col, float v0, float v1, int) {
      const float2 a = *reinterpret_cast<const float2*>(GL + row * AS_LD + col);
      bcast2<CL>(cl, GL, row * AS_LD + col, a.x + v0 * s_in, a.y + v1 * s_in);
    });
    gemm16<CL, X3, false>(GQ, AS_LDS, W.w_q, dp, AS_PL, dp, rank, PART, [&](int row, int col, float v0, float v1, int slot) {
      if (blk == 0) {
        if (P.g_p0) *reinterpret_cast<float2*>(P.g_p0 + (row0 + row) * dp + col) = make_float2(v0 * s_p, v1 * s_p);
      } else {
        GPACC[2 * slot] += v0 * s_p;
        GPACC[2 * slot + 1] += v1 * s_p;
      }
    });
    if (has_proj) {
      gemm16<CL, X3, false>(GX1, AS_LD, W.w_proj, din, AS_C, din, rank, PART, [&](int row, int col, float v0, float v1, int) {
        bcast2<CL>(cl, G2, row * AS_LD + col, v0 * s_in, v1 * s_in);
      });
    }
    cl.sync();

    // ---- layer-norm backward at the block input, plus the residual (or the proj gradient)
    ln_bwd_sample(GL, sv + so.xn + row0 * din, has_proj ? G2 : GX1, G2, din, sv[so.rstd + 2 * b], RED);
    __syncthreads();
    if (blk > 0) {
      store_rows<CL>(P.gws + int64_t(blk - 1) * go.total + go.gout + row0 * AS_C, G2, AS_LD, AS_C, rank);
    } else if (P.g_x0) {
      store_rows<CL>(P.g_x0 + row0 * din, G2, AS_LD, din, rank);
    }
    // The next block's first exchange writes GU, which no CTA reads any more; its first cl.sync() is the barrier
    // in front of the next writes into GL / G2 (read just above).
  }
  if (P.g_p && P.n_blocks > 1) {
    // slot -> (row, col) exactly as gemm16 enumerates the pairs of an N = 512 layer
    for (int idx = threadIdx.x; idx < (AS_C / 8 / CL) * 64; idx += AS_NT) {
      const int li = idx >> 6, r = idx & 63, lp = r >> 1, half = r & 1;
      const int row = (lp >> 2) + 8 * half, col = (rank + CL * li) * 8 + (lp & 3) * 2;
      *reinterpret_cast<float2*>(P.g_p + (row0 + row) * AS_C + col) = make_float2(GPACC[2 * idx], GPACC[2 * idx + 1]);
    }
  }
  cl.sync();  // nobody leaves while a peer may still write into its shared memory
}

// ------------------------------------------------------------------------------------------------ weight gradients
// gW[n][k] = scale * sum_r G[r][n] X[r][k],  gb[n] = lr_mul * sum_r G[r][n]  for a list of layers; one CTA per
// 64 x 64 tile, f32 FMAs, the B*16 rows streamed through shared memory in chunks of 16.
struct AsWgTask {
  const float* G;
  const float* X;
  float* gW;
  float* gb;
  int N, K;
  float scale;
  int tile0;
};
struct AsWgParams {
  AsWgTask t[AS_MAXB * 7];
  int n_tasks, rows, total_tiles;
  float lr_mul;
};

__global__ void __launch_bounds__(256) attn_stack_wgrad_kernel(const __grid_constant__ AsWgParams P) {
  __shared__ __align__(16) float Gs[16][64];
  __shared__ __align__(16) float Xs[16][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr_ = tid >> 4, lc = (tid & 15) * 4;  // loader: row and first column of this thread's float4
  int j = 0;
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    while (j + 1 < P.n_tasks && tile >= P.t[j + 1].tile0) ++j;
    const AsWgTask& T = P.t[j];
    const int ktiles = (T.K + 63) >> 6;
    const int local = tile - T.tile0;
    const int n0 = (local / ktiles) * 64, k0 = (local % ktiles) * 64;
    const bool k_ok = k0 + lc < T.K;  // K is a multiple of 4: a float4 is inside or outside as a whole
    const bool want_bias = k0 == 0 && tx == 0 && T.gb != nullptr;
    float acc[4][4] = {};
    float bs[4] = {0.f, 0.f, 0.f, 0.f};
    const float* gp = T.G + int64_t(lr_) * T.N + n0 + lc;
    const float* xp = T.X + int64_t(lr_) * T.K + k0 + lc;
    float4 gv = *reinterpret_cast<const float4*>(gp);
    float4 xv = k_ok ? *reinterpret_cast<const float4*>(xp) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < P.rows; r0 += 16) {
      __syncthreads();
      *reinterpret_cast<float4*>(&Gs[lr_][lc]) = gv;
      *reinterpret_cast<float4*>(&Xs[lr_][lc]) = xv;
      __syncthreads();
      if (r0 + 16 < P.rows) {
        gv = *reinterpret_cast<const float4*>(gp + int64_t(r0 + 16) * T.N);
        if (k_ok) xv = *reinterpret_cast<const float4*>(xp + int64_t(r0 + 16) * T.K);
      }
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float4 g4 = *reinterpret_cast<const float4*>(&Gs[r][ty * 4]);
        const float4 x4 = *reinterpret_cast<const float4*>(&Xs[r][tx * 4]);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
        const float x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] += g[a] * x[c];
          bs[a] += g[a];
        }
      }
    }
    if (k0 + tx * 4 < T.K) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
        *reinterpret_cast<float4*>(T.gW + int64_t(n0 + ty * 4 + a) * T.K + k0 + tx * 4) =
            make_float4(acc[a][0] * T.scale, acc[a][1] * T.scale, acc[a][2] * T.scale, acc[a][3] * T.scale);
    }
    if (want_bias) {
#pragma unroll
      for (int a = 0; a < 4; ++a) T.gb[n0 + ty * 4 + a] = bs[a] * P.lr_mul;
    }
  }
}

static int as_validate(const te_attn_block* blocks, int n_blocks, int batch, int precision) {
  TE_CHECK_ARG(blocks, "attn_stack: null block table");
  TE_CHECK_ARG(n_blocks >= 1 && n_blocks <= AS_MAXB, "attn_stack: n_blocks %d outside [1, %d]", n_blocks, AS_MAXB);
  TE_CHECK_ARG(batch >= 0, "attn_stack: negative batch");
  TE_CHECK_ARG(precision == 0 || precision == 1, "attn_stack: precision %d (0 = 3xTF32, 1 = TF32)", precision);
  for (int i = 0; i < n_blocks; ++i) {
    const te_attn_block& w = blocks[i];
    TE_CHECK_ARG(w.in_dim % 16 == 0 && w.in_dim >= 64 && w.in_dim <= AS_MAXD,
                 "attn_stack: block %d in_dim %d (need a multiple of 16 in [64, %d])", i, w.in_dim, AS_MAXD);
    TE_CHECK_ARG(w.param_dim % 16 == 0 && w.param_dim >= 64 && w.param_dim <= AS_MAXD,
                 "attn_stack: block %d param_dim %d (need a multiple of 16 in [64, %d])", i, w.param_dim, AS_MAXD);
    TE_CHECK_ARG(i == 0 || (w.in_dim == AS_C && w.param_dim == AS_C),
                 "attn_stack: block %d must be 512 wide (it reads the previous block's output and p)", i);
    TE_CHECK_ARG((w.w_proj != nullptr) == (w.in_dim != AS_C) && (w.b_proj != nullptr) == (w.in_dim != AS_C),
                 "attn_stack: block %d needs w_proj/b_proj exactly when in_dim != 512", i);
    TE_CHECK_ARG(w.w_q && w.b_q && w.w_k && w.b_k && w.w_v && w.b_v && w.w_o && w.b_o && w.w_m1 && w.b_m1 &&
                     w.w_m2 && w.b_m2,
                 "attn_stack: block %d has a null parameter pointer", i);
    const uintptr_t al = reinterpret_cast<uintptr_t>(w.w_proj) | reinterpret_cast<uintptr_t>(w.w_q) |
                         reinterpret_cast<uintptr_t>(w.w_k) | reinterpret_cast<uintptr_t>(w.w_v) |
                         reinterpret_cast<uintptr_t>(w.w_o) | reinterpret_cast<uintptr_t>(w.w_m1) |
                         reinterpret_cast<uintptr_t>(w.w_m2);
    TE_CHECK_ARG((al & 15) == 0, "attn_stack: block %d has a weight matrix that is not 16-byte aligned", i);
  }
  return TE_OK;
}

template <typename Kern>
static int as_max_clusters(Kern kern, int cl, int smem, int* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cl * 64);
  cfg.blockDim = dim3(AS_NT);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  TE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  TE_CHECK_CUDA(cudaOccupancyMaxActiveClusters(out, kern, &cfg));
  return TE_OK;
}

// Resident clusters of 8 and of 4 CTAs for the forward [0..1] and backward [2..3] kernels (queried once).
static int as_occupancy(int (&occ)[4]) {
  static int cached[4] = {0, 0, 0, 0};
  static bool done = false;
  if (!done) {
    if (int rc = as_max_clusters(attn_stack_fwd_kernel<8, true>, 8, AS_FWD_SMEM, &cached[0])) return rc;
    if (int rc = as_max_clusters(attn_stack_fwd_kernel<4, true>, 4, AS_FWD_SMEM, &cached[1])) return rc;
    if (int rc = as_max_clusters(attn_stack_bwd_kernel<8, true>, 8, AS_BWD_SMEM, &cached[2])) return rc;
    if (int rc = as_max_clusters(attn_stack_bwd_kernel<4, true>, 4, AS_BWD_SMEM, &cached[3])) return rc;
    done = true;
  }
  for (int i = 0; i < 4; ++i) occ[i] = cached[i];
  return TE_OK;
}

// 8 CTAs per sample halve the weight bytes each CTA streams, 4 fit twice as many samples at once: take whichever
// needs less time for this batch (a wave of 4-CTA clusters costs ~1.7x a wave of 8-CTA clusters).
static int as_pick_cluster(int batch, int occ8, int occ4) {
  if (occ8 < 1) return 4;
  if (occ4 < 1) return 8;
  const int waves8 = (batch + occ8 - 1) / occ8, waves4 = (batch + occ4 - 1) / occ4;
  return 10 * waves8 <= 17 * waves4 ? 8 : 4;
}

template <typename Kern, typename Params>
static int as_launch(Kern kern, const Params& P, int cl, int batch, int smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(batch * cl);
  cfg.blockDim = dim3(AS_NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  // every instantiation needs its own opt-in to > 48 KB of dynamic shared memory (they share one pointer TYPE,
  // so remember the pointer values)
  static const void* configured[16];
  static int n_configured = 0;
  bool seen = false;
  for (int i = 0; i < n_configured; ++i) seen |= configured[i] == reinterpret_cast<const void*>(kern);
  if (!seen) {
    TE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (n_configured < 16) configured[n_configured++] = reinterpret_cast<const void*>(kern);
  }
  TE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  return TE_OK;
}

}  // namespace te

extern "C" int te_attn_stack_occupancy(int* fwd_clusters, int* bwd_clusters) {
  using namespace te;
  TE_CHECK_ARG(fwd_clusters && bwd_clusters, "attn_stack_occupancy: null pointer");
  int occ[4];
  if (int rc = as_occupancy(occ)) return rc;
  fwd_clusters[0] = occ[0];
  fwd_clusters[1] = occ[1];
  bwd_clusters[0] = occ[2];
  bwd_clusters[1] = occ[3];
  return TE_OK;
}

extern "C" int te_attn_stack_workspace(const te_attn_block* blocks, int n_blocks, int batch, int64_t* save_floats,
                                       int64_t* gws_floats) {
  using namespace te;
  TE_CHECK_ARG(blocks && n_blocks >= 1 && n_blocks <= AS_MAXB && batch >= 0, "attn_stack_workspace: bad arguments");
  int64_t s = 0;
  for (int i = 0; i < n_blocks; ++i) s += as_save_off(batch, blocks[i].in_dim).total;
  if (save_floats) *save_floats = s;
  if (gws_floats) *gws_floats = as_gws_off(batch).total * n_blocks;
  return TE_OK;
}

extern "C" int te_attn_stack_fwd(float* y, const float* x0, const float* p0, const float* p,
                                 const te_attn_block* blocks, int n_blocks, int batch, float lr_mul, int precision,
                                 float* save, void* stream) {
  using namespace te;
  if (int rc = as_validate(blocks, n_blocks, batch, precision)) return rc;
  TE_CHECK_ARG(y && x0 && p0 && (p || n_blocks == 1), "attn_stack_fwd: null pointer");
  if (batch == 0) return TE_OK;
  AsParams P;
  for (int i = 0; i < n_blocks; ++i) P.blk[i] = blocks[i];
  P.n_blocks = n_blocks;
  P.batch = batch;
  P.lr_mul = lr_mul;
  P.x0 = x0;
  P.p0 = p0;
  P.p = p;
  P.y = y;
  P.save = save;
  int occ[4];
  if (int rc = as_occupancy(occ)) return rc;
  const int cl = as_pick_cluster(batch, occ[0], occ[1]);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cl == 8) {
    return precision == 0 ? as_launch(attn_stack_fwd_kernel<8, true>, P, 8, batch, AS_FWD_SMEM, st)
                          : as_launch(attn_stack_fwd_kernel<8, false>, P, 8, batch, AS_FWD_SMEM, st);
  }
  return precision == 0 ? as_launch(attn_stack_fwd_kernel<4, true>, P, 4, batch, AS_FWD_SMEM, st)
                        : as_launch(attn_stack_fwd_kernel<4, false>, P, 4, batch, AS_FWD_SMEM, st);
}

extern "C" int te_attn_stack_bwd(float* g_x0, float* g_p0, float* g_p, const te_attn_block* grads, const float* gy,
                                 const float* x0, const float* p0, const float* p, const te_attn_block* blocks,
                                 int n_blocks, int batch, float lr_mul, int precision, const float* save, float* gws,
                                 void* stream) {
  using namespace te;
  if (int rc = as_validate(blocks, n_blocks, batch, precision)) return rc;
  TE_CHECK_ARG(grads && gy && x0 && p0 && (p || n_blocks == 1) && save && gws, "attn_stack_bwd: null pointer");
  if (batch == 0) return TE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AsBwdParams P;
  for (int i = 0; i < n_blocks; ++i) P.blk[i] = blocks[i];
  P.n_blocks = n_blocks;
  P.batch = batch;
  P.lr_mul = lr_mul;
  P.gy = gy;
  P.save = save;
  P.gws = gws;
  P.g_x0 = g_x0;
  P.g_p0 = g_p0;
  P.g_p = g_p;

  // grouped weight-gradient GEMM: 6 (+1 with proj) layers per block (validated before anything is launched)
  AsWgParams Wp;
  const AsGwsOff go = as_gws_off(batch);
  const int64_t rows = int64_t(batch) * AS_T;
  int nt = 0, tiles = 0;
  int64_t save_base = 0;
  auto add = [&](const float* G, const float* X, const float* gW, const float* gb, int N, int K) {
    AsWgTask& t = Wp.t[nt++];
    t.G = G;
    t.X = X;
    t.gW = const_cast<float*>(gW);
    t.gb = const_cast<float*>(gb);
    t.N = N;
    t.K = K;
    t.scale = lr_mul / sqrtf(float(K));
    t.tile0 = tiles;
    tiles += (N / 64) * ((K + 63) / 64);
  };
  for (int i = 0; i < n_blocks; ++i) {
    const te_attn_block& w = blocks[i];
    const te_attn_block& g = grads[i];
    TE_CHECK_ARG(g.w_q && g.b_q && g.w_k && g.b_k && g.w_v && g.b_v && g.w_o && g.b_o && g.w_m1 && g.b_m1 &&
                     g.w_m2 && g.b_m2 && (w.w_proj == nullptr || (g.w_proj && g.b_proj)),
                 "attn_stack_bwd: block %d has a null gradient pointer", i);
    const uintptr_t al = reinterpret_cast<uintptr_t>(g.w_proj) | reinterpret_cast<uintptr_t>(g.w_q) |
                         reinterpret_cast<uintptr_t>(g.w_k) | reinterpret_cast<uintptr_t>(g.w_v) |
                         reinterpret_cast<uintptr_t>(g.w_o) | reinterpret_cast<uintptr_t>(g.w_m1) |
                         reinterpret_cast<uintptr_t>(g.w_m2);
    TE_CHECK_ARG((al & 15) == 0, "attn_stack_bwd: block %d has a weight-gradient buffer that is not 16-byte aligned", i);
    const AsSaveOff so = as_save_off(batch, w.in_dim);
    const float* sv = save + save_base;
    const float* gw = gws + int64_t(i) * go.total;
    const float* gout = i + 1 == n_blocks ? gy : gw + go.gout;
    add(gout, sv + so.h, g.w_m2, g.b_m2, AS_C, AS_C);
    add(gw + go.gu, sv + so.ln1, g.w_m1, g.b_m1, AS_C, AS_C);
    add(gw + go.gx1, sv + so.att, g.w_o, g.b_o, AS_C, AS_PL);
    if (w.w_proj) add(gw + go.gx1, x0, g.w_proj, g.b_proj, AS_C, w.in_dim);
    add(gw + go.gq, i == 0 ? p0 : p, g.w_q, g.b_q, AS_PL, w.param_dim);
    add(gw + go.gk, sv + so.xn, g.w_k, g.b_k, AS_PL, w.in_dim);
    add(gw + go.gv, sv + so.xn, g.w_v, g.b_v, AS_PL, w.in_dim);
    save_base += so.total;
  }
  Wp.n_tasks = nt;
  Wp.rows = int(rows);
  Wp.total_tiles = tiles;
  Wp.lr_mul = lr_mul;

  int occ[4];
  if (int rc = as_occupancy(occ)) return rc;
  const int cl = as_pick_cluster(batch, occ[2], occ[3]);
  int rc;
  if (cl == 8) {
    rc = precision == 0 ? as_launch(attn_stack_bwd_kernel<8, true>, P, 8, batch, AS_BWD_SMEM, st)
                        : as_launch(attn_stack_bwd_kernel<8, false>, P, 8, batch, AS_BWD_SMEM, st);
  } else {
    rc = precision == 0 ? as_launch(attn_stack_bwd_kernel<4, true>, P, 4, batch, AS_BWD_SMEM, st)
                        : as_launch(attn_stack_bwd_kernel<4, false>, P, 4, batch, AS_BWD_SMEM, st);
  }
  if (rc) return rc;
  const int grid = tiles < 4 * kNumSMs ? tiles : 4 * kNumSMs;
  attn_stack_wgrad_kernel<<<grid, 256, 0, st>>>(Wp);
  TE_CHECK_LAUNCH();
  return TE_OK;
}
