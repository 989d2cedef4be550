// Tensor-core implicit-GEMM convolution for sm_100a: TMA -> shared memory (128B swizzle) ->
// tcgen05.mma (bf16 x bf16 -> f32 accumulators in TMEM) -> tcgen05.ld epilogue.
//
// One kernel covers every convolution and data-gradient of the generator / discriminator through a
// TAP TABLE over an ANCHOR grid (te_tc_conv_desc, include/te_b200.h):
//     y[b, a*os + oo, :] = act( out_scale[b,:] * SUM_t  x[b, a*is + d_t, :] . W[w_t]^T  + bias )
//   stride-1 conv        is=1, os=1, taps d=(ky-p, kx-p)                 (and its data gradient)
//   stride-2 conv        is=2, os=1, taps d=(ky, kx)                     (D down-convs; grad of convT)
//   transposed stride-2  is=1, os=2, one launch per output parity class  (G up-convs; grad of stride-2)
//
// GEMM view: D[anchor, cout] = sum over taps and 64-channel chunks of X_tile . W_tile^T
//   * M tile = 128 anchors = an (NB x TH x TW) patch; its A operand for one tap is ONE TMA box of the
//     4-D tensor [B,H,W,C] at shifted coordinates (traversal stride `is` in the tensor map).  Rows that
//     fall outside the image are zero-filled by the TMA unit, which IS the convolution's zero padding:
//     no im2col buffer, no halo logic.  Channel tails (Cin % 64 != 0) are zero-filled the same way.
//   * B operand = a [BLOCK_N x 64] box of the tap-major weight tensor; every CTA reads the same
//     weights, so they stay L2-resident (the point of the shared-weight formulation).
//   * K loop = taps x ceil(Cin/64) k-blocks of 4 x tcgen05.mma (M128, N=BLOCK_N, K16); a pipeline stage holds
//     TC_KSUB k-blocks (1: measured, tools/mma_rate_probe.py + profiles/r02_mma_rate_probe.md — with operands in
//     shared memory one tcgen05.mma of M = 128 rows per CTA costs ~140 cycles whatever N <= 256 is and however many
//     MMAs share a barrier round trip, so N = 256 tiles run at the cuBLAS rate and N = 128 tiles at ~half of it;
//     two k-blocks per stage only coarsened the pipeline: 0.385 vs 0.329 ms on the 128-channel 256^2 layer).
//   * Persistent: one CTA per SM walks the tile list.  Warp 0 = TMA producer, warp 1 = MMA issuer
//     (+ TMEM alloc), warps 2-5 = epilogue (TMEM lane quarter = warp_id % 4).  A 6-stage (8 at
//     BLOCK_N=64) full/empty mbarrier ring feeds the MMAs; TWO accumulator buffers in TMEM
//     (tmem_full / tmem_empty barriers) let the producer and the MMA issuer run into the next tile
//     while the epilogue warps drain the previous one.
//   * Epilogue: out = act(acc * out_scale[b,cout] + bias[cout]) -> bf16 (or f32), 16-byte stores.
//   * Split-operand mode (te_tc_conv_desc.split = 2 or 3): x and W are stored as 2 (3) bf16 PLANES hi, mid(, lo)
//     of f32 values (te_split_bf16 / te_pack_weights_tc) and the K loop additionally walks the plane pairs
//     (hi,hi) (hi,mid) (mid,hi) [(mid,mid) (hi,lo) (lo,hi)], all accumulated in the same f32 TMEM tile: the f32
//     convolution to ~2^-16 (2^-24) relative accuracy at 3 (6) tensor-core products — the fp32 parity mode.
//     The tensor core's f32 accumulation truncates (measured: the end-to-end error grows with the number of MMAs
//     accumulated into one tile), so the split kernels spread one tile over FOUR TMEM accumulators: the hi*hi
//     k-blocks round-robin over three of them (a third of the roundings each, against sums a third the size) and
//     all correction pairs (2^-8 and smaller) go to the fourth; the epilogue adds the four in registers
//     (round-to-nearest).  Measured image error at 256^2: see profiles/r02_parity_modes*.txt.
#include <stdlib.h>

#include "tc_common.cuh"
#include "conv_tc2.cuh"

namespace te {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;        // bf16 elements = 128 bytes = one swizzle row
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 192;
constexpr int TC_KSUB = 1;              // k-blocks per pipeline stage (see the header)
constexpr int TC_SMEM_BUDGET = 196608;  // operand ring
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;  // 16 KB

struct TcParams {
  int batch, cin, cout, hout, wout;
  int ntaps;
  int tap_dy[9], tap_dx[9], tap_w[9];
  int in_stride, out_stride, out_off_y, out_off_x;
  int grid_h, grid_w;      // anchors per sample
  int tw, th, nb;          // tile patch; nb*th*tw == 128
  int tiles_w, tiles_h, tiles_b, n_tiles;
  int w_slices_per_sample; // 0: shared weights; else slices per sample (weights indexed b*slices + w_t)
  int halo_x0, halo_y0;    // haloed-patch kernel: input offset of the box origin relative to the tile's first anchor
  int halo_bo;             // profiling aid: 1 = leave the descriptor's matrix base offset at 0
  int nseg, npairs;        // split-operand planes and plane pairs (1, 1 = plain bf16)
  int pair_a[6], pair_w[6];
  int act;
  float act_gain;
  const void* residual;    // [B, hout, wout, cout] (the output's dtype) added after the activation, or null
  const float* slope;      // act 3 (PReLU): negative slope per output channel [cout]
  int debug;               // TE_TC_DEBUG bits (profiling aid): 1 no stores, 2 no epilogue work, 4 no MMAs, 8 no loads
  const float* out_scale;  // [B, cout] or null
  const float* bias;       // [cout] or null
  void* y;
};

template <int BLOCK_N>
struct TcSmem {
  static constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
  static constexpr int SUB_BYTES = TC_A_BYTES + B_BYTES;          // one k-block: A tile + B tile
  static constexpr int STAGE_BYTES = TC_KSUB * SUB_BYTES;
  static constexpr int STAGES = TC_SMEM_BUDGET / STAGE_BYTES;     // 8 at N=64, 6 at N=128, 4 at N=256
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFFSET = BAR_OFFSET + 256;     // per-epilogue-warp scale/bias staging
  static constexpr int TOTAL = EPI_OFFSET + 4 * 2 * BLOCK_N * 4 + 1024;  // + slack for 1024-B alignment
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Epilogue of one 32-column chunk of one accumulator row (= one output pixel): scale, bias, activation,
// residual, store.  Shared by the 1-CTA and 2-CTA kernels.
template <bool OUT_F32>
__device__ __forceinline__ void tc_epilogue_chunk(const TcParams& p, const uint32_t (&v)[32], const float* s_osc,
                                                  const float* s_bias, bool staged, const float* osc, bool valid,
                                                  int64_t pix, int n0c0) {
  float f[32];
  if (staged) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 sc = *reinterpret_cast<const float4*>(s_osc + j);
      const float4 bi = *reinterpret_cast<const float4*>(s_bias + j);
      f[j] = __uint_as_float(v[j]) * sc.x + bi.x;
      f[j + 1] = __uint_as_float(v[j + 1]) * sc.y + bi.y;
      f[j + 2] = __uint_as_float(v[j + 2]) * sc.z + bi.z;
      f[j + 3] = __uint_as_float(v[j + 3]) * sc.w + bi.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0c0 + j;
      float a = __uint_as_float(v[j]);
      if (n < p.cout) {
        if (osc) a *= __ldg(osc + n);
        if (p.bias) a += __ldg(p.bias + n);
      }
      f[j] = a;
    }
  }
  if (p.act == 3) {  // PReLU: one slope per output channel (inference side path: plain cached loads)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0c0 + j;
      if (f[j] < 0.f && n < p.cout) f[j] *= __ldg(p.slope + n);
    }
  } else if (p.act != 0) {
    const float gain = p.act_gain, slope = p.act == 2 ? 0.01f : 0.2f;
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = (f[j] > 0.f ? f[j] : slope * f[j]) * gain;
  }
  if (p.residual != nullptr && valid) {
    if (OUT_F32) {
      const float4* rsrc = reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + pix * p.cout + n0c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (n0c0 + 4 * j >= p.cout) continue;
        const float4 r = __ldg(rsrc + j);
        f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
      }
    } else {
      const uint4* rsrc =
          reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + pix * p.cout + n0c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n0c0 + 8 * j >= p.cout) continue;
        const uint4 r = __ldg(rsrc + j);
        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          f[8 * j + 2 * q] += __uint_as_float(rw[q] << 16);
          f[8 * j + 2 * q + 1] += __uint_as_float(rw[q] & 0xffff0000u);
        }
      }
    }
  }
  if (valid && !(p.debug & 1)) {
    if (OUT_F32) {
      float* dst = static_cast<float*>(p.y) + pix * p.cout + n0c0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n0c0 + 4 * j < p.cout)
          reinterpret_cast<float4*>(dst)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    } else {
      __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.y) + pix * p.cout + n0c0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n0c0 + 8 * j >= p.cout) continue;
        uint4 o;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
        __nv_bfloat162 t1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
        __nv_bfloat162 t3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
        o.x = *reinterpret_cast<uint32_t*>(&t0);
        o.y = *reinterpret_cast<uint32_t*>(&t1);
        o.z = *reinterpret_cast<uint32_t*>(&t2);
        o.w = *reinterpret_cast<uint32_t*>(&t3);
        reinterpret_cast<uint4*>(dst)[j] = o;
      }
    }
  }
}

// Persistent kernel: one CTA per SM walks tiles t = blockIdx.x, +gridDim.x, ...; the operand ring and
// the two TMEM accumulator buffers let the TMA producer / MMA issuer run ahead into the next tile
// while the epilogue warps drain the previous one.
// accumulator slots of one tile / tile buffers in TMEM (512 columns): plain 1 x 2; split 4 x (2 at N=64, 1 at N=128)
template <int BLOCK_N, bool SPLIT> struct TcAcc {
  static constexpr int SLOTS = SPLIT ? 4 : 1;
  static constexpr int NBUF = (2 * SLOTS * BLOCK_N <= 512) ? 2 : 1;
  static constexpr uint32_t COLS = NBUF * SLOTS * BLOCK_N < 32 ? 32 : NBUF * SLOTS * BLOCK_N;
  static_assert(COLS <= 512, "TMEM has 512 columns");
};

template <int BLOCK_N, bool OUT_F32, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ TcParams p) {
  using S = TcSmem<BLOCK_N>;
  constexpr int STAGES = S::STAGES;
  using ACC = TcAcc<BLOCK_N, SPLIT>;
  constexpr uint32_t TMEM_COLS = ACC::COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = (p.cout + BLOCK_N - 1) / BLOCK_N;
  const int total_tiles = p.n_tiles * n_blocks;
  const int k_chunks = (p.cin + TC_BLOCK_K - 1) / TC_BLOCK_K;
  const int kb_per_tap = p.npairs * k_chunks;  // plane pairs x 64-channel chunks
  const int num_kb = p.ntaps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation: whole warp, power-of-two columns >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> coordinates (n fastest so CTAs running at the same time share activation tiles in L2)
  auto tile_coords = [&](int tile, int& b0, int& ay0, int& ax0, int& n0) {
    const int n_blk = tile % n_blocks;
    int m_blk = tile / n_blocks;
    const int tile_w = m_blk % p.tiles_w; m_blk /= p.tiles_w;
    const int tile_h = m_blk % p.tiles_h; m_blk /= p.tiles_h;
    b0 = m_blk * p.nb; ay0 = tile_h * p.th; ax0 = tile_w * p.tw; n0 = n_blk * BLOCK_N;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int b0, ay0, ax0, n0;
        tile_coords(tile, b0, ay0, ax0, n0);
        if (tile + int(gridDim.x) < total_tiles) {
          // pull the NEXT tile's activation patch into L2 now: its first touch comes from HBM, and an HBM
          // miss in the middle of the ring stalls every stage behind it (measured: feed-bound at 128 channels)
          int pb, py, px, pn;
          tile_coords(tile + int(gridDim.x), pb, py, px, pn);
          for (int sg = 0; sg < p.nseg; ++sg)
            for (int kc = 0; kc < k_chunks; ++kc)
              tma_prefetch_5d(&map_x, kc * TC_BLOCK_K, px * p.in_stride, py * p.in_stride, pb, sg);
        }
        for (int kb0 = 0; kb0 < num_kb; kb0 += TC_KSUB, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const int nsub = num_kb - kb0 < TC_KSUB ? num_kb - kb0 : TC_KSUB;
          mbar_expect_tx(&full_bar[s], nsub * S::SUB_BYTES);
          for (int sub = 0; sub < nsub; ++sub) {
            const int kb = kb0 + sub;
            const int tap = kb / kb_per_tap, r = kb - tap * kb_per_tap;
            const int pr = r / k_chunks, kc = r - pr * k_chunks;
            uint8_t* a_dst = smem + s * S::STAGE_BYTES + sub * S::SUB_BYTES;
            uint8_t* b_dst = a_dst + TC_A_BYTES;
            tma_load_5d(a_dst, &map_x, &full_bar[s], kc * TC_BLOCK_K, ax0 * p.in_stride + p.tap_dx[tap],
                        ay0 * p.in_stride + p.tap_dy[tap], b0, p.pair_a[pr]);
            const int wsl = p.w_slices_per_sample ? b0 * p.w_slices_per_sample + p.tap_w[tap] : p.tap_w[tap];
            tma_load_4d(b_dst, &map_w, &full_bar[s], kc * TC_BLOCK_K, n0, wsl, p.pair_w[pr]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_bf16(TC_BLOCK_M, BLOCK_N);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        const uint32_t buf = ACC::NBUF == 2 ? (ti & 1) : 0;
        // epilogue has drained this accumulator
        mbar_wait(&tmem_empty_bar[buf], (ACC::NBUF == 2 ? ((ti >> 1) & 1) : (ti & 1)) ^ 1);
        tcgen05_fence_after();
        const uint32_t d_base = tmem_base + buf * ACC::SLOTS * BLOCK_N;
        bool corr_started = false;
        for (int kb0 = 0; kb0 < num_kb; kb0 += TC_KSUB, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const int nsub = num_kb - kb0 < TC_KSUB ? num_kb - kb0 : TC_KSUB;
          for (int sub = 0; sub < nsub; ++sub) {
            const int kb = kb0 + sub;
            const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES + sub * S::SUB_BYTES);
            const uint64_t da = make_sw128_desc(a_addr);
            const uint64_t db = make_sw128_desc(a_addr + TC_A_BYTES);
            uint32_t d_tmem = d_base;
            bool fresh = kb == 0;
            if (SPLIT) {
              // hi*hi k-blocks round-robin over slots 0..2, every correction pair into slot 3 (see the header)
              const int tap = kb / kb_per_tap, r = kb - tap * kb_per_tap;
              const int pr = r / k_chunks, kc = r - pr * k_chunks;
              if (pr == 0) {
                const int idx = tap * k_chunks + kc;
                d_tmem = d_base + (idx % 3) * BLOCK_N;
                fresh = idx < 3;
              } else {
                d_tmem = d_base + 3 * BLOCK_N;
                fresh = !corr_started;
                corr_started = true;
              }
            }
            if (!(p.debug & 4)) {
#pragma unroll
              for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
                umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (!fresh || k != 0) ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
        }
        umma_commit(&tmem_full_bar[buf]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int per_img = p.th * p.tw;
    const int nbi = row / per_img;
    const int rem = row - nbi * per_img;
    const int thi = rem / p.tw, twi = rem - thi * p.tw;
    // per-warp staging of out_scale / bias for the tile (valid when the tile lies in one sample)
    float* s_osc = reinterpret_cast<float*>(smem + S::EPI_OFFSET) + quarter * 2 * BLOCK_N;
    float* s_bias = s_osc + BLOCK_N;
    const bool staged = p.nb == 1;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      int b0, ay0, ax0, n0;
      tile_coords(tile, b0, ay0, ax0, n0);
      const uint32_t buf = ACC::NBUF == 2 ? (ti & 1) : 0;
      if (staged) {
        __syncwarp();
        for (int c = lane; c < BLOCK_N; c += 32) {
          const int n = n0 + c;
          s_osc[c] = (p.out_scale && n < p.cout) ? __ldg(p.out_scale + static_cast<int64_t>(b0) * p.cout + n) : 1.f;
          s_bias[c] = (p.bias && n < p.cout) ? __ldg(p.bias + n) : 0.f;
        }
        __syncwarp();
      }
      const int b = b0 + nbi, ay = ay0 + thi, ax = ax0 + twi;
      const bool valid = b < p.batch && ay < p.grid_h && ax < p.grid_w;
      const int bs = valid ? b : 0;
      const float* osc = p.out_scale ? p.out_scale + static_cast<int64_t>(bs) * p.cout : nullptr;
      const int oy = ay * p.out_stride + p.out_off_y, ox = ax * p.out_stride + p.out_off_x;
      const int64_t pix = (static_cast<int64_t>(bs) * p.hout + oy) * p.wout + ox;

      mbar_wait(&tmem_full_bar[buf], ACC::NBUF == 2 ? ((ti >> 1) & 1) : (ti & 1));
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        if (n0 + c0 >= p.cout || (p.debug & 2)) break;  // warp-uniform
        uint32_t v[32];
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * ACC::SLOTS * BLOCK_N + c0;
        if (SPLIT) {
          // sum the tile's accumulator slots in registers (round-to-nearest): corrections first, then the hi*hi slots
          tmem_ld32(t_row + 3 * BLOCK_N, v);
          const int n_main = num_kb / p.npairs < 3 ? num_kb / p.npairs : 3;
          for (int sl = 0; sl < n_main; ++sl) {
            uint32_t u[32];
            tmem_ld32(t_row + sl * BLOCK_N, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
          }
        } else {
          tmem_ld32(t_row, v);
        }
        tc_epilogue_chunk<OUT_F32>(p, v, s_osc + c0, s_bias + c0, staged, osc, valid, pix, n0 + c0);
      }
      // this warp is done reading the accumulator buffer: hand it back to the MMA issuer
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

// ---- 2-CTA (cta_group::2) kernel: see conv_tc2.cuh for the protocol --------------------------------
template <int BLOCK_N>
struct Tc2Smem {
  static constexpr int B_BYTES = (BLOCK_N / 2) * TC_BLOCK_K * 2;   // this CTA's half of the weight tile
  static constexpr int SUB_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int STAGE_BYTES = TC_KSUB * SUB_BYTES;
  static constexpr int STAGES = TC_SMEM_BUDGET / STAGE_BYTES;      // 8 at N=128, 6 at N=256
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFFSET = BAR_OFFSET + 256;
  static constexpr int TOTAL = EPI_OFFSET + 4 * 2 * BLOCK_N * 4 + 1024;
};

template <int BLOCK_N, bool OUT_F32, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ TcParams p) {
  using S = Tc2Smem<BLOCK_N>;
  constexpr int STAGES = S::STAGES;
  using ACC = TcAcc<BLOCK_N, SPLIT>;
  constexpr uint32_t TMEM_COLS = ACC::COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);   // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2], used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_blocks = (p.cout + BLOCK_N - 1) / BLOCK_N;
  const int m_pairs = (p.n_tiles + 1) / 2;
  const int total_work = m_pairs * n_blocks;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int k_chunks = (p.cin + TC_BLOCK_K - 1) / TC_BLOCK_K;
  const int kb_per_tap = p.npairs * k_chunks;  // plane pairs x 64-channel chunks
  const int num_kb = p.ntaps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 8);  // 4 epilogue warps x 2 CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto work_coords = [&](int w, int& b0, int& ay0, int& ax0, int& n0) {
    const int n_blk = w % n_blocks;
    int m_blk = (w / n_blocks) * 2 + int(rank);   // this CTA's M tile of the pair
    const int tile_w = m_blk % p.tiles_w; m_blk /= p.tiles_w;
    const int tile_h = m_blk % p.tiles_h; m_blk /= p.tiles_h;
    b0 = m_blk * p.nb; ay0 = tile_h * p.th; ax0 = tile_w * p.tw; n0 = n_blk * BLOCK_N;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): own A tile + own half of the B tile, signalling the leader's barrier =====
      uint32_t it = 0;
      for (int w = pair_id; w < total_work; w += num_pairs) {
        int b0, ay0, ax0, n0;
        work_coords(w, b0, ay0, ax0, n0);
        if (w + num_pairs < total_work) {  // L2 prefetch of the next work item's activation patch
          int pb, py, px, pn;
          work_coords(w + num_pairs, pb, py, px, pn);
          for (int sg = 0; sg < p.nseg; ++sg)
            for (int kc = 0; kc < k_chunks; ++kc)
              tma_prefetch_5d(&map_x, kc * TC_BLOCK_K, px * p.in_stride, py * p.in_stride, pb, sg);
        }
        for (int kb0 = 0; kb0 < num_kb; kb0 += TC_KSUB, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const int nsub = num_kb - kb0 < TC_KSUB ? num_kb - kb0 : TC_KSUB;
          if (p.debug & 8) {  // profiling aid: no loads
            if (leader) mbar_arrive(&full_bar[s]);
            continue;
          }
          if (leader) mbar_expect_tx(&full_bar[s], 2 * nsub * S::SUB_BYTES);
          for (int sub = 0; sub < nsub; ++sub) {
            const int kb = kb0 + sub;
            const int tap = kb / kb_per_tap, r = kb - tap * kb_per_tap;
            const int pr = r / k_chunks, kc = r - pr * k_chunks;
            uint8_t* a_dst = smem + s * S::STAGE_BYTES + sub * S::SUB_BYTES;
            uint8_t* b_dst = a_dst + TC_A_BYTES;
            tma2_load_5d(a_dst, &map_x, &full_bar[s], kc * TC_BLOCK_K, ax0 * p.in_stride + p.tap_dx[tap],
                         ay0 * p.in_stride + p.tap_dy[tap], b0, p.pair_a[pr]);
            const int wsl = p.w_slices_per_sample ? b0 * p.w_slices_per_sample + p.tap_w[tap] : p.tap_w[tap];
            tma2_load_4d(b_dst, &map_w, &full_bar[s], kc * TC_BLOCK_K, n0 + int(rank) * (BLOCK_N / 2), wsl,
                         p.pair_w[pr]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader only): M256 x BLOCK_N x K16 across the pair =====
      constexpr uint32_t idesc = make_idesc_bf16(2 * TC_BLOCK_M, BLOCK_N);
      uint32_t it = 0, ti = 0;
      for (int w = pair_id; w < total_work; w += num_pairs, ++ti) {
        const uint32_t buf = ACC::NBUF == 2 ? (ti & 1) : 0;
        mbar_wait(&tmem_empty_bar[buf], (ACC::NBUF == 2 ? ((ti >> 1) & 1) : (ti & 1)) ^ 1);
        tcgen05_fence_after();
        const uint32_t d_base = tmem_base + buf * ACC::SLOTS * BLOCK_N;
        bool corr_started = false;
        for (int kb0 = 0; kb0 < num_kb; kb0 += TC_KSUB, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const int nsub = num_kb - kb0 < TC_KSUB ? num_kb - kb0 : TC_KSUB;
          for (int sub = 0; sub < nsub; ++sub) {
            const int kb = kb0 + sub;
            const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES + sub * S::SUB_BYTES);
            const uint64_t da = make_sw128_desc(a_addr);
            const uint64_t db = make_sw128_desc(a_addr + TC_A_BYTES);
            uint32_t d_tmem = d_base;
            bool fresh = kb == 0;
            if (SPLIT) {
              const int tap = kb / kb_per_tap, r = kb - tap * kb_per_tap;
              const int pr = r / k_chunks, kc = r - pr * k_chunks;
              if (pr == 0) {
                const int idx = tap * k_chunks + kc;
                d_tmem = d_base + (idx % 3) * BLOCK_N;
                fresh = idx < 3;
              } else {
                d_tmem = d_base + 3 * BLOCK_N;
                fresh = !corr_started;
                corr_started = true;
              }
            }
            if (!(p.debug & 4)) {
#pragma unroll
              for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k)
                umma2_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (!fresh || k != 0) ? 1u : 0u);
            }
          }
          umma2_commit_both(&empty_bar[s]);
        }
        umma2_commit_both(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ===== epilogue (both CTAs): own 128 TMEM lanes =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int per_img = p.th * p.tw;
    const int nbi = row / per_img;
    const int rem = row - nbi * per_img;
    const int thi = rem / p.tw, twi = rem - thi * p.tw;
    float* s_osc = reinterpret_cast<float*>(smem + S::EPI_OFFSET) + quarter * 2 * BLOCK_N;
    float* s_bias = s_osc + BLOCK_N;
    const bool staged = p.nb == 1;
    uint32_t ti = 0;
    for (int w = pair_id; w < total_work; w += num_pairs, ++ti) {
      int b0, ay0, ax0, n0;
      work_coords(w, b0, ay0, ax0, n0);
      const uint32_t buf = ACC::NBUF == 2 ? (ti & 1) : 0;
      const int b = b0 + nbi, ay = ay0 + thi, ax = ax0 + twi;
      const bool valid = b < p.batch && ay < p.grid_h && ax < p.grid_w;
      const int bs = valid ? b : 0;
      if (staged) {
        __syncwarp();
        const int bq = b0 < p.batch ? b0 : 0;
        for (int c = lane; c < BLOCK_N; c += 32) {
          const int n = n0 + c;
          s_osc[c] = (p.out_scale && n < p.cout) ? __ldg(p.out_scale + static_cast<int64_t>(bq) * p.cout + n) : 1.f;
          s_bias[c] = (p.bias && n < p.cout) ? __ldg(p.bias + n) : 0.f;
        }
        __syncwarp();
      }
      const float* osc = p.out_scale ? p.out_scale + static_cast<int64_t>(bs) * p.cout : nullptr;
      const int oy = ay * p.out_stride + p.out_off_y, ox = ax * p.out_stride + p.out_off_x;
      const int64_t pix = (static_cast<int64_t>(bs) * p.hout + oy) * p.wout + ox;

      mbar_wait(&tmem_full_bar[buf], ACC::NBUF == 2 ? ((ti >> 1) & 1) : (ti & 1));
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        if (n0 + c0 >= p.cout || (p.debug & 2)) break;  // warp-uniform
        uint32_t v[32];
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * ACC::SLOTS * BLOCK_N + c0;
        if (SPLIT) {
          // sum the tile's accumulator slots in registers (round-to-nearest): corrections first, then the hi*hi slots
          tmem_ld32(t_row + 3 * BLOCK_N, v);
          const int n_main = num_kb / p.npairs < 3 ? num_kb / p.npairs : 3;
          for (int sl = 0; sl < n_main; ++sl) {
            uint32_t u[32];
            tmem_ld32(t_row + sl * BLOCK_N, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
          }
        } else {
          tmem_ld32(t_row, v);
        }
        tc_epilogue_chunk<OUT_F32>(p, v, s_osc + c0, s_bias + c0, staged, osc, valid, pix, n0 + c0);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[buf]);
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();   // the peer's shared memory / TMEM stay alive until the leader's last MMA has retired
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace te
#include "conv_tc_halo.cuh"
#include "conv_tct.cuh"
namespace te {

template <int BLOCK_N, bool OUT_F32, bool SPLIT = false>
static int launch_tc2(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, cudaStream_t st) {
  using S = Tc2Smem<BLOCK_N>;
  auto kern = conv_tc2_kernel<BLOCK_N, OUT_F32, SPLIT>;
  static bool configured = false;
  if (!configured) {
    TE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  const int total_work = ((p.n_tiles + 1) / 2) * ((p.cout + BLOCK_N - 1) / BLOCK_N);
  const int pairs = total_work < kNumSMs / 2 ? total_work : kNumSMs / 2;
  kern<<<2 * pairs, TC_THREADS, S::TOTAL, st>>>(mx, mw, p);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

// ---- host side ---------------------------------------------------------------------------------
template <int BLOCK_N, bool OUT_F32, bool SPLIT = false>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, cudaStream_t st) {
  using S = TcSmem<BLOCK_N>;
  auto kern = conv_tc_kernel<BLOCK_N, OUT_F32, SPLIT>;
  static bool configured = false;
  if (!configured) {
    TE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  const int total = p.n_tiles * ((p.cout + BLOCK_N - 1) / BLOCK_N);
  const int grid = total < kNumSMs ? total : kNumSMs;  // persistent: one CTA per SM
  kern<<<grid, TC_THREADS, S::TOTAL, st>>>(mx, mw, p);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

static int conv_tc_dispatch(void* y, const void* x, const void* w, const float* out_scale, const float* bias,
                            const te_tc_conv_desc& d, cudaStream_t st) {
  TE_CHECK_ARG(y && x && w, "conv_tc: null tensor pointer");
  TE_CHECK_ARG(d.batch > 0 && d.hin > 0 && d.win > 0 && d.hout > 0 && d.wout > 0, "conv_tc: bad shape");
  TE_CHECK_ARG(d.cin % 8 == 0 && d.cin >= 8, "conv_tc: Cin must be a multiple of 8 (got %d)", d.cin);
  TE_CHECK_ARG(d.cout % 8 == 0 && d.cout >= 8, "conv_tc: Cout must be a multiple of 8 (got %d)", d.cout);
  TE_CHECK_ARG(d.ntaps >= 1 && d.ntaps <= 9, "conv_tc: 1..9 taps");
  TE_CHECK_ARG((d.in_stride == 1 || d.in_stride == 2) && (d.out_stride == 1 || d.out_stride == 2),
               "conv_tc: strides must be 1 or 2");
  TE_CHECK_ARG(d.grid_h > 0 && d.grid_w > 0, "conv_tc: empty anchor grid");
  TE_CHECK_ARG((d.grid_h - 1) * d.out_stride + d.out_off_y < d.hout && (d.grid_w - 1) * d.out_stride + d.out_off_x < d.wout,
               "conv_tc: anchor grid exceeds the output tensor");
  TE_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(y) & 15) == 0, "conv_tc: pointers must be 16-byte aligned");
  TcParams p;
  p.batch = d.batch; p.cin = d.cin; p.cout = d.cout; p.hout = d.hout; p.wout = d.wout;
  p.ntaps = d.ntaps;
  for (int t = 0; t < 9; ++t) {
    p.tap_dy[t] = t < d.ntaps ? d.tap_dy[t] : 0;
    p.tap_dx[t] = t < d.ntaps ? d.tap_dx[t] : 0;
    p.tap_w[t] = t < d.ntaps ? d.tap_w[t] : 0;
    if (t < d.ntaps) TE_CHECK_ARG(d.tap_w[t] >= 0 && d.tap_w[t] < d.w_slices, "conv_tc: tap weight index out of range");
  }
  p.in_stride = d.in_stride; p.out_stride = d.out_stride; p.out_off_y = d.out_off_y; p.out_off_x = d.out_off_x;
  p.grid_h = d.grid_h; p.grid_w = d.grid_w;
  p.tw = next_pow2(d.grid_w) < 16 ? next_pow2(d.grid_w) : 16;
  int th = TC_BLOCK_M / p.tw;
  p.th = next_pow2(d.grid_h) < th ? next_pow2(d.grid_h) : th;
  p.nb = TC_BLOCK_M / (p.tw * p.th);
  p.tiles_w = (d.grid_w + p.tw - 1) / p.tw;
  p.tiles_h = (d.grid_h + p.th - 1) / p.th;
  p.tiles_b = (d.batch + p.nb - 1) / p.nb;
  p.n_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  p.act = d.act; p.out_scale = out_scale; p.bias = bias; p.y = y;
  TE_CHECK_ARG(d.act >= 0 && d.act <= 3, "conv_tc: act must be 0, 1, 2 or 3");
  TE_CHECK_ARG(d.act == 0 || d.act == 3 || !(d.act_gain < 0.f), "conv_tc: the activation gain must be positive");
  TE_CHECK_ARG(d.act != 3 || d.slope != nullptr, "conv_tc: act 3 (PReLU) needs the per-channel slopes");
  p.slope = static_cast<const float*>(d.slope);
  p.act_gain = d.act_gain != 0.f ? d.act_gain : (d.act == 2 ? 1.f : 1.4142135623730951f);
  p.residual = d.residual;
  TE_CHECK_ARG((reinterpret_cast<uintptr_t>(d.residual) & 15) == 0, "conv_tc: residual must be 16-byte aligned");
  TE_CHECK_ARG(d.split >= 0 && d.split <= 3, "conv_tc: split must be 0/1 (plain bf16), 2 or 3 planes");
  p.nseg = d.split < 2 ? 1 : d.split;
  {
    const SplitPairs sp = split_pairs(p.nseg);
    p.npairs = sp.n;
    for (int i = 0; i < 6; ++i) { p.pair_a[i] = sp.a[i]; p.pair_w[i] = sp.b[i]; }
  }
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("TE_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  p.halo_x0 = p.halo_y0 = p.halo_bo = 0;
  p.w_slices_per_sample = 0;
  if (d.w_bstride != 0) {
    TE_CHECK_ARG(d.w_bstride == int64_t(d.w_slices) * d.cout * d.cin, "conv_tc: per-sample weights must be densely packed");
    TE_CHECK_ARG(p.nb == 1, "conv_tc: per-sample weights need >= 128 anchors per sample");
    p.w_slices_per_sample = d.w_slices;
  }
  // N tile: 256 halves the A-operand shared-memory reads per FLOP (one-CTA UMMA at M128 x N128 is bound by
  // the 128 B/cycle shared-memory port: every K16 step reads 4 KB of A + 4 KB of B in 64 cycles); used when
  // the layer still yields >= 1 tile per SM.
  const bool split = p.nseg > 1;
  TE_CHECK_ARG(!split || d.out_f32, "conv_tc: the split-operand mode writes f32 output");
  int block_n = (d.cout % 128 == 0) ? 128 : 64;
  if (d.cout % 256 == 0 && !split) {  // split mode: four accumulator slots per tile fit TMEM only up to N = 128
    // waves x cost per tile: an N=256 tile does twice the work of an N=128 tile in ~1.4x the time (measured
    // 1.67 vs 1.10 PFLOP/s at 512 channels), so it wins whenever it does not cost an extra wave of its own
    const int64_t t256 = int64_t(p.n_tiles) * (d.cout / 256), t128 = 2 * t256;
    const int64_t w256 = (t256 + kNumSMs - 1) / kNumSMs, w128 = (t128 + kNumSMs - 1) / kNumSMs;
    if (t256 >= kNumSMs || 14 * w256 < 10 * w128) block_n = 256;
  }
  {  // TE_TC_BLOCK_N=128|256 (profiling aid): force the N tile where the layer allows it
    static int force_n = -1;
    if (force_n < 0) { const char* e = getenv("TE_TC_BLOCK_N"); force_n = e ? atoi(e) : 0; }
    if (force_n == 256 && d.cout % 256 == 0 && !split) block_n = 256;
    if (force_n == 128 && d.cout % 128 == 0) block_n = 128;
  }

  // 2-CTA pairs (cta_group::2) when the layer has enough tiles and, for per-sample weights, both tiles of
  // a pair always belong to one sample.  TE_TC_2CTA=0 forces the single-CTA kernel.
  static int use2 = -1;
  if (use2 < 0) { const char* e = getenv("TE_TC_2CTA"); use2 = e ? atoi(e) : 1; }
  bool two_cta = use2 != 0 && block_n >= 128 && int64_t((p.n_tiles + 1) / 2) * ((d.cout + block_n - 1) / block_n) >= kNumSMs / 2;
  if (two_cta && p.w_slices_per_sample && ((p.tiles_w * p.tiles_h) % 2) != 0) two_cta = false;

  // Haloed-patch kernel (conv_tc_halo.cuh): stride-1 geometries with several taps whose offsets fit the 18 x 16 box
  bool halo = false;
  {
    static int use_halo = -1, halo_bo = 0;
    if (use_halo < 0) {
      // off by default: measured no faster (the kernels are bound by MMA issue, not by operand fill, see header)
      const char* e = getenv("TE_TC_HALO"); use_halo = e ? atoi(e) : 0;
      halo_bo = 1;  // the swizzle is a function of the absolute shared-memory address: base offset stays 0
    }
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    for (int t = 0; t < d.ntaps; ++t) {
      x0 = t == 0 || d.tap_dx[t] < x0 ? d.tap_dx[t] : x0; x1 = t == 0 || d.tap_dx[t] > x1 ? d.tap_dx[t] : x1;
      y0 = t == 0 || d.tap_dy[t] < y0 ? d.tap_dy[t] : y0; y1 = t == 0 || d.tap_dy[t] > y1 ? d.tap_dy[t] : y1;
    }
    const int h_tiles = ((d.grid_w + TH_TILE_W - 1) / TH_TILE_W) * ((d.grid_h + TH_TILE_H - 1) / TH_TILE_H);
    const int64_t h_work = int64_t((int64_t(h_tiles) * d.batch + 1) / 2) * ((d.cout + block_n - 1) / block_n);
    halo = use_halo != 0 && use2 != 0 && d.in_stride == 1 && d.ntaps >= 2 && block_n >= 128 &&
           d.grid_h >= TH_TILE_H && d.grid_w >= TH_TILE_W && x1 - x0 <= TH_BOX_W - TH_TILE_W && y1 - y0 <= 2 &&
           h_work >= kNumSMs / 2 && !(p.w_slices_per_sample && (h_tiles % 2) != 0);
    if (halo) {
      p.tw = TH_TILE_W; p.th = TH_TILE_H; p.nb = 1;
      p.tiles_w = (d.grid_w + TH_TILE_W - 1) / TH_TILE_W;
      p.tiles_h = (d.grid_h + TH_TILE_H - 1) / TH_TILE_H;
      p.tiles_b = d.batch;
      p.n_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
      p.halo_x0 = x0; p.halo_y0 = y0; p.halo_bo = halo_bo;
      two_cta = true;
    }
  }
  // Operand-swapped kernel (conv_tct.cuh): channel tile of 128 only (Cout = 128 layers), 256-pixel tiles, enough of
  // them to fill the chip.  TE_TC_SWAP=0 keeps those layers on the kernels above.
  bool swapped = false;
  if (!halo && !split && block_n == 128) {
    static int use_swap = -1;
    if (use_swap < 0) { const char* e = getenv("TE_TC_SWAP"); use_swap = e ? atoi(e) : 1; }
    const int tw = next_pow2(d.grid_w) < 16 ? next_pow2(d.grid_w) : 16;
    const int th_max = TT_BLOCK_N / tw;
    const int th = next_pow2(d.grid_h) < th_max ? next_pow2(d.grid_h) : th_max;
    const int nb = TT_BLOCK_N / (tw * th);
    const int64_t tiles = int64_t((d.grid_w + tw - 1) / tw) * ((d.grid_h + th - 1) / th) * ((d.batch + nb - 1) / nb);
    const int num_kb = d.ntaps * ((d.cin + TC_BLOCK_K - 1) / TC_BLOCK_K);   // light tiles (1x1 layers) are epilogue-bound:
    swapped = use_swap != 0 && tw * th >= 32 && nb <= 256 && (use_swap == 2 || num_kb >= 8) &&   // they stay on the 4-chunk tiles
              (use_swap == 2 || tiles * (d.cout / TT_BLOCK_M) >= kNumSMs) &&   // 2 = force (tests)
              !(p.w_slices_per_sample && nb != 1);
    if (swapped) {
      p.tw = tw; p.th = th; p.nb = nb;
      p.tiles_w = (d.grid_w + tw - 1) / tw;
      p.tiles_h = (d.grid_h + th - 1) / th;
      p.tiles_b = (d.batch + nb - 1) / nb;
      p.n_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
      two_cta = false;
    }
  }
  CUtensorMap mx, mw;
  if (halo) {
    uint64_t dims[5] = {uint64_t(d.cin), uint64_t(d.win), uint64_t(d.hin), uint64_t(d.batch), uint64_t(p.nseg)};
    uint64_t strides[4] = {uint64_t(d.cin) * 2, uint64_t(d.win) * d.cin * 2, uint64_t(d.hin) * d.win * d.cin * 2,
                           uint64_t(d.batch) * d.hin * d.win * d.cin * 2};
    uint32_t box[5] = {TC_BLOCK_K, TH_BOX_W, TH_BOX_H, 1, 1};
    int rc = encode_map_bf16(&mx, x, 5, dims, strides, box, nullptr);
    if (rc) return rc;
  } else {
    const uint32_t is = uint32_t(d.in_stride);
    // [planes][B][H][W][C]: the split-operand planes are the outermost dimension (1 plane for plain bf16)
    uint64_t dims[5] = {uint64_t(d.cin), uint64_t(d.win), uint64_t(d.hin), uint64_t(d.batch), uint64_t(p.nseg)};
    uint64_t strides[4] = {uint64_t(d.cin) * 2, uint64_t(d.win) * d.cin * 2, uint64_t(d.hin) * d.win * d.cin * 2,
                           uint64_t(d.batch) * d.hin * d.win * d.cin * 2};
    // with a traversal stride e the unit loads ceil(box/e) elements: box = count * e
    uint32_t box[5] = {TC_BLOCK_K, uint32_t(p.tw) * is, uint32_t(p.th) * is, uint32_t(p.nb), 1};
    uint32_t estr[5] = {1, is, is, 1, 1};
    int rc = encode_map_bf16(&mx, x, 5, dims, strides, box, estr);
    if (rc) return rc;
  }
  {
    const uint64_t nw = uint64_t(d.w_slices) * (d.w_bstride ? d.batch : 1);
    uint64_t dims[4] = {uint64_t(d.cin), uint64_t(d.cout), nw, uint64_t(p.nseg)};
    uint64_t strides[3] = {uint64_t(d.cin) * 2, uint64_t(d.cout) * d.cin * 2, nw * d.cout * d.cin * 2};
    uint32_t box[4] = {TC_BLOCK_K, uint32_t(two_cta ? block_n / 2 : block_n), 1, 1};
    int rc = encode_map_bf16(&mw, w, 4, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  if (swapped) return d.out_f32 ? launch_tct<true>(mx, mw, p, st) : launch_tct<false>(mx, mw, p, st);
  if (halo) {
    if (split) return launch_tc2h<128, true, true>(mx, mw, p, st);
    if (block_n == 256)
      return d.out_f32 ? launch_tc2h<256, true>(mx, mw, p, st) : launch_tc2h<256, false>(mx, mw, p, st);
    return d.out_f32 ? launch_tc2h<128, true>(mx, mw, p, st) : launch_tc2h<128, false>(mx, mw, p, st);
  }
  if (split) {
    if (two_cta) return launch_tc2<128, true, true>(mx, mw, p, st);
    if (block_n == 128) return launch_tc<128, true, true>(mx, mw, p, st);
    return launch_tc<64, true, true>(mx, mw, p, st);
  }
  if (two_cta) {
    if (block_n == 256)
      return d.out_f32 ? launch_tc2<256, true>(mx, mw, p, st) : launch_tc2<256, false>(mx, mw, p, st);
    return d.out_f32 ? launch_tc2<128, true>(mx, mw, p, st) : launch_tc2<128, false>(mx, mw, p, st);
  }
  if (block_n == 256)
    return d.out_f32 ? launch_tc<256, true>(mx, mw, p, st) : launch_tc<256, false>(mx, mw, p, st);
  if (block_n == 128)
    return d.out_f32 ? launch_tc<128, true>(mx, mw, p, st) : launch_tc<128, false>(mx, mw, p, st);
  return d.out_f32 ? launch_tc<64, true>(mx, mw, p, st) : launch_tc<64, false>(mx, mw, p, st);
}

static void same_conv_desc(te_tc_conv_desc& d, int batch, int h, int wd, int cin, int cout, int kh, int kw, int act,
                           int64_t w_bstride, int out_f32) {
  d.batch = batch; d.hin = h; d.win = wd; d.cin = cin; d.hout = h; d.wout = wd; d.cout = cout;
  d.ntaps = kh * kw;
  for (int t = 0; t < 9; ++t) { d.tap_dy[t] = d.tap_dx[t] = d.tap_w[t] = 0; }
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx) {
      const int t = ky * kw + kx;
      d.tap_dy[t] = ky - kh / 2; d.tap_dx[t] = kx - kw / 2; d.tap_w[t] = t;
    }
  d.w_slices = kh * kw;
  d.in_stride = 1; d.out_stride = 1; d.out_off_y = 0; d.out_off_x = 0;
  d.grid_h = h; d.grid_w = wd; d.act = act; d.out_f32 = out_f32; d.w_bstride = w_bstride;
  d.act_gain = 0.f; d.wgrad_alpha = 0.f; d.residual = nullptr; d.slope = nullptr; d.split = 0; d.reserved = 0;
}

}  // namespace te

extern "C" int te_conv_tc(void* y, const void* x, const void* w, const float* out_scale, const float* bias,
                          const te_tc_conv_desc* d, void* stream) {
  using namespace te;
  TE_CHECK_ARG(d != nullptr, "conv_tc: null descriptor");
  return conv_tc_dispatch(y, x, w, out_scale, bias, *d, static_cast<cudaStream_t>(stream));
}

extern "C" int te_conv2d_tc(void* y, const void* x, const void* w, const float* out_scale, const float* bias,
                            int batch, int hin, int win, int cin, int cout, int kh, int kw, int act,
                            int64_t w_bstride, void* stream) {
  using namespace te;
  TE_CHECK_ARG(kh == kw && (kh == 1 || kh == 3), "conv2d_tc: kernel must be 1x1 or 3x3");
  te_tc_conv_desc d;
  same_conv_desc(d, batch, hin, win, cin, cout, kh, kw, act, w_bstride, 0);
  return conv_tc_dispatch(y, x, w, out_scale, bias, d, static_cast<cudaStream_t>(stream));
}

// D[M,N] f32 = A[M,K] * B[N,K]^T through the same kernel, as a 1x1 convolution over an [M/16, 16] "image".
extern "C" int te_gemm_tc_selftest(float* dd, const void* a, const void* b, int m, int n, int k, void* stream) {
  using namespace te;
  TE_CHECK_ARG(m % 16 == 0 && m >= 16, "gemm_tc_selftest: M must be a multiple of 16");
  te_tc_conv_desc d;
  same_conv_desc(d, 1, m / 16, 16, k, n, 1, 1, 0, 0, 1);
  return conv_tc_dispatch(dd, a, b, nullptr, nullptr, d, static_cast<cudaStream_t>(stream));
}
