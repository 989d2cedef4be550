// Operand-swapped variant of the tcgen05 implicit-GEMM convolution for layers whose output-channel count only
// allows a 128-wide channel tile (Cout = 128: every 256^2 layer of G and D).
//
// Why: with both operands in shared memory one tcgen05.mma costs ~140 cycles whatever its N <= 256 is
// (profiles/r02_mma_rate_probe.md), so the M128(pixels) x N128(channels) tiles of conv_tc_kernel run at half the
// rate of N = 256 tiles.  Here the roles are exchanged — D^T[cout, pixel] = W[cout, cin] . X[pixel, cin]^T:
//     A operand (M = 128)  = the [128 cout x 64 cin] weight box of one tap
//     B operand (N = 256)  = a [256 pixels x 64 cin] activation box (NB x TH x TW patch, same TMA maps / zero fill)
// Both boxes are K-major SWIZZLE_128B, so swapping them is swapping the two descriptors of the instruction; every
// MMA now carries twice the FLOPs and a weight tile is fetched once per 256 pixels instead of once per 128.
// The accumulator comes out transposed: TMEM lane = output channel, column = pixel.  An epilogue thread owns ONE
// channel (out_scale / bias / PReLU slope are per-thread scalars); each [32 channels x 32 pixels] block is transposed
// through a per-warp staging tile in shared memory and written with 16-byte stores (see the epilogue).
// Pipeline: as conv_tc_kernel — warp 0 TMA producer, warp 1 MMA issuer, 4-stage operand ring (48 KB per stage), two
// 256-column accumulators in TMEM — but EIGHT epilogue warps (two per TMEM lane quarter, even / odd pixel chunks): a
// single warp per scheduler issues ~0.25 instructions per cycle, which made the 4-warp epilogue (2 k cycles per
// 32-pixel chunk, ncu) the bottleneck of the 16 k-cycle tile.
#pragma once
#include "tc_common.cuh"

namespace te {

constexpr int TT_BLOCK_M = 128;   // output channels per tile
constexpr int TT_BLOCK_N = 256;   // pixels (anchors) per tile
constexpr int TT_W_BYTES = TT_BLOCK_M * TC_BLOCK_K * 2;   // 16 KB
constexpr int TT_X_BYTES = TT_BLOCK_N * TC_BLOCK_K * 2;   // 32 KB
constexpr int TT_STAGE_BYTES = TT_W_BYTES + TT_X_BYTES;
constexpr int TT_STAGES = 4;
constexpr int TT_EPI_WARPS = 8;                             // two per TMEM lane quarter (even / odd 32-pixel chunks)
constexpr int TT_THREADS = 64 + 32 * TT_EPI_WARPS;
constexpr int TT_BAR_OFFSET = TT_STAGES * TT_STAGE_BYTES;
constexpr int TT_EPI_OFFSET = TT_BAR_OFFSET + 256;
constexpr int TT_EPI_STRIDE = 36;   // floats per staged pixel row (32 channels + 4 of padding)
constexpr int TT_SMEM_TOTAL = TT_EPI_OFFSET + TT_EPI_WARPS * 16 * TT_EPI_STRIDE * 4 + 1024;  // 16-pixel staging tiles

template <bool OUT_F32>
__global__ void __launch_bounds__(TT_THREADS, 1)
conv_tct_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ TcParams p) {
  constexpr int STAGES = TT_STAGES;
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TT_BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = p.cout / TT_BLOCK_M;
  const int total_tiles = p.n_tiles * m_blocks;
  const int k_chunks = (p.cin + TC_BLOCK_K - 1) / TC_BLOCK_K;
  const int num_kb = p.ntaps * k_chunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], TT_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> coordinates (channel block fastest: CTAs running together share the activation patch in L2)
  auto tile_coords = [&](int tile, int& b0, int& ay0, int& ax0, int& n0) {
    const int m_blk = tile % m_blocks;
    int px = tile / m_blocks;
    const int tile_w = px % p.tiles_w; px /= p.tiles_w;
    const int tile_h = px % p.tiles_h; px /= p.tiles_h;
    b0 = px * p.nb; ay0 = tile_h * p.th; ax0 = tile_w * p.tw; n0 = m_blk * TT_BLOCK_M;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int b0, ay0, ax0, n0;
        tile_coords(tile, b0, ay0, ax0, n0);
        if (tile + int(gridDim.x) < total_tiles) {  // L2 prefetch of the next tile's activation patch
          int pb, py, px, pn;
          tile_coords(tile + int(gridDim.x), pb, py, px, pn);
          for (int kc = 0; kc < k_chunks; ++kc)
            tma_prefetch_5d(&map_x, kc * TC_BLOCK_K, px * p.in_stride, py * p.in_stride, pb, 0);
        }
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (p.debug & 8) {  // profiling aid: no loads
            mbar_arrive(&full_bar[s]);
            continue;
          }
          mbar_expect_tx(&full_bar[s], TT_STAGE_BYTES);
          const int tap = kb / k_chunks, kc = kb - tap * k_chunks;
          uint8_t* w_dst = smem + s * TT_STAGE_BYTES;
          uint8_t* x_dst = w_dst + TT_W_BYTES;
          const int wsl = p.w_slices_per_sample ? b0 * p.w_slices_per_sample + p.tap_w[tap] : p.tap_w[tap];
          tma_load_4d(w_dst, &map_w, &full_bar[s], kc * TC_BLOCK_K, n0, wsl, 0);
          tma_load_5d(x_dst, &map_x, &full_bar[s], kc * TC_BLOCK_K, ax0 * p.in_stride + p.tap_dx[tap],
                      ay0 * p.in_stride + p.tap_dy[tap], b0, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: D^T[128 cout, 256 pixels] += W_tile . X_tile^T =====
      constexpr uint32_t idesc = make_idesc_bf16(TT_BLOCK_M, TT_BLOCK_N);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        const uint32_t buf = ti & 1;
        mbar_wait(&tmem_empty_bar[buf], ((ti >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * TT_BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint32_t w_addr = smem_u32(smem + s * TT_STAGE_BYTES);
          const uint64_t da = make_sw128_desc(w_addr);
          const uint64_t db = make_sw128_desc(w_addr + TT_W_BYTES);
          if (!(p.debug & 4)) {
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k)
              umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb != 0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ===== epilogue: warps 2..5; TMEM lane = output channel n0 + 32*quarter + lane, column = pixel of the tile =====
    // Per 32-pixel chunk a warp holds a [32 channels x 32 pixels] block, one channel per lane.  scale / bias /
    // activation are applied there (per-thread scalars), the block is transposed through a padded per-warp staging
    // tile in shared memory ([pixel][channel] f32, row stride 36 floats: conflict-free both ways), and leaves as
    // 16-byte vectors: lane -> pixel 8*i + lane/4, channels 8*(lane%4) .. +8 (residual added on the way out).
    const int quarter = warp & 3;
    const int chunk_par = (warp - 2) >> 2;                    // this warp takes the 32-pixel chunks of that parity
    float* stage = reinterpret_cast<float*>(smem + TT_EPI_OFFSET) + (warp - 2) * (16 * TT_EPI_STRIDE);
    const int tl = 31 - __clz(p.tw), hl = 31 - __clz(p.th);   // tw, th are powers of two
    const int img_shift = tl + hl;                             // pixels per sample inside a tile = 1 << img_shift (>= 32)
    const float gain = p.act_gain, lrelu = p.act == 2 ? 0.01f : 0.2f;
    const int pl = lane >> 2, cseg = (lane & 3) * 8;           // output role: pixel within a group of 8, channel segment
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      int b0, ay0, ax0, n0;
      tile_coords(tile, b0, ay0, ax0, n0);
      const uint32_t buf = ti & 1;
      const int n = n0 + quarter * 32 + lane;
      const float bias = p.bias ? __ldg(p.bias + n) : 0.f;
      const float slope = p.act == 3 ? __ldg(p.slope + n) : lrelu;
      float osc = 1.f;
      if (p.residual != nullptr) {
        // the tile's residual rows are pulled into L2 while this warp would only wait for the accumulator
        const int esz = OUT_F32 ? 4 : 2;
        for (int c0 = 32 * chunk_par; c0 < TT_BLOCK_N; c0 += 64) {
          const int b = b0 + (c0 >> img_shift);
          if (b >= p.batch) break;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int q = (c0 + 8 * i + pl) & ((1 << img_shift) - 1);
            const int ay = ay0 + (q >> tl), ax = ax0 + (q & (p.tw - 1));
            if (ay >= p.grid_h || ax >= p.grid_w) continue;
            const int oy = ay * p.out_stride + p.out_off_y, ox = ax * p.out_stride + p.out_off_x;
            const int64_t o = ((static_cast<int64_t>(b) * p.hout + oy) * p.wout + ox) * p.cout + n0 + quarter * 32 + cseg;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const char*>(p.residual) + o * esz));
          }
        }
      }
      mbar_wait(&tmem_full_bar[buf], (ti >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = 32 * chunk_par; c0 < TT_BLOCK_N; c0 += 64) {
        if (p.debug & 2) break;
        const int b = b0 + (c0 >> img_shift);                  // a 32-pixel chunk never straddles two samples
        if (b >= p.batch) break;                               // warp-uniform: the remaining chunks are padding
        // this lane's four output pixels of the chunk: offsets first, so that the residual loads are in flight while
        // the accumulator block is read and transposed
        int64_t off[4];
        uint4 res[4][OUT_F32 ? 2 : 1];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int q = (c0 + 8 * i + pl) & ((1 << img_shift) - 1);
          const int ay = ay0 + (q >> tl), ax = ax0 + (q & (p.tw - 1));
          const int oy = ay * p.out_stride + p.out_off_y, ox = ax * p.out_stride + p.out_off_x;
          off[i] = (ay >= p.grid_h || ax >= p.grid_w)
                       ? -1
                       : ((static_cast<int64_t>(b) * p.hout + oy) * p.wout + ox) * p.cout + n0 + quarter * 32 + cseg;
          if (p.residual != nullptr && off[i] >= 0) {
            if (OUT_F32) {
              const uint4* rs = reinterpret_cast<const uint4*>(static_cast<const float*>(p.residual) + off[i]);
              res[i][0] = __ldg(rs);
              res[i][OUT_F32 ? 1 : 0] = __ldg(rs + 1);
            } else {
              res[i][0] = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + off[i]));
            }
          }
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * TT_BLOCK_N + c0, v);
        osc = p.out_scale ? __ldg(p.out_scale + static_cast<int64_t>(b) * p.cout + n) : 1.f;
        if (p.act == 3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float f = __uint_as_float(v[j]) * osc + bias;
            v[j] = __float_as_uint(f < 0.f ? f * slope : f);
          }
        } else if (p.act != 0) {
          // leaky_relu(x) * gain == leaky_relu(x * gain) for gain > 0, and leaky_relu(f) == max(f, slope * f) for slope < 1
          const float sg = osc * gain, bg = bias * gain;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float f = __uint_as_float(v[j]) * sg + bg;
            v[j] = __float_as_uint(fmaxf(f, slope * f));
          }
        } else if (p.out_scale != nullptr || p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * osc + bias);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if ((i & 1) == 0) {                                  // stage the next 16 pixels of the chunk
            __syncwarp();                                      // earlier reads of the staging tile are done
#pragma unroll
            for (int j = 0; j < 16; ++j) stage[j * TT_EPI_STRIDE + lane] = __uint_as_float(v[8 * i + j]);
            __syncwarp();
          }
          const int pj = 8 * (i & 1) + pl;                     // staged row of the pixel this lane writes
          const float4 lo = *reinterpret_cast<const float4*>(stage + pj * TT_EPI_STRIDE + cseg);
          const float4 hi = *reinterpret_cast<const float4*>(stage + pj * TT_EPI_STRIDE + cseg + 4);
          if (off[i] < 0 || (p.debug & 1)) continue;
          float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
          if (OUT_F32) {
            float* dst = static_cast<float*>(p.y) + off[i];
            if (p.residual) {
              const uint4 r0 = res[i][0], r1 = res[i][OUT_F32 ? 1 : 0];
              f[0] += __uint_as_float(r0.x); f[1] += __uint_as_float(r0.y); f[2] += __uint_as_float(r0.z);
              f[3] += __uint_as_float(r0.w); f[4] += __uint_as_float(r1.x); f[5] += __uint_as_float(r1.y);
              f[6] += __uint_as_float(r1.z); f[7] += __uint_as_float(r1.w);
            }
            reinterpret_cast<float4*>(dst)[0] = make_float4(f[0], f[1], f[2], f[3]);
            reinterpret_cast<float4*>(dst)[1] = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            if (p.residual) {
              const uint32_t rw[4] = {res[i][0].x, res[i][0].y, res[i][0].z, res[i][0].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                f[2 * t] += __uint_as_float(rw[t] << 16);
                f[2 * t + 1] += __uint_as_float(rw[t] & 0xffff0000u);
              }
            }
            uint4 o;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(f[2], f[3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(f[6], f[7]);
            o.x = *reinterpret_cast<uint32_t*>(&t0);
            o.y = *reinterpret_cast<uint32_t*>(&t1);
            o.z = *reinterpret_cast<uint32_t*>(&t2);
            o.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.y) + off[i]) = o;
          }
        }
      }
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

template <bool OUT_F32>
static int launch_tct(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, cudaStream_t st) {
  auto kern = conv_tct_kernel<OUT_F32>;
  static bool configured = false;
  if (!configured) {
    TE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM_TOTAL));
    configured = true;
  }
  const int total = p.n_tiles * (p.cout / TT_BLOCK_M);
  const int grid = total < kNumSMs ? total : kNumSMs;
  kern<<<grid, TT_THREADS, TT_SMEM_TOTAL, st>>>(mx, mw, p);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

}  // namespace te
