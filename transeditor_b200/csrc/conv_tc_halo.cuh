// Haloed-patch variant of the 2-CTA tcgen05 implicit-GEMM convolution (stride-1 geometries).
//
// conv_tc2_kernel fetches one [128 anchors x 64 channels] TMA box PER TAP: a 3x3 layer pulls every activation byte
// nine times from L2 into shared memory (ncu, round 1: 4.04 GB of L2->SM traffic for a 268 MB input, tensor pipe
// 53 % active — the shared-memory port is busy with TMA writes).  Here the M tile is a 16-row x 8-column patch of
// anchors and ONE box of [18 rows x 16 columns x 64 channels] (36 KB) per 64-channel chunk carries the patch plus
// its halo; each tap's A operand is a SHIFTED VIEW of that box:
//     row r = (ty, tx) of the tile  ->  box pixel (ty + 1 + dy, tx + 1 + dx)
//   * the box is 16 pixels wide (a 2048-byte image row = two 1024-byte swizzle atoms), so the eight pixels of one
//     tile row are eight consecutive 128-byte rows and the UMMA descriptor walks image rows with
//     stride-byte-offset 2048: one K-major SWIZZLE_128B descriptor per tap, start address
//     box + ((1 + dy) * 16 + 1 + dx) * 128, matrix base offset (1 + dx) & 7 (the start is not 1024-byte aligned).
//   * A-operand fill traffic drops 4x (36 KB per chunk instead of 9 x 16 KB); the weight tiles keep their own ring.
//   * Split-operand mode: the A box of activation plane `ap` is reused for every weight plane it pairs with
//     (hi: hi, mid(, lo); mid: hi(, mid); lo: hi), so A traffic per f32 product drops further.
// Protocol, tile pairing, TMEM accumulator slots and the epilogue are those of conv_tc2_kernel.
#pragma once

namespace te {

constexpr int TH_TILE_H = 16, TH_TILE_W = 8;
constexpr int TH_BOX_H = TH_TILE_H + 2, TH_BOX_W = 16;
constexpr int TH_A_BYTES = TH_BOX_H * TH_BOX_W * 128;     // 36864
constexpr int TH_A_STAGES = 3;

template <int BLOCK_N>
struct TcHaloSmem {
  static constexpr int B_BYTES = (BLOCK_N / 2) * TC_BLOCK_K * 2;   // this CTA's half of one weight tile
  static constexpr int B_STAGES = (TC_SMEM_BUDGET - TH_A_STAGES * TH_A_BYTES) / B_BYTES;   // 10 at N=128, 5 at N=256
  static constexpr int B_OFFSET = TH_A_STAGES * TH_A_BYTES;
  static constexpr int BAR_OFFSET = B_OFFSET + B_STAGES * B_BYTES;
  static constexpr int EPI_OFFSET = BAR_OFFSET + 512;
  static constexpr int TOTAL = EPI_OFFSET + 4 * 2 * BLOCK_N * 4 + 1024;
};

// K-major SWIZZLE_128B descriptor of a shifted view: 8-row groups 2048 bytes apart, start not 1024-aligned
__device__ __forceinline__ uint64_t make_sw128_desc_view(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int BLOCK_N, bool OUT_F32, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
conv_tc2h_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ TcParams p) {
  using S = TcHaloSmem<BLOCK_N>;
  using ACC = TcAcc<BLOCK_N, SPLIT>;
  constexpr int AS = TH_A_STAGES, BS = S::B_STAGES;
  constexpr uint32_t TMEM_COLS = ACC::COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);   // leader only
  uint64_t* a_empty = a_full + AS;
  uint64_t* b_full = a_empty + AS;                                        // leader only
  uint64_t* b_empty = b_full + BS;
  uint64_t* tmem_full_bar = b_empty + BS;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2], leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_blocks = (p.cout + BLOCK_N - 1) / BLOCK_N;
  const int m_pairs = (p.n_tiles + 1) / 2;
  const int total_work = m_pairs * n_blocks;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int k_chunks = (p.cin + TC_BLOCK_K - 1) / TC_BLOCK_K;
  const int items_per_work = k_chunks * p.nseg;          // A boxes per work item: (chunk, activation plane)
  const int my_work = pair_id < total_work ? (total_work - pair_id + num_pairs - 1) / num_pairs : 0;
  const int total_items = my_work * items_per_work;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < AS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < BS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
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
    b0 = m_blk; ay0 = tile_h * TH_TILE_H; ax0 = tile_w * TH_TILE_W; n0 = n_blk * BLOCK_N;
  };
  // weight planes paired with activation plane `ap`: nseg - ap of them (hi: all; mid: all but the last; ...)
  auto n_wplanes = [&](int ap) { return p.nseg - ap; };

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): A boxes run one item ahead of the weight tiles =====
      uint32_t bit = 0;
      auto issue_a = [&](int gi) {
        const int wi = gi / items_per_work, r = gi - wi * items_per_work;
        const int kc = r / p.nseg, ap = r - kc * p.nseg;
        int b0, ay0, ax0, n0;
        work_coords(pair_id + wi * num_pairs, b0, ay0, ax0, n0);
        const int s = gi % AS;
        mbar_wait(&a_empty[s], ((gi / AS) & 1) ^ 1);
        if (p.debug & 8) {  // profiling aid: no loads, the MMAs chew on whatever is in shared memory
          if (leader) mbar_arrive(&a_full[s]);
          return;
        }
        if (leader) mbar_expect_tx(&a_full[s], 2 * TH_A_BYTES);
        tma2_load_5d(smem + s * TH_A_BYTES, &map_x, &a_full[s], kc * TC_BLOCK_K, ax0 + p.halo_x0, ay0 + p.halo_y0, b0, ap);
      };
      if (total_items > 0) issue_a(0);
      for (int gi = 0; gi < total_items; ++gi) {
        if (gi + 1 < total_items) issue_a(gi + 1);
        const int wi = gi / items_per_work, r = gi - wi * items_per_work;
        const int kc = r / p.nseg, ap = r - kc * p.nseg;
        int b0, ay0, ax0, n0;
        work_coords(pair_id + wi * num_pairs, b0, ay0, ax0, n0);
        const int nwp = n_wplanes(ap);
        for (int wp = 0; wp < nwp; ++wp) {
          for (int tap = 0; tap < p.ntaps; ++tap, ++bit) {
            const int s = bit % BS;
            mbar_wait(&b_empty[s], ((bit / BS) & 1) ^ 1);
            if (p.debug & 8) {
              if (leader) mbar_arrive(&b_full[s]);
              continue;
            }
            if (leader) mbar_expect_tx(&b_full[s], 2 * S::B_BYTES);
            const int wsl = p.w_slices_per_sample ? b0 * p.w_slices_per_sample + p.tap_w[tap] : p.tap_w[tap];
            tma2_load_4d(smem + S::B_OFFSET + s * S::B_BYTES, &map_w, &b_full[s], kc * TC_BLOCK_K,
                         n0 + int(rank) * (BLOCK_N / 2), wsl, wp);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader only): M256 x BLOCK_N x K16 across the pair =====
      constexpr uint32_t idesc = make_idesc_bf16(2 * TC_BLOCK_M, BLOCK_N);
      uint32_t bit = 0;
      int gi = 0;
      for (int wi = 0; wi < my_work; ++wi) {
        const uint32_t ti = wi;
        const uint32_t buf = ACC::NBUF == 2 ? (ti & 1) : 0;
        mbar_wait(&tmem_empty_bar[buf], (ACC::NBUF == 2 ? ((ti >> 1) & 1) : (ti & 1)) ^ 1);
        tcgen05_fence_after();
        const uint32_t d_base = tmem_base + buf * ACC::SLOTS * BLOCK_N;
        uint32_t started = 0;   // bit s: accumulator slot s has received its first MMA of this tile
        int main_idx = 0;
        for (int it = 0; it < items_per_work; ++it, ++gi) {
          const int ap = it % p.nseg;
          const int sa = gi % AS;
          mbar_wait(&a_full[sa], (gi / AS) & 1);
          tcgen05_fence_after();
          const uint32_t a_box = smem_u32(smem + sa * TH_A_BYTES);
          const int nwp = n_wplanes(ap);
          for (int wp = 0; wp < nwp; ++wp) {
            for (int tap = 0; tap < p.ntaps; ++tap, ++bit) {
              const int sb = bit % BS;
              mbar_wait(&b_full[sb], (bit / BS) & 1);
              tcgen05_fence_after();
              const int px = p.tap_dx[tap] - p.halo_x0, py = p.tap_dy[tap] - p.halo_y0;   // box pixel of anchor (0, 0)
              const uint32_t a_addr = a_box + (py * TH_BOX_W + px) * 128;
              const uint64_t da = make_sw128_desc_view(a_addr, TH_BOX_W * 128, p.halo_bo ? 0u : uint32_t(px));
              const uint64_t db = make_sw128_desc(smem_u32(smem + S::B_OFFSET + sb * S::B_BYTES));
              int slot = 0;
              if (SPLIT) {
                slot = 3;
                if (ap == 0 && wp == 0) { slot = main_idx % 3; ++main_idx; }
              }
              const bool fresh = ((started >> slot) & 1u) == 0;
              started |= 1u << slot;
              if (!(p.debug & 4)) {
#pragma unroll
                for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k)
                  umma2_bf16(d_base + slot * BLOCK_N, da + 2 * k, db + 2 * k, idesc, (!fresh || k != 0) ? 1u : 0u);
              }
              umma2_commit_both(&b_empty[sb]);
            }
          }
          umma2_commit_both(&a_empty[sa]);
        }
        umma2_commit_both(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ===== epilogue (both CTAs): own 128 TMEM lanes; row r = (ty, tx) = (r / 8, r % 8) =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int thi = row / TH_TILE_W, twi = row - thi * TH_TILE_W;
    float* s_osc = reinterpret_cast<float*>(smem + S::EPI_OFFSET) + quarter * 2 * BLOCK_N;
    float* s_bias = s_osc + BLOCK_N;
    const int num_main = k_chunks * p.ntaps;   // hi*hi k-blocks of a tile
    for (int wi = 0; wi < my_work; ++wi) {
      const uint32_t ti = wi;
      int b0, ay0, ax0, n0;
      work_coords(pair_id + wi * num_pairs, b0, ay0, ax0, n0);
      const uint32_t buf = ACC::NBUF == 2 ? (ti & 1) : 0;
      const int ay = ay0 + thi, ax = ax0 + twi;
      const bool valid = b0 < p.batch && ay < p.grid_h && ax < p.grid_w;
      const int bs = valid ? b0 : 0;
      {
        __syncwarp();
        const int bq = b0 < p.batch ? b0 : 0;
        for (int c = lane; c < BLOCK_N; c += 32) {
          const int n = n0 + c;
          s_osc[c] = (p.out_scale && n < p.cout) ? __ldg(p.out_scale + static_cast<int64_t>(bq) * p.cout + n) : 1.f;
          s_bias[c] = (p.bias && n < p.cout) ? __ldg(p.bias + n) : 0.f;
        }
        __syncwarp();
      }
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
          tmem_ld32(t_row + 3 * BLOCK_N, v);
          const int n_main = num_main < 3 ? num_main : 3;
          for (int sl = 0; sl < n_main; ++sl) {
            uint32_t u[32];
            tmem_ld32(t_row + sl * BLOCK_N, u);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
          }
        } else {
          tmem_ld32(t_row, v);
        }
        tc_epilogue_chunk<OUT_F32>(p, v, s_osc + c0, s_bias + c0, true, nullptr, valid, pix, n0 + c0);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[buf]);
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BLOCK_N, bool OUT_F32, bool SPLIT = false>
static int launch_tc2h(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, cudaStream_t st) {
  using S = TcHaloSmem<BLOCK_N>;
  auto kern = conv_tc2h_kernel<BLOCK_N, OUT_F32, SPLIT>;
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

}  // namespace te
