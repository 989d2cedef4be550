// Tensor-core weight gradient for sm_100a (tcgen05, bf16 operands, f32 accumulate).
//
// For the anchor/tap geometry of te_tc_conv_desc (see conv_tc.cu):
//     gw[w_t][m][n] += SUM_{b, anchors a}  g[b, a*os + oo, m] * x[b, a*is + d_t, n]
// i.e. per tap a GEMM  D[cout, cin] = G^T . X  whose reduction dimension is the PIXELS.  Both operands are
// read by TMA as [anchor rows x 64 channels] boxes straight from the channels-last tensors, which makes
// them MN-major UMMA operands (the channel dimension is contiguous): no transpose pass exists anywhere.
//   * CTA tile: 128 (cout) x 128 (cin), up to 3 taps resident in TMEM (3 x 128 of the 512 columns).  A 128 x 256 tile
//     with 2 taps resident exists behind TE_WG_N=256 (the N = 256 instruction carries twice the FLOPs per issue slot)
//     but measured slower: see the host code.
//   * stage = 64 anchors: ONE G tile (2 x [64 x 64ch]) shared by the taps + one shifted X tile per tap
//     (2 x [64 x 64ch] each) = 16 + 3*16 = 64 KB, 3-stage ring: G is fetched once per anchor tile.
//   * 4 x tcgen05.mma (M128, N128, K16 anchors) per tap per stage; taps are balanced over grid.y
//     (9 taps -> 3 groups of 3, 4 -> 2 x 2).
//   * split-K over anchor tiles (grid.z) so the chip is filled even when Cout*Cin is one tile;
//     partial sums are combined with vectorised f32 reductions (red.global.add.v4.f32).
//   * Out-of-range anchors / padding taps / channel tails are zero-filled by the TMA unit.
//   * Split-operand mode (te_tc_conv_desc.split = 2 or 3, see conv_tc.cu): g and x are bf16 plane stacks of f32
//     values; every anchor tile is walked once per plane pair, all pairs accumulating into the same TMEM tiles.
#include <stdlib.h>

#include "tc_common.cuh"

namespace te {

constexpr int WG_M = 128;
constexpr int WG_KA = 64;              // anchors per stage
constexpr int WG_THREADS = 192;
constexpr int WG_HALF_BYTES = WG_KA * 128;          // one [64 anchors x 64 ch] box = 8 KB
constexpr int WG_TILE_BYTES = 2 * WG_HALF_BYTES;    // G tile [64 anchors x 128 ch] = 16 KB

// Haloed X box of the HALO variant: a 4 x 16 anchor tile plus one column either side, [4][18] pixels x 64 channels per
// half = 72 rows of 128 B (padded to a multiple of the 1024-byte swizzle atom); the three taps of a group (same dy,
// dx = d0, d0+1, d0+2) are row-shifted VIEWS of it.
constexpr int WG_HALO_W = 18, WG_HALO_H = 4;
constexpr int WG_HALO_HALF = ((WG_HALO_W * WG_HALO_H * 128 + 1023) / 1024) * 1024;   // 10240

// WN = cin columns per CTA tile (128 or 256)
template <int WN, bool HALO = false> struct WgCfg {
  static constexpr int TAPS = 512 / WN > 3 ? 3 : 512 / WN;       // taps resident in TMEM: 3 (N=128), 2 (N=256)
  static constexpr int X_BYTES = (WN / 64) * WG_HALF_BYTES;      // one tap's X tile
  static constexpr int STAGE_BYTES = HALO ? WG_TILE_BYTES + (WN / 64) * WG_HALO_HALF   // G + ONE haloed X box: 36 KB
                                          : WG_TILE_BYTES + TAPS * X_BYTES;            // G + TAPS X tiles: 64 / 80 KB
  static constexpr int STAGES = HALO ? 5 : (WN == 128 ? 3 : 2);
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int SMEM_TOTAL = BAR_OFFSET + 128 + 1024;
};

struct WgParams {
  int batch, cin, cout;
  int ntaps, taps_per_group;
  int tap_dy[9], tap_dx[9], tap_w[9];
  int in_stride, out_stride, out_off_y, out_off_x;
  int tw, th, nb;                 // anchor tile patch; nb*th*tw == 64
  int tiles_w, tiles_h, tiles_b, n_tiles;
  int tiles_per_split;
  int splits_per_sample;          // > 0: per-sample gradients (grid.z = batch * splits_per_sample)
  int tiles_per_sample;
  int64_t gw_bstride;             // elements between two samples' gradients (per-sample mode)
  int use_atomics;
  int npairs, pair_g[6], pair_x[6];   // split-operand plane pairs (1 pair = plain bf16)
  float alpha;                    // gw += alpha * partial
  float* gw;                      // [w_slices, cout, cin]
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int WG_N, bool HALO>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x,
                const __grid_constant__ WgParams p) {
  using C = WgCfg<WG_N, HALO>;
  constexpr int WG_STAGES = C::STAGES, WG_STAGE_BYTES = C::STAGE_BYTES, WG_BAR_OFFSET = C::BAR_OFFSET;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_BAR_OFFSET);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* tmem_full_bar = empty_bar + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = (p.cin + WG_N - 1) / WG_N;
  const int m0 = (blockIdx.x / n_blocks) * WG_M, n0 = (blockIdx.x % n_blocks) * WG_N;
  const int tap0 = blockIdx.y * p.taps_per_group;
  const int ntap = min(p.taps_per_group, p.ntaps - tap0);
  int tile_lo, tile_hi;
  float* gw_base = p.gw;
  if (p.splits_per_sample > 0) {  // every sample owns its own gradient slice
    const int smp = blockIdx.z / p.splits_per_sample, part = blockIdx.z % p.splits_per_sample;
    tile_lo = smp * p.tiles_per_sample + part * p.tiles_per_split;
    tile_hi = min((smp + 1) * p.tiles_per_sample, tile_lo + p.tiles_per_split);
    gw_base += smp * p.gw_bstride;
  } else {
    tile_lo = blockIdx.z * p.tiles_per_split;
    tile_hi = min(p.n_tiles, tile_lo + p.tiles_per_split);
  }
  const int num_it = max(0, tile_hi - tile_lo) * p.npairs;  // pipeline iterations: anchor tiles x plane pairs

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_g)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int it = 0; it < num_it; ++it) {
        const int s = it % WG_STAGES;
        const uint32_t ph = (it / WG_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int pr = it % p.npairs;
        const int sg = p.pair_g[pr], sx = p.pair_x[pr];
        int t = tile_lo + it / p.npairs;
        const int tile_w = t % p.tiles_w; t /= p.tiles_w;
        const int tile_h = t % p.tiles_h; t /= p.tiles_h;
        const int b0 = t * p.nb, ay0 = tile_h * p.th, ax0 = tile_w * p.tw;
        uint8_t* dst = smem + s * WG_STAGE_BYTES;
        const int gx = ax0 * p.out_stride + p.out_off_x, gy = ay0 * p.out_stride + p.out_off_y;
        if (HALO) {
          // ONE haloed box per 64-channel half serves every tap of the group (same dy, consecutive dx)
          mbar_expect_tx(&full_bar[s], WG_TILE_BYTES + (WG_N / 64) * (WG_HALO_W * WG_HALO_H * 128));
          tma_load_5d(dst, &map_g, &full_bar[s], m0, gx, gy, b0, sg);
          tma_load_5d(dst + WG_HALF_BYTES, &map_g, &full_bar[s], m0 + 64, gx, gy, b0, sg);
          const int xx = ax0 + p.tap_dx[tap0], xy = ay0 + p.tap_dy[tap0];
#pragma unroll
          for (int h = 0; h < WG_N / 64; ++h)
            tma_load_5d(dst + WG_TILE_BYTES + h * WG_HALO_HALF, &map_x, &full_bar[s], n0 + 64 * h, xx, xy, b0, sx);
          continue;
        }
        mbar_expect_tx(&full_bar[s], WG_TILE_BYTES + ntap * C::X_BYTES);
        tma_load_5d(dst, &map_g, &full_bar[s], m0, gx, gy, b0, sg);
        tma_load_5d(dst + WG_HALF_BYTES, &map_g, &full_bar[s], m0 + 64, gx, gy, b0, sg);
        for (int tp = 0; tp < ntap; ++tp) {
          const int tap = tap0 + tp;
          uint8_t* xd = dst + WG_TILE_BYTES + tp * C::X_BYTES;
          const int xx = ax0 * p.in_stride + p.tap_dx[tap], xy = ay0 * p.in_stride + p.tap_dy[tap];
#pragma unroll
          for (int h = 0; h < WG_N / 64; ++h)
            tma_load_5d(xd + h * WG_HALF_BYTES, &map_x, &full_bar[s], n0 + 64 * h, xx, xy, b0, sx);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_bf16_mn(WG_M, WG_N);
      for (int it = 0; it < num_it; ++it) {
        const int s = it % WG_STAGES;
        const uint32_t ph = (it / WG_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t base = smem_u32(smem + s * WG_STAGE_BYTES);
        for (int tp = 0; tp < ntap; ++tp) {
#pragma unroll
          for (int k = 0; k < WG_KA / 16; ++k) {
            // 16 anchors = two 8-row groups = 2048 bytes further down the tile
            const uint64_t da = make_sw128_mn_desc(base + k * 2048, WG_HALF_BYTES);
            // HALO: K16 step k = image row k of the 4 x 16 tile; the tap's view starts (k * 18 + dx - dx0) rows into the
            // box (the swizzle is a function of the absolute address, so a row-shifted start needs no base offset)
            const uint64_t db = HALO
                ? make_sw128_mn_desc(base + WG_TILE_BYTES + (k * WG_HALO_W + p.tap_dx[tap0 + tp] - p.tap_dx[tap0]) * 128,
                                     WG_HALO_HALF)
                : make_sw128_mn_desc(base + WG_TILE_BYTES + tp * C::X_BYTES + k * 2048, WG_HALF_BYTES);
            umma_bf16(tmem_base + tp * WG_N, da, db, idesc, (it | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===== epilogue =====
    const int quarter = warp & 3;
    const int m = m0 + quarter * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    if (num_it > 0) {
      for (int tp = 0; tp < ntap; ++tp) {
        float* row = gw_base + (static_cast<int64_t>(p.tap_w[tap0 + tp]) * p.cout + m) * p.cin + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < WG_N; c0 += 32) {
          if (n0 + c0 >= p.cin) break;
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + tp * WG_N + c0, v);
          if (m < p.cout) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (n0 + c0 + 4 * j >= p.cin) continue;
              float a0 = __uint_as_float(v[4 * j]) * p.alpha, a1 = __uint_as_float(v[4 * j + 1]) * p.alpha;
              float a2 = __uint_as_float(v[4 * j + 2]) * p.alpha, a3 = __uint_as_float(v[4 * j + 3]) * p.alpha;
              if (p.use_atomics)
                red_add_v4(row + c0 + 4 * j, a0, a1, a2, a3);
              else {
                float4* q = reinterpret_cast<float4*>(row + c0 + 4 * j);
                float4 old = *q;
                *q = make_float4(old.x + a0, old.y + a1, old.z + a2, old.w + a3);
              }
            }
          }
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace te

extern "C" int te_conv_wgrad_tc(float* gw, const void* g, const void* x, const te_tc_conv_desc* dp, void* stream) {
  using namespace te;
  TE_CHECK_ARG(gw && g && x && dp, "conv_wgrad_tc: null pointer");
  const te_tc_conv_desc& d = *dp;
  TE_CHECK_ARG(d.cin % 8 == 0 && d.cout % 8 == 0 && d.cin >= 8 && d.cout >= 8,
               "conv_wgrad_tc: channel counts must be multiples of 8 (got %d, %d)", d.cin, d.cout);
  TE_CHECK_ARG(d.ntaps >= 1 && d.ntaps <= 9, "conv_wgrad_tc: 1..9 taps");
  TE_CHECK_ARG((d.in_stride == 1 || d.in_stride == 2) && (d.out_stride == 1 || d.out_stride == 2),
               "conv_wgrad_tc: strides must be 1 or 2");
  TE_CHECK_ARG(d.grid_h > 0 && d.grid_w > 0 && d.batch > 0, "conv_wgrad_tc: empty anchor grid");
  TE_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(gw) & 15) == 0 && d.cin % 4 == 0,
               "conv_wgrad_tc: pointers must be 16-byte aligned");
  WgParams p;
  p.alpha = d.wgrad_alpha != 0.f ? d.wgrad_alpha : 1.f;
  p.batch = d.batch; p.cin = d.cin; p.cout = d.cout; p.ntaps = d.ntaps;
  for (int t = 0; t < 9; ++t) {
    p.tap_dy[t] = t < d.ntaps ? d.tap_dy[t] : 0;
    p.tap_dx[t] = t < d.ntaps ? d.tap_dx[t] : 0;
    p.tap_w[t] = t < d.ntaps ? d.tap_w[t] : 0;
    if (t < d.ntaps) TE_CHECK_ARG(d.tap_w[t] >= 0 && d.tap_w[t] < d.w_slices, "conv_wgrad_tc: tap weight index out of range");
  }
  p.in_stride = d.in_stride; p.out_stride = d.out_stride; p.out_off_y = d.out_off_y; p.out_off_x = d.out_off_x;
  p.tw = next_pow2(d.grid_w) < 16 ? next_pow2(d.grid_w) : 16;
  int th = WG_KA / p.tw;
  p.th = next_pow2(d.grid_h) < th ? next_pow2(d.grid_h) : th;
  p.nb = WG_KA / (p.tw * p.th);
  p.tiles_w = (d.grid_w + p.tw - 1) / p.tw;
  p.tiles_h = (d.grid_h + p.th - 1) / p.th;
  p.tiles_b = (d.batch + p.nb - 1) / p.nb;
  p.n_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  p.gw = gw;
  // cin tile: 128.  The 256-wide tile (TE_WG_N=256, Cin % 256 == 0) was measured SLOWER on every shape (512->512 @64^2
  // B=16: 1123 -> 918 TFLOP/s; 256->256 @128^2: 1148 -> 1067; tools/wgrad_probe.py, same box): with two taps resident
  // the nine taps need five passes over G instead of three and the 80 KB stages leave a 2-deep ring — this kernel is
  // bound by operand feed, not by MMA issue.
  static int force_n = -1;
  if (force_n < 0) { const char* e = getenv("TE_WG_N"); force_n = e ? atoi(e) : 0; }
  const int wg_n = (d.cin % 256 == 0 && force_n == 256) ? 256 : 128;
  const int wg_taps = wg_n == 256 ? WgCfg<256>::TAPS : WgCfg<128>::TAPS;
  const int out_tiles = ((d.cout + WG_M - 1) / WG_M) * ((d.cin + wg_n - 1) / wg_n);
  const int tap_groups = (d.ntaps + wg_taps - 1) / wg_taps;
  p.taps_per_group = (d.ntaps + tap_groups - 1) / tap_groups;  // balanced: 9 -> 3+3+3, 4 -> 2+2
  // split the anchor reduction so that about one wave of CTAs exists, at least 8 tiles per split
  int splits;
  p.splits_per_sample = 0; p.tiles_per_sample = 0; p.gw_bstride = 0;
  if (d.w_bstride != 0) {
    TE_CHECK_ARG(p.nb == 1, "conv_wgrad_tc: per-sample gradients need >= 64 anchors per sample");
    TE_CHECK_ARG(d.w_bstride == int64_t(d.w_slices) * d.cout * d.cin, "conv_wgrad_tc: per-sample gradients must be densely packed");
    p.tiles_per_sample = p.tiles_w * p.tiles_h;
    // one CTA per SM (196 KB of shared memory): keep the grid within ONE wave — 150 CTAs on 148 SMs would
    // run two waves and double the kernel time — so round the split count DOWN
    // one CTA per SM (shared memory): pick the per-sample split count with the best WAVE efficiency
    // items / (ceil(items / 148) * 148) — batch 16 x 4 tiles x 3 tap groups = 192 CTAs unsplit is 1.3 waves, the second
    // a third full (0.65); 3 splits make 576 = 3.9 waves (0.97).  Fewer splits on near-ties: each is an atomic pass.
    const int max_sps = (p.tiles_per_sample + 7) / 8;
    const int base = out_tiles * tap_groups * d.batch;
    int sps = 1;
    double best = 0.0;
    static int wave_split = -1;   // TE_WG_WAVE_SPLIT=0: the former rule (one wave, rounded down)
    if (wave_split < 0) { const char* e = getenv("TE_WG_WAVE_SPLIT"); wave_split = e ? atoi(e) : 1; }
    if (!wave_split) {
      sps = kNumSMs / base;
      if (sps > max_sps) sps = max_sps;
      if (sps < 1) sps = 1;
    }
    for (int c = 1; wave_split && c <= 12 && c <= max_sps; ++c) {
      const int items = base * c;
      const double eff = double(items) / double(((items + kNumSMs - 1) / kNumSMs) * kNumSMs);
      if (eff > best + 0.03) { best = eff; sps = c; }
    }
    p.tiles_per_split = (p.tiles_per_sample + sps - 1) / sps;
    sps = (p.tiles_per_sample + p.tiles_per_split - 1) / p.tiles_per_split;
    p.splits_per_sample = sps;
    p.gw_bstride = d.w_bstride;
    splits = sps * d.batch;
    p.use_atomics = sps > 1 ? 1 : 0;
    TE_CHECK_ARG(splits <= 65535, "conv_wgrad_tc: grid.z too large");
  } else {
    splits = kNumSMs / (out_tiles * tap_groups);  // rounded down: a single wave of CTAs
    int max_splits = (p.n_tiles + 7) / 8;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
    splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    p.use_atomics = splits > 1 ? 1 : 0;
  }

  TE_CHECK_ARG(d.split >= 0 && d.split <= 3, "conv_wgrad_tc: split must be 0/1 (plain bf16), 2 or 3 planes");
  const int nseg = d.split < 2 ? 1 : d.split;
  {
    const SplitPairs sp = split_pairs(nseg);
    p.npairs = sp.n;
    for (int i = 0; i < 6; ++i) { p.pair_g[i] = sp.a[i]; p.pair_x[i] = sp.b[i]; }
  }
  // Haloed-box variant: stride-1 input, 4 x 16 anchor tiles of ONE sample, every tap group = one row of taps with
  // consecutive dx (the 3x3 convolutions and their transposed/polyphase forms with >= 2 taps per group)
  static int use_halo = -1;
  if (use_halo < 0) { const char* e = getenv("TE_WG_HALO"); use_halo = e ? atoi(e) : 1; }
  bool halo = use_halo != 0 && wg_n == 128 && d.in_stride == 1 && p.tw == 16 && p.th == 4 && p.nb == 1 && d.ntaps >= 2;
  for (int g0 = 0; halo && g0 < d.ntaps; g0 += p.taps_per_group) {
    const int ng = d.ntaps - g0 < p.taps_per_group ? d.ntaps - g0 : p.taps_per_group;
    for (int t = 1; t < ng; ++t)
      if (d.tap_dy[g0 + t] != d.tap_dy[g0] || d.tap_dx[g0 + t] <= d.tap_dx[g0 + t - 1] ||
          d.tap_dx[g0 + t] - d.tap_dx[g0] > WG_HALO_W - 16)
        halo = false;
  }
  CUtensorMap mg, mx;
  {
    const uint32_t os = uint32_t(d.out_stride);
    uint64_t dims[5] = {uint64_t(d.cout), uint64_t(d.wout), uint64_t(d.hout), uint64_t(d.batch), uint64_t(nseg)};
    uint64_t strides[4] = {uint64_t(d.cout) * 2, uint64_t(d.wout) * d.cout * 2, uint64_t(d.hout) * d.wout * d.cout * 2,
                           uint64_t(d.batch) * d.hout * d.wout * d.cout * 2};
    uint32_t box[5] = {64, uint32_t(p.tw) * os, uint32_t(p.th) * os, uint32_t(p.nb), 1};
    uint32_t estr[5] = {1, os, os, 1, 1};
    int rc = encode_map_bf16(&mg, g, 5, dims, strides, box, estr);
    if (rc) return rc;
  }
  {
    const uint32_t is = uint32_t(d.in_stride);
    uint64_t dims[5] = {uint64_t(d.cin), uint64_t(d.win), uint64_t(d.hin), uint64_t(d.batch), uint64_t(nseg)};
    uint64_t strides[4] = {uint64_t(d.cin) * 2, uint64_t(d.win) * d.cin * 2, uint64_t(d.hin) * d.win * d.cin * 2,
                           uint64_t(d.batch) * d.hin * d.win * d.cin * 2};
    uint32_t box[5] = {64, uint32_t(p.tw) * is, uint32_t(p.th) * is, uint32_t(p.nb), 1};
    if (halo) { box[1] = WG_HALO_W; box[2] = WG_HALO_H; }
    uint32_t estr[5] = {1, is, is, 1, 1};
    int rc = encode_map_bf16(&mx, x, 5, dims, strides, box, estr);
    if (rc) return rc;
  }
  static bool configured = false;
  if (!configured) {
    TE_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       WgCfg<128>::SMEM_TOTAL));
    TE_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       WgCfg<256>::SMEM_TOTAL));
    TE_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       WgCfg<128, true>::SMEM_TOTAL));
    configured = true;
  }
  dim3 grid(out_tiles, tap_groups, splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (halo)
    wgrad_tc_kernel<128, true><<<grid, WG_THREADS, WgCfg<128, true>::SMEM_TOTAL, st>>>(mg, mx, p);
  else if (wg_n == 256)
    wgrad_tc_kernel<256, false><<<grid, WG_THREADS, WgCfg<256>::SMEM_TOTAL, st>>>(mg, mx, p);
  else
    wgrad_tc_kernel<128, false><<<grid, WG_THREADS, WgCfg<128>::SMEM_TOTAL, st>>>(mg, mx, p);
  TE_CHECK_LAUNCH();
  return TE_OK;
}
