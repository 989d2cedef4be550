// Tensor-core implicit-GEMM convolution for sm_100a: TMA -> shared memory (128B swizzle) ->
// tcgen05.mma (bf16 x bf16 -> f32 accumulators in TMEM) -> tcgen05.ld epilogue.
//
// GEMM view (stride 1, "same" padding, channels-last):
//     D[pixel, cout] = sum over taps (ky,kx) and 64-channel chunks of
//                      X[b, y+ky-p, x+kx-p, c0:c0+64]  .  W[tap, cout, c0:c0+64]^T
//   * M tile = 128 output pixels = an (NB x TH x TW) patch; its A operand for one tap is ONE TMA box
//     of the 4-D tensor [B,H,W,C] at shifted coordinates.  Out-of-bounds rows are zero-filled by
//     the TMA unit, which IS the convolution's zero padding: no im2col buffer, no halo logic.
//   * B operand = a [BLOCK_N x 64] box of the tap-major weight tensor; every CTA reads the same
//     weights, so they stay L2-resident (the point of the shared-weight formulation).
//   * K loop = taps x (Cin/64); each step is 4 x tcgen05.mma (M128, N=BLOCK_N, K16).
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2-5 = epilogue
//     (TMEM lane quarter = warp_id % 4).  2 CTAs/SM are co-resident (3 stages each) so one CTA's
//     epilogue overlaps the other's main loop.
//   * Epilogue: out = act(acc * out_scale[b,cout] + bias[cout]) -> bf16 (or f32), 16-byte stores.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace te {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;        // bf16 elements = 128 bytes = one swizzle row
constexpr int TC_UMMA_K = 16;
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;  // 16 KB

struct TcParams {
  int batch, h, w, cin, cout;
  int kh, kw, pad;
  int tw, th, nb;          // tile patch; nb*th*tw == 128
  int tiles_w, tiles_h, tiles_b, n_tiles;
  int taps_per_sample;     // 0: shared weights; else kh*kw (weights indexed b*taps + tap)
  int act;
  const float* out_scale;  // [B, cout] or null
  const float* bias;       // [cout] or null
  void* y;
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, f32 accumulate, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, canonical value 1) |
// SBO>>4 [32,46) = 1024 B between 8-row groups | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16
// [10,13)=1, both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

template <int BLOCK_N>
struct TcSmem {
  static constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = TC_STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 128 + 1024;  // barriers + slack for 1024-B alignment
};

template <int BLOCK_N, bool OUT_F32>
__global__ void __launch_bounds__(TC_THREADS, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
               const TcParams p) {
  using S = TcSmem<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* tmem_full_bar = empty_bar + TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates (n fastest so neighbouring CTAs share the activation tile through L2)
  const int n_blocks = p.cout / BLOCK_N;
  const int n_blk = blockIdx.x % n_blocks;
  int m_blk = blockIdx.x / n_blocks;
  const int tile_w = m_blk % p.tiles_w; m_blk /= p.tiles_w;
  const int tile_h = m_blk % p.tiles_h; m_blk /= p.tiles_h;
  const int b0 = m_blk * p.nb, h0 = tile_h * p.th, w0 = tile_w * p.tw, n0 = n_blk * BLOCK_N;
  const int k_chunks = p.cin / TC_BLOCK_K;
  const int num_kb = p.kh * p.kw * k_chunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation: whole warp, power-of-two columns >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(BLOCK_N))
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
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int tap = kb / k_chunks, kc = kb - tap * k_chunks;
        const int ky = tap / p.kw, kx = tap - ky * p.kw;
        uint8_t* a_dst = smem + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + TC_A_BYTES;
        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
        tma_load_4d(a_dst, &map_x, &full_bar[s], kc * TC_BLOCK_K, w0 + kx - p.pad, h0 + ky - p.pad, b0);
        const int wtap = p.taps_per_sample ? b0 * p.taps_per_sample + tap : tap;
        tma_load_3d(b_dst, &map_w, &full_bar[s], kc * TC_BLOCK_K, n0, wtap);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_bf16(TC_BLOCK_M, BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint64_t da = make_sw128_desc(a_addr);
        const uint64_t db = make_sw128_desc(a_addr + TC_A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int per_img = p.th * p.tw;
    const int nbi = row / per_img;
    const int rem = row - nbi * per_img;
    const int thi = rem / p.tw, twi = rem - thi * p.tw;
    const int b = b0 + nbi, yy = h0 + thi, xx = w0 + twi;
    const bool valid = b < p.batch && yy < p.h && xx < p.w;
    const int bs = valid ? b : 0;
    const float* osc = p.out_scale ? p.out_scale + static_cast<int64_t>(bs) * p.cout + n0 : nullptr;
    const float* bias = p.bias ? p.bias + n0 : nullptr;
    const int64_t pix = (static_cast<int64_t>(bs) * p.h + yy) * p.w + xx;

    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v);
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float a = __uint_as_float(v[j]);
        if (osc) a *= __ldg(osc + c0 + j);
        if (bias) a += __ldg(bias + c0 + j);
        if (p.act == 1) a = (a > 0.f ? a : 0.2f * a) * 1.4142135623730951f;
        f[j] = a;
      }
      if (valid) {
        if (OUT_F32) {
          float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.y) + pix * p.cout + n0 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
          uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.y) + pix * p.cout + n0 + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
            o.x = *reinterpret_cast<uint32_t*>(&t0);
            o.y = *reinterpret_cast<uint32_t*>(&t1);
            o.z = *reinterpret_cast<uint32_t*>(&t2);
            o.w = *reinterpret_cast<uint32_t*>(&t3);
            dst[j] = o;
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(BLOCK_N))
                 : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &ptr, 12000, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  return fn;
}

static int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("conv2d_tc: cuTensorMapEncodeTiled entry point unavailable");
    return TE_ERR_CUDA;
  }
  uint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv2d_tc: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return TE_ERR_CUDA;
  }
  return TE_OK;
}

template <int BLOCK_N, bool OUT_F32>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, cudaStream_t st) {
  using S = TcSmem<BLOCK_N>;
  auto kern = conv_tc_kernel<BLOCK_N, OUT_F32>;
  static bool configured = false;
  if (!configured) {
    TE_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  const int grid = p.n_tiles * (p.cout / BLOCK_N);
  kern<<<grid, TC_THREADS, S::TOTAL, st>>>(mx, mw, p);
  TE_CHECK_LAUNCH();
  return TE_OK;
}

static int conv_tc_dispatch(void* y, const void* x, const void* w, const float* out_scale, const float* bias,
                            int batch, int h, int wd, int cin, int cout, int kh, int kw, int act,
                            int64_t w_bstride, bool out_f32, cudaStream_t st) {
  TE_CHECK_ARG(y && x && w, "conv2d_tc: null tensor pointer");
  TE_CHECK_ARG(cin % 64 == 0 && cout % 64 == 0, "conv2d_tc: Cin and Cout must be multiples of 64 (got %d, %d)", cin, cout);
  TE_CHECK_ARG(kh == kw && (kh == 1 || kh == 3), "conv2d_tc: kernel must be 1x1 or 3x3");
  TE_CHECK_ARG((h & (h - 1)) == 0 && (wd & (wd - 1)) == 0 && h >= 4 && wd >= 4 && h * wd >= 16,
               "conv2d_tc: H and W must be powers of two >= 4 (got %d x %d)", h, wd);
  TE_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(y) & 15) == 0, "conv2d_tc: pointers must be 16-byte aligned");
  TcParams p;
  p.batch = batch; p.h = h; p.w = wd; p.cin = cin; p.cout = cout; p.kh = kh; p.kw = kw; p.pad = kh / 2;
  p.tw = wd < 16 ? wd : 16;
  p.th = (TC_BLOCK_M / p.tw) < h ? (TC_BLOCK_M / p.tw) : h;
  p.nb = TC_BLOCK_M / (p.tw * p.th);
  p.tiles_w = wd / p.tw; p.tiles_h = h / p.th; p.tiles_b = (batch + p.nb - 1) / p.nb;
  p.n_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  p.act = act; p.out_scale = out_scale; p.bias = bias; p.y = y;
  const int taps = kh * kw;
  p.taps_per_sample = 0;
  if (w_bstride != 0) {
    TE_CHECK_ARG(w_bstride == int64_t(taps) * cout * cin, "conv2d_tc: per-sample weights must be densely packed");
    TE_CHECK_ARG(p.nb == 1, "conv2d_tc: per-sample weights need H*W >= 128");
    p.taps_per_sample = taps;
  }
  const int block_n = (cout % 128 == 0) ? 128 : 64;

  CUtensorMap mx, mw;
  {
    uint64_t dims[4] = {uint64_t(cin), uint64_t(wd), uint64_t(h), uint64_t(batch)};
    uint64_t strides[3] = {uint64_t(cin) * 2, uint64_t(wd) * cin * 2, uint64_t(h) * wd * cin * 2};
    uint32_t box[4] = {TC_BLOCK_K, uint32_t(p.tw), uint32_t(p.th), uint32_t(p.nb)};
    int rc = encode_map(&mx, x, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t nw = uint64_t(taps) * (w_bstride ? batch : 1);
    uint64_t dims[3] = {uint64_t(cin), uint64_t(cout), nw};
    uint64_t strides[2] = {uint64_t(cin) * 2, uint64_t(cout) * cin * 2};
    uint32_t box[3] = {TC_BLOCK_K, uint32_t(block_n), 1};
    int rc = encode_map(&mw, w, 3, dims, strides, box);
    if (rc) return rc;
  }
  if (block_n == 128)
    return out_f32 ? launch_tc<128, true>(mx, mw, p, st) : launch_tc<128, false>(mx, mw, p, st);
  return out_f32 ? launch_tc<64, true>(mx, mw, p, st) : launch_tc<64, false>(mx, mw, p, st);
}

}  // namespace te

extern "C" int te_conv2d_tc(void* y, const void* x, const void* w, const float* out_scale, const float* bias,
                            int batch, int hin, int win, int cin, int cout, int kh, int kw, int act,
                            int64_t w_bstride, void* stream) {
  return te::conv_tc_dispatch(y, x, w, out_scale, bias, batch, hin, win, cin, cout, kh, kw, act, w_bstride, false,
                              static_cast<cudaStream_t>(stream));
}

// D[M,N] f32 = A[M,K] * B[N,K]^T through the same kernel, as a 1x1 convolution over an [M/16, 16] "image".
extern "C" int te_gemm_tc_selftest(float* d, const void* a, const void* b, int m, int n, int k, void* stream) {
  using namespace te;
  TE_CHECK_ARG(m % 128 == 0 && m >= 128, "gemm_tc_selftest: M must be a multiple of 128");
  int h = m / 16;
  TE_CHECK_ARG((h & (h - 1)) == 0, "gemm_tc_selftest: M/16 must be a power of two");
  return conv_tc_dispatch(d, a, b, nullptr, nullptr, 1, h, 16, k, n, 1, 1, 0, 0, true,
                          static_cast<cudaStream_t>(stream));
}
