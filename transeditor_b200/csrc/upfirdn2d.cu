// upfirdn2d: zero-insert upsample -> pad/crop -> 2-D FIR (true convolution) -> decimate.
// Semantics follow the reference operator (utils/op/upfirdn2d_kernel.cu:52-137,167-168) — restated.
//
// Two kernels:
//   * upfirdn2d_generic: any (up, down, pad, FIR <= 16x16, minor); one thread per output element
//     (minor fastest, so NHWC reads vectorise over channels and NCHW reads coalesce over x).
//     Never leaves the output uninitialised (the reference launches nothing for unmatched modes).
//   * fir_planes_tiled:  the hot case — up = down = 1, minor = 1 (NCHW planes), FIR <= 4x4:
//     every Blur of the generator / discriminator (model_spatial_query.py:137-153).  HBM-bound:
//     input tile staged once in shared memory (zero-filled halo = the padding), each thread
//     produces a 4x4 register block from a 7x7 patch (3 shared-memory words per output),
//     16-byte stores.  TMA tiled loads are not usable here: row pitches such as 257*4 B are not
//     16-byte multiples (cuTensorMapEncodeTiled requirement), so staging uses coalesced LDG.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace te {

struct UpfirdnParams {
  int64_t major;
  int in_h, in_w, minor, kh, kw;
  int up_x, up_y, down_x, down_y, pad_x0, pad_y0;
  int out_h, out_w;
};

__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return (q * b > a) ? q - 1 : q;
}

template <typename T>
__global__ void __launch_bounds__(256)
upfirdn2d_generic(T* __restrict__ out, const T* __restrict__ in, const float* __restrict__ fir,
                  UpfirdnParams p, int64_t total) {
  using A = typename Acc<T>::type;
  __shared__ float sk[256];  // flipped taps: sk[ky][kx] = fir[kh-1-ky][kw-1-kx]
  for (int t = threadIdx.x; t < p.kh * p.kw; t += blockDim.x) {
    int ky = t / p.kw, kx = t - ky * p.kw;
    sk[t] = fir[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
  }
  __syncthreads();
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int64_t r = idx;
    const int mi = int(r % p.minor); r /= p.minor;
    const int ox = int(r % p.out_w); r /= p.out_w;
    const int oy = int(r % p.out_h); r /= p.out_h;
    const int64_t mj = r;
    // position in the zero-inserted, padded grid of the first tap
    const int mid_x = ox * p.down_x + p.up_x - 1 - p.pad_x0;
    const int mid_y = oy * p.down_y + p.up_y - 1 - p.pad_y0;
    const int in_x0 = floordiv(mid_x, p.up_x);
    const int in_y0 = floordiv(mid_y, p.up_y);
    const int tap_x0 = (in_x0 + 1) * p.up_x - mid_x - 1;
    const int tap_y0 = (in_y0 + 1) * p.up_y - mid_y - 1;
    A acc = A(0);
    for (int ty = tap_y0, iy = in_y0; ty < p.kh; ty += p.up_y, ++iy) {
      if (iy < 0 || iy >= p.in_h) continue;
      const T* row = in + ((mj * p.in_h + iy) * int64_t(p.in_w)) * p.minor + mi;
      for (int tx = tap_x0, ix = in_x0; tx < p.kw; tx += p.up_x, ++ix) {
        if (ix < 0 || ix >= p.in_w) continue;
        acc += to_acc(row[int64_t(ix) * p.minor]) * A(sk[ty * p.kw + tx]);
      }
    }
    out[idx] = from_acc<T, A>(acc);
  }
}

// ---- hot case: planes, up = down = 1, FIR zero-extended to 4x4 --------------------------------
// 256 threads arranged threads_x x threads_y (chosen per launch so that odd widths such as 257 split
// into equal tiles: 257 -> 3 tiles of 88 columns, 22 x 11 threads), each thread a 4x4 output block.
constexpr int FT_THREADS = 256;
constexpr int FT_MAX_SMEM_ELEMS = 5632;  // (tile_h + 3) * pitch upper bound for every layout below

struct FirTiling {
  int threads_x, threads_y, tile_w, tile_h, pitch, tiles_x, tiles_y;
};

static FirTiling fir_tiling(int out_w, int out_h) {
  FirTiling t;
  const int nx = (out_w + 127) / 128;
  int tw = ((out_w + nx - 1) / nx + 3) / 4 * 4;  // equalised tile width, multiple of 4
  if (tw < 4) tw = 4;
  t.threads_x = tw / 4;
  t.tile_w = tw;
  t.threads_y = FT_THREADS / t.threads_x;
  int max_ty = ((out_h + 3) / 4);
  if (t.threads_y > max_ty) t.threads_y = max_ty;
  if (t.threads_y < 1) t.threads_y = 1;
  t.tile_h = t.threads_y * 4;
  t.pitch = t.tile_w + 4;
  while ((t.tile_h + 3) * t.pitch > FT_MAX_SMEM_ELEMS) {  // very narrow images: shorten the tile
    t.threads_y -= 1;
    t.tile_h = t.threads_y * 4;
  }
  t.tiles_x = (out_w + t.tile_w - 1) / t.tile_w;
  t.tiles_y = (out_h + t.tile_h - 1) / t.tile_h;
  return t;
}

template <typename T>
__global__ void __launch_bounds__(FT_THREADS)
fir_planes_tiled(T* __restrict__ out, const T* __restrict__ in, const float* __restrict__ fir,
                 UpfirdnParams p, FirTiling tl, int64_t n_tiles) {
  using A = typename Acc<T>::type;
  extern __shared__ __align__(16) unsigned char fir_smem[];
  A* sx = reinterpret_cast<A*>(fir_smem);
  __shared__ A sk[4][4];
  __shared__ float s_row[4], s_col[4];
  __shared__ int s_sep;
  if (threadIdx.x < 16) {
    int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    // flipped taps, zero-extended on the high side when kh/kw < 4
    sk[ky][kx] = (ky < p.kh && kx < p.kw) ? A(fir[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)]) : A(0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // rank-1 (separable) test, pivot at the largest |tap|
    int py = 0, px = 0;
    float best = 0.f;
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 4; ++x)
        if (fabsf(float(sk[y][x])) > best) { best = fabsf(float(sk[y][x])); py = y; px = x; }
    int sepf = best > 0.f && sizeof(A) == 4;
    for (int y = 0; y < 4 && sepf; ++y)
      for (int x = 0; x < 4; ++x)
        if (fabsf(float(sk[y][x]) - float(sk[y][px]) * (float(sk[py][x]) / float(sk[py][px]))) > 1e-6f * best) { sepf = 0; break; }
    for (int i = 0; i < 4; ++i) {
      s_col[i] = float(sk[i][px]);
      s_row[i] = best > 0.f ? float(sk[py][i]) / float(sk[py][px]) : 0.f;
    }
    s_sep = sepf;
  }
  __syncthreads();
  const bool sep = s_sep != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx = threadIdx.x % tl.threads_x, ty = threadIdx.x / tl.threads_x;
  const bool worker = ty < tl.threads_y;
  const int rows_in = tl.tile_h + 3, cols_in = tl.tile_w + 3;
  const bool vec_store = (p.out_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int64_t r = tile;
    const int tcol = int(r % tl.tiles_x); r /= tl.tiles_x;
    const int trow = int(r % tl.tiles_y); r /= tl.tiles_y;
    const int64_t plane = r;
    const int ox0 = tcol * tl.tile_w, oy0 = trow * tl.tile_h;
    const int ix0 = ox0 - p.pad_x0, iy0 = oy0 - p.pad_y0;
    const T* src = in + plane * int64_t(p.in_h) * p.in_w;
    __syncthreads();  // previous tile fully consumed (also orders the sk writes)
    // stage the input tile: one warp per row, 32 consecutive columns per step, loads issued in
    // batches of 5 before the stores so several are in flight per thread
    for (int ry = warp; ry < rows_in; ry += FT_THREADS / 32) {
      const int iy = iy0 + ry;
      const bool row_ok = iy >= 0 && iy < p.in_h;
      const T* srow = src + int64_t(iy) * p.in_w + ix0;
      A v[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int c = lane + 32 * j;
        const int ix = ix0 + c;
        v[j] = (row_ok && c < cols_in && ix >= 0 && ix < p.in_w) ? to_acc(srow[c]) : A(0);
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int c = lane + 32 * j;
        if (c < cols_in) sx[ry * tl.pitch + c] = v[j];
      }
    }
    __syncthreads();
    if (!worker) continue;
    A acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = A(0);
    if (sizeof(A) == 4 && sep) {
      // separable FIR with packed f32x2 FMAs: horizontal 4-tap reduce of the row (2 FFMA2 chains for the
      // 4 outputs), then vertical scatter into the 4 output rows: 16 FFMA2 per input row instead of 64 FFMA
      float2 acc2[4][2];
#pragma unroll
      for (int a = 0; a < 4; ++a) acc2[a][0] = acc2[a][1] = make_float2(0.f, 0.f);
      const float2 r0 = make_float2(s_row[0], s_row[0]), r1 = make_float2(s_row[1], s_row[1]);
      const float2 r2 = make_float2(s_row[2], s_row[2]), r3 = make_float2(s_row[3], s_row[3]);
#pragma unroll
      for (int ry = 0; ry < 7; ++ry) {
        const float* rp = reinterpret_cast<const float*>(sx) + (ty * 4 + ry) * tl.pitch + tx * 4;
        const float4 q0 = *reinterpret_cast<const float4*>(rp);
        const float4 q1 = *reinterpret_cast<const float4*>(rp + 4);
        // outputs (0,1) use x0..x4, outputs (2,3) use x2..x6
        float2 h01 = __fmul2_rn(make_float2(q0.x, q0.y), r0);
        h01 = __ffma2_rn(make_float2(q0.y, q0.z), r1, h01);
        h01 = __ffma2_rn(make_float2(q0.z, q0.w), r2, h01);
        h01 = __ffma2_rn(make_float2(q0.w, q1.x), r3, h01);
        float2 h23 = __fmul2_rn(make_float2(q0.z, q0.w), r0);
        h23 = __ffma2_rn(make_float2(q0.w, q1.x), r1, h23);
        h23 = __ffma2_rn(make_float2(q1.x, q1.y), r2, h23);
        h23 = __ffma2_rn(make_float2(q1.y, q1.z), r3, h23);
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const int a = ry - ky;
          if (a >= 0 && a < 4) {
            const float2 ck = make_float2(s_col[ky], s_col[ky]);
            acc2[a][0] = __ffma2_rn(h01, ck, acc2[a][0]);
            acc2[a][1] = __ffma2_rn(h23, ck, acc2[a][1]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        acc[a][0] = A(acc2[a][0].x); acc[a][1] = A(acc2[a][0].y);
        acc[a][2] = A(acc2[a][1].x); acc[a][3] = A(acc2[a][1].y);
      }
    } else {
#pragma unroll
    for (int ry = 0; ry < 7; ++ry) {
      const A* rp = sx + (ty * 4 + ry) * tl.pitch + tx * 4;
      A rowv[8];
      if (sizeof(A) == 4) {  // two 16-byte shared loads (pitch and tx*4 are multiples of 4 words)
        const float4 q0 = *reinterpret_cast<const float4*>(rp);
        const float4 q1 = *reinterpret_cast<const float4*>(rp + 4);
        rowv[0] = q0.x; rowv[1] = q0.y; rowv[2] = q0.z; rowv[3] = q0.w;
        rowv[4] = q1.x; rowv[5] = q1.y; rowv[6] = q1.z; rowv[7] = q1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 7; ++j) rowv[j] = rp[j];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int ky = ry - a;  // out row a uses input row a+ky
        if (ky >= 0 && ky < 4) {
#pragma unroll
          for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) acc[a][b] += rowv[b + kx] * sk[ky][kx];
        }
      }
    }
    }
    T* dst = out + plane * int64_t(p.out_h) * p.out_w;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int oy = oy0 + ty * 4 + a;
      const int ox = ox0 + tx * 4;
      if (oy >= p.out_h || ox >= p.out_w) continue;
      T* q = dst + int64_t(oy) * p.out_w + ox;
      if (vec_store && ox + 3 < p.out_w) {
        struct alignas(sizeof(T) * 4) V4 { T v[4]; } o;
#pragma unroll
        for (int b = 0; b < 4; ++b) o.v[b] = from_acc<T, A>(acc[a][b]);
        *reinterpret_cast<V4*>(q) = o;
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (ox + b < p.out_w) q[b] = from_acc<T, A>(acc[a][b]);
      }
    }
  }
}

// ---- f32 planes, separable FIR: cp.async double-buffered version of fir_planes_tiled -------------
// The synchronous kernel alternates "load tile -> barrier -> compute -> store" and relies on other CTAs to
// hide the global-load latency (ncu: long-scoreboard stalls, 0.51 of HBM peak).  Here every CTA prefetches
// tile i+1 into the second shared buffer with cp.async (4-byte copies: plane rows such as 257 floats are not
// 16-byte aligned; src-size 0 zero-fills the padding halo) while it computes tile i.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

__global__ void __launch_bounds__(FT_THREADS)
fir_planes_async(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ fir,
                 UpfirdnParams p, FirTiling tl, int64_t n_tiles, int buf_elems) {
  extern __shared__ __align__(16) unsigned char fir_smem[];
  float* sbuf = reinterpret_cast<float*>(fir_smem);
  __shared__ float sk[4][4];
  __shared__ float s_row[4], s_col[4];
  __shared__ int s_sep;
  if (threadIdx.x < 16) {
    int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    sk[ky][kx] = (ky < p.kh && kx < p.kw) ? fir[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // rank-1 test + factorisation, pivot at the largest |tap|
    int py = 0, px = 0;
    float best = 0.f;
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 4; ++x)
        if (fabsf(sk[y][x]) > best) { best = fabsf(sk[y][x]); py = y; px = x; }
    int sepf = best > 0.f;
    for (int y = 0; y < 4 && sepf; ++y)
      for (int x = 0; x < 4; ++x)
        if (fabsf(sk[y][x] - sk[y][px] * (sk[py][x] / sk[py][px])) > 1e-6f * best) { sepf = 0; break; }
    for (int i = 0; i < 4; ++i) {
      s_col[i] = sk[i][px];
      s_row[i] = best > 0.f ? sk[py][i] / sk[py][px] : 0.f;
    }
    s_sep = sepf;
  }
  __syncthreads();
  const bool sep = s_sep != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx = threadIdx.x % tl.threads_x, ty = threadIdx.x / tl.threads_x;
  const bool worker = ty < tl.threads_y;
  const int rows_in = tl.tile_h + 3, cols_in = tl.tile_w + 3;
  const bool vec_store = (p.out_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);

  auto prefetch = [&](int64_t tile, float* sx) {
    int64_t r = tile;
    const int tcol = int(r % tl.tiles_x); r /= tl.tiles_x;
    const int trow = int(r % tl.tiles_y); r /= tl.tiles_y;
    const int ix0 = tcol * tl.tile_w - p.pad_x0, iy0 = trow * tl.tile_h - p.pad_y0;
    const float* src = in + r * int64_t(p.in_h) * p.in_w;
    for (int ry = warp; ry < rows_in; ry += FT_THREADS / 32) {
      const int iy = iy0 + ry;
      const bool row_ok = iy >= 0 && iy < p.in_h;
      const float* srow = src + int64_t(row_ok ? iy : 0) * p.in_w;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int c = lane + 32 * j;
        if (c < cols_in) {
          const int ix = ix0 + c;
          const bool ok = row_ok && ix >= 0 && ix < p.in_w;
          cp_async4(sx + ry * tl.pitch + c, srow + (ok ? ix : 0), ok);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int64_t tile = blockIdx.x;
  if (tile < n_tiles) prefetch(tile, sbuf);
  int it = 0;
  for (; tile < n_tiles; tile += gridDim.x, ++it) {
    float* sx = sbuf + (it & 1) * buf_elems;
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {
      prefetch(next, sbuf + ((it + 1) & 1) * buf_elems);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    int64_t r = tile;
    const int tcol = int(r % tl.tiles_x); r /= tl.tiles_x;
    const int trow = int(r % tl.tiles_y); r /= tl.tiles_y;
    const int64_t plane = r;
    const int ox0 = tcol * tl.tile_w, oy0 = trow * tl.tile_h;
    if (worker) {
      float2 acc2[4][2];
#pragma unroll
      for (int a = 0; a < 4; ++a) acc2[a][0] = acc2[a][1] = make_float2(0.f, 0.f);
      const float2 r0 = make_float2(s_row[0], s_row[0]), r1 = make_float2(s_row[1], s_row[1]);
      const float2 r2 = make_float2(s_row[2], s_row[2]), r3 = make_float2(s_row[3], s_row[3]);
      if (!sep) {  // general 4x4 FIR: 16 taps per output from the same shared tile
#pragma unroll
        for (int ry = 0; ry < 7; ++ry) {
          const float* rp = sx + (ty * 4 + ry) * tl.pitch + tx * 4;
          float rowv[8];
#pragma unroll
          for (int j = 0; j < 7; ++j) rowv[j] = rp[j];
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int ky = ry - a;
            if (ky >= 0 && ky < 4) {
#pragma unroll
              for (int kx = 0; kx < 4; ++kx) {
                const float t = sk[ky][kx];
                acc2[a][0].x += rowv[kx] * t;     acc2[a][0].y += rowv[1 + kx] * t;
                acc2[a][1].x += rowv[2 + kx] * t; acc2[a][1].y += rowv[3 + kx] * t;
              }
            }
          }
        }
      } else
#pragma unroll
      for (int ry = 0; ry < 7; ++ry) {
        const float* rp = sx + (ty * 4 + ry) * tl.pitch + tx * 4;
        const float4 q0 = *reinterpret_cast<const float4*>(rp);
        const float4 q1 = *reinterpret_cast<const float4*>(rp + 4);
        float2 h01 = __fmul2_rn(make_float2(q0.x, q0.y), r0);
        h01 = __ffma2_rn(make_float2(q0.y, q0.z), r1, h01);
        h01 = __ffma2_rn(make_float2(q0.z, q0.w), r2, h01);
        h01 = __ffma2_rn(make_float2(q0.w, q1.x), r3, h01);
        float2 h23 = __fmul2_rn(make_float2(q0.z, q0.w), r0);
        h23 = __ffma2_rn(make_float2(q0.w, q1.x), r1, h23);
        h23 = __ffma2_rn(make_float2(q1.x, q1.y), r2, h23);
        h23 = __ffma2_rn(make_float2(q1.y, q1.z), r3, h23);
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const int a = ry - ky;
          if (a >= 0 && a < 4) {
            const float2 ck = make_float2(s_col[ky], s_col[ky]);
            acc2[a][0] = __ffma2_rn(h01, ck, acc2[a][0]);
            acc2[a][1] = __ffma2_rn(h23, ck, acc2[a][1]);
          }
        }
      }
      float* dst = out + plane * int64_t(p.out_h) * p.out_w;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int oy = oy0 + ty * 4 + a;
        const int ox = ox0 + tx * 4;
        if (oy >= p.out_h || ox >= p.out_w) continue;
        float* q = dst + int64_t(oy) * p.out_w + ox;
        if (vec_store && ox + 3 < p.out_w) {
          *reinterpret_cast<float4*>(q) = make_float4(acc2[a][0].x, acc2[a][0].y, acc2[a][1].x, acc2[a][1].y);
        } else {
          const float v[4] = {acc2[a][0].x, acc2[a][0].y, acc2[a][1].x, acc2[a][1].y};
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (ox + b < p.out_w) q[b] = v[b];
        }
      }
    }
    __syncthreads();  // buffer (it & 1) may be refilled by the prefetch issued in the next iteration
  }
}

// ---- channels-last hot case: up = down = 1, FIR <= 4x4, minor % VEC == 0 --------------------------
// Thread = one 16-byte channel vector of one output column, walking down FN_TY output rows with a
// sliding window: every input row is loaded once (4 horizontally adjacent vectors) and contributes to
// up to 4 output rows.  When the FIR is an outer product (every FIR on this path is make_kernel([1,3,3,1]),
// model_spatial_query.py:84-92) the row is first reduced horizontally (4 FMA) and then scattered
// vertically (4 FMA) instead of 16 FMA per element.
constexpr int FN_TY = 4;

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
fir_nhwc_kernel(T* __restrict__ out, const T* __restrict__ in, const float* __restrict__ fir, UpfirdnParams p,
                int strips, int64_t total) {
  struct alignas(sizeof(T) * VEC) V { T v[VEC]; };
  __shared__ float sk[4][4];   // flipped, zero-extended taps
  __shared__ float s_row[4], s_col[4];
  __shared__ int s_sep;
  if (threadIdx.x < 16) {
    int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    sk[ky][kx] = (ky < p.kh && kx < p.kw) ? fir[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // rank-1 test: k[y][x] == col[y] * row[x] with the pivot at the largest |tap|
    int py = 0, px = 0;
    float best = 0.f;
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 4; ++x)
        if (fabsf(sk[y][x]) > best) { best = fabsf(sk[y][x]); py = y; px = x; }
    int sep = best > 0.f;
    for (int y = 0; y < 4 && sep; ++y)
      for (int x = 0; x < 4; ++x)
        if (fabsf(sk[y][x] - sk[y][px] * (sk[py][x] / sk[py][px])) > 1e-6f * best) { sep = 0; break; }
    for (int i = 0; i < 4; ++i) {
      s_col[i] = sk[i][px];
      s_row[i] = best > 0.f ? sk[py][i] / sk[py][px] : 0.f;
    }
    s_sep = sep;
  }
  __syncthreads();
  const bool sep = s_sep != 0;
  const int cv = p.minor / VEC;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int64_t r = idx;
    const int c = int(r % cv); r /= cv;
    const int ox = int(r % p.out_w); r /= p.out_w;
    const int strip = int(r % strips); r /= strips;
    const int64_t n = r;
    const int oy0 = strip * FN_TY;
    const int ix0 = ox - p.pad_x0, iy0 = oy0 - p.pad_y0;
    const T* src = in + n * int64_t(p.in_h) * p.in_w * p.minor + int64_t(c) * VEC;
    float acc[FN_TY][VEC];
#pragma unroll
    for (int a = 0; a < FN_TY; ++a)
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[a][j] = 0.f;
#pragma unroll
    for (int ry = 0; ry < FN_TY + 3; ++ry) {
      const int iy = iy0 + ry;
      // predicated (branch-free) loads so the compiler can keep several rows in flight
      const bool row_ok = iy >= 0 && iy < p.in_h && (oy0 + ry - 3 < p.out_h);
      float xv[4][VEC];
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int ix = ix0 + kx;
        V v;
        if (row_ok && ix >= 0 && ix < p.in_w) {
          v = *reinterpret_cast<const V*>(src + (int64_t(iy) * p.in_w + ix) * p.minor);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) v.v[j] = from_acc<T, float>(0.f);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) xv[kx][j] = float(to_acc(v.v[j]));
      }
      if (sep) {
        float h[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          h[j] = xv[0][j] * s_row[0] + xv[1][j] * s_row[1] + xv[2][j] * s_row[2] + xv[3][j] * s_row[3];
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const int a = ry - ky;  // output row a uses input row a + ky
          if (a >= 0 && a < FN_TY) {
            const float t = s_col[ky];
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[a][j] += h[j] * t;
          }
        }
      } else {
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const int a = ry - ky;
          if (a >= 0 && a < FN_TY) {
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
              const float t = sk[ky][kx];
#pragma unroll
              for (int j = 0; j < VEC; ++j) acc[a][j] += xv[kx][j] * t;
            }
          }
        }
      }
    }
    T* dst = out + n * int64_t(p.out_h) * p.out_w * p.minor + int64_t(c) * VEC;
#pragma unroll
    for (int a = 0; a < FN_TY; ++a) {
      const int oy = oy0 + a;
      if (oy >= p.out_h) break;
      V o;
#pragma unroll
      for (int j = 0; j < VEC; ++j) o.v[j] = from_acc<T, float>(acc[a][j]);
      *reinterpret_cast<V*>(dst + (int64_t(oy) * p.out_w + ox) * p.minor) = o;
    }
  }
}

// ---- channels-last 16-bit fast path: packed f32x2 math (FFMA2), 2 output columns x 4 rows per thread ----
// ncu on fir_nhwc_kernel showed the bf16 blur to be ISSUE-bound (57 % issue active at 0.22 of HBM peak):
// this variant halves the instruction count per element with Blackwell's packed f32 FMA (fma.rn.f32x2),
// shares each loaded vector between two adjacent output columns and keeps the separable structure
// (horizontal 4-tap reduce, then vertical scatter).  Non-separable FIRs take a plain 16-tap loop.
template <typename T> struct Unpack2;
template <> struct Unpack2<__nv_bfloat16> {
  static __device__ __forceinline__ float2 get(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
  }
  static __device__ __forceinline__ uint32_t put(float2 f) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f.x, f.y);
    return *reinterpret_cast<uint32_t*>(&t);
  }
};
template <> struct Unpack2<__half> {
  static __device__ __forceinline__ float2 get(uint32_t w) {
    return __half22float2(*reinterpret_cast<__half2*>(&w));
  }
  static __device__ __forceinline__ uint32_t put(float2 f) {
    __half2 t = __floats2half2_rn(f.x, f.y);
    return *reinterpret_cast<uint32_t*>(&t);
  }
};

// One 16-byte vector of T as NQ packed float2 values: the TMA-staged kernels below are written over this view, so
// that f32 tensors (a 128-byte channel chunk = 32 floats, 4 per thread) share the code of the 16-bit ones (64, 8).
template <typename T> struct Pack16 {
  static constexpr int NQ = 4;            // float2 per 16 bytes
  static constexpr int ELEMS = 8;         // elements per 16 bytes
  static constexpr int CHUNK = 64;        // elements per 128-byte channel chunk
  static __device__ __forceinline__ float2 get(const uint4& r, int q) { return Unpack2<T>::get((&r.x)[q]); }
  static __device__ __forceinline__ uint4 put(const float2* a) {
    return make_uint4(Unpack2<T>::put(a[0]), Unpack2<T>::put(a[1]), Unpack2<T>::put(a[2]), Unpack2<T>::put(a[3]));
  }
};
template <> struct Pack16<float> {
  static constexpr int NQ = 2;
  static constexpr int ELEMS = 4;
  static constexpr int CHUNK = 32;
  static __device__ __forceinline__ float2 get(const uint4& r, int q) {
    return make_float2(__uint_as_float((&r.x)[2 * q]), __uint_as_float((&r.x)[2 * q + 1]));
  }
  static __device__ __forceinline__ uint4 put(const float2* a) {
    return make_uint4(__float_as_uint(a[0].x), __float_as_uint(a[0].y), __float_as_uint(a[1].x), __float_as_uint(a[1].y));
  }
};
template <typename T> struct IsTmaFir { static constexpr bool value = false; };
template <> struct IsTmaFir<__nv_bfloat16> { static constexpr bool value = true; };
template <> struct IsTmaFir<__half> { static constexpr bool value = true; };
template <> struct IsTmaFir<float> { static constexpr bool value = true; };

constexpr int F2_TY = 4;   // output rows per thread
constexpr int F2_XW = 2;   // output columns per thread

template <typename T>
__global__ void __launch_bounds__(128)
fir_nhwc16_kernel(T* __restrict__ out, const T* __restrict__ in, const float* __restrict__ fir, UpfirdnParams p,
                  int strips, int xpairs, int64_t total) {
  __shared__ float sk[4][4];
  __shared__ float s_row[4], s_col[4];
  __shared__ int s_sep;
  if (threadIdx.x < 16) {
    int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    sk[ky][kx] = (ky < p.kh && kx < p.kw) ? fir[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int py = 0, px = 0;
    float best = 0.f;
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 4; ++x)
        if (fabsf(sk[y][x]) > best) { best = fabsf(sk[y][x]); py = y; px = x; }
    int sep = best > 0.f;
    for (int y = 0; y < 4 && sep; ++y)
      for (int x = 0; x < 4; ++x)
        if (fabsf(sk[y][x] - sk[y][px] * (sk[py][x] / sk[py][px])) > 1e-6f * best) { sep = 0; break; }
    for (int i = 0; i < 4; ++i) {
      s_col[i] = sk[i][px];
      s_row[i] = best > 0.f ? sk[py][i] / sk[py][px] : 0.f;
    }
    s_sep = sep;
  }
  __syncthreads();
  const bool sep = s_sep != 0;
  const int cv = p.minor / 8;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int64_t r = idx;
    const int c = int(r % cv); r /= cv;
    const int xp = int(r % xpairs); r /= xpairs;
    const int strip = int(r % strips); r /= strips;
    const int64_t n = r;
    const int ox0 = xp * F2_XW, oy0 = strip * F2_TY;
    const int ix0 = ox0 - p.pad_x0, iy0 = oy0 - p.pad_y0;
    const T* src = in + n * int64_t(p.in_h) * p.in_w * p.minor + int64_t(c) * 8;
    T* dst = out + n * int64_t(p.out_h) * p.out_w * p.minor + int64_t(c) * 8;
    if (sep) {
      const float2 r0 = make_float2(s_row[0], s_row[0]), r1 = make_float2(s_row[1], s_row[1]);
      const float2 r2 = make_float2(s_row[2], s_row[2]), r3 = make_float2(s_row[3], s_row[3]);
      float2 acc[F2_TY][F2_XW][4];
#pragma unroll
      for (int a = 0; a < F2_TY; ++a)
#pragma unroll
        for (int w = 0; w < F2_XW; ++w)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[a][w][q] = make_float2(0.f, 0.f);
#pragma unroll
      for (int ry = 0; ry < F2_TY + 3; ++ry) {
        const int iy = iy0 + ry;
        const bool row_ok = iy >= 0 && iy < p.in_h && (oy0 + ry - 3 < p.out_h);
        uint4 raw[5];
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          const int ix = ix0 + kx;
          raw[kx] = (row_ok && ix >= 0 && ix < p.in_w)
                        ? *reinterpret_cast<const uint4*>(src + (int64_t(iy) * p.in_w + ix) * p.minor)
                        : make_uint4(0u, 0u, 0u, 0u);
        }
        float2 h[F2_XW][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 v[5];
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) v[kx] = Unpack2<T>::get((&raw[kx].x)[q]);
#pragma unroll
          for (int w = 0; w < F2_XW; ++w) {
            float2 t = __fmul2_rn(v[w], r0);
            t = __ffma2_rn(v[w + 1], r1, t);
            t = __ffma2_rn(v[w + 2], r2, t);
            h[w][q] = __ffma2_rn(v[w + 3], r3, t);
          }
        }
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const int a = ry - ky;  // output row a uses input row a + ky
          if (a >= 0 && a < F2_TY) {
            const float2 ck = make_float2(s_col[ky], s_col[ky]);
#pragma unroll
            for (int w = 0; w < F2_XW; ++w)
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[a][w][q] = __ffma2_rn(h[w][q], ck, acc[a][w][q]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < F2_TY; ++a) {
        const int oy = oy0 + a;
        if (oy >= p.out_h) break;
#pragma unroll
        for (int w = 0; w < F2_XW; ++w) {
          const int ox = ox0 + w;
          if (ox >= p.out_w) continue;
          uint4 o;
          o.x = Unpack2<T>::put(acc[a][w][0]);
          o.y = Unpack2<T>::put(acc[a][w][1]);
          o.z = Unpack2<T>::put(acc[a][w][2]);
          o.w = Unpack2<T>::put(acc[a][w][3]);
          *reinterpret_cast<uint4*>(dst + (int64_t(oy) * p.out_w + ox) * p.minor) = o;
        }
      }
    } else {
      // general FIR: plain 16-tap accumulation per output (rare: every FIR on the path is separable)
      for (int a = 0; a < F2_TY; ++a) {
        const int oy = oy0 + a;
        if (oy >= p.out_h) break;
        for (int w = 0; w < F2_XW; ++w) {
          const int ox = ox0 + w;
          if (ox >= p.out_w) continue;
          float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
          for (int ky = 0; ky < 4; ++ky) {
            const int iy = oy - p.pad_y0 + ky;
            if (iy < 0 || iy >= p.in_h) continue;
            for (int kx = 0; kx < 4; ++kx) {
              const int ix = ox - p.pad_x0 + kx;
              if (ix < 0 || ix >= p.in_w) continue;
              const uint4 raw = *reinterpret_cast<const uint4*>(src + (int64_t(iy) * p.in_w + ix) * p.minor);
              const float2 t = make_float2(sk[ky][kx], sk[ky][kx]);
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[q] = __ffma2_rn(Unpack2<T>::get((&raw.x)[q]), t, acc[q]);
            }
          }
          uint4 o;
          o.x = Unpack2<T>::put(acc[0]); o.y = Unpack2<T>::put(acc[1]);
          o.z = Unpack2<T>::put(acc[2]); o.w = Unpack2<T>::put(acc[3]);
          *reinterpret_cast<uint4*>(dst + (int64_t(oy) * p.out_w + ox) * p.minor) = o;
        }
      }
    }
  }
}

template <typename T> struct Is16 { static constexpr bool value = false; };
template <> struct Is16<__nv_bfloat16> { static constexpr bool value = true; };
template <> struct Is16<__half> { static constexpr bool value = true; };

template <typename T>
static int launch_fir_nhwc16(T* out, const T* in, const float* fir, const UpfirdnParams& p, cudaStream_t st) {
  if constexpr (Is16<T>::value) {
    const int strips = (p.out_h + F2_TY - 1) / F2_TY;
    const int xpairs = (p.out_w + F2_XW - 1) / F2_XW;
    const int64_t work = p.major * strips * int64_t(xpairs) * (p.minor / 8);
    fir_nhwc16_kernel<T><<<grid_for(work, 128, 24), 128, 0, st>>>(out, in, fir, p, strips, xpairs, work);
    return 1;
  } else {
    return 0;
  }
}


// ---- channels-last, 16-bit, TMA-staged ---------------------------------------------------------------------
// The register kernel above tops out near half of HBM bandwidth: 128 registers per thread leave 16 warps per SM and
// every warp alternates between its loads and ~700 instructions of FIR arithmetic, so too few bytes are in flight.
// Here one elected thread streams haloed [rows+3][cols+3][64 channels] boxes into a two-stage shared-memory ring
// with cp.async.bulk.tensor (out-of-bounds = the operator's zero padding, negative start coordinates included) while
// all threads filter the previous box out of shared memory; two such CTAs per SM keep ~200 KB of loads in flight.
struct FirTmaTiling {
  int tw, th;          // output tile (columns even, rows multiple of F2_TY)
  int xpairs, strips;  // per tile: tw / 2, th / F2_TY
  int box_w, box_h;    // tw + 3, th + 3
  int tiles_x, tiles_y, chunks;
  int64_t jobs;
  // optional epilogue (te_upfirdn2d_bias_act): out = leaky_relu(fir(x) + bias[c], slope) * gain
  const float* bias;
  float slope, gain;
  int act;
};
constexpr int FIR_TMA_STAGES = 2;
constexpr int FIR_TMA_MAX_THREADS = 256;  // x 2 CTAs per SM -> 128 registers per thread, no spills

template <typename T>
__global__ void __launch_bounds__(FIR_TMA_MAX_THREADS, 2)
fir_nhwc_tma_kernel(T* __restrict__ out, const __grid_constant__ CUtensorMap in_map, const float* __restrict__ fir,
                    UpfirdnParams p, FirTmaTiling t) {
  extern __shared__ __align__(128) uint8_t fir_smem_raw[];
  __shared__ __align__(8) uint64_t full[FIR_TMA_STAGES];
  __shared__ float sk[4][4];
  __shared__ float s_row[4], s_col[4];
  __shared__ int s_sep;
  uint8_t* stage0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fir_smem_raw) + 127) & ~uintptr_t(127));
  const uint32_t stage_bytes = uint32_t(t.box_h) * t.box_w * 128u;
  const int tid = threadIdx.x;
  if (tid < 16) {
    int ky = tid >> 2, kx = tid & 3;
    sk[ky][kx] = (ky < p.kh && kx < p.kw) ? fir[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.f;
  }
  if (tid == 32) {
    for (int s = 0; s < FIR_TMA_STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int64_t job, int s) {
    const int chunk = int(job % t.chunks);
    int64_t r = job / t.chunks;
    const int tx = int(r % t.tiles_x); r /= t.tiles_x;
    const int ty = int(r % t.tiles_y); r /= t.tiles_y;
    mbar_expect_tx(&full[s], stage_bytes);
    tma_load_4d(stage0 + size_t(s) * stage_bytes, &in_map, &full[s], chunk * Pack16<T>::CHUNK, tx * t.tw - p.pad_x0,
                ty * t.th - p.pad_y0, int(r));
  };
  if (tid == 32) {
    int64_t job = blockIdx.x;
    for (int s = 0; s < FIR_TMA_STAGES && job < t.jobs; ++s, job += gridDim.x) issue(job, s);
  }
  if (tid == 0) {
    int py = 0, px = 0;
    float best = 0.f;
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 4; ++x)
        if (fabsf(sk[y][x]) > best) { best = fabsf(sk[y][x]); py = y; px = x; }
    int sep = best > 0.f;
    for (int y = 0; y < 4 && sep; ++y)
      for (int x = 0; x < 4; ++x)
        if (fabsf(sk[y][x] - sk[y][px] * (sk[py][x] / sk[py][px])) > 1e-6f * best) { sep = 0; break; }
    for (int i = 0; i < 4; ++i) {
      s_col[i] = sk[i][px];
      s_row[i] = best > 0.f ? sk[py][i] / sk[py][px] : 0.f;
    }
    s_sep = sep;
  }
  __syncthreads();
  const bool sep = s_sep != 0;
  const int q8 = tid & 7;                 // 8-channel group inside the 64-channel chunk
  const int xp = (tid >> 3) % t.xpairs;
  const int strip = (tid >> 3) / t.xpairs;
  const bool active = strip < t.strips;
  const float2 r0 = make_float2(s_row[0], s_row[0]), r1 = make_float2(s_row[1], s_row[1]);
  const float2 r2 = make_float2(s_row[2], s_row[2]), r3 = make_float2(s_row[3], s_row[3]);
  const float c0 = s_col[0], c1 = s_col[1], c2 = s_col[2], c3 = s_col[3];

  int it = 0;
  for (int64_t job = blockIdx.x; job < t.jobs; job += gridDim.x, ++it) {
    const int s = it % FIR_TMA_STAGES;
    mbar_wait(&full[s], (it / FIR_TMA_STAGES) & 1);
    if (active) {
      const int chunk = int(job % t.chunks);
      int64_t r = job / t.chunks;
      const int tx = int(r % t.tiles_x); r /= t.tiles_x;
      const int ty = int(r % t.tiles_y); r /= t.tiles_y;
      const int ox0 = tx * t.tw + xp * F2_XW, oy0 = ty * t.th + strip * F2_TY;
      const uint8_t* src = stage0 + size_t(s) * stage_bytes +
                           (size_t(strip * F2_TY) * t.box_w + xp * F2_XW) * 128 + q8 * 16;
      T* dst = out + ((r * p.out_h + oy0) * int64_t(p.out_w) + ox0) * p.minor + chunk * Pack16<T>::CHUNK + q8 * Pack16<T>::ELEMS;
      if (ox0 < p.out_w && oy0 < p.out_h) {
        float2 acc[F2_TY][F2_XW][Pack16<T>::NQ];
#pragma unroll
        for (int a = 0; a < F2_TY; ++a)
#pragma unroll
          for (int w = 0; w < F2_XW; ++w)
#pragma unroll
            for (int q = 0; q < Pack16<T>::NQ; ++q) acc[a][w][q] = make_float2(0.f, 0.f);
        if (sep) {
#pragma unroll
          for (int ry = 0; ry < F2_TY + 3; ++ry) {
            uint4 raw[5];
#pragma unroll
            for (int kx = 0; kx < 5; ++kx)
              raw[kx] = *reinterpret_cast<const uint4*>(src + (size_t(ry) * t.box_w + kx) * 128);
            float2 h[F2_XW][Pack16<T>::NQ];
#pragma unroll
            for (int q = 0; q < Pack16<T>::NQ; ++q) {
              float2 v[5];
#pragma unroll
              for (int kx = 0; kx < 5; ++kx) v[kx] = Pack16<T>::get(raw[kx], q);
#pragma unroll
              for (int w = 0; w < F2_XW; ++w) {
                float2 u = __fmul2_rn(v[w], r0);
                u = __ffma2_rn(v[w + 1], r1, u);
                u = __ffma2_rn(v[w + 2], r2, u);
                h[w][q] = __ffma2_rn(v[w + 3], r3, u);
              }
            }
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
              const int a = ry - ky;
              if (a >= 0 && a < F2_TY) {
                const float cs = ky == 0 ? c0 : ky == 1 ? c1 : ky == 2 ? c2 : c3;
                const float2 ck = make_float2(cs, cs);
#pragma unroll
                for (int w = 0; w < F2_XW; ++w)
#pragma unroll
                  for (int q = 0; q < Pack16<T>::NQ; ++q) acc[a][w][q] = __ffma2_rn(h[w][q], ck, acc[a][w][q]);
              }
            }
          }
        } else {
          // general (non rank-1) FIR: rare, every FIR on the path is separable
          for (int ry = 0; ry < F2_TY + 3; ++ry)
            for (int kx = 0; kx < 5; ++kx) {
              const uint4 raw = *reinterpret_cast<const uint4*>(src + (size_t(ry) * t.box_w + kx) * 128);
#pragma unroll
              for (int a = 0; a < F2_TY; ++a)
#pragma unroll
                for (int w = 0; w < F2_XW; ++w) {
                  const int ky = ry - a, kk = kx - w;
                  if (ky >= 0 && ky < 4 && kk >= 0 && kk < 4) {
                    const float2 tp = make_float2(sk[ky][kk], sk[ky][kk]);
#pragma unroll
                    for (int q = 0; q < Pack16<T>::NQ; ++q)
                      acc[a][w][q] = __ffma2_rn(Pack16<T>::get(raw, q), tp, acc[a][w][q]);
                  }
                }
            }
        }
        if (t.act) {
          // fused bias + leaky ReLU of the layer that follows the blur (StyledConv's upsampling branch)
          const float* bp = t.bias + chunk * Pack16<T>::CHUNK + q8 * Pack16<T>::ELEMS;
          float2 bq[Pack16<T>::NQ];
#pragma unroll
          for (int q = 0; q < Pack16<T>::NQ; ++q) bq[q] = make_float2(__ldg(bp + 2 * q), __ldg(bp + 2 * q + 1));
#pragma unroll
          for (int a = 0; a < F2_TY; ++a)
#pragma unroll
            for (int w = 0; w < F2_XW; ++w)
#pragma unroll
              for (int q = 0; q < Pack16<T>::NQ; ++q) {
                const float u0 = acc[a][w][q].x + bq[q].x, u1 = acc[a][w][q].y + bq[q].y;
                acc[a][w][q] = make_float2(fmaxf(u0, t.slope * u0) * t.gain, fmaxf(u1, t.slope * u1) * t.gain);
              }
        }
#pragma unroll
        for (int a = 0; a < F2_TY; ++a) {
          if (oy0 + a >= p.out_h) break;
#pragma unroll
          for (int w = 0; w < F2_XW; ++w) {
            if (ox0 + w >= p.out_w) continue;
            *reinterpret_cast<uint4*>(dst + (int64_t(a) * p.out_w + w) * p.minor) = Pack16<T>::put(acc[a][w]);
          }
        }
      }
    }
    __syncthreads();  // every thread is done reading stage s
    const int64_t next = job + int64_t(FIR_TMA_STAGES) * gridDim.x;
    if (tid == 32 && next < t.jobs) issue(next, s);
  }
}

template <typename T>
static int launch_fir_nhwc_tma(T* out, const T* in, const float* fir, const UpfirdnParams& p, cudaStream_t st,
                               int* status, const float* bias = nullptr, float slope = 1.f, float gain = 1.f,
                               int act = 0) {
  *status = TE_OK;
  if constexpr (IsTmaFir<T>::value) {
    FirTmaTiling t;
    t.bias = bias; t.slope = slope; t.gain = gain; t.act = act;
    const int nx = (p.out_w + 31) / 32;
    t.tw = (((p.out_w + nx - 1) / nx) + 1) & ~1;
    t.xpairs = t.tw / 2;
    int strips = FIR_TMA_MAX_THREADS / (t.xpairs * 8);
    const int need = (p.out_h + F2_TY - 1) / F2_TY;
    strips = strips < 1 ? 1 : strips > 4 ? 4 : strips;
    if (strips > need) strips = need;
    t.strips = strips;
    t.th = strips * F2_TY;
    t.box_w = t.tw + 3;
    t.box_h = t.th + 3;
    t.tiles_x = (p.out_w + t.tw - 1) / t.tw;
    t.tiles_y = (p.out_h + t.th - 1) / t.th;
    t.chunks = p.minor / Pack16<T>::CHUNK;
    t.jobs = int64_t(p.major) * t.tiles_y * t.tiles_x * t.chunks;
    const int threads = ((t.strips * t.xpairs * 8) + 31) / 32 * 32;
    const size_t smem = size_t(FIR_TMA_STAGES) * t.box_h * t.box_w * 128 + 128;
    if (threads < 64 || threads > FIR_TMA_MAX_THREADS || smem > 110 * 1024) return 0;
    CUtensorMap map;
    const uint64_t dims[4] = {uint64_t(p.minor), uint64_t(p.in_w), uint64_t(p.in_h), uint64_t(p.major)};
    const uint64_t strides[3] = {uint64_t(p.minor) * sizeof(T), uint64_t(p.in_w) * p.minor * sizeof(T),
                                 uint64_t(p.in_h) * p.in_w * p.minor * sizeof(T)};
    const uint32_t box[4] = {Pack16<T>::CHUNK, uint32_t(t.box_w), uint32_t(t.box_h), 1};
    *status = encode_map_linear(&map, in, 4, dims, strides, box, int(sizeof(T)));
    if (*status != TE_OK) return 1;
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(fir_nhwc_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
      attr_done = true;
    }
    const int64_t cap = int64_t(kNumSMs) * 2;
    const int grid = int(t.jobs < cap ? t.jobs : cap);
    fir_nhwc_tma_kernel<T><<<grid, threads, smem, st>>>(out, map, fir, p, t);
    return 1;
  } else {
    return 0;
  }
}


// ---- channels-last, 16-bit, TMA-staged, factor-2 resampling ----------------------------------------------------
// down = 2 (blur + decimate: the ResBlock skip branch evaluated only where the stride-2 1x1 convolution reads it,
// model_spatial_query.py:744-750,788) and up = 2 (its gradient, and the RGB-skip Upsample geometry).  Same
// producer/consumer ring as fir_nhwc_tma_kernel.  Every output is a direct sum over the taps that touch it (16 for
// down = 2, 4 for up = 2 — neighbouring outputs share no horizontal partial sums at stride 2, so a separable
// evaluation would not save anything).
struct FirResampleTiling {
  int tw, th;                    // output tile
  int box_w, box_h;              // input box (pixels)
  int groups_x, groups_y;        // thread groups per tile: (tw, th/2) for down=2, (tw/2, th/4) for up=2
  int tiles_x, tiles_y, chunks;
  int org_x, org_y;              // input coordinate of the box of tile (0, 0)
  int64_t jobs;
};

template <typename T, int MODE>  // MODE 0: down=2; 1: up=2 with even pad0; 2: up=2 with odd pad0
__global__ void __launch_bounds__(256, 2)
fir_nhwc_resample_tma_kernel(T* __restrict__ out, const __grid_constant__ CUtensorMap in_map,
                             const float* __restrict__ fir, UpfirdnParams p, FirResampleTiling t) {
  extern __shared__ __align__(128) uint8_t fir_smem_raw[];
  __shared__ __align__(8) uint64_t full[FIR_TMA_STAGES];
  __shared__ float sk[4][4];
  uint8_t* stage0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fir_smem_raw) + 127) & ~uintptr_t(127));
  const uint32_t stage_bytes = uint32_t(t.box_h) * t.box_w * 128u;
  const int tid = threadIdx.x;
  if (tid < 16) {
    int ky = tid >> 2, kx = tid & 3;
    sk[ky][kx] = (ky < p.kh && kx < p.kw) ? fir[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.f;
  }
  if (tid == 32) {
    for (int s = 0; s < FIR_TMA_STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr int IN_PER_OUT_NUM = MODE == 0 ? 2 : 1, IN_PER_OUT_DEN = MODE == 0 ? 1 : 2;
  auto issue = [&](int64_t job, int s) {
    const int chunk = int(job % t.chunks);
    int64_t r = job / t.chunks;
    const int tx = int(r % t.tiles_x); r /= t.tiles_x;
    const int ty = int(r % t.tiles_y); r /= t.tiles_y;
    mbar_expect_tx(&full[s], stage_bytes);
    tma_load_4d(stage0 + size_t(s) * stage_bytes, &in_map, &full[s], chunk * Pack16<T>::CHUNK,
                t.org_x + tx * t.tw * IN_PER_OUT_NUM / IN_PER_OUT_DEN,
                t.org_y + ty * t.th * IN_PER_OUT_NUM / IN_PER_OUT_DEN, int(r));
  };
  if (tid == 32) {
    int64_t job = blockIdx.x;
    for (int s = 0; s < FIR_TMA_STAGES && job < t.jobs; ++s, job += gridDim.x) issue(job, s);
  }
  float2 wt[4][4];
#pragma unroll
  for (int ky = 0; ky < 4; ++ky)
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) wt[ky][kx] = make_float2(sk[ky][kx], sk[ky][kx]);
  const int q8 = tid & 7;
  const int gx = (tid >> 3) % t.groups_x;
  const int gy = (tid >> 3) / t.groups_x;
  const bool active = gy < t.groups_y;

  int it = 0;
  for (int64_t job = blockIdx.x; job < t.jobs; job += gridDim.x, ++it) {
    const int s = it % FIR_TMA_STAGES;
    mbar_wait(&full[s], (it / FIR_TMA_STAGES) & 1);
    if (active) {
      const int chunk = int(job % t.chunks);
      int64_t r = job / t.chunks;
      const int tx = int(r % t.tiles_x); r /= t.tiles_x;
      const int ty = int(r % t.tiles_y); r /= t.tiles_y;
      const uint8_t* tile = stage0 + size_t(s) * stage_bytes + q8 * 16;
      if (MODE == 0) {
        // thread: output column gx, output rows 2*gy, 2*gy+1 of the tile; window rows 4*gy .. 4*gy+5, cols 2*gx .. +3
        const int ox = tx * t.tw + gx, oy0 = ty * t.th + 2 * gy;
        if (ox < p.out_w && oy0 < p.out_h) {
          float2 acc[2][Pack16<T>::NQ];
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int q = 0; q < Pack16<T>::NQ; ++q) acc[a][q] = make_float2(0.f, 0.f);
#pragma unroll
          for (int rr = 0; rr < 6; ++rr) {
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
              const uint4 raw = *reinterpret_cast<const uint4*>(tile + (size_t(4 * gy + rr) * t.box_w + 2 * gx + kx) * 128);
#pragma unroll
              for (int a = 0; a < 2; ++a) {
                const int ky = rr - 2 * a;
                if (ky >= 0 && ky < 4) {
#pragma unroll
                  for (int q = 0; q < Pack16<T>::NQ; ++q)
                    acc[a][q] = __ffma2_rn(Pack16<T>::get(raw, q), wt[ky][kx], acc[a][q]);
                }
              }
            }
          }
          T* dst = out + ((r * p.out_h + oy0) * int64_t(p.out_w) + ox) * p.minor + chunk * Pack16<T>::CHUNK + q8 * Pack16<T>::ELEMS;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            if (oy0 + a >= p.out_h) break;
            *reinterpret_cast<uint4*>(dst + int64_t(a) * p.out_w * p.minor) = Pack16<T>::put(acc[a]);
          }
        }
      } else {
        // up = 2, pad0 = 2m + E.  Output o = 2i + par reads compact inputs i - m + j + (E == 0 ? par : 0), j = 0, 1,
        // through taps (E ^ par) + 2j.  Thread: output columns 2*gx, 2*gx+1 and rows 4*gy .. 4*gy+3 of the tile;
        // the box starts at compact (i0 - m): rows 2*gy .. 2*gy+3, columns gx .. gx+2.
        constexpr int E = MODE == 1 ? 0 : 1;
        const int ox0 = tx * t.tw + 2 * gx, oy0 = ty * t.th + 4 * gy;
        if (ox0 < p.out_w && oy0 < p.out_h) {
          float2 acc[4][2][Pack16<T>::NQ];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
              for (int q = 0; q < Pack16<T>::NQ; ++q) acc[a][w][q] = make_float2(0.f, 0.f);
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
              if (E == 1 && cc == 2) continue;
              const uint4 raw = *reinterpret_cast<const uint4*>(tile + (size_t(2 * gy + rr) * t.box_w + gx + cc) * 128);
              float2 v[Pack16<T>::NQ];
#pragma unroll
              for (int q = 0; q < Pack16<T>::NQ; ++q) v[q] = Pack16<T>::get(raw, q);
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const int pa = a & 1, h = a >> 1;
                const int jy = rr - h - (E == 0 ? pa : 0);
                if (jy < 0 || jy > 1) continue;
                const int ky = (E ^ pa) + 2 * jy;
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                  const int jx = cc - (E == 0 ? w : 0);
                  if (jx < 0 || jx > 1) continue;
                  const int kx = (E ^ w) + 2 * jx;
#pragma unroll
                  for (int q = 0; q < Pack16<T>::NQ; ++q) acc[a][w][q] = __ffma2_rn(v[q], wt[ky][kx], acc[a][w][q]);
                }
              }
            }
          }
          T* dst = out + ((r * p.out_h + oy0) * int64_t(p.out_w) + ox0) * p.minor + chunk * Pack16<T>::CHUNK + q8 * Pack16<T>::ELEMS;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            if (oy0 + a >= p.out_h) break;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              if (ox0 + w >= p.out_w) continue;
              *reinterpret_cast<uint4*>(dst + (int64_t(a) * p.out_w + w) * p.minor) = Pack16<T>::put(acc[a][w]);
            }
          }
        }
      }
    }
    __syncthreads();
    const int64_t next = job + int64_t(FIR_TMA_STAGES) * gridDim.x;
    if (tid == 32 && next < t.jobs) issue(next, s);
  }
}

template <typename T, int MODE>
static int launch_resample_mode(T* out, const CUtensorMap& map, const float* fir, const UpfirdnParams& p,
                                const FirResampleTiling& t, int threads, size_t smem, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(fir_nhwc_resample_tma_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    attr_done = true;
  }
  const int64_t cap = int64_t(kNumSMs) * 2;
  const int grid = int(t.jobs < cap ? t.jobs : cap);
  fir_nhwc_resample_tma_kernel<T, MODE><<<grid, threads, smem, st>>>(out, map, fir, p, t);
  return 1;
}

// returns 1 when a kernel was launched (or *status carries an error), 0 when the geometry is not covered
template <typename T>
static int launch_fir_nhwc_resample(T* out, const T* in, const float* fir, const UpfirdnParams& p, cudaStream_t st,
                                    int* status) {
  *status = TE_OK;
  if constexpr (IsTmaFir<T>::value) {
    const bool down2 = p.up_x == 1 && p.up_y == 1 && p.down_x == 2 && p.down_y == 2;
    const bool up2 = p.up_x == 2 && p.up_y == 2 && p.down_x == 1 && p.down_y == 1;
    if (!(down2 || up2) || p.pad_x0 != p.pad_y0 || p.kh > 4 || p.kw > 4 || p.minor % Pack16<T>::CHUNK != 0) return 0;
    FirResampleTiling t;
    int mode;
    if (down2) {
      mode = 0;
      t.tw = p.out_w >= 16 ? 16 : p.out_w;
      t.th = 4;
      t.box_w = 2 * t.tw + 2; t.box_h = 2 * t.th + 2;
      t.groups_x = t.tw; t.groups_y = t.th / 2;
      t.org_x = -p.pad_x0; t.org_y = -p.pad_y0;
    } else {
      const int e = p.pad_x0 & 1, m = (p.pad_x0 - e) / 2;  // pad0 = 2m + e, floor semantics for negative pads
      mode = e == 0 ? 1 : 2;
      t.tw = p.out_w >= 32 ? 32 : (p.out_w + 1) & ~1;
      t.th = 8;
      t.box_w = t.tw / 2 + 2; t.box_h = t.th / 2 + 2;
      t.groups_x = t.tw / 2; t.groups_y = t.th / 4;
      t.org_x = -m; t.org_y = -m;
    }
    const int threads = (t.groups_x * t.groups_y * 8 + 31) / 32 * 32;
    if (threads > 256 || t.tw < 2) return 0;
    t.tiles_x = (p.out_w + t.tw - 1) / t.tw;
    t.tiles_y = (p.out_h + t.th - 1) / t.th;
    t.chunks = p.minor / Pack16<T>::CHUNK;
    t.jobs = int64_t(p.major) * t.tiles_y * t.tiles_x * t.chunks;
    const size_t smem = size_t(FIR_TMA_STAGES) * t.box_h * t.box_w * 128 + 128;
    CUtensorMap map;
    const uint64_t dims[4] = {uint64_t(p.minor), uint64_t(p.in_w), uint64_t(p.in_h), uint64_t(p.major)};
    const uint64_t strides[3] = {uint64_t(p.minor) * sizeof(T), uint64_t(p.in_w) * p.minor * sizeof(T),
                                 uint64_t(p.in_h) * p.in_w * p.minor * sizeof(T)};
    const uint32_t box[4] = {Pack16<T>::CHUNK, uint32_t(t.box_w), uint32_t(t.box_h), 1};
    *status = encode_map_linear(&map, in, 4, dims, strides, box, int(sizeof(T)));
    if (*status != TE_OK) return 1;
    if (mode == 0) return launch_resample_mode<T, 0>(out, map, fir, p, t, threads, smem, st);
    if (mode == 1) return launch_resample_mode<T, 1>(out, map, fir, p, t, threads, smem, st);
    return launch_resample_mode<T, 2>(out, map, fir, p, t, threads, smem, st);
  } else {
    return 0;
  }
}

template <typename T>
static int upfirdn2d_typed(void* out_, const void* in_, const float* fir, const UpfirdnParams& p,
                           cudaStream_t st) {
  T* out = static_cast<T*>(out_);
  const T* in = static_cast<const T*>(in_);
  const int64_t total = p.major * p.out_h * int64_t(p.out_w) * p.minor;
  if (total == 0) return TE_OK;
  const bool hot = p.up_x == 1 && p.up_y == 1 && p.down_x == 1 && p.down_y == 1 && p.minor == 1 &&
                   p.kh <= 4 && p.kw <= 4 && p.out_w >= 16 && p.out_h >= 8 && sizeof(T) <= 4;
  constexpr int NVEC = 16 / sizeof(T);
  const bool hot_cl = p.up_x == 1 && p.up_y == 1 && p.down_x == 1 && p.down_y == 1 && p.minor > 1 &&
                      p.minor % NVEC == 0 && p.kh <= 4 && p.kw <= 4 && sizeof(T) <= 4 &&
                      (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  static const bool use_tma = getenv("TE_FIR_TMA") == nullptr || atoi(getenv("TE_FIR_TMA")) != 0;
  int tma_status = TE_OK;
  const bool aligned16 = (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  if (use_tma && aligned16 && p.minor > 1 && total >= (int64_t(1) << 16) &&
      launch_fir_nhwc_resample<T>(out, in, fir, p, st, &tma_status)) {
    // 16-bit channels-last factor-2 resampling: TMA-staged kernel launched
    if (tma_status != TE_OK) return tma_status;
  } else if (hot_cl && use_tma && p.minor % (128 / int(sizeof(T))) == 0 && total >= (int64_t(1) << 20) &&
      launch_fir_nhwc_tma<T>(out, in, fir, p, st, &tma_status)) {
    // 16-bit channels-last, large: TMA-staged kernel launched (or the tensor map could not be encoded)
    if (tma_status != TE_OK) return tma_status;
  } else if (hot_cl && launch_fir_nhwc16<T>(out, in, fir, p, st)) {
    // 16-bit channels-last: packed-f32 register kernel launched
  } else if (hot_cl) {
    const int strips = (p.out_h + FN_TY - 1) / FN_TY;
    const int64_t work = p.major * strips * int64_t(p.out_w) * (p.minor / NVEC);
    fir_nhwc_kernel<T, (sizeof(T) <= 4 ? NVEC : 1)><<<grid_for(work, 256, 8), 256, 0, st>>>(out, in, fir, p, strips, work);
  } else if (hot) {
    const FirTiling tl = fir_tiling(p.out_w, p.out_h);
    const int64_t n_tiles = p.major * tl.tiles_x * tl.tiles_y;
    const size_t smem = size_t(tl.tile_h + 3) * tl.pitch * sizeof(typename Acc<T>::type);
    if constexpr (sizeof(T) == 4) {
      // f32: cp.async double-buffered kernel; ~4 resident CTAs per SM, a few tiles each
      const int buf_elems = (tl.tile_h + 3) * tl.pitch;
      int64_t blocks = n_tiles < int64_t(kNumSMs) * 4 ? n_tiles : int64_t(kNumSMs) * 4;
      fir_planes_async<<<unsigned(blocks), FT_THREADS, 2 * smem, st>>>(
          reinterpret_cast<float*>(out), reinterpret_cast<const float*>(in), fir, p, tl, n_tiles, buf_elems);
    } else {
      int64_t blocks = n_tiles < int64_t(kNumSMs) * 8 ? n_tiles : int64_t(kNumSMs) * 8;
      fir_planes_tiled<T><<<unsigned(blocks), FT_THREADS, smem, st>>>(out, in, fir, p, tl, n_tiles);
    }
  } else {
    upfirdn2d_generic<T><<<grid_for(total, 256, 32), 256, 0, st>>>(out, in, fir, p, total);
  }
  TE_CHECK_LAUNCH();
  return TE_OK;
}

}  // namespace te

extern "C" int te_upfirdn2d(void* out, const void* in, const float* fir, int64_t major, int in_h,
                            int in_w, int minor, int kh, int kw, int up_x, int up_y, int down_x,
                            int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1, int dtype,
                            void* stream) {
  using namespace te;
  if (major == 0) return TE_OK;  // empty batch: nothing to do (empty tensors carry null data pointers)
  TE_CHECK_ARG(out && in && fir, "upfirdn2d: null pointer");
  TE_CHECK_ARG(major >= 0 && in_h > 0 && in_w > 0 && minor > 0, "upfirdn2d: bad input shape");
  TE_CHECK_ARG(kh >= 1 && kw >= 1 && kh <= 16 && kw <= 16, "upfirdn2d: FIR must be 1..16 taps per axis");
  TE_CHECK_ARG(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
  UpfirdnParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  // output size, utils/op/upfirdn2d_kernel.cu:167-168
  p.out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
  p.out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
  TE_CHECK_ARG(p.out_h > 0 && p.out_w > 0, "upfirdn2d: empty output (%d x %d)", p.out_h, p.out_w);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case TE_F32: return upfirdn2d_typed<float>(out, in, fir, p, st);
    case TE_BF16: return upfirdn2d_typed<__nv_bfloat16>(out, in, fir, p, st);
    case TE_F16: return upfirdn2d_typed<__half>(out, in, fir, p, st);
    case TE_F64: return upfirdn2d_typed<double>(out, in, fir, p, st);
  }
  set_error("upfirdn2d: unknown dtype %d", dtype);
  return TE_ERR_INVALID;
}

template <typename T>
static int upfirdn2d_bias_act_typed(void* out_, const void* in_, const float* fir, const float* bias, float slope,
                                    float gain, const te::UpfirdnParams& p, cudaStream_t st) {
  using namespace te;
  T* out = static_cast<T*>(out_);
  const T* in = static_cast<const T*>(in_);
  constexpr int NVEC = 16 / sizeof(T);
  const int64_t total = p.major * p.out_h * int64_t(p.out_w) * p.minor;
  const bool ok = p.minor % NVEC == 0 && p.kh <= 4 && p.kw <= 4 && p.minor % (128 / int(sizeof(T))) == 0 &&
                  total >= (int64_t(1) << 20) && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  int status = TE_OK;
  if (ok && launch_fir_nhwc_tma<T>(out, in, fir, p, st, &status, bias, slope, gain, 1)) {
    if (status != TE_OK) return status;
    TE_CHECK_LAUNCH();
    return TE_OK;
  }
  set_error("upfirdn2d_bias_act: geometry not covered by the TMA-staged channels-last kernel");
  return TE_ERR_UNSUPPORTED;
}

extern "C" int te_upfirdn2d_bias_act(void* out, const void* in, const float* fir, const float* bias, int64_t major,
                                     int in_h, int in_w, int minor, int kh, int kw, int pad_x0, int pad_x1, int pad_y0,
                                     int pad_y1, float slope, float gain, int dtype, void* stream) {
  using namespace te;
  if (major == 0) return TE_OK;
  TE_CHECK_ARG(out && in && fir && bias, "upfirdn2d_bias_act: null pointer");
  TE_CHECK_ARG(major >= 0 && in_h > 0 && in_w > 0 && minor > 1, "upfirdn2d_bias_act: channels-last input expected");
  TE_CHECK_ARG(kh >= 1 && kw >= 1 && kh <= 4 && kw <= 4, "upfirdn2d_bias_act: FIR of up to 4 x 4 taps");
  TE_CHECK_ARG(gain > 0.f && slope >= 0.f && slope <= 1.f, "upfirdn2d_bias_act: gain > 0 and 0 <= slope <= 1");
  UpfirdnParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
  p.up_x = p.up_y = p.down_x = p.down_y = 1;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.out_h = in_h + pad_y0 + pad_y1 - kh + 1;
  p.out_w = in_w + pad_x0 + pad_x1 - kw + 1;
  TE_CHECK_ARG(p.out_h > 0 && p.out_w > 0, "upfirdn2d_bias_act: empty output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case TE_BF16: return upfirdn2d_bias_act_typed<__nv_bfloat16>(out, in, fir, bias, slope, gain, p, st);
    case TE_F16: return upfirdn2d_bias_act_typed<__half>(out, in, fir, bias, slope, gain, p, st);
    case TE_F32: return upfirdn2d_bias_act_typed<float>(out, in, fir, bias, slope, gain, p, st);
  }
  set_error("upfirdn2d_bias_act: dtype must be bf16, f16 or f32");
  return TE_ERR_UNSUPPORTED;
}
