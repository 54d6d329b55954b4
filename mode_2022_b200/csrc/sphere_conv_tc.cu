// sphere_conv_tc.cu -- spherical convolution forward on tcgen05 tensor cores: the tangent-plane bilinear gather is
// fused into the operand staging of an implicit GEMM, so the reference's 151 MB column buffer never exists.
//
// Reference: sphere_conv_forward_cuda (sphere_conv_cuda.cpp:129-210) = sphere_im2col_gpu_kernel
// (sphere_conv_cuda_kernel.cu:195-262, bilinear :83-113) + addmm_ per batch element.
//
//   out[b, pix, o] = sum_{k, c} bilinear(x[b, :, :, c], pos[2k, pix], pos[2k+1, pix]) * W[o, c, k]
//
// GEMM view: M = 128 consecutive pixels (NHWC), N = Co, K = 9 taps x C.  For every (tap, 64-channel half) stage the
// 8 gather warps blend the four corner pixels (fp32 arithmetic, reference expression order w1*v1+w2*v2+w3*v3+w4*v4,
// reference edge rules: tap dropped unless -1 < h < H and -1 < w < W, each corner dropped when outside the image, no
// longitude wrap) into a bf16 A tile laid out for UMMA ("interleave" K-major: [8-channel chunk][pixel][16 B], chunk
// stride padded to 2064 B so the 128-bit shared stores are conflict free), the matching 16 KB weight slab is fetched
// with cp.async, and one elected thread issues 4 x tcgen05.mma (M=128, N=Co, K=16) into a TMEM accumulator.  The same
// warps run the epilogue (BN affine + residual + ReLU, bf16 NHWC store).
//
// The kernel is gather-bound, not MMA-bound: per 128-pixel tile the corner loads alone are 128 x 9 x 4 x 256 B =
// 1.18 MB of L1 traffic (~9.2k wavefront cycles) against 4.6k MMA cycles.  One CTA (16 gather warps) per SM with a
// minimal shared-memory carve-out, so that the tile's ~45 KB input neighbourhood stays L1 resident across the 36 re-reads.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
using namespace mode;

namespace {

constexpr int kGatherWarps = 16;
constexpr int kThreadsS = (kGatherWarps + 2) * 32;  // 576: 16 gather/epilogue warps, MMA warp, weight-loader warp
constexpr int kStagesS = 3;
constexpr int kChunkStrideA = 128 * 16 + 16;             // 2064 B: +16 B pad -> conflict-free STS.128 from 8 chunk-lanes
constexpr int kABytes = ((8 * kChunkStrideA) + 127) & ~127;  // 16640

// Bilinear blend r = w1*v1 + w2*v2 + w3*v3 + w4*v4 of two channels.  The weights live in the gather table as fp16 pairs
// w12 = (w1, w2), w34 = (w3, w4) for BOTH storage formats (weights are in [0, 1]: fp16 keeps 11 bits of them).
//   fp16 storage: packed HFMA2 (the broadcasts fold into operand selectors .H0_H0 / .H1_H1); the three partial sums are rounded
//                 to fp16 (2^-12 each -- together still below ONE bf16 rounding);
//   bf16 storage: unpack, four fp32 FMAs per channel in the reference's expression order (kernel.cu:111), ONE rounding.
template <int FMT>
__device__ __forceinline__ uint32_t blend2(uint32_t w12, uint32_t w34, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t v4) {
  const __half2 a = *reinterpret_cast<__half2*>(&w12), b = *reinterpret_cast<__half2*>(&w34);
  if (FMT == kFmtBF16) {
    const float2 wa = __half22float2(a), wb = __half22float2(b);
    float l1, h1, l2, h2, l3, h3, l4, h4;
    unpack2<kFmtBF16>(v1, l1, h1), unpack2<kFmtBF16>(v2, l2, h2), unpack2<kFmtBF16>(v3, l3, h3), unpack2<kFmtBF16>(v4, l4, h4);
    const float lo = fmaf(wb.y, l4, fmaf(wb.x, l3, fmaf(wa.y, l2, wa.x * l1)));
    const float hi = fmaf(wb.y, h4, fmaf(wb.x, h3, fmaf(wa.y, h2, wa.x * h1)));
    return pack_bf16x2(lo, hi);
  } else {
    const __half2 r = __hfma2(__high2half2(b), *reinterpret_cast<__half2*>(&v4),
                      __hfma2(__low2half2(b), *reinterpret_cast<__half2*>(&v3),
                      __hfma2(__high2half2(a), *reinterpret_cast<__half2*>(&v2),
                      __hmul2(__low2half2(a), *reinterpret_cast<__half2*>(&v1)))));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
}

// 16-byte read-only load executed only when `take` != 0; otherwise v keeps its previous (finite) contents
__device__ __forceinline__ void ldg_if(uint4& v, const void* ptr, uint32_t take) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
               : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)
               : "l"(ptr), "r"(take));
}

struct ScParams {
  const uint16_t* x;    // (B,H,W,C) bf16
  const int4* table;    // [9][H*W] x {top-left corner pixel index, w1|w2, w3|w4 (16-bit, storage format), 0}: built once per grid and format
  const uint16_t* wpk;  // [9][C/64][8][Co][8] bf16
  const float* scale;
  const float* shift;
  const uint16_t* res;  // (B,H,W,Co) bf16 or null
  uint16_t* out;        // (B,H,W,Co) bf16
  int B, C, H, W, Co, relu;
  long long npix;       // B*H*W
  int ntiles;
  int tw, th, tiles_x, tiles_y;  // tile = th x tw pixels (th*tw == 128); tw == 0: linear tiles of 128 consecutive pixels
  int epi_tma;                   // epilogue moves the residual / output tiles with TMA (2-D tiles, Co/ngrp == 32)
  // list mode (the slab kernel below takes every tile whose neighbourhood fits its shared-memory slab; this kernel the rest):
  const int* plist;              // tile positions (ty*tiles_x + tx) this launch processes in every image, or null = all tiles
  const int* pcount;             // device-side length of plist
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// arrive from lane 0 only, as a predicated instruction (an `if (lane == 0)` costs a divergence region per use)
__device__ __forceinline__ void mbar_arrive_lane0(uint32_t bar, int lane) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(bar), "r"(lane) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    if (BACKOFF) __nanosleep(64);  // producer threads far ahead of their consumer: do not steal issue slots from the gather warps
    if (it > (1u << 24)) {
      printf("sphere_conv_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// global pixel index (b*HW + y*W + x) of row r of tile t, or -1 beyond the tensor.  2-D tiles (th x tw) keep the tile's
// input neighbourhood compact (10 x 18 pixels instead of 3 full image rows), which is what makes the 36 corner re-reads hit L1.
struct ScParams;
__device__ __forceinline__ long long tile_pixel(const ScParams& p, int tile, int r, int HW);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFF) | (1u << 14); }
__host__ __device__ constexpr uint32_t make_idesc(int n, int fmt) { return (1u << 4) | ((fmt == 0 ? 1u : 0u) << 7) | ((fmt == 0 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

__device__ __forceinline__ long long tile_pixel(const ScParams& p, int tile, int r, int HW) {
  if (p.tw == 0) {
    const long long gp = (long long)tile * 128 + r;
    return gp < p.npix ? gp : -1;
  }
  const int per_img = p.tiles_x * p.tiles_y;
  const int b = tile / per_img, t = tile - b * per_img;
  const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
  const int y = ty * p.th + r / p.tw, x = tx * p.tw + r % p.tw;
  return (long long)b * HW + (long long)y * p.W + x;
}
// list mode: item i of this launch -> tile id (b * tiles per image + listed position)
__device__ __forceinline__ int tile_of_item(const ScParams& p, int i, int nlist) {
  if (p.plist == nullptr) return i;
  const int b = i / nlist;
  return b * (p.tiles_x * p.tiles_y) + __ldg(p.plist + (i - b * nlist));
}

// CC: compile-time channel count (128 = the model's layer4; 0 = read it from the parameters)
template <int FMT, int CC>
__global__ void __launch_bounds__(kThreadsS, 1) sphere_conv_tc_kernel(const ScParams p, const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t b_bytes = (uint32_t)p.Co * 64 * 2;          // one weight slab: Co x 64 ch
  const uint32_t stage_bytes = kABytes + b_bytes;
  // per-warp 2 KB epilogue tile (32 pixels x 32 channels, 64-byte-swizzled): residual in (TMA load), output out (TMA store)
  uint8_t* epi_s = smem + (((size_t)kStagesS * stage_bytes + 1023) & ~(size_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_s + kGatherWarps * 2048);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStagesS;
  uint64_t* tfull_bar = bars + 2 * kStagesS;       // [2]: the accumulator is double buffered
  uint64_t* tempty_bar = bars + 2 * kStagesS + 2;  // [2]
  uint64_t* res_bar = bars + 2 * kStagesS + 4;     // [kGatherWarps]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * kStagesS + 4 + kGatherWarps);
  const int tmem_cols = p.Co <= 16 ? 32 : p.Co <= 32 ? 64 : p.Co <= 64 ? 128 : p.Co <= 128 ? 256 : 512;  // 2 accumulators

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStagesS; ++i) {
      mbar_init(smem_u32(full_bar + i), kGatherWarps + 1);  // 16 gather warps + the weight loader's expect_tx arrive
      mbar_init(smem_u32(empty_bar + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(tfull_bar + i), 1);
      mbar_init(smem_u32(tempty_bar + i), kGatherWarps);
    }
    for (int i = 0; i < kGatherWarps; ++i) mbar_init(smem_u32(res_bar + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kGatherWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const int HW = p.H * p.W;
  const int C = CC ? CC : p.C;
  const int nhalf = C / 64;
  const int nstage_tile = 9 * nhalf;
  // contiguous tile ranges per CTA: consecutive tiles are consecutive image rows and share 2 of their 3 input rows in L1
  const int nlist = p.pcount ? __ldg(p.pcount) : 0;
  const int nitems = p.plist ? p.B * nlist : p.ntiles;
  const int tiles_per = (nitems + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile0 = min((int)blockIdx.x * tiles_per, nitems), tile1 = min(tile0 + tiles_per, nitems);  // ITEM range of this CTA

  if (warp < kGatherWarps) {
    // =================================================== gather producers (+ epilogue)
    const int tid = threadIdx.x;  // 0..511
    const int kc = tid & 7;       // 8-channel chunk inside the 64-channel half
    uint32_t stage = 0, tile_n = 0;
    const int ngrp = (p.Co % 128 == 0) ? 4 : 2;
    const int q = warp & 3, grp = warp >> 2;
    uint8_t* etile = epi_s + (size_t)warp * 2048;
    // TMA coordinates of this warp's 2 x 16 pixel x 32 channel piece of a tile
    auto epi_coords = [&](int tile, int& cx, int& cy) {
      const int per_img = p.tiles_x * p.tiles_y;
      const int b = tile / per_img, t = tile - b * per_img;
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      cx = tx * p.tw, cy = b * p.H + ty * p.th + (32 / p.tw) * q;
    };
    // ---- epilogue of one tile: warp w reads TMEM lanes 32*(w%4).., column group w/4 of ngrp (4 groups when Co % 128 == 0,
    // else 2).  It runs one stage into the NEXT tile's gather (the accumulator is double buffered), so the tensor pipe's
    // tail and the epilogue's own latency hide behind gather work instead of idling the L1 pipe at every tile boundary.
    // With 2-D tiles the residual piece arrives by TMA and the output piece leaves by TMA through a 64-byte-swizzled 2 KB
    // tile per warp: row-per-thread global accesses (16 B at a 256 B stride = 32 L1 wavefronts per instruction) would
    // spend a fifth of the L1 data pipe -- the kernel's bound -- on the epilogue.
    auto epilogue = [&](int item, uint32_t tn) {
      const int tile = tile_of_item(p, item, nlist);
      const uint32_t buf = tn & 1;
      mbar_wait(smem_u32(tfull_bar + buf), (tn >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * (uint32_t)p.Co + ((uint32_t)(q * 32) << 16);
      const long long gp = p.epi_tma ? 0 : tile_pixel(p, tile, q * 32 + lane, HW);
      for (int c0 = grp * (p.Co / ngrp); grp < ngrp && c0 < (grp + 1) * (p.Co / ngrp); c0 += 32) {  // exactly one pass when epi_tma
        uint32_t v[32];
        tmem_ld32(tacc + c0, v);
        if (p.epi_tma && p.res) mbar_wait(smem_u32(res_bar + warp), tn & 1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (gp >= 0) {
          const uint16_t* rp = p.res ? p.res + gp * p.Co + c0 : nullptr;
          uint16_t* op = p.out + gp * p.Co + c0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint8_t* ep = etile + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4);
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(v[g * 8 + e]);
            if (p.scale) {
              const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + g * 8)), s1 = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + g * 8) + 1);
              y[0] *= s0.x, y[1] *= s0.y, y[2] *= s0.z, y[3] *= s0.w, y[4] *= s1.x, y[5] *= s1.y, y[6] *= s1.z, y[7] *= s1.w;
            }
            if (p.shift) {
              const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + g * 8)), s1 = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + g * 8) + 1);
              y[0] += s0.x, y[1] += s0.y, y[2] += s0.z, y[3] += s0.w, y[4] += s1.x, y[5] += s1.y, y[6] += s1.z, y[7] += s1.w;
            }
            if (p.res) {
              const uint4 r = p.epi_tma ? *reinterpret_cast<const uint4*>(ep) : ld_nc_v4(rp + g * 8);
              const uint32_t r4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float r0, r1;
                unpack2<FMT>(r4[e], r0, r1);
                y[2 * e] += r0, y[2 * e + 1] += r1;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.f);
            }
            const uint4 o = make_uint4(pack2<FMT>(y[0], y[1]), pack2<FMT>(y[2], y[3]), pack2<FMT>(y[4], y[5]), pack2<FMT>(y[6], y[7]));
            if (p.epi_tma)
              *reinterpret_cast<uint4*>(ep) = o;  // same thread, same address as the residual chunk it just consumed
            else
              *reinterpret_cast<uint4*>(op + g * 8) = o;
          }
        }
        if (p.epi_tma) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            int cx, cy;
            epi_coords(tile, cx, cy);
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tm_out), "r"(c0), "r"(cx), "r"(cy), "r"(smem_u32(etile))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tempty_bar + buf));
    };
    uint4 v1[2], v2[2], v3[2], v4[2];  // corner values; loop carried so that skipped loads leave finite data behind
#pragma unroll
    for (int j = 0; j < 2; ++j) v1[j] = v2[j] = v3[j] = v4[j] = make_uint4(0, 0, 0, 0);
    for (int item = tile0; item < tile1; ++item, ++tile_n) {
      const int tile = tile_of_item(p, item, nlist);
      // this thread's two pixels of the tile: table row (pixel inside the image) and batch offset (b*HW)
      int pp[2];
      uint32_t base[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const long long gp = tile_pixel(p, tile, (tid >> 3) + 64 * j, HW);
        const int b = gp >= 0 ? (int)(gp / HW) : 0;
        pp[j] = gp >= 0 ? (int)(gp - (long long)b * HW) : -1;
        base[j] = (uint32_t)b * (uint32_t)HW;
      }
      // Stage order is HALF-major (all 9 taps of channels 0..63, then of 64..127): the 36 corner reads of a phase then
      // touch only ~3 image rows x 128 B per pixel (~50 KB), which stays L1 resident; tap-major order alternated between the
      // two halves and thrashed the ~120 KB L1 (ncu: 44 % hit rate).  Table entries of the next stage are prefetched.
      int4 ten[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) ten[j] = __ldg(p.table + ((size_t)0 * HW + max(pp[j], 0)));
      const uint32_t wmask0 = pp[0] >= 0 ? 0xffffffffu : 0u, wmask1 = pp[1] >= 0 ? 0xffffffffu : 0u;
      const ptrdiff_t rowC = (ptrdiff_t)p.W * C;
      for (int s = 0; s < nstage_tile; ++s, ++stage) {
        if (s == 1 && item > tile0) epilogue(item - 1, tile_n - 1);
        if (s == 4 && p.epi_tma && p.res != nullptr && grp < ngrp && lane == 0) {
          // this tile's residual piece -> the warp's epilogue tile (free once the previous tile's output store has read it)
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          int cx, cy;
          epi_coords(tile, cx, cy);
          const uint32_t bar = smem_u32(res_bar + warp);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2048) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(etile)),
                       "l"(&tm_res), "r"(grp * 32), "r"(cx), "r"(cy), "r"(bar)
                       : "memory");
        }
        const int half = s / 9;
        uint32_t w12[2], w34[2];  // bilinear weights, 16-bit, in the layer's storage format
        const int coff = half * 64 + kc * 8;
#pragma unroll
        for (int j = 0; j < 2; ++j) {  // all 8 corner loads in flight before the first use
          // Corners with weight exactly 0 are not fetched: the dropped ones (reference edge rules), and 7 of the 36 on a
          // Cassini / ERP grid -- the centre tap and the two taps on the pixel's own meridian sample at integer coordinates.
          // Their registers keep an older (finite) value, and 0 * finite == 0.  (The reference multiplies the kept ones by 0;
          // the only observable difference would be a NaN from an Inf/NaN input.)
          w12[j] = (uint32_t)ten[j].y & (j ? wmask1 : wmask0), w34[j] = (uint32_t)ten[j].z & (j ? wmask1 : wmask0);
          const uint16_t* xp = p.x + (ptrdiff_t)(((int)base[j] + ten[j].x) * C + coff);  // may be negative for a dropped top-left corner
          ldg_if(v1[j], xp, w12[j] & 0xffffu);
          ldg_if(v2[j], xp + C, w12[j] >> 16);
          ldg_if(v3[j], xp + rowC, w34[j] & 0xffffu);
          ldg_if(v4[j], xp + rowC + C, w34[j] >> 16);
        }
        if (s + 1 < nstage_tile) {  // prefetch the next stage's table entries (in flight during the blend)
          const int kn = (s + 1) % 9;
#pragma unroll
          for (int j = 0; j < 2; ++j) ten[j] = __ldg(p.table + ((size_t)kn * HW + max(pp[j], 0)));
        }
        {
          const uint32_t slot = stage % kStagesS, phase = (stage / kStagesS) & 1;
          mbar_wait(smem_u32(empty_bar + slot), phase ^ 1);
          uint8_t* a_s = smem + (size_t)slot * stage_bytes;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint4 o;
            o.x = blend2<FMT>(w12[j], w34[j], v1[j].x, v2[j].x, v3[j].x, v4[j].x);
            o.y = blend2<FMT>(w12[j], w34[j], v1[j].y, v2[j].y, v3[j].y, v4[j].y);
            o.z = blend2<FMT>(w12[j], w34[j], v1[j].z, v2[j].z, v3[j].z, v4[j].z);
            o.w = blend2<FMT>(w12[j], w34[j], v1[j].w, v2[j].w, v3[j].w, v4[j].w);
            const int pix_l = (tid >> 3) + 64 * j;
            *reinterpret_cast<uint4*>(a_s + kc * kChunkStrideA + pix_l * 16) = o;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(full_bar + slot));
        }
      }
    }
    if (tile1 > tile0) epilogue(tile1 - 1, tile_n - 1);
    if (p.epi_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp == kGatherWarps + 1) {
    // =================================================== weight loader: one bulk copy (TMA, 1-D) per stage, up to kStagesS ahead
    if (lane == 0) {
      uint32_t stage = 0;
      for (int tile = tile0; tile < tile1; ++tile) {
        for (int s2 = 0; s2 < nstage_tile; ++s2, ++stage) {
          const int half = s2 / 9, k = s2 - half * 9;
          const uint32_t slot = stage % kStagesS, phase = (stage / kStagesS) & 1;
          mbar_wait(smem_u32(empty_bar + slot), phase ^ 1);
          const uint32_t bar = smem_u32(full_bar + slot);
          const uint32_t dst = smem_u32(smem + (size_t)slot * stage_bytes + kABytes);
          const uint16_t* wsrc = p.wpk + ((size_t)(k * nhalf + half) * b_bytes) / 2;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b_bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(wsrc), "r"(b_bytes), "r"(bar)
                       : "memory");
        }
      }
    }
  } else {
    // =================================================== MMA issuer (warp-uniform control flow, one elected lane issues)
    const uint32_t idesc = make_idesc(p.Co, FMT);
    const uint32_t a_hi = desc_hi(128), b_hi = desc_hi(128);
    uint32_t stage = 0, tile_n = 0;
    for (int tile = tile0; tile < tile1; ++tile, ++tile_n) {
      const uint32_t buf = tile_n & 1;
      mbar_wait(smem_u32(tempty_bar + buf), ((tile_n >> 1) & 1) ^ 1);  // epilogue of tile - 2 has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * (uint32_t)p.Co;
      for (int s = 0; s < nstage_tile; ++s, ++stage) {
        const uint32_t slot = stage % kStagesS, phase = (stage / kStagesS) & 1;
        mbar_wait(smem_u32(full_bar + slot), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = desc_lo(smem_u32(smem + (size_t)slot * stage_bytes), kChunkStrideA);
          const uint32_t b0 = desc_lo(smem_u32(smem + (size_t)slot * stage_bytes + kABytes), (uint32_t)p.Co * 16);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma(tacc, a0 + ((ks * 2 * kChunkStrideA) >> 4), a_hi, b0 + ((uint32_t)(ks * 2 * p.Co * 16) >> 4), b_hi, idesc, (s | ks) ? 1u : 0u);
          umma_commit(smem_u32(empty_bar + slot));
          if (s == nstage_tile - 1) umma_commit(smem_u32(tfull_bar + buf));
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kGatherWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

// gather table: for every (tap, pixel) the top-left corner's pixel index (the other three are +1, +W, +W+1), the four
// bilinear weights as fp16 (both storage formats blend with them, see blend2), with the reference's rules folded in
// (kernel.cu:246 tap guard, :97-107 per-corner guards -> weight 0, :109 weights), and the corner's (row, col) packed as two
// int16 for the slab kernel.  A corner of weight 0 is never fetched, so its index may point outside the image.  16 bytes per
// entry; depends only on the sampling grid.
__global__ void sphere_table_kernel(const float* __restrict__ pos, int4* __restrict__ table, int H, int W, int KK) {
  const int HW = H * W;
  const long long n = (long long)KK * HW;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e / HW), pp = (int)(e - (long long)k * HW);
    const float h_im = pos[(size_t)(2 * k) * HW + pp], w_im = pos[(size_t)(2 * k + 1) * HW + pp];
    int idx = 0, hl = pp / W, wl = pp - (pp / W) * W;  // dropped tap: all weights 0, corner = the pixel itself
    float4 wt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h_im > -1 && w_im > -1 && h_im < H && w_im < W) {
      const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
      const int h_high = h_low + 1, w_high = w_low + 1;
      const float lh = h_im - h_low, lw = w_im - w_low, hh = 1 - lh, hw = 1 - lw;
      idx = h_low * W + w_low, hl = h_low, wl = w_low;
      if (h_low >= 0 && w_low >= 0) wt.x = hh * hw;
      if (h_low >= 0 && w_high <= W - 1) wt.y = hh * lw;
      if (h_high <= H - 1 && w_low >= 0) wt.z = lh * hw;
      if (h_high <= H - 1 && w_high <= W - 1) wt.w = lh * lw;
    }
    const uint32_t w12 = (uint32_t)float_to_h16_bits(wt.x, kFmtFP16) | ((uint32_t)float_to_h16_bits(wt.y, kFmtFP16) << 16);
    const uint32_t w34 = (uint32_t)float_to_h16_bits(wt.z, kFmtFP16) | ((uint32_t)float_to_h16_bits(wt.w, kFmtFP16) << 16);
    table[e] = make_int4(idx, (int)w12, (int)w34, (int)(((uint32_t)hl << 16) | ((uint32_t)wl & 0xffffu)));
  }
}

// ---- tile classes -------------------------------------------------------------------------------------------------------
// The slab kernel stages the input neighbourhood of a 128-pixel tile in shared memory.  Per tile position (the same for every
// image) this kernel measures that neighbourhood from the gather table: along the LONG image axis (longitude: rows of a Cassini
// map, columns of an ERP map -- the axis the sampling grid wraps around and along which the footprint grows like 1/cos(lat))
// the first line l0 and the line count L, modulo the axis length; along the short axis the first pixel s0 of a kSlabShort window.
// Tiles with L <= kSlabLines whose short extent fits go to the slab kernel ("fast"), the rest (the polar tile columns, where
// one tap reaches half-way round the sphere) to the direct-gather kernel above.
constexpr int kSlabShort = 12;                        // tile short side 8 + 2 + 2: taps reach [-2, +2] columns (rows for ERP)
constexpr int kSlabLineBytes = kSlabShort * 128;      // one LOADED line of the slab: 12 pixels x 64 channels x 2 B
// Slab line pitch in pixels and line capacity, by orientation.  Gather lanes of a quarter-warp are 8 neighbouring pixels along the
// SHORT axis for a Cassini map: a pitch that is a multiple of 8 pixels makes their eight 16-byte swizzle positions distinct even
// when neighbouring columns sample different lines (the footprint grows towards the poles), i.e. conflict-free LDS.128.  For an
// ERP map the lanes run along the long axis (one LINE per lane): an odd pitch walks the eight positions instead.
__host__ __device__ constexpr int slab_pitch(bool cassini) { return cassini ? 16 : 13; }
__host__ __device__ constexpr int slab_lines(bool cassini) { return cassini ? 28 : 32; }   // 16 + 2*6: footprints up to +-5.8 lines (7.5 for ERP)
constexpr int kSlabBufBytes = 28 * 16 * 128;          // 57344 per (tile, channel half) >= 32 * 13 * 128, 1024-aligned

__global__ void sphere_tileinfo_kernel(const int4* __restrict__ table, int4* __restrict__ info, int H, int W, int KK, int TH, int TW) {
  const int kSlabLines = slab_lines(TH > TW);
  __shared__ int s_min_l, s_max_l, s_min_s, s_max_s, s_bad;
  const int tiles_x = W / TW;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const bool cassini = TH > TW;  // long axis = rows
  const int LEN = cassini ? H : W, tl0 = cassini ? ty * TH : tx * TW;
  if (threadIdx.x == 0) s_min_l = 1 << 30, s_max_l = -(1 << 30), s_min_s = 1 << 30, s_max_s = -(1 << 30), s_bad = 0;
  __syncthreads();
  const int m = threadIdx.x, h = ty * TH + m / TW, w = tx * TW + m % TW;
  int min_l = 1 << 30, max_l = -(1 << 30), min_s = 1 << 30, max_s = -(1 << 30), bad = 0;
  for (int k = 0; k < KK; ++k) {
    const int4 e = table[(size_t)k * H * W + (size_t)h * W + w];
    const uint32_t wb[4] = {(uint32_t)e.y & 0xffffu, (uint32_t)e.y >> 16, (uint32_t)e.z & 0xffffu, (uint32_t)e.z >> 16};
    if ((wb[0] | wb[1] | wb[2] | wb[3]) == 0) continue;  // dropped tap: nothing is fetched
    const int hl = e.w >> 16, wl = (int)(short)(e.w & 0xffff);
    if (hl < 0 || wl < 0) bad = 1;  // top-left corner outside the image: the slab address arithmetic assumes it is inside
    for (int c = 0; c < 4; ++c) {
      if (c != 0 && wb[c] == 0) continue;  // the top-left corner anchors the address arithmetic: always inside the slab
      const int ch = hl + (c >> 1), cw = wl + (c & 1);
      const int lng = cassini ? ch : cw, sht = cassini ? cw : ch;
      int dl = (lng - tl0 + LEN / 2) % LEN;  // signed offset from the tile's first line, wrapped into [-LEN/2, LEN/2)
      if (dl < 0) dl += LEN;
      dl -= LEN / 2;
      min_l = min(min_l, dl), max_l = max(max_l, dl), min_s = min(min_s, sht), max_s = max(max_s, sht);
    }
  }
  atomicMin(&s_min_l, min_l), atomicMax(&s_max_l, max_l), atomicMin(&s_min_s, min_s), atomicMax(&s_max_s, max_s);
  if (bad) atomicOr(&s_bad, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int4 r = make_int4(0, 0, 0, (ty << 16) | tx);  // z = 0: not a slab tile
    if (s_max_l < s_min_l) {
      r.z = 1;  // no tap at all in this tile: one (unused) line
    } else if (!s_bad && s_max_l - s_min_l + 1 <= kSlabLines && 2 * (s_max_l - s_min_l + 1) <= LEN && s_max_s - s_min_s + 1 <= kSlabShort) {
      // (2L <= LEN: the lines of the slab are distinct and consecutive modulo the axis length, so the (line + 1) corner of a tap is
      //  the NEXT slab line; a footprint that wraps more than half-way round a small map goes to the direct-gather kernel)
      int l0 = (tl0 + s_min_l) % LEN;
      if (l0 < 0) l0 += LEN;
      r.x = l0, r.y = s_min_s, r.z = s_max_l - s_min_l + 1;
    }
    info[blockIdx.x] = r;
  }
}
// header {n_fast, n_rest, TH, TW} + the two position lists (deterministic order)
__global__ void sphere_tilelist_kernel(const int4* __restrict__ info, int* __restrict__ hdr, int* __restrict__ fast, int* __restrict__ rest, int npos, int TH, int TW) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int nf = 0, nr = 0;
  for (int i = 0; i < npos; ++i) {
    if (info[i].z > 0)
      fast[nf++] = i;
    else
      rest[nr++] = i;
  }
  hdr[0] = nf, hdr[1] = nr, hdr[2] = TH, hdr[3] = TW;
}

// (Co, C, 3, 3) fp32 -> [tap 9][half C/64][chunk 8][n Co][8] bf16
__global__ void pack_wsphere_kernel(const float* __restrict__ w, uint16_t* __restrict__ wp, int C, int Co, int fmt, long long total) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long r = e;
    const int j = (int)(r % 8);
    r /= 8;
    const int n = (int)(r % Co);
    r /= Co;
    const int kc = (int)(r % 8);
    r /= 8;
    const int half = (int)(r % (C / 64));
    const int k = (int)(r / (C / 64));
    const int c = half * 64 + kc * 8 + j;
    wp[e] = float_to_h16_bits(w[((size_t)n * C + c) * 9 + k], fmt);
  }
}

decltype(&cuTensorMapEncodeTiled) tmap_encoder() {
  static decltype(&cuTensorMapEncodeTiled) encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      set_error("sphere_conv_tc: cuTensorMapEncodeTiled is not available from this driver");
      return nullptr;
    }
    encode = reinterpret_cast<decltype(&cuTensorMapEncodeTiled)>(fn);
  }
  return encode;
}

// epilogue tiles: (Co, W, B*H) view of an NHWC tensor, box 32 channels x tw columns x 32/tw rows (one warp's 32 pixels), 64-byte swizzle
int make_epi_tmap(CUtensorMap* tm, const void* ptr, int fmt, int Co, int W, long long rows, int tw) {
  auto encode = tmap_encoder();
  if (!encode) return MODE_ECUDA;
  const cuuint64_t gdim[3] = {(cuuint64_t)Co, (cuuint64_t)W, (cuuint64_t)rows};
  const cuuint64_t gstr[2] = {(cuuint64_t)Co * 2, (cuuint64_t)W * Co * 2};
  const cuuint32_t box[3] = {32, (cuuint32_t)tw, (cuuint32_t)(32 / tw)}, estr[3] = {1, 1, 1};
  const CUresult r = encode(tm, fmt == kFmtBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("sphere_conv_tc: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return MODE_ECUDA;
  }
  return MODE_OK;
}

// slab lines: (C, W, H, B) view of the NHWC input, box = 64 channels x kSlabShort pixels along the SHORT image axis x one line of
// the long axis, 128-byte swizzle, zero fill outside the image (= the reference's zero corners, kernel.cu:97-107)
int make_slab_tmap(CUtensorMap* tm, const void* ptr, int fmt, int B, int C, int H, int W, bool cassini) {
  auto encode = tmap_encoder();
  if (!encode) return MODE_ECUDA;
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {64, cassini ? (cuuint32_t)kSlabShort : 1u, cassini ? 1u : (cuuint32_t)kSlabShort, 1}, estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(tm, fmt == kFmtBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("sphere_conv_tc: cuTensorMapEncodeTiled (slab) failed (CUresult %d)", (int)r);
    return MODE_ECUDA;
  }
  return MODE_OK;
}

// =============================================================================================================================
// Slab kernel: the gather reads SHARED memory and the A operand lives in TENSOR memory.
//
// The direct-gather kernel above spends its time in the L1 pipe and in instruction issue: per 128-pixel tile it issues 65 k warp
// instructions (225 per warp and stage: 64-bit address arithmetic, a table entry per 8-channel chunk, predicates) and its 36 corner
// loads per pixel are L1/L2 round trips (ncu r01: issue 52 %, L1 data pipe 57 %, long-scoreboard stalls, tensor pipe 10 %).  Here:
//   * a loader thread stages the tile's input neighbourhood (<= 32 lines x 12 pixels x 64 channels, 128-byte-swizzled) in shared
//     memory with one TMA box per line of the long (wrapping) image axis -- the longitude wrap of the sampling grid is a different
//     line coordinate, out-of-image pixels are TMA zero fill;
//   * gather threads own one PIXEL each (= one TMEM lane) and 16 channels: one table entry per (pixel, tap), 8 conflict-free
//     LDS.128 (corner x chunk; zero-weight corners predicated off), the blend, and ONE tcgen05.st of the 16 blended channels
//     straight into the A operand in tensor memory -- no A tile in shared memory, so the 128 B/clk shared-memory port carries only
//     the corner loads, the weight slabs (B operand) and their TMA fills;
//   * tcgen05.mma reads A from TMEM (.kind::f16 "TS" form), B from the shared-memory ring; accumulators double-buffered in TMEM.
// Tiles whose neighbourhood does not fit (polar tile columns: a tap reaches +-128 lines) are left to the direct-gather kernel.
// =============================================================================================================================
constexpr int kFStages = 5;                              // ring depth: A slots in TMEM (32 columns each) and weight slabs in smem
constexpr int kFSets = 3;                                // warp sets of the gather role: set j produces the stages j, j + kFSets, ...
constexpr int kFGatherWarps = 4 * kFSets;                // a set = 4 warps = the 4 TMEM lane quadrants (32 pixels each), 64 channels per thread
constexpr int kFPieces = 16;                             // epilogue pieces: 4 lane quadrants x 4 groups of 32 output channels, 2 KB each
constexpr int kFThreads = (kFGatherWarps + 3) * 32;      // gather/epilogue warps, MMA warp, weight loader, slab loader
constexpr int kFBBytes = 128 * 64 * 2;                   // one weight slab: Co = 128 x 64 channels
constexpr int kTmemAcc = 256;                            // two 128-column accumulators, then the A ring

struct FcParams {
  const int4* table;    // [9][H*W] gather entries
  const int* hdr;       // {n_fast, n_rest, TH, TW}, then int4 info[npos], then the fast / rest position lists
  const uint16_t* wpk;  // [9][C/64][8][128][8]
  const float* scale;
  const float* shift;
  int has_res, relu;
  int B, C, H, W, npos;
  int th, tw, tw_shift, cassini;
};

__device__ __forceinline__ void lds_if(uint4& v, uint32_t addr, uint32_t take) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
               : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)
               : "r"(addr), "r"(take));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
      "%26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
      "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),
      "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}

template <int FMT, int CASSINI>
__global__ void __launch_bounds__(kFThreads, 1) sphere_conv_slab_kernel(const FcParams p, const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_out,
                                                                        const __grid_constant__ CUtensorMap tm_res) {
  extern __shared__ uint8_t smem_raw[];
  // The two launches of one layer call write disjoint tiles of the output and only read the layer's input: the direct-gather launch that
  // follows is launched with programmatic stream serialization and fills SMs as soon as this grid's CTAs retire (238 -> 231 us per call).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // shared-memory map (byte offsets from a 1024-aligned base: TMA swizzle atoms are 1024-byte aligned); all addressing below is
  // 32-bit shared-window arithmetic on `sm`
  const uint32_t sm = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t kOffB = 2 * kSlabBufBytes;                    // weight ring: kFStages x kFBBytes
  constexpr uint32_t kOffEpi = kOffB + kFStages * kFBBytes;        // kFPieces x 2 KB epilogue tiles (32 pixels x 32 channels, 64-byte-swizzled)
  constexpr uint32_t kOffBar = kOffEpi + kFPieces * 2048;
  const uint32_t full_bar = sm + kOffBar;                          // [kFStages] A slot written (4 warps) + weight slab landed (tx)
  const uint32_t empty_bar = full_bar + 8 * kFStages;              // [kFStages] the MMAs that read the slot have completed
  const uint32_t sfull_bar = empty_bar + 8 * kFStages;             // [2] slab landed (tx)
  const uint32_t sempty_bar = sfull_bar + 16;                      // [2] slab consumed (all gather warps)
  const uint32_t tfull_bar = sempty_bar + 16;                      // [2] accumulator complete
  const uint32_t tempty_bar = tfull_bar + 16;                      // [2] accumulator drained (all gather warps)
  const uint32_t res_bar = tempty_bar + 16;                        // [kFPieces] residual piece landed
  const uint32_t tmem_ptr_u32 = res_bar + 8 * kFPieces;
  uint8_t* const sm_gen = smem_raw + (sm - smem_u32(smem_raw));    // generic pointer to the aligned base

  if (threadIdx.x == 0) {
    for (int i = 0; i < kFStages; ++i) {
      mbar_init(full_bar + 8 * i, 4 + 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(sfull_bar + 8 * i, 1);
      mbar_init(sempty_bar + 8 * i, kFGatherWarps);
      mbar_init(tfull_bar + 8 * i, 1);
      mbar_init(tempty_bar + 8 * i, kFGatherWarps);
    }
    for (int i = 0; i < kFPieces; ++i) mbar_init(res_bar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kFGatherWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_u32), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(sm_gen + (tmem_ptr_u32 - sm));
  const int HW = p.H * p.W;
  const int nhalf = p.C / 64;
  const int nst = 9 * nhalf;                     // stages per tile, HALF-major: all 9 taps of channels 0..63, then of 64..127
  const int nfast = __ldg(p.hdr);
  const int nitems = p.B * nfast;
  const int4* info = reinterpret_cast<const int4*>(p.hdr + 4);
  const int* flist = p.hdr + 4 + 4 * p.npos;
  const int LEN = CASSINI ? p.H : p.W;
  constexpr int kSlabPitch = slab_pitch(CASSINI != 0), kSlabPitchBytes = kSlabPitch * 128;
  const int my_tiles = (int)blockIdx.x < nitems ? (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp < kFGatherWarps) {
    // =================================================== gather producers (+ epilogue)
    // Gather role: kFSets sets of 4 warps; set j produces the stages j, j + kFSets, ... of the CTA's stage sequence.  A thread owns
    // one PIXEL (TMEM lane) and all 64 channels of the stage: per (pixel, tap) ONE table entry, ONE address computation, 32
    // predicated LDS.128 (corner x chunk), the blend and ONE tcgen05.st of 32 packed columns.
    // Epilogue role: all warps; pieces = lane quadrant q x output-channel group eg (32 channels), groups dealt out over the sets.
    const int q = warp & 3, set = warp >> 2;
    const int eg0 = set * 4 / kFSets, eg1 = (set + 1) * 4 / kFSets;
    const int m = q * 32 + lane;                 // GEMM row = TMEM lane = pixel of the tile, short image axis fastest
    const int mr = m >> p.tw_shift, mc = m & (p.tw - 1);
    constexpr int dW = CASSINI ? 1 : kSlabPitch, dH = CASSINI ? kSlabPitch : 1;  // slab pixel steps of the (row, col+1) and (row+1, col) corners
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    auto decode = [&](int tn, int& b, int4& inf) {
      const int item = (int)blockIdx.x + tn * (int)gridDim.x;
      b = item / nfast;
      inf = __ldg(info + __ldg(flist + (item - b * nfast)));
    };
    auto epi_coords = [&](int b, const int4& inf, int& cx, int& cy) {
      const int ty = inf.w >> 16, tx = inf.w & 0xffff;
      cx = tx * p.tw, cy = b * p.H + ty * p.th + (32 >> p.tw_shift) * q;
    };
    // epilogue of tile tn: TMEM lanes 32q.., accumulator columns 32eg..32eg+31 -> affine + residual + ReLU -> the piece's swizzled 2 KB
    // tile -> TMA store.  Called >= 6 stages into the NEXT tile's gather: the A ring is kFStages = 5 deep, so by then every MMA of
    // tile tn has completed (no wait), and the double-buffered accumulator keeps the tensor pipe busy meanwhile.
    auto epilogue = [&](int tn) {
      int b;
      int4 inf;
      decode(tn, b, inf);
      const uint32_t buf = (uint32_t)tn & 1u;
      mbar_wait(tfull_bar + 8 * buf, ((uint32_t)tn >> 1) & 1u);
      tc_fence_after();
      for (int eg = eg0; eg < eg1; ++eg) {
        const int piece = eg * 4 + q, c0 = eg * 32;
        uint8_t* const etile_gen = sm_gen + kOffEpi + (size_t)piece * 2048;
        uint32_t v[32];
        tmem_ld32(tmem_base + buf * 128u + (uint32_t)c0 + lane_off, v);
        if (p.has_res) mbar_wait(res_bar + 8 * piece, (uint32_t)tn & 1u);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          uint8_t* ep = etile_gen + lane * 64 + ((gg ^ ((lane >> 1) & 3)) << 4);
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(v[gg * 8 + e]);
          if (p.scale) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + gg * 8)), s1 = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + gg * 8) + 1);
            y[0] *= s0.x, y[1] *= s0.y, y[2] *= s0.z, y[3] *= s0.w, y[4] *= s1.x, y[5] *= s1.y, y[6] *= s1.z, y[7] *= s1.w;
          }
          if (p.shift) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + gg * 8)), s1 = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + gg * 8) + 1);
            y[0] += s0.x, y[1] += s0.y, y[2] += s0.z, y[3] += s0.w, y[4] += s1.x, y[5] += s1.y, y[6] += s1.z, y[7] += s1.w;
          }
          if (p.has_res) {
            const uint4 r = *reinterpret_cast<const uint4*>(ep);
            const uint32_t r4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float r0, r1;
              unpack2<FMT>(r4[e], r0, r1);
              y[2 * e] += r0, y[2 * e + 1] += r1;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.f);
          }
          *reinterpret_cast<uint4*>(ep) = make_uint4(pack2<FMT>(y[0], y[1]), pack2<FMT>(y[2], y[3]), pack2<FMT>(y[4], y[5]), pack2<FMT>(y[6], y[7]));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        int cx, cy;
        epi_coords(b, inf, cx, cy);
        for (int eg = eg0; eg < eg1; ++eg)
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tm_out), "r"(eg * 32), "r"(cx), "r"(cy),
                       "r"(sm + kOffEpi + (uint32_t)(eg * 4 + q) * 2048u)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * buf);
    };
    // this tile's residual pieces -> the warp's epilogue tiles (free once the previous tile's output stores have read them)
    auto load_residual = [&](int b, const int4& inf) {
      if (p.has_res && lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        int cx, cy;
        epi_coords(b, inf, cx, cy);
        for (int eg = eg0; eg < eg1; ++eg) {
          const uint32_t bar = res_bar + 8 * (eg * 4 + q);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2048) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                           sm + kOffEpi + (uint32_t)(eg * 4 + q) * 2048u),
                       "l"(&tm_res), "r"(eg * 32), "r"(cx), "r"(cy), "r"(bar)
                       : "memory");
        }
      }
    };

    uint4 v1[2], v2[2], v3[2], v4[2];  // corner values of one chunk pair; a skipped load leaves an older finite value, times weight 0
#pragma unroll
    for (int j = 0; j < 2; ++j) v1[j] = v2[j] = v3[j] = v4[j] = make_uint4(0, 0, 0, 0);
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)nst;
    int tn = 0, s = set;                         // (tile, stage in tile) of this warp's current stage (nst >= 9 > kFSets)
    int b = 0, l0 = 0, s0 = 0, pix = 0;
    int4 inf = make_int4(0, 0, 0, 0), ent = make_int4(0, 0, 0, 0);
    int cur_tn = -1;
    bool epi_done = true, res_issued = true;
    uint32_t pend_bar = 0;                       // full barrier of my previous stage, not yet arrived on
    uint32_t slot = (uint32_t)set, phase = 0;    // A ring slot / parity of stage gs (gs % kFStages, (gs / kFStages) & 1), kept incrementally
    for (uint32_t gs = (uint32_t)set; gs < total; gs += kFSets) {
      if (tn != cur_tn) {  // first stage of mine in a new tile
        cur_tn = tn;
        decode(tn, b, inf);
        l0 = inf.x, s0 = inf.y;
        pix = ((inf.w >> 16) * p.th + mr) * p.W + (inf.w & 0xffff) * p.tw + mc;  // this thread's pixel inside the image
        const int k0 = s >= 9 ? s - 9 : s;
        ent = __ldg(p.table + (size_t)k0 * HW + pix);
        epi_done = false, res_issued = false;
      }
      const int hf = s >= 9 ? 1 : 0, k = s - 9 * hf;
      const uint32_t slabn = (uint32_t)(tn * nhalf + hf), sb = slabn & 1u;
      if (k < kFSets) mbar_wait(sfull_bar + 8 * sb, (slabn >> 1) & 1u);  // my first stage in this (tile, half): its slab has landed
      if (!epi_done && s >= 6) {
        if (tn > 0) epilogue(tn - 1);
        epi_done = true;
        if (s + kFSets >= nst) load_residual(b, inf), res_issued = true;
      } else if (epi_done && !res_issued) {
        load_residual(b, inf), res_issued = true;
      }
      // slab address of the top-left corner: line (long-axis coordinate - l0, modulo the axis: the grid wraps), then the short axis
      const int hl = ent.w >> 16, wl = (int)(short)(ent.w & 0xffff);
      int d = (CASSINI ? hl : wl) - l0;
      d += (d >> 31) & LEN;
      const int p1 = d * kSlabPitch + ((CASSINI ? wl : hl) - s0);
      const uint32_t w12 = (uint32_t)ent.y, w34 = (uint32_t)ent.z;
      {  // next stage of mine in this tile: tap k + kFSets (mod 9: the following half starts over)
        const int kn = k + kFSets >= 9 ? k + kFSets - 9 : k + kFSets;
        if (s + kFSets < nst) ent = __ldg(p.table + (size_t)kn * HW + pix);
      }
      const uint32_t slab = sm + sb * kSlabBufBytes;
      const int pc[4] = {p1, p1 + dW, p1 + dH, p1 + dH + dW};
      const uint32_t take[4] = {w12 & 0xffffu, w12 >> 16, w34 & 0xffffu, w34 >> 16};
      uint32_t ca[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)  // pixel pc = 128 bytes; its 16-byte chunk cc sits at position cc ^ (pixel & 7) (TMA 128-byte swizzle); this is chunk 0
        ca[c] = slab + ((uint32_t)pc[c] << 7) + ((uint32_t)(pc[c] & 7) << 4);
      uint32_t o[32];
#pragma unroll
      for (int h2 = 0; h2 < 4; ++h2) {  // chunk pairs (2 h2, 2 h2 + 1): positions ca ^ 32 h2, ca ^ (32 h2 + 16)
        lds_if(v1[0], ca[0] ^ (32u * h2), take[0]);
        lds_if(v1[1], ca[0] ^ (32u * h2 + 16u), take[0]);
        lds_if(v2[0], ca[1] ^ (32u * h2), take[1]);
        lds_if(v2[1], ca[1] ^ (32u * h2 + 16u), take[1]);
        lds_if(v3[0], ca[2] ^ (32u * h2), take[2]);
        lds_if(v3[1], ca[2] ^ (32u * h2 + 16u), take[2]);
        lds_if(v4[0], ca[3] ^ (32u * h2), take[3]);
        lds_if(v4[1], ca[3] ^ (32u * h2 + 16u), take[3]);
        if (h2 == 0 && pend_bar != 0) {
          // publish my PREVIOUS stage: its tcgen05.st has had a whole loop turn-around (and these loads) to complete
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          mbar_arrive_lane0(pend_bar, lane);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          o[8 * h2 + 4 * j + 0] = blend2<FMT>(w12, w34, v1[j].x, v2[j].x, v3[j].x, v4[j].x);
          o[8 * h2 + 4 * j + 1] = blend2<FMT>(w12, w34, v1[j].y, v2[j].y, v3[j].y, v4[j].y);
          o[8 * h2 + 4 * j + 2] = blend2<FMT>(w12, w34, v1[j].z, v2[j].z, v3[j].z, v4[j].z);
          o[8 * h2 + 4 * j + 3] = blend2<FMT>(w12, w34, v1[j].w, v2[j].w, v3[j].w, v4[j].w);
        }
      }
      if (k >= 9 - kFSets) {  // my last stage in this (tile, half): every load of the slab has been consumed by a blend
        __syncwarp();
        mbar_arrive_lane0(sempty_bar + 8 * sb, lane);
      }
      mbar_wait(empty_bar + 8 * slot, phase ^ 1u);  // the MMAs that read this A slot kFStages stages ago are complete
      tc_fence_after();
      // 64 channels of this pixel = 32 packed columns of TMEM lane m, A-operand slot `slot`; published one stage later (above)
      tmem_st32(tmem_base + kTmemAcc + slot * 32u + lane_off, o);
      pend_bar = full_bar + 8 * slot;
      slot += kFSets;
      if (slot >= kFStages) slot -= kFStages, phase ^= 1u;
      s += kFSets;
      if (s >= nst) s -= nst, ++tn;
    }
    if (pend_bar != 0) {  // publish my last stage
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      mbar_arrive_lane0(pend_bar, lane);
    }
    // drain: the last tile's epilogue (and, for a warp set that never reached stage 6 of it, the one before)
    if (my_tiles > 0) {
      if (!epi_done && my_tiles > 1) epilogue(my_tiles - 2);
      if (!res_issued) {
        decode(my_tiles - 1, b, inf);
        load_residual(b, inf);
      }
      epilogue(my_tiles - 1);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp == kFGatherWarps + 1) {
    // =================================================== weight loader: one 16 KB bulk copy per stage, up to kFStages ahead
    if (lane == 0) {
      const uint32_t total = (uint32_t)my_tiles * (uint32_t)nst;
      int s = 0;
      for (uint32_t gs = 0; gs < total; ++gs) {
        const int hf = s >= 9 ? 1 : 0, k = s - 9 * hf;
        const uint32_t slot = gs % kFStages, phase = (gs / kFStages) & 1u;
        mbar_wait<true>(empty_bar + 8 * slot, phase ^ 1u);
        const uint32_t bar = full_bar + 8 * slot;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kFBBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm + kOffB + slot * kFBBytes),
                     "l"(p.wpk + (size_t)(k * nhalf + hf) * (kFBBytes / 2)), "r"(kFBBytes), "r"(bar)
                     : "memory");
        if (++s == nst) s = 0;
      }
    }
  } else if (warp == kFGatherWarps + 2) {
    // =================================================== slab loader: one TMA box per line of the long axis, one (tile, half) ahead
    if (lane == 0) {
      uint32_t slabn = 0;
      for (int tn = 0; tn < my_tiles; ++tn) {
        const int item = (int)blockIdx.x + tn * (int)gridDim.x;
        const int b = item / nfast;
        const int4 inf = __ldg(info + __ldg(flist + (item - b * nfast)));
        const int l0 = inf.x, s0 = inf.y, L = inf.z;
        for (int hf = 0; hf < nhalf; ++hf, ++slabn) {
          const uint32_t sb = slabn & 1u;
          mbar_wait<true>(sempty_bar + 8 * sb, ((slabn >> 1) & 1u) ^ 1u);
          const uint32_t bar = sfull_bar + 8 * sb;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(L * kSlabLineBytes) : "memory");
          const uint32_t dst0 = sm + sb * kSlabBufBytes;
          for (int l = 0; l < L; ++l) {
            int line = l0 + l;
            line -= line >= LEN ? LEN : 0;  // the sampling grid wraps around the long axis (sphere_conv.py:225)
            const int cw = CASSINI ? s0 : line, chh = CASSINI ? line : s0;
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst0 + l * kSlabPitchBytes),
                         "l"(&tm_x), "r"(hf * 64), "r"(cw), "r"(chh), "r"(b), "r"(bar)
                         : "memory");
          }
        }
      }
    }
  } else {
    // =================================================== MMA issuer (warp-uniform control flow, one elected lane issues)
    const uint32_t idesc = make_idesc(128, FMT);
    const uint32_t b_hi = desc_hi(128);
    uint32_t gs = 0;
    for (int tn = 0; tn < my_tiles; ++tn) {
      const uint32_t buf = (uint32_t)tn & 1u;
      mbar_wait(tempty_bar + 8 * buf, (((uint32_t)tn >> 1) & 1u) ^ 1u);  // epilogue of tile - 2 has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * 128u;
      for (int s = 0; s < nst; ++s, ++gs) {
        const uint32_t slot = gs % kFStages, phase = (gs / kFStages) & 1u;
        mbar_wait(full_bar + 8 * slot, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = tmem_base + kTmemAcc + slot * 32u;                     // 64 channels = 32 packed columns
          const uint32_t b0 = desc_lo(sm + kOffB + slot * kFBBytes, 128u * 16u);   // [8 chunks][128 rows][16 B]
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ts(tacc, a0 + ks * 8u, b0 + ((uint32_t)(ks * 2 * 128 * 16) >> 4), b_hi, idesc, (s | ks) ? 1u : 0u);
          umma_commit(empty_bar + 8 * slot);
          if (s == nst - 1) umma_commit(tfull_bar + 8 * buf);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kFGatherWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// launch with programmatic stream serialization: the grid may start while the previous kernel of the stream is still draining (the
// previous kernel executes griddepcontrol.launch_dependents); used ONLY between the two launches of one layer call, which are
// independent of each other.  The next layer's first launch is a normal one: it waits for everything before it.
template <typename... KArgs, typename... Args>
cudaError_t launch_overlapped(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, bool overlap, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = overlap ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace

extern "C" int mode_sphere_conv_pack_weights(const float* w, mode_h16* w_packed, int C, int Co, int fmt, void* stream) {
  MODE_CHECK_ARG(w && w_packed, "sphere_conv_pack_weights: null pointer");
  MODE_CHECK_ARG(C > 0 && C % 64 == 0 && Co >= 16 && Co <= 256 && Co % 16 == 0, "sphere_conv_pack_weights: need C %% 64 == 0 and Co %% 16 == 0, Co <= 256");
  const long long total = 9LL * C * Co;
  pack_wsphere_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, w_packed, C, Co, fmt, total);
  MODE_CHECK_LAUNCH("sphere_conv_pack_weights");
  return MODE_OK;
}

// tile geometry of the slab kernel for an H x W map: 16 x 8 tiles when the long (wrapping) axis is the rows (Cassini), 8 x 16 when
// it is the columns (ERP); 0 tile positions when the map is not tileable
static void slab_tiling(int H, int W, int& th, int& tw, int& npos) {
  th = H >= W ? 16 : 8, tw = H >= W ? 8 : 16;
  npos = (H % th == 0 && W % tw == 0) ? (H / th) * (W / tw) : 0;
}

extern "C" size_t mode_sphere_conv_table_bytes(int H, int W, int Kh, int Kw) {
  int th, tw, npos;
  slab_tiling(H, W, th, tw, npos);
  return (size_t)16 * Kh * Kw * H * W + 16 + (size_t)npos * (16 + 4 + 4);  // entries | header | tile info | fast list | rest list
}

extern "C" int mode_sphere_conv_build_table(const float* pos, void* table, int H, int W, int Kh, int Kw, int fmt, void* stream) {
  MODE_CHECK_ARG(pos && table && H > 0 && W > 0 && Kh > 0 && Kw > 0, "sphere_conv_build_table: bad arguments");
  MODE_CHECK_ARG(fmt == kFmtBF16 || fmt == kFmtFP16, "sphere_conv_build_table: fmt must be 0 (bf16) or 1 (fp16)");
  MODE_CHECK_ARG(H < 32768 && W < 32768, "sphere_conv_build_table: map too large");
  (void)fmt;  // the table no longer depends on the storage format (fp16 blend weights for both)
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = (long long)Kh * Kw * H * W;
  sphere_table_kernel<<<ceil_div(n, 256), 256, 0, s>>>(pos, (int4*)table, H, W, Kh * Kw);
  MODE_CHECK_LAUNCH("sphere_conv_build_table");
  int th, tw, npos;
  slab_tiling(H, W, th, tw, npos);
  int* hdr = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(table) + (size_t)16 * n);
  int4* info = reinterpret_cast<int4*>(hdr + 4);
  if (npos > 0 && Kh == 3 && Kw == 3) {
    sphere_tileinfo_kernel<<<npos, 128, 0, s>>>((const int4*)table, info, H, W, Kh * Kw, th, tw);
    MODE_CHECK_LAUNCH("sphere_conv_build_table (tile classes)");
  }
  sphere_tilelist_kernel<<<1, 32, 0, s>>>(info, hdr, hdr + 4 + 4 * npos, hdr + 4 + 5 * npos, (Kh == 3 && Kw == 3) ? npos : 0, th, tw);
  MODE_CHECK_LAUNCH("sphere_conv_build_table (tile lists)");
  return MODE_OK;
}

extern "C" int mode_sphere_conv_tc(const mode_h16* x, const void* table, const mode_h16* w_packed, const float* scale, const float* shift,
                                   const mode_h16* residual, mode_h16* out, int B, int C, int H, int W, int Co, int relu, int fmt, void* stream) {
  MODE_CHECK_ARG(fmt == kFmtBF16 || fmt == kFmtFP16, "sphere_conv_tc: fmt must be 0 (bf16) or 1 (fp16)");
  MODE_CHECK_ARG(x && table && w_packed && out, "sphere_conv_tc: null pointer");
  MODE_CHECK_ARG(B > 0 && H > 0 && W > 0, "sphere_conv_tc: bad shape");
  MODE_CHECK_ARG(C > 0 && C % 64 == 0, "sphere_conv_tc: C (%d) must be a multiple of 64 (use the f32 kernel otherwise)", C);
  MODE_CHECK_ARG(Co >= 64 && Co <= 256 && Co % 64 == 0, "sphere_conv_tc: Co (%d) must be 64, 128, 192 or 256", Co);
  cudaStream_t cs = (cudaStream_t)stream;
  ScParams p;
  p.x = x, p.table = (const int4*)table, p.wpk = w_packed, p.scale = scale, p.shift = shift, p.res = residual, p.out = out;
  p.B = B, p.C = C, p.H = H, p.W = W, p.Co = Co, p.relu = relu;
  p.npix = (long long)B * H * W;
  p.plist = nullptr, p.pcount = nullptr;
  MODE_CHECK_ARG((p.npix + W + 1) * C < 2147483647LL, "sphere_conv_tc: activation tensor too large for 32-bit offsets");
  const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  // ---- slab kernel for every tile whose neighbourhood fits shared memory (all but the polar tile columns on a 256 x 128 map)
  int th, tw, npos;
  slab_tiling(H, W, th, tw, npos);
  const int* hdr = reinterpret_cast<const int*>(reinterpret_cast<const uint8_t*>(table) + (size_t)16 * 9 * H * W);
  static const bool slab_disabled = [] {
    const char* e = getenv("MODE_B200_SPHERE_SLAB");  // diagnostics / A-B timing: MODE_B200_SPHERE_SLAB=0 forces the direct-gather kernel
    return e != nullptr && e[0] == '0';
  }();
  const bool slab_ok = !slab_disabled && npos > 0 && Co == 128 && (C == 64 || C == 128) && aligned && B < 65536;
  CUtensorMap tm_out, tm_res, tm_x;
  memset(&tm_out, 0, sizeof(tm_out));
  memset(&tm_res, 0, sizeof(tm_res));
  if (slab_ok) {
    int rc = make_epi_tmap(&tm_out, out, fmt, Co, W, (long long)B * H, tw);
    if (rc == MODE_OK && residual) rc = make_epi_tmap(&tm_res, residual, fmt, Co, W, (long long)B * H, tw);
    if (rc == MODE_OK) rc = make_slab_tmap(&tm_x, x, fmt, B, C, H, W, th > tw);
    if (rc != MODE_OK) return rc;
    FcParams f;
    f.table = (const int4*)table, f.hdr = hdr, f.wpk = w_packed, f.scale = scale, f.shift = shift;
    f.has_res = residual != nullptr, f.relu = relu, f.B = B, f.C = C, f.H = H, f.W = W, f.npos = npos;
    f.th = th, f.tw = tw, f.tw_shift = tw == 8 ? 3 : 4, f.cassini = th > tw;
    const size_t fsmem = 1024 + 2 * (size_t)kSlabBufBytes + (size_t)kFStages * kFBBytes + kFPieces * 2048 + (2 * kFStages + 8 + kFPieces) * 8 + 16;
    static thread_local size_t fattr_dev[kMaxDevices] = {};
    size_t& fattr = fattr_dev[current_device()];
    if (fsmem > fattr) {
      const void* kernels[4] = {(const void*)sphere_conv_slab_kernel<kFmtBF16, 0>, (const void*)sphere_conv_slab_kernel<kFmtBF16, 1>,
                                (const void*)sphere_conv_slab_kernel<kFmtFP16, 0>, (const void*)sphere_conv_slab_kernel<kFmtFP16, 1>};
      for (const void* k : kernels) MODE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem), "sphere_conv_tc (slab)");
      fattr = fsmem;
    }
    const int fgrid = std::min(B * npos, kNumSMs);
    if (fmt == kFmtBF16) {
      if (f.cassini)
        sphere_conv_slab_kernel<kFmtBF16, 1><<<fgrid, kFThreads, fsmem, cs>>>(f, tm_x, tm_out, tm_res);
      else
        sphere_conv_slab_kernel<kFmtBF16, 0><<<fgrid, kFThreads, fsmem, cs>>>(f, tm_x, tm_out, tm_res);
    } else {
      if (f.cassini)
        sphere_conv_slab_kernel<kFmtFP16, 1><<<fgrid, kFThreads, fsmem, cs>>>(f, tm_x, tm_out, tm_res);
      else
        sphere_conv_slab_kernel<kFmtFP16, 0><<<fgrid, kFThreads, fsmem, cs>>>(f, tm_x, tm_out, tm_res);
    }
    MODE_CHECK_LAUNCH("sphere_conv_tc (slab)");
    // the remaining tile positions go through the direct-gather kernel below, in list mode, with the same tile geometry
    p.tw = tw, p.th = th, p.tiles_x = W / tw, p.tiles_y = H / th;
    p.ntiles = B * npos;  // upper bound; the kernel reads the real count from the table
    p.plist = hdr + 4 + 5 * npos, p.pcount = hdr + 1;
  } else if (W % 16 == 0 && H % 8 == 0) {
    p.tw = 16, p.th = 8, p.tiles_x = W / 16, p.tiles_y = H / 8;
    p.ntiles = B * p.tiles_x * p.tiles_y;
  } else {
    p.tw = 0, p.th = 0, p.tiles_x = p.tiles_y = 0;
    p.ntiles = (int)((p.npix + 127) / 128);
  }
  const size_t smem = (size_t)kStagesS * (kABytes + (size_t)Co * 128) + 1024 + kGatherWarps * 2048 + (2 * kStagesS + 4 + kGatherWarps) * 8 + 16;
  static thread_local size_t attr_dev[kMaxDevices] = {};  // the attribute is per device and per function
  size_t& attr = attr_dev[current_device()];
  if (smem > attr) {
    // one CTA per SM: leave the rest of the 228 KB to L1 -- the 9 taps x 4 corners of a tile re-read the same ~45 KB
    // of input, which must stay L1 resident (with a maximal carve-out the kernel was L2-bandwidth bound: 3.6 GB/launch)
    const int carve = (int)((smem + 8 * 1024) * 100 / (228 * 1024)) + 1;
    const void* kernels[4] = {(const void*)sphere_conv_tc_kernel<kFmtBF16, 0>, (const void*)sphere_conv_tc_kernel<kFmtFP16, 0>,
                              (const void*)sphere_conv_tc_kernel<kFmtBF16, 128>, (const void*)sphere_conv_tc_kernel<kFmtFP16, 128>};
    for (const void* k : kernels) {
      MODE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "sphere_conv_tc");
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    attr = smem;
  }
  // epilogue tiles through TMA: (Co, W, B*H) view of the NHWC output / residual, box 32 ch x tw x 32/tw, 64-byte swizzle
  p.epi_tma = (p.tw != 0 && (Co == 64 || Co == 128) && aligned) ? 1 : 0;
  if (p.epi_tma && !slab_ok) {
    int rc = make_epi_tmap(&tm_out, out, fmt, Co, W, (long long)B * H, p.tw);
    if (rc == MODE_OK && residual) rc = make_epi_tmap(&tm_res, residual, fmt, Co, W, (long long)B * H, p.tw);
    if (rc != MODE_OK) return rc;
  }
  const int grid = std::min(p.ntiles, kNumSMs);
  cudaError_t le;
  if (C == 128) {
    if (fmt == kFmtBF16)
      le = launch_overlapped(sphere_conv_tc_kernel<kFmtBF16, 128>, grid, kThreadsS, smem, cs, slab_ok, p, tm_out, tm_res);
    else
      le = launch_overlapped(sphere_conv_tc_kernel<kFmtFP16, 128>, grid, kThreadsS, smem, cs, slab_ok, p, tm_out, tm_res);
  } else {
    if (fmt == kFmtBF16)
      le = launch_overlapped(sphere_conv_tc_kernel<kFmtBF16, 0>, grid, kThreadsS, smem, cs, slab_ok, p, tm_out, tm_res);
    else
      le = launch_overlapped(sphere_conv_tc_kernel<kFmtFP16, 0>, grid, kThreadsS, smem, cs, slab_ok, p, tm_out, tm_res);
  }
  MODE_CHECK_CUDA(le, "sphere_conv_tc");
  MODE_CHECK_LAUNCH("sphere_conv_tc");
  return MODE_OK;
}
