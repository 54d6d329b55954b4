// conv3d_cls_tc.cu -- the 32 -> 1 classifier convolution of the 3-D stack (3x3x3, pad 1, no bias, fp32 logits + fp32 residual
// chain) as a POINTWISE tensor-core GEMM followed by a shifted sum.
//
// Reference: classif1/2/3[2] = nn.Conv3d(32, 1, 3, padding=1, bias=False) and cost2 = classif2(out2) + cost1,
// cost3 = classif3(out3) + cost2 (models/mode_disparity.py:72-80, 126-129).
//
// An implicit GEMM (conv3d_tc with N = 16 padded output channels x 3 stacked depth taps) spends 18 MMAs per 128 output voxels
// on a layer that has ONE real output channel: 0.29 ms per 6 pairs against a 0.10 ms HBM floor.  With a single output
// channel the 27 taps can be applied BEFORE the spatial shift:
//     T[v][tap] = sum_c x[v][c] * w[c][tap]                 one GEMM per input voxel, M = voxels, N = 27 (padded to 32), K = 32
//     out[d,h,w] = sum_{kd,kh,kw} T[(d-1+kd, h-1+kh, w-1+kw)][(kd,kh,kw)]
// i.e. 4 MMAs (2 M tiles x 2 K steps, N = 32) per halo'd 18 x 10 input plane instead of 18.  The depth part of the sum involves no
// spatial shift -- S(od)[v][kh,kw] = T(od-1)[v][0,kh,kw] + T(od)[v][1,kh,kw] + T(od+1)[v][2,kh,kw] for the SAME input voxel v -- so
// the 4 EXTRACT warps (TMEM lane = voxel) carry it in registers across consecutive planes and spill only the 9 completed values per
// voxel into a shared-memory ring ([kh,kw][voxel], conflict-free); the 4 SUM warps add the 9 spatially shifted values per output
// voxel (+ fp32 residual, coalesced fp32 stores).  A CTA walks a 16 x 8 tile column along depth: TMA box (zero-filled halo) ->
// 4 MMAs -> T in TMEM -> extract -> sum.  (First version: all 27 taps went through shared memory -- 27 STS + 27 LDS per voxel and
// plane; the kernel sat at 0.49 of the HBM bound with the shared-memory port 76 % busy.)
#include <cuda.h>
#include <string.h>

#include "common.cuh"
using namespace mode;

namespace {

constexpr int kXW = 6, kSW = 4;                    // extract warps (0-3: M tile 0, 4-5: rows 128..191 of M tile 1; TMEM lane quadrant = warp % 4), sum warps
constexpr int kClsThreads = (kXW + kSW + 2) * 32;  // + MMA warp, TMA producer
constexpr int kBoxH = 18, kBoxW = 10, kBoxVox = kBoxH * kBoxW;  // halo'd input plane of a 16 x 8 tile
constexpr int kSlotBytes = 12288;      // one box (11520 B) rounded up to the swizzle-atom alignment; the second M tile reads on
                                       // into whatever follows (rows >= 180 are never used)
constexpr int kSlots = 5;
constexpr int kTStride = 184;          // floats per (kh,kw) row of an S plane (180 voxels padded)
constexpr int kSPlane = 9 * kTStride;  // floats per S plane: the depth-summed taps of one OUTPUT plane, still un-shifted in (h, w)
constexpr int kSRing = 4;              // S planes in shared memory
constexpr int kTmemRing = 4;           // input planes whose T sits in TMEM (64 columns each)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {  // bounded: a protocol bug must trap, never hang the box
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    if (it > (1u << 24)) {
      printf("conv3d_cls_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
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
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); }
// descriptor version 1 at bit 46 (bit 14 of the high word), layout type at bits 61-63 (0 = none, 4 = 64-byte swizzle)
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo, uint32_t layout) { return ((sbo >> 4) & 0x3FFF) | (1u << 14) | (layout << 29); }
__host__ __device__ constexpr uint32_t make_idesc(int n, int fmt) {
  return (1u << 4) | ((fmt == 0 ? 1u : 0u) << 7) | ((fmt == 0 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
}

struct ClsParams {
  const float* w;    // (1, 32, 3, 3, 3) fp32
  const float* res;  // (B, D, H, W) fp32 or null
  float* out;        // (B, D, H, W) fp32
  int B, D, H, W;
  int tiles_h, tiles_w, nchunks, chunk, nitems;
};

template <int FMT>
__global__ void __launch_bounds__(kClsThreads, 2) conv3d_cls_tc_kernel(const ClsParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((base + 1023u) & ~1023u) - base);  // swizzled TMA boxes: 1024-byte aligned
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  uint8_t* slots_s = smem;
  uint8_t* b_s = slots_s + kSlots * kSlotBytes;       // weights [K chunk 4][n 32][8 x 16 bit] = 2 KB
  float* s_s = reinterpret_cast<float*>(b_s + 2048);  // [kSRing planes][9 (kh,kw)][kTStride]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_s + kSRing * kSPlane);
  uint64_t* full_bar = bars;                     // [kSlots]    TMA -> MMA
  uint64_t* empty_bar = bars + kSlots;           // [kSlots]    MMA (commit) -> TMA
  uint64_t* tfull_bar = bars + 2 * kSlots;       // [kTmemRing] MMA (commit) -> extract warps
  uint64_t* tempty_bar = tfull_bar + kTmemRing;  // [kTmemRing] extract warps -> MMA
  uint64_t* sready_bar = tempty_bar + kTmemRing; // [kSRing] extract warps -> sum warps: S plane in shared memory
  uint64_t* sfree_bar = sready_bar + kSRing;     // [kSRing] sum warps -> extract warps: ring slot may be overwritten
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(sfree_bar + kSRing);

  // weights: B[n = tap][k = channel], K-major un-swizzled core matrices [k chunk][n][8]
  for (int e = threadIdx.x; e < 4 * 32 * 8; e += kClsThreads) {
    const int el = e & 7, n = (e >> 3) & 31, kc = e >> 8;
    const int c = kc * 8 + el;
    reinterpret_cast<uint16_t*>(b_s)[e] = float_to_h16_bits(n < 27 ? p.w[c * 27 + n] : 0.f, FMT);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(smem_u32(full_bar + i), 1);
      mbar_init(smem_u32(empty_bar + i), 1);
    }
    for (int i = 0; i < kTmemRing; ++i) {
      mbar_init(smem_u32(tfull_bar + i), 1);
      mbar_init(smem_u32(tempty_bar + i), kXW);
    }
    for (int i = 0; i < kSRing; ++i) {
      mbar_init(smem_u32(sready_bar + i), kXW);
      mbar_init(smem_u32(sfree_bar + i), kSW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kXW + kSW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(kTmemRing * 64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  // item = (batch, tile, depth chunk [o0, o1)); its input planes are [max(o0-1, 0), min(o1, D-1)]
  auto decode = [&](int item, int& b, int& th, int& tw, int& o0, int& o1) {
    const int ch = item % p.nchunks;
    item /= p.nchunks;
    tw = item % p.tiles_w;
    item /= p.tiles_w;
    th = item % p.tiles_h;
    b = item / p.tiles_h;
    o0 = ch * p.chunk, o1 = min(o0 + p.chunk, p.D);
  };

  if (warp == kXW + kSW + 1) {
    // =========================================================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
      uint32_t n = 0;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        int b, th, tw, o0, o1;
        decode(item, b, th, tw, o0, o1);
        for (int pl = max(o0 - 1, 0); pl <= min(o1, p.D - 1); ++pl, ++n) {
          const uint32_t slot = n % kSlots, phase = (n / kSlots) & 1;
          mbar_wait(smem_u32(empty_bar + slot), phase ^ 1);
          const uint32_t bar = smem_u32(full_bar + slot);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBoxVox * 64) : "memory");
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                           smem_u32(slots_s + (size_t)slot * kSlotBytes)),
                       "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(tw * 8 - 1), "r"(th * 16 - 1), "r"(b * p.D + pl), "r"(bar)
                       : "memory");
        }
      }
    }
  } else if (warp == kXW + kSW) {
    // =========================================================== MMA issuer: 4 MMAs per input plane
    const uint32_t idesc = make_idesc(32, FMT);
    const uint32_t a_hi = desc_hi(512, 4), b_hi = desc_hi(128, 0);
    const uint32_t b_lo = desc_lo(smem_u32(b_s), 512);
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      int b, th, tw, o0, o1;
      decode(item, b, th, tw, o0, o1);
      for (int pl = max(o0 - 1, 0); pl <= min(o1, p.D - 1); ++pl, ++n) {
        const uint32_t slot = n % kSlots, phase = (n / kSlots) & 1;
        const uint32_t ts = n % kTmemRing, tphase = (n / kTmemRing) & 1;
        mbar_wait(smem_u32(tempty_bar + ts), tphase ^ 1);
        mbar_wait(smem_u32(full_bar + slot), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = smem_u32(slots_s + (size_t)slot * kSlotBytes);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma(tmem_base + ts * 64 + mt * 32, desc_lo(a0 + mt * 8192 + ks * 32, 16), a_hi, b_lo + ((uint32_t)(ks * 2 * 512) >> 4), b_hi, idesc, ks ? 1u : 0u);
          umma_commit(smem_u32(empty_bar + slot));
          umma_commit(smem_u32(tfull_bar + ts));
        }
        __syncwarp();
      }
    }
  } else if (warp < kXW) {
    // =========================================================== extract: T of every input plane out of TMEM; depth sum in registers
    // P1 = the kd = 0 taps of the previous plane (partial S of the current plane), P0 = kd 0 of two planes ago + kd 1 of the previous
    // one (partial S of the previous plane); the current plane's kd = 2 taps complete S(ip - 1).  Tap index = kd * 9 + kh * 3 + kw.
    // Warps 0-3 own the 128 rows of the first M tile, warps 4-5 rows 128..191 of the second (voxels 128..179 are real): one TMEM load
    // per warp and plane.
    uint32_t n = 0, m = 0;  // input planes consumed / output planes produced by this CTA so far
    const int vox = warp * 32 + lane;  // warps 4, 5: 128 + (warp - 4) * 32 + lane
    const bool vox_ok = vox < kBoxVox;
    const uint32_t tsub = ((uint32_t)((warp & 3) * 32) << 16) + (warp >= 4 ? 32u : 0u);  // lane quadrant, column block of the M tile
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      int b, th, tw, o0, o1;
      decode(item, b, th, tw, o0, o1);
      float P0[9], P1[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) P0[j] = P1[j] = 0.f;
      for (int pl = max(o0 - 1, 0); pl <= min(o1, p.D - 1); ++pl, ++n) {
        const uint32_t ts = n % kTmemRing, tphase = (n / kTmemRing) & 1;
        const bool emit = pl - 1 >= o0;  // S(pl - 1) is an output plane of this item (pl - 1 < o1 always holds)
        float* sp = s_s + (size_t)(m % kSRing) * kSPlane;
        mbar_wait(smem_u32(tfull_bar + ts), tphase);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + tsub + ts * 64, v);
        if (emit) mbar_wait(smem_u32(sfree_bar + m % kSRing), ((m / kSRing) & 1) ^ 1);  // the sum warps are done with the plane that lived here
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(tempty_bar + ts));
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          if (emit && vox_ok) sp[j * kTStride + vox] = P0[j] + __uint_as_float(v[18 + j]);
          P0[j] = P1[j] + __uint_as_float(v[9 + j]);
          P1[j] = __uint_as_float(v[j]);
        }
        if (emit) {
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(sready_bar + m % kSRing));  // (mbarrier arrive has release semantics for the stores above)
          ++m;
        }
        if (pl == p.D - 1) {
          // last plane of the volume: S(D - 1) has no kd = 2 term (zero padding along depth) -- it is complete now
          float* sq = s_s + (size_t)(m % kSRing) * kSPlane;
          mbar_wait(smem_u32(sfree_bar + m % kSRing), ((m / kSRing) & 1) ^ 1);
          if (vox_ok) {
#pragma unroll
            for (int j = 0; j < 9; ++j) sq[j * kTStride + vox] = P0[j];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(sready_bar + m % kSRing));
          ++m;
        }
      }
    }
  } else {
    // =========================================================== sum: out[od][h][w] = sum over (kh, kw) of S(od)[(h-1+kh, w-1+kw)][kh,kw]
    // output voxel of the tile: sum warp s takes tile rows s, s+4, s+8, s+12 (8 columns each).  An S plane keeps the 18 x 10 halo'd
    // box with a row pitch of 10 floats: rows 4 apart start 40 floats = 8 banks apart, so the warp's four 8-wide windows fall on 32
    // distinct banks for every tap (four CONSECUTIVE rows overlap on 6 banks: every LDS a 2-way conflict).
    const int hl = (warp - kXW) + 4 * (lane >> 3), wl = lane & 7;
    uint32_t m = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      int b, th, tw, o0, o1;
      decode(item, b, th, tw, o0, o1);
      const int oh = th * 16 + hl, ow = tw * 8 + wl;
      const bool ok = oh < p.H && ow < p.W;
      // The fp32 residual of a plane is fetched TWO planes ahead: a plane's sum is a few hundred cycles of work, a DRAM round trip
      // several times that.
      const size_t plane = (size_t)p.H * p.W;
      const size_t o_first = (((size_t)b * p.D + o0) * p.H + oh) * p.W + ow;
      const bool has_r = ok && p.res != nullptr;
      float r0 = has_r ? __ldg(p.res + o_first) : 0.f;
      float r1 = (has_r && o0 + 1 < o1) ? __ldg(p.res + o_first + plane) : 0.f;
      for (int od = o0; od < o1; ++od, ++m) {
        const size_t o = o_first + (size_t)(od - o0) * plane;
        const float r = r0;
        r0 = r1;
        r1 = (has_r && od + 2 < o1) ? __ldg(p.res + o + 2 * plane) : 0.f;
        const uint32_t rs = m % kSRing;
        mbar_wait(smem_u32(sready_bar + rs), (m / kSRing) & 1);
        const float* sq = s_s + (size_t)rs * kSPlane + hl * kBoxW + wl;
        float acc[3];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          acc[kh] = sq[(kh * 3) * kTStride + kh * kBoxW];
#pragma unroll
          for (int kw = 1; kw < 3; ++kw) acc[kh] += sq[(kh * 3 + kw) * kTStride + kh * kBoxW + kw];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(sfree_bar + rs));
        if (ok) p.out[o] = (acc[0] + acc[1]) + acc[2] + r;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kXW + kSW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemRing * 64) : "memory");
}

}  // namespace

extern "C" int mode_conv3d_classifier_tc(const mode_h16* x, const float* w, const float* residual_f32, float* out_f32, int B, int D, int H, int W, int fmt,
                                         void* stream) {
  MODE_CHECK_ARG(fmt == kFmtBF16 || fmt == kFmtFP16, "conv3d_classifier_tc: fmt must be 0 (bf16) or 1 (fp16)");
  MODE_CHECK_ARG(x && w && out_f32, "conv3d_classifier_tc: null pointer");
  MODE_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0, "conv3d_classifier_tc: bad shape");
  static decltype(&cuTensorMapEncodeTiled) encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      set_error("conv3d_classifier_tc: cuTensorMapEncodeTiled is not available from this driver");
      return MODE_ECUDA;
    }
    encode = reinterpret_cast<decltype(&cuTensorMapEncodeTiled)>(fn);
  }
  CUtensorMap tm;
  {
    const cuuint64_t gdim[4] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * D};
    const cuuint64_t gstr[3] = {64, (cuuint64_t)W * 64, (cuuint64_t)H * W * 64};
    const cuuint32_t box[4] = {32, kBoxW, kBoxH, 1}, estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(&tm, fmt == kFmtBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<mode_h16*>(x), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv3d_classifier_tc: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
      return MODE_ECUDA;
    }
  }
  ClsParams p;
  p.w = w, p.res = residual_f32, p.out = out_f32;
  p.B = B, p.D = D, p.H = H, p.W = W;
  p.tiles_h = ceil_div(H, 16), p.tiles_w = ceil_div(W, 8);
  const int cols = B * p.tiles_h * p.tiles_w;
  const int ctas = 2 * kNumSMs;
  // depth chunk: minimise (items per CTA) x (planes per item incl. the two halo planes)
  long long best = -1;
  int best_chunk = D;
  for (int nch = 1; nch <= D; ++nch) {
    const int chunk = ceil_div(D, nch);
    if (ceil_div(D, chunk) != nch) continue;
    const long long cost = (long long)ceil_div(cols * nch, ctas) * (chunk + 2);
    if (best < 0 || cost < best) best = cost, best_chunk = chunk;
  }
  p.chunk = best_chunk, p.nchunks = ceil_div(D, best_chunk);
  p.nitems = cols * p.nchunks;
  const size_t smem = 1024 + (size_t)kSlots * kSlotBytes + 2048 + (size_t)kSRing * kSPlane * 4 + (2 * kSlots + 2 * kTmemRing + 2 * kSRing) * 8 + 16;
  static thread_local bool attr_dev[kMaxDevices] = {};  // the attribute is per device and per function
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    MODE_CHECK_CUDA(cudaFuncSetAttribute(conv3d_cls_tc_kernel<kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d_classifier_tc");
    MODE_CHECK_CUDA(cudaFuncSetAttribute(conv3d_cls_tc_kernel<kFmtFP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d_classifier_tc");
    attr = true;
  }
  const int grid = std::min(p.nitems, ctas);
  if (fmt == kFmtBF16)
    conv3d_cls_tc_kernel<kFmtBF16><<<grid, kClsThreads, smem, (cudaStream_t)stream>>>(p, tm);
  else
    conv3d_cls_tc_kernel<kFmtFP16><<<grid, kClsThreads, smem, (cudaStream_t)stream>>>(p, tm);
  MODE_CHECK_LAUNCH("conv3d_classifier_tc");
  return MODE_OK;
}
