// stem_conv_tc.cu -- first convolution of the MODE feature extractor on tcgen05 tensor cores:
// 3 -> 32 channels, 7x7, stride 2, pad 3, + folded eval-BatchNorm + ReLU.
//
// Reference: sphere_feature_extraction.firstconv[0] = convbn(3, 32, 7, 2, 3, 1) + ReLU (models/submodule.py:155, 15-17),
// applied to the left and the right image (models/mode_disparity.py:99-100).
//
// Reads the fp32 NCHW camera images as the caller holds them (left and right batches given separately: no torch.cat,
// no dtype / layout conversion pass) and writes the NHWC 16-bit activation the rest of the 16-bit plan consumes.
// 4704 MACs per output pixel would keep the FP32 pipe busy for ~200 us per 12 images at 100 % FFMA issue, and a
// library implicit GEMM pads K = 147 and C = 3 up to tensor-core tiles (0.5 ms + two conversion passes, measured).
//
// Formulation: an image row is kept in shared memory as 16-bit pixels of 4 channels (c0, c1, c2, 0) = 8 bytes, with a
// 3-pixel zero halo.  For output row oy, kernel row ky and a group g of 4 kernel columns (kx = 4g .. 4g+3), the A operand
// of one M=128 x N=32 x K=16 MMA is a pure VIEW of that row: GEMM row m = output column ox0 + m starts at pixel
// 2*(ox0+m) + 4g, i.e. 16 bytes further per m -- which is exactly the row pitch of an un-swizzled K-major core matrix
// (8 rows x 16 B, rows 16 B apart): the stride-2 of the convolution is absorbed by the 2-pixel = 16-byte row pitch.
// The second 16-byte K chunk (LBO = 16 B) is the next pixel pair, so one MMA covers 4 kernel columns x 4 channels, and
// 7 rows x 2 column groups = 14 MMAs produce 128 output pixels x 32 channels (the 8th column and the 4th channel carry
// zero weights).  No im2col, no data movement besides the fp32 -> 16-bit row conversion (each input row converted once).
//
// CTA = one strip of output rows of one image.  16 warps convert input rows into a ring of 24 row buffers and drain the
// accumulators (BN affine + ReLU + 16-bit NHWC stores); one warp issues the MMAs of a group of G output rows (G x tiles x
// 14) into one of two TMEM buffers while the other is being drained.
#include "common.cuh"
using namespace mode;

namespace {

constexpr int kWorkWarps = 16;
constexpr int kStemThreads = (kWorkWarps + 1) * 32;  // + MMA warp
constexpr int kRing = 24;                            // input row buffers (needs 4G + 5 <= kRing)
constexpr int kBBytes = 7 * 2 * 2 * 32 * 16;         // weights [ky][col group][K chunk][n][8 x 16 bit] = 14336 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {  // bounded: a protocol bug must trap, never hang the box
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    if (it > (1u << 24)) {
      printf("stem_conv_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
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
// un-swizzled K-major descriptor halves: start address + LBO (distance between the two 16-byte K chunks) | SBO (distance
// between 8-row groups), descriptor version 1
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFF) | (1u << 14); }
// instruction descriptor: D fp32, A/B bf16 (fmt 0) or fp16 (fmt 1), K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n, int fmt) {
  return (1u << 4) | ((fmt == 0 ? 1u : 0u) << 7) | ((fmt == 0 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
}

struct StemParams {
  const float* x0;  // (B0, 3, H, W) fp32
  const float* x1;  // (B1, 3, H, W) fp32 or null
  const float* w;   // (32, 3, 7, 7) fp32
  const float* scale;
  const float* shift;
  uint16_t* out;  // (B0 + B1, Ho, Wo, 32) 16-bit
  int B0, B, H, W, Ho, Wo, relu;
  int tiles_x;     // ceil(Wo / 128)
  int G;           // output rows per MMA group: G * tiles_x * 32 TMEM columns per buffer
  int rows_strip;  // output rows per CTA (multiple of G)
  int strips;      // strips per image
  int row_bytes;   // one shared-memory row: (256 * tiles_x + 8) pixels x 8 B
};

template <int FMT>
__global__ void __launch_bounds__(kStemThreads, 1) stem_conv_tc_kernel(const StemParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  uint8_t* b_s = smem;                                                 // weights
  float4* ss_s = reinterpret_cast<float4*>(smem + kBBytes);            // {scale x4, shift x4} x 8
  uint64_t* tfull_bar = reinterpret_cast<uint64_t*>(smem + kBBytes + 256);  // [2]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tfull_bar + 2);
  uint8_t* ring_s = smem + kBBytes + 256 + 128;
  const int buf_cols = p.G * p.tiles_x * 32;
  const uint32_t tmem_cols = buf_cols <= 16 ? 32 : buf_cols <= 32 ? 64 : buf_cols <= 64 ? 128 : buf_cols <= 128 ? 256 : 512;

  // ---- one-time setup: weights -> 16-bit MMA layout, BN affine table, zeroed row ring (halo pixels stay zero for good)
  for (int e = threadIdx.x; e < kBBytes / 2; e += kStemThreads) {
    const int el = e & 7, n = (e >> 3) & 31, kc = (e >> 8) & 1, g = (e >> 9) & 1, ky = e >> 10;
    const int kx = 4 * g + 2 * kc + (el >> 2), c = el & 3;
    const float v = (c < 3 && kx < 7) ? p.w[((n * 3 + c) * 7 + ky) * 7 + kx] : 0.f;
    reinterpret_cast<uint16_t*>(b_s)[e] = float_to_h16_bits(v, FMT);
  }
  if (threadIdx.x < 32) {
    const int c = threadIdx.x;
    float* e = reinterpret_cast<float*>(ss_s + 2 * (c >> 2)) + (c & 3);
    e[0] = p.scale ? p.scale[c] : 1.f;
    e[4] = p.shift ? p.shift[c] : 0.f;
  }
  for (int e = threadIdx.x; e < kRing * p.row_bytes / 16; e += kStemThreads) reinterpret_cast<uint4*>(ring_s)[e] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(tfull_bar), 1);
    mbar_init(smem_u32(tfull_bar + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWorkWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int b = blockIdx.x / p.strips, strip = blockIdx.x - b * p.strips;
  const int r0 = strip * p.rows_strip, r1 = min(r0 + p.rows_strip, p.Ho);
  const float* xb = b < p.B0 ? p.x0 + (size_t)b * 3 * p.H * p.W : p.x1 + (size_t)(b - p.B0) * 3 * p.H * p.W;
  const size_t plane = (size_t)p.H * p.W;
  const int ngroups = (r1 - r0 + p.G - 1) / p.G;
  const uint32_t idesc = make_idesc(32, FMT);
  const uint32_t a_hi = desc_hi(128), b_hi = desc_hi(128);

  // input row iy lives in ring slot (iy + 3) % kRing
  auto convert_rows = [&](int iy_lo, int iy_hi) {  // work warps: fp32 planar -> (c0, c1, c2, 0) 16-bit pixels, rows [iy_lo, iy_hi]
    // four rows per pass: their 12 loads per pixel column are all issued before the first conversion (one row at a time left 3 loads in
    // flight per thread, and the kernel waited on DRAM latency once per row: 0.18 of the HBM bound)
    for (int iy4 = iy_lo; iy4 <= iy_hi; iy4 += 4) {
      for (int ix = threadIdx.x; ix < p.W; ix += kWorkWarps * 32) {
        float c[4][3];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int iy = iy4 + r;
          const bool inside = iy <= iy_hi && iy >= 0 && iy < p.H;
          const float* src = xb + (size_t)(inside ? iy : 0) * p.W + ix;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) c[r][ch] = inside ? __ldg(src + ch * plane) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int iy = iy4 + r;
          if (iy <= iy_hi) {
            uint8_t* row = ring_s + (size_t)((iy + 3) % kRing) * p.row_bytes + 3 * 8;
            *reinterpret_cast<uint2*>(row + (size_t)ix * 8) = make_uint2(pack2<FMT>(c[r][0], c[r][1]), pack2<FMT>(c[r][2], 0.f));
          }
        }
      }
    }
  };
  auto drain_group = [&](int g) {  // work warps: accumulators of group g -> BN affine + ReLU -> 16-bit NHWC
    const uint32_t buf = (uint32_t)g & 1;
    mbar_wait(smem_u32(tfull_bar + buf), ((uint32_t)g >> 1) & 1);
    tc_fence_after();
    const int q = warp & 3;
    const int ntile = p.G * p.tiles_x;
    for (int t = warp >> 2; t < ntile; t += kWorkWarps / 4) {
      const int ry = t / p.tiles_x, tx = t - ry * p.tiles_x;
      const int oy = r0 + g * p.G + ry, ox = tx * 128 + q * 32 + lane;
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)buf_cols + (uint32_t)t * 32, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (oy < r1 && ox < p.Wo) {
        uint4* op = reinterpret_cast<uint4*>(p.out + (((size_t)b * p.Ho + oy) * p.Wo + ox) * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float y[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 sc = ss_s[2 * (2 * j + h)], sh = ss_s[2 * (2 * j + h) + 1];
            ffma2(y[4 * h], y[4 * h + 1], __uint_as_float(v[8 * j + 4 * h]), __uint_as_float(v[8 * j + 4 * h + 1]), sc.x, sc.y, sh.x, sh.y);
            ffma2(y[4 * h + 2], y[4 * h + 3], __uint_as_float(v[8 * j + 4 * h + 2]), __uint_as_float(v[8 * j + 4 * h + 3]), sc.z, sc.w, sh.z, sh.w);
          }
          uint4 o = make_uint4(pack2<FMT>(y[0], y[1]), pack2<FMT>(y[2], y[3]), pack2<FMT>(y[4], y[5]), pack2<FMT>(y[6], y[7]));
          if (p.relu) o.x = relu2<FMT>(o.x), o.y = relu2<FMT>(o.y), o.z = relu2<FMT>(o.z), o.w = relu2<FMT>(o.w);
          op[j] = o;
        }
      }
    }
    tc_fence_before();
  };

  for (int g = 0; g <= ngroups; ++g) {
    if (warp < kWorkWarps && g < ngroups) {
      // rows needed by group g: 2*(r0 + g*G) - 3 .. 2*(r0 + g*G + G - 1) + 3; all but the last 2G were converted for group g-1
      const int oy0 = r0 + g * p.G;
      const int lo = 2 * oy0 - 3, hi = 2 * (oy0 + p.G - 1) + 3;
      convert_rows(g == 0 ? lo : hi - 2 * p.G + 1, hi);
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();  // rows of group g are in place; every warp has finished draining group g-2 (frees TMEM buffer g & 1)
    tc_fence_after();
    if (warp == kWorkWarps) {
      if (g < ngroups) {
        if (elect_one()) {
          const int oy0 = r0 + g * p.G;
          const uint32_t buf = (uint32_t)g & 1;
          for (int ry = 0; ry < p.G; ++ry) {
            for (int tx = 0; tx < p.tiles_x; ++tx) {
              const uint32_t d = tmem_base + buf * (uint32_t)buf_cols + (uint32_t)(ry * p.tiles_x + tx) * 32;
#pragma unroll
              for (int ky = 0; ky < 7; ++ky) {
                const int iy = 2 * (oy0 + ry) - 3 + ky;
                const uint32_t row = smem_u32(ring_s + (size_t)((iy + 3) % kRing) * p.row_bytes);
#pragma unroll
                for (int cg = 0; cg < 2; ++cg) {
                  const uint32_t a_lo = desc_lo(row + (uint32_t)(256 * tx + 4 * cg) * 8, 16);
                  const uint32_t b_lo = desc_lo(smem_u32(b_s) + (uint32_t)((ky * 2 + cg) * 1024), 512);
                  umma(d, a_lo, a_hi, b_lo, b_hi, idesc, (ky | cg) ? 1u : 0u);
                }
              }
            }
          }
          umma_commit(smem_u32(tfull_bar + buf));
        }
        __syncwarp();
      }
    } else if (g > 0) {
      drain_group(g - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWorkWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

}  // namespace

extern "C" int mode_stem_conv_tc(const float* x0, const float* x1, const float* w, const float* scale, const float* shift, mode_h16* out, int B0, int B1, int H,
                                 int W, int relu, int fmt, void* stream) {
  MODE_CHECK_ARG(fmt == kFmtBF16 || fmt == kFmtFP16, "stem_conv_tc: fmt must be 0 (bf16) or 1 (fp16)");
  MODE_CHECK_ARG(x0 && w && out && (B1 == 0 || x1), "stem_conv_tc: null pointer");
  MODE_CHECK_ARG(B0 > 0 && B1 >= 0 && H > 0 && W > 0, "stem_conv_tc: bad shape");
  StemParams p;
  p.x0 = x0, p.x1 = x1, p.w = w, p.scale = scale, p.shift = shift, p.out = out;
  p.B0 = B0, p.B = B0 + B1, p.H = H, p.W = W, p.relu = relu;
  p.Ho = (H + 6 - 7) / 2 + 1, p.Wo = (W + 6 - 7) / 2 + 1;
  p.tiles_x = (p.Wo + 127) / 128;
  MODE_CHECK_ARG(p.tiles_x <= 8, "stem_conv_tc: image wider than 2048 pixels is not supported (W = %d)", W);
  p.G = p.tiles_x <= 2 ? 4 : p.tiles_x <= 4 ? 2 : 1;
  p.row_bytes = (256 * p.tiles_x + 8) * 8;
  const size_t smem = (size_t)kBBytes + 256 + 128 + (size_t)kRing * p.row_bytes;
  MODE_CHECK_ARG(smem <= 227 * 1024, "stem_conv_tc: image too wide for the shared-memory row ring (W = %d)", W);
  // strips: about one CTA per SM; a strip start re-converts 5 + 2G - 1 rows of overlap, so strips should not be tiny
  const int want = std::max(1, kNumSMs / p.B);
  int rows = (p.Ho + want - 1) / want;
  rows = std::max(p.G, (rows + p.G - 1) / p.G * p.G);
  p.rows_strip = rows;
  p.strips = (p.Ho + rows - 1) / rows;
  static thread_local size_t attr_dev[kMaxDevices] = {};  // the attribute is per device and per function
  size_t& attr = attr_dev[current_device()];
  if (smem > attr) {
    MODE_CHECK_CUDA(cudaFuncSetAttribute(stem_conv_tc_kernel<kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "stem_conv_tc");
    MODE_CHECK_CUDA(cudaFuncSetAttribute(stem_conv_tc_kernel<kFmtFP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "stem_conv_tc");
    attr = smem;
  }
  const int grid = p.B * p.strips;
  if (fmt == kFmtBF16)
    stem_conv_tc_kernel<kFmtBF16><<<grid, kStemThreads, smem, (cudaStream_t)stream>>>(p);
  else
    stem_conv_tc_kernel<kFmtFP16><<<grid, kStemThreads, smem, (cudaStream_t)stream>>>(p);
  MODE_CHECK_LAUNCH("stem_conv_tc");
  return MODE_OK;
}
