// conv3d_tc.cu -- 3x3x3 conv / strided conv / transposed conv on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Reference: nn.Conv3d / nn.ConvTranspose3d + BatchNorm3d(eval) [+ residual] [+ ReLU] in the stacked hourglass
// (models/submodule.py:20-22, models/mode_disparity.py:11-46,66-80,115-129) -- 1013.8 GFLOP per pair, the only
// dense contraction on the hot path.
//
// Formulation: implicit GEMM, M = 128 output positions (a 16(h) x 8(w) tile of one depth plane), N = NT output
// channels, K = 27 taps x Cin.  Activations are NDHWC bf16; accumulation is fp32 in TMEM; BN affine, residual and
// ReLU are applied in the epilogue straight out of TMEM.
//
// Data movement (why it is not a plain im2col GEMM): with N = 32 the MMA is shared-memory-read bound and a per-tap
// re-fetch of A from L2 would exceed the ~42 B/clk/SM L2 budget by 2.4x, so every input byte is brought on chip
// ONCE per tile column and reused by all 27 taps:
//   * a CTA walks a column of tiles along depth.  Each *input* plane (with its h/w halo) is staged once in shared
//     memory and immediately consumed by every output plane it contributes to ("plane-major" order): plane p feeds
//     outputs p-1, p, p+1 (stride 1), so three TMEM accumulators are in flight and each stage is used exactly once.
//   * the halo tile is stored un-swizzled as [8-channel chunk][h][w][16 B] (UMMA "interleave" K-major layout: a core
//     matrix = 8 consecutive w positions x 16 B).  A tap shift (kh, kw) is then just a different descriptor start
//     address -- no data is moved or duplicated for the 9 in-plane taps.
//   * stride-2 convs de-interleave the halo by (h, w) parity while staging so that the strided rows become
//     contiguous again; transposed convs keep 4 accumulators per output plane (one per output parity class).
//   * all 27 x Cin x NT weights stay resident in shared memory for the CTA's lifetime.
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lane quadrant = warp id), warp 4 = MMA issuer (one elected lane),
// warp 5 = TMA producer: one cp.async.bulk.tensor (4-D box {8 ch, w, h, 1 plane}) per 8-channel chunk lands each halo
// plane directly in the UMMA layout; out-of-bounds box elements are zero-filled by the TMA unit = the conv padding.
// Pipelines: smem ring full/empty (producer <-> MMA), TMEM set full/empty (MMA <-> epilogue).
#include <cuda.h>
#include <string.h>

#include "common.cuh"
using namespace mode;

namespace {

// warp roles: [0, EW) epilogue (TMEM lane quadrant = warp & 3), warp EW = MMA issuer, warp EW+1 = TMA producer.
// Transposed convs write 8x more output voxels than they read and add a same-sized residual: their epilogue is the
// bottleneck (per-CTA timers: the MMA warp waited 80 % of the time for TMEM blocks), so they get 16 epilogue warps
// (one per lane quadrant x output parity class) instead of 4.
template <int MODE>
constexpr int epi_warps() { return MODE == 2 ? 16 : 4; }
template <int MODE>
constexpr int num_threads() { return (epi_warps<MODE>() + 2 + (MODE == 2 ? 1 : 0)) * 32; }  // + MMA warp, TMA producer, (transposed conv) residual loader
constexpr int kMaxSlots = 8;
constexpr int kMaxSets = 32;  // TMEM accumulator blocks (512 columns / NT) or sets (transposed conv)
constexpr int kWHalf = 32;  // the packed weights are organised in 32-input-channel halves

struct TcParams {
  const uint16_t* x;         // (B, Di, Hi, Wi, Cin) bf16
  const uint16_t* wpk;       // packed weights [nblk][khalf][27][4][NT][8] bf16
  const float* scale;        // [Co] or null
  const float* shift;        // [Co] or null
  const uint16_t* res;       // (B, Do, Ho, Wo, Co) bf16 or null
  const float* res_f32;      // (B, Do, Ho, Wo, CoReal) fp32 or null (classifier chain)
  uint16_t* out;             // (B, Do, Ho, Wo, Co) bf16 (or null when out_f32)
  float* out_f32;            // (B, Do, Ho, Wo, CoReal) fp32 or null
  int B, Cin, Co, CoReal;
  int Di, Hi, Wi, Do, Ho, Wo;
  int tiles_h, tiles_w, nchunks, chunk, nblk;  // chunk: output planes per item (mode 0/1), input planes (mode 2)
  int total_items;
  int nslots, relu;
  int tmem_cols;             // 512 (one CTA per SM) or 256 (two co-resident CTAs: two independent MMA issue streams)
  long long* dbg;            // optional per-CTA timing records [grid][4]: smid, t_start, t_end, n_items (profiling aid)
};

template <int MODE>
struct Geo;
template <>
struct Geo<0> {  // stride 1: halo 18 x 10
  static constexpr int HV = 18, WV = 10, ROW = 10, ACCS = 1, NBOX = 1, NVOX = 180;
};
template <>
struct Geo<1> {  // stride 2: halo 33 x 17 loaded as 4 parity sub-planes of 17 x 9 (TMA traversal stride 2), layout [parity][chunk]
  static constexpr int HV = 33, WV = 17, ROW = 9, ACCS = 1, NBOX = 4, NVOX = 153;
};
template <>
struct Geo<2> {  // transposed stride 2: halo 17 x 9 (one extra row/col on the high side), 4 parity accumulators
  static constexpr int HV = 17, WV = 9, ROW = 9, ACCS = 4, NBOX = 1, NVOX = 153;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must abort the kernel (trap -> launch error), never hang the GPU box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
    if (it > (1u << 24)) {
      printf("conv3d_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// zero 32 (or 16) consecutive TMEM columns of this warp's 32 lanes
template <int NCOL>
__device__ __forceinline__ void tmem_zero(uint32_t taddr) {
  const uint32_t z = 0;
  if (NCOL == 32) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
        : "memory");
  } else {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
  }
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// MMA with descriptors passed as (lo, hi) halves so that per-tap updates are single 32-bit adds
__device__ __forceinline__ void umma_bf16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout = 0) {  // version 1 at bit 46, layout type at bits 61-63
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout << 29);
}

// UMMA shared-memory descriptor, K-major, no swizzle ("interleave"): core matrix = 8 rows x 16 B, rows 16 B apart;
// LBO = byte distance between the two 16-byte K chunks of one MMA, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46);  // descriptor version 1 (Blackwell); base_offset 0, layout_type 0 = SWIZZLE_NONE
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = NT
__host__ __device__ constexpr uint32_t make_idesc(int n, int fmt = 0) {
  return (1u << 4) | ((fmt == 0 ? 1u : 0u) << 7) | ((fmt == 0 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Item {
  int nb, b, th, tw, ch;
};
__device__ __forceinline__ Item decode_item(const TcParams& p, int item) {
  Item it;
  it.ch = item % p.nchunks;
  item /= p.nchunks;
  it.tw = item % p.tiles_w;
  item /= p.tiles_w;
  it.th = item % p.tiles_h;
  item /= p.tiles_h;
  it.b = item % p.B;
  it.nb = item / p.B;
  return it;
}

// range of input planes [p0, p1] staged for chunk `ch`, and range of output planes [o0, o1) it produces
template <int MODE>
__device__ __forceinline__ void chunk_ranges(const TcParams& p, int ch, int& p0, int& p1, int& o0, int& o1) {
  if (MODE == 0) {
    o0 = ch * p.chunk, o1 = min(o0 + p.chunk, p.Do);
    p0 = max(o0 - 1, 0), p1 = min(o1, p.Di - 1);
  } else if (MODE == 1) {
    o0 = ch * p.chunk, o1 = min(o0 + p.chunk, p.Do);
    p0 = max(2 * o0 - 1, 0), p1 = min(2 * (o1 - 1) + 1, p.Di - 1);
  } else {
    const int j0 = ch * p.chunk, j1 = min(j0 + p.chunk, p.Di);
    o0 = 2 * j0, o1 = 2 * j1;
    p0 = j0, p1 = min(j1, p.Di - 1);
  }
}
// first / last input plane contributing to output plane od
template <int MODE>
__device__ __forceinline__ void contrib_range(const TcParams& p, int od, int& first, int& last) {
  if (MODE == 0) {
    first = max(od - 1, 0), last = min(od + 1, p.Di - 1);
  } else if (MODE == 1) {
    first = max(2 * od - 1, 0), last = min(2 * od + 1, p.Di - 1);
  } else {
    first = od >> 1, last = (od & 1) ? min((od >> 1) + 1, p.Di - 1) : (od >> 1);
  }
}
// output plane fed by input plane pl through depth tap kd (or -1)
template <int MODE>
__device__ __forceinline__ int out_plane(int pl, int kd) {
  if (MODE == 0) return pl + 1 - kd;
  if (MODE == 1) {
    const int n = pl + 1 - kd;
    return (n & 1) ? -1 : (n >> 1);
  }
  return 2 * pl - 1 + kd;
}

// Weight block order inside one (tap, k-chunk) group for the depth-stacked MMAs (mode 0/1): blocks are sorted by
// ascending OUTPUT plane so that one MMA with N = nb*NT updates nb adjacent TMEM accumulator blocks at once.
//   mode 0: block = 2 - kd  (kd = 2 -> oldest output plane p-1, kd = 0 -> newest p+1)
//   mode 1: odd input planes feed kd = 2 (block 0) and kd = 0 (block 1); even planes feed kd = 1 (block 2)
//   mode 2: block = kd (input plane p feeds output planes 2p-1, 2p, 2p+1)
__host__ __device__ inline int wblock_of_kd(int mode, int kd) { return mode == 0 ? 2 - kd : mode == 2 ? kd : (kd == 2 ? 0 : (kd == 0 ? 1 : 2)); }

template <int MODE, int NT, int FMT, int SC>
__global__ void __launch_bounds__(num_threads<MODE>(), MODE == 2 ? 1 : 2) conv3d_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_res) {
  using G = Geo<MODE>;
  constexpr int kEpiWarps = epi_warps<MODE>();
  constexpr int kThreads = num_threads<MODE>();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle patterns are anchored at 1024-byte boundaries
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform by construction
  long long t_start = 0;
  if (p.dbg != nullptr && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
  const int lane = threadIdx.x & 31;
  // A stage = SC channels (SC*2-byte rows) of one halo plane, laid out [h][w][SC] with the TMA/UMMA hardware swizzle of the row
  // width (64 B -> SWIZZLE_64B, 128 B -> SWIZZLE_128B): one TMA element per voxel instead of one per 16-byte chunk (with
  // un-swizzled 16-byte elements the TMA writes alone kept the shared-memory port busy ~720 cycles per 1010-cycle stage).
  // A tap view is still just a shifted descriptor start address: the swizzle XOR is a function of the absolute shared-memory
  // address for both the TMA write and the UMMA read, so any row-aligned start inside a 1024-byte-aligned box is consistent.
  constexpr uint32_t RB = SC * 2;                                        // row bytes
  constexpr uint32_t kLayoutA = (SC == 32) ? 4u : 2u;                    // UMMA layout type: SWIZZLE_64B / SWIZZLE_128B
  constexpr uint32_t kBoxStride = (G::NVOX * RB + 1023) & ~1023u;
  constexpr uint32_t kSlotBytes = G::NBOX * kBoxStride;
  constexpr int KS = SC / 16;                                            // MMA K steps per tap and stage
  const int KH = p.Cin / SC;                                             // stages per plane (2 only for stride-2 layers with Cin = 64)
  const uint32_t w_bytes = (uint32_t)(p.Cin / kWHalf) * 27 * 4 * NT * 16;
  // TMEM: mode 0/1 use a ring of R = 512/NT accumulator blocks (one per output plane in flight); mode 2 a ring of 3
  // sets x 4 parity classes.
  constexpr int kSetCols = G::ACCS * NT;
  const int R = (MODE == 2) ? 4 : p.tmem_cols / NT;  // mode 2: ring of 4 output planes x 4 parity classes x NT columns
  const uint32_t kTmemCols = (uint32_t)p.tmem_cols;
  static_assert((MODE != 2 || 4 * 4 * NT <= 512) && 512 / NT <= kMaxSets, "TMEM ring overflow");
  constexpr int NTP = NT > 32 ? 32 : NT;  // accumulator columns drained per pass (NT = 64: two passes of 32)
  constexpr int NPASS = NT / NTP;

  // ---- shared memory carve-up: [weights][slots][barriers][tmem ptr]
  uint8_t* w_s = smem;
  uint8_t* slots_s = smem + ((w_bytes + 1023) & ~1023u);  // swizzled boxes: 1024-byte aligned
  // transposed conv: staging tiles [parity class 4][buffer 2][quadrant 4] x 2 KB (32 voxels x 32 channels, 64-byte swizzle),
  // one per epilogue warp and buffer.  The residual of a plane arrives there by TMA -- one stride-2 box per class over the
  // output-shaped tensor, issued by a dedicated loader warp up to two planes ahead, so its HBM latency and the issue cost
  // are off the epilogue's per-plane critical path -- the output piece is staged in the same tile and leaves through
  // lane-contiguous 16-byte stores (row-per-thread global accesses, 64 B per thread at a 128 B stride, cost 32 L1
  // wavefronts per instruction).
  uint8_t* epi_s = slots_s + (size_t)p.nslots * kSlotBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_s + (MODE == 2 ? kEpiWarps * 4096 : 0));
  uint64_t* full_bar = bars;                               // [nslots] producers -> MMA
  uint64_t* empty_bar = bars + kMaxSlots;                  // [nslots] MMA (commit) -> producers
  uint64_t* tfull_bar = bars + 2 * kMaxSlots;              // [R]      MMA (commit) -> epilogue
  uint64_t* tempty_bar = bars + 2 * kMaxSlots + kMaxSets;  // [R]      epilogue -> MMA
  uint64_t* rfull_bar = bars + 2 * kMaxSlots + 2 * kMaxSets;  // [class 4][buffer 2] residual loader -> epilogue (transposed conv)
  uint64_t* rempty_bar = rfull_bar + 8;                       // [class 4][buffer 2] epilogue (4 warps) -> residual loader
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * kMaxSlots + 2 * kMaxSets + (MODE == 2 ? 2 * kEpiWarps : 0));
  // BN affine of all output channels, interleaved per 4 channels as {scale x4, shift x4} (1 / 0 where the caller passed NULL):
  // the epilogue applies it unconditionally with packed FFMA2 fed by broadcast LDS.128
  float4* ss_s = reinterpret_cast<float4*>(tmem_ptr_s + 4);
  for (int c = threadIdx.x; c < p.Co; c += kThreads) {
    float* e = reinterpret_cast<float*>(ss_s + 2 * (c >> 2)) + (c & 3);
    e[0] = (p.scale && c < p.CoReal) ? p.scale[c] : 1.f;
    e[4] = (p.shift && c < p.CoReal) ? p.shift[c] : 0.f;
  }

  if (threadIdx.x == 0) {
    if (MODE == 2)
      for (int i = 0; i < 8; ++i) {
        mbar_init(smem_u32(rfull_bar + i), 1);
        mbar_init(smem_u32(rempty_bar + i), 4);
      }
    for (int i = 0; i < p.nslots; ++i) {
      mbar_init(smem_u32(full_bar + i), 1);
      mbar_init(smem_u32(empty_bar + i), 1);
    }
    for (int i = 0; i < R; ++i) {
      mbar_init(smem_u32(tfull_bar + i), 1);
      mbar_init(smem_u32(tempty_bar + i), kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps) {  // the MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Work distribution: items are nb-major (nb = output-channel block); the grid is split into nblk equal groups of
  // CTAs so that a CTA only ever sees one nb and keeps that block's weights resident for its whole lifetime.
  const int per_nb = p.total_items / p.nblk;
  const int ctas_per_nb = gridDim.x / p.nblk;
  const int nb_of_cta = blockIdx.x / ctas_per_nb;
  const int lane_cta = blockIdx.x - nb_of_cta * ctas_per_nb;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.wpk) + (size_t)nb_of_cta * (w_bytes / 16);
    uint4* dst = reinterpret_cast<uint4*>(w_s);
    for (uint32_t i = threadIdx.x; i < w_bytes / 16; i += kThreads) dst[i] = __ldg(src + i);
  }
  fence_proxy_async();  // generic-proxy smem writes (weights) -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  if (warp < kEpiWarps) {
    // depth-stacked MMAs always accumulate: every accumulator block starts at zero (and is re-zeroed by the epilogue)
    for (int c = (warp >> 2) * 32; c < p.tmem_cols; c += 32 * (kEpiWarps / 4)) tmem_zero<32>(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (MODE == 2 && warp == kEpiWarps + 2) {
    // =========================================================== RESIDUAL LOADER (transposed conv, one lane)
    if (lane == 0 && p.res != nullptr) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_res)) : "memory");
      uint32_t n = 0;  // output planes in the order the epilogue drains them
      for (int li = lane_cta; li < per_nb; li += ctas_per_nb) {
        const Item it = decode_item(p, nb_of_cta * per_nb + li);
        int p0, p1, o0, o1;
        chunk_ranges<MODE>(p, it.ch, p0, p1, o0, o1);
        for (int od = o0; od < o1; ++od, ++n) {
          const uint32_t buf = n & 1, use = n >> 1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            mbar_wait(smem_u32(rempty_bar + c * 2 + buf), (use & 1) ^ 1);
            const uint32_t bar = smem_u32(rfull_bar + c * 2 + buf);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(8192) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                             smem_u32(epi_s + (size_t)(c * 2 + buf) * 8192)),
                         "l"(reinterpret_cast<uint64_t>(&tmap_res)), "r"(it.nb * NT), "r"(2 * (it.tw * 8) + (c & 1)), "r"(2 * (it.th * 16) + (c >> 1)), "r"(it.b * p.Do + od),
                         "r"(bar)
                         : "memory");
          }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // =========================================================== TMA PRODUCER (one lane)
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
      uint32_t slot = 0, phase = 0;
      for (int li = lane_cta; li < per_nb; li += ctas_per_nb) {
        const Item it = decode_item(p, nb_of_cta * per_nb + li);
        int p0, p1, o0, o1;
        chunk_ranges<MODE>(p, it.ch, p0, p1, o0, o1);
        const int h0 = it.th * 16, w0 = it.tw * 8;
        const int gh0 = (MODE == 0) ? h0 - 1 : (MODE == 1) ? 2 * h0 - 1 : h0;
        const int gw0 = (MODE == 0) ? w0 - 1 : (MODE == 1) ? 2 * w0 - 1 : w0;
        for (int pl = p0; pl <= p1; ++pl) {
          const int plane = it.b * p.Di + pl;
          for (int kh = 0; kh < KH; ++kh) {
            mbar_wait(smem_u32(empty_bar + slot), phase ^ 1);
            const uint32_t bar = smem_u32(full_bar + slot);
            const uint32_t sbase = smem_u32(slots_s + (size_t)slot * kSlotBytes);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(G::NBOX * G::NVOX * RB)) : "memory");
#pragma unroll
            for (int q = 0; q < G::NBOX; ++q) {  // q: parity sub-plane (stride-2 only)
              const int cw = gw0 + (q & 1), chh = gh0 + (q >> 1);
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                      sbase + (uint32_t)q * kBoxStride),
                  "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(kh * SC), "r"(cw), "r"(chh), "r"(plane), "r"(bar)
                  : "memory");
            }
            if (++slot == (uint32_t)p.nslots) slot = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kEpiWarps) {
    // =========================================================== MMA ISSUER
    // The whole warp runs the (warp-uniform) control flow so that descriptors live in uniform registers; a single
    // elected lane issues tcgen05.mma / tcgen05.commit.
    const uint32_t w_base = smem_u32(w_s);
    const uint32_t a_hi = desc_hi(G::ROW * RB, kLayoutA);
    const uint32_t b_hi = desc_hi(128);
    // mode 0/1: accumulator block of output plane od = od - o0 (chunk <= R, so the depth-stacked window never wraps:
    // switching the accumulator address between consecutive MMAs costs ~230 cycles, measured with tools/umma_bench.cu).
    // use_mask holds the mbarrier phase parity of every block (toggled per use).  mode 2: set = job % R.
    uint32_t stage = 0, job_base = 0, use_mask = 0;
    uint32_t slot = 0, phase = 0;
    long long dbg_tempty = 0, dbg_full = 0, dbg_grp = 0, dbg_t0 = p.dbg ? clock64() : 0;
    int dbg_nst = 0;
    const uint32_t a_lo0 = desc_lo(smem_u32(slots_s), 16);
    const uint32_t b_lo0 = desc_lo(w_base, 3 * NT * 16);
    uint32_t a_lo = a_lo0;
    if (MODE != 2) {
      // ---- depth-stacked issue, software pipelined.  One stage = (input plane, 32-channel half) = 18 MMAs (9 in-plane taps x
      // 2 K steps) with N = nb*NT covering the nb <= 3 output planes fed by this input plane (weight blocks [wb, wb+nb)).
      // The tensor pipe drains one N=96 MMA per ~56 cycles and queues only ~8 of them, while a stage's bookkeeping (window
      // arithmetic, two mbarrier try_waits of ~90 cycles each, commits, iterator advance) costs the issuing warp ~1000 cycles.
      // Run back to back (issue 18, then bookkeep) the pipe idled ~45 % of the time.  Here the bookkeeping of stage s+1 is cut
      // into three slices that execute between the three 6-MMA groups of stage s, i.e. while the queue is still full.
      int li = lane_cta;
      bool have = li < per_nb;
      Item it;
      int p0 = 0, p1 = -1, o0 = 0, o1 = 0, nout = 0, pl = 0, kh = 0, next_new = 0, next_done = 0;
      auto begin_item = [&]() {
        it = decode_item(p, nb_of_cta * per_nb + li);
        chunk_ranges<MODE>(p, it.ch, p0, p1, o0, o1);
        nout = o1 - o0, pl = p0, kh = 0, next_new = 0, next_done = 0;
      };
      auto window = [&](int& oa, int& ob, int& wb) {
        if (MODE == 0) {
          const int rel = pl - o0;
          oa = max(rel - 1, 0), ob = min(rel + 1, nout - 1), wb = oa - rel + 1;
        } else {
          const int rel2 = pl - (2 * o0 - 1);
          if (!(rel2 & 1)) {
            const int od1 = (rel2 >> 1) - 1;
            oa = max(od1, 0), ob = min(od1 + 1, nout - 1), wb = oa - od1;
          } else {
            oa = ob = rel2 >> 1, wb = 2;
          }
        }
      };
      auto wait_new_blocks = [&](int ob) {  // first touch of a block: the epilogue must have drained + zeroed it
        const long long c0 = p.dbg ? clock64() : 0;
        for (; next_new <= ob; ++next_new) {
          mbar_wait(smem_u32(tempty_bar + next_new), ((use_mask >> next_new) & 1) ^ 1);
          use_mask ^= 1u << next_new;
        }
        if (p.dbg) dbg_tempty += clock64() - c0;
      };
      auto wait_full = [&](uint32_t sl, uint32_t ph) {
        const long long c0 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(full_bar + sl), ph);
        if (p.dbg) {
          dbg_full += clock64() - c0;
          if (blockIdx.x < 4 && dbg_nst < 60 && lane == 0) p.dbg[8 * gridDim.x + blockIdx.x * 64 + dbg_nst] = clock64() - dbg_t0;
          ++dbg_nst;
        }
        tc_fence_after();
      };
      auto issue_taps = [&](int t0, uint32_t d1, uint32_t i1, uint32_t a_cur, uint32_t b0) {
#pragma unroll
        for (int t = t0; t < t0 + 3; ++t) {
          const int th_ = t / 3, tw_ = t % 3;
          // byte offset of the tap's view inside the stage (stride 2: parity sub-plane q, then (kh>>1, kw>>1))
          const uint32_t voff = (MODE == 0) ? (uint32_t)(th_ * 10 + tw_) * RB
                                            : (uint32_t)(((th_ & 1) << 1) | (tw_ & 1)) * kBoxStride + (uint32_t)((th_ >> 1) * 9 + (tw_ >> 1)) * RB;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            // weight slab of this K step: 32-channel half (ks >> 1), 8-channel chunk pair (ks & 1) * 2
            const uint32_t boff = (uint32_t)(ks >> 1) * (27 * 4 * NT * 16) + (uint32_t)(t * 4 + (ks & 1) * 2) * (3 * NT * 16);
            umma_bf16_lh(d1, a_cur + ((voff + ks * 32) >> 4), a_hi, b0 + (boff >> 4), b_hi, i1, 1u);
          }
        }
      };
      // ---- prologue: first stage prepared with blocking waits
      int c_oa = 0, c_ob = -1, c_wb = 0;
      if (have) {
        begin_item();
        window(c_oa, c_ob, c_wb);
        wait_new_blocks(c_ob);
        wait_full(slot, phase);
      }
      while (have) {
        // parameters of the CURRENT stage
        const uint32_t i1 = make_idesc((c_ob - c_oa + 1) * NT, FMT);
        const uint32_t d1 = tmem_base + (uint32_t)c_oa * NT;
        const uint32_t b0 = b_lo0 + (uint32_t)(kh * (SC / kWHalf)) * ((27 * 4 * NT * 16) >> 4) + (uint32_t)c_wb * ((NT * 16) >> 4);
        const uint32_t a_cur = a_lo, cur_slot = slot;
        const int cur_pl = pl, cur_p1 = p1, cur_o0 = o0, cur_ob = c_ob, cur_nd0 = next_done;
        const bool cur_last_kh = (kh == KH - 1);
        long long g0 = p.dbg ? clock64() : 0;
        if (elect_one()) issue_taps(0, d1, i1, a_cur, b0);
        __syncwarp();
        if (p.dbg) dbg_grp += clock64() - g0;
        // ---- slice A (queue full): completed-block count of the current stage, iterator advance, next window
        int ncommit = 0;
        if (cur_last_kh) {
          // block b is complete after its last contributing plane: o0+b+1 (mode 0) / 2(o0+b)+1 (mode 1), or the item's last plane
          while (next_done <= cur_ob && (cur_pl == cur_p1 || cur_pl == ((MODE == 0) ? cur_o0 + next_done + 1 : 2 * (cur_o0 + next_done) + 1))) ++next_done, ++ncommit;
        }
        bool new_plane = false, new_item = false;
        if (++kh == KH) {
          kh = 0, new_plane = true;
          if (++pl > p1) {
            li += ctas_per_nb;
            have = li < per_nb;
            new_item = true;
            if (have) begin_item();
          }
        }
        if (++slot == (uint32_t)p.nslots) slot = 0, phase ^= 1, a_lo = a_lo0;
        else a_lo += kSlotBytes >> 4;
        int n_oa = c_oa, n_ob = c_ob, n_wb = c_wb;
        if (have && new_plane) window(n_oa, n_ob, n_wb);
        g0 = p.dbg ? clock64() : 0;
        if (elect_one()) issue_taps(3, d1, i1, a_cur, b0);
        __syncwarp();
        if (p.dbg) dbg_grp += clock64() - g0;
        // ---- slice B: TMEM blocks first touched by the next stage.  (Across an item boundary the block may be one whose
        // completion is only committed below -- short chunks -- so that wait moves behind the commits.)
        if (have && new_plane && !new_item) wait_new_blocks(n_ob);
        g0 = p.dbg ? clock64() : 0;
        if (elect_one()) {
          issue_taps(6, d1, i1, a_cur, b0);
          for (int nd = cur_nd0; nd < cur_nd0 + ncommit; ++nd) umma_commit(smem_u32(tfull_bar + nd));  // accumulator complete -> epilogue
          umma_commit(smem_u32(empty_bar + cur_slot));                                                 // stage consumed -> TMA may refill it
        }
        __syncwarp();
        if (p.dbg) dbg_grp += clock64() - g0;
        // ---- slice C: operands of the next stage
        if (have && new_item) wait_new_blocks(n_ob);
        if (have) wait_full(slot, phase);
        c_oa = n_oa, c_ob = n_ob, c_wb = n_wb;
      }
    } else {
    for (int li = lane_cta; li < per_nb; li += ctas_per_nb) {
      const Item it = decode_item(p, nb_of_cta * per_nb + li);
      int p0, p1, o0, o1;
      chunk_ranges<MODE>(p, it.ch, p0, p1, o0, o1);
      for (int pl = p0; pl <= p1; ++pl) {
        {
          // ---- transposed conv, depth-stacked: accumulator block of (parity class c, output plane od) sits at TMEM column
          // (c*4 + job%4)*NT, so for a fixed class the <= 3 output planes fed by this input plane (kd = 0,1,2 -> od = 2p-1,
          // 2p, 2p+1) are adjacent column blocks and one MMA with N = nb*NT updates them all (two MMAs when the ring of 4
          // wraps).  9 (class, in-plane shift) views x KS K-steps = 36..72 MMAs per plane instead of 108 N=32 ones.
          const uint32_t slot_ = stage % p.nslots, phase_ = (stage / p.nslots) & 1;
          ++stage;
          const int od_lo = max(2 * pl - 1, o0), od_hi = min(2 * pl + 1, o1 - 1);
          const int nb = od_hi - od_lo + 1, kd_a = od_lo - (2 * pl - 1);
          const uint32_t ja = job_base + (uint32_t)(od_lo - o0);
          const uint32_t ca = ja % 4;
          const int n1 = min(nb, (int)(4 - ca)), n2 = nb - n1;
          long long c0 = p.dbg ? clock64() : 0;
          for (int od = od_lo; od <= od_hi; ++od) {  // first touch: the epilogue must have drained + zeroed the 4 class blocks
            int first, last;
            contrib_range<MODE>(p, od, first, last);
            if (first == pl) {
              const uint32_t job = job_base + (uint32_t)(od - o0);
              mbar_wait(smem_u32(tempty_bar + job % 4), ((job / 4) & 1) ^ 1);
            }
          }
          if (p.dbg) dbg_tempty += clock64() - c0;
          c0 = p.dbg ? clock64() : 0;
          mbar_wait(smem_u32(full_bar + slot_), phase_);
          if (p.dbg) dbg_full += clock64() - c0;
          tc_fence_after();
          c0 = p.dbg ? clock64() : 0;
          if (nb > 0 && elect_one()) {
            const uint32_t a0 = desc_lo(smem_u32(slots_s + (size_t)slot_ * kSlotBytes), 16);
            const uint32_t b0 = b_lo0 + (uint32_t)kd_a * ((NT * 16) >> 4);
            const uint32_t i1 = make_idesc(n1 * NT, FMT), i2 = make_idesc(max(n2, 1) * NT, FMT);
#pragma unroll
            for (int cls = 0; cls < 4; ++cls) {
              const int ph = cls >> 1, pw = cls & 1;
              const uint32_t d1 = tmem_base + (uint32_t)(cls * 4 + ca) * NT, d2 = tmem_base + (uint32_t)(cls * 4) * NT;
#pragma unroll
              for (int ih = 0; ih <= ph; ++ih) {
#pragma unroll
                for (int iw = 0; iw <= pw; ++iw) {
                  // output parity 0 <- tap 1 (same index); parity 1 <- tap 2 (same index) and tap 0 (index + 1)
                  const int th_ = ph ? (ih ? 0 : 2) : 1, tw_ = pw ? (iw ? 0 : 2) : 1;
#pragma unroll
                  for (int ks = 0; ks < KS; ++ks) {
                    const uint32_t a_lo = a0 + (((uint32_t)(ih * 9 + iw) * RB + ks * 32) >> 4);
                    const uint32_t b_lo = b0 + (((uint32_t)(ks >> 1) * (27 * 4 * NT * 16) + (uint32_t)((th_ * 3 + tw_) * 4 + (ks & 1) * 2) * (3 * NT * 16)) >> 4);
                    umma_bf16_lh(d1, a_lo, a_hi, b_lo, b_hi, i1, 1u);
                    if (n2 > 0) umma_bf16_lh(d2, a_lo, a_hi, b_lo + (((uint32_t)n1 * NT * 16) >> 4), b_hi, i2, 1u);
                  }
                }
              }
            }
            for (int od = od_lo; od <= od_hi; ++od) {
              int first, last;
              contrib_range<MODE>(p, od, first, last);
              if (last == pl) umma_commit(smem_u32(tfull_bar + (job_base + (uint32_t)(od - o0)) % 4));  // all 4 class blocks complete
            }
          }
          __syncwarp();
          if (elect_one()) umma_commit(smem_u32(empty_bar + slot_));
          __syncwarp();
          if (p.dbg) dbg_grp += clock64() - c0;
        }
      }
      job_base += (uint32_t)(o1 - o0);
    }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 8 + 4] = dbg_tempty;
      p.dbg[blockIdx.x * 8 + 5] = dbg_full;
      p.dbg[blockIdx.x * 8 + 3] = dbg_grp;
      p.dbg[blockIdx.x * 8 + 6] = clock64() - dbg_t0;
    }
  } else {
    // =========================================================== EPILOGUE (4 warps = 128 rows)
    const int row = (warp & 3) * 32 + lane;  // TMEM lane == GEMM row == tile position
    const int cls = (MODE == 2) ? (warp >> 2) : 0;  // transposed conv: this warp's output parity class
    const int hl = row >> 3, wl = row & 7;
    uint32_t job_base = 0, use_mask = 0;
    long long dbg_tfull = 0;
    const bool res_tma = MODE == 2 && p.res != nullptr;
    uint32_t nplane = 0;  // transposed conv: output planes drained so far (staging buffer = nplane & 1)
    for (int li = lane_cta; li < per_nb; li += ctas_per_nb) {
      const Item it = decode_item(p, nb_of_cta * per_nb + li);
      int p0, p1, o0, o1;
      chunk_ranges<MODE>(p, it.ch, p0, p1, o0, o1);
      const int n0 = it.nb * NT;
      for (int od = o0; od < o1; ++od) {
        const uint32_t job = job_base + (uint32_t)(od - o0);
        uint32_t set, par;
        if (MODE != 2) {
          set = (uint32_t)(od - o0), par = (use_mask >> set) & 1;
          use_mask ^= 1u << set;
        } else {
          set = job % R, par = (job / R) & 1;
        }
        // this warp's output voxel; residual operands do not depend on the accumulator: fetch them BEFORE waiting for it
        int oh, ow;
        bool ok;
        if (MODE == 2) {
          const int ih = it.th * 16 + hl, iw = it.tw * 8 + wl;
          oh = 2 * ih + (cls >> 1), ow = 2 * iw + (cls & 1);
          ok = ih < p.Hi && iw < p.Wi;
        } else {
          oh = it.th * 16 + hl, ow = it.tw * 8 + wl;
          ok = oh < p.Ho && ow < p.Wo;
        }
        const size_t vox = (((size_t)it.b * p.Do + od) * p.Ho + oh) * p.Wo + ow;
        uint4 rpre[NTP / 8];
        const float rpre_f32 = (ok && p.res_f32) ? __ldg(p.res_f32 + vox * p.CoReal) : 0.f;
        // lane-contiguous view of this warp's 32 rows x 64 B: instruction k, lane t <-> row 8k + (t >> 2), 16-byte chunk t & 3
        uint8_t* tile_res = epi_s + (size_t)(((cls * 2 + (nplane & 1)) * 4 + (warp & 3)) * 2048);
        uint8_t* tile_out = tile_res;  // a thread overwrites exactly the chunks it has just read
        size_t vox_c[4];
        bool ok_c[4];
        if (MODE == 2 && NT == 32) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int ih = it.th * 16 + (warp & 3) * 4 + k, iw = it.tw * 8 + (lane >> 2);
            ok_c[k] = ih < p.Hi && iw < p.Wi;
            vox_c[k] = (((size_t)it.b * p.Do + od) * p.Ho + (2 * ih + (cls >> 1))) * p.Wo + (2 * iw + (cls & 1));
          }
          if (res_tma) {
            mbar_wait(smem_u32(rfull_bar + cls * 2 + (nplane & 1)), (nplane >> 1) & 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) rpre[q] = *reinterpret_cast<const uint4*>(tile_res + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4));
          }
        } else {
#pragma unroll
          for (int q = 0; q < NTP / 8; ++q) rpre[q] = (ok && p.res) ? ld_nc_v4(p.res + vox * p.Co + n0 + q * 8) : make_uint4(0, 0, 0, 0);
        }
        const long long e0 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(tfull_bar + set), par);
        if (p.dbg) dbg_tfull += clock64() - e0;
        tc_fence_after();
#pragma unroll
        for (int ps = 0; ps < NPASS; ++ps) {
          uint32_t v[32];
          const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (MODE == 2 ? (uint32_t)(cls * 4 + set) * NT : set * NT) + ps * 32;
          const int nn0 = n0 + ps * 32;
          if (NTP == 32)
            tmem_ld32(taddr, v);
          else
            tmem_ld16(taddr, v);
          tmem_ld_wait();
          tmem_zero<NTP>(taddr);  // hand the block back zeroed (completion awaited below)
          if (ok) {
            if (p.out_f32 != nullptr) {
              // classifier: only the first CoReal (=1) channels are real
#pragma unroll
              for (int c = 0; c < NTP; ++c) {
                if (c >= p.CoReal) break;
                float y = __uint_as_float(v[c]);
                if (p.scale) y *= __ldg(p.scale + c);
                if (p.shift) y += __ldg(p.shift + c);
                if (p.res_f32) y += (c == 0) ? rpre_f32 : __ldg(p.res_f32 + vox * p.CoReal + c);
                if (p.relu) y = fmaxf(y, 0.f);
                p.out_f32[vox * p.CoReal + c] = y;
              }
            } else {
              uint16_t* op = p.out + vox * p.Co + nn0;
#pragma unroll
              for (int q = 0; q < NTP / 8; ++q) {
                float y[8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // y = acc * scale + shift (+ residual), two channels per instruction
                  const float4 sc = ss_s[2 * ((nn0 >> 2) + q * 2 + h)], sh = ss_s[2 * ((nn0 >> 2) + q * 2 + h) + 1];
                  ffma2(y[4 * h], y[4 * h + 1], __uint_as_float(v[q * 8 + 4 * h]), __uint_as_float(v[q * 8 + 4 * h + 1]), sc.x, sc.y, sh.x, sh.y);
                  ffma2(y[4 * h + 2], y[4 * h + 3], __uint_as_float(v[q * 8 + 4 * h + 2]), __uint_as_float(v[q * 8 + 4 * h + 3]), sc.z, sc.w, sh.z, sh.w);
                }
                if (p.res) {
                  const uint4 r = (NPASS == 1) ? rpre[q] : ld_nc_v4(p.res + vox * p.Co + nn0 + q * 8);
                  const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float r0, r1;
                    unpack2<FMT>(rr[j], r0, r1);
                    fadd2(y[2 * j], y[2 * j + 1], y[2 * j], y[2 * j + 1], r0, r1);
                  }
                }
                uint4 o;
                o.x = pack2<FMT>(y[0], y[1]), o.y = pack2<FMT>(y[2], y[3]), o.z = pack2<FMT>(y[4], y[5]), o.w = pack2<FMT>(y[6], y[7]);
                if (p.relu) o.x = relu2<FMT>(o.x), o.y = relu2<FMT>(o.y), o.z = relu2<FMT>(o.z), o.w = relu2<FMT>(o.w);
                if (MODE == 2 && NT == 32)
                  *reinterpret_cast<uint4*>(tile_out + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = o;  // staged, stored below
                else
                  *reinterpret_cast<uint4*>(op + q * 8) = o;
              }
            }
          }
        }
        if (MODE == 2 && NT == 32 && p.out_f32 == nullptr) {
          __syncwarp();
          uint4 o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = 8 * k + (lane >> 2);
            o[k] = *reinterpret_cast<const uint4*>(tile_out + r * 64 + (((lane & 3) ^ ((r >> 1) & 3)) << 4));
          }
          if (res_tma) {  // the tile is free again (before the global stores, so that the fence does not wait on them)
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(rempty_bar + cls * 2 + (nplane & 1)));
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (ok_c[k]) *reinterpret_cast<uint4*>(p.out + vox_c[k] * p.Co + n0 + (lane & 3) * 8) = o[k];
          ++nplane;
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(tempty_bar + set));
      }
      job_base += (uint32_t)(o1 - o0);
    }
    if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 8 + 7] = dbg_tfull;
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (p.dbg != nullptr && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    p.dbg[blockIdx.x * 8 + 0] = smid;
    p.dbg[blockIdx.x * 8 + 1] = t_start;
    p.dbg[blockIdx.x * 8 + 2] = t1;
    // slot 3: cycles the MMA warp spent inside its issue groups (written by the MMA warp)
  }
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// weights (Co,Ci,3,3,3) [mode 0/1] or (Ci,Co,3,3,3) [mode 2], fp32 -> [nblk][khalf][27][4][NT][8] bf16, zero padded
__global__ void pack_w3d_kernel(const float* __restrict__ w, uint16_t* __restrict__ wp, int Ci, int Co, int NT, int nblk, int mode, int fmt,
                                long long total) {
  // mode 2:   [nblk][khalf][tap 27][kc 4][NT][8]
  // mode 0/1: [nblk][khalf][tap 9 (kh,kw)][kc 4][depth block 3][NT][8]   (see wblock_of_kd)
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long r = e;
    const int j = (int)(r % 8);
    r /= 8;
    const int n = (int)(r % NT);
    r /= NT;
    int kd, t9, kc;
    {
      const int blk = (int)(r % 3);
      r /= 3;
      kc = (int)(r % 4);
      r /= 4;
      t9 = (int)(r % 9);
      r /= 9;
      kd = 0;
      for (int k = 0; k < 3; ++k)
        if (wblock_of_kd(mode, k) == blk) kd = k;
    }
    const int KH = Ci / kWHalf;
    const int kh = (int)(r % KH);
    const int nb = (int)(r / KH);
    const int co = nb * NT + n, ci = kh * kWHalf + kc * 8 + j, t = kd * 9 + t9;
    float v = 0.f;
    if (co < Co) v = (mode == 2) ? w[((size_t)ci * Co + co) * 27 + t] : w[((size_t)co * Ci + ci) * 27 + t];
    wp[e] = float_to_h16_bits(v, fmt);
  }
}

// accumulator block width: 32 output channels (Co = 64 runs as two resident blocks on disjoint halves of the grid), 16 for the
// classifier; the 32 -> 64 stride-2 layer takes all 64 at once (its 110 KB of weights fit: the input is then read once and
// the depth-stacked MMAs are N = 128 instead of 64, i.e. at the full math rate and long enough to hide the issue bookkeeping)
int pick_nt(int Ci, int Co, int mode) { return (mode == 1 && Ci == 32 && Co == 64) ? 64 : (Co >= 32 ? 32 : 16); }

// 4-D tensor map over the NDHWC activation: dims (C, W, H, B*D), box {8 ch, WV, HV, 1}; stride-2 layers traverse w/h with
// element stride 2 (one box per parity).  Out-of-bounds elements (halo outside the volume) are zero-filled by the TMA unit.
int make_tmap(CUtensorMap* tm, const void* x, int fmt, int C, int Wi, int Hi, long long planes, int mode, int SC) {
  static decltype(&cuTensorMapEncodeTiled) encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      set_error("conv3d_tc: cuTensorMapEncodeTiled is not available from this driver");
      return MODE_ECUDA;
    }
    encode = reinterpret_cast<decltype(&cuTensorMapEncodeTiled)>(fn);
  }
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)planes};
  const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wi * C * 2, (cuuint64_t)Hi * Wi * C * 2};
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  if (mode == 0) {
    box[0] = SC, box[1] = 10, box[2] = 18, box[3] = 1;
  } else if (mode == 1) {
    box[0] = SC, box[1] = 18, box[2] = 34, box[3] = 1;  // ceil(18/2) = 9 columns, ceil(34/2) = 17 rows
    estr[1] = 2, estr[2] = 2;
  } else {
    box[0] = SC, box[1] = 9, box[2] = 17, box[3] = 1;
  }
  const CUresult r = encode(tm, fmt == kFmtBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, SC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv3d_tc: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return MODE_ECUDA;
  }
  return MODE_OK;
}

int make_tmap_res(CUtensorMap* tm, const void* ptr, int fmt, int Co, int Wo, int Ho, long long planes) {
  static decltype(&cuTensorMapEncodeTiled) encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      set_error("conv3d_tc: cuTensorMapEncodeTiled is not available from this driver");
      return MODE_ECUDA;
    }
    encode = reinterpret_cast<decltype(&cuTensorMapEncodeTiled)>(fn);
  }
  const cuuint64_t gdim[4] = {(cuuint64_t)Co, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)planes};
  const cuuint64_t gstr[3] = {(cuuint64_t)Co * 2, (cuuint64_t)Wo * Co * 2, (cuuint64_t)Ho * Wo * Co * 2};
  const cuuint32_t box[4] = {32, 16, 32, 1}, estr[4] = {1, 2, 2, 1};  // 8 columns x 16 rows at element stride 2 = one parity class of a tile
  const CUresult r = encode(tm, fmt == kFmtBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv3d_tc: cuTensorMapEncodeTiled (residual) failed (CUresult %d)", (int)r);
    return MODE_ECUDA;
  }
  return MODE_OK;
}

template <int MODE, int NT, int FMT, int SC>
int launch_tc3(const TcParams& p, const CUtensorMap& tm, const CUtensorMap& tmr, int grid, size_t smem, cudaStream_t s) {
  static thread_local size_t attr_dev[kMaxDevices] = {};  // the attribute is per device and per function
  size_t& attr = attr_dev[current_device()];
  if (smem > attr) {
    MODE_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<MODE, NT, FMT, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "conv3d_tc");
    attr = smem;
  }
  conv3d_tc_kernel<MODE, NT, FMT, SC><<<grid, num_threads<MODE>(), smem, s>>>(p, tm, tmr);
  MODE_CHECK_LAUNCH("conv3d_tc");
  return MODE_OK;
}
template <int MODE, int NT, int FMT>
int launch_tc2(const TcParams& p, const CUtensorMap& tm, const CUtensorMap& tmr, int sc, int grid, size_t smem, cudaStream_t s) {
  return sc == 32 ? launch_tc3<MODE, NT, FMT, 32>(p, tm, tmr, grid, smem, s) : launch_tc3<MODE, NT, FMT, 64>(p, tm, tmr, grid, smem, s);
}
template <int MODE, int NT>
int launch_tc(const TcParams& p, const CUtensorMap& tm, const CUtensorMap& tmr, int sc, int fmt, int grid, size_t smem, cudaStream_t s) {
  return fmt == kFmtBF16 ? launch_tc2<MODE, NT, kFmtBF16>(p, tm, tmr, sc, grid, smem, s) : launch_tc2<MODE, NT, kFmtFP16>(p, tm, tmr, sc, grid, smem, s);
}

}  // namespace

static long long* g_tc_dbg = nullptr;
// profiling aid (not part of the reference surface): per-CTA {smid, start ns, end ns, items} records of the next launches
extern "C" int mode_conv3d_set_debug_buffer(void* dev_ptr) {
  g_tc_dbg = (long long*)dev_ptr;
  return MODE_OK;
}

extern "C" size_t mode_conv3d_packed_weight_elems(int Ci, int Co, int mode) {
  const int NT = pick_nt(Ci, Co, mode);
  const int nblk = (Co + NT - 1) / NT;
  return (size_t)nblk * (Ci / kWHalf) * 27 * 4 * NT * 8;
}

extern "C" int mode_conv3d_pack_weights(const float* w, mode_h16* w_packed, int Ci, int Co, int mode, int fmt, void* stream) {
  MODE_CHECK_ARG(w && w_packed, "conv3d_pack_weights: null pointer");
  MODE_CHECK_ARG(Ci == 32 || Ci == 64, "conv3d_pack_weights: Ci must be 32 or 64 (got %d)", Ci);
  MODE_CHECK_ARG(Co >= 1 && mode >= 0 && mode <= 2, "conv3d_pack_weights: bad Co/mode");
  MODE_CHECK_ARG(fmt == kFmtBF16 || fmt == kFmtFP16, "conv3d_pack_weights: fmt must be 0 (bf16) or 1 (fp16)");
  const int NT = pick_nt(Ci, Co, mode);
  const int nblk = (Co + NT - 1) / NT;
  const long long total = (long long)mode_conv3d_packed_weight_elems(Ci, Co, mode);
  pack_w3d_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, w_packed, Ci, Co, NT, nblk, mode, fmt, total);
  MODE_CHECK_LAUNCH("conv3d_pack_weights");
  return MODE_OK;
}

extern "C" int mode_conv3d_tc(const mode_h16* x, const mode_h16* w_packed, const float* scale, const float* shift, const mode_h16* residual,
                              const float* residual_f32, mode_h16* out, float* out_f32, int B, int Ci, int Co, int Di, int Hi, int Wi, int mode,
                              int relu, int fmt, void* stream) {
  MODE_CHECK_ARG(fmt == kFmtBF16 || fmt == kFmtFP16, "conv3d_tc: fmt must be 0 (bf16) or 1 (fp16)");
  MODE_CHECK_ARG(x && w_packed && (out || out_f32), "conv3d_tc: null pointer");
  MODE_CHECK_ARG(Ci == 32 || Ci == 64, "conv3d_tc: Ci must be 32 or 64 (got %d)", Ci);
  MODE_CHECK_ARG(B > 0 && Di > 0 && Hi > 0 && Wi > 0 && mode >= 0 && mode <= 2, "conv3d_tc: bad shape/mode");
  const int NT = pick_nt(Ci, Co, mode);
  MODE_CHECK_ARG(out_f32 ? (Co <= 16) : (Co % 32 == 0), "conv3d_tc: Co=%d unsupported (bf16 out needs Co %% 32 == 0, fp32 out needs Co <= 16)", Co);
  MODE_CHECK_ARG(!(mode == 2 && NT != 32), "conv3d_tc: transposed conv needs Co %% 32 == 0");
  TcParams p;
  p.x = x, p.wpk = w_packed, p.scale = scale, p.shift = shift, p.res = residual, p.res_f32 = residual_f32, p.out = out, p.out_f32 = out_f32;
  p.dbg = g_tc_dbg;
  p.B = B, p.Cin = Ci, p.Co = Co, p.CoReal = Co, p.Di = Di, p.Hi = Hi, p.Wi = Wi, p.relu = relu;
  if (mode == 0) {
    p.Do = Di, p.Ho = Hi, p.Wo = Wi;
  } else if (mode == 1) {
    p.Do = (Di - 1) / 2 + 1, p.Ho = (Hi - 1) / 2 + 1, p.Wo = (Wi - 1) / 2 + 1;
  } else {
    p.Do = 2 * Di, p.Ho = 2 * Hi, p.Wo = 2 * Wi;
  }
  const int th_dim = (mode == 2) ? Hi : p.Ho, tw_dim = (mode == 2) ? Wi : p.Wo, d_dim = (mode == 2) ? Di : p.Do;
  p.tiles_h = ceil_div(th_dim, 16), p.tiles_w = ceil_div(tw_dim, 8);
  p.nblk = ceil_div(Co, NT);
  // shared memory: resident weights + stage ring
  // stage channels: the whole Cin, except stride-2 layers with Cin = 64 (four parity boxes of 128-byte rows would need 80 KB per stage)
  const int SC = (mode == 1) ? 32 : Ci;
  const size_t w_bytes = ((size_t)(Ci / kWHalf) * 27 * 4 * NT * 16 + 1023) & ~(size_t)1023;
  const size_t nvox = (mode == 0) ? Geo<0>::NVOX : Geo<1>::NVOX, nbox = (mode == 1) ? 4 : 1;
  const size_t slot_bytes = nbox * (((nvox * SC * 2) + 1023) & ~(size_t)1023);
  const size_t misc = (2 * kMaxSlots + 2 * kMaxSets + 32) * 8 + 16 + 128 + 1024 + 2 * 64 * 4 + (mode == 2 ? 16 * 4096 : 0);  // barriers, smem base alignment, epilogue staging
  const size_t budget = 227 * 1024;
  // Two co-resident CTAs per SM when they fit (stride-1 / stride-2 layers with <= 55 KB of weights): the MMA-issuing warp is
  // bound by its own instruction latency (~1700 cycles of waits + bookkeeping + issue per 18-MMA stage against ~1000 cycles of
  // tensor-pipe work, measured with the per-CTA timers), so two independent issue streams keep the pipe fed.  Each CTA then
  // owns 256 TMEM columns (8 accumulator blocks of 32) instead of 512.
  int ctas_per_sm = 1;
  if (mode != 2) {
    const size_t half = (budget - 2048) / 2;
    if (w_bytes + misc + 3 * slot_bytes <= half) ctas_per_sm = 2;
  }
  const size_t cta_budget = ctas_per_sm == 2 ? (budget - 2048) / 2 : budget;
  int nslots = (int)std::min<size_t>(kMaxSlots, (cta_budget - w_bytes - misc) / slot_bytes);
  MODE_CHECK_ARG(nslots >= 2, "conv3d_tc: not enough shared memory for a 2-stage pipeline");
  p.nslots = nslots;
  p.tmem_cols = ctas_per_sm == 2 ? 256 : 512;
  const int sm_slots = kNumSMs * ctas_per_sm;
  // depth chunking: split the depth range into nch BALANCED chunks (chunk = ceil(d / nch)); pick the nch that minimises
  // the critical-path stage count  rounds x (chunk + halo)  (persistent CTAs, round-robin items).
  const long long cols = (long long)B * p.tiles_h * p.tiles_w;
  const int halo = (mode == 0) ? 2 : 1;
  const int max_chunk = (mode == 2) ? d_dim : p.tmem_cols / NT;  // mode 0/1: one TMEM accumulator block per output plane of the chunk
  int best_chunk = std::min(d_dim, max_chunk);
  double best = 1e30;
  for (int nch = 1; nch <= d_dim; ++nch) {
    const int c = ceil_div(d_dim, nch);
    if (c > max_chunk) continue;
    const int nch_eff = ceil_div(d_dim, c);
    const long long items = cols * nch_eff;
    const long long ctas = std::min<long long>(items, sm_slots / p.nblk);
    const long long rounds = (items + ctas - 1) / ctas;
    const double stages = (double)(mode == 1 ? 2 * c + halo : c + halo) * (Ci / SC);
    const double cost = (double)rounds * (stages + 1.5);  // +1.5: per-item pipeline ramp
    if (cost < best - 1e-9) best = cost, best_chunk = c;
  }
  p.chunk = best_chunk;
  p.nchunks = ceil_div(d_dim, p.chunk);
  const long long per_nb = cols * p.nchunks;
  MODE_CHECK_ARG(per_nb * p.nblk < 2147483647LL, "conv3d_tc: too many work items");
  p.total_items = (int)(per_nb * p.nblk);
  const int ctas_per_nb = (int)std::min<long long>(per_nb, sm_slots / p.nblk);
  const int grid = ctas_per_nb * p.nblk;
  const size_t smem = w_bytes + (size_t)nslots * slot_bytes + misc;
  cudaStream_t s = (cudaStream_t)stream;
  CUtensorMap tm;
  {
    const int rc = make_tmap(&tm, x, fmt, Ci, Wi, Hi, (long long)B * Di, mode, SC);
    if (rc != MODE_OK) return rc;
  }
  // transposed conv with a residual: stride-2 box (one parity class of a 16 x 8 tile -> 128 output voxels x 32 channels)
  // over the output-shaped residual tensor, (Co, Wo, Ho, B*Do)
  CUtensorMap tmr;
  memset(&tmr, 0, sizeof(tmr));
  if (mode == 2 && residual != nullptr) {
    const int rc = make_tmap_res(&tmr, residual, fmt, Co, p.Wo, p.Ho, (long long)B * p.Do);
    if (rc != MODE_OK) return rc;
  }
  if (NT == 64) return launch_tc<1, 64>(p, tm, tmr, SC, fmt, grid, smem, s);
  if (NT == 32) {
    if (mode == 0) return launch_tc<0, 32>(p, tm, tmr, SC, fmt, grid, smem, s);
    if (mode == 1) return launch_tc<1, 32>(p, tm, tmr, SC, fmt, grid, smem, s);
    return launch_tc<2, 32>(p, tm, tmr, SC, fmt, grid, smem, s);
  }
  if (mode == 0) return launch_tc<0, 16>(p, tm, tmr, SC, fmt, grid, smem, s);
  if (mode == 1) return launch_tc<1, 16>(p, tm, tmr, SC, fmt, grid, smem, s);
  set_error("conv3d_tc: unsupported configuration");
  return MODE_ENOSUP;
}
