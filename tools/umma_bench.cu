// umma_bench.cu -- micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N and of the shared
// memory operand layout (no-swizzle "interleave" with arbitrary LBO/SBO vs canonical 128B swizzle).  Data are zeros;
// only the issue/operand-fetch timing matters.  nvcc -arch=sm_100a -o build/umma_bench tools/umma_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}

struct Cfg {
  int n;        // MMA N
  int layout;   // 0 = no swizzle, 2 = 128B swizzle
  int a_lbo, a_sbo, b_lbo, b_sbo;
  int a_step, b_step;  // descriptor start-address advance (bytes) between consecutive MMAs (cycled over 8 positions)
  int rot_d;    // rotate accumulator column block between MMAs (0 = same D)
  int iters;
  int win;      // >0: sliding window: D = 32 * ((i / win) % 13)  (conv kernel pattern)
  int commit;   // >0: tcgen05.commit to a dummy barrier every `commit` MMAs
  int epi;      // 1: warps 1-3 hammer TMEM with tcgen05.ld/st on columns 448.. while the MMAs run
};

__global__ void __launch_bounds__(160, 1) bench_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, dummy_bar;
  __shared__ volatile int done_flag;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (c.rot_d == 7) {  // hash -> two bf16 values with random sign / mantissa, exponent in [2^-3, 2)
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
      const uint32_t lo = ((h & 0x8000u) | ((0x7c + ((h >> 7) & 3)) << 7) | (h & 0x7f)) & 0xffffu;
      const uint32_t hi = (((h >> 16) & 0x8000u) | ((0x7c + ((h >> 23) & 3)) << 7) | ((h >> 16) & 0x7f)) & 0xffffu;
      v = lo | (hi << 16);
    } else if (c.rot_d == 8) {
      v = 0x3f803f80u;  // all ones
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&dummy_bar)));
    done_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if (warp == 0) {
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 96 * 1024;
    const uint32_t idesc = make_idesc(c.n);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      const uint64_t ad0 = make_desc(a_base, c.a_lbo, c.a_sbo, c.layout);
      const uint64_t bd0 = make_desc(b_base, c.b_lbo, c.b_sbo, c.layout);
      const int nst = c.iters / 18;
      uint32_t w = 0;
      for (int st = 0; st < nst; ++st) {
        const uint32_t d = tmem + (c.win ? 32 * w : 0);
        if (c.win && ++w == 13) w = 0;
#pragma unroll
        for (int t = 0; t < 18; ++t) {
          uint32_t dd = d;
          umma(dd, ad0 + (uint64_t)(t & 7) * (c.a_step >> 4), bd0 + (uint64_t)(t & 7) * (c.b_step >> 4), idesc, 1u);
        }
        if (c.commit && c.commit < 1000) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy_bar)) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    t1 = clock64();
    long long tt = __shfl_sync(0xffffffffu, t0, 0);  // elected lane is lane 0 in practice; take max below anyway
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - (t0 ? t0 : tt);
    done_flag = 1;
  } else if (c.epi) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c.b_sbo * 0 + (uint32_t)c.a_step * 0 + (uint32_t)(c.commit >= 1000 ? c.commit - 1000 : 448);
    while (!done_flag) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v[0] & 0u) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (c.epi > 1) __nanosleep(c.epi);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Named {
    const char* name;
    Cfg c;
  } cfgs[] = {
      // name, {n, layout, a_lbo, a_sbo, b_lbo, b_sbo, a_step, b_step, data(0 zeros, 7 random, 8 ones), iters, win, commit, epi}
      {"N=96  zeros", {96, 0, 2944, 160, 1536, 128, 16, 6144, 0, 9216, 18, 18, 0}},
      {"N=96  ones", {96, 0, 2944, 160, 1536, 128, 16, 6144, 8, 9216, 18, 18, 0}},
      {"N=96  random bf16", {96, 0, 2944, 160, 1536, 128, 16, 6144, 7, 9216, 18, 18, 0}},
      {"N=256 zeros", {256, 0, 2048, 128, 4096, 128, 4096, 8192, 0, 9216, 0, 0, 0}},
      {"N=256 random bf16", {256, 0, 2048, 128, 4096, 128, 4096, 8192, 7, 9216, 0, 0, 0}},
      {"N=32  random bf16", {32, 0, 2944, 160, 1536, 128, 16, 6144, 7, 9216, 18, 18, 0}},
      {"N=64  random bf16", {64, 0, 2944, 160, 1536, 128, 16, 6144, 7, 9216, 18, 18, 0}},
      {"N=128 random bf16", {128, 0, 2944, 160, 2048, 128, 16, 8192, 7, 9216, 18, 18, 0}},
  };
  for (auto& nc : cfgs) {
    for (int grid : {148}) {
      bench_kernel<<<grid, 160, 200 * 1024>>>(nc.c, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%-45s grid=%3d ERROR %s\n", nc.name, grid, cudaGetErrorString(e));
        return 1;
      }
      long long h[148];
      cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const double cyc = (double)mx / nc.c.iters;
      printf("%-45s grid=%3d  %7.1f cyc/MMA  (math floor %5.1f)  eff %.2f\n", nc.name, grid, cyc, 128.0 * nc.c.n / 256.0, (128.0 * nc.c.n / 256.0) / cyc);
    }
  }
  return 0;
}
