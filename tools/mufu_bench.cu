// mufu_bench.cu -- measured MUFU.EX2 issue rate of one B200 (the roofline denominator of disp_regress).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/mufu_bench.cu -o build/mufu_bench && build/mufu_bench
// Every thread runs ILP independent ex2 chains; blocks of 256 threads, `bps` blocks per SM.  Prints exponentials per second for
// pure-MUFU loops and for the disp_regress inner-loop mix (FFMA + EX2 + FADD + FFMA per plane).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool MIX>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  float x[ILP], s = 0.f, w = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = seed * (threadIdx.x + i) * 1e-3f - 1.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      float e;
      if (MIX) {
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(0.25f, x[(i + 1) % ILP], x[i])));
        s += e;
        w = fmaf(e, (float)i, w);
      } else {
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x[i]));
        x[i] = e - 1.5f;
      }
    }
  }
  float r = s + w;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += x[i];
  if (r == 123.456f) out[0] = r;
}

template <int ILP, bool MIX>
void run(const char* name, int bps) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 4096;
  k<ILP, MIX><<<sms * bps, 256>>>(out, 16, 1.f);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  cudaEventRecord(a);
  k<ILP, MIX><<<sms * bps, 256>>>(out, iters, 1.f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double n = (double)sms * bps * 256 * iters * ILP;
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s ILP %d, %d blocks/SM: %.3f ms, %.2f T exp/s, %.2f exp/clk/SM at %d MHz\n", name, ILP, bps, ms, n / ms / 1e9, n / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  run<8, false>("pure MUFU.EX2", 4);
  run<8, false>("pure MUFU.EX2", 8);
  run<16, false>("pure MUFU.EX2", 8);
  run<8, true>("FFMA+EX2+FADD+FFMA mix", 4);
  run<8, true>("FFMA+EX2+FADD+FFMA mix", 8);
  run<16, true>("FFMA+EX2+FADD+FFMA mix", 3);
  return 0;
}
