// geometry.cu -- stage-boundary geometry on the GPU (reference: host numpy/numba + grid_sample round trips).
//   a8  disp -> depth, sine rule            save_output_disparity_stage.py:118-133
//   a9  rotateCassini / a11 cassini2Equirec  utils/geometry.py:48-91, 7-45   (constant grid + bilinear, border)
//   a10 depthViewTransWithConf               utils/geometry.py:94-156        (fp64 geometry + serial z-buffer)
// All constant tables/grids (functions of the shape and the camera angles only) are generated on the host in
// numpy exactly as the reference does and uploaded once; the kernels do the per-pixel work.
#include "common.cuh"
using namespace mode;

// ---------------------------------------------------------------------------------------------------------
// a8.  depth = b * sin(pi/2 - phi_r) / sin(phi_r - phi_l),  phi_r = disp*pi/W + phi_l;  disp == 0 -> 1000;
//      > 1000 -> 1000; < 0 -> 0.
// Precision follows the reference AS IT EXECUTES under NumPy >= 2 (the container's 2.3): the masked-array
// product `disp_not_0 * math.pi` promotes to float64 (np.ma wraps the Python scalar in a 0-d float64 array,
// which is not a weak scalar under NEP 50), so the whole triangulation runs in fp64 on the fp32 inputs.
// The kernel writes the fp64 result (consumed un-rounded by the forward warp, as in the reference) and its
// fp32 rounding (what rotateCassini / the fusion loader see: torch.FloatTensor / astype(float32)).
__global__ void disp_to_depth_kernel(const float* __restrict__ disp, const float* __restrict__ phi_l, float* __restrict__ depth,
                                     double* __restrict__ depth64, long long n, int W, float baseline) {
  const double PI = 3.141592653589793;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = disp[i];
    const double pl = (double)__ldg(phi_l + (int)(i % W));
    double out;
    if (d == 0.f) {
      out = 1000.0;
    } else {
      const double pr = __dadd_rn(__ddiv_rn(__dmul_rn((double)d, PI), (double)W), pl);
      const double num = __dmul_rn((double)baseline, sin(__dsub_rn(PI / 2, pr)));
      out = __ddiv_rn(num, sin(__dsub_rn(pr, pl)));
      if (out > 1000.0) out = 1000.0;
      if (out < 0.0) out = 0.0;
    }
    if (depth) depth[i] = (float)out;
    if (depth64) depth64[i] = out;
  }
}

extern "C" int mode_disp_to_depth(const float* disp, const float* phi_l, float* depth, double* depth64, int B, int H, int W, float baseline,
                                  void* stream) {
  MODE_CHECK_ARG(disp && phi_l && (depth || depth64) && B > 0 && H > 0 && W > 0, "disp_to_depth: bad arguments");
  const long long n = (long long)B * H * W;
  const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)kNumSMs * 16);
  disp_to_depth_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(disp, phi_l, depth, depth64, n, W, baseline);
  MODE_CHECK_LAUNCH("disp_to_depth");
  return MODE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// a9/a11.  F.grid_sample(bilinear, align_corners=True, padding_mode='border'), grid shared by all N*C planes.
// ATen arithmetic: ix = ((x+1)/2)*(W-1), clip to [0, W-1]; x0=floor; weights (x1-ix)*(y1-iy) ...; corners outside
// the image contribute 0 (only reachable at the clipped upper edge where their weight is 0).
__global__ void grid_sample_border_kernel(const float* __restrict__ src, const float* __restrict__ grid, float* __restrict__ out,
                                          int NC, int Hs, int Ws, int HoWo) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HoWo) return;
  const float2 g = __ldg(reinterpret_cast<const float2*>(grid) + p);
  float ix = ((g.x + 1.f) / 2.f) * (float)(Ws - 1);
  float iy = ((g.y + 1.f) / 2.f) * (float)(Hs - 1);
  ix = fminf(fmaxf(ix, 0.f), (float)(Ws - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(Hs - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - x0f, wy1 = iy - y0f;
  const float wx0 = (x0f + 1.f) - ix, wy0 = (y0f + 1.f) - iy;
  const float nw = wx0 * wy0, ne = wx1 * wy0, sw = wx0 * wy1, se = wx1 * wy1;
  const bool vx1 = x1 <= Ws - 1, vy1 = y1 <= Hs - 1;
  for (int c = blockIdx.y; c < NC; c += gridDim.y) {
    const float* s = src + (size_t)c * Hs * Ws;
    float acc = 0.f;
    acc += __ldg(s + y0 * Ws + x0) * nw;
    if (vx1) acc += __ldg(s + y0 * Ws + x1) * ne;
    if (vy1) acc += __ldg(s + y1 * Ws + x0) * sw;
    if (vx1 && vy1) acc += __ldg(s + y1 * Ws + x1) * se;
    out[(size_t)c * HoWo + p] = acc;
  }
}

extern "C" int mode_grid_sample_border(const float* src, const float* grid, float* out, int N, int C, int Hs, int Ws, int Ho, int Wo,
                                       void* stream) {
  MODE_CHECK_ARG(src && grid && out && N > 0 && C > 0 && Hs > 0 && Ws > 0 && Ho > 0 && Wo > 0, "grid_sample_border: bad arguments");
  dim3 g(ceil_div((long long)Ho * Wo, 256), std::min(N * C, 65535));
  grid_sample_border_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(src, grid, out, N * C, Hs, Ws, Ho * Wo);
  MODE_CHECK_LAUNCH("grid_sample_border");
  return MODE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// a10.  Forward warp with z-buffer.  Reference semantics (geometry.py:148-156): pixels visited in row-major
// order; a pixel with r_1 > 0 replaces the target iff r_2 (fp64) < view_2[target] (fp32 buffer).  Writing the
// state as (v fp32, conf), a candidate (b fp64, b' = fp32(b)) replaces iff b' < v, or b' == v and b < b'
// (b was rounded up).  Hence the final depth is min b', and the final conf comes from
//     n* = max( first index with b' == vmin , last index with b' == vmin and b < b' ).
// Three deterministic passes reproduce this exactly: atomicMin(depth bits), atomicMin/atomicMax(indices), resolve.
struct RT {
  double r[9];
  double t[3];
};

// back-projection exactly as numpy evaluates geometry.py:122-124: the trig tables are fp32; the products are
// fp32 when the depth map is fp32 and fp64 when it is fp64 (numpy promotion), left to right.
__device__ __forceinline__ void back_project(float r1, float sp, float cp, float st, float ct, double& x1, double& y1, double& z1) {
  const float rc = __fmul_rn(r1, cp);
  x1 = (double)__fmul_rn(r1, sp);
  y1 = (double)__fmul_rn(rc, st);
  z1 = (double)__fmul_rn(rc, ct);
}
__device__ __forceinline__ void back_project(double r1, float sp, float cp, float st, float ct, double& x1, double& y1, double& z1) {
  const double rc = __dmul_rn(r1, (double)cp);
  x1 = __dmul_rn(r1, (double)sp);
  y1 = __dmul_rn(rc, (double)st);
  z1 = __dmul_rn(rc, (double)ct);
}

template <typename T>
__device__ __forceinline__ void warp_target(T r1, float sp, float cp, float st, float ct, const RT& rt, int H, int W, double& r2, int& tgt) {
  double x1, y1, z1;
  back_project(r1, sp, cp, st, ct, x1, y1, z1);
  const double a = x1 - rt.t[0], b = y1 - rt.t[1], c = z1 - rt.t[2];
  // np.matmul(R, X_1 - t) (geometry.py:127) for stacked 3x3 @ 3x1 runs OpenBLAS' FMA micro-kernel: measured on the AVX-512 hosts of this
  // pool (exact-arithmetic emulation, tools: tests/test_oracle_golden.py::test_numpy_matmul_rounding_model) every row is
  //     fma(r2, c, fma(r0, a, r1 * b))          -- product of the MIDDLE column first, then two fused multiply-adds.
  // One ulp in X/Y/Z moves r_2 by an ulp, which decides z-buffer ties between neighbouring source pixels (constant-depth regions),
  // so the same operation order is used here.
  const double X = __fma_rn(rt.r[2], c, __fma_rn(rt.r[0], a, __dmul_rn(rt.r[1], b)));
  const double Y = __fma_rn(rt.r[5], c, __fma_rn(rt.r[3], a, __dmul_rn(rt.r[4], b)));
  const double Z = __fma_rn(rt.r[8], c, __fma_rn(rt.r[6], a, __dmul_rn(rt.r[7], b)));
  r2 = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(X, X), __dmul_rn(Y, Y)), __dmul_rn(Z, Z)));
  const double theta = atan2(Y, Z);
  double q = X / r2;
  q = fmin(fmax(q, -1.0), 1.0);
  const double phi = asin(q);
  const double PI = 3.141592653589793;
  double I = rint((double)H / 2 - (double)H * theta / (2 * PI));
  double J = rint((double)W / 2 - (double)W * phi / PI);
  I = fmin(fmax(I, 0.0), (double)(H - 1));
  J = fmin(fmax(J, 0.0), (double)(W - 1));
  tgt = (int)I * W + (int)J;
}

__global__ void warp_init_kernel(uint32_t* ws, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    ws[i] = 0xFFFFFFFFu;        // min depth bits
    ws[n + i] = 0xFFFFFFFFu;    // first index
    ws[2 * n + i] = 0u;         // last rounded-up index + 1 (0 = none)
  }
}

template <int PASS, typename T>
__global__ void warp_scatter_kernel(const T* __restrict__ depth, const float* __restrict__ sp, const float* __restrict__ cp,
                                    const float* __restrict__ st, const float* __restrict__ ct, RT rt, uint32_t* ws, int H, int W,
                                    long long n) {
  const int HW = H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T r1 = depth[i];
    if (!(r1 > (T)0)) continue;
    const int pix = (int)(i % HW);
    const long long base = i - pix;
    const int h = pix / W, w = pix - h * W;
    double r2;
    int tgt;
    warp_target(r1, __ldg(sp + w), __ldg(cp + w), __ldg(st + h), __ldg(ct + h), rt, H, W, r2, tgt);
    const float r2f = (float)r2;
    if (!(r2 < 100000.0)) continue;  // the fp64 candidate never beats the initial 100000 sentinel (also drops NaN); a candidate just
                                      // below it that ROUNDS to 100000.f still wins, sets the confidence, and is zeroed in the resolve pass
    const uint32_t bits = __float_as_uint(r2f);  // r2 >= 0: unsigned order == float order
    if (PASS == 0) {
      atomicMin(ws + base + tgt, bits);
    } else {
      if (ws[base + tgt] == bits) {
        atomicMin(ws + n + base + tgt, (uint32_t)pix);
        if (r2 < (double)r2f) atomicMax(ws + 2 * n + base + tgt, (uint32_t)pix + 1u);
      }
    }
  }
}

__global__ void warp_resolve_kernel(const float* __restrict__ conf, const uint32_t* __restrict__ ws, float* __restrict__ view2,
                                    float* __restrict__ conf2, int HW, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t bits = ws[i];
    float v = 0.f, c = 0.f;
    if (bits != 0xFFFFFFFFu) {
      v = __uint_as_float(bits);
      if (v == 100000.f) v = 0.f;  // view_2[view_2 == 100000] = 0 (geometry.py:141), before the clip
      if (v > 1000.f) v = 1000.f;
      uint32_t first = ws[n + i], lastup = ws[2 * n + i];
      uint32_t src = first;
      if (lastup != 0u && lastup - 1u > first) src = lastup - 1u;
      c = conf[(i - (i % HW)) + src];
    }
    view2[i] = v;
    conf2[i] = c;
  }
}

extern "C" int mode_depth_view_trans(const float* depth, const double* depth64, const float* conf, const float* sin_phi, const float* cos_phi,
                                     const float* sin_theta, const float* cos_theta, const double* Rt_host, uint32_t* workspace,
                                     float* view2, float* conf2, int B, int H, int W, void* stream) {
  MODE_CHECK_ARG((depth || depth64) && conf && sin_phi && cos_phi && sin_theta && cos_theta && Rt_host && workspace && view2 && conf2,
                 "depth_view_trans: null pointer");
  MODE_CHECK_ARG(B > 0 && H > 0 && W > 0 && (long long)H * W < 2147483647LL, "depth_view_trans: bad shape");
  RT rt;
  for (int i = 0; i < 9; ++i) rt.r[i] = Rt_host[i];
  for (int i = 0; i < 3; ++i) rt.t[i] = Rt_host[9 + i];
  const long long n = (long long)B * H * W;
  const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)kNumSMs * 16);
  cudaStream_t s = (cudaStream_t)stream;
  warp_init_kernel<<<blocks, 256, 0, s>>>(workspace, n);
  MODE_CHECK_LAUNCH("depth_view_trans/init");
  if (depth64) {
    warp_scatter_kernel<0, double><<<blocks, 256, 0, s>>>(depth64, sin_phi, cos_phi, sin_theta, cos_theta, rt, workspace, H, W, n);
    MODE_CHECK_LAUNCH("depth_view_trans/min");
    warp_scatter_kernel<1, double><<<blocks, 256, 0, s>>>(depth64, sin_phi, cos_phi, sin_theta, cos_theta, rt, workspace, H, W, n);
    MODE_CHECK_LAUNCH("depth_view_trans/index");
  } else {
    warp_scatter_kernel<0, float><<<blocks, 256, 0, s>>>(depth, sin_phi, cos_phi, sin_theta, cos_theta, rt, workspace, H, W, n);
    MODE_CHECK_LAUNCH("depth_view_trans/min");
    warp_scatter_kernel<1, float><<<blocks, 256, 0, s>>>(depth, sin_phi, cos_phi, sin_theta, cos_theta, rt, workspace, H, W, n);
    MODE_CHECK_LAUNCH("depth_view_trans/index");
  }
  warp_resolve_kernel<<<blocks, 256, 0, s>>>(conf, workspace, view2, conf2, H * W, n);
  MODE_CHECK_LAUNCH("depth_view_trans/resolve");
  return MODE_OK;
}
