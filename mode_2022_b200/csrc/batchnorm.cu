// batchnorm.cu -- training-mode BatchNorm2d / BatchNorm3d (batch statistics) forward and backward for fp32 NCHW / NCDHW tensors.
//
// Reference: every conv of the stereo stage is followed by nn.BatchNorm2d / nn.BatchNorm3d (models/submodule.py:14-30,
// models/mode_disparity.py:11-46,66-80); train_disparity.py:147-163 runs them in training mode (batch statistics, running-stat
// update with momentum 0.1, unbiased running variance).  On the GPU the reference goes through cuDNN's spatial BN kernels; at the
// training shape (1 pair, 1024x512, D=192: 3.8 GB of normalised activations per step, thirteen of them 201 MB NCDHW tensors) those
// take 49 ms of a 182 ms step -- bn_bw_1C11 runs at ~0.5 TB/s.  The op is a per-channel reduction plus an elementwise pass, i.e.
// HBM-bound:
//   forward : pass 1 reads x (per-channel shifted sums), pass 2 reads x and writes y                     3 x tensor bytes
//   backward: pass 1 reads x, dy (sum dy, sum dy*xhat), pass 2 reads x, dy and writes dx                 5 x tensor bytes
// Layout: x is (N, C, S) contiguous, S = H*W or D*H*W.  A channel's N*S elements are split over `nsplit` blocks; every thread
// accumulates fp32 sums of (x - K) and (x - K)^2 with the per-channel shift K = x[0, c, 0] (|x - K| ~ sigma: no cancellation in
// E[x^2] - E[x]^2), blocks reduce in fp64 into a [C][nsplit] partial array, and a finalise kernel adds the partials in a fixed
// order -- the statistics are bit-reproducible from run to run.
#include "common.cuh"
using namespace mode;

namespace {

constexpr int kBnThreads = 256;

struct BnSplit {
  int nsplit;          // blocks per channel
  long long per;       // elements of the channel's flattened (n, s) range per block (multiple of 4 when S % 4 == 0)
};

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < kBnThreads / 32 ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  }
  return t;  // valid in thread 0
}

// walks the elements [e0, e1) of channel c's flattened (n, s) index space; f(x index) is called with float4 granularity when VEC
template <bool VEC, typename F4, typename F1>
__device__ __forceinline__ void for_range(long long e0, long long e1, long long S, int C, int c, F4 f4, F1 f1) {
  if (VEC) {  // S % 4 == 0 and e0, e1 multiples of 4: a float4 never straddles two images
    for (long long e = e0 + 4LL * threadIdx.x; e < e1; e += 4LL * kBnThreads) {
      const long long n = e / S, s = e - n * S;
      f4(((n * C + c) * S + s));
    }
  } else {
    for (long long e = e0 + threadIdx.x; e < e1; e += kBnThreads) {
      const long long n = e / S, s = e - n * S;
      f1(((n * C + c) * S + s));
    }
  }
}

// ---- forward pass 1: partial[c][split] = {sum (x - K), sum (x - K)^2}
template <bool VEC>
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ x, double2* __restrict__ partial, int C, long long S, long long total, BnSplit sp) {
  __shared__ double red[kBnThreads / 32];
  const int c = blockIdx.x, split = blockIdx.y;
  const float K = __ldg(x + (size_t)c * S);
  const long long e0 = split * sp.per, e1 = min(e0 + sp.per, total);
  float s1 = 0.f, s2 = 0.f;
  for_range<VEC>(
      e0, e1, S, C, c,
      [&](long long i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + i));
        const float a = v.x - K, b = v.y - K, cc = v.z - K, d = v.w - K;
        s1 += (a + b) + (cc + d);
        s2 = fmaf(a, a, fmaf(b, b, fmaf(cc, cc, fmaf(d, d, s2))));
      },
      [&](long long i) {
        const float a = __ldg(x + i) - K;
        s1 += a;
        s2 = fmaf(a, a, s2);
      });
  const double t1 = block_sum((double)s1, red);
  const double t2 = block_sum((double)s2, red);
  if (threadIdx.x == 0) partial[(size_t)c * sp.nsplit + split] = make_double2(t1, t2);
}

// ---- forward finalise: mean, biased variance -> save_mean / save_invstd, running statistics (momentum update, unbiased variance)
__global__ void bn_finalize_kernel(const float* __restrict__ x, const double2* __restrict__ partial, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                   float* __restrict__ batch_var, float* __restrict__ running_mean, float* __restrict__ running_var, int C, long long kstride, long long total, int nsplit,
                                   float eps, float momentum) {
  // one warp per channel: lanes stride over the channel's partials, then a shuffle tree -- a fixed summation order
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int i = lane; i < nsplit; i += 32) {
    const double2 p = partial[(size_t)c * nsplit + i];
    s1 += p.x, s2 += p.y;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, d), s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  if (lane != 0) return;
  const double K = (double)__ldg(x + (size_t)c * kstride);  // the shift the partial sums were taken around: x[0, c, 0]
  const double n = (double)total;
  const double m1 = s1 / n;
  const double mean = K + m1;
  const double var = fmax(s2 / n - m1 * m1, 0.0);
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  const double var_u = total > 1 ? var * n / (n - 1.0) : var;  // unbiased: what the running variance tracks
  if (batch_var) batch_var[c] = (float)var_u;
  if (running_mean) running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
  if (running_var) running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * var_u);
}

// ---- forward pass 2: y = (x - mean) * invstd * gamma + beta, one (n, c) row segment per block
template <bool VEC>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ y, int C, long long S,
                                                              long long seg) {
  const long long row = blockIdx.y;  // n * C + c
  const int c = (int)(row % C);
  const float m = __ldg(mean + c), a = __ldg(invstd + c) * (gamma ? __ldg(gamma + c) : 1.f), b = beta ? __ldg(beta + c) : 0.f;
  const long long s0 = blockIdx.x * seg, s1 = min(s0 + seg, S);
  const float* xr = x + row * S;
  float* yr = y + row * S;
  if (VEC) {
    for (long long s = s0 + 4LL * threadIdx.x; s < s1; s += 4LL * kBnThreads) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xr + s));
      *reinterpret_cast<float4*>(yr + s) = make_float4(fmaf(v.x - m, a, b), fmaf(v.y - m, a, b), fmaf(v.z - m, a, b), fmaf(v.w - m, a, b));
    }
  } else {
    for (long long s = s0 + threadIdx.x; s < s1; s += kBnThreads) yr[s] = fmaf(__ldg(xr + s) - m, a, b);
  }
}

// ---- backward pass 1: partial[c][split] = {sum dy, sum dy * (x - mean)}
template <bool VEC>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                                                                   double2* __restrict__ partial, int C, long long S, long long total, BnSplit sp) {
  __shared__ double red[kBnThreads / 32];
  const int c = blockIdx.x, split = blockIdx.y;
  const float m = __ldg(mean + c);
  const long long e0 = split * sp.per, e1 = min(e0 + sp.per, total);
  float s1 = 0.f, s2 = 0.f;
  for_range<VEC>(
      e0, e1, S, C, c,
      [&](long long i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + i)), g = __ldg(reinterpret_cast<const float4*>(dy + i));
        s1 += (g.x + g.y) + (g.z + g.w);
        s2 = fmaf(g.x, v.x - m, fmaf(g.y, v.y - m, fmaf(g.z, v.z - m, fmaf(g.w, v.w - m, s2))));
      },
      [&](long long i) {
        const float g = __ldg(dy + i);
        s1 += g;
        s2 = fmaf(g, __ldg(x + i) - m, s2);
      });
  const double t1 = block_sum((double)s1, red);
  const double t2 = block_sum((double)s2, red);
  if (threadIdx.x == 0) partial[(size_t)c * sp.nsplit + split] = make_double2(t1, t2);
}

// ---- backward finalise: dbeta = sum dy, dgamma = invstd * sum dy (x - mean); coefficients of dx = ca * dy + cb * (x - mean) + cc
__global__ void bn_bwd_finalize_kernel(const double2* __restrict__ partial, const float* __restrict__ gamma, const float* __restrict__ invstd, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ coef, int C, long long total, int nsplit) {
  // one warp per channel: lanes stride over the channel's partials, then a shuffle tree -- a fixed summation order
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int i = lane; i < nsplit; i += 32) {
    const double2 p = partial[(size_t)c * nsplit + i];
    s1 += p.x, s2 += p.y;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, d), s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  if (lane != 0) return;
  const double is = (double)invstd[c], g = gamma ? (double)gamma[c] : 1.0, n = (double)total;
  if (dbeta) dbeta[c] = (float)s1;
  if (dgamma) dgamma[c] = (float)(s2 * is);
  // dx = g * is * (dy - s1 / n - (x - mean) * is^2 * s2 / n)
  coef[3 * c + 0] = (float)(g * is);
  coef[3 * c + 1] = (float)(-g * is * is * is * s2 / n);
  coef[3 * c + 2] = (float)(-g * is * s1 / n);
}

// ---- backward pass 2: dx = ca * dy + cb * (x - mean) + cc
template <bool VEC>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                                                                  const float* __restrict__ coef, float* __restrict__ dx, int C, long long S, long long seg) {
  const long long row = blockIdx.y;
  const int c = (int)(row % C);
  const float m = __ldg(mean + c), ca = __ldg(coef + 3 * c), cb = __ldg(coef + 3 * c + 1), cc = __ldg(coef + 3 * c + 2);
  const long long s0 = blockIdx.x * seg, s1 = min(s0 + seg, S);
  const float* xr = x + row * S;
  const float* gr = dy + row * S;
  float* dr = dx + row * S;
  if (VEC) {
    for (long long s = s0 + 4LL * threadIdx.x; s < s1; s += 4LL * kBnThreads) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xr + s)), g = __ldg(reinterpret_cast<const float4*>(gr + s));
      *reinterpret_cast<float4*>(dr + s) =
          make_float4(fmaf(ca, g.x, fmaf(cb, v.x - m, cc)), fmaf(ca, g.y, fmaf(cb, v.y - m, cc)), fmaf(ca, g.z, fmaf(cb, v.z - m, cc)), fmaf(ca, g.w, fmaf(cb, v.w - m, cc)));
    }
  } else {
    for (long long s = s0 + threadIdx.x; s < s1; s += kBnThreads) dr[s] = fmaf(ca, __ldg(gr + s), fmaf(cb, __ldg(xr + s) - m, cc));
  }
}

// =============================================================================================================================
// channels-last layout: x is (N, S, C) in memory (torch.channels_last / channels_last_3d), R = N * S rows of C contiguous floats,
// C % 4 == 0 and C <= 256.  A thread owns one float4 column group (the same one on every row it visits), a block a strided set of
// rows; coalesced 16-byte accesses, per-channel shifted sums reduced through shared memory into fp64 partials [block][C].
// =============================================================================================================================
constexpr int kClMaxC = 256;

template <bool BWD>
__global__ void __launch_bounds__(kBnThreads) bn_cl_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                                                                  double2* __restrict__ partial, int C, long long R) {
  __shared__ float red[kBnThreads][9];
  const int G = C >> 2;                       // float4 groups per row
  const int g = threadIdx.x % G, r0 = threadIdx.x / G, rstep = kBnThreads / G;  // G divides 256 for C in {4, 8, ..., 256} powers of two; else the tail threads idle
  const bool live = r0 < rstep;
  // forward: shift K = first row of the tensor; backward: the batch mean
  const float4 K = BWD ? __ldg(reinterpret_cast<const float4*>(mean) + g) : __ldg(reinterpret_cast<const float4*>(x) + g);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    for (long long r = (long long)blockIdx.x * rstep + r0; r < R; r += (long long)gridDim.x * rstep) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C) + g);
      const float a[4] = {v.x - K.x, v.y - K.y, v.z - K.z, v.w - K.w};
      if (BWD) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(dy + r * C) + g);
        const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) s1[i] += dd[i], s2[i] = fmaf(dd[i], a[i], s2[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) s1[i] += a[i], s2[i] = fmaf(a[i], a[i], s2[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) red[threadIdx.x][i] = s1[i], red[threadIdx.x][4 + i] = s2[i];
  __syncthreads();
  if (threadIdx.x < C) {  // thread c sums the rstep threads that own channel c (fixed order)
    const int c = threadIdx.x, gg = c >> 2, i = c & 3;
    double t1 = 0.0, t2 = 0.0;
    for (int q = 0; q < rstep; ++q) t1 += (double)red[q * G + gg][i], t2 += (double)red[q * G + gg][4 + i];
    partial[(size_t)c * gridDim.x + blockIdx.x] = make_double2(t1, t2);
  }
}

// y = (x - mean) * a + b   or   dx = ca * dy + cb * (x - mean) + cc; one float4 group per thread, fixed over the grid-stride loop
template <bool BWD>
__global__ void __launch_bounds__(kBnThreads) bn_cl_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 const float* __restrict__ coef, float* __restrict__ out, int C, long long R) {
  const int G = C >> 2;
  const int g = threadIdx.x % G, r0 = threadIdx.x / G, rstep = kBnThreads / G;
  if (r0 >= rstep) return;
  float m[4], a[4], b[4], cc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = 4 * g + i;
    m[i] = __ldg(mean + c);
    if (BWD) {
      a[i] = __ldg(coef + 3 * c), b[i] = __ldg(coef + 3 * c + 1), cc[i] = __ldg(coef + 3 * c + 2);
    } else {
      a[i] = __ldg(invstd + c) * (gamma ? __ldg(gamma + c) : 1.f), b[i] = beta ? __ldg(beta + c) : 0.f, cc[i] = 0.f;
    }
  }
  for (long long r = (long long)blockIdx.x * rstep + r0; r < R; r += (long long)gridDim.x * rstep) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C) + g);
    float4 o;
    if (BWD) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(dy + r * C) + g);
      o = make_float4(fmaf(a[0], d.x, fmaf(b[0], v.x - m[0], cc[0])), fmaf(a[1], d.y, fmaf(b[1], v.y - m[1], cc[1])), fmaf(a[2], d.z, fmaf(b[2], v.z - m[2], cc[2])),
                      fmaf(a[3], d.w, fmaf(b[3], v.w - m[3], cc[3])));
    } else {
      o = make_float4(fmaf(v.x - m[0], a[0], b[0]), fmaf(v.y - m[1], a[1], b[1]), fmaf(v.z - m[2], a[2], b[2]), fmaf(v.w - m[3], a[3], b[3]));
    }
    *(reinterpret_cast<float4*>(out + r * C) + g) = o;
  }
}

bool cl_supported(int C) { return C >= 4 && C <= kClMaxC && (C & (C - 1)) == 0; }  // power of two: C/4 groups tile the 256 threads
int cl_blocks(long long R, int C) {
  const long long rstep = kBnThreads / (C >> 2);
  return (int)std::max<long long>(1, std::min<long long>(4LL * kNumSMs, (R + rstep - 1) / rstep));
}

BnSplit pick_split(int C, long long total, bool vec) {
  // enough blocks to fill the GPU several times over, at least ~8 k elements per block
  long long want = std::max<long long>(1, (8LL * kNumSMs + C - 1) / C);
  want = std::min<long long>(want, std::max<long long>(1, total / 8192));
  long long per = (total + want - 1) / want;
  if (vec) per = (per + 3) / 4 * 4;
  BnSplit sp;
  sp.per = per;
  sp.nsplit = (int)((total + per - 1) / per);
  return sp;
}

// elementwise passes: grid (segments of a row, rows); segment length a multiple of 4 * threads
void pick_rows(long long rows, long long S, long long& seg, int& nseg) {
  const long long unit = 4LL * kBnThreads;
  long long want = std::max<long long>(1, (16LL * kNumSMs + rows - 1) / rows);  // segments per row
  seg = (S + want - 1) / want;
  seg = std::max(unit, (seg + unit - 1) / unit * unit);
  nseg = (int)((S + seg - 1) / seg);
}

}  // namespace

extern "C" size_t mode_batchnorm_workspace_bytes(int C, long long N, long long S, int channels_last) {
  const int nsplit = channels_last ? cl_blocks(N * S, C) : pick_split(C, N * S, S % 4 == 0).nsplit;
  return (size_t)C * nsplit * sizeof(double2) + (size_t)3 * C * sizeof(float);
}

extern "C" int mode_batchnorm_train_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* save_mean, float* save_invstd, float* batch_var,
                                            float* running_mean, float* running_var, void* workspace, long long N, int C, long long S, int channels_last, float eps, float momentum, void* stream) {
  MODE_CHECK_ARG(x && y && save_mean && save_invstd && workspace, "batchnorm_train_fwd_f32: null pointer");
  MODE_CHECK_ARG(N > 0 && C > 0 && S > 0 && N * C < 2147483647LL, "batchnorm_train_fwd_f32: bad shape");
  const long long total = N * S;
  if (channels_last) {
    MODE_CHECK_ARG(cl_supported(C), "batchnorm_train_fwd_f32: channels-last layout needs a power-of-two C in [4, 256] (got %d)", C);
    MODE_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "batchnorm_train_fwd_f32: channels-last tensors must be 16-byte aligned");
    cudaStream_t cs = (cudaStream_t)stream;
    double2* part = reinterpret_cast<double2*>(workspace);
    const int nb = cl_blocks(total, C);
    bn_cl_reduce_kernel<false><<<nb, kBnThreads, 0, cs>>>(x, nullptr, nullptr, part, C, total);
    MODE_CHECK_LAUNCH("batchnorm_train_fwd_f32 (statistics, channels-last)");
    bn_finalize_kernel<<<ceil_div(C, 4), 128, 0, cs>>>(x, part, save_mean, save_invstd, batch_var, running_mean, running_var, C, 1, total, nb, eps, momentum);
    MODE_CHECK_LAUNCH("batchnorm_train_fwd_f32 (finalise)");
    bn_cl_apply_kernel<false><<<nb, kBnThreads, 0, cs>>>(x, nullptr, gamma, beta, save_mean, save_invstd, nullptr, y, C, total);
    MODE_CHECK_LAUNCH("batchnorm_train_fwd_f32 (normalise, channels-last)");
    return MODE_OK;
  }
  const bool vec = S % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
  const BnSplit sp = pick_split(C, total, S % 4 == 0);
  cudaStream_t s = (cudaStream_t)stream;
  double2* partial = reinterpret_cast<double2*>(workspace);
  dim3 g1(C, sp.nsplit);
  if (vec)
    bn_stats_kernel<true><<<g1, kBnThreads, 0, s>>>(x, partial, C, S, total, sp);
  else
    bn_stats_kernel<false><<<g1, kBnThreads, 0, s>>>(x, partial, C, S, total, sp);
  MODE_CHECK_LAUNCH("batchnorm_train_fwd_f32 (statistics)");
  bn_finalize_kernel<<<ceil_div(C, 4), 128, 0, s>>>(x, partial, save_mean, save_invstd, batch_var, running_mean, running_var, C, S, total, sp.nsplit, eps, momentum);  // K = x[c * S]
  MODE_CHECK_LAUNCH("batchnorm_train_fwd_f32 (finalise)");
  long long seg;
  int nseg;
  pick_rows(N * C, S, seg, nseg);
  dim3 g2(nseg, (unsigned)(N * C));
  if (vec)
    bn_apply_kernel<true><<<g2, kBnThreads, 0, s>>>(x, gamma, beta, save_mean, save_invstd, y, C, S, seg);
  else
    bn_apply_kernel<false><<<g2, kBnThreads, 0, s>>>(x, gamma, beta, save_mean, save_invstd, y, C, S, seg);
  MODE_CHECK_LAUNCH("batchnorm_train_fwd_f32 (normalise)");
  return MODE_OK;
}

extern "C" int mode_batchnorm_train_bwd_f32(const float* x, const float* dy, const float* gamma, const float* save_mean, const float* save_invstd, float* dx,
                                            float* dgamma, float* dbeta, void* workspace, long long N, int C, long long S, int channels_last, void* stream) {
  MODE_CHECK_ARG(x && dy && dx && save_mean && save_invstd && workspace, "batchnorm_train_bwd_f32: null pointer");
  MODE_CHECK_ARG(N > 0 && C > 0 && S > 0 && N * C < 2147483647LL, "batchnorm_train_bwd_f32: bad shape");
  const long long total = N * S;
  if (channels_last) {
    MODE_CHECK_ARG(cl_supported(C), "batchnorm_train_bwd_f32: channels-last layout needs a power-of-two C in [4, 256] (got %d)", C);
    MODE_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
                   "batchnorm_train_bwd_f32: channels-last tensors must be 16-byte aligned");
    cudaStream_t cs = (cudaStream_t)stream;
    const int nb = cl_blocks(total, C);
    double2* part = reinterpret_cast<double2*>(workspace);
    float* cf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + (size_t)C * nb * sizeof(double2));
    bn_cl_reduce_kernel<true><<<nb, kBnThreads, 0, cs>>>(x, dy, save_mean, part, C, total);
    MODE_CHECK_LAUNCH("batchnorm_train_bwd_f32 (reductions, channels-last)");
    bn_bwd_finalize_kernel<<<ceil_div(C, 4), 128, 0, cs>>>(part, gamma, save_invstd, dgamma, dbeta, cf, C, total, nb);
    MODE_CHECK_LAUNCH("batchnorm_train_bwd_f32 (finalise)");
    bn_cl_apply_kernel<true><<<nb, kBnThreads, 0, cs>>>(x, dy, nullptr, nullptr, save_mean, nullptr, cf, dx, C, total);
    MODE_CHECK_LAUNCH("batchnorm_train_bwd_f32 (dx, channels-last)");
    return MODE_OK;
  }
  const bool vec = S % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0;
  const BnSplit sp = pick_split(C, total, S % 4 == 0);
  cudaStream_t s = (cudaStream_t)stream;
  double2* partial = reinterpret_cast<double2*>(workspace);
  float* coef = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + (size_t)C * sp.nsplit * sizeof(double2));
  dim3 g1(C, sp.nsplit);
  if (vec)
    bn_bwd_reduce_kernel<true><<<g1, kBnThreads, 0, s>>>(x, dy, save_mean, partial, C, S, total, sp);
  else
    bn_bwd_reduce_kernel<false><<<g1, kBnThreads, 0, s>>>(x, dy, save_mean, partial, C, S, total, sp);
  MODE_CHECK_LAUNCH("batchnorm_train_bwd_f32 (reductions)");
  bn_bwd_finalize_kernel<<<ceil_div(C, 4), 128, 0, s>>>(partial, gamma, save_invstd, dgamma, dbeta, coef, C, total, sp.nsplit);
  MODE_CHECK_LAUNCH("batchnorm_train_bwd_f32 (finalise)");
  long long seg;
  int nseg;
  pick_rows(N * C, S, seg, nseg);
  dim3 g2(nseg, (unsigned)(N * C));
  if (vec)
    bn_bwd_apply_kernel<true><<<g2, kBnThreads, 0, s>>>(x, dy, save_mean, coef, dx, C, S, seg);
  else
    bn_bwd_apply_kernel<false><<<g2, kBnThreads, 0, s>>>(x, dy, save_mean, coef, dx, C, S, seg);
  MODE_CHECK_LAUNCH("batchnorm_train_bwd_f32 (dx)");
  return MODE_OK;
}
