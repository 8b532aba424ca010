// HBM-bound helpers: BatchNorm statistics / apply / backward, embedding gather + scatter, weight-norm, dropout/ReLU
// backward, speaker-style reparameterisation, GRU-input assembly, Adam, Philox RNG, FGD statistics.
// All loops are grid-stride with coalesced (vectorised where alignment allows) accesses.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";
void tg_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* tg_last_error(void) { return g_err; }
extern "C" int tg_version(void) { return 100; }

static int g_sms = 0, g_smem = 0;
static void query_dev() {
  int dev = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&g_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (g_sms <= 0) g_sms = 148;
}
int tg_num_sms() { if (!g_sms) query_dev(); return g_sms; }
int tg_max_smem_optin() { if (!g_sms) query_dev(); return g_smem; }
extern "C" int tg_device_info(int* out2) { out2[0] = tg_num_sms(); out2[1] = tg_max_smem_optin(); return 0; }

static inline int ew_blocks(long long n, int per_block = 256) {
  long long b = (n + per_block - 1) / per_block;
  long long cap = (long long)tg_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
#define GRID_STRIDE(i, n) for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

namespace {

// ------------------------------------------------------------------ column statistics (BatchNorm)
// block = 256 threads viewed as (256/CP) row lanes x CP column lanes, CP = pow2 >= C (<= 256).
template <bool BWD>
__global__ void __launch_bounds__(256) col_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, int ld,
                                                         long long M, int C, int CP, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, float slope,
                                                         double* __restrict__ sums) {
  __shared__ double s0[256], s1[256];
  const int c = threadIdx.x % CP, rl = threadIdx.x / CP, RL = 256 / CP;
  double a0 = 0.0, a1 = 0.0;
  if (c < C) {
    float mu = 0.f, rs = 0.f, sc = 0.f, sh = 0.f;
    if (BWD) { mu = mean[c]; rs = rstd[c]; sc = scale[c]; sh = shift[c]; }
    float f0 = 0.f, f1 = 0.f;
    int cnt = 0;
    for (long long r = (long long)blockIdx.x * RL + rl; r < M; r += (long long)gridDim.x * RL) {
      const float v = x[r * ld + c];
      if (BWD) {
        const float z = v * sc + sh;
        const float dz = dy[r * C + c] * (z >= 0.f ? 1.f : slope);
        f0 += dz; f1 += dz * (v - mu) * rs;
      } else {
        f0 += v; f1 += v * v;
      }
      if (++cnt == 64) { a0 += f0; a1 += f1; f0 = f1 = 0.f; cnt = 0; }
    }
    a0 += f0; a1 += f1;
  }
  s0[threadIdx.x] = a0; s1[threadIdx.x] = a1;
  __syncthreads();
  if (rl == 0 && c < C) {
    for (int k = 1; k < RL; ++k) { a0 += s0[k * CP + c]; a1 += s1[k * CP + c]; }
    atomicAdd(sums + c, a0);
    atomicAdd(sums + C + c, a1);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long M, int C, float eps, float momentum, int n_updates,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                                   float* running_var, long long* nbt, float* mean, float* rstd, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += n_updates;
  if (c >= C) return;
  const double mu = sums[c] / (double)M;
  double var = sums[C + c] / (double)M - mu * mu;
  if (var < 0) var = 0;
  const float rs = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean[c] = (float)mu; rstd[c] = rs;
  scale[c] = g * rs; shift[c] = b - (float)mu * g * rs;
  if (running_mean) {
    const float unbiased = (float)(var * ((double)M / (double)(M > 1 ? M - 1 : 1)));
    float rm = running_mean[c], rv = running_var[c];
    for (int i = 0; i < n_updates; ++i) {
      rm = (1.f - momentum) * rm + momentum * (float)mu;
      rv = (1.f - momentum) * rv + momentum * unbiased;
    }
    running_mean[c] = rm; running_var[c] = rv;
  }
}

__global__ void affine_lrelu_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int C,
                                    const float* __restrict__ scale, const float* __restrict__ shift, float slope) {
  GRID_STRIDE(i, n) {
    const int c = (int)(i % C);
    const float v = x[i] * scale[c] + shift[c];
    y[i] = v >= 0.f ? v : v * slope;
  }
}

__global__ void bn_bwd_apply_kernel(const float* dy, const float* __restrict__ x, float* dx, long long M,
                                    int C, const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ scale, const float* __restrict__ shift, float slope,
                                    const float* __restrict__ gamma, const double* __restrict__ sums, float* dgamma, float* dbeta) {
  const long long n = M * C;
  const double invM = 1.0 / (double)M;
  GRID_STRIDE(i, n) {
    const int c = (int)(i % C);
    const float v = x[i];
    const float z = v * scale[c] + shift[c];
    const float dz = dy[i] * (z >= 0.f ? 1.f : slope);
    const float xh = (v - mean[c]) * rstd[c];
    const float g = gamma ? gamma[c] : 1.f;
    dx[i] = g * rstd[c] * (dz - (float)(sums[c] * invM) - xh * (float)(sums[C + c] * invM));
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dgamma) dgamma[c] += (float)sums[C + c];
      if (dbeta) dbeta[c] += (float)sums[c];
    }
  }
}

// ---- float4 variants of the BatchNorm passes for power-of-two channel counts (WavEncoder: 16 / 32 / 64 channels over up to 1M rows).
// A thread's grid stride (gridDim * 256 * 4 elements) is a multiple of C, so it keeps the SAME four channels for its whole loop: the
// per-channel constants live in registers and every memory instruction is a 16-byte access (the scalar kernels above reach 31-47 % of
// the HBM peak, limited by instruction issue: one 4-byte access + an integer modulo per element).
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <bool BWD>
__global__ void __launch_bounds__(256) col_reduce_v4_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long n4, int C,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ scale, const float* __restrict__ shift, float slope,
                                                            double* __restrict__ sums) {
  __shared__ double sh[256][8];
  const long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
  const int c0 = (int)((i0 * 4) % C);
  float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu, sc = mu, sf = mu;
  if (BWD) { mu = ldg4(mean + c0); rs = ldg4(rstd + c0); sc = ldg4(scale + c0); sf = ldg4(shift + c0); }
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)gridDim.x * 256;
  // batches of 4 independent 16-byte loads per array (the loads of a batch are issued before the first one is consumed); fp32 partial sums
  // are flushed into the fp64 accumulators every 16 batches
  long long i = i0;
  while (i < n4) {
#pragma unroll 1
    for (int batch = 0; batch < 16 && i < n4; ++batch) {
      float4 v[4], g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long iu = i + u * stride;
        v[u] = iu < n4 ? ldg4(x + 4 * iu) : make_float4(BWD ? mu.x : 0.f, BWD ? mu.y : 0.f, BWD ? mu.z : 0.f, BWD ? mu.w : 0.f);
        if (BWD) g[u] = iu < n4 ? ldg4(dy + 4 * iu) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (BWD) {
          const float d0 = g[u].x * (v[u].x * sc.x + sf.x >= 0.f ? 1.f : slope), d1 = g[u].y * (v[u].y * sc.y + sf.y >= 0.f ? 1.f : slope);
          const float d2 = g[u].z * (v[u].z * sc.z + sf.z >= 0.f ? 1.f : slope), d3 = g[u].w * (v[u].w * sc.w + sf.w >= 0.f ? 1.f : slope);
          f[0] += d0; f[1] += d1; f[2] += d2; f[3] += d3;
          f[4] += d0 * (v[u].x - mu.x) * rs.x; f[5] += d1 * (v[u].y - mu.y) * rs.y; f[6] += d2 * (v[u].z - mu.z) * rs.z; f[7] += d3 * (v[u].w - mu.w) * rs.w;
        } else {
          f[0] += v[u].x; f[1] += v[u].y; f[2] += v[u].z; f[3] += v[u].w;
          f[4] += v[u].x * v[u].x; f[5] += v[u].y * v[u].y; f[6] += v[u].z * v[u].z; f[7] += v[u].w * v[u].w;
        }
      }
      i += 4 * stride;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] += f[k]; f[k] = 0.f; }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sh[threadIdx.x][k] = a[k] + (double)f[k];
  __syncthreads();
  // threads tid, tid + C/4, tid + 2C/4, ... hold the same four channels
  const int lanes = C / 4;
  for (int o = threadIdx.x; o < lanes * 8; o += 256) {
    const int l = o >> 3, k = o & 7;
    double t = 0.0;
    for (int q = l; q < 256; q += lanes) t += sh[q][k];
    const int c = (int)(((long long)blockIdx.x * 256 + l) * 4 % C) + (k & 3);
    atomicAdd(sums + (k < 4 ? 0 : C) + c, t);
  }
}

__global__ void __launch_bounds__(256) affine_lrelu_v4_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4, int C,
                                                              const float* __restrict__ scale, const float* __restrict__ shift, float slope) {
  const long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
  const int c0 = (int)((i0 * 4) % C);
  const float4 sc = ldg4(scale + c0), sf = ldg4(shift + c0);
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = i0; i < n4; i += stride) {
    const float4 v = ldg4(x + 4 * i);
    float4 o;
    o.x = v.x * sc.x + sf.x; o.y = v.y * sc.y + sf.y; o.z = v.z * sc.z + sf.z; o.w = v.w * sc.w + sf.w;
    o.x = o.x >= 0.f ? o.x : o.x * slope; o.y = o.y >= 0.f ? o.y : o.y * slope;
    o.z = o.z >= 0.f ? o.z : o.z * slope; o.w = o.w >= 0.f ? o.w : o.w * slope;
    *reinterpret_cast<float4*>(y + 4 * i) = o;
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_v4_kernel(const float* dy, const float* __restrict__ x, float* dx, long long n4, long long M,
                                                              int C, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ scale, const float* __restrict__ shift, float slope,
                                                              const float* __restrict__ gamma, const double* __restrict__ sums, float* dgamma,
                                                              float* dbeta) {
  const long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
  const int c0 = (int)((i0 * 4) % C);
  const double invM = 1.0 / (double)M;
  const float4 mu = ldg4(mean + c0), rs = ldg4(rstd + c0), sc = ldg4(scale + c0), sf = ldg4(shift + c0);
  const float4 g = gamma ? ldg4(gamma + c0) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float m0x = (float)(sums[c0] * invM), m0y = (float)(sums[c0 + 1] * invM), m0z = (float)(sums[c0 + 2] * invM), m0w = (float)(sums[c0 + 3] * invM);
  const float m1x = (float)(sums[C + c0] * invM), m1y = (float)(sums[C + c0 + 1] * invM), m1z = (float)(sums[C + c0 + 2] * invM),
              m1w = (float)(sums[C + c0 + 3] * invM);
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = i0; i < n4; i += stride) {
    const float4 v = ldg4(x + 4 * i);
    const float4 d = *reinterpret_cast<const float4*>(dy + 4 * i);         // dx may alias dy: plain load
    float4 o;
    o.x = g.x * rs.x * (d.x * (v.x * sc.x + sf.x >= 0.f ? 1.f : slope) - m0x - (v.x - mu.x) * rs.x * m1x);
    o.y = g.y * rs.y * (d.y * (v.y * sc.y + sf.y >= 0.f ? 1.f : slope) - m0y - (v.y - mu.y) * rs.y * m1y);
    o.z = g.z * rs.z * (d.z * (v.z * sc.z + sf.z >= 0.f ? 1.f : slope) - m0z - (v.z - mu.z) * rs.z * m1z);
    o.w = g.w * rs.w * (d.w * (v.w * sc.w + sf.w >= 0.f ? 1.f : slope) - m0w - (v.w - mu.w) * rs.w * m1w);
    *reinterpret_cast<float4*>(dx + 4 * i) = o;
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dgamma) dgamma[c] += (float)sums[C + c];
      if (dbeta) dbeta[c] += (float)sums[c];
    }
  }
}

// the float4 variants apply to: contiguous rows, C a power of two in [4, 256], 16-byte aligned pointers
static inline bool v4_ok(int C, int ld, const void* a, const void* b = nullptr, const void* c = nullptr) {
  return ld == C && C >= 4 && C <= 256 && (C & (C - 1)) == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}
static inline int v4_blocks(long long n4) {
  long long b = (n4 + 255) / 256, cap = (long long)tg_num_sms() * 8;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

// ------------------------------------------------------------------ embedding
__global__ void embedding_gather_kernel(const float* __restrict__ table, const long long* __restrict__ idx, int idx_mod,
                                        const float* __restrict__ mask, float* __restrict__ out, long long M, int E) {
  const long long n = M * E;
  GRID_STRIDE(i, n) {
    const long long m = i / E;
    const int e = (int)(i - m * E);
    const long long row = idx[idx_mod > 0 ? m % idx_mod : m];
    float v = table[row * E + e];
    if (mask) v *= mask[i];
    out[i] = v;
  }
}
__global__ void embedding_scatter_kernel(const float* __restrict__ dout, const long long* __restrict__ idx,
                                         const float* __restrict__ mask, float* dtable, long long M, int E) {
  const long long n = M * E;
  GRID_STRIDE(i, n) {
    const long long m = i / E;
    const int e = (int)(i - m * E);
    float v = dout[i];
    if (mask) v *= mask[i];
    if (v != 0.f) atomicAdd(dtable + idx[m] * E + e, v);
  }
}

// ------------------------------------------------------------------ weight norm: one warp per output row
// v [N, Cin, taps] (nn.Conv1d layout).  Effective weight is written TAP-MAJOR: w[tap][n][c] (each tap a K-contiguous
// [N, Cin] GEMM operand) and optionally transposed per tap: wT[tap][c][n] (the data-gradient operand).
__global__ void weight_norm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ w,
                                       float* __restrict__ wT, float* __restrict__ inv_norm, int N, int Cin, int taps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const int K = Cin * taps;
  const float* vr = v + (long long)row * K;
  float ss = 0.f;
  for (int k = lane; k < K; k += 32) ss += vr[k] * vr[k];
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss);
  const float s = g[row] * inv;
  for (int k = lane; k < K; k += 32) {
    const int c = k / taps, j = k - c * taps;
    const float val = vr[k] * s;
    w[((long long)j * N + row) * Cin + c] = val;
    if (wT) wT[((long long)j * Cin + c) * N + row] = val;
  }
  if (lane == 0) inv_norm[row] = inv;
}
// one 128-thread block per output row (was one warp per row: 19 dependent load round trips per pass, 12-27 us for a 300 x 600 filter on
// 38 CTAs; the TCN's eight filters sit on the tail of the iteration)
__global__ void __launch_bounds__(128) weight_norm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ g,
                                                              const float* __restrict__ inv_norm, float* dv, float* dg, int N, int Cin, int taps) {
  __shared__ float red[4];
  const int row = blockIdx.x, tid = threadIdx.x;
  const int K = Cin * taps;
  const float* vr = v + (long long)row * K;
  float dot = 0.f;
#pragma unroll 4
  for (int k = tid; k < K; k += 128) {
    const int c = k / taps, j = k - c * taps;
    dot = fmaf(dw[((long long)j * N + row) * Cin + c], vr[k], dot);
  }
  dot = warp_sum(dot);
  if ((tid & 31) == 0) red[tid >> 5] = dot;
  __syncthreads();
  dot = (red[0] + red[1]) + (red[2] + red[3]);
  const float inv = inv_norm[row], gg = g[row];
  // w = g v / |v| : dg = dot/|v| ; dv = g/|v| * (dw - v * dot / |v|^2)
  const float c1 = gg * inv, c2 = dot * inv * inv;
#pragma unroll 4
  for (int k = tid; k < K; k += 128) {
    const int c = k / taps, j = k - c * taps;
    dv[(long long)row * K + k] += c1 * (dw[((long long)j * N + row) * Cin + c] - vr[k] * c2);
  }
  if (tid == 0) dg[row] += dot * inv;
}

// ------------------------------------------------------------------ small elementwise
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n) {
  GRID_STRIDE(i, n) o[i] = a[i] * b[i];
}
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n, int relu) {
  GRID_STRIDE(i, n) {
    const float v = a[i] + b[i];
    o[i] = (relu && v < 0.f) ? 0.f : v;
  }
}
__global__ void bn_eval_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rm,
                                    const float* __restrict__ rv, float eps, const float* __restrict__ cb, float* __restrict__ scale,
                                    float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = (gamma ? gamma[c] : 1.f) / sqrtf(rv[c] + eps);
  scale[c] = s;
  shift[c] = (beta ? beta[c] : 0.f) + ((cb ? cb[c] : 0.f) - rm[c]) * s;
}
__global__ void relu_mask_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ mask,
                                     float* __restrict__ dx, long long n) {
  GRID_STRIDE(i, n) {
    float v = y[i] > 0.f ? dy[i] : 0.f;
    if (mask) v *= mask[i];
    dx[i] = v;
  }
}
// float4 variants of the small elementwise kernels on the TextEncoderTCN / head chains (n % 4 == 0, 16-byte aligned): one 16-byte access
// per array per thread instead of four scalar ones
__global__ void relu_mask_bwd_v4_kernel(const float4* __restrict__ dy, const float4* __restrict__ y, const float4* __restrict__ mask,
                                        float4* __restrict__ dx, long long n4) {
  GRID_STRIDE(i, n4) {
    const float4 g = dy[i], a = y[i];
    float4 v = make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f, a.w > 0.f ? g.w : 0.f);
    if (mask) { const float4 m = mask[i]; v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w; }
    dx[i] = v;
  }
}
// TemporalBlock backward head when the block output was produced by conv2's epilogue (xo = relu(y2 + x), y2 = relu(conv2 + b) * m2 never
// stored): dpre = dxo * (xo > 0), dc2 = dpre * m2 * (y2 > 0) with y2 > 0 read off as xo - x > 0 (exact whenever dpre != 0: there xo = y2 + x)
__device__ __forceinline__ void tcn_res_bwd_one(float g, float o, float xi, float m, float& dp, float& dc) {
  dp = o > 0.f ? g : 0.f;
  dc = (o - xi > 0.f) ? dp * m : 0.f;
}
__global__ void tcn_res_bwd_kernel(const float* __restrict__ dxo, const float* __restrict__ xo, const float* __restrict__ x,
                                   const float* __restrict__ mask, float* __restrict__ dpre, float* __restrict__ dc2, long long n) {
  GRID_STRIDE(i, n) {
    float dp, dc;
    tcn_res_bwd_one(dxo[i], xo[i], x[i], mask ? mask[i] : 1.f, dp, dc);
    dpre[i] = dp; dc2[i] = dc;
  }
}
__global__ void tcn_res_bwd_v4_kernel(const float4* __restrict__ dxo, const float4* __restrict__ xo, const float4* __restrict__ x,
                                      const float4* __restrict__ mask, float4* __restrict__ dpre, float4* __restrict__ dc2, long long n4) {
  GRID_STRIDE(i, n4) {
    const float4 g = dxo[i], o = xo[i], xi = x[i];
    const float4 m = mask ? mask[i] : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 dp, dc;
    tcn_res_bwd_one(g.x, o.x, xi.x, m.x, dp.x, dc.x); tcn_res_bwd_one(g.y, o.y, xi.y, m.y, dp.y, dc.y);
    tcn_res_bwd_one(g.z, o.z, xi.z, m.z, dp.z, dc.z); tcn_res_bwd_one(g.w, o.w, xi.w, m.w, dp.w, dc.w);
    dpre[i] = dp; dc2[i] = dc;
  }
}
__global__ void sum_halves_v4_kernel(const float4* __restrict__ x, float4* __restrict__ o, long long M, int H4) {
  const long long n = M * H4;
  GRID_STRIDE(i, n) {
    const long long m = i / H4; const int h = (int)(i - m * H4);
    const float4 a = x[m * 2 * H4 + h], b = x[m * 2 * H4 + H4 + h];
    o[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}
__global__ void dup_halves_v4_kernel(const float4* __restrict__ d, float4* __restrict__ dx, long long M, int H4) {
  const long long n = M * H4;
  GRID_STRIDE(i, n) {
    const long long m = i / H4; const int h = (int)(i - m * H4);
    const float4 v = d[i];
    dx[m * 2 * H4 + h] = v; dx[m * 2 * H4 + H4 + h] = v;
  }
}
__global__ void sum_halves_kernel(const float* __restrict__ x, float* __restrict__ o, long long M, int H) {
  const long long n = M * H;
  GRID_STRIDE(i, n) {
    const long long m = i / H; const int h = (int)(i - m * H);
    o[i] = x[m * 2 * H + h] + x[m * 2 * H + H + h];
  }
}
__global__ void dup_halves_kernel(const float* __restrict__ d, float* __restrict__ dx, long long M, int H) {
  const long long n = M * H;
  GRID_STRIDE(i, n) {
    const long long m = i / H; const int h = (int)(i - m * H);
    const float v = d[i];
    dx[m * 2 * H + h] = v; dx[m * 2 * H + H + h] = v;
  }
}
__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                                   float* __restrict__ z, long long n) {
  GRID_STRIDE(i, n) z[i] = mu[i] + eps[i] * expf(0.5f * lv[i]);
}
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ lv, const float* __restrict__ eps,
                                   float* dmu, float* dlv, long long n) {
  GRID_STRIDE(i, n) {
    dmu[i] += dz[i];
    dlv[i] += dz[i] * eps[i] * 0.5f * expf(0.5f * lv[i]);
  }
}
__global__ void make_pre_seq_kernel(const float* __restrict__ target, float* __restrict__ pre, int B, int T, int D, int n_pre) {
  const long long n = (long long)B * T * (D + 1);
  GRID_STRIDE(i, n) {
    const long long m = i / (D + 1); const int d = (int)(i - m * (D + 1));
    const int t = (int)(m % T);
    float v = 0.f;
    if (t < n_pre) v = d < D ? target[m * D + d] : 1.f;
    pre[i] = v;
  }
}
__global__ void gru_input_concat_kernel(const float* __restrict__ pre, const float* __restrict__ audio, const float* __restrict__ text,
                                        const float* __restrict__ z, float* __restrict__ out, int B, int Ba, int T, int Dp, int Da,
                                        int Dt, int Dz) {
  const int D = Dp + Da + Dt + Dz;
  const long long n = (long long)B * T * D;
  GRID_STRIDE(i, n) {
    const long long m = i / D; int d = (int)(i - m * D);
    const int b = (int)(m / T), t = (int)(m - (long long)b * T);
    float v;
    if (d < Dp) v = pre[((long long)(b % Ba) * T + t) * Dp + d];
    else if ((d -= Dp) < Da) v = audio[((long long)(b % Ba) * T + t) * Da + d];
    else if ((d -= Da) < Dt) v = text[m * Dt + d];
    else v = z[(long long)b * Dz + (d - Dt)];
    out[i] = v;
  }
}
__global__ void gru_input_split_bwd_kernel(const float* __restrict__ din, float* __restrict__ daudio, float* __restrict__ dtext,
                                           float* __restrict__ dz, int B, int T, int Dp, int Da, int Dt, int Dz) {
  const int D = Dp + Da + Dt + Dz;
  const long long na = (long long)B * T * Da, nt = (long long)B * T * Dt, nz = (long long)B * Dz;
  GRID_STRIDE(i, na + nt + nz) {
    if (i < na) {
      const long long m = i / Da; const int d = (int)(i - m * Da);
      daudio[i] = din[m * D + Dp + d];
    } else if (i < na + nt) {
      const long long k = i - na; const long long m = k / Dt; const int d = (int)(k - m * Dt);
      dtext[k] = din[m * D + Dp + Da + d];
    } else {
      const long long k = i - na - nt; const int b = (int)(k / Dz), d = (int)(k - (long long)b * Dz);
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += din[((long long)b * T + t) * D + Dp + Da + Dt + d];
      dz[k] = s;
    }
  }
}

// ------------------------------------------------------------------ Adam
__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 long long n, float lr, float b1, float b2, float eps, float gscale,
                                 const long long* __restrict__ step_dev) {
  const double step = (double)(*step_dev);
  const float bc1 = (float)(1.0 - pow((double)b1, step));
  const float bc2s = (float)sqrt(1.0 - pow((double)b2, step));
  const float step_size = lr / bc1;
  const long long n4 = n >> 2;
  GRID_STRIDE(i, n4) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define ADAM1(P, G, Mm, V)                                   \
    { const float gr = G * gscale;                           \
      Mm = b1 * Mm + (1.f - b1) * gr;                        \
      V = b2 * V + (1.f - b2) * gr * gr;                     \
      P -= step_size * Mm / (sqrtf(V) / bc2s + eps); }
    ADAM1(pp.x, gg.x, mm.x, vv.x) ADAM1(pp.y, gg.y, mm.y, vv.y) ADAM1(pp.z, gg.z, mm.z, vv.z) ADAM1(pp.w, gg.w, mm.w, vv.w)
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float P = p[i], Mm = m[i], V = v[i]; const float G = g[i];
    ADAM1(P, G, Mm, V)
    p[i] = P; m[i] = Mm; v[i] = V;
  }
}
__global__ void increment_kernel(long long* x, long long by) { *x += by; }

// ------------------------------------------------------------------ Philox RNG
__global__ void philox_normal_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                                     const long long* __restrict__ offset_dev, int stream_id) {
  const unsigned long long off = (unsigned long long)(offset_dev ? *offset_dev : 0);
  const long long n4 = (n + 3) >> 2;
  GRID_STRIDE(i, n4) {
    Philox ph(seed, ((unsigned long long)stream_id << 40) + (unsigned long long)i, off);
    const uint4 r = ph.next();
    float u[4] = {u32_to_unit(r.x), u32_to_unit(r.y), u32_to_unit(r.z), u32_to_unit(r.w)};
    float o[4];
    // Box-Muller on (u0,u1), (u2,u3); 1-u in (0,1]
    float r0 = sqrtf(-2.f * logf(1.f - u[0])), r1 = sqrtf(-2.f * logf(1.f - u[2]));
    float s0, c0, s1, c1;
    sincosf(6.2831853071795864f * u[1], &s0, &c0);
    sincosf(6.2831853071795864f * u[3], &s1, &c1);
    o[0] = r0 * c0; o[1] = r0 * s0; o[2] = r1 * c1; o[3] = r1 * s1;
    for (int k = 0; k < 4; ++k) if (i * 4 + k < n) out[i * 4 + k] = o[k];
  }
}
__global__ void philox_dropout_kernel(float* __restrict__ out, long long n, float p, unsigned long long seed,
                                      const long long* __restrict__ offset_dev, int stream_id) {
  const unsigned long long off = (unsigned long long)(offset_dev ? *offset_dev : 0);
  const float keep = 1.f / (1.f - p);
  const long long n4 = (n + 3) >> 2;
  GRID_STRIDE(i, n4) {
    Philox ph(seed, ((unsigned long long)stream_id << 40) + (unsigned long long)i, off);
    const uint4 r = ph.next();
    const float u[4] = {u32_to_unit(r.x), u32_to_unit(r.y), u32_to_unit(r.z), u32_to_unit(r.w)};
    if (i * 4 + 3 < n && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
      reinterpret_cast<float4*>(out)[i] = make_float4(u[0] >= p ? keep : 0.f, u[1] >= p ? keep : 0.f, u[2] >= p ? keep : 0.f,
                                                      u[3] >= p ? keep : 0.f);
    } else {
      for (int k = 0; k < 4; ++k) if (i * 4 + k < n) out[i * 4 + k] = u[k] >= p ? keep : 0.f;
    }
  }
}
// one CTA, n <= 2048: sort (key,idx) pairs with a bitonic network
__global__ void __launch_bounds__(1024) philox_randperm_kernel(long long* __restrict__ out, int n, unsigned long long seed,
                                                               const long long* __restrict__ offset_dev, int stream_id) {
  __shared__ unsigned long long keys[2048];
  const unsigned long long off = (unsigned long long)(offset_dev ? *offset_dev : 0);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
    if (i < n) {
      Philox ph(seed, ((unsigned long long)stream_id << 40) + (unsigned long long)i, off);
      const uint4 r = ph.next();
      keys[i] = ((unsigned long long)r.x << 32) | ((unsigned long long)(r.y & 0xFFFFF000u)) | (unsigned long long)i;
    } else keys[i] = ~0ull;
  }
  __syncthreads();
  for (int k = 2; k <= 2048; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const unsigned long long a = keys[i], b = keys[ixj];
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = (long long)(keys[i] & 0xFFFull);
}
__global__ void gather_i64_kernel(const long long* __restrict__ src, const long long* __restrict__ idx, long long* __restrict__ out, int n) {
  GRID_STRIDE(i, n) out[i] = src[idx[i]];
}

// ------------------------------------------------------------------ FGD sufficient statistics
// block: 256 threads; each block reduces a chunk of rows of a [n,F] matrix (F <= 64) into fp64 sum / outer-product sums
__global__ void __launch_bounds__(256) feature_stats_kernel(const float* __restrict__ feat, long long n, int F, double* acc) {
  extern __shared__ float rows[];   // [32][F]
  const int npair = F * F;
  double s2[16];                    // up to 16 (i,j) pairs per thread: F*F <= 4096
  double s1 = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s2[k] = 0.0;
  for (long long r0 = (long long)blockIdx.x * 32; r0 < n; r0 += (long long)gridDim.x * 32) {
    const int nr = (int)min((long long)32, n - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * F; i += 256) rows[i] = feat[r0 * F + i];
    __syncthreads();
    for (int r = 0; r < nr; ++r) {
      const float* x = rows + r * F;
      if (threadIdx.x < F) s1 += (double)x[threadIdx.x];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int pidx = threadIdx.x + 256 * k;
        if (pidx < npair) s2[k] += (double)x[pidx / F] * (double)x[pidx % F];
      }
    }
  }
  if (threadIdx.x < F) atomicAdd(acc + 1 + threadIdx.x, s1);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int pidx = threadIdx.x + 256 * k;
    if (pidx < npair) atomicAdd(acc + 1 + F + pidx, s2[k]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(acc, (double)n);
}
__global__ void l1_dist_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, double* acc) {
  __shared__ double red[8];
  double s = 0.0;
  GRID_STRIDE(i, n) s += (double)fabsf(a[i] - b[i]);
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) t += red[k];
    atomicAdd(acc, t);
  }
}

}  // namespace

static int next_pow2(int c) { int p = 1; while (p < c) p <<= 1; return p; }

extern "C" int tg_col_stats_f64(const float* x, int ld, long long M, int C, double* sums, tg_stream stream) {
  TG_REQUIRE(x && sums && C > 0 && C <= 256 && M > 0, "tg_col_stats_f64");
  if (v4_ok(C, ld, x) && M * C >= 4096) {
    const long long n4 = M * C / 4;
    col_reduce_v4_kernel<false><<<v4_blocks(n4), 256, 0, (cudaStream_t)stream>>>(x, nullptr, n4, C, nullptr, nullptr, nullptr, nullptr, 0.f, sums);
    TG_CHECK_LAUNCH("tg_col_stats_f64");
    return 0;
  }
  const int CP = next_pow2(C), RL = 256 / CP;
  long long blocks = (M + RL * 16 - 1) / (RL * 16);
  if (blocks > tg_num_sms() * 8) blocks = tg_num_sms() * 8;
  if (blocks < 1) blocks = 1;
  col_reduce_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, nullptr, ld, M, C, CP, nullptr, nullptr, nullptr, nullptr, 0.f, sums);
  TG_CHECK_LAUNCH("tg_col_stats_f64");
  return 0;
}
extern "C" int tg_bn_finalize(const double* sums, long long M, int C, float eps, float momentum, int n_updates,
                              const float* gamma, const float* beta, float* running_mean, float* running_var,
                              long long* nbt, float* mean, float* rstd, float* scale, float* shift, tg_stream stream) {
  TG_REQUIRE(sums && mean && rstd && scale && shift && C > 0, "tg_bn_finalize");
  bn_finalize_kernel<<<tg_ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, M, C, eps, momentum, n_updates, gamma, beta,
                                                                        running_mean, running_var, nbt, mean, rstd, scale, shift);
  TG_CHECK_LAUNCH("tg_bn_finalize");
  return 0;
}
extern "C" int tg_affine_lrelu(const float* x, float* y, long long M, int C, const float* scale, const float* shift, float slope,
                               tg_stream stream) {
  TG_REQUIRE(x && y && scale && shift, "tg_affine_lrelu");
  if (v4_ok(C, C, x, y, scale) && (((uintptr_t)shift) & 15) == 0 && M * C >= 4096) {
    const long long n4 = M * C / 4;
    affine_lrelu_v4_kernel<<<v4_blocks(n4), 256, 0, (cudaStream_t)stream>>>(x, y, n4, C, scale, shift, slope);
    TG_CHECK_LAUNCH("tg_affine_lrelu");
    return 0;
  }
  affine_lrelu_kernel<<<ew_blocks(M * C), 256, 0, (cudaStream_t)stream>>>(x, y, M * C, C, scale, shift, slope);
  TG_CHECK_LAUNCH("tg_affine_lrelu");
  return 0;
}
extern "C" int tg_bn_bwd_reduce(const float* dy, const float* x, long long M, int C, const float* mean, const float* rstd,
                                const float* scale, const float* shift, float slope, double* sums, tg_stream stream) {
  TG_REQUIRE(dy && x && sums && C > 0 && C <= 256, "tg_bn_bwd_reduce");
  if (v4_ok(C, C, x, dy, mean) && M * C >= 4096 && ((((uintptr_t)rstd | (uintptr_t)scale | (uintptr_t)shift) & 15) == 0)) {
    const long long n4 = M * C / 4;
    col_reduce_v4_kernel<true><<<v4_blocks(n4), 256, 0, (cudaStream_t)stream>>>(x, dy, n4, C, mean, rstd, scale, shift, slope, sums);
    TG_CHECK_LAUNCH("tg_bn_bwd_reduce");
    return 0;
  }
  const int CP = next_pow2(C), RL = 256 / CP;
  long long blocks = (M + RL * 16 - 1) / (RL * 16);
  if (blocks > tg_num_sms() * 8) blocks = tg_num_sms() * 8;
  if (blocks < 1) blocks = 1;
  col_reduce_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, dy, C, M, C, CP, mean, rstd, scale, shift, slope, sums);
  TG_CHECK_LAUNCH("tg_bn_bwd_reduce");
  return 0;
}
extern "C" int tg_bn_bwd_apply(const float* dy, const float* x, float* dx, long long M, int C, const float* mean, const float* rstd,
                               const float* scale, const float* shift, float slope, const float* gamma, const double* sums,
                               float* dgamma, float* dbeta, tg_stream stream) {
  TG_REQUIRE(dy && x && dx && sums, "tg_bn_bwd_apply");
  if (v4_ok(C, C, x, dy, dx) && M * C >= 4096 && mean && rstd && scale && shift &&
      ((((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)scale | (uintptr_t)shift | (uintptr_t)gamma) & 15) == 0)) {
    const long long n4 = M * C / 4;
    bn_bwd_apply_v4_kernel<<<v4_blocks(n4), 256, 0, (cudaStream_t)stream>>>(dy, x, dx, n4, M, C, mean, rstd, scale, shift, slope, gamma, sums,
                                                                            dgamma, dbeta);
    TG_CHECK_LAUNCH("tg_bn_bwd_apply");
    return 0;
  }
  bn_bwd_apply_kernel<<<ew_blocks(M * C), 256, 0, (cudaStream_t)stream>>>(dy, x, dx, M, C, mean, rstd, scale, shift, slope, gamma, sums,
                                                                       dgamma, dbeta);
  TG_CHECK_LAUNCH("tg_bn_bwd_apply");
  return 0;
}
extern "C" int tg_embedding_gather(const float* table, const long long* idx, int idx_mod, const float* mask, float* out,
                                   long long M, int E, tg_stream stream) {
  TG_REQUIRE(table && idx && out, "tg_embedding_gather");
  embedding_gather_kernel<<<ew_blocks(M * E), 256, 0, (cudaStream_t)stream>>>(table, idx, idx_mod, mask, out, M, E);
  TG_CHECK_LAUNCH("tg_embedding_gather");
  return 0;
}
extern "C" int tg_embedding_scatter_add(const float* dout, const long long* idx, const float* mask, float* dtable, long long M, int E,
                                        tg_stream stream) {
  TG_REQUIRE(dout && idx && dtable, "tg_embedding_scatter_add");
  embedding_scatter_kernel<<<ew_blocks(M * E), 256, 0, (cudaStream_t)stream>>>(dout, idx, mask, dtable, M, E);
  TG_CHECK_LAUNCH("tg_embedding_scatter_add");
  return 0;
}
extern "C" int tg_weight_norm_fwd(const float* v, const float* g, float* w, float* wT, float* inv_norm, int N, int Cin, int taps,
                                  tg_stream stream) {
  TG_REQUIRE(v && g && w && inv_norm && taps > 0, "tg_weight_norm_fwd");
  weight_norm_fwd_kernel<<<tg_ceil_div(N, 8), 256, 0, (cudaStream_t)stream>>>(v, g, w, wT, inv_norm, N, Cin, taps);
  TG_CHECK_LAUNCH("tg_weight_norm_fwd");
  return 0;
}
extern "C" int tg_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dv, float* dg, int N,
                                  int Cin, int taps, tg_stream stream) {
  TG_REQUIRE(dw && v && g && inv_norm && dv && dg && taps > 0, "tg_weight_norm_bwd");
  weight_norm_bwd_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(dw, v, g, inv_norm, dv, dg, N, Cin, taps);
  TG_CHECK_LAUNCH("tg_weight_norm_bwd");
  return 0;
}
extern "C" int tg_mul(const float* a, const float* b, float* out, long long n, tg_stream stream) {
  mul_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  TG_CHECK_LAUNCH("tg_mul"); return 0;
}
extern "C" int tg_bn_eval_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                               const float* conv_bias, float* scale, float* shift, int C, tg_stream stream) {
  TG_REQUIRE(running_mean && running_var && scale && shift && C > 0, "tg_bn_eval_fold");
  bn_eval_fold_kernel<<<tg_ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, eps, conv_bias, scale,
                                                                         shift, C);
  TG_CHECK_LAUNCH("tg_bn_eval_fold"); return 0;
}
extern "C" int tg_add(const float* a, const float* b, float* out, long long n, int relu, tg_stream stream) {
  add_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, out, n, relu);
  TG_CHECK_LAUNCH("tg_add"); return 0;
}
extern "C" int tg_relu_mask_bwd(const float* dy, const float* y, const float* mask, float* dx, long long n, tg_stream stream) {
  if ((n & 3) == 0 && (((uintptr_t)dy | (uintptr_t)y | (uintptr_t)mask | (uintptr_t)dx) & 15) == 0)
    relu_mask_bwd_v4_kernel<<<ew_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y),
                                                                               reinterpret_cast<const float4*>(mask), reinterpret_cast<float4*>(dx), n / 4);
  else
    relu_mask_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(dy, y, mask, dx, n);
  TG_CHECK_LAUNCH("tg_relu_mask_bwd"); return 0;
}
extern "C" int tg_tcn_res_bwd(const float* dxo, const float* xo, const float* x, const float* mask, float* dpre, float* dc2, long long n,
                              tg_stream stream) {
  TG_REQUIRE(dxo && xo && x && dpre && dc2 && n > 0, "tg_tcn_res_bwd");
  if ((n & 3) == 0 && (((uintptr_t)dxo | (uintptr_t)xo | (uintptr_t)x | (uintptr_t)mask | (uintptr_t)dpre | (uintptr_t)dc2) & 15) == 0)
    tcn_res_bwd_v4_kernel<<<ew_blocks(n / 4), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(dxo), reinterpret_cast<const float4*>(xo), reinterpret_cast<const float4*>(x),
        reinterpret_cast<const float4*>(mask), reinterpret_cast<float4*>(dpre), reinterpret_cast<float4*>(dc2), n / 4);
  else
    tcn_res_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(dxo, xo, x, mask, dpre, dc2, n);
  TG_CHECK_LAUNCH("tg_tcn_res_bwd"); return 0;
}
extern "C" int tg_sum_halves(const float* x, float* out, long long M, int H, tg_stream stream) {
  if ((H & 3) == 0 && (((uintptr_t)x | (uintptr_t)out) & 15) == 0)
    sum_halves_v4_kernel<<<ew_blocks(M * H / 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), M, H / 4);
  else
    sum_halves_kernel<<<ew_blocks(M * H), 256, 0, (cudaStream_t)stream>>>(x, out, M, H);
  TG_CHECK_LAUNCH("tg_sum_halves"); return 0;
}
extern "C" int tg_dup_halves(const float* d, float* dx, long long M, int H, tg_stream stream) {
  if ((H & 3) == 0 && (((uintptr_t)d | (uintptr_t)dx) & 15) == 0)
    dup_halves_v4_kernel<<<ew_blocks(M * H / 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(d), reinterpret_cast<float4*>(dx), M, H / 4);
  else
    dup_halves_kernel<<<ew_blocks(M * H), 256, 0, (cudaStream_t)stream>>>(d, dx, M, H);
  TG_CHECK_LAUNCH("tg_dup_halves"); return 0;
}
extern "C" int tg_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z, long long n, tg_stream stream) {
  reparam_fwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(mu, logvar, eps, z, n);
  TG_CHECK_LAUNCH("tg_reparam_fwd"); return 0;
}
extern "C" int tg_reparam_bwd(const float* dz, const float* logvar, const float* eps, float* dmu, float* dlogvar, long long n,
                              tg_stream stream) {
  reparam_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(dz, logvar, eps, dmu, dlogvar, n);
  TG_CHECK_LAUNCH("tg_reparam_bwd"); return 0;
}
extern "C" int tg_make_pre_seq(const float* target, float* pre, int B, int T, int D, int n_pre, tg_stream stream) {
  TG_REQUIRE(target && pre, "tg_make_pre_seq");
  make_pre_seq_kernel<<<ew_blocks((long long)B * T * (D + 1)), 256, 0, (cudaStream_t)stream>>>(target, pre, B, T, D, n_pre);
  TG_CHECK_LAUNCH("tg_make_pre_seq"); return 0;
}
extern "C" int tg_gru_input_concat(const float* pre, const float* audio, const float* text, const float* z, float* out, int B, int Ba,
                                   int T, int Dp, int Da, int Dt, int Dz, tg_stream stream) {
  TG_REQUIRE(out && B > 0 && (Da == 0 || Ba > 0), "tg_gru_input_concat");
  gru_input_concat_kernel<<<ew_blocks((long long)B * T * (Dp + Da + Dt + Dz)), 256, 0, (cudaStream_t)stream>>>(pre, audio, text, z, out, B, Ba,
                                                                                                          T, Dp, Da, Dt, Dz);
  TG_CHECK_LAUNCH("tg_gru_input_concat"); return 0;
}
extern "C" int tg_gru_input_split_bwd(const float* din, float* daudio, float* dtext, float* dz, int B, int T, int Dp, int Da, int Dt,
                                      int Dz, tg_stream stream) {
  const long long n = (long long)B * T * (Da + Dt) + (long long)B * Dz;
  gru_input_split_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(din, daudio, dtext, dz, B, T, Dp, Da, Dt, Dz);
  TG_CHECK_LAUNCH("tg_gru_input_split_bwd"); return 0;
}
extern "C" int tg_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                            float grad_scale, const long long* step_dev, tg_stream stream) {
  TG_REQUIRE(p && g && m && v && step_dev, "tg_adam_flat");
  TG_REQUIRE(((uintptr_t)p & 15) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)m & 15) == 0 && ((uintptr_t)v & 15) == 0, "tg_adam_flat");
  adam_flat_kernel<<<ew_blocks(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, b1, b2, eps, grad_scale, step_dev);
  TG_CHECK_LAUNCH("tg_adam_flat"); return 0;
}
extern "C" int tg_increment_i64(long long* x, long long by, tg_stream stream) {
  increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(x, by);
  TG_CHECK_LAUNCH("tg_increment_i64"); return 0;
}
extern "C" int tg_philox_normal(float* out, long long n, unsigned long long seed, const long long* offset_dev, int stream_id,
                                tg_stream stream) {
  philox_normal_kernel<<<ew_blocks(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset_dev, stream_id);
  TG_CHECK_LAUNCH("tg_philox_normal"); return 0;
}
extern "C" int tg_philox_dropout_mask(float* out, long long n, float p, unsigned long long seed, const long long* offset_dev,
                                      int stream_id, tg_stream stream) {
  philox_dropout_kernel<<<ew_blocks(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(out, n, p, seed, offset_dev, stream_id);
  TG_CHECK_LAUNCH("tg_philox_dropout_mask"); return 0;
}
extern "C" int tg_philox_randperm(long long* out, int n, unsigned long long seed, const long long* offset_dev, int stream_id,
                                  tg_stream stream) {
  TG_REQUIRE(n > 0 && n <= 2048, "tg_philox_randperm");
  philox_randperm_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(out, n, seed, offset_dev, stream_id);
  TG_CHECK_LAUNCH("tg_philox_randperm"); return 0;
}
extern "C" int tg_gather_i64(const long long* src, const long long* idx, long long* out, int n, tg_stream stream) {
  gather_i64_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(src, idx, out, n);
  TG_CHECK_LAUNCH("tg_gather_i64"); return 0;
}
extern "C" int tg_feature_stats_f64(const float* feat, long long n, int F, double* acc, tg_stream stream) {
  TG_REQUIRE(feat && acc && F > 0 && F <= 64 && n > 0, "tg_feature_stats_f64");
  long long blocks = (n + 31) / 32;
  if (blocks > tg_num_sms() * 2) blocks = tg_num_sms() * 2;
  feature_stats_kernel<<<(int)blocks, 256, 32 * F * sizeof(float), (cudaStream_t)stream>>>(feat, n, F, acc);
  TG_CHECK_LAUNCH("tg_feature_stats_f64"); return 0;
}
extern "C" int tg_l1_dist_f64(const float* a, const float* b, long long n, double* acc, tg_stream stream) {
  l1_dist_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, acc);
  TG_CHECK_LAUNCH("tg_l1_dist_f64"); return 0;
}
