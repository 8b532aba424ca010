// ConvDiscriminator convolution stack as ONE launch per pass, forward and backward (multimodal_context_net.py:212-220,233-236):
//   Conv1d(27,16,3) -> BatchNorm1d(16) -> LeakyReLU(True) == identity -> Conv1d(16,8,3) -> BatchNorm1d(8) -> identity -> Conv1d(8,8,3)
// on channels-last poses [B,34,27] -> [B,28,8].  Train-mode BatchNorm needs statistics over ALL clips, so the batch (B <= 128) is spread
// over ONE cluster of 8 CTAs (16 clips each, activations in shared memory) and the per-channel sums are all-reduced through distributed
// shared memory: every CTA publishes its partial sums, a hardware cluster barrier, every CTA adds the 8 partials in rank order (so all
// CTAs hold bit-identical statistics).  The whole stack is ~7.5 MFLOP: the per-operator plan it replaces (3 implicit-GEMM launches +
// 2 x [memset + column statistics + finalize] forward, 3 weight-gradient + 3 data-gradient + 2 x [reduce + apply] backward) spent its
// time in launch latency (15-20 us per launch).
#include <cuda_runtime.h>
#include "common.cuh"

namespace {

constexpr int CL = 8, NT = 512, CPB = 16;            // CTAs per cluster, threads per CTA, clips per CTA
constexpr int T0 = 34, C0 = 27, C1 = 16, C2 = 8, C3 = 8, KW = 3;
constexpr int T1 = T0 - 2, T2 = T1 - 2, T3 = T2 - 2;  // 32, 30, 28

__device__ __forceinline__ void cluster_sync_() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank_() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ double ld_cluster_f64(const double* local, uint32_t rank) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sa), "r"(rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
// sum of `mine[0..n)` over the 8 CTAs of the cluster, in rank order, into tot[0..n) (every CTA gets the same bits)
__device__ __forceinline__ void cluster_allreduce(double* mine, double* tot, int n) {
  cluster_sync_();                                   // every CTA's partial sums are in its `mine`
  for (int i = threadIdx.x; i < n; i += NT) {
    double a = 0.0;
    for (uint32_t r = 0; r < CL; ++r) a += ld_cluster_f64(mine + i, r);
    tot[i] = a;
  }
  cluster_sync_();                                   // nobody overwrites `mine` before every peer has read it; tot visible CTA-wide
}

struct ConvFwdP {
  const float* x;                                                   // [B,34,27]
  const float* w1; const float* b1; const float* w2; const float* b2; const float* w3; const float* b3;     // [N,Cin,3] as stored
  const float* g1; const float* be1; float* rm1; float* rv1; long long* nbt1;                              // BatchNorm1d(16)
  const float* g2; const float* be2; float* rm2; float* rv2; long long* nbt2;                              // BatchNorm1d(8)
  float* y0; float* y1; float* y2;                                  // [B*32,16], [B*30,8], [B*28,8]  (y0 / y1: pre-BatchNorm conv outputs)
  float* st1; float* st2;                                           // per BN: mean | rstd | scale | shift  (4 x C floats)
  int B, training; float eps, momentum;
};

// per-channel (scale, shift) of a BatchNorm from batch sums (train) or running statistics (eval); rank 0 publishes and updates the buffers
__device__ __forceinline__ void bn_affine(const double* tot, int C, long long M, const float* gamma, const float* beta, float* rm, float* rv,
                                          long long* nbt, float* st, float* s_scale, float* s_shift, int training, float eps, float momentum,
                                          bool leader) {
  const int c = threadIdx.x;
  if (c < C) {
    float mu, rs;
    if (training) {
      const double m = tot[c] / (double)M;
      double var = tot[C + c] / (double)M - m * m;
      if (var < 0) var = 0;
      mu = (float)m; rs = (float)(1.0 / sqrt(var + (double)eps));
      if (leader) {
        const float unbiased = (float)(var * ((double)M / (double)(M > 1 ? M - 1 : 1)));
        rm[c] = (1.f - momentum) * rm[c] + momentum * mu;
        rv[c] = (1.f - momentum) * rv[c] + momentum * unbiased;
        if (c == 0) *nbt += 1;
      }
    } else {
      mu = rm[c]; rs = 1.f / sqrtf(rv[c] + eps);
    }
    const float sc = gamma[c] * rs, sh = beta[c] - mu * gamma[c] * rs;
    s_scale[c] = sc; s_shift[c] = sh;
    if (leader) { st[c] = mu; st[C + c] = rs; st[2 * C + c] = sc; st[3 * C + c] = sh; }
  }
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) dconv_stack_fwd_kernel(const ConvFwdP p) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                                    // [CPB][34][27]
  float* y0s = xs + CPB * T0 * C0;                   // [CPB][32][16]
  float* y1s = y0s + CPB * T1 * C1;                  // [CPB][30][8]
  float* w1s = y1s + CPB * T2 * C2;                  // [3][27][16]  tap-major, output channel fastest
  float* w2s = w1s + KW * C0 * C1;                   // [3][16][8]
  float* w3s = w2s + KW * C1 * C2;                   // [3][8][8]
  float* aff = w3s + KW * C2 * C3;                   // scale1[16] shift1[16] scale2[8] shift2[8]
  double* mine = reinterpret_cast<double*>(aff + 48);  // [32] partial sums of this CTA
  double* tot = mine + 32;                             // [32]
  const int tid = threadIdx.x;
  const int rank = (int)cluster_rank_();
  const int c_lo = rank * CPB;
  const int nclip = max(0, min(CPB, p.B - c_lo));
  const bool leader = rank == 0;

  for (int i = tid; i < nclip * T0 * C0; i += NT) xs[i] = __ldg(p.x + (long long)c_lo * T0 * C0 + i);
  for (int i = tid; i < KW * C0 * C1; i += NT) { const int j = i / (C0 * C1), r = i - j * C0 * C1, ci = r / C1, n = r - ci * C1; w1s[i] = __ldg(p.w1 + (n * C0 + ci) * KW + j); }
  for (int i = tid; i < KW * C1 * C2; i += NT) { const int j = i / (C1 * C2), r = i - j * C1 * C2, ci = r / C2, n = r - ci * C2; w2s[i] = __ldg(p.w2 + (n * C1 + ci) * KW + j); }
  for (int i = tid; i < KW * C2 * C3; i += NT) { const int j = i / (C2 * C3), r = i - j * C2 * C3, ci = r / C3, n = r - ci * C3; w3s[i] = __ldg(p.w3 + (n * C2 + ci) * KW + j); }
  if (tid < 32) mine[tid] = 0.0;
  __syncthreads();

  // ---- conv1: thread = (output channel n = tid % 16, position lane tid / 16); x reads are broadcasts, weight reads conflict-free
  {
    const int n = tid & (C1 - 1), pl = tid >> 4;
    const float bias = __ldg(p.b1 + n);
    float s0 = 0.f, s1 = 0.f;
    for (int pos = pl; pos < nclip * T1; pos += NT / C1) {
      const int c = pos / T1, t = pos - c * T1;
      const float* xr = xs + (c * T0 + t) * C0;
      float a = bias;
#pragma unroll
      for (int j = 0; j < KW; ++j)
#pragma unroll 9
        for (int ci = 0; ci < C0; ++ci) a = fmaf(xr[j * C0 + ci], w1s[(j * C0 + ci) * C1 + n], a);
      y0s[pos * C1 + n] = a;
      p.y0[((long long)c_lo * T1 + pos) * C1 + n] = a;
      s0 += a; s1 += a * a;
    }
    if (p.training) { atomicAdd(&mine[n], (double)s0); atomicAdd(&mine[C1 + n], (double)s1); }
  }
  __syncthreads();
  if (p.training) cluster_allreduce(mine, tot, 2 * C1);
  bn_affine(tot, C1, (long long)p.B * T1, p.g1, p.be1, p.rm1, p.rv1, p.nbt1, p.st1, aff, aff + 16, p.training, p.eps, p.momentum, leader);
  if (tid < 32) mine[tid] = 0.0;
  __syncthreads();

  // ---- conv2 on a0 = scale1 * y0 + shift1 (LeakyReLU(True) is the identity): thread = (n = tid % 8, position lane tid / 8)
  {
    const int n = tid & (C2 - 1), pl = tid >> 3;
    const float bias = __ldg(p.b2 + n);
    float s0 = 0.f, s1 = 0.f;
    for (int pos = pl; pos < nclip * T2; pos += NT / C2) {
      const int c = pos / T2, t = pos - c * T2;
      const float* yr = y0s + (c * T1 + t) * C1;
      float a = bias;
#pragma unroll
      for (int j = 0; j < KW; ++j)
#pragma unroll
        for (int ci = 0; ci < C1; ++ci) a = fmaf(fmaf(yr[j * C1 + ci], aff[ci], aff[16 + ci]), w2s[(j * C1 + ci) * C2 + n], a);
      y1s[pos * C2 + n] = a;
      p.y1[((long long)c_lo * T2 + pos) * C2 + n] = a;
      s0 += a; s1 += a * a;
    }
    if (p.training) { atomicAdd(&mine[n], (double)s0); atomicAdd(&mine[C2 + n], (double)s1); }
  }
  __syncthreads();
  if (p.training) cluster_allreduce(mine, tot, 2 * C2);
  bn_affine(tot, C2, (long long)p.B * T2, p.g2, p.be2, p.rm2, p.rv2, p.nbt2, p.st2, aff + 32, aff + 40, p.training, p.eps, p.momentum, leader);
  __syncthreads();

  // ---- conv3 on a1 = scale2 * y1 + shift2
  {
    const int n = tid & (C3 - 1), pl = tid >> 3;
    const float bias = __ldg(p.b3 + n);
    for (int pos = pl; pos < nclip * T3; pos += NT / C3) {
      const int c = pos / T3, t = pos - c * T3;
      const float* yr = y1s + (c * T2 + t) * C2;
      float a = bias;
#pragma unroll
      for (int j = 0; j < KW; ++j)
#pragma unroll
        for (int ci = 0; ci < C2; ++ci) a = fmaf(fmaf(yr[j * C2 + ci], aff[32 + ci], aff[40 + ci]), w3s[(j * C2 + ci) * C3 + n], a);
      p.y2[((long long)c_lo * T3 + pos) * C3 + n] = a;
    }
  }
  cluster_sync_();                                   // no CTA exits while a peer may still read its partial sums
}

// =====================================================================================================================
// backward
// =====================================================================================================================
struct ConvBwdP {
  const float* dy2;                                  // [B*28,8] gradient w.r.t. the stack output
  const float* x; const float* y0; const float* y1;  // saved by the forward
  const float* st1; const float* st2;                // mean | rstd | scale | shift
  const float* w1; const float* w2; const float* w3; const float* g1; const float* g2;
  float* dw1; float* db1; float* dw2; float* db2; float* dw3; float* db3; float* dg1; float* dbe1; float* dg2; float* dbe2;   // accumulated
  float* dx;                                         // [B,34,27] or NULL
  int B;
};

// dW[n][ci][j] += sum over the CTA's (clip, t) of dy[c][t][n] * a[c][t + j][ci]  (a = affine(src) if sc != NULL), db[n] += sum dy
template <int CIN, int COUT, int TIN>
__device__ __forceinline__ void conv_wgrad_part(const float* dys, const float* src, const float* sc, const float* sh, int nclip, float* dW, float* db) {
  constexpr int TOUT = TIN - 2;
  for (int o = threadIdx.x; o < COUT * CIN * KW + COUT; o += NT) {
    float a = 0.f;
    if (o < COUT * CIN * KW) {
      const int n = o / (CIN * KW), r = o - n * CIN * KW, ci = r / KW, j = r - ci * KW;
      const float s = sc ? sc[ci] : 1.f, h = sc ? sh[ci] : 0.f;
      for (int c = 0; c < nclip; ++c)
        for (int t = 0; t < TOUT; ++t) a = fmaf(dys[(c * TOUT + t) * COUT + n], fmaf(src[(c * TIN + t + j) * CIN + ci], s, h), a);
      atomicAdd(dW + o, a);
    } else {
      const int n = o - COUT * CIN * KW;
      for (int i = 0; i < nclip * TOUT; ++i) a += dys[i * COUT + n];
      atomicAdd(db + n, a);
    }
  }
}
// da[c][u][ci] = sum_{n, j} dy[c][u - j][n] * W[n][ci][j]   (ws: [3][CIN][COUT] tap-major copy of W)
template <int CIN, int COUT, int TIN>
__device__ __forceinline__ void conv_dgrad_part(const float* dys, const float* ws, int nclip, float* das) {
  constexpr int TOUT = TIN - 2;
  for (int i = threadIdx.x; i < nclip * TIN * CIN; i += NT) {
    const int ci = i % CIN, u = (i / CIN) % TIN, c = i / (CIN * TIN);
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < KW; ++j) {
      const int t = u - j;
      if (t >= 0 && t < TOUT) {
        const float* dr = dys + (c * TOUT + t) * COUT;
#pragma unroll
        for (int n = 0; n < COUT; ++n) a = fmaf(dr[n], ws[(j * CIN + ci) * COUT + n], a);
      }
    }
    das[i] = a;
  }
}
// BatchNorm (identity activation) backward in place on d [npos][C]: partial sums -> cluster all-reduce -> dy = g * rstd * (d - S0/M - xhat * S1/M)
template <int C>
__device__ __forceinline__ void bn_bwd_part(float* d, const float* ysrc, int npos, long long M, const float* st, const float* gamma, double* mine,
                                            double* tot, float* dgamma, float* dbeta, bool leader) {
  __syncthreads();
  if (threadIdx.x < 2 * C) mine[threadIdx.x] = 0.0;
  __syncthreads();
  {
    const int c = threadIdx.x % C;
    const float mu = st[c], rs = st[C + c];
    float s0 = 0.f, s1 = 0.f;
    for (int i = threadIdx.x; i < npos * C; i += NT) {          // NT % C == 0: a thread keeps its channel
      const float dz = d[i];
      s0 += dz; s1 += dz * (ysrc[i] - mu) * rs;
    }
    atomicAdd(&mine[c], (double)s0); atomicAdd(&mine[C + c], (double)s1);
  }
  __syncthreads();
  cluster_allreduce(mine, tot, 2 * C);
  {
    const int c = threadIdx.x % C;
    const float mu = st[c], rs = st[C + c], g = gamma[c];
    const float m0 = (float)(tot[c] / (double)M), m1 = (float)(tot[C + c] / (double)M);
    for (int i = threadIdx.x; i < npos * C; i += NT) d[i] = g * rs * (d[i] - m0 - (ysrc[i] - mu) * rs * m1);
    if (leader && threadIdx.x < C) { dgamma[c] += (float)tot[C + c]; dbeta[c] += (float)tot[c]; }
  }
  __syncthreads();
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) dconv_stack_bwd_kernel(const ConvBwdP p) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                                    // [CPB][34][27]
  float* y0s = xs + CPB * T0 * C0;                   // [CPB][32][16]   y0, later reused? (kept: conv2 weight gradient needs a0)
  float* y1s = y0s + CPB * T1 * C1;                  // [CPB][30][8]
  float* d2s = y1s + CPB * T2 * C2;                  // [CPB][28][8]    dy2
  float* d1s = d2s + CPB * T3 * C3;                  // [CPB][30][8]    d a1 -> d y1
  float* d0s = d1s + CPB * T2 * C2;                  // [CPB][32][16]   d a0 -> d y0
  float* w1s = d0s + CPB * T1 * C1;                  // [3][27][16]
  float* w2s = w1s + KW * C0 * C1;                   // [3][16][8]
  float* w3s = w2s + KW * C1 * C2;                   // [3][8][8]
  float* st1 = w3s + KW * C2 * C3;                   // 64 floats
  float* st2 = st1 + 64;                             // 32 floats
  double* mine = reinterpret_cast<double*>(st2 + 32);
  double* tot = mine + 32;
  const int tid = threadIdx.x;
  const int rank = (int)cluster_rank_();
  const int c_lo = rank * CPB;
  const int nclip = max(0, min(CPB, p.B - c_lo));
  const bool leader = rank == 0;

  for (int i = tid; i < nclip * T0 * C0; i += NT) xs[i] = __ldg(p.x + (long long)c_lo * T0 * C0 + i);
  for (int i = tid; i < nclip * T1 * C1; i += NT) y0s[i] = __ldg(p.y0 + (long long)c_lo * T1 * C1 + i);
  for (int i = tid; i < nclip * T2 * C2; i += NT) y1s[i] = __ldg(p.y1 + (long long)c_lo * T2 * C2 + i);
  for (int i = tid; i < nclip * T3 * C3; i += NT) d2s[i] = __ldg(p.dy2 + (long long)c_lo * T3 * C3 + i);
  for (int i = tid; i < KW * C0 * C1; i += NT) { const int j = i / (C0 * C1), r = i - j * C0 * C1, ci = r / C1, n = r - ci * C1; w1s[i] = __ldg(p.w1 + (n * C0 + ci) * KW + j); }
  for (int i = tid; i < KW * C1 * C2; i += NT) { const int j = i / (C1 * C2), r = i - j * C1 * C2, ci = r / C2, n = r - ci * C2; w2s[i] = __ldg(p.w2 + (n * C1 + ci) * KW + j); }
  for (int i = tid; i < KW * C2 * C3; i += NT) { const int j = i / (C2 * C3), r = i - j * C2 * C3, ci = r / C3, n = r - ci * C3; w3s[i] = __ldg(p.w3 + (n * C2 + ci) * KW + j); }
  if (tid < 64) st1[tid] = __ldg(p.st1 + tid);
  if (tid < 32) st2[tid] = __ldg(p.st2 + tid);
  __syncthreads();

  const long long M1 = (long long)p.B * T1, M2 = (long long)p.B * T2;
  // conv3: weight gradient on a1 = affine2(y1), data gradient -> d a1
  conv_wgrad_part<C2, C3, T2>(d2s, y1s, st2 + 2 * C2, st2 + 3 * C2, nclip, p.dw3, p.db3);
  conv_dgrad_part<C2, C3, T2>(d2s, w3s, nclip, d1s);
  bn_bwd_part<C2>(d1s, y1s, nclip * T2, M2, st2, p.g2, mine, tot, p.dg2, p.dbe2, leader);
  // conv2
  conv_wgrad_part<C1, C2, T1>(d1s, y0s, st1 + 2 * C1, st1 + 3 * C1, nclip, p.dw2, p.db2);
  conv_dgrad_part<C1, C2, T1>(d1s, w2s, nclip, d0s);
  bn_bwd_part<C1>(d0s, y0s, nclip * T1, M1, st1, p.g1, mine, tot, p.dg1, p.dbe1, leader);
  // conv1
  conv_wgrad_part<C0, C1, T0>(d0s, xs, nullptr, nullptr, nclip, p.dw1, p.db1);
  if (p.dx) {
    __syncthreads();
    float* dxs = xs;                                 // x is no longer needed once its weight gradient is done
    __syncthreads();
    conv_dgrad_part<C0, C1, T0>(d0s, w1s, nclip, dxs);
    __syncthreads();
    for (int i = tid; i < nclip * T0 * C0; i += NT) p.dx[(long long)c_lo * T0 * C0 + i] = dxs[i];
  }
  cluster_sync_();
}

size_t fwd_smem() { return (size_t)(CPB * T0 * C0 + CPB * T1 * C1 + CPB * T2 * C2 + KW * C0 * C1 + KW * C1 * C2 + KW * C2 * C3 + 48) * 4 + 64 * 8; }
size_t bwd_smem() {
  return (size_t)(CPB * T0 * C0 + 2 * CPB * T1 * C1 + 2 * CPB * T2 * C2 + CPB * T3 * C3 + KW * C0 * C1 + KW * C1 * C2 + KW * C2 * C3 + 96) * 4 + 64 * 8;
}

}  // namespace

extern "C" int tg_dconv_stack_fwd(const float* x, const float* w1, const float* b1, const float* g1, const float* be1, float* rm1, float* rv1,
                                  long long* nbt1, const float* w2, const float* b2, const float* g2, const float* be2, float* rm2, float* rv2,
                                  long long* nbt2, const float* w3, const float* b3, float* y0, float* y1, float* y2, float* st1, float* st2,
                                  int B, int T, int D, int training, float eps, float momentum, tg_stream stream) {
  TG_REQUIRE(x && w1 && b1 && g1 && be1 && rm1 && rv1 && w2 && b2 && g2 && be2 && rm2 && rv2 && w3 && b3 && y0 && y1 && y2 && st1 && st2,
             "tg_dconv_stack_fwd");
  TG_REQUIRE(B > 0 && B <= CL * CPB && T == T0 && D == C0, "tg_dconv_stack_fwd(shape)");
  ConvFwdP p;
  p.x = x; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3;
  p.g1 = g1; p.be1 = be1; p.rm1 = rm1; p.rv1 = rv1; p.nbt1 = nbt1; p.g2 = g2; p.be2 = be2; p.rm2 = rm2; p.rv2 = rv2; p.nbt2 = nbt2;
  p.y0 = y0; p.y1 = y1; p.y2 = y2; p.st1 = st1; p.st2 = st2; p.B = B; p.training = training; p.eps = eps; p.momentum = momentum;
  const size_t smem = fwd_smem();
  cudaError_t e = cudaFuncSetAttribute(dconv_stack_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("tg_dconv_stack_fwd: smem attr: %s", cudaGetErrorString(e)); return -3; }
  dconv_stack_fwd_kernel<<<CL, NT, smem, (cudaStream_t)stream>>>(p);
  TG_CHECK_LAUNCH("tg_dconv_stack_fwd");
  return 0;
}

extern "C" int tg_dconv_stack_bwd(const float* dy2, const float* x, const float* y0, const float* y1, const float* st1, const float* st2,
                                  const float* w1, const float* w2, const float* w3, const float* g1, const float* g2, float* dw1, float* db1,
                                  float* dw2, float* db2, float* dw3, float* db3, float* dg1, float* dbe1, float* dg2, float* dbe2, float* dx,
                                  int B, int T, int D, tg_stream stream) {
  TG_REQUIRE(dy2 && x && y0 && y1 && st1 && st2 && w1 && w2 && w3 && g1 && g2 && dw1 && db1 && dw2 && db2 && dw3 && db3 && dg1 && dbe1 && dg2 && dbe2,
             "tg_dconv_stack_bwd");
  TG_REQUIRE(B > 0 && B <= CL * CPB && T == T0 && D == C0, "tg_dconv_stack_bwd(shape)");
  ConvBwdP p;
  p.dy2 = dy2; p.x = x; p.y0 = y0; p.y1 = y1; p.st1 = st1; p.st2 = st2; p.w1 = w1; p.w2 = w2; p.w3 = w3; p.g1 = g1; p.g2 = g2;
  p.dw1 = dw1; p.db1 = db1; p.dw2 = dw2; p.db2 = db2; p.dw3 = dw3; p.db3 = db3; p.dg1 = dg1; p.dbe1 = dbe1; p.dg2 = dg2; p.dbe2 = dbe2;
  p.dx = dx; p.B = B;
  const size_t smem = bwd_smem();
  cudaError_t e = cudaFuncSetAttribute(dconv_stack_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("tg_dconv_stack_bwd: smem attr: %s", cudaGetErrorString(e)); return -3; }
  dconv_stack_bwd_kernel<<<CL, NT, smem, (cudaStream_t)stream>>>(p);
  TG_CHECK_LAUNCH("tg_dconv_stack_bwd");
  return 0;
}
