// ConvDiscriminator convolution stack as ONE launch per pass, forward and backward (multimodal_context_net.py:212-220,233-236):
//   Conv1d(27,16,3) -> BatchNorm1d(16) -> LeakyReLU(True) == identity -> Conv1d(16,8,3) -> BatchNorm1d(8) -> identity -> Conv1d(8,8,3)
// on channels-last poses [B,34,27] -> [B,28,8].  Train-mode BatchNorm needs statistics over ALL clips, so the batch (B <= 128) is spread
// over ONE cluster of 8 CTAs (16 clips each, activations in shared memory) and the per-channel sums are all-reduced through distributed
// shared memory: every CTA publishes its partial sums, a hardware cluster barrier, every CTA adds the 8 partials in rank order (so all
// CTAs hold bit-identical statistics).  The whole stack is ~7.5 MFLOP: the per-operator plan it replaces (3 implicit-GEMM launches +
// 2 x [memset + column statistics + finalize] forward, 3 weight-gradient + 3 data-gradient + 2 x [reduce + apply] backward) spent its
// time in launch latency (15-20 us per launch).
// Version 2 (after profiles/r02_ncu_dfused_v1_*.txt: 46 us forward / 138 us backward, all of it dependent-latency chains in 8 CTAs):
// inputs arrive by cp.async in one round trip, the statistics are reduced with warp shuffles + per-warp partials instead of contended
// fp64 shared-memory atomics (CAS loops), every all-reduce owns its buffer (one cluster barrier each instead of two), the convolutions
// keep four independent accumulators per thread, and the weight gradients run as (output, clip-slice) threads with a rolling 3-tap
// window instead of one 512-long dependent chain per tap.
#include <cuda_runtime.h>
#include "common.cuh"

namespace {

constexpr int CL = 8, NT = 512, NW = NT / 32, CPB = 16;   // CTAs per cluster, threads per CTA, warps per CTA, clips per CTA
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
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
// n floats global -> shared: 16-byte asynchronous copies for the aligned body (4-byte ones when the source is not 16-byte aligned: the
// poses may be a view at any clip offset, e.g. pass 2 of a 3-clip sweep), scalar loads for a tail of < 4 floats
__device__ __forceinline__ void stage_flat(float* dst, const float* src, int n) {
  const int n4 = n & ~3;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    for (int i = threadIdx.x * 4; i < n4; i += NT * 4) cp_async16(dst + i, src + i);
  } else {
    for (int i = threadIdx.x; i < n4; i += NT) cp_async4(dst + i, src + i);
  }
  if (threadIdx.x < n - n4) dst[n4 + threadIdx.x] = __ldg(src + n4 + threadIdx.x);
}
// nn.Conv1d filter [N][CIN][3] -> [3][CIN][N] (tap-major, output channel fastest)
template <int CIN, int N>
__device__ __forceinline__ void load_filter(float* ws, const float* w) {
  for (int i = threadIdx.x; i < KW * CIN * N; i += NT) {
    const int j = i / (CIN * N), r = i - j * CIN * N, ci = r / N, n = r - ci * N;
    ws[i] = __ldg(w + (n * CIN + ci) * KW + j);
  }
}
// nn.Conv1d filter [N][CIN][3] -> [3][N][CIN] (tap-major, INPUT channel fastest): the data gradient's threads differ in the input channel
template <int CIN, int N>
__device__ __forceinline__ void load_filter_dg(float* wd, const float* w) {
  for (int i = threadIdx.x; i < KW * CIN * N; i += NT) {
    const int j = i / (CIN * N), r = i - j * CIN * N, n = r / CIN, ci = r - n * CIN;
    wd[i] = __ldg(w + (n * CIN + ci) * KW + j);
  }
}
// sum of `mine[0..n)` over the 8 CTAs of the cluster, in rank order, into tot[0..n) (every CTA gets the same bits).  `mine` must not be
// written again by its owner (a peer may read it at any time up to the kernel's final cluster barrier).
__device__ __forceinline__ void cluster_allreduce(const double* mine, double* tot, int n) {
  cluster_sync_();                                   // every CTA's partial sums are in its `mine`
  for (int i = threadIdx.x; i < n; i += NT) {
    double a = 0.0;
    for (uint32_t r = 0; r < CL; ++r) a += ld_cluster_f64(mine + i, r);
    tot[i] = a;
  }
  __syncthreads();
}
// per-channel sums of a thread's (s0, s1) over the CTA: threads with equal (threadIdx.x % C) own the same channel.  Shuffle inside the
// warp, one partial per warp in shared memory, fp64 sum over the 16 warps -> mine[0..C) (sum) and mine[C..2C) (second sum).
template <int C>
__device__ __forceinline__ void cta_channel_sums(float s0, float s1, float* part, double* mine) {
#pragma unroll
  for (int o = 16; o >= C; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < C) { part[warp * 32 + lane] = s0; part[warp * 32 + C + lane] = s1; }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) a += (double)part[w * 32 + threadIdx.x];
    mine[threadIdx.x] = a;
  }
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
        if (c == 0 && nbt) *nbt += 1;
      }
    } else {
      mu = rm[c]; rs = 1.f / sqrtf(rv[c] + eps);
    }
    const float sc = gamma[c] * rs, sh = beta[c] - mu * gamma[c] * rs;
    s_scale[c] = sc; s_shift[c] = sh;
    if (leader) { st[c] = mu; st[C + c] = rs; st[2 * C + c] = sc; st[3 * C + c] = sh; }
  }
}

// valid k=3 convolution of this CTA's clips from shared memory: in [nclip][TIN][CIN] -> out [nclip][TIN-2][COUT] (shared) and gout
// (global, may be NULL).  thread = (output channel n = tid % COUT, position lane tid / COUT), four positions in flight per thread;
// a window (3 taps x CIN) is CONTIGUOUS in the channels-last layout.  Returns the thread's sum / sum of squares of what it produced.
template <int CIN, int COUT, int TIN>
__device__ __forceinline__ void conv3_part(const float* in, const float* ws, float bias, int nclip, float* out, float* gout, float& s0, float& s1) {
  constexpr int TOUT = TIN - 2, LANES = NT / COUT, WIN = KW * CIN;
  const int n = threadIdx.x % COUT, pl = threadIdx.x / COUT;
  const int npos = nclip * TOUT;
  s0 = 0.f; s1 = 0.f;
  for (int p0 = pl; p0 < npos; p0 += 4 * LANES) {
    const float* xr[4];
    float a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pos = min(p0 + i * LANES, npos - 1);
      const int c = pos / TOUT, t = pos - c * TOUT;
      xr[i] = in + (c * TIN + t) * CIN;
      a[i] = bias;
    }
#pragma unroll 9
    for (int q = 0; q < WIN; ++q) {
      const float w = ws[q * COUT + n];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = fmaf(xr[i][q], w, a[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pos = p0 + i * LANES;
      if (pos < npos) {
        if (out) out[pos * COUT + n] = a[i];
        if (gout) gout[(long long)pos * COUT + n] = a[i];
        s0 += a[i]; s1 += a[i] * a[i];
      }
    }
  }
}
// x[i] = x[i] * scale[c] + shift[c] in place over [npos][C]  (NT % C == 0: a thread keeps its channel)
template <int C>
__device__ __forceinline__ void affine_inplace(float* x, int npos, const float* scale, const float* shift) {
  const float sc = scale[threadIdx.x % C], sh = shift[threadIdx.x % C];
  for (int i = threadIdx.x; i < npos * C; i += NT) x[i] = fmaf(x[i], sc, sh);
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
  float* part = aff + 48;                            // [16 warps][32] per-warp partial sums
  double* mineA = reinterpret_cast<double*>(part + NW * 32);   // [32] this CTA's sums for BatchNorm 1 (read by the peers)
  double* mineB = mineA + 32;                                  // [16] ... BatchNorm 2
  double* tot = mineB + 16;                                    // [32]
  const int tid = threadIdx.x;
  const int rank = (int)cluster_rank_();
  const int c_lo = rank * CPB;
  const int nclip = max(0, min(CPB, p.B - c_lo));
  const bool leader = rank == 0;

  stage_flat(xs, p.x + (long long)c_lo * T0 * C0, nclip * T0 * C0);
  load_filter<C0, C1>(w1s, p.w1);
  load_filter<C1, C2>(w2s, p.w2);
  load_filter<C2, C3>(w3s, p.w3);
  cp_async_wait_all();
  __syncthreads();

  float s0, s1;
  // ---- conv1 (the pre-BatchNorm output goes to global memory for the backward)
  conv3_part<C0, C1, T0>(xs, w1s, __ldg(p.b1 + (tid % C1)), nclip, y0s, p.y0 + (long long)c_lo * T1 * C1, s0, s1);
  if (p.training) {
    cta_channel_sums<C1>(s0, s1, part, mineA);
    cluster_allreduce(mineA, tot, 2 * C1);
  } else {
    __syncthreads();
  }
  bn_affine(tot, C1, (long long)p.B * T1, p.g1, p.be1, p.rm1, p.rv1, p.nbt1, p.st1, aff, aff + 16, p.training, p.eps, p.momentum, leader);
  __syncthreads();
  affine_inplace<C1>(y0s, nclip * T1, aff, aff + 16);          // LeakyReLU(True) has slope 1: the identity
  __syncthreads();
  // ---- conv2
  conv3_part<C1, C2, T1>(y0s, w2s, __ldg(p.b2 + (tid % C2)), nclip, y1s, p.y1 + (long long)c_lo * T2 * C2, s0, s1);
  if (p.training) {
    cta_channel_sums<C2>(s0, s1, part, mineB);
    cluster_allreduce(mineB, tot, 2 * C2);
  } else {
    __syncthreads();
  }
  bn_affine(tot, C2, (long long)p.B * T2, p.g2, p.be2, p.rm2, p.rv2, p.nbt2, p.st2, aff + 32, aff + 40, p.training, p.eps, p.momentum, leader);
  __syncthreads();
  affine_inplace<C2>(y1s, nclip * T2, aff + 32, aff + 40);
  __syncthreads();
  // ---- conv3
  conv3_part<C2, C3, T2>(y1s, w3s, __ldg(p.b3 + (tid % C3)), nclip, nullptr, p.y2 + (long long)c_lo * T3 * C3, s0, s1);
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

// dW[n][ci][j] += sum over the CTA's (clip, t) of dy[c][t][n] * a[c][t + j][ci],  db[n] += sum dy,  a = sc[ci] * src + sh[ci] (the
// BatchNorm in front of the convolution, applied on the fly; sc == NULL: a = src).
// thread = ((n, ci), clip slice): the three taps share a rolling window over t (one load of `a` and one of `dy` per position for three
// FMAs on independent accumulators); the SL slices are summed through shared memory (`red`, SL * COUT*CIN*3 floats) before ONE atomic
// per weight and CTA.
template <int CIN, int COUT, int TIN>
__device__ __forceinline__ void conv_wgrad_part(const float* dys, const float* src, const float* sc, const float* sh, int nclip, float* dW,
                                                float* db, float* red) {
  constexpr int TOUT = TIN - 2, NOUT = COUT * CIN, SL = (NT / NOUT) >= 8 ? 8 : (NT / NOUT) >= 4 ? 4 : (NT / NOUT) >= 2 ? 2 : 1, CPS = CPB / SL;
  const int o = threadIdx.x % NOUT, sl = threadIdx.x / NOUT;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (sl < SL) {
    const int n = o / CIN, ci = o - n * CIN;
    const float s_ = sc ? sc[ci] : 1.f, h_ = sc ? sh[ci] : 0.f;
    for (int c = sl * CPS; c < min(nclip, (sl + 1) * CPS); ++c) {
      const float* sr = src + c * TIN * CIN + ci;
      const float* dr = dys + c * TOUT * COUT + n;
      float x0 = fmaf(sr[0], s_, h_), x1 = fmaf(sr[CIN], s_, h_);
#pragma unroll 4
      for (int t = 0; t < TOUT; ++t) {
        const float x2 = fmaf(sr[(t + 2) * CIN], s_, h_), g = dr[t * COUT];
        a0 = fmaf(g, x0, a0); a1 = fmaf(g, x1, a1); a2 = fmaf(g, x2, a2);
        x0 = x1; x1 = x2;
      }
    }
  }
  if (SL == 1) {
    if (sl < SL) { atomicAdd(dW + o * KW, a0); atomicAdd(dW + o * KW + 1, a1); atomicAdd(dW + o * KW + 2, a2); }
  } else {
    __syncthreads();                                 // `red` may still be in use by the previous caller
    if (sl < SL) { red[(sl * NOUT + o) * KW] = a0; red[(sl * NOUT + o) * KW + 1] = a1; red[(sl * NOUT + o) * KW + 2] = a2; }
    __syncthreads();
    for (int i = threadIdx.x; i < NOUT * KW; i += NT) {
      float a = 0.f;
#pragma unroll
      for (int q = 0; q < SL; ++q) a += red[q * NOUT * KW + i];
      atomicAdd(dW + i, a);
    }
  }
  // bias gradient: one warp per output channel (COUT <= 16 warps)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < COUT) {
    float a = 0.f;
    for (int i = lane; i < nclip * TOUT; i += 32) a += dys[i * COUT + warp];
    a = warp_sum(a);
    if (lane == 0) atomicAdd(db + warp, a);
  }
}
// da[c][u][ci] = sum_{n, j} dy[c][u - j][n] * W[n][ci][j]   (ws: [3][COUT][CIN] copy of W - consecutive threads = consecutive input
// channels read consecutive words; the first version read [3][CIN][COUT] at a 64-byte stride: 16-way bank conflicts, 70 % of the kernel)
template <int CIN, int COUT, int TIN>
__device__ __forceinline__ void conv_dgrad_part(const float* dys, const float* ws, int nclip, float* das) {
  constexpr int TOUT = TIN - 2;
  for (int i = threadIdx.x; i < nclip * TIN * CIN; i += NT) {
    const int ci = i % CIN, u = (i / CIN) % TIN, c = i / (CIN * TIN);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    {
      const int t = u;
      if (t < TOUT) { const float* dr = dys + (c * TOUT + t) * COUT;
#pragma unroll
        for (int n = 0; n < COUT; ++n) a0 = fmaf(dr[n], ws[(0 * COUT + n) * CIN + ci], a0); }
    }
    {
      const int t = u - 1;
      if (t >= 0 && t < TOUT) { const float* dr = dys + (c * TOUT + t) * COUT;
#pragma unroll
        for (int n = 0; n < COUT; ++n) a1 = fmaf(dr[n], ws[(1 * COUT + n) * CIN + ci], a1); }
    }
    {
      const int t = u - 2;
      if (t >= 0) { const float* dr = dys + (c * TOUT + t) * COUT;
#pragma unroll
        for (int n = 0; n < COUT; ++n) a2 = fmaf(dr[n], ws[(2 * COUT + n) * CIN + ci], a2); }
    }
    das[i] = a0 + a1 + a2;
  }
}
// BatchNorm (identity activation) backward in place on d [npos][C]: partial sums -> cluster all-reduce -> dy = g * rstd * (d - S0/M - xhat * S1/M)
template <int C>
__device__ __forceinline__ void bn_bwd_part(float* d, const float* ysrc, int npos, long long M, const float* st, const float* gamma, float* part,
                                            double* mine, double* tot, float* dgamma, float* dbeta, bool leader) {
  __syncthreads();                                   // d complete
  const int c = threadIdx.x % C;
  const float mu = st[c], rs = st[C + c];
  {
    float s0 = 0.f, s1 = 0.f;
    for (int i = threadIdx.x; i < npos * C; i += NT) {          // NT % C == 0: a thread keeps its channel
      const float dz = d[i];
      s0 += dz; s1 += dz * (ysrc[i] - mu) * rs;
    }
    cta_channel_sums<C>(s0, s1, part, mine);
  }
  cluster_allreduce(mine, tot, 2 * C);
  {
    const float g = __ldg(gamma + c);
    const float m0 = (float)(tot[c] / (double)M), m1 = (float)(tot[C + c] / (double)M);
    for (int i = threadIdx.x; i < npos * C; i += NT) d[i] = g * rs * (d[i] - m0 - (ysrc[i] - mu) * rs * m1);
    if (leader && threadIdx.x < C) { dgamma[c] += (float)tot[C + c]; dbeta[c] += (float)tot[c]; }
  }
  __syncthreads();
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) dconv_stack_bwd_kernel(const ConvBwdP p) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                                    // [CPB][34][27]   x, later d x
  float* y0s = xs + CPB * T0 * C0;                   // [CPB][32][16]   y0 (pre-BatchNorm)
  float* y1s = y0s + CPB * T1 * C1;                  // [CPB][30][8]    y1
  float* d2s = y1s + CPB * T2 * C2;                  // [CPB][28][8]    dy2
  float* d1s = d2s + CPB * T3 * C3;                  // [CPB][30][8]    d a1 -> d y1
  float* d0s = d1s + CPB * T2 * C2;                  // [CPB][32][16]   d a0 -> d y0
  float* w1s = d0s + CPB * T1 * C1;                  // [3][16][27]  (data-gradient layout: input channel fastest)
  float* w2s = w1s + KW * C0 * C1;                   // [3][16][8]
  float* w3s = w2s + KW * C1 * C2;                   // [3][8][8]
  float* st1 = w3s + KW * C2 * C3;                   // 64 floats
  float* st2 = st1 + 64;                             // 32 floats
  float* part = st2 + 32;                            // [16][32]
  float* red = part + NW * 32;                       // [8 * 64 * 3] slice partials of the weight gradients (conv2: 4 * 128 * 3)
  double* mineA = reinterpret_cast<double*>(red + 1536);
  double* mineB = mineA + 32;
  double* tot = mineB + 32;
  const int tid = threadIdx.x;
  const int rank = (int)cluster_rank_();
  const int c_lo = rank * CPB;
  const int nclip = max(0, min(CPB, p.B - c_lo));
  const bool leader = rank == 0;

  stage_flat(xs, p.x + (long long)c_lo * T0 * C0, nclip * T0 * C0);
  stage_flat(y0s, p.y0 + (long long)c_lo * T1 * C1, nclip * T1 * C1);
  stage_flat(y1s, p.y1 + (long long)c_lo * T2 * C2, nclip * T2 * C2);
  stage_flat(d2s, p.dy2 + (long long)c_lo * T3 * C3, nclip * T3 * C3);
  load_filter_dg<C0, C1>(w1s, p.w1);
  load_filter_dg<C1, C2>(w2s, p.w2);
  load_filter_dg<C2, C3>(w3s, p.w3);
  if (tid < 64) st1[tid] = __ldg(p.st1 + tid);
  if (tid < 32) st2[tid] = __ldg(p.st2 + tid);
  cp_async_wait_all();
  __syncthreads();

  const long long M1 = (long long)p.B * T1, M2 = (long long)p.B * T2;
  // conv3: weight gradient on a1 = affine2(y1), data gradient -> d a1
  conv_wgrad_part<C2, C3, T2>(d2s, y1s, st2 + 2 * C2, st2 + 3 * C2, nclip, p.dw3, p.db3, red);
  conv_dgrad_part<C2, C3, T2>(d2s, w3s, nclip, d1s);
  bn_bwd_part<C2>(d1s, y1s, nclip * T2, M2, st2, p.g2, part, mineA, tot, p.dg2, p.dbe2, leader);
  // conv2
  conv_wgrad_part<C1, C2, T1>(d1s, y0s, st1 + 2 * C1, st1 + 3 * C1, nclip, p.dw2, p.db2, red);
  conv_dgrad_part<C1, C2, T1>(d1s, w2s, nclip, d0s);
  bn_bwd_part<C1>(d0s, y0s, nclip * T1, M1, st1, p.g1, part, mineB, tot, p.dg1, p.dbe1, leader);
  // conv1
  conv_wgrad_part<C0, C1, T0>(d0s, xs, nullptr, nullptr, nclip, p.dw1, p.db1, red);
  if (p.dx) {
    __syncthreads();                                 // x is no longer needed once its weight gradient is done
    conv_dgrad_part<C0, C1, T0>(d0s, w1s, nclip, xs);
    __syncthreads();
    const int n = nclip * T0 * C0;
    float* gx = p.dx + (long long)c_lo * T0 * C0;
    for (int i = tid; i < n; i += NT) gx[i] = xs[i];
  }
  cluster_sync_();
}

size_t fwd_smem() {
  return (size_t)(CPB * T0 * C0 + CPB * T1 * C1 + CPB * T2 * C2 + KW * C0 * C1 + KW * C1 * C2 + KW * C2 * C3 + 48 + NW * 32) * 4 + (32 + 16 + 32) * 8;
}
size_t bwd_smem() {
  return (size_t)(CPB * T0 * C0 + 2 * CPB * T1 * C1 + 2 * CPB * T2 * C2 + CPB * T3 * C3 + KW * C0 * C1 + KW * C1 * C2 + KW * C2 * C3 + 96 + NW * 32 +
                  1536) * 4 + (32 + 32 + 32) * 8;
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
