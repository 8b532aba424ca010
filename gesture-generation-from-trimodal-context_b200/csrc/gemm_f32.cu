// Strict-fp32 implicit-GEMM convolution / linear kernels (CUDA-core FFMA path).
// Forward/dgrad: 128 x BN output tile, 256 threads, 8 x TN register micro-tile, BK = 16, register prefetch of the
// next K slab.  A rows are gathered (tap / stride / dilation / zero padding) straight from the channels-last
// activation, so no im2col, chomp or concat copy ever exists (cf. tcn.py:13 `.contiguous()` in the reference).
#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 16, TM = 8;

template <int BN, int TN>
__global__ void __launch_bounds__(256) conv_gemm_kernel(const tg_conv_gemm_t p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int M = p.B * p.Tout, K = p.taps * p.Cin;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // ---- A loader: thread owns row ar, 8 consecutive k
  const int ar = tid >> 1, ak = (tid & 1) * 8;
  const int am = m0 + ar;
  const bool arow_ok = am < M;
  const int ab = arow_ok ? am / p.Tout : 0;
  const int at = am - ab * p.Tout;
  const int tin0 = at * p.stride - p.pad;
  const long long abstride = p.a_bstride ? p.a_bstride : (long long)p.Tin * p.lda;
  const int asc = p.asc ? p.asc : 1;
  const float* abase = p.A + (long long)ab * abstride;
  const bool has_pro = p.pscale != nullptr;
  float areg[8];
  constexpr int BE = BN * BK / 256;
  float breg[BE];

  auto load_tiles = [&](int k0) {
    int kg = k0 + ak;
    int j = kg / p.Cin;
    int c = kg - j * p.Cin;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = 0.f;
      if (arow_ok && kg + e < K) {
        int tin = tin0 + j * p.dil;
        if (tin >= 0 && tin < p.Tin) {
          v = __ldg(abase + (long long)tin * p.lda + (long long)c * asc);
          if (has_pro) {
            v = v * __ldg(p.pscale + c) + __ldg(p.pshift + c);
            v = v >= 0.f ? v : v * p.pslope;
          }
        }
      }
      areg[e] = v;
      if (++c == p.Cin) { c = 0; ++j; }
    }
#pragma unroll
    for (int e = 0; e < BE; ++e) {
      int i = tid + 256 * e;
      int n = i / BK, kk = i % BK;
      int kg2 = k0 + kk;
      float v = 0.f;
      if (n0 + n < p.N && kg2 < K) {
        int j2 = kg2 / p.Cin;
        int c2 = kg2 - j2 * p.Cin;
        v = __ldg(p.W + (long long)(n0 + n) * p.ldw + (long long)j2 * p.wsj + (long long)c2 * p.wsc);
      }
      breg[e] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int e = 0; e < 8; ++e) As[ak + e][ar] = areg[e];
#pragma unroll
    for (int e = 0; e < BE; ++e) {
      int i = tid + 256 * e;
      Bs[i % BK][i / BK] = breg[e];
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_tiles(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < K) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if constexpr (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
      } else if constexpr (TN == 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      } else {
        const float2 b0 = *reinterpret_cast<const float2*>(&Bs[kk][tx * TN]);
        b[0] = b0.x; b[1] = b0.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
    const int b = m / p.Tout, t = m - b * p.Tout;
    const long long orow = (long long)b * p.ToutFull + (long long)t * p.ostride + p.ooff;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.escale) v *= __ldg(p.escale + n);
      if (p.bias) v += __ldg(p.bias + n);
      v = tg_act(v, p.act1, p.slope1);
      if (p.mask) v *= __ldg(p.mask + orow * p.ldmask + n);
      if (p.residual) v += __ldg(p.residual + orow * p.ldres + n);
      v = tg_act(v, p.act2, 0.f);
      float* dst = p.Y + orow * p.ldc + n;
      if (p.accumulate) v += *dst;
      *dst = v;
    }
  }
}

// ---- weight gradient: dW[n, kw] += sum_m G[m, n] * A[row(m, j(kw)), c(kw)], split over m, atomically reduced.
constexpr int WN = 64, WK = 64, WM = 16;

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const tg_conv_wgrad_t p, int rows_per_split) {
  __shared__ __align__(16) float Gs[WM][WN + 4];
  __shared__ __align__(16) float As[WM][WK + 4];
  __shared__ float bred[4][WN];
  const int tid = threadIdx.x;
  const int M = p.B * p.Tout, KW = p.taps * p.Cin;
  const int kw0 = blockIdx.x * WK, n0 = blockIdx.y * WN;
  const int m_begin = blockIdx.z * rows_per_split;
  const int m_end = min(M, m_begin + rows_per_split);
  const int col = tid & 63, rsub = tid >> 6;   // loader: fixed column, rows rsub + 4e
  const int gn = n0 + col;
  const bool gn_ok = gn < p.N;
  const int akw = kw0 + col;
  const bool akw_ok = akw < KW;
  const int aj = akw_ok ? akw / p.Cin : 0;
  const int ac = akw - aj * p.Cin;
  const bool has_pro = p.pscale != nullptr;
  float psc = 1.f, psh = 0.f;
  if (has_pro && akw_ok) { psc = __ldg(p.pscale + ac); psh = __ldg(p.pshift + ac); }
  const bool do_bias = (p.dbias != nullptr) && blockIdx.x == 0;
  float bsum = 0.f;

  const int ty = tid >> 4, tx = tid & 15;   // compute: n micro ty*4, kw micro tx*4
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float greg[4], areg[4];
  auto load = [&](int mb) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int m = mb + rsub + 4 * e;
      float g = 0.f, a = 0.f;
      if (m < m_end) {
        if (gn_ok) g = __ldg(p.G + (long long)m * p.ldg + gn);
        if (akw_ok) {
          const int b = m / p.Tout, t = m - b * p.Tout;
          const int tin = t * p.stride + aj * p.dil - p.pad;
          if (tin >= 0 && tin < p.Tin) {
            a = __ldg(p.A + ((long long)b * p.Tin + tin) * p.lda + ac);
            if (has_pro) { a = a * psc + psh; a = a >= 0.f ? a : a * p.pslope; }
          }
        }
      }
      greg[e] = g; areg[e] = a;
    }
  };

  load(m_begin);
  for (int mb = m_begin; mb < m_end; mb += WM) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      Gs[rsub + 4 * e][col] = greg[e];
      As[rsub + 4 * e][col] = areg[e];
      bsum += greg[e];
    }
    __syncthreads();
    if (mb + WM < m_end) load(mb + WM);
#pragma unroll
    for (int mm = 0; mm < WM; ++mm) {
      const float4 g4 = *reinterpret_cast<const float4*>(&Gs[mm][ty * 4]);
      const float4 a4 = *reinterpret_cast<const float4*>(&As[mm][tx * 4]);
      const float g[4] = {g4.x, g4.y, g4.z, g4.w};
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(g[i], a[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= p.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kw = kw0 + tx * 4 + j;
      if (kw >= KW) continue;
      const int jj = kw / p.Cin, cc = kw - jj * p.Cin;
      atomicAdd(p.dW + (long long)n * p.ldw + (long long)jj * p.wsj + (long long)cc * p.wsc, acc[i][j]);
    }
  }
  if (do_bias) {
    bred[rsub][col] = bsum;
    __syncthreads();
    if (tid < WN && n0 + tid < p.N)
      atomicAdd(p.dbias + n0 + tid, bred[0][tid] + bred[1][tid] + bred[2][tid] + bred[3][tid]);
  }
}

// ---- WavEncoder conv1: one input channel, HBM-bound.  One thread = one output frame x all N channels.
template <int NP>
__global__ void __launch_bounds__(256) conv1_direct_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ y,
                                                           int B, int Tin, int Tout, int N, int taps, int stride, int pad) {
  __shared__ float ws[32 * NP];
  __shared__ float bs[NP];
  for (int i = threadIdx.x; i < 32 * NP; i += blockDim.x) ws[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < N * taps; i += blockDim.x) {
    int n = i / taps, j = i - n * taps;
    ws[j * NP + n] = w[i];
  }
  if (threadIdx.x < NP) bs[threadIdx.x] = (threadIdx.x < N && bias) ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const long long total = (long long)B * Tout;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < total; m += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(m / Tout), t = (int)(m - (long long)b * Tout);
    const int t0 = t * stride - pad;
    const float* xb = x + (long long)b * Tin;
    float acc[NP];
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[n] = bs[n];
    for (int j = 0; j < taps; ++j) {
      const int ti = t0 + j;
      const float xv = (ti >= 0 && ti < Tin) ? __ldg(xb + ti) : 0.f;
#pragma unroll
      for (int n = 0; n < NP; ++n) acc[n] = fmaf(xv, ws[j * NP + n], acc[n]);
    }
    float* yo = y + m * N;
    if ((N & 3) == 0) {
#pragma unroll
      for (int n = 0; n < NP; n += 4)
        if (n < N) *reinterpret_cast<float4*>(yo + n) = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
    } else {
#pragma unroll
      for (int n = 0; n < NP; ++n)
        if (n < N) yo[n] = acc[n];
    }
  }
}

}  // namespace

extern "C" int tg_struct_sizes(int* out2) {
  out2[0] = (int)sizeof(tg_conv_gemm_t); out2[1] = (int)sizeof(tg_conv_wgrad_t);
  return 0;
}

extern "C" int tg_conv_gemm_f32(const tg_conv_gemm_t* pp, tg_stream stream) {
  const tg_conv_gemm_t& p = *pp;
  TG_REQUIRE(p.A && p.W && p.Y, "tg_conv_gemm_f32");
  TG_REQUIRE(p.B > 0 && p.Tout > 0 && p.N > 0 && p.Cin > 0 && p.taps > 0, "tg_conv_gemm_f32");
  const long long M = (long long)p.B * p.Tout;
  TG_REQUIRE(M < (1ll << 31), "tg_conv_gemm_f32");
  cudaStream_t s = (cudaStream_t)stream;
  int bn;
  if (p.N <= 32) bn = 32;
  else {
    int pad64 = tg_ceil_div(p.N, 64) * 64, pad128 = tg_ceil_div(p.N, 128) * 128;
    bn = (pad128 <= pad64) ? 128 : 64;
  }
  // few row tiles (the per-step GEMMs of the seq2seq decoder, M = batch): a wide tile would leave all but 2-5 SMs idle while each of
  // those grinds through a 128x128 FFMA tile - narrow the tile until the grid covers a good part of the chip
  const int mt = tg_ceil_div(M, BM);
  while (bn > 32 && (long long)mt * tg_ceil_div(p.N, bn) < tg_num_sms() / 2) bn >>= 1;
  dim3 grid(mt, tg_ceil_div(p.N, bn));
  TG_REQUIRE(grid.y < 65536, "tg_conv_gemm_f32");
  if (bn == 128) conv_gemm_kernel<128, 8><<<grid, 256, 0, s>>>(p);
  else if (bn == 64) conv_gemm_kernel<64, 4><<<grid, 256, 0, s>>>(p);
  else conv_gemm_kernel<32, 2><<<grid, 256, 0, s>>>(p);
  TG_CHECK_LAUNCH("tg_conv_gemm_f32");
  return 0;
}

extern "C" int tg_conv_wgrad_f32(const tg_conv_wgrad_t* pp, tg_stream stream) {
  const tg_conv_wgrad_t& p = *pp;
  TG_REQUIRE(p.A && p.G && p.dW, "tg_conv_wgrad_f32");
  const long long M = (long long)p.B * p.Tout;
  TG_REQUIRE(M > 0 && M < (1ll << 31), "tg_conv_wgrad_f32");
  const int KW = p.taps * p.Cin;
  const int tk = tg_ceil_div(KW, WK), tn = tg_ceil_div(p.N, WN);
  int splits = (4 * tg_num_sms()) / (tk * tn);
  if (splits < 1) splits = 1;
  int max_splits = tg_ceil_div(M, 64);
  if (splits > max_splits) splits = max_splits;
  if (splits > 65535) splits = 65535;
  int rows = tg_ceil_div(M, splits);
  rows = tg_ceil_div(rows, WM) * WM;
  splits = tg_ceil_div(M, rows);
  dim3 grid(tk, tn, splits);
  conv_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, rows);
  TG_CHECK_LAUNCH("tg_conv_wgrad_f32");
  return 0;
}

extern "C" int tg_conv1_direct_f32(const float* x, const float* w, const float* bias, float* y, int B, int Tin, int Tout,
                                   int N, int taps, int stride, int pad, tg_stream stream) {
  TG_REQUIRE(x && w && y && N <= 32 && taps <= 32 && N > 0, "tg_conv1_direct_f32");
  const long long total = (long long)B * Tout;
  int blocks = (int)((total + 255) / 256);
  int cap = tg_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (N <= 16) conv1_direct_kernel<16><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, B, Tin, Tout, N, taps, stride, pad);
  else conv1_direct_kernel<32><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, B, Tin, Tout, N, taps, stride, pad);
  TG_CHECK_LAUNCH("tg_conv1_direct_f32");
  return 0;
}
