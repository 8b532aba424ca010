// Helpers of the tensor-core WavEncoder path (fast mode): the strided convolutions conv2-4
// (multimodal_context_net.py:16-22) run as TF32 GEMMs whose A operand is the overlapping-window view of the
// channels-last activation (tg_gemm_tf32 clip mode / tg_wgrad_tf32 x_clip_pitch), so what is left for CUDA cores is
//   * (BatchNorm-apply + LeakyReLU is one elementwise pass, tg_affine_lrelu: TMA-fed MMAs have no operand prologue),
//   * the [N,Cin,k] <-> [N,k*Cin] filter re-layouts,
//   * col2im: the data gradient gathers <= ceil(k/stride) entries of the [M, k*Cin] "column" GEMM result, no atomics,
//   * conv1's weight gradient (Cin = 1: an HBM-bound correlation of dy with the raw waveform).
#include "common.cuh"

namespace {

// w [N][Cin][k] (nn.Conv1d) -> w2 [N][k*Cin] (tap-major, channel-minor: the order of a window row) and w2t [k*Cin][N]
__global__ void window_weights_kernel(const float* __restrict__ w, float* __restrict__ w2, float* __restrict__ w2t, int N, int Cin, int k) {
  const int total = N * Cin * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / (k * Cin), r = i - n * (k * Cin);
    const int j = r / Cin, c = r - j * Cin;
    const float v = w[(n * Cin + c) * k + j];
    w2[i] = v;
    if (w2t) w2t[r * N + n] = v;
  }
}
// w [N][Cin][k] -> wd [ceil(k/stride)][stride*Cin][N]: wd[j][r*Cin + c][n] = w[n][c][r + stride*j] (0 when that tap does not exist)
__global__ void window_dgrad_weights_kernel(const float* __restrict__ w, float* __restrict__ wd, int N, int Cin, int k, int stride) {
  const int ntap = (k + stride - 1) / stride, NP = stride * Cin;
  const int total = ntap * NP * N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % N, rc = (i / N) % NP, j = i / (N * NP);
    const int r = rc / Cin, c = rc - r * Cin;
    const int tap = r + stride * j;
    wd[i] = tap < k ? w[(n * Cin + c) * k + tap] : 0.f;
  }
}
// dw [N][Cin][k] += dw2 [N][k*Cin]
__global__ void window_wgrad_add_kernel(const float* __restrict__ dw2, float* __restrict__ dw, int N, int Cin, int k) {
  const int total = N * Cin * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / (Cin * k), r = i - n * (Cin * k);
    const int c = r / k, j = r - c * k;
    dw[i] += dw2[n * (k * Cin) + j * Cin + c];
  }
}

// da[(b*Tin + s)*Cin + c] = sum_{j = s mod stride, += stride, < k; t = (s-j)/stride in [0,Tout)} col[(b*Tout + t)*(k*Cin) + j*Cin + c]
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ col, float* __restrict__ da, int B, int Tin, int Tout, int Cin,
                                                     int k, int stride) {
  const int c4n = Cin >> 2;
  const long long total = (long long)B * Tin * c4n;
  const long long K = (long long)k * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    const long long bs = i / c4n;
    const int b = (int)(bs / Tin), s = (int)(bs - (long long)b * Tin);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = s % stride; j < k; j += stride) {
      const int t = (s - j) / stride;
      if (t < 0) break;
      if (t >= Tout) continue;
      const float4 v = *reinterpret_cast<const float4*>(col + ((long long)b * Tout + t) * K + (long long)j * Cin + c4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(da + bs * Cin + c4 * 4) = acc;
  }
}

// conv1 weight gradient, Cin = 1, N = 16, taps <= 15 (the bias gradient rides along as "tap 15" with x == 1):
//   dW[n][j] += sum_{b,t} dy[(b*Tout+t)*16 + n] * x[b*Tin + t*stride + j - pad]
// persistent blocks; a tile = TT output positions of one clip staged in shared memory; thread = (tap j, channel quad, tile quarter)
constexpr int C1_TT = 256;
__global__ void __launch_bounds__(256) conv1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dW,
                                                          float* __restrict__ dbias, int B, int Tin, int Tout, int taps, int stride, int pad) {
  extern __shared__ float sm[];
  float* dyS = sm;                                   // [TT][16]
  float* xS = sm + C1_TT * 16;                       // [TT*stride + 16]
  const int xlen = C1_TT * stride + 16;
  const int j = threadIdx.x & 15, nq = (threadIdx.x >> 4) & 3, part = threadIdx.x >> 6;
  const int tiles_per_clip = (Tout + C1_TT - 1) / C1_TT;
  const int tiles = B * tiles_per_clip;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_clip, t0 = (tile - b * tiles_per_clip) * C1_TT;
    const int nt = min(C1_TT, Tout - t0);
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(dy + ((long long)b * Tout + t0) * 16);
    for (int i = threadIdx.x; i < C1_TT * 4; i += 256)
      reinterpret_cast<float4*>(dyS)[i] = (i < nt * 4) ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int x0 = t0 * stride - pad;
    const float* xb = x + (long long)b * Tin;
    for (int i = threadIdx.x; i < xlen; i += 256) {
      const int xi = x0 + i;
      xS[i] = (xi >= 0 && xi < Tin) ? xb[xi] : 0.f;
    }
    __syncthreads();
    const int tb = part * (C1_TT / 4);
#pragma unroll 4
    for (int tt = 0; tt < C1_TT / 4; ++tt) {
      const int t = tb + tt;
      const float4 g = *reinterpret_cast<const float4*>(dyS + t * 16 + nq * 4);
      const float xv = j < taps ? xS[t * stride + j] : 1.f;
      acc[0] = fmaf(g.x, xv, acc[0]); acc[1] = fmaf(g.y, xv, acc[1]); acc[2] = fmaf(g.z, xv, acc[2]); acc[3] = fmaf(g.w, xv, acc[3]);
    }
  }
  __syncthreads();
  float* red = sm;                                   // [4 parts][16 n][16 j]
#pragma unroll
  for (int e = 0; e < 4; ++e) red[(part * 16 + nq * 4 + e) * 16 + j] = acc[e];
  __syncthreads();
  {
    const int n = threadIdx.x >> 4, jj = threadIdx.x & 15;
    const float v = red[(0 * 16 + n) * 16 + jj] + red[(1 * 16 + n) * 16 + jj] + red[(2 * 16 + n) * 16 + jj] + red[(3 * 16 + n) * 16 + jj];
    if (jj < taps) atomicAdd(dW + n * taps + jj, v);
    else if (jj == 15 && dbias) atomicAdd(dbias + n, v);
  }
}

// Register-tiled version for the WavEncoder's actual conv1 (15 taps, stride 5): thread = (chunk of 32 consecutive output frames, channel
// quad) keeps ALL 15 x 4 weight-gradient partial sums in registers and slides a 15-sample window of the waveform along its frames (5 new
// samples per frame), so a frame costs one 16-byte load of dy + 5 cached waveform loads for 60 FMAs and nothing is staged through shared
// memory.  The tile-staging kernel above spends its time waiting on its own load -> __syncthreads -> compute rounds (2 CTAs per SM):
// 82-130 us for 83 MB of traffic, the last kernel of the iteration's audio chain.
template <int TAPS, int STRIDE>
__global__ void __launch_bounds__(256) conv1_wgrad_reg_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dW,
                                                              float* __restrict__ dbias, int Tin, int Tout, int pad, int chunks_per_clip,
                                                              long long total_chunks) {
  constexpr int FPC = 32, NV = TAPS * 4 + 4;
  __shared__ float red[8][4][NV];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, quad = tid & 3;
  const long long chunk = (long long)blockIdx.x * 64 + (tid >> 2);
  float acc[TAPS][4], accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < TAPS; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  if (chunk < total_chunks) {
    const int b = (int)(chunk / chunks_per_clip), f0 = (int)(chunk - (long long)b * chunks_per_clip) * FPC;
    const int nf = min(FPC, Tout - f0);
    const float* xb = x + (long long)b * Tin;
    const float* dyp = dy + ((long long)b * Tout + f0) * 16 + quad * 4;
    const int xi0 = f0 * STRIDE - pad;
    auto ldx = [&](int xi) { return (xi >= 0 && xi < Tin) ? __ldg(xb + xi) : 0.f; };
    float win[TAPS];
#pragma unroll
    for (int j = 0; j < TAPS; ++j) win[j] = 0.f;
#pragma unroll
    for (int j = 0; j < TAPS - STRIDE; ++j) win[j + STRIDE] = ldx(xi0 + j);
#pragma unroll 3
    for (int f = 0; f < nf; ++f) {
#pragma unroll
      for (int j = 0; j < TAPS - STRIDE; ++j) win[j] = win[j + STRIDE];
#pragma unroll
      for (int j = TAPS - STRIDE; j < TAPS; ++j) win[j] = ldx(xi0 + f * STRIDE + j);
      const float4 g = __ldg(reinterpret_cast<const float4*>(dyp + (long long)f * 16));
#pragma unroll
      for (int j = 0; j < TAPS; ++j) {
        acc[j][0] = fmaf(g.x, win[j], acc[j][0]); acc[j][1] = fmaf(g.y, win[j], acc[j][1]);
        acc[j][2] = fmaf(g.z, win[j], acc[j][2]); acc[j][3] = fmaf(g.w, win[j], acc[j][3]);
      }
      accb[0] += g.x; accb[1] += g.y; accb[2] += g.z; accb[3] += g.w;
    }
  }
  // lanes with equal (lane & 3) hold the same channel quad: sum the warp's 8 chunks, then the block's 8 warps, then one atomic per value
#pragma unroll
  for (int j = 0; j < TAPS; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = acc[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) red[warp][lane][j * 4 + e] = v;
    }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v = accb[e];
    v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (lane < 4) red[warp][lane][TAPS * 4 + e] = v;
  }
  __syncthreads();
  for (int o = tid; o < 4 * NV; o += 256) {
    const int qd = o / NV, idx = o - qd * NV;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][qd][idx];
    if (idx < TAPS * 4) atomicAdd(dW + (qd * 4 + (idx & 3)) * TAPS + (idx >> 2), v);
    else if (dbias) atomicAdd(dbias + qd * 4 + (idx - TAPS * 4), v);
  }
}

inline int ew_grid(long long n, int per_block = 256) {
  long long b = (n + per_block - 1) / per_block;
  const long long cap = (long long)tg_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int tg_window_weights(const float* w, float* w2, float* w2t, int N, int Cin, int k, tg_stream stream) {
  TG_REQUIRE(w && w2 && N > 0 && Cin > 0 && k > 0, "tg_window_weights");
  window_weights_kernel<<<ew_grid((long long)N * Cin * k), 256, 0, (cudaStream_t)stream>>>(w, w2, w2t, N, Cin, k);
  TG_CHECK_LAUNCH("tg_window_weights");
  return 0;
}

extern "C" int tg_window_dgrad_weights(const float* w, float* wd, int N, int Cin, int k, int stride, tg_stream stream) {
  TG_REQUIRE(w && wd && N > 0 && Cin > 0 && k > 0 && stride > 0, "tg_window_dgrad_weights");
  window_dgrad_weights_kernel<<<ew_grid((long long)tg_ceil_div(k, stride) * stride * Cin * N), 256, 0, (cudaStream_t)stream>>>(w, wd, N, Cin, k, stride);
  TG_CHECK_LAUNCH("tg_window_dgrad_weights");
  return 0;
}

extern "C" int tg_window_wgrad_add(const float* dw2, float* dw, int N, int Cin, int k, tg_stream stream) {
  TG_REQUIRE(dw2 && dw && N > 0 && Cin > 0 && k > 0, "tg_window_wgrad_add");
  window_wgrad_add_kernel<<<ew_grid((long long)N * Cin * k), 256, 0, (cudaStream_t)stream>>>(dw2, dw, N, Cin, k);
  TG_CHECK_LAUNCH("tg_window_wgrad_add");
  return 0;
}

extern "C" int tg_col2im(const float* col, float* da, int B, int Tin, int Tout, int Cin, int k, int stride, tg_stream stream) {
  TG_REQUIRE(col && da && B > 0 && Tin > 0 && Tout > 0 && k > 0 && stride > 0 && Cin > 0 && (Cin & 3) == 0, "tg_col2im");
  TG_REQUIRE(((reinterpret_cast<uintptr_t>(col) | reinterpret_cast<uintptr_t>(da)) & 15) == 0, "tg_col2im(alignment)");
  col2im_kernel<<<ew_grid((long long)B * Tin * (Cin / 4)), 256, 0, (cudaStream_t)stream>>>(col, da, B, Tin, Tout, Cin, k, stride);
  TG_CHECK_LAUNCH("tg_col2im");
  return 0;
}

extern "C" int tg_conv1_wgrad(const float* x, const float* dy, float* dW, float* dbias, int B, int Tin, int Tout, int N, int taps, int stride,
                              int pad, tg_stream stream) {
  TG_REQUIRE(x && dy && dW && B > 0 && Tin > 0 && Tout > 0, "tg_conv1_wgrad");
  TG_REQUIRE(N == 16 && taps >= 1 && taps <= 15 && stride >= 1 && stride <= 8, "tg_conv1_wgrad(shape)");
  TG_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0, "tg_conv1_wgrad(alignment)");
  if (taps == 15 && stride == 5) {
    const int cpc = tg_ceil_div(Tout, 32);
    const long long total = (long long)B * cpc;
    conv1_wgrad_reg_kernel<15, 5><<<(unsigned)((total + 63) / 64), 256, 0, (cudaStream_t)stream>>>(x, dy, dW, dbias, Tin, Tout, pad, cpc, total);
    TG_CHECK_LAUNCH("tg_conv1_wgrad");
    return 0;
  }
  const size_t smem = (size_t)(C1_TT * 16 + C1_TT * stride + 16) * sizeof(float);
  const int tiles = B * tg_ceil_div(Tout, C1_TT);
  int grid = 2 * tg_num_sms();
  if (grid > tiles) grid = tiles;
  conv1_wgrad_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, dy, dW, dbias, B, Tin, Tout, taps, stride, pad);
  TG_CHECK_LAUNCH("tg_conv1_wgrad");
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Input staging (train_eval/staging.py): host->device copy of a pinned batch done by a KERNEL that reads the mapped pinned memory
// directly over PCIe.  A cudaMemcpyAsync would occupy a copy engine for ~0.35 ms per 18.6 MB audio batch, and every copy-engine
// operation of the concurrently running training step (graph memset / memcpy nodes) queues behind it; a handful of otherwise idle
// CTAs do the same transfer without touching the copy engines.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) copy_bytes_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, long long n16,
                                                         unsigned char* __restrict__ dst_tail, const unsigned char* __restrict__ src_tail, int ntail) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}
}  // namespace

extern "C" int tg_copy_bytes(void* dst, const void* src, long long nbytes, int max_ctas, tg_stream stream) {
  TG_REQUIRE(dst && src && nbytes > 0, "tg_copy_bytes");
  TG_REQUIRE(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0, "tg_copy_bytes(alignment)");
  const long long n16 = nbytes / 16;
  const int ntail = (int)(nbytes - n16 * 16);
  long long blocks = (n16 + 255) / 256;
  if (max_ctas < 1) max_ctas = 64;
  if (blocks > max_ctas) blocks = max_ctas;
  if (blocks < 1) blocks = 1;
  copy_bytes_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), n16,
                                                                    reinterpret_cast<unsigned char*>(dst) + n16 * 16,
                                                                    reinterpret_cast<const unsigned char*>(src) + n16 * 16, ntail);
  TG_CHECK_LAUNCH("tg_copy_bytes");
  return 0;
}
