// Speech2Gesture baseline (scripts/model/speech2gesture.py, scripts/train_eval/train_speech2gesture.py): the pieces that are not GEMMs.
// The Conv2d / Conv1d layers of its U-Net run as  im2col (this file)  ->  tcgen05 TF32 GEMM (gemm_tf32.cu) / fp32 GEMM (gemm_f32.cu)
// forward,  weight-gradient GEMM on the same column matrix and  column GEMM -> col2im (this file)  backward; activations are channels-last
// ([B,H,W,C], a 1-D sequence is H = 1), TensorFlow "SAME" padding (speech2gesture.py:9-52) is the (pad_top, pad_left) pair plus implicit
// zeros past the far edge.  All kernels here are HBM-bound gathers with 16-byte accesses where the channel count allows.
#include "common.cuh"

namespace {

#define GRID_STRIDE(i, n) for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

inline int s2g_blocks(long long n) {
  long long b = (n + 255) / 256, cap = (long long)tg_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// col[((b*Ho + ho)*Wo + wo), (i*kw + j)*C + c] = x[b, ho*sh + i - pt, wo*sw + j - pl, c]  (0 outside the image)
template <int V>
__global__ void im2col2d_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int H, int W, int C, int kh, int kw, int sh, int sw,
                                int pt, int pl, int Ho, int Wo) {
  const int CV = C / V;
  const long long n = (long long)B * Ho * Wo * kh * kw * CV;
  GRID_STRIDE(idx, n) {
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int j = (int)(r % kw); r /= kw;
    const int i = (int)(r % kh); r /= kh;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const int h = ho * sh + i - pt, w = wo * sw + j - pl;
    const bool in = h >= 0 && h < H && w >= 0 && w < W;
    const long long src = (((long long)b * H + h) * W + w) * C + (long long)cv * V;
    if (V == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in) v = __ldg(reinterpret_cast<const float4*>(x + src));
      reinterpret_cast<float4*>(col)[idx] = v;
    } else {
      col[idx] = in ? __ldg(x + src) : 0.f;
    }
  }
}

// dx[b, h, w, c] = sum over (i, j) with (h + pt - i) % sh == 0, (w + pl - j) % sw == 0 of col[(b, (h+pt-i)/sh, (w+pl-j)/sw), (i*kw + j)*C + c]
template <int V>
__global__ void col2im2d_kernel(const float* __restrict__ col, float* __restrict__ dx, int B, int H, int W, int C, int kh, int kw, int sh, int sw,
                                int pt, int pl, int Ho, int Wo) {
  const int CV = C / V;
  const long long n = (long long)B * H * W * CV;
  const long long K = (long long)kh * kw * C;
  GRID_STRIDE(idx, n) {
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int b = (int)(r / H);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int i = 0; i < kh; ++i) {
      const int hn = h + pt - i;
      if (hn < 0 || hn % sh) continue;
      const int ho = hn / sh;
      if (ho >= Ho) continue;
      for (int j = 0; j < kw; ++j) {
        const int wn = w + pl - j;
        if (wn < 0 || wn % sw) continue;
        const int wo = wn / sw;
        if (wo >= Wo) continue;
        const float* p = col + (((long long)b * Ho + ho) * Wo + wo) * K + (long long)(i * kw + j) * C + (long long)cv * V;
        if (V == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p));
          a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
        } else {
          a0 += __ldg(p);
        }
      }
    }
    if (V == 4) reinterpret_cast<float4*>(dx)[idx] = make_float4(a0, a1, a2, a3);
    else dx[idx] = a0;
  }
}

// torch.nn.Upsample(size, mode='bilinear', align_corners=False) on channels-last [B,H,W,C] -> [B,Ho,Wo,C]  (speech2gesture.py:147,172)
__device__ __forceinline__ void bilin_src(int o, int in_size, int out_size, int& i0, int& i1, float& l1) {
  float s = ((float)o + 0.5f) * ((float)in_size / (float)out_size) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}
__global__ void resize_bilinear_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo) {
  const long long n = (long long)B * Ho * Wo * C;
  GRID_STRIDE(idx, n) {
    const int c = (int)(idx % C);
    long long r = idx / C;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    int h0, h1, w0, w1; float lh, lw;
    bilin_src(ho, H, Ho, h0, h1, lh);
    bilin_src(wo, W, Wo, w0, w1, lw);
    const float* xb = x + (long long)b * H * W * C + c;
    const float v00 = __ldg(xb + ((long long)h0 * W + w0) * C), v01 = __ldg(xb + ((long long)h0 * W + w1) * C);
    const float v10 = __ldg(xb + ((long long)h1 * W + w0) * C), v11 = __ldg(xb + ((long long)h1 * W + w1) * C);
    y[idx] = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
  }
}
// dx (zeroed by the caller) += scatter of dy with the same weights
__global__ void resize_bilinear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C, int Ho, int Wo) {
  const long long n = (long long)B * Ho * Wo * C;
  GRID_STRIDE(idx, n) {
    const int c = (int)(idx % C);
    long long r = idx / C;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    int h0, h1, w0, w1; float lh, lw;
    bilin_src(ho, H, Ho, h0, h1, lh);
    bilin_src(wo, W, Wo, w0, w1, lw);
    const float g = dy[idx];
    float* xb = dx + (long long)b * H * W * C + c;
    atomicAdd(xb + ((long long)h0 * W + w0) * C, g * (1.f - lh) * (1.f - lw));
    atomicAdd(xb + ((long long)h0 * W + w1) * C, g * (1.f - lh) * lw);
    atomicAdd(xb + ((long long)h1 * W + w0) * C, g * lh * (1.f - lw));
    atomicAdd(xb + ((long long)h1 * W + w1) * C, g * lh * lw);
  }
}

// UnetUp (speech2gesture.py:120-130): y[b,t,c] = x1[b, t / 2, c] + x2[b,t,c], t < T2 <= 2 T1
__global__ void upsample2_add_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, float* __restrict__ y, int B, int T1, int T2,
                                         int C) {
  const long long n = (long long)B * T2 * C;
  GRID_STRIDE(idx, n) {
    const int c = (int)(idx % C);
    long long r = idx / C;
    const int t = (int)(r % T2);
    const int b = (int)(r / T2);
    y[idx] = __ldg(x1 + ((long long)b * T1 + (t >> 1)) * C + c) + __ldg(x2 + idx);
  }
}
// dx1[b,s,c] (+)= dy[b,2s,c] + dy[b,2s+1,c] (the second only if 2s+1 < T2);  accumulate != 0 adds to what dx1 holds
__global__ void upsample2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx1, int B, int T1, int T2, int C, int accumulate) {
  const long long n = (long long)B * T1 * C;
  GRID_STRIDE(idx, n) {
    const int c = (int)(idx % C);
    long long r = idx / C;
    const int s = (int)(r % T1);
    const int b = (int)(r / T1);
    float v = 0.f;
    if (2 * s < T2) v += __ldg(dy + ((long long)b * T2 + 2 * s) * C + c);
    if (2 * s + 1 < T2) v += __ldg(dy + ((long long)b * T2 + 2 * s + 1) * C + c);
    dx1[idx] = accumulate ? dx1[idx] + v : v;
  }
}

// x[:, 1:] - x[:, :-1] over [B,T,D] (train_speech2gesture.py:12-13, speech2gesture.py:245) and its adjoint
__global__ void time_diff_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int T, int D) {
  const long long n = (long long)B * (T - 1) * D;
  GRID_STRIDE(idx, n) {
    const int d = (int)(idx % D);
    long long r = idx / D;
    const int t = (int)(r % (T - 1));
    const int b = (int)(r / (T - 1));
    const float* p = x + ((long long)b * T + t) * D + d;
    y[idx] = __ldg(p + D) - __ldg(p);
  }
}
// dx[b,t,d] (+)= dy[b,t-1,d] - dy[b,t,d]
__global__ void time_diff_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int T, int D, int accumulate) {
  const long long n = (long long)B * T * D;
  GRID_STRIDE(idx, n) {
    const int d = (int)(idx % D);
    long long r = idx / D;
    const int t = (int)(r % T);
    const int b = (int)(r / T);
    const float* p = dy + ((long long)b * (T - 1)) * D + d;
    float v = 0.f;
    if (t > 0) v += __ldg(p + (long long)(t - 1) * D);
    if (t < T - 1) v -= __ldg(p + (long long)t * D);
    dx[idx] = accumulate ? dx[idx] + v : v;
  }
}

// feat[b,t,:] = [a[b,t,0:Ca] | p[b,0:Cp]]  (speech2gesture.py:219-221) and its adjoint (da = d[:, :, :Ca];  dp[b] = sum_t d[b,t,Ca:])
__global__ void concat_bcast_fwd_kernel(const float* __restrict__ a, const float* __restrict__ p, float* __restrict__ y, int B, int T, int Ca, int Cp) {
  const int C = Ca + Cp;
  const long long n = (long long)B * T * C;
  GRID_STRIDE(idx, n) {
    const int c = (int)(idx % C);
    const long long bt = idx / C;
    y[idx] = c < Ca ? __ldg(a + bt * Ca + c) : __ldg(p + (bt / T) * Cp + (c - Ca));
  }
}
__global__ void concat_bcast_bwd_kernel(const float* __restrict__ d, float* __restrict__ da, float* __restrict__ dp, int B, int T, int Ca, int Cp) {
  const int C = Ca + Cp;
  const long long na = (long long)B * T * Ca, np = (long long)B * Cp;
  GRID_STRIDE(idx, na + np) {
    if (idx < na) {
      const int c = (int)(idx % Ca);
      da[idx] = __ldg(d + (idx / Ca) * C + c);
    } else {
      const long long q = idx - na;
      const int c = (int)(q % Cp);
      const long long b = q / Cp;
      float v = 0.f;
      for (int t = 0; t < T; ++t) v += __ldg(d + (b * T + t) * C + Ca + c);
      dp[q] = v;
    }
  }
}

// dx = dy * (x >= 0 ? 1 : slope)   (LeakyReLU without a BatchNorm in front: the discriminator's first layer, speech2gesture.py:237)
__global__ void lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, long long n, float slope) {
  GRID_STRIDE(i, n) dx[i] = dy[i] * (x[i] >= 0.f ? 1.f : slope);
}

// F.mse_loss(const, x) / nn.L1Loss: *scalar += mean; dx (+)= w * d mean / dx
__global__ void __launch_bounds__(256) mse_const_kernel(const float* __restrict__ x, long long n, float target, float w, double* scalar,
                                                        float* __restrict__ dx) {
  float s = 0.f;
  GRID_STRIDE(i, n) {
    const float e = x[i] - target;
    s += e * e;
    if (dx) dx[i] = w * 2.f * e / (float)n;
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(scalar, (double)t / (double)n);
  }
}
__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, float w, double* scalar,
                                                      float* __restrict__ dx) {
  float s = 0.f;
  GRID_STRIDE(i, n) {
    const float e = x[i] - y[i];
    s += fabsf(e);
    if (dx) dx[i] = w * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) / (float)n;
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(scalar, (double)t / (double)n);
  }
}

}  // namespace

extern "C" int tg_im2col2d(const float* x, float* col, int B, int H, int W, int C, int kh, int kw, int sh, int sw, int pt, int pl, int Ho, int Wo,
                           tg_stream stream) {
  TG_REQUIRE(x && col && B > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && Ho > 0 && Wo > 0, "tg_im2col2d");
  const bool v4 = (C & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(col)) & 15) == 0;
  const long long n = (long long)B * Ho * Wo * kh * kw * (v4 ? C / 4 : C);
  if (v4) im2col2d_kernel<4><<<s2g_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, col, B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo);
  else im2col2d_kernel<1><<<s2g_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, col, B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo);
  TG_CHECK_LAUNCH("tg_im2col2d");
  return 0;
}
extern "C" int tg_col2im2d(const float* col, float* dx, int B, int H, int W, int C, int kh, int kw, int sh, int sw, int pt, int pl, int Ho, int Wo,
                           tg_stream stream) {
  TG_REQUIRE(col && dx && B > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && Ho > 0 && Wo > 0, "tg_col2im2d");
  const bool v4 = (C & 3) == 0 && ((reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(col)) & 15) == 0;
  const long long n = (long long)B * H * W * (v4 ? C / 4 : C);
  if (v4) col2im2d_kernel<4><<<s2g_blocks(n), 256, 0, (cudaStream_t)stream>>>(col, dx, B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo);
  else col2im2d_kernel<1><<<s2g_blocks(n), 256, 0, (cudaStream_t)stream>>>(col, dx, B, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo);
  TG_CHECK_LAUNCH("tg_col2im2d");
  return 0;
}
extern "C" int tg_resize_bilinear_fwd(const float* x, float* y, int B, int H, int W, int C, int Ho, int Wo, tg_stream stream) {
  TG_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "tg_resize_bilinear_fwd");
  resize_bilinear_fwd_kernel<<<s2g_blocks((long long)B * Ho * Wo * C), 256, 0, (cudaStream_t)stream>>>(x, y, B, H, W, C, Ho, Wo);
  TG_CHECK_LAUNCH("tg_resize_bilinear_fwd");
  return 0;
}
extern "C" int tg_resize_bilinear_bwd(const float* dy, float* dx, int B, int H, int W, int C, int Ho, int Wo, tg_stream stream) {
  TG_REQUIRE(dy && dx && B > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, "tg_resize_bilinear_bwd");
  cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * H * W * C, (cudaStream_t)stream);
  if (e != cudaSuccess) { tg_set_error("tg_resize_bilinear_bwd: memset: %s", cudaGetErrorString(e)); return -2; }
  resize_bilinear_bwd_kernel<<<s2g_blocks((long long)B * Ho * Wo * C), 256, 0, (cudaStream_t)stream>>>(dy, dx, B, H, W, C, Ho, Wo);
  TG_CHECK_LAUNCH("tg_resize_bilinear_bwd");
  return 0;
}
extern "C" int tg_upsample2_add_fwd(const float* x1, const float* x2, float* y, int B, int T1, int T2, int C, tg_stream stream) {
  TG_REQUIRE(x1 && x2 && y && B > 0 && T1 > 0 && T2 > 0 && T2 <= 2 * T1 && C > 0, "tg_upsample2_add_fwd");
  upsample2_add_fwd_kernel<<<s2g_blocks((long long)B * T2 * C), 256, 0, (cudaStream_t)stream>>>(x1, x2, y, B, T1, T2, C);
  TG_CHECK_LAUNCH("tg_upsample2_add_fwd");
  return 0;
}
extern "C" int tg_upsample2_bwd(const float* dy, float* dx1, int B, int T1, int T2, int C, int accumulate, tg_stream stream) {
  TG_REQUIRE(dy && dx1 && B > 0 && T1 > 0 && T2 > 0 && T2 <= 2 * T1 && C > 0, "tg_upsample2_bwd");
  upsample2_bwd_kernel<<<s2g_blocks((long long)B * T1 * C), 256, 0, (cudaStream_t)stream>>>(dy, dx1, B, T1, T2, C, accumulate);
  TG_CHECK_LAUNCH("tg_upsample2_bwd");
  return 0;
}
extern "C" int tg_time_diff_fwd(const float* x, float* y, int B, int T, int D, tg_stream stream) {
  TG_REQUIRE(x && y && B > 0 && T > 1 && D > 0, "tg_time_diff_fwd");
  time_diff_fwd_kernel<<<s2g_blocks((long long)B * (T - 1) * D), 256, 0, (cudaStream_t)stream>>>(x, y, B, T, D);
  TG_CHECK_LAUNCH("tg_time_diff_fwd");
  return 0;
}
extern "C" int tg_time_diff_bwd(const float* dy, float* dx, int B, int T, int D, int accumulate, tg_stream stream) {
  TG_REQUIRE(dy && dx && B > 0 && T > 1 && D > 0, "tg_time_diff_bwd");
  time_diff_bwd_kernel<<<s2g_blocks((long long)B * T * D), 256, 0, (cudaStream_t)stream>>>(dy, dx, B, T, D, accumulate);
  TG_CHECK_LAUNCH("tg_time_diff_bwd");
  return 0;
}
extern "C" int tg_concat_bcast_fwd(const float* a, const float* p, float* y, int B, int T, int Ca, int Cp, tg_stream stream) {
  TG_REQUIRE(a && p && y && B > 0 && T > 0 && Ca > 0 && Cp > 0, "tg_concat_bcast_fwd");
  concat_bcast_fwd_kernel<<<s2g_blocks((long long)B * T * (Ca + Cp)), 256, 0, (cudaStream_t)stream>>>(a, p, y, B, T, Ca, Cp);
  TG_CHECK_LAUNCH("tg_concat_bcast_fwd");
  return 0;
}
extern "C" int tg_concat_bcast_bwd(const float* d, float* da, float* dp, int B, int T, int Ca, int Cp, tg_stream stream) {
  TG_REQUIRE(d && da && dp && B > 0 && T > 0 && Ca > 0 && Cp > 0, "tg_concat_bcast_bwd");
  concat_bcast_bwd_kernel<<<s2g_blocks((long long)B * T * Ca + (long long)B * Cp), 256, 0, (cudaStream_t)stream>>>(d, da, dp, B, T, Ca, Cp);
  TG_CHECK_LAUNCH("tg_concat_bcast_bwd");
  return 0;
}
extern "C" int tg_lrelu_bwd(const float* dy, const float* x, float* dx, long long n, float slope, tg_stream stream) {
  TG_REQUIRE(dy && x && dx && n > 0, "tg_lrelu_bwd");
  lrelu_bwd_kernel<<<s2g_blocks(n), 256, 0, (cudaStream_t)stream>>>(dy, x, dx, n, slope);
  TG_CHECK_LAUNCH("tg_lrelu_bwd");
  return 0;
}
extern "C" int tg_mse_const(const float* x, long long n, float target, float w, double* scalar, float* dx, tg_stream stream) {
  TG_REQUIRE(x && scalar && n > 0, "tg_mse_const");
  long long b = (n + 255) / 256;
  if (b > 256) b = 256;
  mse_const_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(x, n, target, w, scalar, dx);
  TG_CHECK_LAUNCH("tg_mse_const");
  return 0;
}
extern "C" int tg_l1_loss(const float* x, const float* y, long long n, float w, double* scalar, float* dx, tg_stream stream) {
  TG_REQUIRE(x && y && scalar && n > 0, "tg_l1_loss");
  long long b = (n + 255) / 256;
  if (b > 256) b = 256;
  l1_loss_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(x, y, n, w, scalar, dx);
  TG_CHECK_LAUNCH("tg_l1_loss");
  return 0;
}
