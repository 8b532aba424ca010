// Kernels that only the FGD auto-encoder trainer needs (train_feature_extractor.py:54-97, train_joint_embed.py:5-65):
// the L1 reconstruction loss with its frame-difference term (value + gradient in one pass, one CTA per clip) and the
// layout switch between the reference's channel-major flatten / view ([B,C,T], embedding_net.py:71,213) and the
// channels-last activations every other kernel works on.  Everything is a few hundred KB: launch-latency bound.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }   // torch: d|x|/dx = sign(x), 0 at 0

__global__ void __launch_bounds__(256) ae_recon_loss_kernel(const float* __restrict__ recon, const float* __restrict__ target, int T,
                                                            int D, int use_diff, float weight, double* __restrict__ acc,
                                                            float* __restrict__ d_recon) {
  __shared__ float red0[8], red1[8];
  const long long base = (long long)blockIdx.x * T * D;
  const float* r = recon + base;
  const float* y = target + base;
  const int TD = T * D;
  const float inv0 = 1.f / (float)TD, inv1 = (T > 1) ? 1.f / (float)((T - 1) * D) : 0.f;
  float l0 = 0.f, l1 = 0.f;
  for (int e = threadIdx.x; e < TD; e += blockDim.x) {
    const int t = e / D;
    const float re = r[e], ye = y[e];
    const float u = re - ye;
    l0 += fabsf(u);
    float g = sgn(u) * inv0;
    if (use_diff) {
      if (t + 1 < T) {                               // e_t = (r[t+1]-r[t]) - (y[t+1]-y[t]): this element is the "-r[t]" term
        const float et = (r[e + D] - re) - (y[e + D] - ye);
        l1 += fabsf(et);
        g -= sgn(et) * inv1;
      }
      if (t > 0) {                                   // ... and the "+r[t]" term of e_{t-1}
        const float ep = (re - r[e - D]) - (ye - y[e - D]);
        g += sgn(ep) * inv1;
      }
    }
    if (d_recon) d_recon[base + e] = weight * g;
  }
  l0 = warp_sum(l0); l1 = warp_sum(l1);
  if ((threadIdx.x & 31) == 0) { red0[threadIdx.x >> 5] = l0; red1[threadIdx.x >> 5] = l1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { s0 += red0[k]; s1 += red1[k]; }
    atomicAdd(acc + 0, (double)(s0 * inv0) + (double)(s1 * inv1));
    atomicAdd(acc + 1, (double)(s0 * inv0));
  }
}

__global__ void transpose_batched_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int R, int C) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % R);
    const long long bc = i / R;
    const int c = (int)(bc % C);
    const long long b = bc / C;
    out[i] = in[(b * R + r) * C + c];
  }
}

}  // namespace

extern "C" int tg_ae_recon_loss(const float* recon, const float* target, int B, int T, int D, int use_diff, float weight, double* acc,
                                float* d_recon, tg_stream stream) {
  TG_REQUIRE(recon && target && acc && B > 0 && T > 0 && D > 0, "tg_ae_recon_loss");
  ae_recon_loss_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(recon, target, T, D, use_diff, weight, acc, d_recon);
  TG_CHECK_LAUNCH("tg_ae_recon_loss");
  return 0;
}

extern "C" int tg_transpose_batched_f32(const float* in, float* out, int B, int R, int C, tg_stream stream) {
  TG_REQUIRE(in && out && B > 0 && R > 0 && C > 0, "tg_transpose_batched_f32");
  const long long n = (long long)B * R * C;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)tg_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  transpose_batched_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n, R, C);
  TG_CHECK_LAUNCH("tg_transpose_batched_f32");
  return 0;
}
