// Fused losses of one G+D iteration (train_gan.py:41,53-56,67-82): value and gradient in a single pass, one CTA per
// clip, 128-bit-free scalar loads are fine here (the whole working set is 3 x [B,34,27] floats = 1.4 MB at B=128:
// launch-latency bound, not HBM bound - see DESIGN.md).
#include "common.cuh"

namespace {

__device__ __forceinline__ void huber_term(float x, float y, float beta, float& val, float& grad) {
  // F.smooth_l1_loss(x/beta, y/beta) * beta with torch's beta=1 (train_gan.py:53-54,68-69)
  const float uu = x / beta - y / beta;
  const float d = fabsf(uu);
  if (d < 1.f) { val = 0.5f * d * d * beta; grad = uu; }
  else { val = (d - 0.5f) * beta; grad = uu > 0.f ? 1.f : -1.f; }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
  return t;
}

__global__ void __launch_bounds__(256) gen_losses_kernel(const float* __restrict__ out, const float* __restrict__ target,
                                                         const float* __restrict__ out_rand, const float* __restrict__ z,
                                                         const float* __restrict__ z_rand, const float* __restrict__ mu,
                                                         const float* __restrict__ logvar, int B, int TD, int Z, float w_reg,
                                                         float w_div, float w_kld, double* scalars, float* __restrict__ d_out,
                                                         float* __restrict__ dmu, float* __restrict__ dlogvar) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* o = out + (long long)b * TD;
  const float* tg = target + (long long)b * TD;
  const float* orr = out_rand ? out_rand + (long long)b * TD : nullptr;
  float hub = 0.f, pl1 = 0.f;
  for (int e = threadIdx.x; e < TD; e += blockDim.x) {
    float v, g;
    huber_term(o[e], tg[e], 0.1f, v, g);
    hub += v;
    if (orr) { huber_term(o[e], orr[e], 0.05f, v, g); pl1 += v; }
  }
  hub = block_sum(hub, red);
  float coef = 0.f, div_i = 0.f;
  if (orr) {
    pl1 = block_sum(pl1, red);
    float zl = 0.f;
    for (int e = threadIdx.x; e < Z; e += blockDim.x) zl += fabsf(z[(long long)b * Z + e] - z_rand[(long long)b * Z + e]);
    zl = block_sum(zl, red) / (float)Z;
    const float denom = zl + 1.0e-5f;
    const float raw = -(pl1 / denom);
    div_i = raw < -1000.f ? -1000.f : raw;
    coef = raw >= -1000.f ? -1.f / denom : 0.f;
  }
  float kl = 0.f;
  if (mu) {
    for (int e = threadIdx.x; e < Z; e += blockDim.x) {
      const float m = mu[(long long)b * Z + e], lv = logvar[(long long)b * Z + e];
      const float ex = expf(lv);
      kl += 1.f + lv - m * m - ex;
      if (dmu) {
        dmu[(long long)b * Z + e] = w_kld * m / ((float)B * (float)Z);
        dlogvar[(long long)b * Z + e] = -0.5f * w_kld * (1.f - ex) / ((float)B * (float)Z);
      }
    }
    kl = block_sum(kl, red);
  }
  if (d_out) {
    const float s_reg = w_reg / ((float)B * (float)TD), s_div = w_div * coef / (float)B;
    for (int e = threadIdx.x; e < TD; e += blockDim.x) {
      float v, g;
      huber_term(o[e], tg[e], 0.1f, v, g);
      float gr = s_reg * g;
      if (orr) { huber_term(o[e], orr[e], 0.05f, v, g); gr += s_div * g; }
      d_out[(long long)b * TD + e] = gr;
    }
  }
  if (threadIdx.x == 0) {
    atomicAdd(scalars + 0, (double)hub);
    atomicAdd(scalars + 1, (double)div_i);
    atomicAdd(scalars + 2, (double)kl);
  }
}

__global__ void __launch_bounds__(256) bce_sigmoid_kernel(const float* __restrict__ p, int n, float s, float o, float w,
                                                          double* scalar, float* __restrict__ dlogit) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float pv = p[i];
    const float a = (s * pv + o) + 1e-8f;
    acc += -logf(a);
    if (dlogit) dlogit[i] = w * (-s / ((float)n * a)) * pv * (1.f - pv);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(scalar, (double)acc / (double)n);
}

}  // namespace

extern "C" int tg_gen_losses(const float* out, const float* target, const float* out_rand, const float* z, const float* z_rand,
                             const float* mu, const float* logvar, int B, int TD, int Z, float w_reg, float w_div, float w_kld,
                             double* scalars, float* d_out, float* dmu, float* dlogvar, tg_stream stream) {
  TG_REQUIRE(out && target && scalars && B > 0, "tg_gen_losses");
  TG_REQUIRE(!out_rand || (z && z_rand), "tg_gen_losses");
  gen_losses_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(out, target, out_rand, z, z_rand, mu, logvar, B, TD, Z, w_reg, w_div, w_kld,
                                                        scalars, d_out, dmu, dlogvar);
  TG_CHECK_LAUNCH("tg_gen_losses");
  return 0;
}

extern "C" int tg_bce_sigmoid(const float* p, int n, float s, float o, float w, double* scalar, float* dlogit, tg_stream stream) {
  TG_REQUIRE(p && scalar && n > 0, "tg_bce_sigmoid");
  bce_sigmoid_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p, n, s, o, w, scalar, dlogit);
  TG_CHECK_LAUNCH("tg_bce_sigmoid");
  return 0;
}
