// Persistent bidirectional-GRU recurrence, strict-fp32 path (nn.GRU, multimodal_context_net.py:98-99,155,221-222,241).
//
// One cooperative launch per layer covers both directions and all T steps.  CTA (chunk c, batch group y, dir d) keeps
// the W_hh rows of its `u` hidden units (3u gate rows x H) resident in shared memory for the whole sequence, computes
// gh = W_hh h_{t-1} for its rows and its batch tile as a register-tiled outer-product GEMM, applies the gate
// non-linearities and writes h_t for its units.  The UC chunk-CTAs of one (batch group, direction) exchange h_t through
// L2 (st.global + fence, ld.global.cg) and step together with a release/acquire counter - no grid-wide barrier, no
// kernel launch per time step.  The backward kernel mirrors it: every CTA multiplies its own dgh rows with the same
// resident W_hh slice into a partial dh_{t-1}[all H], partials are summed by the owner of each unit after the barrier.
#include "common.cuh"

namespace {

constexpr int RP = 128;        // padded gate rows per CTA (3*u <= RP)
constexpr int UMAX = 42;

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void group_wait(const int* counter, int target) {
  if (threadIdx.x == 0) {
    while (ld_acquire(counter) < target) { __nanosleep(20); }
  }
  __syncthreads();
}
__device__ __forceinline__ void group_signal(int* counter) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(counter, 1);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct GruFwdP {
  const float* gi;          // [B*T, 6H]
  const float* whhT[2];     // [H, 3H] per direction (transposed W_hh)
  const float* bhh[2];      // [3H]
  float* out;               // [B, T, 2H]
  float* saved;             // [4][..] planes of [B*T, 2H], may be null
  long long saved_qstride;
  int* sync;
  int B, T, H, u, UC, NB, ntiles;
};

template <int BT>
__global__ void __launch_bounds__(256, 1) gru_fwd_kernel(const GruFwdP p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int BC = BT / 16;
  constexpr int HS = BT + 1;
  float* Ws = smem;                       // [H][RP]
  float* hs = smem + (size_t)p.H * RP;    // [H][HS]   (aliased by ghs [RP][HS] after the k loop)
  float* ghs = hs;
  const int tid = threadIdx.x;
  const int c = blockIdx.x, by = blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, u = p.u;
  const int u0 = c * u;
  int* counter = p.sync + dir * p.NB + by;

  // resident weights: Ws[k][g*u + jj] = W_hh[g*H + u0 + jj][k] = whhT[k][g*H + u0 + jj]
  const float* wT = p.whhT[dir];
  for (int i = tid; i < H * RP; i += 256) {
    const int k = i / RP, lr = i - k * RP;
    const int g = lr / u, jj = lr - g * u;
    float v = 0.f;
    if (g < 3 && u0 + jj < H) v = __ldg(wT + (long long)k * 3 * H + g * H + u0 + jj);
    Ws[i] = v;
  }
  __syncthreads();

  const int rg = tid >> 4, bcol = tid & 15;
  const float* bhh = p.bhh[dir];
  const long long row2H = 2ll * H;

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    if (s > 0) group_wait(counter, p.UC * s);
    for (int tile = by; tile < p.ntiles; tile += p.NB) {
      const int b0 = tile * BT;
      if (s > 0) {
        // h_{t-1} tile, transposed into hs[k][bb]
        for (int i = tid; i < BT * H; i += 256) {
          const int bb = i / H, k = i - bb * H;
          const int b = b0 + bb;
          float v = 0.f;
          if (b < p.B) v = __ldcg(p.out + ((long long)b * T + tp) * row2H + dir * H + k);
          hs[k * HS + bb] = v;
        }
        __syncthreads();
        float acc[8][BC];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int q = 0; q < BC; ++q) acc[i][q] = 0.f;
#pragma unroll 4
        for (int k = 0; k < H; ++k) {
          const float4 w0 = *reinterpret_cast<const float4*>(Ws + k * RP + rg * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(Ws + k * RP + rg * 8 + 4);
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          float hv[BC];
#pragma unroll
          for (int q = 0; q < BC; ++q) hv[q] = hs[k * HS + bcol + 16 * q];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < BC; ++q) acc[i][q] = fmaf(w[i], hv[q], acc[i][q]);
        }
        __syncthreads();   // everyone done reading hs before it is overwritten as ghs
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int q = 0; q < BC; ++q) ghs[(rg * 8 + i) * HS + bcol + 16 * q] = acc[i][q];
        __syncthreads();
      }
      // gates for (bb, jj): lanes along units -> coalesced global traffic
      for (int i = tid; i < BT * u; i += 256) {
        const int bb = i / u, jj = i - bb * u;
        const int b = b0 + bb, unit = u0 + jj;
        if (b >= p.B || unit >= H) continue;
        float ghr = bhh[unit], ghz = bhh[H + unit], ghn = bhh[2 * H + unit];
        float hprev = 0.f;
        if (s > 0) {
          ghr += ghs[jj * HS + bb]; ghz += ghs[(u + jj) * HS + bb]; ghn += ghs[(2 * u + jj) * HS + bb];
          hprev = __ldcg(p.out + ((long long)b * T + tp) * row2H + dir * H + unit);
        }
        const long long row = (long long)b * T + t;
        const float* gip = p.gi + row * 6 * H + dir * 3 * H + unit;
        const float r = sigmoidf_(__ldg(gip) + ghr);
        const float z = sigmoidf_(__ldg(gip + H) + ghz);
        const float n = tanhf(__ldg(gip + 2 * H) + r * ghn);
        const float h = (1.f - z) * n + z * hprev;
        const long long o = row * row2H + dir * H + unit;
        p.out[o] = h;
        if (p.saved) {
          p.saved[o] = r; p.saved[p.saved_qstride + o] = z; p.saved[2 * p.saved_qstride + o] = n;
          p.saved[3 * p.saved_qstride + o] = ghn;
        }
      }
      __syncthreads();
    }
    if (s + 1 < T) group_signal(counter);
  }
}

struct GruBwdP {
  const float* dout;        // [B, T, 2H]
  const float* out;         // [B, T, 2H]
  const float* saved; long long saved_qstride;
  const float* whh[2];      // [3H, H]
  float* dgi;               // [B*T, 6H]
  float* dgh;               // [B*T, 6H]
  float* partial;           // [2 parity][2 dir][B][UC][HP]
  int* sync;
  int B, T, H, u, UC, NB, ntiles, HP, tiles_per_cta;
};

template <int BT>
__global__ void __launch_bounds__(256, 1) gru_bwd_kernel(const GruBwdP p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NBC = BT / 8;
  constexpr int DS = BT + 1;
  const int H = p.H, T = p.T, u = p.u, HP = p.HP;
  float* Wb = smem;                             // [RP][HP]
  float* ds = Wb + (size_t)RP * HP;             // [RP][DS]
  float* dhc = ds + (size_t)RP * DS;            // [tiles_per_cta][BT][u]
  const int tid = threadIdx.x;
  const int c = blockIdx.x, by = blockIdx.y, dir = blockIdx.z;
  const int u0 = c * u;
  int* counter = p.sync + dir * p.NB + by;

  const float* W = p.whh[dir];
  for (int i = tid; i < RP * HP; i += 256) {
    const int lr = i / HP, k = i - lr * HP;
    const int g = lr / u, jj = lr - g * u;
    float v = 0.f;
    if (g < 3 && u0 + jj < H && k < H) v = __ldg(W + (long long)(g * H + u0 + jj) * H + k);
    Wb[i] = v;
  }
  for (int i = tid; i < RP * DS; i += 256) ds[i] = 0.f;
  for (int i = tid; i < p.tiles_per_cta * BT * u; i += 256) dhc[i] = 0.f;
  __syncthreads();

  const int kg = tid & 31, bg = tid >> 5;
  const long long row2H = 2ll * H;
  const long long pstride_parity = 2ll * p.B * p.UC * HP;

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? T - 1 - s : s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    const bool tp_ok = tp >= 0 && tp < T;
    if (s > 0) group_wait(counter, p.UC * s);
    const float* Pin = p.partial + (long long)((s - 1) & 1) * pstride_parity + (long long)dir * p.B * p.UC * HP;
    float* Pout = p.partial + (long long)(s & 1) * pstride_parity + (long long)dir * p.B * p.UC * HP;
    int tl = 0;
    for (int tile = by; tile < p.ntiles; tile += p.NB, ++tl) {
      const int b0 = tile * BT;
      float* dh_t = dhc + (size_t)tl * BT * u;
      for (int i = tid; i < BT * u; i += 256) {
        const int bb = i / u, jj = i - bb * u;
        const int b = b0 + bb, unit = u0 + jj;
        float drp = 0.f, dzp = 0.f, dnr = 0.f;
        if (b < p.B && unit < H) {
          float carry = 0.f;
          if (s > 0) {
            carry = dh_t[i];
            const float* pp = Pin + ((long long)b * p.UC) * HP + unit;
            for (int cc = 0; cc < p.UC; ++cc) carry += __ldcg(pp + (long long)cc * HP);
          }
          const long long row = (long long)b * T + t;
          const long long o = row * row2H + dir * H + unit;
          const float dh = __ldg(p.dout + o) + carry;
          const float r = __ldg(p.saved + o), z = __ldg(p.saved + p.saved_qstride + o);
          const float n = __ldg(p.saved + 2 * p.saved_qstride + o), hn = __ldg(p.saved + 3 * p.saved_qstride + o);
          const float hprev = tp_ok ? __ldg(p.out + ((long long)b * T + tp) * row2H + dir * H + unit) : 0.f;
          const float dn = dh * (1.f - z) * (1.f - n * n);
          dzp = dh * (hprev - n) * z * (1.f - z);
          drp = dn * hn * r * (1.f - r);
          dnr = dn * r;
          float* gp = p.dgi + row * 6 * H + dir * 3 * H + unit;
          gp[0] = drp; gp[H] = dzp; gp[2 * H] = dn;
          float* hp = p.dgh + row * 6 * H + dir * 3 * H + unit;
          hp[0] = drp; hp[H] = dzp; hp[2 * H] = dnr;
          dh_t[i] = dh * z;
        }
        ds[jj * DS + bb] = drp; ds[(u + jj) * DS + bb] = dzp; ds[(2 * u + jj) * DS + bb] = dnr;
      }
      __syncthreads();
      if (s + 1 < T) {
        // partial dh_{prev}[bb, k] = sum_{own rows i} dgh[bb, i] * W_hh[i, k]
        float acc[12][NBC];
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
          for (int q = 0; q < NBC; ++q) acc[i][q] = 0.f;
        const int nrows = 3 * u;
        for (int i = 0; i < nrows; ++i) {
          float w[12];
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const int col = kg * 4 + 128 * q;
            float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col < HP) w4 = *reinterpret_cast<const float4*>(Wb + (size_t)i * HP + col);
            w[q * 4] = w4.x; w[q * 4 + 1] = w4.y; w[q * 4 + 2] = w4.z; w[q * 4 + 3] = w4.w;
          }
          float d[NBC];
#pragma unroll
          for (int q = 0; q < NBC; ++q) d[q] = ds[i * DS + bg + 8 * q];
#pragma unroll
          for (int e = 0; e < 12; ++e)
#pragma unroll
            for (int q = 0; q < NBC; ++q) acc[e][q] = fmaf(w[e], d[q], acc[e][q]);
        }
#pragma unroll
        for (int q = 0; q < NBC; ++q) {
          const int b = b0 + bg + 8 * q;
          if (b >= p.B) continue;
          float* po = Pout + ((long long)b * p.UC + c) * HP;
#pragma unroll
          for (int qq = 0; qq < 3; ++qq) {
            const int col = kg * 4 + 128 * qq;
            if (col < HP)
              __stcg(reinterpret_cast<float4*>(po + col),
                     make_float4(acc[qq * 4][q], acc[qq * 4 + 1][q], acc[qq * 4 + 2][q], acc[qq * 4 + 3][q]));
          }
        }
      }
      __syncthreads();
    }
    if (s + 1 < T) group_signal(counter);
  }
}


// =====================================================================================================================
// Small hidden size (H <= 64: the ConvDiscriminator's GRU, multimodal_context_net.py:221-222).  A CTA owns SB = 2 clips of one
// direction for all T steps (B/2 x 2 CTAs: one wave for 128 clips), no inter-CTA exchange.  The recurrent weights live in REGISTERS:
// thread r < 3H keeps row r of W_hh (forward) / column j of one gate block of W_hh (backward) - 64 floats - so the per-step product is
// 64 FMAs per clip fed by broadcast LDS.128 of the 256-byte state vector, instead of streaming 48 KB of weights from shared memory every
// step (the first version: 1.3 us / 2.1 us per step, LDS-bound; ncu r01).  Two __syncthreads per step; the gi / saved-gate loads of a
// step are issued before the product so their latency hides behind it.
// =====================================================================================================================
constexpr int SB = 2;             // clips per CTA
constexpr int HS = 64;            // register-resident weights per thread (H <= 64, zero padded)
constexpr int SMALL_NT = 192;     // 3 * 64 row threads

struct GruSmallFwdP {
  const float* gi; const float* whhT[2]; const float* bhh[2]; float* out; float* saved; long long saved_qstride;
  int B, T, H;
};

__global__ void __launch_bounds__(SMALL_NT) gru_small_fwd_kernel(const GruSmallFwdP p) {
  __shared__ __align__(16) float hs[SB][HS];          // h_{t-1}
  __shared__ __align__(16) float ghs[SB][3 * HS];     // W_hh h_{t-1}
  const int H = p.H, T = p.T, G3 = 3 * H;
  const int tid = threadIdx.x;
  const int dir = blockIdx.y, b0 = blockIdx.x * SB;
  const float* wT = dir ? p.whhT[1] : p.whhT[0];      // [H][3H]: wT[k][r] = W_hh[r][k]
  float w[HS];
#pragma unroll
  for (int k = 0; k < HS; ++k) w[k] = (tid < G3 && k < H) ? __ldg(wT + (long long)k * G3 + tid) : 0.f;
  for (int i = tid; i < SB * HS; i += SMALL_NT) (&hs[0][0])[i] = 0.f;
  const float* bhh = dir ? p.bhh[1] : p.bhh[0];
  const long long row2H = 2ll * H;
  // the (clip, unit) pair this thread owns in the gate phase
  const int pbb = tid / H, punit = tid - pbb * H;
  const int pb = b0 + pbb;
  const bool pok = tid < SB * H && pb < p.B;
  float b_r = 0.f, b_z = 0.f, b_n = 0.f;
  if (pok) { b_r = __ldg(bhh + punit); b_z = __ldg(bhh + H + punit); b_n = __ldg(bhh + 2 * H + punit); }
  __syncthreads();
  // input projections are prefetched ONE STEP AHEAD: a step is ~0.3 us, shorter than a global-memory round trip, so loads issued at the top
  // of the step they belong to stalled the gate phase (measured: 1.0 us per step with the product already down to ~0.2 us)
  float nx_r = 0.f, nx_z = 0.f, nx_n = 0.f;
  if (pok) {
    const float* gp = p.gi + ((long long)pb * T + (dir == 0 ? 0 : T - 1)) * 6 * H + dir * 3 * H + punit;
    nx_r = __ldg(gp); nx_z = __ldg(gp + H); nx_n = __ldg(gp + 2 * H);
  }
  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    const float gir = nx_r, giz = nx_z, gin = nx_n;
    if (pok && s + 1 < T) {
      const float* gp = p.gi + ((long long)pb * T + (dir == 0 ? t + 1 : t - 1)) * 6 * H + dir * 3 * H + punit;
      nx_r = __ldg(gp); nx_z = __ldg(gp + H); nx_n = __ldg(gp + 2 * H);
    }
    if (s > 0) {
      // four independent partial sums per clip: one accumulator per clip is a chain of 64 dependent FMAs (~260 cycles of pure latency)
      float acc[SB][4];
#pragma unroll
      for (int c = 0; c < SB; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
#pragma unroll
      for (int k = 0; k < HS; k += 4) {
#pragma unroll
        for (int c = 0; c < SB; ++c) {
          const float4 h4 = *reinterpret_cast<const float4*>(&hs[c][k]);      // same address in every lane: one broadcast wavefront
          acc[c][0] = fmaf(w[k], h4.x, acc[c][0]); acc[c][1] = fmaf(w[k + 1], h4.y, acc[c][1]);
          acc[c][2] = fmaf(w[k + 2], h4.z, acc[c][2]); acc[c][3] = fmaf(w[k + 3], h4.w, acc[c][3]);
        }
      }
      if (tid < G3) {
#pragma unroll
        for (int c = 0; c < SB; ++c) ghs[c][tid] = (acc[c][0] + acc[c][1]) + (acc[c][2] + acc[c][3]);
      }
      __syncthreads();
    }
    if (pok) {
      float ghr = b_r, ghz = b_z, ghn = b_n;
      if (s > 0) { ghr += ghs[pbb][punit]; ghz += ghs[pbb][H + punit]; ghn += ghs[pbb][2 * H + punit]; }
      const float hprev = hs[pbb][punit];
      const float r = sigmoidf_(gir + ghr);
      const float z = sigmoidf_(giz + ghz);
      const float n = tanhf(gin + r * ghn);
      const float h = (1.f - z) * n + z * hprev;
      hs[pbb][punit] = h;
      const long long o = ((long long)pb * T + t) * row2H + dir * H + punit;
      p.out[o] = h;
      if (p.saved) {
        p.saved[o] = r; p.saved[p.saved_qstride + o] = z; p.saved[2 * p.saved_qstride + o] = n; p.saved[3 * p.saved_qstride + o] = ghn;
      }
    }
    __syncthreads();
  }
}

struct GruSmallBwdP {
  const float* dout; const float* out; const float* saved; long long saved_qstride; const float* whh[2];
  float* dgi; float* dgh; int B, T, H;
};

__global__ void __launch_bounds__(SMALL_NT) gru_small_bwd_kernel(const GruSmallBwdP p) {
  __shared__ __align__(16) float ds[SB][3 * HS];      // dgh of the tile: [clip][gate block g][unit], gate blocks HS apart
  __shared__ float dhp[3][SB][HS];                    // per-gate-block partial of W_hh^T dgh
  const int H = p.H, T = p.T;
  const int tid = threadIdx.x;
  const int dir = blockIdx.y, b0 = blockIdx.x * SB;
  const float* W = dir ? p.whh[1] : p.whh[0];         // [3H][H] as stored
  // product thread (g, j): partial_g[c][j] = sum_r' W_hh[g*H + r'][j] * dgh[c][g*H + r']
  const int g = tid / HS, j = tid - g * HS;
  float w[HS];
#pragma unroll
  for (int r = 0; r < HS; ++r) w[r] = (j < H && r < H) ? __ldg(W + ((long long)g * H + r) * H + j) : 0.f;
  for (int i = tid; i < SB * 3 * HS; i += SMALL_NT) (&ds[0][0])[i] = 0.f;
  for (int i = tid; i < 3 * SB * HS; i += SMALL_NT) (&dhp[0][0][0])[i] = 0.f;
  const long long row2H = 2ll * H;
  const int pbb = tid / H, punit = tid - pbb * H;
  const int pb = b0 + pbb;
  const bool pin = tid < SB * H;
  const bool pok = pin && pb < p.B;
  float dhz = 0.f;                                    // dh * z carried by the owner of the (clip, unit) pair
  // the six recurrence-independent operands of a step are prefetched one step ahead (see the forward kernel)
  float n_do = 0.f, n_r = 0.f, n_z = 0.f, n_n = 0.f, n_hn = 0.f, n_hp = 0.f;
  auto prefetch = [&](int t) {
    const int tp = dir == 0 ? t - 1 : t + 1;
    const long long o = ((long long)pb * T + t) * row2H + dir * H + punit;
    n_do = __ldg(p.dout + o); n_r = __ldg(p.saved + o); n_z = __ldg(p.saved + p.saved_qstride + o);
    n_n = __ldg(p.saved + 2 * p.saved_qstride + o); n_hn = __ldg(p.saved + 3 * p.saved_qstride + o);
    n_hp = (tp >= 0 && tp < T) ? __ldg(p.out + ((long long)pb * T + tp) * row2H + dir * H + punit) : 0.f;
  };
  if (pok) prefetch(dir == 0 ? T - 1 : 0);
  __syncthreads();
  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? T - 1 - s : s;
    if (pin) {
      float drp = 0.f, dzp = 0.f, dnr = 0.f;
      if (pok) {
        const float l_do = n_do, r = n_r, z = n_z, n = n_n, hn = n_hn, hprev = n_hp;
        if (s + 1 < T) prefetch(dir == 0 ? t - 1 : t + 1);
        const long long row = (long long)pb * T + t;
        const float dh = l_do + dhz + dhp[0][pbb][punit] + dhp[1][pbb][punit] + dhp[2][pbb][punit];
        const float dn = dh * (1.f - z) * (1.f - n * n);
        dzp = dh * (hprev - n) * z * (1.f - z);
        drp = dn * hn * r * (1.f - r);
        dnr = dn * r;
        dhz = dh * z;
        float* gp = p.dgi + row * 6 * H + dir * 3 * H + punit;
        gp[0] = drp; gp[H] = dzp; gp[2 * H] = dn;
        float* hp = p.dgh + row * 6 * H + dir * 3 * H + punit;
        hp[0] = drp; hp[H] = dzp; hp[2 * H] = dnr;
      }
      ds[pbb][punit] = drp; ds[pbb][HS + punit] = dzp; ds[pbb][2 * HS + punit] = dnr;
    }
    __syncthreads();
    if (s + 1 < T) {
      float acc[SB][4];                                  // independent partial sums (see the forward kernel)
#pragma unroll
      for (int c = 0; c < SB; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
#pragma unroll
      for (int r = 0; r < HS; r += 4) {
#pragma unroll
        for (int c = 0; c < SB; ++c) {
          const float4 d4 = *reinterpret_cast<const float4*>(&ds[c][g * HS + r]);   // warp-uniform address (one gate block per warp pair)
          acc[c][0] = fmaf(w[r], d4.x, acc[c][0]); acc[c][1] = fmaf(w[r + 1], d4.y, acc[c][1]);
          acc[c][2] = fmaf(w[r + 2], d4.z, acc[c][2]); acc[c][3] = fmaf(w[r + 3], d4.w, acc[c][3]);
        }
      }
#pragma unroll
      for (int c = 0; c < SB; ++c) dhp[g][c][j] = (acc[c][0] + acc[c][1]) + (acc[c][2] + acc[c][3]);
    }
    __syncthreads();
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(long long)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(long long)c * R + r] = tile[threadIdx.x][i];
  }
}

struct GruPlan { int u, UC, BT, ntiles, NB, tiles_per_cta; };

static int make_plan(int B, int H, GruPlan* pl, const char* name) {
  TG_REQUIRE(H > 0 && H <= 384 && B > 0, name);
  pl->UC = tg_ceil_div(H, UMAX);
  pl->u = tg_ceil_div(H, pl->UC);
  const int sms = tg_num_sms();
  int max_nb = sms / (2 * pl->UC);
  if (max_nb < 1) max_nb = 1;
  // smallest batch tile whose tile count fits the co-resident grid; otherwise the widest tile and several tiles per CTA
  int bt = 16;
  while (bt < 48 && tg_ceil_div(B, bt) > max_nb) bt += 16;
  pl->BT = bt;
  pl->ntiles = tg_ceil_div(B, bt);
  pl->NB = pl->ntiles < max_nb ? pl->ntiles : max_nb;
  pl->tiles_per_cta = tg_ceil_div(pl->ntiles, pl->NB);
  return 0;
}

template <typename K, typename P>
static int coop_launch(K kernel, const P& params, dim3 grid, size_t smem, cudaStream_t s, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("%s: smem attr (%zu B): %s", name, smem, cudaGetErrorString(e)); return -3; }
  void* args[] = {(void*)&params};
  // Cooperative launch = the runtime checks that all CTAs can be co-resident (they step together through L2 counters).
  // Stream capture does not accept cooperative launches: inside a CUDA graph the same kernel is launched normally - the
  // grid is sized to at most one CTA per SM, and no two of these stepping kernels ever run concurrently in our schedule.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &cap);
  if (cap != cudaStreamCaptureStatusNone) e = cudaLaunchKernel((const void*)kernel, grid, dim3(256), args, smem, s);
  else e = cudaLaunchCooperativeKernel((const void*)kernel, grid, dim3(256), args, smem, s);
  if (e != cudaSuccess) { tg_set_error("%s: cooperative launch grid (%u,%u,%u) smem %zu: %s", name, grid.x, grid.y, grid.z, smem,
                                       cudaGetErrorString(e)); return -2; }
  return 0;
}

}  // namespace

extern "C" int tg_transpose_f32(const float* in, float* out, int R, int C, tg_stream stream) {
  dim3 grid(tg_ceil_div(C, 32), tg_ceil_div(R, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, out, R, C);
  TG_CHECK_LAUNCH("tg_transpose_f32");
  return 0;
}

extern "C" int tg_gru_sync_ints(int B, int H) {
  GruPlan pl;
  if (make_plan(B, H, &pl, "tg_gru_sync_ints")) return -1;
  return 2 * pl.NB;
}

extern "C" int tg_gru_layer_fwd(const float* gi, const float* whhT_f, const float* whhT_r, const float* bhh_f, const float* bhh_r,
                                float* out, float* saved, long long saved_qstride, int* sync, int B, int T, int H, tg_stream stream) {
  TG_REQUIRE(gi && whhT_f && whhT_r && bhh_f && bhh_r && out && sync && T > 0, "tg_gru_layer_fwd");
  cudaStream_t s = (cudaStream_t)stream;
  if (H <= 64 && (H & 3) == 0) {
    GruSmallFwdP q;
    q.gi = gi; q.whhT[0] = whhT_f; q.whhT[1] = whhT_r; q.bhh[0] = bhh_f; q.bhh[1] = bhh_r; q.out = out; q.saved = saved;
    q.saved_qstride = saved_qstride; q.B = B; q.T = T; q.H = H;
    gru_small_fwd_kernel<<<dim3(tg_ceil_div(B, SB), 2), SMALL_NT, 0, s>>>(q);
    TG_CHECK_LAUNCH("tg_gru_layer_fwd(small)");
    return 0;
  }
  GruPlan pl;
  if (make_plan(B, H, &pl, "tg_gru_layer_fwd")) return -1;
  cudaError_t e = cudaMemsetAsync(sync, 0, sizeof(int) * 2 * pl.NB, s);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_fwd: memset: %s", cudaGetErrorString(e)); return -2; }
  GruFwdP p;
  p.gi = gi; p.whhT[0] = whhT_f; p.whhT[1] = whhT_r; p.bhh[0] = bhh_f; p.bhh[1] = bhh_r;
  p.out = out; p.saved = saved; p.saved_qstride = saved_qstride; p.sync = sync;
  p.B = B; p.T = T; p.H = H; p.u = pl.u; p.UC = pl.UC; p.NB = pl.NB; p.ntiles = pl.ntiles;
  dim3 grid(pl.UC, pl.NB, 2);
  const size_t smem = ((size_t)H * RP + (size_t)(H > RP ? H : RP) * (pl.BT + 1)) * sizeof(float);
  TG_REQUIRE(smem <= (size_t)tg_max_smem_optin(), "tg_gru_layer_fwd");
  if (pl.BT == 16) return coop_launch(gru_fwd_kernel<16>, p, grid, smem, s, "tg_gru_layer_fwd");
  if (pl.BT == 32) return coop_launch(gru_fwd_kernel<32>, p, grid, smem, s, "tg_gru_layer_fwd");
  return coop_launch(gru_fwd_kernel<48>, p, grid, smem, s, "tg_gru_layer_fwd");
}

extern "C" size_t tg_gru_bwd_scratch_floats(int B, int H) {
  GruPlan pl;
  if (make_plan(B, H, &pl, "tg_gru_bwd_scratch_floats")) return 0;
  const int HP = (H + 3) & ~3;
  return (size_t)2 * 2 * B * pl.UC * HP;
}

extern "C" int tg_gru_layer_bwd(const float* dout, const float* out, const float* saved, long long saved_qstride,
                                const float* whh_f, const float* whh_r, float* dgi, float* dgh, float* partial, int* sync,
                                int B, int T, int H, tg_stream stream) {
  TG_REQUIRE(dout && out && saved && whh_f && whh_r && dgi && dgh && partial && sync && T > 0, "tg_gru_layer_bwd");
  cudaStream_t s = (cudaStream_t)stream;
  if (H <= 64 && (H & 3) == 0) {
    GruSmallBwdP q;
    q.dout = dout; q.out = out; q.saved = saved; q.saved_qstride = saved_qstride; q.whh[0] = whh_f; q.whh[1] = whh_r;
    q.dgi = dgi; q.dgh = dgh; q.B = B; q.T = T; q.H = H;
    gru_small_bwd_kernel<<<dim3(tg_ceil_div(B, SB), 2), SMALL_NT, 0, s>>>(q);
    TG_CHECK_LAUNCH("tg_gru_layer_bwd(small)");
    return 0;
  }
  GruPlan pl;
  if (make_plan(B, H, &pl, "tg_gru_layer_bwd")) return -1;
  cudaError_t e = cudaMemsetAsync(sync, 0, sizeof(int) * 2 * pl.NB, s);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_bwd: memset: %s", cudaGetErrorString(e)); return -2; }
  GruBwdP p;
  p.dout = dout; p.out = out; p.saved = saved; p.saved_qstride = saved_qstride; p.whh[0] = whh_f; p.whh[1] = whh_r;
  p.dgi = dgi; p.dgh = dgh; p.partial = partial; p.sync = sync;
  p.B = B; p.T = T; p.H = H; p.u = pl.u; p.UC = pl.UC; p.NB = pl.NB; p.ntiles = pl.ntiles;
  p.HP = (H + 3) & ~3; p.tiles_per_cta = pl.tiles_per_cta;
  dim3 grid(pl.UC, pl.NB, 2);
  const size_t smem = ((size_t)RP * p.HP + (size_t)RP * (pl.BT + 1) + (size_t)pl.tiles_per_cta * pl.BT * pl.u) * sizeof(float);
  TG_REQUIRE(smem <= (size_t)tg_max_smem_optin(), "tg_gru_layer_bwd");
  if (pl.BT == 16) return coop_launch(gru_bwd_kernel<16>, p, grid, smem, s, "tg_gru_layer_bwd");
  if (pl.BT == 32) return coop_launch(gru_bwd_kernel<32>, p, grid, smem, s, "tg_gru_layer_bwd");
  return coop_launch(gru_bwd_kernel<48>, p, grid, smem, s, "tg_gru_layer_bwd");
}
