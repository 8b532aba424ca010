// ConvDiscriminator recurrent stack as ONE launch per pass (multimodal_context_net.py:221-222,241-251):
//   4-layer bidirectional GRU(8 -> 64) over the 28 conv frames, inter-layer dropout, sum of the two directions, Linear(64,1) per frame,
//   Linear(28,1) over the frames, sigmoid.
// Clips are independent of each other through this whole stack, so a CTA owns ONE clip for all layers and both directions (B CTAs, one
// wave for 128 clips) and nothing is exchanged between CTAs; the layer-to-layer activations (28 x 128 floats) never leave shared memory.
// Per layer: (1) input projection gi = W_ih x + b_ih for all 28 frames (thread = gate row of one direction, weights in registers, 64 at a
// time), (2) the recurrence with W_hh in registers (same scheme as gru_small_*_kernel in gru.cu).  Everything the backward needs is
// written in the layouts the per-layer kernels use (layer outputs, r / z / n / W_hn h planes, masked layer inputs, hsum, per-frame head).
// Replaces, per pass, 4 x [projection GEMM + recurrence + dropout multiply] + sum + 2 head GEMMs = 15 launches of 3-25 us each.
#include "common.cuh"

// Round-2 profile of the first version (ncu --set full, profiles/r02_ncu_dfused_v1_*.txt): 279 us forward / 326 us backward per pass.
// Two things cost almost all of it and neither was arithmetic: (1) every thread fetched ITS weight row straight from global memory -
// 32 sectors per warp instruction, 30 sectors per request, the load/store queue saturated (stall_lg); (2) the backward's data gradient
// streamed W_ih from L2 four loads at a time with the latency exposed 96 times per layer (stall_long_sb, 62 % of all instructions), and
// the recurrences waited on one-step-ahead global prefetches at every step.  Now: weights are staged through shared memory with
// cp.async (coalesced 16-byte chunks, XOR-swizzled so that "thread = row" reads are conflict-free) one stage AHEAD of their use, and
// everything a recurrence step reads (masks; saved gate planes and layer outputs in the backward) is bulk-copied to shared memory before
// the first step, so the step loop touches no global memory except its (fire-and-forget) stores.
namespace {

constexpr int H = 64, G3 = 192, NT = 384, MAXL = 4, MAXT = 32, WROW = 64;

// FAST (the tensor-core arithmetic mode, tolerance 1e-2): ex2.approx / rcp.approx forms, as in the generator's recurrence (gru_cl.cu) - the
// precise expf / IEEE division / tanhf are ~100 dependent instructions per step on the gate warps, half of the step's latency.
template <bool FAST> __device__ __forceinline__ float sigmoidf_(float x) {
  return FAST ? __fdividef(1.f, 1.f + __expf(-x)) : 1.f / (1.f + expf(-x));
}
template <bool FAST> __device__ __forceinline__ float tanhf_(float x) {
  return FAST ? 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)) : tanhf(x);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// rows [0, 384) x features [k0, k0 + kn) of a row-major [384][K] weight block -> wst (row pitch 64 floats; the 16-byte chunk c of row r
// sits at chunk position c ^ (r & 15), so the 8 threads of a quarter warp reading "their" rows hit 8 different bank groups)
__device__ __forceinline__ void stage_rows(float* wst, const float* W, int K, int k0, int kn) {
  const int cpr = kn >> 2;
  for (int i = threadIdx.x; i < 2 * G3 * cpr; i += NT) {
    const int row = i / cpr, c = i - row * cpr;
    cp_async16(wst + row * WROW + 4 * (c ^ (row & 15)), W + (long long)row * K + k0 + 4 * c);
  }
}
// this thread's row (row = threadIdx.x) of the staged block -> registers (features >= kn read as zero)
__device__ __forceinline__ void load_row(const float* wst, int kn, float (&w)[64]) {
  const int row = threadIdx.x;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (4 * c < kn) v = *reinterpret_cast<const float4*>(wst + row * WROW + 4 * (c ^ (row & 15)));
    w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
  }
}
// n floats (n % 4 == 0, both pointers 16-byte aligned) global -> shared, asynchronously
__device__ __forceinline__ void stage_flat(float* dst, const float* src, int n) {
  for (int i = threadIdx.x * 4; i < n; i += NT * 4) cp_async16(dst + i, src + i);
}

struct StackFwdP {
  const float* x;                 // [B, T, I0]
  const float* params;            // the GRU parameter block of the flat arena (gru_arena_order)
  const float* mask[MAXL];        // [B*T, 2H] keep-masks (already scaled) applied to the output of layer l < L-1, or NULL
  float* out[MAXL];               // [B*T, 2H] layer outputs
  float* saved[MAXL];             // 4 planes [B*T, 2H] (r, z, n, W_hn h + b_hn), plane stride saved_qstride; or NULL
  float* drop[MAXL];              // [B*T, 2H] masked outputs (= input of layer l+1), l < L-1; or NULL (no mask: the next layer reads out[l])
  long long saved_qstride;
  const float* w_out; const float* b_out; const float* w_out2; const float* b_out2;
  float* hsum; float* o1; float* prob;
  int B, T, I0, L;
};

// offsets (floats) of layer l's tensors inside the arena block: [W_ih f | W_ih r | b_ih f | b_ih r | W_hh f | W_hh r | b_hh f | b_hh r]
__host__ __device__ inline long long layer_base(int l, int I0) {
  const long long l0 = 2ll * G3 * I0 + 2 * G3 + 2ll * G3 * H + 2 * G3;
  const long long ln = 2ll * G3 * (2 * H) + 2 * G3 + 2ll * G3 * H + 2 * G3;
  return l == 0 ? 0 : l0 + (long long)(l - 1) * ln;
}

template <bool FAST>
__global__ void __launch_bounds__(NT, 1) dgru_stack_fwd_kernel(const StackFwdP p) {
  extern __shared__ __align__(16) float sm[];
  const int T = p.T;
  float* gi = sm;                              // [2 dirs][T][192]
  float* xin = gi + 2 * T * G3;                // [2 buffers][T][128]  layer input (layer 0: [T][I0])
  float* hs = xin + 2 * T * 2 * H;             // [2][64]
  float* ghs = hs + 2 * H;                     // [2][192]
  float* o1s = ghs + 2 * G3;                   // [32]
  float* mks = o1s + MAXT;                     // [T][128]  dropout mask of the current layer's output
  float* wst = mks + T * 2 * H;                // [384][64] weight staging (swizzled)
  const int tid = threadIdx.x;
  const int d = tid / G3, r = tid - d * G3;    // direction, gate row  (tid == row of the stacked [2][192][.] weight blocks)
  const int b = blockIdx.x;
  const int lt = r;                            // gate phase: threads r < 64 of each direction own hidden unit r
  const long long row0 = (long long)b * T;

  stage_flat(xin, p.x + row0 * p.I0, T * p.I0);
  stage_rows(wst, p.params, p.I0, 0, min(64, p.I0));          // layer 0, W_ih chunk 0
  cp_async_commit();

  for (int l = 0; l < p.L; ++l) {
    const int K = l == 0 ? p.I0 : 2 * H;
    const float* base = p.params + layer_base(l, p.I0);
    const float* bih = base + 2ll * G3 * K + d * G3;
    const float* whh0 = base + 2ll * G3 * K + 2 * G3;          // [2][192][64]
    const float* bhh = base + 2ll * G3 * K + 2 * G3 + 2ll * G3 * H + d * G3;
    const float* xi = xin + (l & 1) * T * 2 * H;           // this layer's input  [T][K]
    float* xo = xin + ((l + 1) & 1) * T * 2 * H;           // next layer's input  [T][128]
    float* gid = gi + d * T * G3;
    const float* mk = (l + 1 < p.L) ? p.mask[l] : nullptr;

    // ---- (1) input projection, 64 input features at a time; the NEXT weight block is in flight while this one is multiplied
    float w[64];
    for (int k0 = 0; k0 < K; k0 += 64) {
      const int kn = min(64, K - k0);
      cp_async_wait_all();
      __syncthreads();                                     // this chunk (and, first time round, x) has landed for every thread
      load_row(wst, kn, w);
      __syncthreads();                                     // wst is free again
      if (k0 + 64 < K) {
        stage_rows(wst, base, K, k0 + 64, min(64, K - k0 - 64));
      } else {
        stage_rows(wst, whh0, H, 0, H);
        if (mk) stage_flat(mks, mk + row0 * 2 * H, T * 2 * H);
      }
      cp_async_commit();
      const float bias = k0 == 0 ? __ldg(bih + r) : 0.f;
      for (int t0 = 0; t0 < T; t0 += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        // fully unrolled with a (uniform) guard on the feature count: a run-time index into w[] would push the array to local memory
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          if (j < kn) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (t0 + i < T) {
                const float4 x4 = *reinterpret_cast<const float4*>(xi + (t0 + i) * K + k0 + j);      // same address in every lane: broadcast
                acc[i] = fmaf(w[j], x4.x, acc[i]); acc[i] = fmaf(w[j + 1], x4.y, acc[i]);
                acc[i] = fmaf(w[j + 2], x4.z, acc[i]); acc[i] = fmaf(w[j + 3], x4.w, acc[i]);
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (t0 + i < T) gid[(t0 + i) * G3 + r] = (k0 == 0 ? bias : gid[(t0 + i) * G3 + r]) + acc[i];
      }
    }
    // ---- (2) recurrence: W_hh row in registers, masks in shared memory
    cp_async_wait_all();
    __syncthreads();                                       // W_hh and the masks have landed; gi complete
    load_row(wst, H, w);
    float b_r = 0.f, b_z = 0.f, b_n = 0.f;
    if (lt < H) { b_r = __ldg(bhh + lt); b_z = __ldg(bhh + H + lt); b_n = __ldg(bhh + 2 * H + lt); hs[d * H + lt] = 0.f; }
    __syncthreads();                                       // wst free, hs zeroed
    if (l + 1 < p.L) {                                     // next layer's first W_ih chunk travels under this layer's recurrence
      stage_rows(wst, p.params + layer_base(l + 1, p.I0), 2 * H, 0, 64);
      cp_async_commit();
    }
    for (int s = 0; s < T; ++s) {
      const int t = d == 0 ? s : T - 1 - s;
      if (s > 0) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int k = 0; k < H; k += 4) {
          const float4 h4 = *reinterpret_cast<const float4*>(hs + d * H + k);
          a0 = fmaf(w[k], h4.x, a0); a1 = fmaf(w[k + 1], h4.y, a1); a2 = fmaf(w[k + 2], h4.z, a2); a3 = fmaf(w[k + 3], h4.w, a3);
        }
        ghs[d * G3 + r] = (a0 + a1) + (a2 + a3);
        __syncthreads();
      }
      if (lt < H) {
        float ghr = b_r, ghz = b_z, ghn = b_n;
        if (s > 0) { ghr += ghs[d * G3 + lt]; ghz += ghs[d * G3 + H + lt]; ghn += ghs[d * G3 + 2 * H + lt]; }
        const float hprev = hs[d * H + lt];
        const float rg = sigmoidf_<FAST>(gid[t * G3 + lt] + ghr);
        const float zg = sigmoidf_<FAST>(gid[t * G3 + H + lt] + ghz);
        const float ng = tanhf_<FAST>(gid[t * G3 + 2 * H + lt] + rg * ghn);
        const float h = (1.f - zg) * ng + zg * hprev;
        hs[d * H + lt] = h;
        const long long o = (row0 + t) * 2 * H + d * H + lt;
        p.out[l][o] = h;
        if (p.saved[l]) {
          float* sv = p.saved[l];
          sv[o] = rg; sv[p.saved_qstride + o] = zg; sv[2 * p.saved_qstride + o] = ng; sv[3 * p.saved_qstride + o] = ghn;
        }
        const float hm = mk ? h * mks[t * 2 * H + d * H + lt] : h;
        xo[t * 2 * H + d * H + lt] = hm;
        if (mk && p.drop[l]) p.drop[l][o] = hm;
      }
      __syncthreads();
    }
  }
  // ---- heads: hsum[t][j] = fwd + rev, o1[t] = hsum[t] . w_out + b_out, prob = sigmoid(o1 . w_out2 + b_out2)
  const float* xl = xin + (p.L & 1) * T * 2 * H;           // output of the last layer (unmasked)
  const int warp = tid >> 5, lane = tid & 31;
  for (int t = warp; t < T; t += NT / 32) {
    const float s0 = xl[t * 2 * H + lane] + xl[t * 2 * H + H + lane];
    const float s1 = xl[t * 2 * H + 32 + lane] + xl[t * 2 * H + H + 32 + lane];
    p.hsum[(row0 + t) * H + lane] = s0;
    p.hsum[(row0 + t) * H + 32 + lane] = s1;
    float v = warp_sum(s0 * __ldg(p.w_out + lane) + s1 * __ldg(p.w_out + 32 + lane));
    if (lane == 0) { v += __ldg(p.b_out); o1s[t] = v; p.o1[row0 + t] = v; }
  }
  __syncthreads();
  if (warp == 0) {
    float v = lane < T ? o1s[lane] * __ldg(p.w_out2 + lane) : 0.f;
    v = warp_sum(v);
    if (lane == 0) p.prob[b] = sigmoidf_<false>(v + __ldg(p.b_out2));
  }
}

size_t fwd_smem_bytes(int T) { return ((size_t)2 * T * G3 + (size_t)2 * T * 2 * H + 2 * H + 2 * G3 + MAXT + (size_t)T * 2 * H + 2 * G3 * WROW) * sizeof(float); }

}  // namespace

extern "C" int tg_dgru_stack_fwd(const float* x, const float* gru_params, const float* const* masks, float* const* outs, float* const* saved,
                                 long long saved_qstride, float* const* drops, const float* w_out, const float* b_out, const float* w_out2,
                                 const float* b_out2, float* hsum, float* o1, float* prob, int B, int T, int I0, int Hh, int L, int fast,
                                 tg_stream stream) {
  TG_REQUIRE(x && gru_params && outs && w_out && b_out && w_out2 && b_out2 && hsum && o1 && prob, "tg_dgru_stack_fwd");
  TG_REQUIRE(Hh == H && L >= 1 && L <= MAXL && T >= 1 && T <= MAXT && I0 >= 4 && I0 <= 32 && (I0 & 3) == 0 && B > 0, "tg_dgru_stack_fwd(shape)");
  TG_REQUIRE((((uintptr_t)x | (uintptr_t)gru_params) & 15) == 0, "tg_dgru_stack_fwd(16-byte alignment)");
  StackFwdP p;
  p.x = x; p.params = gru_params; p.saved_qstride = saved_qstride;
  for (int l = 0; l < MAXL; ++l) {
    p.mask[l] = (masks && l < L - 1) ? masks[l] : nullptr;
    p.out[l] = l < L ? outs[l] : nullptr;
    p.saved[l] = (saved && l < L) ? saved[l] : nullptr;
    p.drop[l] = (drops && l < L - 1) ? drops[l] : nullptr;
    TG_REQUIRE(l >= L || p.out[l], "tg_dgru_stack_fwd(out)");
    TG_REQUIRE((((uintptr_t)p.mask[l]) & 15) == 0, "tg_dgru_stack_fwd(mask alignment)");
  }
  p.w_out = w_out; p.b_out = b_out; p.w_out2 = w_out2; p.b_out2 = b_out2; p.hsum = hsum; p.o1 = o1; p.prob = prob;
  p.B = B; p.T = T; p.I0 = I0; p.L = L;
  const size_t smem = fwd_smem_bytes(T);
  auto kern = fast ? dgru_stack_fwd_kernel<true> : dgru_stack_fwd_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("tg_dgru_stack_fwd: smem attr: %s", cudaGetErrorString(e)); return -3; }
  kern<<<B, NT, smem, (cudaStream_t)stream>>>(p);
  TG_CHECK_LAUNCH("tg_dgru_stack_fwd");
  return 0;
}

// =====================================================================================================================
// backward of the same stack: heads -> 4 x [recurrence backward + data gradient through W_ih (+ dropout mask)] in one launch.
// Weight gradients of the recurrent layers stay with the tensor-core wgrad kernels (they need dgi / dgh of ALL clips and run on the
// weight-gradient stream, off the critical path): this kernel writes dgi / dgh per layer for them.  The four head gradients are summed
// with atomics (94 floats per clip).
// =====================================================================================================================
namespace {

struct StackBwdP {
  const float* dlogit;            // [B]  d loss / d (pre-sigmoid output)
  const float* params;
  const float* mask[MAXL];        // mask[l] multiplies the output of layer l (l < L-1), or NULL
  const float* out[MAXL];         // layer outputs [B*T, 2H]
  const float* saved[MAXL];       // gate planes
  long long saved_qstride;
  const float* hsum; const float* o1; const float* w_out; const float* w_out2;
  float* dgi[MAXL]; float* dgh[MAXL];      // [B*T, 6H] per layer
  float* dx0;                     // [B*T, I0] gradient w.r.t. the stack input, or NULL
  float* g_w_out; float* g_b_out; float* g_w_out2; float* g_b_out2;
  int B, T, I0, L;
};

constexpr int DXR = 64;                        // W_ih rows per streamed chunk of the data-gradient phase

__global__ void __launch_bounds__(NT, 1) dgru_stack_bwd_kernel(const StackBwdP p) {
  extern __shared__ __align__(16) float sm[];
  const int T = p.T;
  float* dcur = sm;                            // [T][128]  gradient w.r.t. the current layer's output (both directions)
  float* dgs = dcur + T * 2 * H;               // [T][384]  dgi of the current layer, both directions (operand of the W_ih data gradient)
  float* ds = dgs + T * 2 * G3;                // [2][192]  dgh of the current step
  float* dhp = ds + 2 * G3;                    // [2][3][64] per-gate-block partials of W_hh^T dgh
  float* do1 = dhp + 2 * 3 * H;                // [32]
  float* svs = do1 + MAXT;                     // [4][T][128] saved gate planes of the current layer (r, z, n, W_hn h + b_hn)
  float* ous = svs + 4 * T * 2 * H;            // [T][128]  outputs of the current layer (h_prev of every step)
  float* wst = ous + T * 2 * H;                // [2][DXR][128] streamed W_ih chunks (data-gradient phase)
  const int tid = threadIdx.x;
  const int d = tid / G3, r = tid - d * G3;
  const int g = r / H, j = r - g * H;          // product thread (gate block g, output unit j)
  const int lt = r;                            // gate phase: r < 64 owns hidden unit r of direction d
  const int b = blockIdx.x;
  const long long row0 = (long long)b * T;

  // the top layer's planes start travelling before the head arithmetic
  {
    const int l = p.L - 1;
    for (int q = 0; q < 4; ++q) stage_flat(svs + q * T * 2 * H, p.saved[l] + q * p.saved_qstride + row0 * 2 * H, T * 2 * H);
    stage_flat(ous, p.out[l] + row0 * 2 * H, T * 2 * H);
    cp_async_commit();
  }
  // ---- heads: d o1[t] = dlogit * w2[t];  d hsum[t][j] = d o1[t] * w_out[j]  (the same for both directions)
  const float dl = __ldg(p.dlogit + b);
  if (tid < T) {
    const float v = dl * __ldg(p.w_out2 + tid);
    do1[tid] = v;
    atomicAdd(p.g_w_out2 + tid, dl * __ldg(p.o1 + row0 + tid));
    atomicAdd(p.g_b_out, v);
  }
  if (tid == 0) atomicAdd(p.g_b_out2, dl);
  __syncthreads();
  for (int i = tid; i < T * 2 * H; i += NT) {
    const int t = i / (2 * H), c = i - t * 2 * H;
    dcur[i] = do1[t] * __ldg(p.w_out + (c & (H - 1)));
  }
  if (tid < H) {
    float a = 0.f;
    for (int t = 0; t < T; ++t) a = fmaf(do1[t], __ldg(p.hsum + (row0 + t) * H + tid), a);
    atomicAdd(p.g_w_out + tid, a);
  }

  for (int l = p.L - 1; l >= 0; --l) {
    const int K = l == 0 ? p.I0 : 2 * H;
    const float* base = p.params + layer_base(l, p.I0);
    const float* wih = base;                             // [2][192][K] = 384 rows of K
    const float* whh = base + 2ll * G3 * K + 2 * G3 + (long long)d * G3 * H;
    const int nchunk = l == 0 ? 1 : 2 * G3 / DXR;        // layer 0: the whole [384][I0] block is one chunk (I0 <= 32)
    // ---- recurrence backward: thread (g, j) keeps W_hh[g*64 + r'][j], r' < 64, of its direction in registers (coalesced over j)
    float w[64];
#pragma unroll
    for (int q = 0; q < 64; ++q) w[q] = __ldg(whh + ((long long)g * H + q) * H + j);
    for (int i = tid; i < 2 * 3 * H; i += NT) dhp[i] = 0.f;
    // first W_ih chunk of the data-gradient phase travels under the recurrence
    stage_flat(wst, wih, l == 0 ? 2 * G3 * K : DXR * K);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();                                     // planes, outputs, dcur, dhp ready
    float dhz = 0.f;
    for (int s = 0; s < T; ++s) {
      const int t = d == 0 ? T - 1 - s : s;
      if (lt < H) {
        const int o = t * 2 * H + d * H + lt;
        const int tp = d == 0 ? t - 1 : t + 1;
        const float rg = svs[o], zg = svs[T * 2 * H + o], ng = svs[2 * T * 2 * H + o], hn = svs[3 * T * 2 * H + o];
        const float hprev = (tp >= 0 && tp < T) ? ous[tp * 2 * H + d * H + lt] : 0.f;
        const float dh = dcur[o] + dhz + dhp[(d * 3 + 0) * H + lt] + dhp[(d * 3 + 1) * H + lt] + dhp[(d * 3 + 2) * H + lt];
        const float dn = dh * (1.f - zg) * (1.f - ng * ng);
        const float dzp = dh * (hprev - ng) * zg * (1.f - zg);
        const float drp = dn * hn * rg * (1.f - rg);
        const float dnr = dn * rg;
        dhz = dh * zg;
        const long long o6 = (row0 + t) * 6 * H + d * 3 * H + lt;
        float* gp = p.dgi[l] + o6; float* hp = p.dgh[l] + o6;
        gp[0] = drp; gp[H] = dzp; gp[2 * H] = dn;
        hp[0] = drp; hp[H] = dzp; hp[2 * H] = dnr;
        ds[d * G3 + lt] = drp; ds[d * G3 + H + lt] = dzp; ds[d * G3 + 2 * H + lt] = dnr;
        float* gs = dgs + t * 2 * G3 + d * G3 + lt;
        gs[0] = drp; gs[H] = dzp; gs[2 * H] = dn;
      }
      __syncthreads();
      if (s + 1 < T) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int q = 0; q < H; q += 4) {
          const float4 d4 = *reinterpret_cast<const float4*>(ds + d * G3 + g * H + q);
          a0 = fmaf(w[q], d4.x, a0); a1 = fmaf(w[q + 1], d4.y, a1); a2 = fmaf(w[q + 2], d4.z, a2); a3 = fmaf(w[q + 3], d4.w, a3);
        }
        dhp[(d * 3 + g) * H + j] = (a0 + a1) + (a2 + a3);
      }
      __syncthreads();
    }
    // the layer below: its planes and outputs travel under the data-gradient phase (svs / ous are dead after the last step)
    if (l > 0) {
      for (int q = 0; q < 4; ++q) stage_flat(svs + q * T * 2 * H, p.saved[l - 1] + q * p.saved_qstride + row0 * 2 * H, T * 2 * H);
      stage_flat(ous, p.out[l - 1] + row0 * 2 * H, T * 2 * H);
    }
    cp_async_commit();
    // ---- data gradient through W_ih: dx[t][k] = sum over both directions and the 192 gate rows of dgi[t][d][r] * W_ih[d][r][k], times the
    // dropout mask of the layer below.
    if (l > 0) {
      // thread = (input feature k, one third of the frames); W_ih streams through shared memory in 64-row chunks, double-buffered
      const int k = tid & 127, tg = tid >> 7;            // 128 features x 3 frame groups
      const int t_lo = tg * ((T + 2) / 3), t_hi = min(T, t_lo + (T + 2) / 3);
      constexpr int TG = 11;                             // frames per thread (T <= 32 -> ceil(32 / 3))
      float acc[TG];
#pragma unroll
      for (int i = 0; i < TG; ++i) acc[i] = 0.f;
      for (int ch = 0; ch < nchunk; ++ch) {
        if (ch > 0) cp_async_wait_all();                 // chunk 0 landed before the recurrence; the planes may stay in flight
        __syncthreads();                                 // chunk ch visible to all; everyone is done with the other buffer
        if (ch + 1 < nchunk) {
          stage_flat(wst + ((ch + 1) & 1) * DXR * 2 * H, wih + (long long)(ch + 1) * DXR * K, DXR * K);
          cp_async_commit();
        }
        const float* Ws = wst + (ch & 1) * DXR * 2 * H;
        for (int rr = 0; rr < DXR; rr += 4) {
          const float w0 = Ws[rr * 2 * H + k], w1 = Ws[(rr + 1) * 2 * H + k], w2 = Ws[(rr + 2) * 2 * H + k], w3 = Ws[(rr + 3) * 2 * H + k];
#pragma unroll
          for (int i = 0; i < TG; ++i) {
            if (t_lo + i < t_hi) {
              const float4 g4 = *reinterpret_cast<const float4*>(dgs + (t_lo + i) * 2 * G3 + ch * DXR + rr);      // broadcast
              acc[i] = fmaf(g4.x, w0, acc[i]); acc[i] = fmaf(g4.y, w1, acc[i]); acc[i] = fmaf(g4.z, w2, acc[i]); acc[i] = fmaf(g4.w, w3, acc[i]);
            }
          }
        }
      }
      __syncthreads();                                   // every thread has finished reading dcur / dgs of this layer
      const float* mk = p.mask[l - 1];
#pragma unroll
      for (int i = 0; i < TG; ++i) {
        const int t = t_lo + i;
        if (t < t_hi) dcur[t * 2 * H + k] = mk ? acc[i] * __ldg(mk + (row0 + t) * 2 * H + k) : acc[i];
      }
    } else {
      // layer 0: [T][384] x [384][I0], thread = (frame, input feature)
      cp_async_wait_all();
      __syncthreads();
      if (p.dx0) {
        for (int idx = tid; idx < T * K; idx += NT) {
          const int t = idx / K, k = idx - t * K;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          for (int rr = 0; rr < 2 * G3; rr += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(dgs + t * 2 * G3 + rr);
            a0 = fmaf(g4.x, wst[rr * K + k], a0); a1 = fmaf(g4.y, wst[(rr + 1) * K + k], a1);
            a2 = fmaf(g4.z, wst[(rr + 2) * K + k], a2); a3 = fmaf(g4.w, wst[(rr + 3) * K + k], a3);
          }
          p.dx0[(row0 + t) * K + k] = (a0 + a1) + (a2 + a3);
        }
      }
    }
    // the next layer's top-of-loop __syncthreads (after its cp.async wait) orders these dcur writes before the recurrence reads them
  }
}

size_t bwd_smem_bytes(int T) {
  return ((size_t)T * 2 * H + (size_t)T * 2 * G3 + 2 * G3 + 2 * 3 * H + MAXT + (size_t)5 * T * 2 * H + (size_t)2 * DXR * 2 * H) * sizeof(float);
}

}  // namespace

extern "C" int tg_dgru_stack_bwd(const float* dlogit, const float* gru_params, const float* const* masks, const float* const* outs,
                                 const float* const* saved, long long saved_qstride, const float* hsum, const float* o1, const float* w_out,
                                 const float* w_out2, float* const* dgi, float* const* dgh, float* dx0, float* g_w_out, float* g_b_out,
                                 float* g_w_out2, float* g_b_out2, int B, int T, int I0, int Hh, int L, tg_stream stream) {
  TG_REQUIRE(dlogit && gru_params && outs && saved && hsum && o1 && w_out && w_out2 && dgi && dgh && g_w_out && g_b_out && g_w_out2 && g_b_out2,
             "tg_dgru_stack_bwd");
  TG_REQUIRE(Hh == H && L >= 1 && L <= MAXL && T >= 1 && T <= MAXT && I0 >= 4 && I0 <= 32 && (I0 & 3) == 0 && B > 0, "tg_dgru_stack_bwd(shape)");
  TG_REQUIRE((((uintptr_t)gru_params | (uintptr_t)saved_qstride * 4) & 15) == 0, "tg_dgru_stack_bwd(16-byte alignment)");
  StackBwdP p;
  p.dlogit = dlogit; p.params = gru_params; p.saved_qstride = saved_qstride;
  for (int l = 0; l < MAXL; ++l) {
    p.mask[l] = (masks && l < L - 1) ? masks[l] : nullptr;
    p.out[l] = l < L ? outs[l] : nullptr;
    p.saved[l] = l < L ? saved[l] : nullptr;
    p.dgi[l] = l < L ? dgi[l] : nullptr;
    p.dgh[l] = l < L ? dgh[l] : nullptr;
    TG_REQUIRE(l >= L || (p.out[l] && p.saved[l] && p.dgi[l] && p.dgh[l]), "tg_dgru_stack_bwd(buffers)");
    TG_REQUIRE((((uintptr_t)p.out[l] | (uintptr_t)p.saved[l] | (uintptr_t)p.mask[l]) & 15) == 0, "tg_dgru_stack_bwd(buffer alignment)");
  }
  p.hsum = hsum; p.o1 = o1; p.w_out = w_out; p.w_out2 = w_out2; p.dx0 = dx0;
  p.g_w_out = g_w_out; p.g_b_out = g_b_out; p.g_w_out2 = g_w_out2; p.g_b_out2 = g_b_out2;
  p.B = B; p.T = T; p.I0 = I0; p.L = L;
  const size_t smem = bwd_smem_bytes(T);
  cudaError_t e = cudaFuncSetAttribute(dgru_stack_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("tg_dgru_stack_bwd: smem attr: %s", cudaGetErrorString(e)); return -3; }
  dgru_stack_bwd_kernel<<<B, NT, smem, (cudaStream_t)stream>>>(p);
  TG_CHECK_LAUNCH("tg_dgru_stack_bwd");
  return 0;
}
