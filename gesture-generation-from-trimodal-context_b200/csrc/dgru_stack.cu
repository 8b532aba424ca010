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

constexpr int H = 64, G3 = 192, NT = 384, MAXL = 4, MAXT = 32;
constexpr int WCH = 32;                        // features per staged weight chunk (forward)
constexpr int WBUF = 2 * G3 * WCH;             // floats per chunk buffer: [384 rows][32]
constexpr int XS = 2 * H + 4;                  // row pitch of a 128-wide layer input in shared memory: (4 g + t) % 32 distinct over a warp's
                                               // MMA fragment loads (g = lane / 4, t = lane % 4)

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
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// fp32 -> nearest TF32, as the b32 operand of mma.sync
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// D (16x8, fp32) += A (16x8, row) * B (8x8, col), TF32 operands: the warp-level tensor-core instruction (one clip's GEMMs are far too
// small for a tcgen05 tile - M = 28 frames - and this kernel's CTAs own one clip each)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// rows [0, 384) x features [k0, k0 + kn) (kn <= 32) of a row-major [384][K] weight block -> buf ([384][32] floats; the 16-byte chunk c of
// row r sits at chunk position c ^ (r & 7): "thread = row" float4 reads and the MMA's B-fragment reads are both bank-conflict free)
__device__ __forceinline__ void stage_rows(float* buf, const float* W, int K, int k0, int kn) {
  const int cpr = kn >> 2;
  for (int i = threadIdx.x; i < 2 * G3 * cpr; i += NT) {
    const int row = i / cpr, c = i - row * cpr;
    cp_async16(buf + row * WCH + 4 * (c ^ (row & 7)), W + (long long)row * K + k0 + 4 * c);
  }
}
// this thread's row (row = threadIdx.x) of a staged chunk -> 32 registers (features >= kn read as zero)
__device__ __forceinline__ void load_row32(const float* buf, int kn, float* w) {
  const int row = threadIdx.x;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (4 * c < kn) v = *reinterpret_cast<const float4*>(buf + row * WCH + 4 * (c ^ (row & 7)));
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
  float* xin = gi + 2 * T * G3;                // [2 buffers][32][XS]  layer input (layer 0: [T][I0], tight)
  float* hs = xin + 2 * MAXT * XS;             // [2][64]
  float* ghs = hs + 2 * H;                     // [2][192]
  float* o1s = ghs + 2 * G3;                   // [32]
  float* mks = o1s + MAXT;                     // [T][128]  dropout mask of the current layer's output
  float* wst = mks + T * 2 * H;                // [2][384][32] weight chunks (swizzled); both together hold W_hh [384][64] as two halves
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g8 = lane >> 2, t4 = lane & 3;
  const int d = tid / G3, r = tid - d * G3;    // direction, gate row  (tid == row of the stacked [2][192][.] weight blocks)
  const int b = blockIdx.x;
  const int lt = r;                            // gate phase: threads r < 64 of each direction own hidden unit r
  const long long row0 = (long long)b * T;

  // prologue: x and layer 0's first chunk (group 0); its second chunk - or, for a single-chunk layer, W_hh half 1 - (group 1)
  {
    const int K0 = p.I0, nch0 = (K0 + WCH - 1) / WCH;
    stage_flat(xin, p.x + row0 * K0, T * K0);
    stage_rows(wst, p.params, K0, 0, min(WCH, K0));
    cp_async_commit();
    if (nch0 > 1) stage_rows(wst + WBUF, p.params, K0, WCH, min(WCH, K0 - WCH));
    else stage_rows(wst + WBUF, p.params + 2ll * G3 * K0 + 2 * G3, H, WCH, WCH);
    cp_async_commit();
  }

  for (int l = 0; l < p.L; ++l) {
    const int K = l == 0 ? p.I0 : 2 * H;
    const int xs = l == 0 ? p.I0 : XS;                     // row pitch of this layer's input
    const int nch = (K + WCH - 1) / WCH;
    const float* base = p.params + layer_base(l, p.I0);
    const float* bih = base + 2ll * G3 * K;                // [2][192]
    const float* whh0 = base + 2ll * G3 * K + 2 * G3;      // [2][192][64]
    const float* bhh = base + 2ll * G3 * K + 2 * G3 + 2ll * G3 * H + d * G3;
    const float* xi = xin + (l & 1) * MAXT * XS;           // this layer's input  [T][xs]
    float* xo = xin + ((l + 1) & 1) * MAXT * XS;           // next layer's input  [T][XS]
    float* gid = gi + d * T * G3;
    const float* mk = (l + 1 < p.L) ? p.mask[l] : nullptr;

    // ---- (1) input projection gi[t][row] = W_ih[row] . x[t] + b_ih[row], 32 input features per staged chunk, two chunks in flight.
    // FAST: warp w owns gate rows [32 w, 32 w + 32) (4 n-tiles) x 32 frames (2 m-tiles) of mma.sync m16n8k8 tiles; B fragments come
    // straight from the swizzled chunk, A fragments from the padded layer input.  Otherwise: thread = gate row, FFMA from registers.
    float acc[2][4][4];
    if (FAST) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) { acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f; }
    }
    for (int c = 0; c < nch; ++c) {
      const int k0 = c * WCH, kn = min(WCH, K - k0);
      const float* buf = wst + (c & 1) * WBUF;
      cp_async_wait_1();
      __syncthreads();                                     // chunk c (and, first time round, x) has landed for every thread
      if (FAST) {
        for (int ks = 0; ks < kn; ks += 8) {
          uint32_t a[2][4];
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const float* xr = xi + (m * 16 + g8) * xs + k0 + ks + t4;
            a[m][0] = tf32_rna(xr[0]); a[m][1] = tf32_rna(xr[8 * xs]); a[m][2] = tf32_rna(xr[4]); a[m][3] = tf32_rna(xr[8 * xs + 4]);
          }
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int row = warp * 32 + n * 8 + g8;
            const uint32_t b0 = tf32_rna(buf[row * WCH + 4 * (((ks >> 2)) ^ (row & 7)) + t4]);
            const uint32_t b1 = tf32_rna(buf[row * WCH + 4 * (((ks >> 2) + 1) ^ (row & 7)) + t4]);
            mma_tf32(acc[0][n], a[0], b0, b1);
            mma_tf32(acc[1][n], a[1], b0, b1);
          }
        }
      } else {
        float w[WCH];
        load_row32(buf, kn, w);
        const float bias = c == 0 ? __ldg(bih + tid) : 0.f;
        for (int t0 = 0; t0 < T; t0 += 4) {
          float a4[4] = {0.f, 0.f, 0.f, 0.f};
          // fully unrolled with a (uniform) guard on the feature count: a run-time index into w[] would push the array to local memory
#pragma unroll
          for (int j = 0; j < WCH; j += 4) {
            if (j < kn) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (t0 + i < T) {
                  const float4 x4 = *reinterpret_cast<const float4*>(xi + (t0 + i) * xs + k0 + j);      // same address in every lane: broadcast
                  a4[i] = fmaf(w[j], x4.x, a4[i]); a4[i] = fmaf(w[j + 1], x4.y, a4[i]);
                  a4[i] = fmaf(w[j + 2], x4.z, a4[i]); a4[i] = fmaf(w[j + 3], x4.w, a4[i]);
                }
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (t0 + i < T) gid[(t0 + i) * G3 + r] = (c == 0 ? bias : gid[(t0 + i) * G3 + r]) + a4[i];
        }
      }
      __syncthreads();                                     // everyone is done with this chunk buffer
      // refill it: chunk c + 2 of W_ih, or - when W_ih is exhausted - the half of W_hh that lives in this buffer (+ the masks, once)
      if (c + 2 < nch) {
        stage_rows(wst + (c & 1) * WBUF, base, K, (c + 2) * WCH, min(WCH, K - (c + 2) * WCH));
      } else {
        const int half = c & 1;                            // nch is 1 or even: the last two chunks sit in buffers 0 and 1 in that order
        if (!(nch == 1 && half == 1)) stage_rows(wst + half * WBUF, whh0, H, half * WCH, WCH);
        if (c == nch - 1 && mk) stage_flat(mks, mk + row0 * 2 * H, T * 2 * H);
      }
      cp_async_commit();
    }
    if (FAST) {
      // accumulators -> gi (+ bias): c0/c1 = (frame g8, rows 2 t4, 2 t4 + 1), c2/c3 = frame g8 + 8
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int row = warp * 32 + n * 8 + 2 * t4;        // stacked row; direction = row / 192 (a warp never straddles: 192 = 6 * 32)
        const int dd = row / G3, rr = row - dd * G3;
        const float b0 = __ldg(bih + row), b1 = __ldg(bih + row + 1);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const int f0 = m * 16 + g8, f1 = f0 + 8;
          if (f0 < T) *reinterpret_cast<float2*>(gi + dd * T * G3 + f0 * G3 + rr) = make_float2(acc[m][n][0] + b0, acc[m][n][1] + b1);
          if (f1 < T) *reinterpret_cast<float2*>(gi + dd * T * G3 + f1 * G3 + rr) = make_float2(acc[m][n][2] + b0, acc[m][n][3] + b1);
        }
      }
    }
    // ---- (2) recurrence: W_hh row in registers, masks in shared memory
    cp_async_wait_all();
    __syncthreads();                                       // both halves of W_hh and the masks have landed; gi complete
    float w[64];
    load_row32(wst, WCH, w);
    load_row32(wst + WBUF, WCH, w + WCH);
    float b_r = 0.f, b_z = 0.f, b_n = 0.f;
    if (lt < H) { b_r = __ldg(bhh + lt); b_z = __ldg(bhh + H + lt); b_n = __ldg(bhh + 2 * H + lt); hs[d * H + lt] = 0.f; }
    __syncthreads();                                       // chunk buffers free, hs zeroed
    if (l + 1 < p.L) {                                     // the next layer's first two W_ih chunks travel under this layer's recurrence
      const float* nb = p.params + layer_base(l + 1, p.I0);
      stage_rows(wst, nb, 2 * H, 0, WCH);
      cp_async_commit();
      stage_rows(wst + WBUF, nb, 2 * H, WCH, WCH);
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
        xo[t * XS + d * H + lt] = hm;
        if (mk && p.drop[l]) p.drop[l][o] = hm;
      }
      __syncthreads();
    }
  }
  // ---- heads: hsum[t][j] = fwd + rev, o1[t] = hsum[t] . w_out + b_out, prob = sigmoid(o1 . w_out2 + b_out2)
  const float* xl = xin + (p.L & 1) * MAXT * XS;           // output of the last layer (unmasked)
  for (int t = warp; t < T; t += NT / 32) {
    const float s0 = xl[t * XS + lane] + xl[t * XS + H + lane];
    const float s1 = xl[t * XS + 32 + lane] + xl[t * XS + H + 32 + lane];
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

size_t fwd_smem_bytes(int T) { return ((size_t)2 * T * G3 + (size_t)2 * MAXT * XS + 2 * H + 2 * G3 + MAXT + (size_t)T * 2 * H + 2 * WBUF) * sizeof(float); }

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
constexpr int WS2 = 2 * H + 8;                 // row pitch of a staged W_ih chunk: (8 t + g) % 32 distinct over a warp's B-fragment loads
constexpr int GS = 2 * G3 + 4;                 // row pitch of dgs: (4 g + t) % 32 distinct over a warp's A-fragment loads

// rows [r0, r0 + nrows) of a row-major [.][K] block (K % 4 == 0) -> dst with row pitch `pitch` floats
__device__ __forceinline__ void stage_rows_pitched(float* dst, const float* W, int K, int nrows, int pitch) {
  const int cpr = K >> 2;
  for (int i = threadIdx.x; i < nrows * cpr; i += NT) {
    const int row = i / cpr, c = i - row * cpr;
    cp_async16(dst + row * pitch + 4 * c, W + (long long)row * K + 4 * c);
  }
}

template <bool FAST>
__global__ void __launch_bounds__(NT, 1) dgru_stack_bwd_kernel(const StackBwdP p) {
  extern __shared__ __align__(16) float sm[];
  const int T = p.T;
  float* dcur = sm;                            // [T][128]  gradient w.r.t. the current layer's output (both directions)
  float* dgs = dcur + T * 2 * H;               // [32][GS]  dgi of the current layer, both directions (operand of the W_ih data gradient)
  float* ds = dgs + MAXT * GS;                 // [2][192]  dgh of the current step
  float* dhp = ds + 2 * G3;                    // [2][3][64] per-gate-block partials of W_hh^T dgh
  float* do1 = dhp + 2 * 3 * H;                // [32]
  float* svs = do1 + MAXT;                     // [4][T][128] saved gate planes of the current layer (r, z, n, W_hn h + b_hn)
  float* ous = svs + 4 * T * 2 * H;            // [T][128]  outputs of the current layer (h_prev of every step)
  float* wst = ous + T * 2 * H;                // [2][DXR][WS2] streamed W_ih chunks (data-gradient phase)
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g8 = lane >> 2, t4 = lane & 3;
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
  // frames T..31 of dgs are MMA padding: keep them finite
  if (FAST) for (int i = T * GS + tid; i < MAXT * GS; i += NT) dgs[i] = 0.f;
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
    if (l == 0) stage_flat(wst, wih, 2 * G3 * K);
    else stage_rows_pitched(wst, wih, K, DXR, WS2);
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
        float* gs = dgs + t * GS + d * G3 + lt;
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
    // dropout mask of the layer below.  W_ih streams through shared memory in 64-row chunks, double-buffered.
    if (l > 0) {
      const float* mk = p.mask[l - 1];
      if (FAST) {
        // [32 frames x 384] x [384 x 128] on mma.sync m16n8k8 tiles: warps 0..7 own 16 input features (2 n-tiles) x 32 frames each
        float acc[2][2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int n = 0; n < 2; ++n) { acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f; }
        for (int ch = 0; ch < nchunk; ++ch) {
          if (ch > 0) cp_async_wait_all();                 // chunk 0 landed before the recurrence; the planes may stay in flight
          __syncthreads();                                 // chunk ch visible to all; everyone is done with the other buffer
          if (ch + 1 < nchunk) {
            stage_rows_pitched(wst + ((ch + 1) & 1) * DXR * WS2, wih + (long long)(ch + 1) * DXR * K, K, DXR, WS2);
            cp_async_commit();
          }
          const float* Ws = wst + (ch & 1) * DXR * WS2;
          if (warp < 8) {
#pragma unroll 2
            for (int ks = 0; ks < DXR; ks += 8) {
              uint32_t a[2][4];
#pragma unroll
              for (int m = 0; m < 2; ++m) {
                const float* gr = dgs + (m * 16 + g8) * GS + ch * DXR + ks + t4;
                a[m][0] = tf32_rna(gr[0]); a[m][1] = tf32_rna(gr[8 * GS]); a[m][2] = tf32_rna(gr[4]); a[m][3] = tf32_rna(gr[8 * GS + 4]);
              }
#pragma unroll
              for (int n = 0; n < 2; ++n) {
                const int f = warp * 16 + n * 8 + g8;
                const uint32_t b0 = tf32_rna(Ws[(ks + t4) * WS2 + f]), b1 = tf32_rna(Ws[(ks + t4 + 4) * WS2 + f]);
                mma_tf32(acc[0][n], a[0], b0, b1);
                mma_tf32(acc[1][n], a[1], b0, b1);
              }
            }
          }
        }
        // dcur is only read by the recurrence: free to overwrite
        if (warp < 8) {
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            const int f = warp * 16 + n * 8 + 2 * t4;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              const int f0 = m * 16 + g8, f1 = f0 + 8;
              if (f0 < T) {
                float2 v = make_float2(acc[m][n][0], acc[m][n][1]);
                if (mk) { const float2 q = __ldg(reinterpret_cast<const float2*>(mk + (row0 + f0) * 2 * H + f)); v.x *= q.x; v.y *= q.y; }
                *reinterpret_cast<float2*>(dcur + f0 * 2 * H + f) = v;
              }
              if (f1 < T) {
                float2 v = make_float2(acc[m][n][2], acc[m][n][3]);
                if (mk) { const float2 q = __ldg(reinterpret_cast<const float2*>(mk + (row0 + f1) * 2 * H + f)); v.x *= q.x; v.y *= q.y; }
                *reinterpret_cast<float2*>(dcur + f1 * 2 * H + f) = v;
              }
            }
          }
        }
      } else {
        // thread = (input feature k, one third of the frames)
        const int k = tid & 127, tg = tid >> 7;            // 128 features x 3 frame groups
        const int t_lo = tg * ((T + 2) / 3), t_hi = min(T, t_lo + (T + 2) / 3);
        constexpr int TG = 11;                             // frames per thread (T <= 32 -> ceil(32 / 3))
        float acc[TG];
#pragma unroll
        for (int i = 0; i < TG; ++i) acc[i] = 0.f;
        for (int ch = 0; ch < nchunk; ++ch) {
          if (ch > 0) cp_async_wait_all();
          __syncthreads();
          if (ch + 1 < nchunk) {
            stage_rows_pitched(wst + ((ch + 1) & 1) * DXR * WS2, wih + (long long)(ch + 1) * DXR * K, K, DXR, WS2);
            cp_async_commit();
          }
          const float* Ws = wst + (ch & 1) * DXR * WS2;
          for (int rr = 0; rr < DXR; rr += 4) {
            const float w0 = Ws[rr * WS2 + k], w1 = Ws[(rr + 1) * WS2 + k], w2 = Ws[(rr + 2) * WS2 + k], w3 = Ws[(rr + 3) * WS2 + k];
#pragma unroll
            for (int i = 0; i < TG; ++i) {
              if (t_lo + i < t_hi) {
                const float4 g4 = *reinterpret_cast<const float4*>(dgs + (t_lo + i) * GS + ch * DXR + rr);      // broadcast
                acc[i] = fmaf(g4.x, w0, acc[i]); acc[i] = fmaf(g4.y, w1, acc[i]); acc[i] = fmaf(g4.z, w2, acc[i]); acc[i] = fmaf(g4.w, w3, acc[i]);
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < TG; ++i) {
          const int t = t_lo + i;
          if (t < t_hi) dcur[t * 2 * H + k] = mk ? acc[i] * __ldg(mk + (row0 + t) * 2 * H + k) : acc[i];
        }
      }
    } else {
      // layer 0: [T][384] x [384][I0], thread = (frame, input feature); W_ih (tight [384][I0]) was staged whole before the recurrence
      cp_async_wait_all();
      __syncthreads();
      if (p.dx0) {
        for (int idx = tid; idx < T * K; idx += NT) {
          const int t = idx / K, k = idx - t * K;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          for (int rr = 0; rr < 2 * G3; rr += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(dgs + t * GS + rr);
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
  return ((size_t)T * 2 * H + (size_t)MAXT * GS + 2 * G3 + 2 * 3 * H + MAXT + (size_t)5 * T * 2 * H + (size_t)2 * DXR * WS2) * sizeof(float);
}

}  // namespace

extern "C" int tg_dgru_stack_bwd(const float* dlogit, const float* gru_params, const float* const* masks, const float* const* outs,
                                 const float* const* saved, long long saved_qstride, const float* hsum, const float* o1, const float* w_out,
                                 const float* w_out2, float* const* dgi, float* const* dgh, float* dx0, float* g_w_out, float* g_b_out,
                                 float* g_w_out2, float* g_b_out2, int B, int T, int I0, int Hh, int L, int fast, tg_stream stream) {
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
  auto kern = fast ? dgru_stack_bwd_kernel<true> : dgru_stack_bwd_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("tg_dgru_stack_bwd: smem attr: %s", cudaGetErrorString(e)); return -3; }
  kern<<<B, NT, smem, (cudaStream_t)stream>>>(p);
  TG_CHECK_LAUNCH("tg_dgru_stack_bwd");
  return 0;
}
