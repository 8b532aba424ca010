// Persistent bidirectional-GRU recurrence on the 5th-generation tensor cores ("fast" mode; nn.GRU at
// multimodal_context_net.py:98-99,155,221-222,241).  Same CTA decomposition and L2 release/acquire stepping as the
// fp32 kernels in gru.cu, but the per-step product runs as tcgen05.mma kind::tf32 with the accumulator in TMEM:
//
//  forward : D[gate row (128 lanes), batch (BT cols)] = W_hh slice [3u x H] (A operand, loaded ONCE by TMA and resident
//            in shared memory for all T steps) x h_{t-1} tile [BT x H] (B operand, TMA-loaded from the layer output each
//            step, K-chunk by K-chunk so the MMAs start while later chunks are still in flight).  Epilogue warps pull
//            the accumulator out of TMEM, regroup r/z/n per hidden unit through shared memory, apply the gate
//            non-linearities and the state update in registers and publish h_t.
//  backward: D[batch (128 lanes), H cols] = dgh tile [128 x 3u] (A operand, written by the epilogue threads straight
//            into the 128B-swizzled UMMA layout) x W_hh^T slice [H x 3u] (B operand, resident); the per-chunk partial
//            dh_{t-1} are exchanged through L2 and summed by the owner of each hidden unit.
#include <cudaTypedefs.h>
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until(const int* counter, int target) {
  while (ld_acquire_gpu(counter) < target) { __nanosleep(20); }
}
__device__ __forceinline__ void epi_bar512() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// fast-mode gate non-linearities (absolute error ~1e-6, far below the TF32 rounding of the products they follow)
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }
// optional per-step phase stamps of CTA (0,0,0) (tg_debug_gru_trace): trace[step*16 + slot] = %globaltimer
__device__ __forceinline__ void stamp(long long* trace, int step, int slot) {
  if (trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)::"memory");
    trace[step * 16 + slot] = (long long)t;
  }
}

// Loads as volatile asm: the compiler may neither sink them to their first use nor move them across the (volatile) barrier
// waits, so a batch of loads issued before a wait really is in flight during the wait.
__device__ __forceinline__ float ldv_nc(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldv_cg(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldv_nc4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldv_cg4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// =====================================================================================================================
// forward
// =====================================================================================================================
struct FwdP {
  const float* gi; const float* bhh[2]; float* out; float* saved; long long saved_qstride; int* sync;
  int B, T, H, u, UC, NB, ntiles, nkc;
  long long* trace;
};

template <int BT>
__global__ void __launch_bounds__(576, 1) gru_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                                                            const __grid_constant__ CUtensorMap tmH0, const __grid_constant__ CUtensorMap tmH1,
                                                            const FwdP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int W_CHUNK = 128 * 128;          // 128 rows x 128 B
  constexpr int H_CHUNK = BT * 128;
  constexpr int GS = BT + 1;
  constexpr int NP = (BT * 40 + 511) / 512;     // (clip, unit) pairs per epilogue thread (u <= 40, 512 epilogue threads)
  const int nkc = p.nkc;
  uint8_t* Wt = smem;                          // [nkc][128 rows][128 B]
  uint8_t* Ht = smem + (size_t)nkc * W_CHUNK;  // [nkc][BT rows][128 B]; aliased by ghs[128][GS] once the MMAs are done
  float* ghs = reinterpret_cast<float*>(Ht);
  size_t ht_bytes = (size_t)nkc * H_CHUNK;
  if (ht_bytes < (size_t)128 * GS * 4) ht_bytes = (size_t)128 * GS * 4;
  ht_bytes = (ht_bytes + 15) & ~(size_t)15;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(Ht + ht_bytes);
  uint64_t* h_full = w_full + 1;               // [nkc] (<= 12)
  uint64_t* tmem_full = h_full + 12;
  uint64_t* epi_done = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x, by = blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, u = p.u;
  const int u0 = c * u;
  const CUtensorMap* tmW = dir ? &tmW1 : &tmW0;
  const CUtensorMap* tmH = dir ? &tmH1 : &tmH0;
  int* counter = p.sync + dir * p.NB + by;
  constexpr uint32_t TMEM_COLS = BT <= 32 ? 32 : 64;
  int my_tiles = 0;
  for (int tile = by; tile < p.ntiles; tile += p.NB) ++my_tiles;

  // zero the whole operand region once (A rows >= 3u are never written by TMA but are read by the MMA)
  for (int i = threadIdx.x; i < (nkc * W_CHUNK) / 16; i += blockDim.x) reinterpret_cast<float4*>(Wt)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmW); tma_prefetch_desc(tmH);
    mbar_init(w_full, 1);
    for (int k = 0; k < 12; ++k) mbar_init(&h_full[k], 1);
    mbar_init(tmem_full, 1); mbar_init(epi_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // resident recurrent weights: rows [g*H + u0, +u) of W_hh for g = r, z, n  ->  A rows [g*u, +u)
      mbar_expect_tx(w_full, (uint32_t)(nkc * 3 * u * 128));
      for (int kc = 0; kc < nkc; ++kc)
        for (int g = 0; g < 3; ++g) tma_load_2d(Wt + (size_t)kc * W_CHUNK + (size_t)g * u * 128, tmW, w_full, kc * 32, g * H + u0);
    }
    // per step: lane 0 waits for h_{t-1} of every CTA of the group, then the nkc K-chunks of the h tile are issued by nkc lanes in
    // parallel (one TMA + one mbarrier each): serial issue from one thread cost ~1.2 us of the ~9 us step chain
    int it = 0;
    for (int s = 1; s < T; ++s) {
      const int t = dir == 0 ? s : T - 1 - s;
      const int tp = dir == 0 ? t - 1 : t + 1;
      if (lane == 0) {
        stamp(p.trace, s, 0);
        spin_until(counter, p.UC * s);
        stamp(p.trace, s, 1);
      }
      __syncwarp();
      for (int tile = by; tile < p.ntiles; tile += p.NB, ++it) {
        if (lane == 0 && it > 0) mbar_wait(epi_done, (uint32_t)((it - 1) & 1));
        __syncwarp();
        if (lane < nkc) {
          // h_{t-1} was written through the generic proxy by other CTAs: each writer ran fence.proxy.async after its stores and before
          // its release (below), lane 0 acquired the counter and __syncwarp ordered this lane behind it - no reader-side proxy fence
          // (it cost ~1 us per step: TGB200_GRU_READER_FENCE=1 at build time restores it)
#ifdef TGB200_GRU_READER_FENCE
          fence_proxy_async_all();
#endif
          mbar_expect_tx(&h_full[lane], (uint32_t)H_CHUNK);
          tma_load_3d(Ht + (size_t)lane * H_CHUNK, tmH, &h_full[lane], lane * 32, tp, tile * BT);
        }
        __syncwarp();
        if (lane == 0) stamp(p.trace, s, 2);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32(128, BT, 0, 0);
      mbar_wait(w_full, 0);
      int it = 0;
      for (int s = 1; s < T; ++s) {
        for (int tl = 0; tl < my_tiles; ++tl, ++it) {
          for (int kc = 0; kc < nkc; ++kc) {
            mbar_wait(&h_full[kc], (uint32_t)(it & 1));
            tc_fence_after();
            const uint32_t sa = smem_u32(Wt + (size_t)kc * W_CHUNK);
            const uint32_t sb = smem_u32(Ht + (size_t)kc * H_CHUNK);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              mma_tf32(tmem_base, smem_desc_sw128(sa + k4 * 32, 16, 1024), smem_desc_sw128(sb + k4 * 32, 16, 1024), idesc,
                       (kc > 0 || k4 > 0) ? 1u : 0u);
          }
          tc_commit(tmem_full);
        }
      }
    }
  } else {
    // ---- epilogue: 16 warps.  Warps 2-5 move the accumulator TMEM -> shared memory (thread = TMEM lane = gate row
    // lr = g*u + jj); all 512 threads then evaluate the gates (the transcendental-heavy part) for (clip, unit) pairs.
    const int etid = threadIdx.x - 64;
    const int q = warp & 3;
    const int lr = q * 32 + lane;
    const float* bhh = p.bhh[dir];
    const long long row2H = 2ll * H;
    // (clip, unit) pair owned by this thread in slot e: loop invariant (the integer divisions were a quarter of all executed
    // instructions when they sat inside the step loop - ncu source view, profiles/README.md)
    int ppk[NP];                                   // (clip << 16) | unit, or -1
#pragma unroll
    for (int e = 0; e < NP; ++e) {
      const int i = etid + 512 * e;
      const int bb = i / u;
      ppk[e] = (i < BT * u) ? ((bb << 16) | (i - bb * u)) : -1;
    }
    int it = 0;
    for (int s = 0; s < T; ++s) {
      const int t = dir == 0 ? s : T - 1 - s;
      const int tp = dir == 0 ? t - 1 : t + 1;
      for (int tile = by; tile < p.ntiles; tile += p.NB) {
        const int b0 = tile * BT;
        // pass 1: every global load of this (step, tile) in one batch
        float gir[NP], giz[NP], gin[NP], hpv[NP], bhh_r_[NP], bhh_z_[NP], bhh_n_[NP];
#pragma unroll
        for (int e = 0; e < NP; ++e) {
          gir[e] = giz[e] = gin[e] = hpv[e] = bhh_r_[e] = bhh_z_[e] = bhh_n_[e] = 0.f;
          if (ppk[e] >= 0) {
            const int bb = ppk[e] >> 16, jj = ppk[e] & 0xffff;
            const int b = b0 + bb, unit = u0 + jj;
            if (b < p.B && unit < H) {
              const float* gip = p.gi + ((long long)b * T + t) * 6 * H + dir * 3 * H + unit;
              gir[e] = ldv_nc(gip); giz[e] = ldv_nc(gip + H); gin[e] = ldv_nc(gip + 2 * H);
              bhh_r_[e] = ldv_nc(bhh + unit); bhh_z_[e] = ldv_nc(bhh + H + unit); bhh_n_[e] = ldv_nc(bhh + 2 * H + unit);
              if (s > 0) hpv[e] = ldv_cg(p.out + ((long long)b * T + tp) * row2H + dir * H + unit);
            }
          }
        }
        if (s > 0) {
          if (warp < 6) {
            mbar_wait(tmem_full, (uint32_t)(it & 1));
            if (etid == 0) stamp(p.trace, s, 3);
            tc_fence_after();
            float v[BT];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
            if constexpr (BT == 16) { tmem_ld16(taddr, v); }
            else if constexpr (BT == 32) { tmem_ld32(taddr, v); }
            else { tmem_ld32(taddr, v); tmem_ld16(taddr + 32, v + 32); }
            tmem_ld_wait();
#pragma unroll
            for (int bb = 0; bb < BT; ++bb) ghs[lr * GS + bb] = v[bb];
            tc_fence_before();
          }
          epi_bar512();
          if (etid == 0) stamp(p.trace, s, 4);
        }
        // pass 2: gates.  The global loads (gi, h_{t-1}) were issued in pass 1 (before the TMEM wait), so nothing
        // here depends on a fresh memory round trip.
#pragma unroll
        for (int e = 0; e < NP; ++e) {
          if (ppk[e] < 0) continue;
          const int bb = ppk[e] >> 16, jj = ppk[e] & 0xffff;
          const int b = b0 + bb, unit = u0 + jj;
          if (b >= p.B || unit >= H) continue;
          float ghr = bhh_r_[e], ghz = bhh_z_[e], ghn = bhh_n_[e];
          if (s > 0) { ghr += ghs[jj * GS + bb]; ghz += ghs[(u + jj) * GS + bb]; ghn += ghs[(2 * u + jj) * GS + bb]; }
          const float r = sigmoidf_(gir[e] + ghr);
          const float z = sigmoidf_(giz[e] + ghz);
          const float n = tanhf_(gin[e] + r * ghn);
          const float h = (1.f - z) * n + z * hpv[e];
          const long long o = ((long long)b * T + t) * row2H + dir * H + unit;
          p.out[o] = h;
          // r, z, n, hn are only needed by the backward pass: kept in registers, stored after h_t has been published
          gir[e] = r; giz[e] = z; gin[e] = n; hpv[e] = ghn;
        }
        if (s > 0) {
          fence_proxy_async_smem();      // ghs (generic proxy) is about to be overwritten by the next TMA (async proxy)
          epi_bar512();
          if (etid == 0) mbar_arrive(epi_done);
          ++it;
        }
        const bool last_tile = tile + p.NB >= p.ntiles;
        if (last_tile && s + 1 < T) {
          // publish h_t before the (off-critical-path) stores of the saved gates of this tile
          if (etid == 0) stamp(p.trace, s, 5);
          __threadfence();
          fence_proxy_async_all();         // h_t is read by other CTAs through TMA (async proxy)
          if (etid == 0) stamp(p.trace, s, 6);
          epi_bar512();
          if (etid == 0) { atomicAdd(counter, 1); stamp(p.trace, s, 7); }
        }
        if (p.saved) {
#pragma unroll
          for (int e = 0; e < NP; ++e) {
            if (ppk[e] < 0) continue;
            const int bb = ppk[e] >> 16, jj = ppk[e] & 0xffff;
            const int b = b0 + bb, unit = u0 + jj;
            if (b >= p.B || unit >= H) continue;
            const long long o = ((long long)b * T + t) * row2H + dir * H + unit;
            p.saved[o] = gir[e]; p.saved[p.saved_qstride + o] = giz[e]; p.saved[2 * p.saved_qstride + o] = gin[e];
            p.saved[3 * p.saved_qstride + o] = hpv[e];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// =====================================================================================================================
// backward
// =====================================================================================================================
struct BwdP {
  const float* dout; const float* out; const float* saved; long long saved_qstride;
  float* dgi; float* dgh; float* partial; int* sync;
  int B, T, H, UC, NB, HP, nh, Nh;      // nh N-halves of Nh columns each (nh*Nh >= H)
  long long* trace;
};
constexpr int BU = 32;                   // hidden units per CTA in the backward kernel (one 128-byte K chunk per gate)
constexpr int MAXUC = 12;                // H <= 384

__device__ __forceinline__ void epi_bar256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 320 threads: warp 0 = TMA (weights, once), warp 1 = TMEM + MMA issuer, warps 2-9 = 256 epilogue threads.
// RB = clips owned by one CTA (rows 0..RB-1 of the 128-row MMA; the rest stay zero).  Small RB spreads the L2 exchange of
// the partial dh and the operand streams over more SMs (the MMA itself is far from being the bottleneck): with B = 128,
// RB = 32 gives 80 CTAs and each epilogue thread handles one float4 of units - one round of loads per step.
template <int RB>
__global__ void __launch_bounds__(320, 1) gru_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmT0, const __grid_constant__ CUtensorMap tmT1,
                                                            const BwdP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nrows = p.nh * p.Nh;
  const size_t B_CHUNK = (size_t)nrows * 128;
  constexpr int A_CHUNK = 128 * 128;
  uint8_t* Bt = smem;                                   // [3 gates][nrows][128 B]   W_hh^T slice, resident
  uint8_t* At = smem + 3 * B_CHUNK;                     // [3 gates][128 rows][128 B] dgh tile (swizzled)
  uint64_t* w_full = reinterpret_cast<uint64_t*>(At + 3 * A_CHUNK);
  uint64_t* a_ready = w_full + 1;
  uint64_t* tmem_full = a_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x, by = blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, HP = p.HP;
  const int u0 = c * BU;
  const CUtensorMap* tmT = dir ? &tmT1 : &tmT0;
  int* counter = p.sync + dir * p.NB + by;
  const uint32_t tmem_cols = nrows <= 32 ? 32 : nrows <= 64 ? 64 : nrows <= 128 ? 128 : nrows <= 256 ? 256 : 512;

  for (int i = threadIdx.x; i < (3 * A_CHUNK) / 16; i += blockDim.x) reinterpret_cast<float4*>(At)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmT);
    mbar_init(w_full, 1); mbar_init(a_ready, 1); mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // B[n = k_out, kk = own gate row] = W_hh[g*H + u0 + kk, n] = whhT[n][g*H + u0 + kk]
      mbar_expect_tx(w_full, (uint32_t)(3 * B_CHUNK));
      for (int g = 0; g < 3; ++g)
        for (int hf = 0; hf < p.nh; ++hf)
          tma_load_2d(Bt + g * B_CHUNK + (size_t)hf * p.Nh * 128, tmT, w_full, g * H + u0, hf * p.Nh);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, p.Nh, 0, 0);
      mbar_wait(w_full, 0);
      for (int s = 0; s + 1 < T; ++s) {
        mbar_wait(a_ready, (uint32_t)(s & 1));
        tc_fence_after();
        for (int hf = 0; hf < p.nh; ++hf) {
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const uint32_t sa = smem_u32(At + g * A_CHUNK);
            const uint32_t sb = smem_u32(Bt + g * B_CHUNK + (size_t)hf * p.Nh * 128);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              mma_tf32(tmem_base + (uint32_t)(hf * p.Nh), smem_desc_sw128(sa + k4 * 32, 16, 1024), smem_desc_sw128(sb + k4 * 32, 16, 1024), idesc,
                       (g > 0 || k4 > 0) ? 1u : 0u);
          }
        }
        tc_commit(tmem_full);
      }
    }
  } else {
    const int etid = threadIdx.x - 64;                  // 0..255
    constexpr int TPR = 256 / RB;                       // threads per clip row
    constexpr int HU = BU / TPR;                        // hidden units per thread (4, 8 or 16)
    constexpr int NG = HU / 4;                          // float4 groups per thread
    const int rl = etid / TPR, sub = etid - rl * TPR;   // clip row of the tile, unit sub-range
    const int b = by * RB + rl;
    const bool b_ok = b < p.B;
    const long long row2H = 2ll * H;
    const long long pstride_parity = 2ll * p.B * p.UC * HP;
    const long long k4n = HP / 4;
    const int uu0 = u0 + sub * HU;
    float dhc[HU];
#pragma unroll
    for (int j = 0; j < HU; ++j) dhc[j] = 0.f;
    int nu = H - uu0;                                   // valid units of this thread (multiple of 4)
    nu = nu < 0 ? 0 : (nu > HU ? HU : nu);
    // accumulator read-out: TMEM lanes 32q..32q+31 are only visible to warps with (warp & 3) == q
    const int q = warp & 3, half = (warp - 2) >> 2;
    const bool t_warp = q * 32 < RB;
    const int tb = by * RB + q * 32 + lane;             // clip whose accumulator row this thread reads
    const bool tb_ok = t_warp && (q * 32 + lane) < RB && tb < p.B;
    const int nchunk = (nrows + 31) / 32;
    const int ch_lo = half == 0 ? 0 : (nchunk + 1) / 2, ch_hi = half == 0 ? (nchunk + 1) / 2 : nchunk;
    for (int s = 0; s < T; ++s) {
      const int t = dir == 0 ? T - 1 - s : s;
      const int tp = dir == 0 ? t - 1 : t + 1;
      const bool tp_ok = tp >= 0 && tp < T;
      const long long row = (long long)b * T + t;
      const long long o = row * row2H + dir * H + uu0;
      const long long op = ((long long)b * T + tp) * row2H + dir * H + uu0;
      // recurrence-independent operands of this step: one batch of loads issued BEFORE waiting for the other CTAs
      float4 ld_do[NG], ld_r[NG], ld_z[NG], ld_n[NG], ld_hn[NG], ld_hp[NG];
#pragma unroll
      for (int g4 = 0; g4 < NG; ++g4) {
        ld_do[g4] = ld_r[g4] = ld_z[g4] = ld_n[g4] = ld_hn[g4] = ld_hp[g4] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b_ok && g4 * 4 < nu) {
          ld_do[g4] = ldv_nc4(p.dout + o + g4 * 4);
          ld_r[g4] = ldv_nc4(p.saved + o + g4 * 4);
          ld_z[g4] = ldv_nc4(p.saved + p.saved_qstride + o + g4 * 4);
          ld_n[g4] = ldv_nc4(p.saved + 2 * p.saved_qstride + o + g4 * 4);
          ld_hn[g4] = ldv_nc4(p.saved + 3 * p.saved_qstride + o + g4 * 4);
          if (tp_ok) ld_hp[g4] = ldv_nc4(p.out + op + g4 * 4);
        }
      }
      if (etid == 0) stamp(p.trace, s, 0);
      if (s > 0) {
        if (etid == 0) spin_until(counter, p.UC * s);
        epi_bar256();
      }
      if (etid == 0) stamp(p.trace, s, 1);
      float carry[HU];
#pragma unroll
      for (int j = 0; j < HU; ++j) carry[j] = dhc[j];
      if (s > 0 && b_ok) {
        // partial layout [parity][dir][chunk cc][k4 = unit/4][clip b][4]: consecutive lanes touch consecutive clips
        const float* Pin = p.partial + (long long)((s - 1) & 1) * pstride_parity + (long long)dir * p.UC * k4n * p.B * 4 +
                           ((long long)(uu0 / 4) * p.B + b) * 4;
#pragma unroll
        for (int g4 = 0; g4 < NG; ++g4) {
          if (g4 * 4 < nu) {
            float4 tv[MAXUC];
#pragma unroll
            for (int cc = 0; cc < MAXUC; ++cc) {
              tv[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (cc < p.UC) tv[cc] = ldv_cg4(Pin + ((long long)cc * k4n + g4) * p.B * 4);
            }
#pragma unroll
            for (int cc = 0; cc < MAXUC; ++cc) {
              carry[g4 * 4 + 0] += tv[cc].x; carry[g4 * 4 + 1] += tv[cc].y; carry[g4 * 4 + 2] += tv[cc].z; carry[g4 * 4 + 3] += tv[cc].w;
            }
          }
        }
      }
      if (etid == 0) stamp(carry[0] == 12345.678f ? nullptr : p.trace, s, 2);
      float* gp = p.dgi + row * 6 * H + dir * 3 * H + uu0;
      float* hp = p.dgh + row * 6 * H + dir * 3 * H + uu0;
#pragma unroll
      for (int g4 = 0; g4 < NG; ++g4) {
        float dr4[4] = {0.f, 0.f, 0.f, 0.f}, dz4[4] = {0.f, 0.f, 0.f, 0.f}, dn4[4] = {0.f, 0.f, 0.f, 0.f}, dnr4[4] = {0.f, 0.f, 0.f, 0.f};
        if (b_ok && g4 * 4 < nu) {
          const float dov[4] = {ld_do[g4].x, ld_do[g4].y, ld_do[g4].z, ld_do[g4].w}, rv[4] = {ld_r[g4].x, ld_r[g4].y, ld_r[g4].z, ld_r[g4].w};
          const float zv[4] = {ld_z[g4].x, ld_z[g4].y, ld_z[g4].z, ld_z[g4].w}, nv[4] = {ld_n[g4].x, ld_n[g4].y, ld_n[g4].z, ld_n[g4].w};
          const float hnv[4] = {ld_hn[g4].x, ld_hn[g4].y, ld_hn[g4].z, ld_hn[g4].w}, hpv[4] = {ld_hp[g4].x, ld_hp[g4].y, ld_hp[g4].z, ld_hp[g4].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float dh = dov[e] + carry[g4 * 4 + e];
            const float dn = dh * (1.f - zv[e]) * (1.f - nv[e] * nv[e]);
            dz4[e] = dh * (hpv[e] - nv[e]) * zv[e] * (1.f - zv[e]);
            dr4[e] = dn * hnv[e] * rv[e] * (1.f - rv[e]);
            dn4[e] = dn;
            dnr4[e] = dn * rv[e];
            dhc[g4 * 4 + e] = dh * zv[e];
          }
          *reinterpret_cast<float4*>(gp + g4 * 4) = make_float4(dr4[0], dr4[1], dr4[2], dr4[3]);
          *reinterpret_cast<float4*>(gp + H + g4 * 4) = make_float4(dz4[0], dz4[1], dz4[2], dz4[3]);
          *reinterpret_cast<float4*>(gp + 2 * H + g4 * 4) = make_float4(dn4[0], dn4[1], dn4[2], dn4[3]);
          *reinterpret_cast<float4*>(hp + g4 * 4) = make_float4(dr4[0], dr4[1], dr4[2], dr4[3]);
          *reinterpret_cast<float4*>(hp + H + g4 * 4) = make_float4(dz4[0], dz4[1], dz4[2], dz4[3]);
          *reinterpret_cast<float4*>(hp + 2 * H + g4 * 4) = make_float4(dnr4[0], dnr4[1], dnr4[2], dnr4[3]);
        }
        // A operand: K-major, 128B swizzle: 16-byte unit index XOR (row & 7)
        const uint32_t unit16 = (uint32_t)(sub * NG + g4) ^ (uint32_t)(rl & 7);
        const size_t off = (size_t)(rl >> 3) * 1024 + (size_t)(rl & 7) * 128 + (size_t)unit16 * 16;
        *reinterpret_cast<float4*>(At + 0 * A_CHUNK + off) = make_float4(dr4[0], dr4[1], dr4[2], dr4[3]);
        *reinterpret_cast<float4*>(At + 1 * A_CHUNK + off) = make_float4(dz4[0], dz4[1], dz4[2], dz4[3]);
        *reinterpret_cast<float4*>(At + 2 * A_CHUNK + off) = make_float4(dnr4[0], dnr4[1], dnr4[2], dnr4[3]);
      }
      if (s + 1 < T) {
        if (etid == 0) stamp(p.trace, s, 3);
        fence_proxy_async_smem();
        tc_fence_before();
        epi_bar256();
        if (etid == 0) mbar_arrive(a_ready);
        if (etid == 0) stamp(p.trace, s, 4);
        if (t_warp) {
          // partial dh_{prev}[clip, 0..H) of this chunk -> L2 (the two warps of a lane quarter split the columns)
          mbar_wait(tmem_full, (uint32_t)(s & 1));
          if (etid == 0) stamp(p.trace, s, 5);
          tc_fence_after();
          float* Pout = p.partial + (long long)(s & 1) * pstride_parity + ((long long)dir * p.UC + c) * k4n * p.B * 4 + (long long)tb * 4;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
          for (int ch = ch_lo; ch < ch_hi; ++ch) {
            float v[32];
            tmem_ld32(taddr + (uint32_t)(ch * 32), v);
            tmem_ld_wait();
            if (tb_ok) {
#pragma unroll
              for (int j4 = 0; j4 < 32; j4 += 4)
                if (ch * 32 + j4 < HP)
                  __stcg(reinterpret_cast<float4*>(Pout + (long long)(ch * 8 + (j4 >> 2)) * p.B * 4), make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]));
            }
            __syncwarp();
          }
          tc_fence_before();
        }
        if (etid == 0) stamp(p.trace, s, 6);
        __threadfence();
        if (etid == 0) stamp(p.trace, s, 7);
        epi_bar256();
        if (etid == 0) { atomicAdd(counter, 1); stamp(p.trace, s, 8); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

int map_2d(CUtensorMap* m, const float* base, long long rows, long long cols, long long ld, int box_rows, const char* name) {
  auto enc = get_encode();
  if (!enc) { tg_set_error("%s: cuTensorMapEncodeTiled unavailable", name); return -4; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 4) & 15)) { tg_set_error("%s: TMA alignment (ld=%lld)", name, ld); return -1; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { tg_set_error("%s: cuTensorMapEncodeTiled failed (%d)", name, (int)r); return -4; }
  return 0;
}
// out [B, T, 2H] seen as (k < H) x T x B for one direction: box = 32 floats x 1 step x box_b clips
int map_h3d(CUtensorMap* m, const float* base, int B, int T, int H, int box_b, const char* name) {
  auto enc = get_encode();
  if (!enc) { tg_set_error("%s: cuTensorMapEncodeTiled unavailable", name); return -4; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((2ll * H * 4) & 15)) { tg_set_error("%s: TMA alignment (H=%d)", name, H); return -1; }
  cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)2 * H * 4, (cuuint64_t)T * 2 * H * 4};
  cuuint32_t box[3] = {32, 1, (cuuint32_t)box_b};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { tg_set_error("%s: cuTensorMapEncodeTiled(3d) failed (%d)", name, (int)r); return -4; }
  return 0;
}

long long* g_trace = nullptr;

struct FwdPlan { int u, UC, BT, ntiles, NB, nkc; };
int fwd_plan(int B, int H, FwdPlan* pl) {
  if (H < 32 || H > 384 || (H & 3)) return -1;
  pl->u = 40;
  if (H <= 64) pl->u = 32;
  pl->UC = tg_ceil_div(H, pl->u);
  pl->nkc = tg_ceil_div(H, 32);
  int max_nb = tg_num_sms() / (2 * pl->UC);
  if (max_nb < 1) max_nb = 1;
  int bt = 16;
  while (bt < 48 && tg_ceil_div(B, bt) > max_nb) bt += 16;
  pl->BT = bt;
  pl->ntiles = tg_ceil_div(B, bt);
  pl->NB = pl->ntiles < max_nb ? pl->ntiles : max_nb;
  return 0;
}

template <int BT>
int launch_fwd(const CUtensorMap* maps, const FwdP& p, const FwdPlan& pl, cudaStream_t s) {
  size_t ht = (size_t)pl.nkc * BT * 128;
  if (ht < (size_t)128 * (BT + 1) * 4) ht = (size_t)128 * (BT + 1) * 4;
  ht = (ht + 15) & ~(size_t)15;
  const size_t smem = (size_t)pl.nkc * 128 * 128 + ht + 16 * 8 + 16 + 1024;
  if (smem > (size_t)tg_max_smem_optin()) { tg_set_error("tg_gru_layer_fwd_tf32: %zu B of shared memory needed", smem); return -3; }
  cudaError_t e = cudaFuncSetAttribute(gru_fwd_tc_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_fwd_tf32: smem attr: %s", cudaGetErrorString(e)); return -3; }
  void* args[] = {(void*)&maps[0], (void*)&maps[1], (void*)&maps[2], (void*)&maps[3], (void*)&p};
  dim3 grid(pl.UC, pl.NB, 2);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &cap);       // see gru.cu: cooperative launches cannot be captured into a CUDA graph
  if (cap != cudaStreamCaptureStatusNone) e = cudaLaunchKernel((const void*)gru_fwd_tc_kernel<BT>, grid, dim3(576), args, smem, s);
  else e = cudaLaunchCooperativeKernel((const void*)gru_fwd_tc_kernel<BT>, grid, dim3(576), args, smem, s);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_fwd_tf32: cooperative launch (%u,%u,2) smem %zu: %s", grid.x, grid.y, smem, cudaGetErrorString(e)); return -2; }
  return 0;
}

bool use_legacy_gru() {          // TGB200_GRU_LEGACY=1: the L2-counter-stepped kernels of this file even where a cluster plan exists
  static int v = -1;
  if (v < 0) { const char* e = getenv("TGB200_GRU_LEGACY"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

}  // namespace

// cluster / multicast kernels (gru_cl.cu): the default wherever they have a plan (H <= 320)
struct TgGruClPlan { int u, bt_f, bt_b, ntiles_f, ntiles_b; };
bool tg_gru_cl_plan(int B, int H, TgGruClPlan* pl);
size_t tg_gru_cl_xchg_floats(int B, int H);
int tg_gru_cl_fwd(const float* gi, const float* whh_f, const float* whh_r, const float* bhh_f, const float* bhh_r, float* out, float* saved,
                  long long saved_qstride, const float* mask, float* drop, float* xchg, int B, int T, int H, long long* trace, cudaStream_t s);
int tg_gru_cl_bwd(const float* dout, const float* out, const float* saved, long long saved_qstride, const float* whhT_f, const float* whhT_r,
                  float* dgi, float* dgh, float* xchg, int B, int T, int H, long long* trace, cudaStream_t s);

extern "C" int tg_debug_gru_trace(long long* device_buf) {
  g_trace = device_buf;
  return 0;
}

extern "C" int tg_gru_tf32_sync_ints(int B, int H) {
  FwdPlan pl;
  if (fwd_plan(B, H, &pl)) return -1;
  int nb_bwd = tg_ceil_div(B, 32);
  const size_t legacy = (size_t)2 * (pl.NB > nb_bwd ? pl.NB : nb_bwd);
  const size_t xchg = tg_gru_cl_xchg_floats(B, H);      // the cluster kernels use the same scratch as their exchange images
  return (int)(legacy > xchg ? legacy : xchg);
}

extern "C" int tg_gru_layer_fwd_tf32_drop(const float* gi, const float* whh_f, const float* whh_r, const float* bhh_f, const float* bhh_r,
                                          float* out, float* saved, long long saved_qstride, const float* mask, float* drop, int* sync, int B,
                                          int T, int H, tg_stream stream) {
  TG_REQUIRE(mask && drop, "tg_gru_layer_fwd_tf32_drop");
  TgGruClPlan cpl;
  if (!use_legacy_gru() && tg_gru_cl_plan(B, H, &cpl)) {
    TG_REQUIRE(gi && whh_f && whh_r && bhh_f && bhh_r && out && sync && T > 0 && B > 0, "tg_gru_layer_fwd_tf32_drop");
    return tg_gru_cl_fwd(gi, whh_f, whh_r, bhh_f, bhh_r, out, saved, saved_qstride, mask, drop, reinterpret_cast<float*>(sync), B, T, H, g_trace,
                         (cudaStream_t)stream);
  }
  const int rc = tg_gru_layer_fwd_tf32(gi, whh_f, whh_r, bhh_f, bhh_r, out, saved, saved_qstride, sync, B, T, H, stream);
  return rc ? rc : tg_mul(out, mask, drop, (long long)B * T * 2 * H, stream);
}

extern "C" int tg_gru_layer_fwd_tf32(const float* gi, const float* whh_f, const float* whh_r, const float* bhh_f, const float* bhh_r,
                                     float* out, float* saved, long long saved_qstride, int* sync, int B, int T, int H, tg_stream stream) {
  TG_REQUIRE(gi && whh_f && whh_r && bhh_f && bhh_r && out && sync && T > 0 && B > 0, "tg_gru_layer_fwd_tf32");
  cudaStream_t s = (cudaStream_t)stream;
  TgGruClPlan cpl;
  if (!use_legacy_gru() && tg_gru_cl_plan(B, H, &cpl))
    return tg_gru_cl_fwd(gi, whh_f, whh_r, bhh_f, bhh_r, out, saved, saved_qstride, nullptr, nullptr, reinterpret_cast<float*>(sync), B, T, H,
                         g_trace, s);
  FwdPlan pl;
  TG_REQUIRE(fwd_plan(B, H, &pl) == 0, "tg_gru_layer_fwd_tf32");
  cudaError_t e = cudaMemsetAsync(sync, 0, sizeof(int) * 2 * pl.NB, s);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_fwd_tf32: memset: %s", cudaGetErrorString(e)); return -2; }
  CUtensorMap maps[4];
  int rc;
  if ((rc = map_2d(&maps[0], whh_f, 3ll * H, H, H, pl.u, "tg_gru_layer_fwd_tf32(W)"))) return rc;
  if ((rc = map_2d(&maps[1], whh_r, 3ll * H, H, H, pl.u, "tg_gru_layer_fwd_tf32(W)"))) return rc;
  if ((rc = map_h3d(&maps[2], out, B, T, H, pl.BT, "tg_gru_layer_fwd_tf32(h)"))) return rc;
  if ((rc = map_h3d(&maps[3], out + H, B, T, H, pl.BT, "tg_gru_layer_fwd_tf32(h)"))) return rc;
  FwdP p;
  p.gi = gi; p.bhh[0] = bhh_f; p.bhh[1] = bhh_r; p.out = out; p.saved = saved; p.saved_qstride = saved_qstride; p.sync = sync;
  p.B = B; p.T = T; p.H = H; p.u = pl.u; p.UC = pl.UC; p.NB = pl.NB; p.ntiles = pl.ntiles; p.nkc = pl.nkc;
  p.trace = g_trace;
  if (pl.BT == 16) return launch_fwd<16>(maps, p, pl, s);
  if (pl.BT == 32) return launch_fwd<32>(maps, p, pl, s);
  return launch_fwd<48>(maps, p, pl, s);
}

extern "C" size_t tg_gru_bwd_tf32_scratch_floats(int B, int H) {
  const int UC = tg_ceil_div(H, BU), HP = (H + 3) & ~3;
  return (size_t)2 * 2 * B * UC * HP;
}

static int bwd_rows_per_cta(int B, int H) {
  const int UC = tg_ceil_div(H, BU);
  const int cands[3] = {32, 64, 128};
  for (int i = 0; i < 3; ++i)
    if (2 * UC * tg_ceil_div(B, cands[i]) <= tg_num_sms()) return cands[i];
  return 0;
}

template <int RB>
static int launch_bwd(const CUtensorMap* maps, const BwdP& p, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(gru_bwd_tc_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_bwd_tf32: smem attr: %s", cudaGetErrorString(e)); return -3; }
  void* args[] = {(void*)&maps[0], (void*)&maps[1], (void*)&p};
  dim3 grid(p.UC, p.NB, 2);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &cap);
  if (cap != cudaStreamCaptureStatusNone) e = cudaLaunchKernel((const void*)gru_bwd_tc_kernel<RB>, grid, dim3(320), args, smem, s);
  else e = cudaLaunchCooperativeKernel((const void*)gru_bwd_tc_kernel<RB>, grid, dim3(320), args, smem, s);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_bwd_tf32: cooperative launch (%u,%u,2): %s", grid.x, grid.y, cudaGetErrorString(e)); return -2; }
  return 0;
}

extern "C" int tg_gru_layer_bwd_tf32(const float* dout, const float* out, const float* saved, long long saved_qstride,
                                     const float* whhT_f, const float* whhT_r, float* dgi, float* dgh, float* partial, int* sync,
                                     int B, int T, int H, tg_stream stream) {
  TG_REQUIRE(dout && out && saved && whhT_f && whhT_r && dgi && dgh && partial && sync && T > 0 && B > 0, "tg_gru_layer_bwd_tf32");
  TG_REQUIRE(H >= 32 && H <= 384 && (H & 3) == 0, "tg_gru_layer_bwd_tf32");
  TgGruClPlan cpl;
  if (!use_legacy_gru() && tg_gru_cl_plan(B, H, &cpl))
    return tg_gru_cl_bwd(dout, out, saved, saved_qstride, whhT_f, whhT_r, dgi, dgh, reinterpret_cast<float*>(sync), B, T, H, g_trace,
                         (cudaStream_t)stream);
  BwdP p;
  p.UC = tg_ceil_div(H, BU);
  const int RB = bwd_rows_per_cta(B, H);
  TG_REQUIRE(RB > 0, "tg_gru_layer_bwd_tf32");
  p.NB = tg_ceil_div(B, RB);
  p.nh = H > 256 ? 2 : 1;
  p.Nh = ((tg_ceil_div(H, p.nh) + 15) / 16) * 16;
  p.HP = (H + 3) & ~3;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sync, 0, sizeof(int) * 2 * p.NB, s);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_bwd_tf32: memset: %s", cudaGetErrorString(e)); return -2; }
  CUtensorMap maps[2];
  int rc;
  if ((rc = map_2d(&maps[0], whhT_f, H, 3ll * H, 3ll * H, p.Nh, "tg_gru_layer_bwd_tf32(W^T)"))) return rc;
  if ((rc = map_2d(&maps[1], whhT_r, H, 3ll * H, 3ll * H, p.Nh, "tg_gru_layer_bwd_tf32(W^T)"))) return rc;
  p.dout = dout; p.out = out; p.saved = saved; p.saved_qstride = saved_qstride; p.dgi = dgi; p.dgh = dgh; p.partial = partial; p.sync = sync;
  p.B = B; p.T = T; p.H = H; p.trace = g_trace;
  const size_t smem = (size_t)3 * p.nh * p.Nh * 128 + 3 * 128 * 128 + 8 * 8 + 16 + 1024;
  TG_REQUIRE(smem <= (size_t)tg_max_smem_optin(), "tg_gru_layer_bwd_tf32");
  if (RB == 32) return launch_bwd<32>(maps, p, smem, s);
  if (RB == 64) return launch_bwd<64>(maps, p, smem, s);
  return launch_bwd<128>(maps, p, smem, s);
}
