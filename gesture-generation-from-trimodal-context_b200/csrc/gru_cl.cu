// Bidirectional-GRU recurrence as a thread-block-cluster kernel (fast mode; nn.GRU at multimodal_context_net.py:98-99,155-156
// and its backward).  One cluster of 8 CTAs owns one (batch tile, direction) chain for all T steps; clusters never talk to
// each other, so there is no grid-wide co-residency requirement, no cooperative launch and no spin-wait on L2 counters.
//
//  per step, every CTA of a cluster needs the WHOLE exchanged tile (forward: h_{t-1} [BT x H]; backward: dgh_{t+1} [BT x 3H]) as the
//  B operand of its tcgen05.mma, while it produces only its own 1/8 slice of the next one (forward: H/8 hidden units; backward: the
//  gate gradients of H/8 hidden units).  The all-gather runs through L2:
//    owners st.global their 16-byte / 4-byte pieces into an IMAGE of the swizzled K-major operand tile (double buffered)
//    -> fence.proxy.async -> barrier.cluster.arrive.release / wait.acquire (one hardware barrier per step)
//    -> CTA g of the cluster issues ONE cp.async.bulk ... .multicast::cluster of chunk group g: L2 is read once per cluster and the
//       bytes land in all 8 CTAs' operand tiles, completing each CTA's own mbarrier -> the MMA warp starts on group 0 while later
//       groups are still in flight.
//  A operand (recurrent weights, forward: rows r|z|n of this CTA's units, 3u x H; backward: W_hh^T rows of this CTA's units, u x 3H)
//  is loaded once by TMA and stays resident in shared memory for all T steps.
//
//  K is cut into 32-float chunks (one 128-byte swizzle row each); a remainder of <= 8 / <= 16 floats is kept as ONE narrower chunk in
//  the 32-byte / 64-byte swizzle layout instead of a zero-padded 128-byte one.  That is what makes 64 clips (forward, H = 300: 223 KB)
//  and 24 clips (backward, K = 900: 226 KB) per cluster fit in shared memory - a B200 keeps 15 clusters of 8 CTAs resident (measured:
//  one GPC has fewer than 16 free SMs), so 384 x 2 / 64 = 12 and 128 x 2 / 24 = 12 chains run as ONE wave where 48 / 16 clips need two.
//
//  forward : D[gate row (TMEM lane), clip] -> regrouped per hidden unit through shared memory -> sigmoid / tanh / state update in
//            registers (thread = clip x 4 consecutive units, h_{t-1} kept in registers across steps).
//  backward: M = 64 accumulator (rows 16q..16q+15 live in TMEM lanes 32q..32q+15): thread = (hidden unit, BT/8 clips) takes
//            (W_hh^T dgh_{t+1})[unit, clip] straight from TMEM (+ one shuffle) - no shared-memory staging, no partial sums through L2.
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int CL = 8;                 // CTAs per cluster = unit chunks per (tile, direction)

__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// global -> the same shared-memory offset of every CTA in `mask`; each destination CTA's mbarrier (same offset) gets complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t bar_saddr, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst_saddr),
               "l"(src), "r"(bytes), "r"(bar_saddr), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void stamp(long long* trace, int step, int slot) {
  if (trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    trace[step * 16 + slot] = (long long)t;
  }
}
__device__ __forceinline__ float ldv_nc(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldv_nc4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// 32 lanes x N consecutive columns of TMEM -> N registers per thread (thread = TMEM lane of its warp's quarter)
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
  static_assert(N == 4 || N == 6 || N == 8 || N == 12 || N == 16, "column count");
  if constexpr (N == 4) { tmem_ld4(taddr, v); }
  else if constexpr (N == 6) { tmem_ld4(taddr, v); tmem_ld2(taddr + 4, v + 4); }
  else if constexpr (N == 8) { tmem_ld8(taddr, v); }
  else if constexpr (N == 12) { tmem_ld8(taddr, v); tmem_ld4(taddr + 8, v + 8); }
  else { tmem_ld8(taddr, v); tmem_ld8(taddr + 8, v + 8); }
}

// ---- K chunking of one operand pair: nfull 128-byte-swizzle chunks of 32 floats + an optional narrower tail chunk
struct KLay {
  int nfull;          // full chunks
  int last_ksteps;    // tcgen05.mma K-steps (8 floats each) of the last full chunk (1..4)
  int tail_w;         // floats per row of the tail chunk: 0 (none), 8 (32-byte swizzle) or 16 (64-byte swizzle)
  int tail_ksteps;    // K-steps of the tail chunk
  int gsz, ngf;       // full chunks per multicast group, number of full groups (ngf + (tail_w > 0) <= 8)
};
// byte offset of float k of row `row` inside the IMAGE of a B operand tile: [nfull][rows][128 B] then [rows][tail_w * 4 B]; every
// 16-byte unit j of a row sits at j ^ (row bits selected by the swizzle width)
__device__ __forceinline__ uint32_t image_off(int row, int k, int rows, const KLay& L) {
  const int kc = k >> 5;
  if (kc < L.nfull)
    return (uint32_t)(kc * rows * 128 + (row >> 3) * 1024 + (row & 7) * 128 + ((((k & 31) >> 2) ^ (row & 7)) << 4) + (k & 3) * 4);
  const int kl = k - L.nfull * 32;
  const uint32_t base = (uint32_t)(L.nfull * rows * 128);
  if (L.tail_w == 16) return base + (uint32_t)(row * 64 + (((kl >> 2) ^ ((row >> 1) & 3)) << 4) + (kl & 3) * 4);
  return base + (uint32_t)(row * 32 + (((kl >> 2) ^ ((row >> 2) & 1)) << 4) + (kl & 3) * 4);
}

// One elected thread: MMAs of full-chunk group g (or of the tail chunk) of D[tmem] (+)= A * B.  a_chunk / b_chunk = bytes per full chunk of
// the A / B tile; K-major operands, descriptor halves per umma.cuh (lo walks the tile: +2 per 8-float K-step, + chunk bytes / 16 per chunk).
template <int NK>
__device__ __forceinline__ void issue_ksteps(int nk, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  if (nk == 4) mma_tf32_seq<4, 2>(tmem_d, a_lo, hi, b_lo, hi, idesc, acc);
  else if (nk == 3) mma_tf32_seq<3, 2>(tmem_d, a_lo, hi, b_lo, hi, idesc, acc);
  else if (nk == 2) mma_tf32_seq<2, 2>(tmem_d, a_lo, hi, b_lo, hi, idesc, acc);
  else mma_tf32_seq<1, 2>(tmem_d, a_lo, hi, b_lo, hi, idesc, acc);
}
__device__ __forceinline__ void issue_full_group(uint32_t tmem_d, uint32_t a_base, uint32_t a_chunk, uint32_t b_base, uint32_t b_chunk, int g,
                                                 const KLay& L, uint32_t idesc) {
  constexpr uint32_t hi = desc_hi(1024, 2);                 // SWIZZLE_128B, 8 rows x 128 B between row groups
  const int kc_lo = g * L.gsz, kc_hi = min(L.nfull, (g + 1) * L.gsz);
  uint32_t a_lo = desc_lo(a_base + (uint32_t)kc_lo * a_chunk, 16), b_lo = desc_lo(b_base + (uint32_t)kc_lo * b_chunk, 16);
  for (int kc = kc_lo; kc < kc_hi; ++kc, a_lo += a_chunk >> 4, b_lo += b_chunk >> 4) {
    if (kc < L.nfull - 1) mma_tf32_seq<4, 2>(tmem_d, a_lo, hi, b_lo, hi, idesc, kc > 0 ? 1u : 0u);
    else issue_ksteps<0>(L.last_ksteps, tmem_d, a_lo, b_lo, hi, idesc, kc > 0 ? 1u : 0u);
  }
}
__device__ __forceinline__ void issue_tail(uint32_t tmem_d, uint32_t a_tail, uint32_t b_tail, const KLay& L, uint32_t idesc) {
  // rows of tail_w floats: 64-byte swizzle (layout 4, SBO 512) for 16 floats, 32-byte swizzle (layout 6, SBO 256) for 8
  const uint32_t hi = L.tail_w == 16 ? desc_hi(512, 4) : desc_hi(256, 6);
  issue_ksteps<0>(L.tail_ksteps, tmem_d, desc_lo(a_tail, 16), desc_lo(b_tail, 16), hi, idesc, L.nfull > 0 ? 1u : 0u);
}
// The same with the A operand resident in tensor memory: K-step ks of the whole reduction reads A columns a_tmem + 8 ks.
template <int NK>
__device__ __forceinline__ void issue_ts_fixed(uint32_t tmem_base, uint32_t a_tmem, int ks0, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  mma_tf32_ts_seq<NK>(tmem_base, a_tmem + 8u * ks0, b_lo, hi, idesc, ks0 > 0 ? 1u : 0u);
}
__device__ __forceinline__ void issue_ts_n(int nk, uint32_t tmem_base, uint32_t a_tmem, int ks0, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  if (nk == 4) issue_ts_fixed<4>(tmem_base, a_tmem, ks0, b_lo, hi, idesc);
  else if (nk == 3) issue_ts_fixed<3>(tmem_base, a_tmem, ks0, b_lo, hi, idesc);
  else if (nk == 2) issue_ts_fixed<2>(tmem_base, a_tmem, ks0, b_lo, hi, idesc);
  else issue_ts_fixed<1>(tmem_base, a_tmem, ks0, b_lo, hi, idesc);
}
__device__ __forceinline__ void issue_full_group_ts(uint32_t tmem_base, uint32_t a_tmem, uint32_t b_base, uint32_t b_chunk, int g, const KLay& L,
                                                    uint32_t idesc) {
  constexpr uint32_t hi = desc_hi(1024, 2);
  const int kc_lo = g * L.gsz, kc_hi = min(L.nfull, (g + 1) * L.gsz);
  for (int kc = kc_lo; kc < kc_hi; ++kc) {
    const uint32_t b_lo = desc_lo(b_base + (uint32_t)kc * b_chunk, 16);
    const int nk = kc < L.nfull - 1 ? 4 : L.last_ksteps;
    issue_ts_n(nk, tmem_base, a_tmem, kc * 4, b_lo, hi, idesc);
  }
}
__device__ __forceinline__ void issue_tail_ts(uint32_t tmem_base, uint32_t a_tmem, uint32_t b_tail, const KLay& L, uint32_t idesc) {
  const uint32_t hi = L.tail_w == 16 ? desc_hi(512, 4) : desc_hi(256, 6);
  issue_ts_n(L.tail_ksteps, tmem_base, a_tmem, L.nfull * 4, desc_lo(b_tail, 16), hi, idesc);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t bar_saddr) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr), "l"(src), "r"(bytes),
               "r"(bar_saddr)
               : "memory");
}
// One elected thread, after the cluster barrier of step s: arm every group's mbarrier of THIS CTA, multicast the group this CTA's rank owns
// (unicast: fetch every group for this CTA alone)
__device__ __forceinline__ void issue_loads(uint32_t b_base, uint32_t b_chunk, uint32_t b_tail, uint32_t bar0, const uint8_t* img, int rows,
                                            const KLay& L, int rank, bool unicast) {
  for (int g = 0; g < L.ngf; ++g) {
    const int kc_lo = g * L.gsz, kc_hi = min(L.nfull, (g + 1) * L.gsz);
    const uint32_t bytes = (uint32_t)(kc_hi - kc_lo) * b_chunk;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * g), "r"(bytes) : "memory");
    if (unicast) bulk_g2s(b_base + (uint32_t)kc_lo * b_chunk, img + (size_t)kc_lo * b_chunk, bytes, bar0 + 8u * g);
    else if (g == rank) bulk_g2s_multicast(b_base + (uint32_t)kc_lo * b_chunk, img + (size_t)kc_lo * b_chunk, bytes, bar0 + 8u * g, (uint16_t)0xFF);
  }
  if (L.tail_w > 0) {
    const uint32_t bytes = (uint32_t)(rows * L.tail_w * 4);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * L.ngf), "r"(bytes) : "memory");
    if (unicast) bulk_g2s(b_tail, img + (size_t)L.nfull * b_chunk, bytes, bar0 + 8u * L.ngf);
    else if (L.ngf == rank) bulk_g2s_multicast(b_tail, img + (size_t)L.nfull * b_chunk, bytes, bar0 + 8u * L.ngf, (uint16_t)0xFF);
  }
}

// shared-memory offset of the barrier block = size of the operand regions
// forward: [h full chunks | gate scratch][h tail] - the recurrent weights live in tensor memory
__host__ __device__ inline size_t fwd_bar_off(const KLay& L, int u, int BT) {
  (void)u;
  size_t hreg = (size_t)L.nfull * BT * 128;
  if (hreg < (size_t)128 * (BT + 1) * 4) hreg = (size_t)128 * (BT + 1) * 4;
  hreg = (hreg + 1023) & ~(size_t)1023;
  return (hreg + (size_t)BT * L.tail_w * 4 + 1023) & ~(size_t)1023;
}
// backward: [W^T full chunks][dgh full chunks][W^T tail][dgh tail], including what the M = 64 MMA reads past the u packed A rows
__host__ __device__ inline size_t bwd_bar_off(const KLay& L, int u, int BT) {
  size_t greg = (size_t)L.nfull * BT * 128;
  if (greg < (size_t)64 * 128) greg = (size_t)64 * 128;
  greg = (greg + 1023) & ~(size_t)1023;
  const size_t w_full_bytes = (size_t)L.nfull * u * 128;
  const size_t wtail_off = w_full_bytes + greg;
  size_t end = wtail_off + (size_t)(u + BT) * L.tail_w * 4;
  if (L.nfull > 0 && w_full_bytes - (size_t)u * 128 + 64 * 128 > end) end = w_full_bytes - (size_t)u * 128 + 64 * 128;
  if (wtail_off + (size_t)64 * L.tail_w * 4 > end) end = wtail_off + (size_t)64 * L.tail_w * 4;
  return (end + 1023) & ~(size_t)1023;
}

// =====================================================================================================================
// forward
// =====================================================================================================================
struct FwdP {
  const float* gi; const float* whh0; const float* whh1; const float* bhh0; const float* bhh1; float* out; float* saved; long long saved_qstride;
  const float* mask; float* drop;      // optional: drop = out * mask (the inter-layer dropout, stored beside `out`; was a separate 19 us pass)
  float* xchg;
  int B, T, H, u, flags;      // flags bit 0: every CTA loads the whole tile itself (unicast) instead of the 8-way multicast
  KLay L;
  long long* trace;
};

template <int BT> struct FwdCfg {
  static constexpr int EPI_WARPS = BT == 64 ? 20 : 16;      // owners = BT clips x (u/4 <= 10) quads
  static constexpr int NT = 64 + 32 * EPI_WARPS;
};

// tensor-memory columns: the accumulator (64 columns, BT <= 64 used), then the resident weights (one column per k, 16-column stores)
constexpr uint32_t FWD_A0 = 64;
__host__ __device__ inline uint32_t fwd_tmem_cols(int H) {
  const uint32_t need = FWD_A0 + (uint32_t)((H + 15) / 16) * 16;
  uint32_t c = 32;
  while (c < need) c <<= 1;
  return c;
}

// HC = compile-time hidden size (300: the generator; K layout 9 full chunks + a 16-float tail, 5 + 1 multicast groups) or 0 (any H <= 320)
template <int BT, int HC>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(FwdCfg<BT>::NT, 1) gru_fwd_cl_kernel(const FwdP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NT = FwdCfg<BT>::NT, EPI = 32 * FwdCfg<BT>::EPI_WARPS;
  constexpr int H_CHUNK = BT * 128;
  constexpr int GS = BT + 1;
  constexpr int CPP = BT / 4;                  // accumulator columns moved by one staging warp
  const KLay L = p.L;
  const int H = p.H, T = p.T, u = p.u;
  // [h full chunks | gate scratch][h tail][barriers]
  size_t hreg_bytes = (size_t)L.nfull * H_CHUNK;
  if (hreg_bytes < (size_t)128 * GS * 4) hreg_bytes = (size_t)128 * GS * 4;
  hreg_bytes = (hreg_bytes + 1023) & ~(size_t)1023;
  uint8_t* Ht = smem;
  float* ghs = reinterpret_cast<float*>(Ht);
  uint8_t* Htail = Ht + hreg_bytes;
  const size_t bar_off = fwd_bar_off(L, u, BT);
  uint64_t* h_full = reinterpret_cast<uint64_t*>(smem + bar_off);   // [ngf + tail <= 8]
  uint64_t* tmem_full = h_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();     // == blockIdx.x (grid.x == CL)
  const int tile = blockIdx.y, dir = blockIdx.z;
  const int u0 = rank * u;
  const uint32_t tmem_cols = fwd_tmem_cols(H);
  const size_t img_bytes = (size_t)L.nfull * H_CHUNK + (size_t)BT * L.tail_w * 4;
  // image of this cluster's operand tile in global memory: [parity][full chunks][tail]
  uint8_t* img = reinterpret_cast<uint8_t*>(p.xchg) + ((size_t)dir * gridDim.y + tile) * 2 * img_bytes;

  if (smem_u32(smem) & 1023) __trap();
  // the operand tile starts finite; the image starts zero (K padding, dead clips)
  for (int i = threadIdx.x; i < (int)(bar_off / 16); i += NT) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  {
    const int n16 = (int)(2 * img_bytes / 16);
    for (int i = rank * NT + threadIdx.x; i < n16; i += CL * NT) reinterpret_cast<float4*>(img)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (warp == 0 && elect_one()) {
    for (int g = 0; g < 8; ++g) mbar_init(&h_full[g], 1);
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  fence_proxy_async_all();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (warp >= 2 && warp < 6) {
    // resident recurrent weights -> tensor memory, once: TMEM lane lr = g*u + jj holds row g*H + u0 + jj of W_hh (g = r, z, n), one
    // column per k; lanes >= 3u, units >= H and columns >= H are zero.  Each thread streams its own 1.2 KB row (L2-resident).
    const int lr0 = (warp & 3) * 32 + lane;
    const int g = lr0 / u, jj = lr0 - g * u;
    const bool row_live = lr0 < 3 * u && u0 + jj < H;
    const float* wrow = (dir ? p.whh1 : p.whh0) + ((long long)g * H + u0 + jj) * H;
    const int kcols = ((H + 7) / 8) * 8;
    for (int k0 = 0; k0 < kcols; k0 += 16) {
      float v[16];
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_live && k0 + e < H) x = ldv_nc4(wrow + k0 + e);
        v[e] = round_tf32(x.x); v[e + 1] = round_tf32(x.y); v[e + 2] = round_tf32(x.z); v[e + 3] = round_tf32(x.w);   // nearest, not truncated
      }
      tmem_st16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + FWD_A0 + (uint32_t)k0, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  cluster_arrive();          // every CTA's barriers are initialised and the image is zero before any peer multicasts / writes
  cluster_wait();

  // ---- ownership: epilogue thread etid < BT * u/4 owns (clip bb, units 4*quad .. 4*quad+3 of this CTA's chunk) for all steps
  const int etid = (int)threadIdx.x - 64;
  const int nquad = u >> 2;
  const bool epi = etid >= 0;
  const bool owner = epi && etid < BT * nquad;
  const int bb = owner ? etid / nquad : 0;
  const int quad = owner ? etid - bb * nquad : 0;
  const int b = tile * BT + bb;
  const int unit0 = u0 + quad * 4;
  const bool live = owner && unit0 < H && b < p.B;                   // H % 4 == 0: a quad is entirely valid or entirely out of range
  const float* bhh = dir ? p.bhh1 : p.bhh0;
  float4 bh_r = make_float4(0.f, 0.f, 0.f, 0.f), bh_z = bh_r, bh_n = bh_r;
  if (live) { bh_r = ldv_nc4(bhh + unit0); bh_z = ldv_nc4(bhh + H + unit0); bh_n = ldv_nc4(bhh + 2 * H + unit0); }
  float hreg[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t slice_off = live ? image_off(bb, unit0, BT, L) : 0u;
  const int q = warp & 3, part = epi ? (warp - 2) >> 2 : 0;
  const bool stager = epi && part < 4;                               // 16 warps move the accumulator; BT = 64 has 4 more that only compute gates
  const int lr = q * 32 + lane;
  const long long row2H = 2ll * H;
  constexpr uint32_t idesc = idesc_tf32(128, BT, 0, 0);
  const uint32_t ht_s = smem_u32(Ht), htail_s = smem_u32(Htail), hbar_s = smem_u32(h_full);
  const int ngroups = L.ngf + (L.tail_w > 0 ? 1 : 0);

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    float4 gi_r = make_float4(0.f, 0.f, 0.f, 0.f), gi_z = gi_r, gi_n = gi_r;
    if (live) {                                                      // issued before the accumulator wait: in flight during it
      const float* gip = p.gi + ((long long)b * T + t) * 6 * H + dir * 3 * H + unit0;
      gi_r = ldv_nc4(gip); gi_z = ldv_nc4(gip + H); gi_n = ldv_nc4(gip + 2 * H);
    }
    float4 mk4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (live && p.mask) mk4 = ldv_nc4(p.mask + ((long long)b * T + t) * row2H + dir * H + unit0);
    if (s > 0) {
      if (warp == 1) {
        if constexpr (HC == 300) {
          // straight-line issue with immediate offsets (tests/ubench/mma_issue.cu: ~34 cycles per 128 x 64 x 8 MMA, the tensor pipe's own rate;
          // the same MMAs issued through run-time loops over chunks and K-steps took 80-110 cycles each)
          constexpr uint32_t hi = desc_hi(1024, 2), hi_tail = desc_hi(512, 4);
          const uint32_t a0 = tmem_base + FWD_A0, b0 = desc_lo(ht_s, 16), bt0 = desc_lo(htail_s, 16);
#pragma unroll
          for (int g = 0; g < 6; ++g) {
            mbar_wait(&h_full[g], (uint32_t)((s - 1) & 1));
            tc_fence_after();
            if (elect_one()) {
              stamp(p.trace, s, 8 + g);
              if (g < 5) {
#pragma unroll
                for (int kc = 2 * g; kc < (2 * g + 2 < 9 ? 2 * g + 2 : 9); ++kc)
                  mma_tf32_ts_seq<4>(tmem_base, a0 + 32u * kc, b0 + (uint32_t)kc * (H_CHUNK / 16), hi, idesc, kc > 0 ? 1u : 0u);
              } else {
                mma_tf32_ts_seq<2>(tmem_base, a0 + 32u * 9, bt0, hi_tail, idesc, 1u);
              }
            }
            __syncwarp();
          }
        } else {
          for (int g = 0; g < ngroups; ++g) {
            mbar_wait(&h_full[g], (uint32_t)((s - 1) & 1));
            tc_fence_after();
            if (elect_one()) {
              stamp(p.trace, s, 8 + g);
              if (g < L.ngf) issue_full_group_ts(tmem_base, tmem_base + FWD_A0, ht_s, H_CHUNK, g, L, idesc);
              else issue_tail_ts(tmem_base, tmem_base + FWD_A0, htail_s, L, idesc);
            }
            __syncwarp();
          }
        }
        if (elect_one()) { tc_commit(tmem_full); stamp(p.trace, s, 7); }
        __syncwarp();
      }
      if (epi) {
        // 16 epilogue warps move the accumulator TMEM -> gate scratch: warp (q, part) takes lanes 32q.. and columns part*CPP..
        if (stager) {
          mbar_wait(tmem_full, (uint32_t)((s - 1) & 1));
          if (etid == 0) stamp(p.trace, s, 1);
          tc_fence_after();
          float v[CPP];
          tmem_ld_cols<CPP>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * CPP), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CPP; ++j) ghs[lr * GS + part * CPP + j] = v[j];
          tc_fence_before();
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
        if (etid == 0) stamp(p.trace, s, 2);
      }
    }
    float rr[4], zz[4], nn[4], hn[4];
    if (live) {
      const float gir[4] = {gi_r.x, gi_r.y, gi_r.z, gi_r.w}, giz[4] = {gi_z.x, gi_z.y, gi_z.z, gi_z.w}, gin[4] = {gi_n.x, gi_n.y, gi_n.z, gi_n.w};
      const float br[4] = {bh_r.x, bh_r.y, bh_r.z, bh_r.w}, bz[4] = {bh_z.x, bh_z.y, bh_z.z, bh_z.w}, bn[4] = {bh_n.x, bh_n.y, bh_n.z, bh_n.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int jj = quad * 4 + e;
        float ghr = br[e], ghz = bz[e], ghn = bn[e];
        if (s > 0) { ghr += ghs[jj * GS + bb]; ghz += ghs[(u + jj) * GS + bb]; ghn += ghs[(2 * u + jj) * GS + bb]; }
        rr[e] = sigmoidf_(gir[e] + ghr);
        zz[e] = sigmoidf_(giz[e] + ghz);
        nn[e] = tanhf_(gin[e] + rr[e] * ghn);
        hn[e] = ghn;
        hreg[e] = (1.f - zz[e]) * nn[e] + zz[e] * hreg[e];
      }
    }
    if (etid == 0) stamp(p.trace, s, 3);
    if (s + 1 < T) {
      // h_t slice -> image[s & 1] (dead clips keep the zeros the image started with)
      // the MMA operand copy of h_t is rounded to the nearest TF32 (the state carried in registers and the stored output stay fp32)
      if (live)
        *reinterpret_cast<float4*>(img + (size_t)(s & 1) * img_bytes + slice_off) =
            make_float4(round_tf32(hreg[0]), round_tf32(hreg[1]), round_tf32(hreg[2]), round_tf32(hreg[3]));
      fence_proxy_async_all();   // generic-proxy writes (image; gate scratch aliasing the operand tile) -> async-proxy readers / writers
      cluster_arrive();          // release: this CTA's slice is out and it no longer reads its gate scratch
      if (etid == 0) stamp(p.trace, s, 4);
    }
    if (live) {                  // stores for the next layer / the backward pass: between arrive and wait, off the step chain
      const long long o = ((long long)b * T + t) * row2H + dir * H + unit0;
      *reinterpret_cast<float4*>(p.out + o) = make_float4(hreg[0], hreg[1], hreg[2], hreg[3]);
      if (p.drop) *reinterpret_cast<float4*>(p.drop + o) = make_float4(hreg[0] * mk4.x, hreg[1] * mk4.y, hreg[2] * mk4.z, hreg[3] * mk4.w);
      if (p.saved) {
        *reinterpret_cast<float4*>(p.saved + o) = make_float4(rr[0], rr[1], rr[2], rr[3]);
        *reinterpret_cast<float4*>(p.saved + p.saved_qstride + o) = make_float4(zz[0], zz[1], zz[2], zz[3]);
        *reinterpret_cast<float4*>(p.saved + 2 * p.saved_qstride + o) = make_float4(nn[0], nn[1], nn[2], nn[3]);
        *reinterpret_cast<float4*>(p.saved + 3 * p.saved_qstride + o) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      }
    }
    if (s + 1 < T) {
      cluster_wait();            // acquire: every slice of h_t is in the image, every peer's operand tile is free
      if (etid == 0) stamp(p.trace, s, 5);
      if (warp == 0) {
        if (elect_one()) issue_loads(ht_s, H_CHUNK, htail_s, hbar_s, img + (size_t)(s & 1) * img_bytes, BT, L, rank, (p.flags & 1) != 0);
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
  cluster_arrive();            // no CTA exits while a peer's multicast may still address its shared memory / barriers
  cluster_wait();
}

// =====================================================================================================================
// backward
// =====================================================================================================================
struct BwdP {
  const float* dout; const float* out; const float* saved; long long saved_qstride;
  float* dgi; float* dgh; float* xchg;
  int B, T, H, u, flags;
  KLay L;
  long long* trace;
};
constexpr int BWD_NT = 576;           // warp 0: bulk-copy issue, warp 1: TMEM + MMA issue, warps 2-17: epilogue

template <int BT>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(BWD_NT, 1)
    gru_bwd_cl_kernel(const __grid_constant__ CUtensorMap tmT0, const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmV0,
                      const __grid_constant__ CUtensorMap tmV1, const BwdP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int G_CHUNK = BT * 128;
  constexpr int CPP = BT / 4;                  // clips per epilogue thread
  const KLay L = p.L;
  const int u = p.u;
  const uint32_t a_chunk = (uint32_t)u * 128;  // u % 8 == 0: every chunk base stays 1024-byte aligned
  // [W^T full chunks][dgh full chunks][W^T tail][dgh tail][pad][barriers]; the M = 64 MMA reads 64 rows from each A chunk base: rows >= u
  // belong to the next region and land in ignored TMEM lanes
  size_t greg_bytes = (size_t)L.nfull * G_CHUNK;
  if (greg_bytes < (size_t)64 * 128) greg_bytes = (size_t)64 * 128;
  greg_bytes = (greg_bytes + 1023) & ~(size_t)1023;
  uint8_t* Wt = smem;
  uint8_t* Gt = Wt + (size_t)L.nfull * a_chunk;
  uint8_t* Wtail = Gt + greg_bytes;
  uint8_t* Gtail = Wtail + (size_t)u * L.tail_w * 4;
  const size_t bar_off = bwd_bar_off(L, u, BT);
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + bar_off);
  uint64_t* g_full = w_full + 1;               // [ngf + tail <= 8]
  uint64_t* tmem_full = g_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int tile = blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T;
  const int u0 = rank * u;
  constexpr uint32_t TMEM_COLS = 32;
  const size_t img_bytes = (size_t)L.nfull * G_CHUNK + (size_t)BT * L.tail_w * 4;
  uint8_t* img = reinterpret_cast<uint8_t*>(p.xchg) + ((size_t)dir * gridDim.y + tile) * 2 * img_bytes;

  if (smem_u32(smem) & 1023) __trap();
  for (int i = threadIdx.x; i < (int)(bar_off / 16); i += BWD_NT) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  {
    const int n16 = (int)(2 * img_bytes / 16);
    for (int i = rank * BWD_NT + threadIdx.x; i < n16; i += CL * BWD_NT) reinterpret_cast<float4*>(img)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (warp == 0 && elect_one()) {
    mbar_init(w_full, 1);
    for (int g = 0; g < 8; ++g) mbar_init(&g_full[g], 1);
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_proxy_async_all();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (warp == 0 && elect_one()) {
    const CUtensorMap* tmT = dir ? &tmT1 : &tmT0;
    const CUtensorMap* tmV = dir ? &tmV1 : &tmV0;
    tma_prefetch_desc(tmT);
    // A[m = own unit j, k = gate row] = W_hh[k, u0 + j] = whhT[u0 + j][k]; rows >= H and columns >= 3H are zero-filled by TMA
    mbar_expect_tx(w_full, (uint32_t)(L.nfull * a_chunk + u * L.tail_w * 4));
    for (int kc = 0; kc < L.nfull; ++kc) tma_load_2d(Wt + (size_t)kc * a_chunk, tmT, w_full, kc * 32, u0);
    if (L.tail_w > 0) tma_load_2d(Wtail, tmV, w_full, L.nfull * 32, u0);
  }
  __syncwarp();
  if (warp == 1) mbar_wait(w_full, 0);
  cluster_arrive();
  cluster_wait();

  // ---- ownership: the M = 64 accumulator keeps row j in TMEM lane 32*(j/16) + j%16, so only lanes 0-15 of epilogue warp (q = warp & 3, part)
  // can read rows 16q .. 16q+15 for the CPP clips of `part`; lanes 16-31 get the second half of those clips by shuffle, so that every
  // lane owns (hidden unit j = 16q + lane%16 of this CTA's chunk) x (CPH = CPP/2 clips)
  const int etid = (int)threadIdx.x - 64;
  const bool epi = etid >= 0;
  const int q = warp & 3, part = epi ? (warp - 2) >> 2 : 0;
  const int jl = lane & 15, half = lane >> 4;
  const int j = q * 16 + jl;
  const int unit = u0 + j;
  constexpr int CPH = CPP / 2;
  const int c0 = part * CPP + half * CPH;                           // first clip (row of the tile) of this thread
  const bool row_ok = epi && j < u && unit < H;
  const bool q_used = epi && q * 16 < u;                            // warp-uniform: this warp's lane quarter holds accumulator rows
  const long long row2H = 2ll * H;
  float dhc[CPH];
#pragma unroll
  for (int i = 0; i < CPH; ++i) dhc[i] = 0.f;
  constexpr uint32_t idesc = idesc_tf32(64, BT, 0, 0);
  const uint32_t wt_s = smem_u32(Wt), gt_s = smem_u32(Gt), wtail_s = smem_u32(Wtail), gtail_s = smem_u32(Gtail), gbar_s = smem_u32(g_full);
  const int ngroups = L.ngf + (L.tail_w > 0 ? 1 : 0);

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? T - 1 - s : s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    const bool tp_ok = tp >= 0 && tp < T;
    // recurrence-independent operands of this step: loads issued before the accumulator wait
    float v_do[CPH], v_r[CPH], v_z[CPH], v_n[CPH], v_hn[CPH], v_hp[CPH];
#pragma unroll
    for (int i = 0; i < CPH; ++i) {
      v_do[i] = v_r[i] = v_z[i] = v_n[i] = v_hn[i] = v_hp[i] = 0.f;
      const int b = tile * BT + c0 + i;
      if (row_ok && b < p.B) {
        const long long o = ((long long)b * T + t) * row2H + dir * H + unit;
        v_do[i] = ldv_nc(p.dout + o);
        v_r[i] = ldv_nc(p.saved + o);
        v_z[i] = ldv_nc(p.saved + p.saved_qstride + o);
        v_n[i] = ldv_nc(p.saved + 2 * p.saved_qstride + o);
        v_hn[i] = ldv_nc(p.saved + 3 * p.saved_qstride + o);
        if (tp_ok) v_hp[i] = ldv_nc(p.out + ((long long)b * T + tp) * row2H + dir * H + unit);
      }
    }
    float acc[CPH];
#pragma unroll
    for (int i = 0; i < CPH; ++i) acc[i] = 0.f;
    if (s > 0) {
      if (warp == 1) {
        for (int g = 0; g < ngroups; ++g) {
          mbar_wait(&g_full[g], (uint32_t)((s - 1) & 1));
          tc_fence_after();
          if (elect_one()) {
            stamp(p.trace, s, 8 + g);
            if (g < L.ngf) issue_full_group(tmem_base, wt_s, a_chunk, gt_s, G_CHUNK, g, L, idesc);
            else issue_tail(tmem_base, wtail_s, gtail_s, L, idesc);
          }
          __syncwarp();
        }
        if (elect_one()) { tc_commit(tmem_full); stamp(p.trace, s, 7); }
        __syncwarp();
      }
      if (q_used) {
        mbar_wait(tmem_full, (uint32_t)((s - 1) & 1));
        if (etid == 0) stamp(p.trace, s, 1);
        tc_fence_after();
        float full[CPP];
        tmem_ld_cols<CPP>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * CPP), full);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int i = 0; i < CPH; ++i) {
          const float hi = __shfl_sync(0xffffffffu, full[CPH + i], jl);     // lanes 16-31 fetch the second half of the clips from lane - 16
          acc[i] = half ? hi : full[i];
        }
      }
    }
    float dr[CPH], dz[CPH], dn[CPH], dnr[CPH];
#pragma unroll
    for (int i = 0; i < CPH; ++i) {
      const float dh = v_do[i] + dhc[i] + acc[i];
      const float dnv = dh * (1.f - v_z[i]) * (1.f - v_n[i] * v_n[i]);
      dz[i] = dh * (v_hp[i] - v_n[i]) * v_z[i] * (1.f - v_z[i]);
      dr[i] = dnv * v_hn[i] * v_r[i] * (1.f - v_r[i]);
      dn[i] = dnv;
      dnr[i] = dnv * v_r[i];
      dhc[i] = dh * v_z[i];
    }
    if (etid == 0) stamp(p.trace, s, 3);
    if (s + 1 < T) {
      if (row_ok) {
        uint8_t* im = img + (size_t)(s & 1) * img_bytes;
#pragma unroll
        for (int i = 0; i < CPH; ++i) {
          const int c = c0 + i;
          if (tile * BT + c < p.B) {
            *reinterpret_cast<float*>(im + image_off(c, unit, BT, L)) = dr[i];
            *reinterpret_cast<float*>(im + image_off(c, H + unit, BT, L)) = dz[i];
            *reinterpret_cast<float*>(im + image_off(c, 2 * H + unit, BT, L)) = dnr[i];
          }
        }
      }
      fence_proxy_async_all();
      cluster_arrive();
      if (etid == 0) stamp(p.trace, s, 4);
    }
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < CPH; ++i) {
        const int b = tile * BT + c0 + i;
        if (b < p.B) {
          const long long row = (long long)b * T + t;
          float* gp = p.dgi + row * 6 * H + dir * 3 * H + unit;
          float* hp = p.dgh + row * 6 * H + dir * 3 * H + unit;
          gp[0] = dr[i]; gp[H] = dz[i]; gp[2 * H] = dn[i];
          hp[0] = dr[i]; hp[H] = dz[i]; hp[2 * H] = dnr[i];
        }
      }
    }
    if (s + 1 < T) {
      cluster_wait();
      if (etid == 0) stamp(p.trace, s, 5);
      if (warp == 0) {
        if (elect_one()) issue_loads(gt_s, G_CHUNK, gtail_s, gbar_s, img + (size_t)(s & 1) * img_bytes, BT, L, rank, (p.flags & 1) != 0);
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  cluster_arrive();
  cluster_wait();
}

// =====================================================================================================================
// backward, K-split ("own gates") variant
//
// dh_t[clip, j] needs sum_r dgh_{t+1}[clip, r] * W_hh[r, j] over all 3H gate rows r.  Instead of all-gathering dgh (3H values per clip,
// above), CTA c keeps the product for ITS OWN gate rows only - the 3u rows of the u hidden units whose gate gradients its threads have
// just computed - but for ALL H outputs j:  P_c[j, clip] = sum_{r in R_c} W_hh[r, j] * dgh[clip, r].  That is a [ceil(H/128) x 128] x BT
// x 3u product whose A operand (W_hh^T restricted to R_c: H x 3u) fits TENSOR MEMORY next to the accumulators, so no recurrent weight
// sits in shared memory at all and the B operand (BT x 3u) never leaves the CTA.  The partials are then reduce-scattered: the thread
// holding accumulator row j pushes its BT values straight into the shared memory of the CTA that owns unit j (st.shared::cluster),
// one hardware cluster barrier per step publishes them, and the owner adds the 8 contributions.
// =====================================================================================================================
struct BwdKsP {
  const float* dout; const float* out; const float* saved; long long saved_qstride; const float* whhT0; const float* whhT1;
  float* dgi; float* dgh;
  int B, T, H, u;
  long long* trace;
};
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
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
constexpr int KS_ACC_STRIDE = 32;     // tensor-memory columns reserved per accumulator tile (BT <= 32)
__host__ __device__ inline int ks_ka(int u) { return ((3 * u + 15) / 16) * 16; }          // A columns per M tile (16-column stores)
__host__ __device__ inline uint32_t ks_tmem_cols(int H, int u) {
  const uint32_t need = (uint32_t)((H + 127) / 128) * (KS_ACC_STRIDE + ks_ka(u));
  uint32_t c = 32;
  while (c < need) c <<= 1;
  return c;
}
// shared memory: [B tile: ceil(3u/32) chunks x BT rows x 128 B][stage: 2 parities x 8 destinations x u rows x (BT+4) floats]
//                [slots: 2 parities x 8 sources x u rows x (BT+4) floats][barriers]
__host__ __device__ inline size_t ks_btile_bytes(int u, int BT) { return (size_t)((3 * u + 31) / 32) * BT * 128; }
__host__ __device__ inline size_t ks_stage_off(int u, int BT) { return (ks_btile_bytes(u, BT) + 1023) & ~(size_t)1023; }
__host__ __device__ inline size_t ks_slots_off(int u, int BT) { return ks_stage_off(u, BT) + (((size_t)2 * CL * u * (BT + 4) * 4 + 127) & ~(size_t)127); }
__host__ __device__ inline size_t ks_bar_off(int u, int BT) { return (ks_slots_off(u, BT) + (size_t)2 * CL * u * (BT + 4) * 4 + 15) & ~(size_t)15; }

// shared::cta -> shared memory of CTA `dst_saddr` lives in (a mapa address), completing `bar_raddr` (a mapa address in the same CTA)
__device__ __forceinline__ void bulk_s2s_cluster(uint32_t dst_raddr, uint32_t src_saddr, uint32_t bytes, uint32_t bar_raddr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_raddr), "r"(src_saddr),
               "r"(bytes), "r"(bar_raddr)
               : "memory");
}

// the nmt x ksteps MMAs of one step, consecutive MMAs on different accumulator tiles; KSTEPS / NMT > 0: compile-time shape (straight-line
// issue with immediate offsets: ~10 cycles per MMA instead of ~50 through run-time loops - tests/ubench/mma_issue.cu)
template <int BT, int KSTEPS, int NMT>
__device__ __forceinline__ void ks_issue(uint32_t tmem_base, uint32_t a_col0, int ka, uint32_t bt_lo, uint32_t idesc, int ksteps_rt, int nmt_rt) {
  constexpr uint32_t hi = desc_hi(1024, 2);
  if constexpr (KSTEPS > 0) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const uint32_t b_lo = bt_lo + (uint32_t)(ks >> 2) * (BT * 128 / 16) + (uint32_t)(ks & 3) * 2;
#pragma unroll
      for (int m = 0; m < NMT; ++m)
        mma_tf32_ts(tmem_base + (uint32_t)(m * KS_ACC_STRIDE), tmem_base + a_col0 + (uint32_t)(m * ka) + 8u * ks, b_lo, hi, idesc, ks > 0 ? 1u : 0u);
    }
  } else {
    for (int ks = 0; ks < ksteps_rt; ++ks) {
      const uint32_t b_lo = bt_lo + (uint32_t)(ks >> 2) * (BT * 128 / 16) + (uint32_t)(ks & 3) * 2;
      for (int m = 0; m < nmt_rt; ++m)
        mma_tf32_ts(tmem_base + (uint32_t)(m * KS_ACC_STRIDE), tmem_base + a_col0 + (uint32_t)(m * ka) + 8u * ks, b_lo, hi, idesc, ks > 0 ? 1u : 0u);
    }
  }
}

// HC = compile-time hidden size (300: the generator) or 0 (any H <= 320)
template <int BT, int HC>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(BWD_NT, 1) gru_bwd_ks_kernel(const BwdKsP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  static_assert(BT == 16 || BT == 32, "batch tile");
  constexpr int SP = BT + 4;                   // slot / stage row pitch (floats): 16-byte aligned rows
  constexpr int CG = BT / 4;                   // clip groups of 4
  const int H = HC > 0 ? HC : p.H, T = p.T;
  const int u = HC > 0 ? ((HC + CL - 1) / CL + 7) / 8 * 8 : p.u;
  const int K = 3 * u;                         // own gate rows: k = g*u + jl  <->  W_hh row g*H + u0 + jl
  const int ksteps = K >> 3;                   // u % 8 == 0
  const int nmt = (H + 127) >> 7;              // accumulator row tiles (M = 128 each)
  const int ka = ks_ka(u);
  uint8_t* Bt = smem;
  float* stage = reinterpret_cast<float*>(smem + ks_stage_off(u, BT));
  float* slots = reinterpret_cast<float*>(smem + ks_slots_off(u, BT));
  uint64_t* tmem_full = reinterpret_cast<uint64_t*>(smem + ks_bar_off(u, BT));
  uint64_t* slot_full = tmem_full + 1;         // [2 parities]: 8 x (u x SP x 4) bytes of partials have landed in slots[parity]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int tile = blockIdx.y, dir = blockIdx.z;
  const int u0 = rank * u;
  const uint32_t tmem_cols = ks_tmem_cols(H, u);
  const uint32_t a_col0 = (uint32_t)nmt * KS_ACC_STRIDE;              // accumulators first, then the nmt A tiles of ka columns
  const uint32_t slice_bytes = (uint32_t)(u * SP * 4);                // one (source, destination) slice of partials

  if (smem_u32(smem) & 1023) __trap();
  for (int i = threadIdx.x; i < (int)(ks_bar_off(u, BT) / 16); i += BWD_NT) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0 && elect_one()) {
    mbar_init(tmem_full, 1); mbar_init(&slot_full[0], 1); mbar_init(&slot_full[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  fence_proxy_async_all();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (warp == 0 && elect_one()) {              // arm both parities for their first use (steps 1 and 2)
    mbar_expect_tx(&slot_full[0], CL * slice_bytes);
    mbar_expect_tx(&slot_full[1], CL * slice_bytes);
  }
  const int etid = (int)threadIdx.x - 64;
  const bool epi = etid >= 0;
  const int q = warp & 3, mt = epi ? (warp - 2) >> 2 : 0;               // epilogue warp (q, mt): TMEM lanes 32q.., accumulator / A tile mt
  const bool tile_warp = epi && mt < nmt;
  const int jrow = mt * 128 + q * 32 + lane;                            // output unit j held in this thread's TMEM lane
  if (tile_warp) {
    // resident weights -> tensor memory, once: lane j of tile mt, column k = g*u + jl holds W_hh[g*H + u0 + jl][j] = whhT[j][g*H + u0 + jl]
    const float* wrow = (dir ? p.whhT1 : p.whhT0) + (long long)jrow * 3 * H;
    for (int k0 = 0; k0 < ka; k0 += 16) {
      float v[16];
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        const int k = k0 + e, g = k / u, jl = k - g * u;               // u % 4 == 0: a group of 4 never straddles two gates
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (jrow < H && k < K && u0 + jl < H) x = ldv_nc4(wrow + (long long)g * H + u0 + jl);
        v[e] = round_tf32(x.x); v[e + 1] = round_tf32(x.y); v[e + 2] = round_tf32(x.z); v[e + 3] = round_tf32(x.w);
      }
      tmem_st16(tmem_base + ((uint32_t)(q * 32) << 16) + a_col0 + (uint32_t)(mt * ka + k0), v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  cluster_arrive();          // every CTA's slots are zeroed and its barriers armed before any peer copies into them
  cluster_wait();

  // ---- elementwise ownership: epilogue thread etid < CG * u owns (hidden unit u0 + jl, clips 4cg .. 4cg+3 of the tile)
  const int cg = epi ? etid / u : 0, jl = epi ? etid - cg * u : 0;
  const int unit = u0 + jl;
  const bool own = epi && cg < CG && unit < H;
  const long long row2H = 2ll * H;
  float dhc[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t idesc = idesc_tf32(128, BT, 0, 0);
  const uint32_t bt_lo = desc_lo(smem_u32(Bt), 16);
  const uint32_t stage_s = smem_u32(stage), slots_s = smem_u32(slots), sfull_s = smem_u32(slot_full);
  // where this thread's accumulator row is staged: destination = owner CTA of unit jrow, row jrow - owner*u of stage[owner]
  const int owner = jrow / u, jo = jrow - owner * u;
  float* stage_row = stage + ((size_t)(owner < CL ? owner : 0) * u + jo) * SP;      // + parity * CL * u * SP
  // Both staging and slots are double buffered by step parity.  No per-step cluster barrier: a CTA overwrites stage[p] / has slots[p]
  // overwritten two steps after their last use, and in between it has received every peer's next slice - which a peer only sends after
  // this CTA's previous slice has fully landed there (so the copy engine is done with stage[p]) and after it consumed its own slots[p].

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? T - 1 - s : s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    const bool tp_ok = tp >= 0 && tp < T;
    float v_do[4], v_r[4], v_z[4], v_n[4], v_hn[4], v_hp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v_do[i] = v_r[i] = v_z[i] = v_n[i] = v_hn[i] = v_hp[i] = 0.f;
      const int b = tile * BT + cg * 4 + i;
      if (own && b < p.B) {
        const long long o = ((long long)b * T + t) * row2H + dir * H + unit;
        v_do[i] = ldv_nc(p.dout + o);
        v_r[i] = ldv_nc(p.saved + o);
        v_z[i] = ldv_nc(p.saved + p.saved_qstride + o);
        v_n[i] = ldv_nc(p.saved + 2 * p.saved_qstride + o);
        v_hn[i] = ldv_nc(p.saved + 3 * p.saved_qstride + o);
        if (tp_ok) v_hp[i] = ldv_nc(p.out + ((long long)b * T + tp) * row2H + dir * H + unit);
      }
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (s > 0) {
      // all 8 partial slices of this step have landed in slots[s & 1] (bulk copies complete the barrier; phase (s-1)/2 of that parity)
      mbar_wait(&slot_full[s & 1], (uint32_t)(((s - 1) >> 1) & 1));
      if (etid == 0) stamp(p.trace, s, 5);
      if (own) {
        const float* sl = slots + (size_t)(s & 1) * CL * u * SP + (size_t)jl * SP + cg * 4;
#pragma unroll
        for (int src = 0; src < CL; ++src) {
          const float4 x = *reinterpret_cast<const float4*>(sl + (size_t)src * u * SP);
          acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
        }
      }
    }
    float dr[4], dz[4], dn[4], dnr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float dh = v_do[i] + dhc[i] + acc[i];
      const float dnv = dh * (1.f - v_z[i]) * (1.f - v_n[i] * v_n[i]);
      dz[i] = dh * (v_hp[i] - v_n[i]) * v_z[i] * (1.f - v_z[i]);
      dr[i] = dnv * v_hn[i] * v_r[i] * (1.f - v_r[i]);
      dn[i] = dnv;
      dnr[i] = dnv * v_r[i];
      dhc[i] = dh * v_z[i];
    }
    if (etid == 0) stamp(p.trace, s, 3);
    if (s + 1 < T) {
      if (epi) {
        if (own) {
          // B operand of the next product: this CTA's own gate gradients, K-major [clip][k = g*u + jl], 128-byte swizzle
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = cg * 4 + i;
            const uint32_t rowoff = (uint32_t)((c >> 3) * 1024 + (c & 7) * 128);
#pragma unroll
            for (int g = 0; g < 3; ++g) {
              const int k = g * u + jl;
              const float val = round_tf32(g == 0 ? dr[i] : g == 1 ? dz[i] : dnr[i]);
              *reinterpret_cast<float*>(Bt + (size_t)(k >> 5) * BT * 128 + rowoff + ((((k & 31) >> 2) ^ (c & 7)) << 4) + (k & 3) * 4) = val;
            }
          }
        }
        fence_proxy_async_smem();
      }
      if (warp >= 1) {
        tc_fence_before();
        asm volatile("bar.sync 1, 544;" ::: "memory");     // 16 epilogue warps + the MMA warp: the B tile is complete, slots[s & 1] are consumed
      }
      if (warp == 1) {
        tc_fence_after();
        if (elect_one()) {
          // re-arm this step's parity for its next use (step s + 2); its producers cannot send before they have this CTA's step s + 1 slice
          if (s > 0 && s + 2 < T) mbar_expect_tx(&slot_full[s & 1], CL * slice_bytes);
          stamp(p.trace, s, 6);
          ks_issue<BT, HC == 300 ? 15 : 0, HC == 300 ? 3 : 0>(tmem_base, a_col0, ka, bt_lo, idesc, ksteps, nmt);
          tc_commit(tmem_full);
          stamp(p.trace, s, 7);
        }
        __syncwarp();
      }
    }
    if (own) {                 // stores for the weight-gradient GEMMs: overlap the MMAs
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int b = tile * BT + cg * 4 + i;
        if (b < p.B) {
          const long long row = (long long)b * T + t;
          float* gp = p.dgi + row * 6 * H + dir * 3 * H + unit;
          float* hp = p.dgh + row * 6 * H + dir * 3 * H + unit;
          gp[0] = dr[i]; gp[H] = dz[i]; gp[2 * H] = dn[i];
          hp[0] = dr[i]; hp[H] = dz[i]; hp[2 * H] = dnr[i];
        }
      }
    }
    if (s + 1 < T) {
      if (epi) {
        if (tile_warp) {
          // partial dh of THIS CTA's gate rows for output unit jrow, all BT clips -> staged per destination CTA (the owner of unit jrow)
          mbar_wait(tmem_full, (uint32_t)(s & 1));
          if (etid == 0) stamp(p.trace, s, 1);
          tc_fence_after();
          float v[BT];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * KS_ACC_STRIDE);
          tmem_ld16(taddr, v);
          if constexpr (BT == 32) tmem_ld16(taddr + 16, v + 16);
          tmem_ld_wait();
          tc_fence_before();
          if (jrow < H) {
            float* sr = stage_row + (size_t)((s + 1) & 1) * CL * u * SP;
#pragma unroll
            for (int c4 = 0; c4 < BT / 4; ++c4) *reinterpret_cast<float4*>(sr + 4 * c4) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
          }
          fence_proxy_async_smem();      // generic-proxy staging writes -> the bulk copies' async-proxy reads
        }
        asm volatile("bar.sync 2, 512;" ::: "memory");        // the 16 epilogue warps: staging complete
        if (warp == 2 && elect_one()) {
          // reduce-scatter: slice d of the staged partials -> CTA d's slots[(s+1) & 1][source = this rank]; each copy completes CTA d's barrier
          const uint32_t par = (uint32_t)((s + 1) & 1);
          for (int d = 0; d < CL; ++d)
            bulk_s2s_cluster(mapa_cluster(slots_s + (par * CL + (uint32_t)rank) * slice_bytes, (uint32_t)d), stage_s + (par * CL + (uint32_t)d) * slice_bytes,
                             slice_bytes, mapa_cluster(sfull_s + 8u * par, (uint32_t)d));
          stamp(p.trace, s, 4);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
  cluster_arrive();            // no CTA exits while a peer's copy may still address its shared memory
  cluster_wait();
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

// row-major [rows, cols] fp32 matrix with leading dimension ld: box = box_cols (32 / 16 / 8 floats = one 128 / 64 / 32-byte swizzle row)
// columns x box_rows rows
int map_2d(CUtensorMap* m, const float* base, long long rows, long long cols, long long ld, int box_cols, int box_rows, const char* name) {
  auto enc = get_encode();
  if (!enc) { tg_set_error("%s: cuTensorMapEncodeTiled unavailable", name); return -4; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 4) & 15)) { tg_set_error("%s: TMA alignment (ld=%lld)", name, ld); return -1; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { tg_set_error("%s: cuTensorMapEncodeTiled failed (%d)", name, (int)r); return -4; }
  return 0;
}

KLay klay(int K) {
  KLay L;
  const int rem = K % 32;
  L.nfull = K / 32; L.last_ksteps = 4; L.tail_w = 0; L.tail_ksteps = 0;
  if (rem > 16) { L.nfull += 1; L.last_ksteps = (rem + 7) / 8; }
  else if (rem > 0) { L.tail_w = rem <= 8 ? 8 : 16; L.tail_ksteps = (rem + 7) / 8; }
  const int gmax = L.tail_w > 0 ? 7 : 8;
  L.gsz = L.nfull > 0 ? tg_ceil_div(L.nfull, gmax) : 1;
  L.ngf = L.nfull > 0 ? tg_ceil_div(L.nfull, L.gsz) : 0;
  return L;
}

// the forward kernel holds up to 512 tensor-memory columns for its whole life: ask for 180 KB of shared memory so that no second CTA
// (of this kernel or of a concurrently running tensor-core GEMM, all > 48 KB) is ever placed on the same SM, where it would spin in
// tcgen05.alloc until this kernel ends
size_t fwd_smem(int H, int u, int BT) {
  const size_t need = fwd_bar_off(klay(H), u, BT) + 10 * 8 + 16;
  return need > (size_t)180 * 1024 ? need : (size_t)180 * 1024;
}
size_t bwd_smem(int H, int u, int BT) { return bwd_bar_off(klay(3 * H), u, BT) + 10 * 8 + 16; }

template <typename Kern>
cudaError_t cluster_launch_cfg(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, Kern kern, int nthreads, size_t smem, int ntiles, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(CL, ntiles, 2); cfg->blockDim = dim3(nthreads); cfg->dynamicSmemBytes = smem; cfg->stream = s;
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg->attrs = at; cfg->numAttrs = 1;
  return cudaSuccess;
}

// clusters of 8 CTAs (one CTA per SM: the kernels use > 113 KB of shared memory) the device keeps resident; measured 15 on a B200
int max_resident_clusters() {
  static int cached = -1;
  if (cached >= 0) return cached;
  cudaLaunchConfig_t cfg; cudaLaunchAttribute at[1];
  int n = 0;
  const size_t smem = fwd_smem(300, 40, 16);
  if (cluster_launch_cfg(&cfg, at, gru_fwd_cl_kernel<16, 0>, FwdCfg<16>::NT, smem, 32, nullptr) == cudaSuccess &&
      cudaOccupancyMaxActiveClusters(&n, gru_fwd_cl_kernel<16, 0>, &cfg) == cudaSuccess && n > 0)
    cached = n;
  else { cudaGetLastError(); cached = 15; }
  return cached;
}

template <int BT>
int launch_fwd(const FwdP& p, int ntiles, cudaStream_t s) {
  const size_t smem = fwd_smem(p.H, p.u, BT);
  auto kern = p.H == 300 ? gru_fwd_cl_kernel<BT, 300> : gru_fwd_cl_kernel<BT, 0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_fwd_tf32: smem attr (%zu B): %s", smem, cudaGetErrorString(e)); return -3; }
  kern<<<dim3(CL, ntiles, 2), FwdCfg<BT>::NT, smem, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_fwd_tf32: cluster launch (8,%d,2) smem %zu: %s", ntiles, smem, cudaGetErrorString(e)); return -2; }
  return 0;
}
template <int BT>
int launch_bwd(const CUtensorMap* maps, const BwdP& p, int ntiles, cudaStream_t s) {
  const size_t smem = bwd_smem(p.H, p.u, BT);
  cudaError_t e = cudaFuncSetAttribute(gru_bwd_cl_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_bwd_tf32: smem attr (%zu B): %s", smem, cudaGetErrorString(e)); return -3; }
  gru_bwd_cl_kernel<BT><<<dim3(CL, ntiles, 2), BWD_NT, smem, s>>>(maps[0], maps[1], maps[2], maps[3], p);
  e = cudaGetLastError();
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_bwd_tf32: cluster launch (8,%d,2) smem %zu: %s", ntiles, smem, cudaGetErrorString(e)); return -2; }
  return 0;
}

int cl_flags() {             // TGB200_GRU_UNICAST=1: A/B switch of the operand all-gather (development aid)
  static int v = -1;
  if (v < 0) { const char* e = getenv("TGB200_GRU_UNICAST"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}

size_t ks_smem(int u, int BT) {
  const size_t need = ks_bar_off(u, BT) + 32;
  return need > (size_t)180 * 1024 ? need : (size_t)180 * 1024;     // holds most of tensor memory: one CTA per SM (see fwd_smem)
}
int cl_bwd_variant() {       // TGB200_GRU_BWD=ag: the all-gather backward kernel (default: the K-split / reduce-scatter kernel)
  static int v = -1;
  if (v < 0) { const char* e = getenv("TGB200_GRU_BWD"); v = (e && e[0] == 'a') ? 1 : 0; }
  return v;
}
template <int BT>
int launch_bwd_ks(const BwdKsP& p, int ntiles, cudaStream_t s) {
  const size_t smem = ks_smem(p.u, BT);
  auto kern = p.H == 300 ? gru_bwd_ks_kernel<BT, 300> : gru_bwd_ks_kernel<BT, 0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_bwd_tf32: smem attr (%zu B): %s", smem, cudaGetErrorString(e)); return -3; }
  kern<<<dim3(CL, ntiles, 2), BWD_NT, smem, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) { tg_set_error("tg_gru_layer_bwd_tf32: cluster launch (8,%d,2) smem %zu: %s", ntiles, smem, cudaGetErrorString(e)); return -2; }
  return 0;
}

}  // namespace

// ---- plan shared with gru_tc.cu (which owns the C-ABI entries and falls back to its L2-stepped kernels outside this range)
struct TgGruClPlan { int u, bt_f, bt_b, ntiles_f, ntiles_b; };

bool tg_gru_cl_plan(int B, int H, TgGruClPlan* pl) {
  if (H < 8 || H > 320 || (H & 3)) return false;
  pl->u = ((tg_ceil_div(H, CL) + 7) / 8) * 8;          // units per CTA: multiple of 8 keeps every operand chunk 1024-byte aligned
  if (3 * pl->u > 128) return false;
  const size_t cap = (size_t)tg_max_smem_optin();
  const int room = max_resident_clusters();
  // batch tile: the smallest that runs all 2 * ceil(B / BT) chains as ONE wave of resident clusters (fewer clips per CTA = shorter step),
  // else the largest that fits shared memory - more chains than resident clusters simply run as further waves
  const int cf[4] = {16, 32, 48, 64};
  int bt = 0;
  for (int i = 0; i < 4; ++i)
    if (fwd_smem(H, pl->u, cf[i]) <= cap) { bt = cf[i]; if (2 * tg_ceil_div(B, cf[i]) <= room) break; }
  if (!bt) return false;
  pl->bt_f = bt; pl->ntiles_f = tg_ceil_div(B, bt);
  const int cb[2] = {16, 24};
  bt = 0;
  for (int i = 0; i < 2; ++i)
    if (bwd_smem(H, pl->u, cb[i]) <= cap) { bt = cb[i]; if (2 * tg_ceil_div(B, cb[i]) <= room) break; }
  if (!bt) return false;
  pl->bt_b = bt; pl->ntiles_b = tg_ceil_div(B, bt);
  return true;
}

// floats of exchange scratch (the double-buffered operand-tile images) one launch over B clips may touch
size_t tg_gru_cl_xchg_floats(int B, int H) {
  TgGruClPlan pl;
  if (!tg_gru_cl_plan(B, H, &pl)) return 0;
  const KLay Lf = klay(H), Lb = klay(3 * H);
  const size_t f = (size_t)2 * pl.ntiles_f * 2 * ((size_t)Lf.nfull * pl.bt_f * 32 + (size_t)pl.bt_f * Lf.tail_w);
  const size_t b = (size_t)2 * pl.ntiles_b * 2 * ((size_t)Lb.nfull * pl.bt_b * 32 + (size_t)pl.bt_b * Lb.tail_w);
  return f > b ? f : b;
}

int tg_gru_cl_fwd(const float* gi, const float* whh_f, const float* whh_r, const float* bhh_f, const float* bhh_r, float* out, float* saved,
                  long long saved_qstride, const float* mask, float* drop, float* xchg, int B, int T, int H, long long* trace, cudaStream_t s) {
  TgGruClPlan pl;
  if (!tg_gru_cl_plan(B, H, &pl)) { tg_set_error("tg_gru_layer_fwd_tf32: no cluster plan for B=%d H=%d", B, H); return -1; }
  if (reinterpret_cast<uintptr_t>(xchg) & 127) { tg_set_error("tg_gru_layer_fwd_tf32: exchange scratch must be 128-byte aligned"); return -1; }
  if ((reinterpret_cast<uintptr_t>(whh_f) | reinterpret_cast<uintptr_t>(whh_r)) & 15) { tg_set_error("tg_gru_layer_fwd_tf32: W_hh must be 16-byte aligned"); return -1; }
  FwdP p;
  p.L = klay(H);
  p.gi = gi; p.whh0 = whh_f; p.whh1 = whh_r; p.bhh0 = bhh_f; p.bhh1 = bhh_r; p.out = out; p.saved = saved; p.saved_qstride = saved_qstride; p.xchg = xchg;
  p.mask = (mask && drop) ? mask : nullptr; p.drop = (mask && drop) ? drop : nullptr;
  if ((reinterpret_cast<uintptr_t>(p.mask) | reinterpret_cast<uintptr_t>(p.drop)) & 15) { tg_set_error("tg_gru_layer_fwd_tf32: mask / drop must be 16-byte aligned"); return -1; }
  p.B = B; p.T = T; p.H = H; p.u = pl.u; p.flags = cl_flags();
  p.trace = trace;
  if (pl.bt_f == 16) return launch_fwd<16>(p, pl.ntiles_f, s);
  if (pl.bt_f == 32) return launch_fwd<32>(p, pl.ntiles_f, s);
  if (pl.bt_f == 48) return launch_fwd<48>(p, pl.ntiles_f, s);
  return launch_fwd<64>(p, pl.ntiles_f, s);
}

int tg_gru_cl_bwd(const float* dout, const float* out, const float* saved, long long saved_qstride, const float* whhT_f, const float* whhT_r,
                  float* dgi, float* dgh, float* xchg, int B, int T, int H, long long* trace, cudaStream_t s) {
  TgGruClPlan pl;
  if (!tg_gru_cl_plan(B, H, &pl)) { tg_set_error("tg_gru_layer_bwd_tf32: no cluster plan for B=%d H=%d", B, H); return -1; }
  if (reinterpret_cast<uintptr_t>(xchg) & 127) { tg_set_error("tg_gru_layer_bwd_tf32: exchange scratch must be 128-byte aligned"); return -1; }
  if (!cl_bwd_variant() && ((reinterpret_cast<uintptr_t>(whhT_f) | reinterpret_cast<uintptr_t>(whhT_r)) & 15) == 0 && ks_tmem_cols(H, pl.u) <= 512) {
    // K-split kernel: batch tile 16 when all 2 * ceil(B/16) chains fit one wave of resident clusters, else 32
    BwdKsP k;
    k.dout = dout; k.out = out; k.saved = saved; k.saved_qstride = saved_qstride; k.whhT0 = whhT_f; k.whhT1 = whhT_r; k.dgi = dgi; k.dgh = dgh;
    k.B = B; k.T = T; k.H = H; k.u = pl.u; k.trace = trace;
    if (2 * tg_ceil_div(B, 16) <= max_resident_clusters()) return launch_bwd_ks<16>(k, tg_ceil_div(B, 16), s);
    return launch_bwd_ks<32>(k, tg_ceil_div(B, 32), s);
  }
  BwdP p;
  p.L = klay(3 * H);
  CUtensorMap maps[4];
  int rc;
  if ((rc = map_2d(&maps[0], whhT_f, H, 3ll * H, 3ll * H, 32, pl.u, "tg_gru_layer_bwd_tf32(W^T)"))) return rc;
  if ((rc = map_2d(&maps[1], whhT_r, H, 3ll * H, 3ll * H, 32, pl.u, "tg_gru_layer_bwd_tf32(W^T)"))) return rc;
  maps[2] = maps[0]; maps[3] = maps[1];
  if (p.L.tail_w > 0) {
    if ((rc = map_2d(&maps[2], whhT_f, H, 3ll * H, 3ll * H, p.L.tail_w, pl.u, "tg_gru_layer_bwd_tf32(W^T tail)"))) return rc;
    if ((rc = map_2d(&maps[3], whhT_r, H, 3ll * H, 3ll * H, p.L.tail_w, pl.u, "tg_gru_layer_bwd_tf32(W^T tail)"))) return rc;
  }
  p.dout = dout; p.out = out; p.saved = saved; p.saved_qstride = saved_qstride; p.dgi = dgi; p.dgh = dgh; p.xchg = xchg;
  p.B = B; p.T = T; p.H = H; p.u = pl.u; p.flags = cl_flags();
  p.trace = trace;
  if (pl.bt_b == 16) return launch_bwd<16>(maps, p, pl.ntiles_b, s);
  return launch_bwd<24>(maps, p, pl.ntiles_b, s);
}

// how many 8-CTA clusters of the recurrence kernels the device keeps resident at once (cudaOccupancyMaxActiveClusters); measured: 15
extern "C" int tg_debug_gru_cluster_occupancy(int H, int BT) {
  (void)H; (void)BT;
  return max_resident_clusters();
}

// the (forward, backward) batch tiles the plan picks for B clips: packed as bt_f * 1000 + bt_b, or -1 without a cluster plan
extern "C" int tg_debug_gru_cluster_tiles(int B, int H) {
  TgGruClPlan pl;
  if (!tg_gru_cl_plan(B, H, &pl)) return -1;
  return pl.bt_f * 1000 + pl.bt_b;
}
