// Two-tap (causal dilated, k = 2) convolution of the TextEncoderTCN and its anti-causal data gradient as ONE-accumulator tcgen05 GEMMs
// (tcn.py:19-31; the `taps == 2` case of tg_gemm_tf32, dispatched from gemm_tf32.cu when the shape qualifies).
//
//   C[m, n] = epi( sum_k A[m + shift0, k] * Bw[n, k]  (only where the shifted row stays inside its clip)  +  sum_k A[m, k] * Bw[N + n, k] )
//
// The general kernel (gemm_tf32.cu) keeps the shifted tap in a SECOND TMEM accumulator and masks it per row in the epilogue, which caps
// the tile at 128 x 128 (2 x 2 x 128 columns for two resident CTAs).  For the TCN's N = 300 that meant 3 N-tiles per M-tile (22 % padding,
// every A tile fetched three times) and 306 CTAs on 296 resident slots: a 10-CTA tail wave.  Measured 51 us per layer at 92 TFLOP/s
// (profiles/r02_kernels_by_shape.txt), L2 -> SM operand traffic 188 MB per launch.
//
// Here the A operand is a 3-D TMA map {K, T rows, clips} and a tile is a GROUP OF WHOLE CLIPS (floor(128 / T) of them: 3 x 34 = 102 rows):
// the shifted tap is loaded at row coordinate `shift0`, so the rows that fall outside a clip are ZERO-FILLED BY THE TMA UNIT, clip by clip -
// both taps can accumulate into the same TMEM columns.  One accumulator of up to 304 columns (two MMAs: N = 160 + 144) then covers the
// whole N = 300 in one CTA: every A tile is read once, 128 CTAs = one wave on 148 SMs, 123 MB of operand traffic.  Rows 102..127 of the
// MMA tile are stale shared memory; their accumulator rows are never read.
// Small problems (the data gradient of one pass: 128 clips -> 43 tiles) split N over two CTAs instead (2 x 160 columns).
#include <cudaTypedefs.h>
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int BM = 128, BKF = 32, NSTAGE = 4;
constexpr int A_STAGE_BYTES = BM * 128;

__device__ __forceinline__ void stamp(long long* trace, int slot) {
  if (trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)::"memory");
    trace[slot] = (long long)t;
  }
}

struct TcnP {
  long long* trace;
  float* C; int ldc;
  int M, N, K, shift0;
  int tile_rows, cpt;                 // rows / clips per tile
  int npc;                            // accumulator columns per CTA (multiple of 16, <= 320)
  int n1, n2;                         // MMA N of the two column parts (n2 may be 0)
  uint32_t idesc1, idesc2, tmem_cols;
  int bhalf;                          // rows per B TMA box (two boxes per stage): npc / 2
  uint32_t stage_bytes, tx_bytes;
  const float* escale; const float* bias; int act1; float slope1;
  const float* mask; int ldmask; const float* residual; int ldres; int act2; int accumulate;
};

constexpr int EPI_WARPS = 16, NTHREADS = 64 + 32 * EPI_WARPS;

__global__ void __launch_bounds__(NTHREADS, 1) gemm_tcn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                          const TcnP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * p.stage_bytes);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tmem_full_bar = empty_bar + NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int clip0 = blockIdx.x * p.cpt;
  const int m0 = blockIdx.x * p.tile_rows;
  const int m_end = min(p.M, m0 + p.tile_rows);
  const int n0 = blockIdx.y * p.npc;
  const int nkb = (p.K + BKF - 1) / BKF;
  const int iters = 2 * nkb;

  if (threadIdx.x == 0) stamp(p.trace, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (threadIdx.x == 0) stamp(p.trace, 1);

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int tap = it / nkb, kb = it - tap * nkb;
        uint8_t* sa = smem + s * p.stage_bytes;
        uint8_t* sb = sa + A_STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], p.tx_bytes);
        // tap 0 reads rows t + shift0 of every clip of the group: rows outside [0, T) arrive as zeros
        tma_load_3d(sa, &tmA, &full_bar[s], kb * BKF, tap == 0 ? p.shift0 : 0, clip0);
        tma_load_2d(sb, &tmB, &full_bar[s], kb * BKF, tap * p.N + n0);
        tma_load_2d(sb + p.bhalf * 128, &tmB, &full_bar[s], kb * BKF, tap * p.N + n0 + p.bhalf);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (it == 0) stamp(p.trace, 2);
        if (it == nkb) stamp(p.trace, 7);
        const uint32_t sa = smem_u32(smem + s * p.stage_bytes);
        const uint32_t sb = sa + A_STAGE_BYTES;
        constexpr uint32_t hi = desc_hi(1024, 2);                // SWIZZLE_128B; the 4 K-steps of a stage walk the 128-byte row: +32 B each
        mma_tf32_seq<4, 2>(tmem_base, desc_lo(sa, 16), hi, desc_lo(sb, 16), hi, p.idesc1, it > 0 ? 1u : 0u);
        if (p.n2 > 0)
          mma_tf32_seq<4, 2>(tmem_base + (uint32_t)p.n1, desc_lo(sa, 16), hi, desc_lo(sb + (uint32_t)p.n1 * 128u, 16), hi, p.idesc2, it > 0 ? 1u : 0u);
        tc_commit(&empty_bar[s]);
      }
      tc_commit(tmem_full_bar);
      stamp(p.trace, 3);
    }
  } else {
    // ---- epilogue: 16 warps (gemm_tf32.cu has 8; here nothing else is resident on the SM to hide the epilogue behind, and a warp's chunk
    // is one round trip of mask / residual loads: measured 10.8 us of a 26 us CTA with 8 warps).  Warps w, w + 4, w + 8, w + 12 share a TMEM
    // lane quarter and interleave its 32-column chunks.
    const int q = warp & 3, half = (warp - 2) >> 2;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) stamp(p.trace, 4);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool vec_ok = (p.N & 3) == 0 && (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 &&
                        (!p.mask || ((p.ldmask & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0)) &&
                        (!p.residual || ((p.ldres & 3) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0));
    constexpr int SP = 36;
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * SP);
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    const float s1 = p.act1 == 0 ? 1.f : (p.act1 == 1 ? 0.f : p.slope1);
    const float s2 = p.act2 == 0 ? 1.f : 0.f;
    const int row0 = m0 + q * 32 + rsub;             // this lane's first row; it handles rows row0 + 4*i
    const int rows_left = m_end - row0;              // row i is valid iff 4*i < rows_left
#pragma unroll 1
    for (int c0 = half * 32; c0 < p.npc; c0 += 32 * (EPI_WARPS / 4)) {
      const int nb = n0 + c0;
      if (nb >= p.N) break;
      float v[32];
      tmem_ld32(lane_addr + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4) *reinterpret_cast<float4*>(stg + lane * SP + j4) = make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
      __syncwarp();
      const int n = nb + c4;                      // this lane's 4 columns
      if (vec_ok) {
        if (n < p.N) {                            // N % 4 == 0: the lane's four columns are all valid or all invalid
          float4 es = make_float4(1.f, 1.f, 1.f, 1.f), bs = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.escale) es = __ldg(reinterpret_cast<const float4*>(p.escale + n));
          if (p.bias) bs = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          float* crow = p.C + (long long)row0 * p.ldc + n;
          const float* mrow = p.mask ? p.mask + (long long)row0 * p.ldmask + n : nullptr;
          const float* rrow = p.residual ? p.residual + (long long)row0 * p.ldres + n : nullptr;
          const float* srow = stg + rsub * SP + c4;
          const long long cstep = 4ll * p.ldc, mstep = 4ll * p.ldmask, rstep = 4ll * p.ldres;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (4 * i < rows_left) {
              float4 x = *reinterpret_cast<const float4*>(srow + i * 4 * SP);
              x.x = fmaf(x.x, es.x, bs.x); x.y = fmaf(x.y, es.y, bs.y); x.z = fmaf(x.z, es.z, bs.z); x.w = fmaf(x.w, es.w, bs.w);
              x.x = fmaxf(x.x, x.x * s1); x.y = fmaxf(x.y, x.y * s1); x.z = fmaxf(x.z, x.z * s1); x.w = fmaxf(x.w, x.w * s1);
              if (mrow) {
                const float4 t4 = *reinterpret_cast<const float4*>(mrow + i * mstep);
                x.x *= t4.x; x.y *= t4.y; x.z *= t4.z; x.w *= t4.w;
              }
              if (rrow) {
                const float4 t4 = *reinterpret_cast<const float4*>(rrow + i * rstep);
                x.x += t4.x; x.y += t4.y; x.z += t4.z; x.w += t4.w;
              }
              x.x = fmaxf(x.x, x.x * s2); x.y = fmaxf(x.y, x.y * s2); x.z = fmaxf(x.z, x.z * s2); x.w = fmaxf(x.w, x.w * s2);
              if (p.accumulate) {
                const float4 t4 = *reinterpret_cast<const float4*>(crow + i * cstep);
                x.x += t4.x; x.y += t4.y; x.z += t4.z; x.w += t4.w;
              }
              *reinterpret_cast<float4*>(crow + i * cstep) = x;
            }
          }
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          if (4 * i >= rows_left) break;
          const int mr = row0 + 4 * i;
#pragma unroll 1
          for (int e = 0; e < 4; ++e) {
            if (n + e >= p.N) break;
            float y = stg[(rsub + 4 * i) * SP + c4 + e];
            if (p.escale) y *= __ldg(p.escale + n + e);
            if (p.bias) y += __ldg(p.bias + n + e);
            y = fmaxf(y, y * s1);
            if (p.mask) y *= p.mask[(long long)mr * p.ldmask + n + e];
            if (p.residual) y += p.residual[(long long)mr * p.ldres + n + e];
            y = fmaxf(y, y * s2);
            float* dst = p.C + (long long)mr * p.ldc + n + e;
            if (p.accumulate) y += *dst;
            *dst = y;
          }
        }
      }
      __syncwarp();
    }
  }
  if (threadIdx.x == 64) stamp(p.trace, 5);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  if (threadIdx.x == 32) stamp(p.trace, 6);
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode_tcn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

}  // namespace

// Does the clip-group kernel take this two-tap problem?  (flat A of whole clips, T <= 128, N <= 2 x 304 columns, aligned operands)
long long* tg_gemm_trace_ptr();

bool tg_gemm_tcn_applies(const tg_gemm_tf32_t& g) {
  if (g.taps != 2 || g.clip_rows != 0 || g.T <= 0 || g.T > BM || g.M % g.T != 0) return false;
  if (g.a_rows != 0 && g.a_rows != g.M) return false;
  if (g.N > 608 || g.N < 16 || abs(g.shift0) >= g.T) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.Bw) & 15) || (g.lda & 3) || (g.ldb & 3)) return false;
  return true;
}

int tg_gemm_tcn_launch(const tg_gemm_tf32_t& g, cudaStream_t s) {
  auto enc = get_encode_tcn();
  if (!enc) { tg_set_error("tg_gemm_tf32(tcn): cuTensorMapEncodeTiled unavailable"); return -4; }
  const int T = g.T, clips = g.M / T;
  const int cpt = BM / T, tile_rows = cpt * T;
  const int tiles = tg_ceil_div(clips, cpt);
  // all of N in one CTA when that fills the machine or N does not split evenly anyway; otherwise two column halves
  const int np_full = tg_ceil_div(g.N, 16) * 16;
  int nsplit = (np_full <= 304 && tiles * 2 > tg_num_sms()) ? 1 : 2;
  if (np_full <= 160 && tiles * 2 > tg_num_sms()) nsplit = 1;
  int npc = tg_ceil_div(tg_ceil_div(g.N, nsplit), 16) * 16;
  if (npc > 304) { tg_set_error("tg_gemm_tf32(tcn): N = %d too wide", g.N); return -1; }
  TcnP p;
  p.trace = tg_gemm_trace_ptr();
  p.C = g.C; p.ldc = g.ldc; p.M = g.M; p.N = g.N; p.K = g.K; p.shift0 = g.shift0;
  p.tile_rows = tile_rows; p.cpt = cpt; p.npc = npc;
  if (npc <= 256) { p.n1 = npc; p.n2 = 0; }
  else { p.n1 = tg_ceil_div(npc / 2, 16) * 16; p.n2 = npc - p.n1; }
  p.idesc1 = idesc_tf32(BM, p.n1, 0, 0);
  p.idesc2 = p.n2 > 0 ? idesc_tf32(BM, p.n2, 0, 0) : 0;
  p.tmem_cols = npc <= 32 ? 32 : npc <= 64 ? 64 : npc <= 128 ? 128 : npc <= 256 ? 256 : 512;
  p.bhalf = npc / 2;                                     // npc % 16 == 0: a multiple of the 8-row swizzle atom
  p.stage_bytes = (uint32_t)(A_STAGE_BYTES + npc * 128);
  p.tx_bytes = (uint32_t)(128 * T * cpt + npc * 128);
  p.escale = g.escale; p.bias = g.bias; p.act1 = g.act1; p.slope1 = g.slope1; p.mask = g.mask; p.ldmask = g.ldmask;
  p.residual = g.residual; p.ldres = g.ldres; p.act2 = g.act2; p.accumulate = g.accumulate;

  CUtensorMap ta, tb;
  {
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)T, (cuuint64_t)clips};
    cuuint64_t strides[2] = {(cuuint64_t)g.lda * 4, (cuuint64_t)T * g.lda * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)T, (cuuint32_t)cpt};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(g.A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tg_set_error("tg_gemm_tf32(tcn A): cuTensorMapEncodeTiled failed (%d) K=%d T=%d clips=%d lda=%d", (int)r, g.K, T, clips, g.lda); return -4; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)g.K, (cuuint64_t)(2ll * g.N)};
    cuuint64_t strides[1] = {(cuuint64_t)g.ldb * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)p.bhalf};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g.Bw), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tg_set_error("tg_gemm_tf32(tcn B): cuTensorMapEncodeTiled failed (%d) K=%d N=%d ldb=%d", (int)r, g.K, g.N, g.ldb); return -4; }
  }
  const size_t smem = (size_t)NSTAGE * p.stage_bytes + (2 * NSTAGE + 1) * 8 + 16 + 1024;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { tg_set_error("tg_gemm_tf32(tcn): smem attr (%zu bytes): %s", smem, cudaGetErrorString(e)); return -3; }
    attr_smem = smem;
  }
  dim3 grid((unsigned)tiles, (unsigned)tg_ceil_div(g.N, npc));
  gemm_tcn_kernel<<<grid, NTHREADS, smem, s>>>(ta, tb, p);
  TG_CHECK_LAUNCH("tg_gemm_tf32(tcn)");
  return 0;
}
