// Tensor-core GEMM for the "fast" mode: tcgen05.mma kind::tf32 (fp32 operands read straight from the fp32 activations
// and weights by TMA - no conversion pass, 10-bit-mantissa products, fp32 accumulation in TMEM).
//
//   C[m, n] = epi( sum_tap sum_k A[m + shift_tap, k] * B[tap*N + n, k] )        A, B row-major (K contiguous)
//
// 128 x BN tile per CTA, BK = 32 floats (one 128-byte swizzle atom), NSTAGE-deep TMA->MMA mbarrier pipeline,
// warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-9 = epilogue
// (tcgen05.ld 32x32b, one TMEM lane = one output row per thread).
// Two-tap mode (the TCN's causal dilated convolution, tcn.py:19-31, and its anti-causal data gradient): the shifted
// tap accumulates in a second TMEM accumulator and is masked per row in the epilogue (t + shift outside the clip),
// so the shifted operand is a plain 2-D TMA load with a row offset - no im2col, no padded copy, no chomp.
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int BM = 128, BKF = 32;            // BKF floats = 128 bytes
constexpr int A_STAGE_BYTES = BM * 128;

// optional phase stamps of CTA (0,0) (tg_debug_gemm_trace): trace[slot] = %globaltimer
long long* g_trace = nullptr;
__device__ __forceinline__ void stamp(long long* trace, int slot) {
  if (trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)::"memory");
    trace[slot] = (long long)t;
  }
}

struct GemmP {
  long long* trace;
  float* C; int ldc;
  int M, N, K, taps, shift0, T;
  int clip_rows, tiles_per_clip;     // clip mode (clip_rows > 0): tile = 128 rows of one clip; C row = clip*clip_rows + t
  int acc_taps;                      // > 1 (clip mode, taps == 1): tap j reads A rows t - j (rows outside the clip: TMA zero fill) against
                                     // Bw rows j*N + n, every tap into the ONE accumulator - the data gradient of a strided convolution
  long long c_clip_pitch, c_clip_len;  // acc_taps mode: clip c of C starts at C + c*c_clip_pitch and holds c_clip_len floats (a ragged last row)
  int ksplit;                        // > 1: blockIdx.z owns a contiguous share of the K blocks and ADDS its partial tile to a zeroed C with
                                     // red.global.add (long-K, few-tile problems: the GRU data gradient [4352 x 600] x K 1800 ran 57
                                     // dependent stages on 170 CTAs, 72 us for 31 us worth of operand traffic)
  const float* escale; const float* bias; int act1; float slope1;
  const float* mask; int ldmask; const float* residual; int ldres; int act2; int accumulate;
};

// MH = 2 (flat single-tap problems with wide outputs): the CTA owns 256 rows as two 128-row accumulators fed by the same B tile, so a
// 256 x 240 x 32 block costs (256 + 240) x 128 bytes of operand fetch where two 128-row CTAs pay (256 + 480) x 128 - these GEMMs run at
// the L2 -> SM fabric's ~10 TB/s, not at the tensor pipe's rate.  Both accumulators fill tensor memory (2 x 256 columns), one CTA per
// SM, 16 epilogue warps; the epilogue of one SM overlaps the main loops of the others.
template <int BN, int NSTAGE, int MH>
__global__ void __launch_bounds__(MH == 2 ? 576 : 320, MH == 2 ? 1 : 2) gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                                        const __grid_constant__ CUtensorMap tmB, const GemmP p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int B_STAGE_BYTES = BN * 128;
  constexpr int STAGE_BYTES = MH * A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;          // TMEM column offset of the second 128-row accumulator (MH == 2)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tmem_full_bar = empty_bar + NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // flat mode: tile rows m0.. of the [M,K] matrix (A map = {K, a_rows, 1}); clip mode: rows t0.. of clip `clip`
  // (A map = {K, clip_rows, clips} with its own row / clip pitches, e.g. overlapping convolution windows)
  const int clip = p.clip_rows > 0 ? blockIdx.x / p.tiles_per_clip : 0;
  const int t0 = p.clip_rows > 0 ? (blockIdx.x - clip * p.tiles_per_clip) * BM : blockIdx.x * (BM * MH);
  const int m0 = clip * p.clip_rows + t0;                    // first output row of the tile
  const int m_end = p.clip_rows > 0 ? clip * p.clip_rows + p.clip_rows : p.M;
  const int n0 = blockIdx.y * BN;
  const int nkb_all = (p.K + BKF - 1) / BKF;
  // split-K (taps == 1 only): this CTA's K blocks are [kb_lo, kb_lo + nkb)
  const int kb_per = p.ksplit > 1 ? (nkb_all + p.ksplit - 1) / p.ksplit : nkb_all;
  const int kb_lo = p.ksplit > 1 ? (int)blockIdx.z * kb_per : 0;
  const int nkb = p.ksplit > 1 ? max(0, min(kb_per, nkb_all - kb_lo)) : nkb_all;
  const int ntap = p.acc_taps > 1 ? p.acc_taps : p.taps;
  const int iters = nkb * ntap;
  constexpr uint32_t TMEM_COLS_1 = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  constexpr uint32_t TMEM_COLS_2 = 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  const uint32_t tmem_cols = MH == 2 ? 2 * ACC_STRIDE : (p.taps == 2 ? TMEM_COLS_2 : TMEM_COLS_1);

  if (threadIdx.x == 0) stamp(p.trace, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // warp-uniform for ptxas (feeds uniform registers)
  if (threadIdx.x == 0) stamp(p.trace, 1);

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int tap = it / nkb, kb = it - tap * nkb;
        const int shift = p.acc_taps > 1 ? -tap : ((p.taps == 2 && tap == 0) ? p.shift0 : 0);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + MH * A_STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_3d(sa, &tmA, &full_bar[s], (kb_lo + kb) * BKF, t0 + shift, clip);
        if (MH == 2) tma_load_3d(sa + A_STAGE_BYTES, &tmA, &full_bar[s], (kb_lo + kb) * BKF, t0 + BM, clip);      // rows past the matrix: zero fill
        tma_load_2d(sb, &tmB, &full_bar[s], (kb_lo + kb) * BKF, tap * p.N + n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = idesc_tf32(BM, BN, 0, 0);
      for (int it = 0; it < iters; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (it == 0) stamp(p.trace, 2);
        const int tap = it / nkb, kb = it - tap * nkb;
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + MH * A_STAGE_BYTES;
        const uint32_t dcol = tmem_base + (uint32_t)(p.acc_taps > 1 ? 0 : tap * BN);
        const uint32_t accum = (p.acc_taps > 1 ? it > 0 : kb > 0) ? 1u : 0u;
        constexpr uint32_t hi = desc_hi(1024, 2);                // SWIZZLE_128B; the 4 K-steps of a stage walk the 128-byte row: +32 B each
        mma_tf32_seq<4, 2>(dcol, desc_lo(sa, 16), hi, desc_lo(sb, 16), hi, idesc, accum);
        if (MH == 2) mma_tf32_seq<4, 2>(dcol + ACC_STRIDE, desc_lo(sa + A_STAGE_BYTES, 16), hi, desc_lo(sb, 16), hi, idesc, accum);
        tc_commit(&empty_bar[s]);
      }
      tc_commit(tmem_full_bar);
      stamp(p.trace, 3);
    }
  } else {
    // ---- epilogue: 8 warps (16 with two accumulators).  Warp w may only touch TMEM lanes 32*(w%4) .. +31, so the warps of one lane
    // quarter split its (accumulator, 32-column chunk) items round-robin.  One epilogue warp per scheduler cannot hide its own instruction latency, so
    // the per-element work is kept branch-free: act(v) = max(v, v*s) (s = 1 none, 0 relu, slope leaky), every option folded
    // into per-chunk constants, row addresses advanced by pointer increments.
    const int q = warp & 3, grp = (warp - 2) >> 2;
    constexpr int NGRP = MH == 2 ? 4 : 2, NCHUNK = (BN + 31) / 32;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) stamp(p.trace, 4);
    const int m = m0 + q * 32 + lane;          // (two-tap mode is MH == 1)
    bool tap0_ok = true;
    if (p.taps == 2) {
      const int t = m % p.T;
      const int ts = t + p.shift0;
      tap0_ok = ts >= 0 && ts < p.T;
    }
    const uint32_t lane_addr0 = tmem_base + ((uint32_t)(q * 32) << 16);
    // (N need not be a multiple of 4: a lane whose four columns straddle N takes the scalar path below - the 150-wide head with a 152-float
    // row pitch stores 37 vectors and one half-quad per row instead of 150 scalars)
    const bool vec_ok = (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 &&
                        (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) && (!p.escale || (reinterpret_cast<uintptr_t>(p.escale) & 15) == 0) &&
                        (!p.mask || ((p.ldmask & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0)) &&
                        (!p.residual || ((p.ldres & 3) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0));
    // All TMA loads have landed and every MMA has retired once tmem_full_bar fires, so the pipeline stages are free:
    // each epilogue warp transposes its 32x32 chunk through a padded staging tile there (row pitch 36 floats: the
    // float4 writes of 8 consecutive rows and the float4 reads of one row are both bank-conflict free), after which a
    // warp instruction touches 4 rows x 128 contiguous bytes of C / mask / residual instead of 32 rows x 16 bytes.
    constexpr int SP = 36;
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * SP);
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    const float s1 = p.act1 == 0 ? 1.f : (p.act1 == 1 ? 0.f : p.slope1);
    const float s2 = p.act2 == 0 ? 1.f : 0.f;
#pragma unroll 1
    for (int item = grp; item < MH * NCHUNK; item += NGRP) {
      const int h = MH == 2 ? item / NCHUNK : 0;
      const int c0 = (item - h * NCHUNK) * 32;
      const int nb = n0 + c0;
      if (nb >= p.N) continue;
      const uint32_t lane_addr = lane_addr0 + (uint32_t)(h * ACC_STRIDE);
      const int row0 = m0 + h * BM + q * 32 + rsub;    // this lane's first row; it handles rows row0 + 4*i
      const int rows_left = m_end - row0;              // row i is valid iff 4*i < rows_left
      float v[32];
      if (p.taps == 2) {
        float v0[32];
        tmem_ld32(lane_addr + (uint32_t)c0, v0);
        tmem_ld32(lane_addr + (uint32_t)(BN + c0), v);
        tmem_ld_wait();
        if (tap0_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v0[j];
        }
      } else {
        tmem_ld32(lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
      }
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4) *reinterpret_cast<float4*>(stg + lane * SP + j4) = make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
      __syncwarp();
      const int n = nb + c4;                      // this lane's 4 columns
      if (BN % 32 != 0 && c0 + c4 >= BN) {        // BN = 240: the last 32-column chunk is half a tile wide
      } else if (vec_ok && (n + 3 < p.N || n >= p.N)) {
        if (n < p.N) {                            // the lane's four columns are all valid
          float4 es = make_float4(1.f, 1.f, 1.f, 1.f), bs = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.escale) es = __ldg(reinterpret_cast<const float4*>(p.escale + n));
          if (p.bias && (p.ksplit <= 1 || blockIdx.z == 0)) bs = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          float* crow = p.C + (long long)row0 * p.ldc + n;
          int cmax = 8;                                          // rows i < cmax may be stored
          if (p.acc_taps > 1) {
            // ragged clips: row t of clip `clip` lives at C + clip*c_clip_pitch + t*ldc and only the clip's first c_clip_len floats exist
            const int tl = t0 + q * 32 + rsub;
            crow = p.C + (long long)clip * p.c_clip_pitch + (long long)tl * p.ldc + n;
            const long long left = p.c_clip_len - ((long long)tl * p.ldc + n);          // > 0 and a multiple of 4 when the vector is valid
            cmax = left <= 0 ? 0 : (int)min((long long)8, (left - 4) / (4ll * p.ldc) + 1);
          }
          const float* mrow = p.mask ? p.mask + (long long)row0 * p.ldmask + n : nullptr;
          const float* rrow = p.residual ? p.residual + (long long)row0 * p.ldres + n : nullptr;
          const float* srow = stg + rsub * SP + c4;
          const long long cstep = 4ll * p.ldc, mstep = 4ll * p.ldmask, rstep = 4ll * p.ldres;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (4 * i < rows_left && i < cmax) {
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
              if (p.accumulate && p.ksplit <= 1) {              // split-K: the red.add below IS the accumulation
                const float4 t4 = *reinterpret_cast<const float4*>(crow + i * cstep);
                x.x += t4.x; x.y += t4.y; x.z += t4.z; x.w += t4.w;
              }
              if (p.ksplit > 1)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + i * cstep), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
              else
                *reinterpret_cast<float4*>(crow + i * cstep) = x;
            }
          }
        }
      } else {
        // unaligned / N % 4 != 0 (e.g. the 150-wide head): scalar path, not unrolled
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
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
  if (threadIdx.x == 32) stamp(p.trace, 6);
}

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

// row-major fp32 matrix [rows, cols], pitch ld floats; box = 32 floats x box_rows, 128-byte swizzle, zero OOB fill
int make_map_2d(CUtensorMap* m, const float* base, long long rows, long long cols, long long ld, int box_rows, const char* name) {
  auto enc = get_encode();
  if (!enc) { tg_set_error("%s: cuTensorMapEncodeTiled unavailable", name); return -4; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 4) & 15)) { tg_set_error("%s: TMA needs 16-byte aligned base and pitch (ld=%lld)", name, ld); return -1; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { tg_set_error("%s: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", name, (int)r, rows, cols, ld); return -4; }
  return 0;
}

// A operand: [clips][rows][cols] with row pitch ld and clip pitch cp (floats); box = 32 floats x 128 rows x 1 clip
int make_map_a(CUtensorMap* m, const float* base, long long clips, long long rows, long long cols, long long ld, long long cp, const char* name) {
  auto enc = get_encode();
  if (!enc) { tg_set_error("%s: cuTensorMapEncodeTiled unavailable", name); return -4; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 4) & 15) || ((cp * 4) & 15)) {
    tg_set_error("%s: TMA needs 16-byte aligned base and pitches (ld=%lld clip pitch=%lld)", name, ld, cp); return -1;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)clips};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)cp * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tg_set_error("%s: cuTensorMapEncodeTiled(3d) failed (%d) clips=%lld rows=%lld cols=%lld ld=%lld cp=%lld", name, (int)r, clips, rows, cols, ld, cp);
    return -4;
  }
  return 0;
}

// accumulating-taps mode (tg_conv_dgrad_tf32): A rows per clip differ from the output rows per clip, C clips are ragged
struct TapExt { int acc_taps; long long a_rows, c_clip_pitch, c_clip_len; };

template <int BN, int NSTAGE, int MH = 1>
int launch(const tg_gemm_tf32_t& g, cudaStream_t s, const TapExt* x = nullptr) {
  CUtensorMap ta, tb;
  const bool clipm = g.clip_rows > 0;
  const long long clips = clipm ? g.M / g.clip_rows : 1;
  const long long arows = x ? x->a_rows : (clipm ? g.clip_rows : g.a_rows);
  const long long cp = clipm ? g.a_clip_pitch : arows * (long long)g.lda;
  int rc = make_map_a(&ta, g.A, clips, arows, g.K, g.lda, cp > 0 ? cp : 4, "tg_gemm_tf32(A)");
  if (rc) return rc;
  rc = make_map_2d(&tb, g.Bw, (long long)(x ? x->acc_taps : g.taps) * g.N, g.K, g.ldb, BN, "tg_gemm_tf32(B)");
  if (rc) return rc;
  GemmP p;
  p.trace = g_trace;
  p.C = g.C; p.ldc = g.ldc; p.M = g.M; p.N = g.N; p.K = g.K; p.taps = g.taps; p.shift0 = g.shift0; p.T = g.T > 0 ? g.T : 1;
  p.escale = g.escale; p.bias = g.bias; p.act1 = g.act1; p.slope1 = g.slope1; p.mask = g.mask; p.ldmask = g.ldmask;
  p.residual = g.residual; p.ldres = g.ldres; p.act2 = g.act2; p.accumulate = g.accumulate;
  p.clip_rows = clipm ? g.clip_rows : 0;
  p.tiles_per_clip = clipm ? tg_ceil_div(g.clip_rows, BM) : 0;
  p.acc_taps = x ? x->acc_taps : 1;
  p.c_clip_pitch = x ? x->c_clip_pitch : 0;
  p.c_clip_len = x ? x->c_clip_len : 0;
  // split-K: a linear epilogue (scale / bias / mask only), vector-aligned contiguous C, few tiles and a long K loop
  p.ksplit = 1;
  {
    const int tiles = tg_ceil_div(g.M, BM * MH) * tg_ceil_div(g.N, BN), nkb_all = tg_ceil_div(g.K, BKF);
    // (accumulate = 1 with a linear epilogue is split-K without the zero fill: the caller's C - e.g. a buffer zeroed ahead of time on a side
    // stream - receives every partial tile by red.add.  A memset NODE between two kernels of a captured chain costs ~25 us of gaps.)
    const bool linear = g.taps == 1 && !clipm && g.act1 == 0 && g.act2 == 0 && !g.residual;
    const bool vec = (g.N & 3) == 0 && g.ldc == g.N && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 &&
                     (!g.mask || ((g.ldmask & 3) == 0 && (reinterpret_cast<uintptr_t>(g.mask) & 15) == 0));
    static const bool off = getenv("TGB200_NO_SPLITK") != nullptr;
    if (!off && linear && vec && nkb_all >= 24 && tiles < 2 * tg_num_sms()) {
      int ks = (3 * tg_num_sms() + tiles - 1) / tiles;         // aim at ~3 resident CTAs' worth of tiles per SM pair
      if (MH == 2) ks = tg_num_sms() / tiles;                  // 256-row tiles hold a whole SM: size the split for exactly one wave
      if (ks > 4) ks = 4;
      if (ks > nkb_all / 8) ks = nkb_all / 8;
      while (ks > 1 && ((nkb_all + ks - 1) / ks) * (ks - 1) >= nkb_all) --ks;      // every share non-empty
      if (ks > 1) {
        p.ksplit = ks;
        if (!g.accumulate) {
          cudaError_t e = cudaMemsetAsync(g.C, 0, sizeof(float) * (size_t)g.M * g.N, s);
          if (e != cudaSuccess) { tg_set_error("tg_gemm_tf32: memset: %s", cudaGetErrorString(e)); return -2; }
        }
      }
    }
  }
  constexpr size_t smem = (size_t)NSTAGE * (MH * A_STAGE_BYTES + BN * 128) + (2 * NSTAGE + 1) * 8 + 16 + 1024;
  static_assert(MH == 1 || (size_t)NSTAGE * (MH * A_STAGE_BYTES + BN * 128) >= 16 * 32 * 36 * 4, "epilogue staging lives in the pipeline stages");
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<BN, NSTAGE, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { tg_set_error("tg_gemm_tf32: smem attr: %s", cudaGetErrorString(e)); return -3; }
    attr_done = true;
  }
  dim3 grid(clipm ? (unsigned)(clips * p.tiles_per_clip) : tg_ceil_div(g.M, BM * MH), tg_ceil_div(g.N, BN), p.ksplit);
  gemm_tf32_kernel<BN, NSTAGE, MH><<<grid, MH == 2 ? 576 : 320, smem, s>>>(ta, tb, p);
  TG_CHECK_LAUNCH("tg_gemm_tf32");
  return 0;
}

}  // namespace

// Data gradient of a strided, unpadded Conv1d as ceil(k/stride) accumulating taps (header: tg_conv_dgrad_tf32).  With t = stride*q + r:
//   da[b, t, c] = sum_j dy[b, q - j, :] . w[:, c, r + stride*j]        (taps with r + stride*j >= k are zero rows of wd)
// Row q of the output is the stride*Cin contiguous floats da[b, stride*q .. stride*q + stride - 1, :] of the channels-last gradient.
extern "C" int tg_conv_dgrad_tf32(const float* dy, const float* wd, float* da, int B, int Tin, int Tout, int Cin, int N, int k, int stride,
                                  tg_stream stream) {
  TG_REQUIRE(dy && wd && da && B > 0 && Tin > 0 && Tout > 0 && Cin > 0 && N >= 8 && k > 0 && stride > 0, "tg_conv_dgrad_tf32");
  TG_REQUIRE((Cin & 3) == 0 && (N & 3) == 0 && Tin >= (Tout - 1) * stride + k, "tg_conv_dgrad_tf32(shape)");
  TG_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(wd) | reinterpret_cast<uintptr_t>(da)) & 15) == 0,
             "tg_conv_dgrad_tf32(alignment)");
  const int Q = tg_ceil_div(Tin, stride), NP = stride * Cin;
  tg_gemm_tf32_t g;
  memset(&g, 0, sizeof(g));
  g.A = dy; g.lda = N; g.a_rows = Tout;
  g.Bw = wd; g.ldb = N;
  g.C = da; g.ldc = NP;
  g.M = B * Q; g.N = NP; g.K = N; g.taps = 1; g.T = 1;
  g.clip_rows = Q; g.a_clip_pitch = (long long)Tout * N;
  TapExt x;
  x.acc_taps = tg_ceil_div(k, stride);
  x.a_rows = Tout;
  x.c_clip_pitch = (long long)Tin * Cin;
  x.c_clip_len = (long long)Tin * Cin;
  cudaStream_t s = (cudaStream_t)stream;
  if (NP <= 32) return launch<32, 4>(g, s, &x);
  if (NP <= 64) return launch<64, 4>(g, s, &x);
  const int p128 = tg_ceil_div(NP, 128) * 128, p160 = tg_ceil_div(NP, 160) * 160;
  if (p160 < p128) return launch<160, 3>(g, s, &x);
  return launch<128, 3>(g, s, &x);
}

extern "C" int tg_debug_gemm_trace(long long* device_buf) {
  g_trace = device_buf;
  return 0;
}
long long* tg_gemm_trace_ptr() { return g_trace; }

// gemm_tcn.cu: the two-tap case as a one-accumulator clip-group kernel
bool tg_gemm_tcn_applies(const tg_gemm_tf32_t& g);
int tg_gemm_tcn_launch(const tg_gemm_tf32_t& g, cudaStream_t s);

extern "C" int tg_gemm_tf32(const tg_gemm_tf32_t* gp, tg_stream stream) {
  const tg_gemm_tf32_t& g = *gp;
  TG_REQUIRE(g.A && g.Bw && g.C && g.M > 0 && g.N > 0 && g.K > 0, "tg_gemm_tf32");
  TG_REQUIRE(g.taps == 1 || g.taps == 2, "tg_gemm_tf32");
  // epilogue activations are evaluated as max(v, v*s): none, relu, leaky-relu with 0 <= slope <= 1 (no sigmoid on this path)
  TG_REQUIRE(g.act1 >= 0 && g.act1 <= 2 && (g.act1 != 2 || (g.slope1 >= 0.f && g.slope1 <= 1.f)) && (g.act2 == 0 || g.act2 == 1), "tg_gemm_tf32(activation)");
  TG_REQUIRE(g.clip_rows == 0 || (g.clip_rows > 0 && g.taps == 1 && g.M % g.clip_rows == 0 && g.a_clip_pitch > 0), "tg_gemm_tf32(clip mode)");
  cudaStream_t s = (cudaStream_t)stream;
  static const bool tcn_old = getenv("TGB200_TCN_TWO_ACC") != nullptr;      // A/B switch: keep the two-accumulator tiles for taps == 2
  if (!tcn_old && tg_gemm_tcn_applies(g)) return tg_gemm_tcn_launch(g, s);
  // tile width: minimise padded N, prefer wider tiles on ties (fewer A re-reads)
  int bn;
  if (g.N <= 32) bn = 32;
  else if (g.N <= 64) bn = 64;
  else {
    const int p128 = tg_ceil_div(g.N, 128) * 128, p160 = tg_ceil_div(g.N, 160) * 160, p240 = tg_ceil_div(g.N, 240) * 240;
    bn = (p160 < p128) ? 160 : 128;
    // wide outputs (the GRU input projections, N = 1800): a tile's operand fetch, (128 + BN) x 128 bytes per 128 x BN x 32 block, is what
    // bounds these GEMMs (an SM ingests ~46 bytes/clk from L2 while a 128 x 128 block needs 128 bytes per MMA clock), so take the widest
    // tile that does not add padding: 240 columns with two stages still leaves two CTAs per SM for the epilogue overlap
    static const bool no240 = getenv("TGB200_NO_BN240") != nullptr;
    if (!no240 && g.taps == 1 && g.N >= 960 && p240 <= (bn == 160 ? p160 : p128)) bn = 240;
  }
  // 3-4 stages keep two CTAs resident per SM (shared memory and 2 x 256 TMEM columns), so one CTA's epilogue overlaps
  // the other's main loop; the two-accumulator mode is limited to BN <= 128 for the same reason (2 x 2 x 128 = 512 columns)
  if (g.taps == 2 && bn == 160) bn = 128;
  // long-K problems with few output tiles (the GRU data gradient [4352 x 1800] x [1800 x 600]): 256-row x 160-column tiles, one CTA per
  // SM, split-K sized to ONE wave.  The 128-row tiles ran 1.7 waves of 2 CTAs per SM (510 CTAs) and took 63 us for ~30 us of operand
  // traffic; the conditions mirror launch()'s split-K test (linear epilogue, vector-aligned contiguous C)
  {
    static const bool off = getenv("TGB200_NO_WIDE_SPLITK") != nullptr || getenv("TGB200_NO_SPLITK") != nullptr;
    const bool linear = g.taps == 1 && g.clip_rows == 0 && g.act1 == 0 && g.act2 == 0 && !g.residual;
    const bool vec = (g.N & 3) == 0 && g.ldc == g.N && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 &&
                     (!g.mask || ((g.ldmask & 3) == 0 && (reinterpret_cast<uintptr_t>(g.mask) & 15) == 0));
    const int tiles = tg_ceil_div(g.M, 256) * tg_ceil_div(g.N, 160);
    if (!off && linear && vec && g.N > 128 && g.M >= 256 && tg_ceil_div(g.K, BKF) >= 24 && tiles <= tg_num_sms()) return launch<160, 3, 2>(g, s);
  }
  if (bn == 32) return launch<32, 4>(g, s);
  if (bn == 64) return launch<64, 4>(g, s);
  if (bn == 128) return launch<128, 3>(g, s);
  if (bn == 240) {
    static const bool no256 = getenv("TGB200_NO_BM256") != nullptr;
    if (!no256 && g.clip_rows == 0 && g.M >= 256) return launch<240, 3, 2>(g, s);
    return launch<240, 2>(g, s);
  }
  return launch<160, 3>(g, s);
}
