// Weight-gradient GEMM on the tensor cores (fast mode):  dW[n, c] += sum_{b,t} G[(b,t), n] * X[(b, t + shift), c]
// Both operands are read in their natural row-major layout, i.e. MN-major for the MMA (the reduction index runs over
// rows): TMA (3-D maps: channel x time x clip, so rows outside a clip are zero-filled by the hardware - that is the
// whole causal / shifted-tap boundary handling) lands [TK rows x 32 channels] boxes in the MN-major atom layout of 32-bit
// operands (4 rows x 128 B, 32-byte-chunk swizzle: TMA SWIZZLE_128B_ATOM_32B == UMMA SWIZZLE_128B_BASE32B) that
// tcgen05.mma kind::tf32 consumes with a_major = b_major = MN.  128(n) x 128(c) accumulator in TMEM,
// split-K over clips across CTAs, fp32 atomic accumulation into the flat gradient arena (grads accumulate like .grad).
#include <cudaTypedefs.h>
#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;


struct WgP {
  float* dW; int ldw; int N, Cin;
  int B, T, TKrows, nt;       // nt = chunks per clip
  int shift;
  int chunks_per_split;
};

// MN-major operand descriptor: 8 x (8 rows x 128 B) atoms; LBO = bytes between 32-element groups along MN, SBO = bytes
// between 8-row groups along K
template <int TK, int NSTAGE, int MINB>
__global__ void __launch_bounds__(320, MINB) wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                                                            const WgP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int BOX = TK * 128;                 // one [TK rows x 32 floats] box
  constexpr int OPER = 4 * BOX;                 // 128 MN elements
  constexpr int STAGE = 2 * OPER;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tmem_full = empty_bar + NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
  const int total_chunks = p.B * p.nt;
  const int ch0 = blockIdx.z * p.chunks_per_split;
  const int ch1 = min(total_chunks, ch0 + p.chunks_per_split);
  const int iters = ch1 - ch0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmX);
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // warp-uniform for ptxas (feeds uniform registers)

  if (iters > 0) {
    if (warp == 0) {
      if (elect_one()) {
        for (int it = 0; it < iters; ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const int ch = ch0 + it;
          const int b = ch / p.nt, t0 = (ch - b * p.nt) * TK;
          uint8_t* sa = smem + s * STAGE;
          uint8_t* sb = sa + OPER;
          mbar_expect_tx(&full_bar[s], STAGE);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tma_load_3d(sa + j * BOX, &tmG, &full_bar[s], n0 + 32 * j, t0, b);
            tma_load_3d(sb + j * BOX, &tmX, &full_bar[s], c0 + 32 * j, t0 + p.shift, b);
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = idesc_tf32(128, 128, 1, 1);
        for (int it = 0; it < iters; ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE);
          const uint32_t sb = sa + OPER;
          // MN-major 32-bit atoms (layout 1 = SWIZZLE_128B_BASE32B): LBO = BOX bytes between 32-element groups, SBO = 512 B between
          // 4-row K groups; each K-step of 8 rows advances the start address by 1024 B (64 x 16 B)
          constexpr uint32_t hi = desc_hi(512, 1);
          static_assert(TK % 8 == 0 && TK / 8 >= 1 && TK / 8 <= 8, "K-steps per stage");
          if constexpr (TK / 8 <= 4) {
            mma_tf32_seq<TK / 8, 64>(tmem_base, desc_lo(sa, BOX), hi, desc_lo(sb, BOX), hi, idesc, it > 0 ? 1u : 0u);
          } else {
            mma_tf32_seq<4, 64>(tmem_base, desc_lo(sa, BOX), hi, desc_lo(sb, BOX), hi, idesc, it > 0 ? 1u : 0u);
            mma_tf32_seq<TK / 8 - 4, 64>(tmem_base, desc_lo(sa + 4096, BOX), hi, desc_lo(sb + 4096, BOX), hi, idesc, 1u);
          }
          tc_commit(&empty_bar[s]);
        }
        tc_commit(tmem_full);
      }
    } else {
      // 8 epilogue warps: warps w and w+4 share TMEM lane quarter w%4 and split its four 32-column chunks (even / odd)
      const int q = warp & 3, half = (warp - 2) >> 2;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      // every stage is free once tmem_full fires: transpose each 32x32 chunk through a padded staging tile so that one
      // warp-wide reduction covers 4 rows x 128 contiguous bytes of dW (vector red.global.add.v4.f32 where aligned)
      constexpr int SP = 36;
      float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * SP);
      const int rsub = lane >> 3, c4 = (lane & 7) * 4;
      const bool vec_ok = (p.ldw & 3) == 0 && (p.Cin & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dW) & 15) == 0;
      const int row0 = n0 + q * 32 + rsub;
      const int rows_left = p.N - row0;
#pragma unroll 1
      for (int cc = half * 32; cc < 128; cc += 64) {
        if (c0 + cc >= p.Cin) break;
        float v[32];
        tmem_ld32(lane_addr + (uint32_t)cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) *reinterpret_cast<float4*>(stg + lane * SP + j4) = make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
        __syncwarp();
        const int c = c0 + cc + c4;
        if (c < p.Cin) {
          float* dst = p.dW + (long long)row0 * p.ldw + c;
          const float* srow = stg + rsub * SP + c4;
          const long long dstep = 4ll * p.ldw;
          if (vec_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (4 * i < rows_left) {
                const float4 x = *reinterpret_cast<const float4*>(srow + i * 4 * SP);
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i * dstep), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
              }
            }
          } else {
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              if (4 * i >= rows_left) break;
              for (int e = 0; e < 4 && c + e < p.Cin; ++e) atomicAdd(dst + i * dstep + e, srow[i * 4 * SP + e]);
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

__global__ void col_sum_kernel(const float* __restrict__ g, int ld, long long M, int N, float* out) {
  // grid.x over column blocks of 32, grid.y over row slabs; block (32, 8)
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const long long rows_per = (M + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.f;
  if (n < N)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += g[r * ld + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(out + n, t);
  }
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

// [B clips][T rows][cols] row-major, pitch ld floats, clip stride Tfull*ld: box = 32 cols x TK rows x 1 clip
int map_3d(CUtensorMap* m, const float* base, int B, long long T, long long clip_pitch, long long cols, long long ld, int TK, const char* name) {
  auto enc = get_encode();
  if (!enc) { tg_set_error("%s: cuTensorMapEncodeTiled unavailable", name); return -4; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 4) & 15) || ((clip_pitch * 4) & 15)) {
    tg_set_error("%s: TMA alignment (ld=%lld clip pitch=%lld)", name, ld, clip_pitch); return -1;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)clip_pitch * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)TK, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { tg_set_error("%s: cuTensorMapEncodeTiled(3d) failed (%d) B=%d T=%lld cols=%lld ld=%lld", name, (int)r, B, T, cols, ld); return -4; }
  return 0;
}

template <int TK, int NSTAGE, int MINB>
int launch(const tg_wgrad_tf32_t& g, int B, int T, cudaStream_t s) {
  CUtensorMap tg_, tx_;
  int rc;
  const long long xcp = (g.x_clip_pitch > 0 && B > 1) ? g.x_clip_pitch : (long long)T * g.ldx;
  if ((rc = map_3d(&tg_, g.G, B, T, (long long)T * g.ldg, g.N, g.ldg, TK, "tg_wgrad_tf32(G)"))) return rc;
  if ((rc = map_3d(&tx_, g.X, B, T, xcp, g.Cin, g.ldx, TK, "tg_wgrad_tf32(X)"))) return rc;
  WgP p;
  p.dW = g.dW; p.ldw = g.ldw; p.N = g.N; p.Cin = g.Cin; p.B = B; p.T = T; p.TKrows = TK; p.nt = tg_ceil_div(T, TK); p.shift = g.shift;
  const int tiles = tg_ceil_div(g.Cin, 128) * tg_ceil_div(g.N, 128);
  const int total_chunks = B * p.nt;
  // one wave of resident CTAs (MINB per SM); at least 4 chunks per CTA so that the fixed per-CTA cost (TMEM allocation,
  // pipeline fill, 128x128 atomic epilogue) and the number of atomics per gradient element stay bounded for small matrices
  int splits = (MINB * tg_num_sms()) / tiles;
  if (splits < 1) splits = 1;
  if (splits > total_chunks) splits = total_chunks;
  p.chunks_per_split = tg_ceil_div(total_chunks, splits);
  if (p.chunks_per_split < 4) p.chunks_per_split = total_chunks < 4 ? total_chunks : 4;
  splits = tg_ceil_div(total_chunks, p.chunks_per_split);
  constexpr size_t smem = (size_t)NSTAGE * 2 * 4 * TK * 128 + (2 * NSTAGE + 1) * 8 + 16 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tf32_kernel<TK, NSTAGE, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { tg_set_error("tg_wgrad_tf32: smem attr: %s", cudaGetErrorString(e)); return -3; }
    attr_done = true;
  }
  dim3 grid(tg_ceil_div(g.Cin, 128), tg_ceil_div(g.N, 128), splits);
  wgrad_tf32_kernel<TK, NSTAGE, MINB><<<grid, 320, smem, s>>>(tg_, tx_, p);
  TG_CHECK_LAUNCH("tg_wgrad_tf32");
  return 0;
}

}  // namespace

extern "C" int tg_col_sum_f32(const float* g, int ld, long long M, int N, float* out, tg_stream stream) {
  TG_REQUIRE(g && out && M > 0 && N > 0, "tg_col_sum_f32");
  int slabs = (int)((M + 255) / 256);
  int cap = (4 * tg_num_sms()) / tg_ceil_div(N, 32);
  if (cap < 1) cap = 1;
  if (slabs > cap) slabs = cap;
  col_sum_kernel<<<dim3(tg_ceil_div(N, 32), slabs), dim3(32, 8), 0, (cudaStream_t)stream>>>(g, ld, M, N, out);
  TG_CHECK_LAUNCH("tg_col_sum_f32");
  return 0;
}

extern "C" int tg_wgrad_tf32(const tg_wgrad_tf32_t* gp, tg_stream stream) {
  const tg_wgrad_tf32_t& g = *gp;
  TG_REQUIRE(g.G && g.X && g.dW && g.B > 0 && g.T > 0 && g.N > 0 && g.Cin > 0, "tg_wgrad_tf32");
  cudaStream_t s = (cudaStream_t)stream;
  if (g.dbias) {
    int rc = tg_col_sum_f32(g.G, g.ldg, (long long)g.B * g.T, g.N, g.dbias, stream);
    if (rc) return rc;
  }
  if (g.x_clip_pitch > 0) {
    // X rows are per-clip windows (row pitch ldx < Cin allowed: strided-convolution windows read in place): 32-row chunks per clip
    return launch<32, 3, 2>(g, g.B, g.T, s);
  }
  if (g.shift == 0 || g.B == 1) {
    // no clip boundaries to respect: one flat sequence of rows, 32-row chunks
    return launch<32, 3, 2>(g, 1, g.B * g.T, s);      // 3 stages x 32 KB: two CTAs per SM
  }
  TG_REQUIRE(g.T <= 40, "tg_wgrad_tf32");       // shifted taps: one zero-padded chunk per clip
  return launch<40, 4, 1>(g, g.B, g.T, s);
}
