/* tg_b200.h — C ABI of libtg_b200.so: hand-written sm_100a kernels for the trimodal-gesture hot path.
 *
 * The reference (ai4r/Gesture-Generation-from-Trimodal-Context) has no FFI of its own: every FLOP of its hot
 * path is a stock PyTorch op.  Each entry point below therefore cites the reference call site
 * (scripts/... file:line) whose PyTorch op(s) it replaces.  Conventions:
 *   - plain C types only; device pointers are borrowed (never freed, never retained after the call returns);
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns immediately:
 *     no allocation, no synchronisation, safe under CUDA-graph capture;
 *   - return value 0 = ok, negative = error; tg_last_error() gives a thread-local message;
 *   - activations are "channels-last": a [B, T, C] clip tensor is a row-major matrix of B*T rows x C columns.
 */
#ifndef TG_B200_H
#define TG_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* tg_stream;

const char* tg_last_error(void);
int tg_version(void);
/* device properties the host side sizes persistent grids with: out[0]=SM count, out[1]=max opt-in smem/block */
int tg_device_info(int* out2);
/* out[0] = sizeof(tg_conv_gemm_t), out[1] = sizeof(tg_conv_wgrad_t): lets a foreign-language binding verify its layout */
int tg_struct_sizes(int* out2);

/* ---------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear layer, fp32 SIMT ("strict fp32" mode).
 *   Y[orow(b,t), n] = epi( sum_{j<taps} sum_{c<Cin} pro(A[b*Tin + t*stride + j*dil - pad, c]) * W(n,j,c) )
 * with m = b*Tout + t, out-of-range input rows reading as zero (causal left pad of tcn.py:7-13,19-31; the
 * pad=1600 of multimodal_context_net.py:13).  W(n,j,c) = W[n*ldw + j*wsj + c*wsc] so that nn.Conv1d weights
 * [N,Cin,k] (wsj=1, wsc=k), nn.Linear weights [N,K] (taps=1, wsc=1) and transposed views are read in place.
 * Output row orow = b*ToutFull + t*ostride + ooff lets a strided-conv dgrad be issued as `stride` dense phases.
 * Replaces: nn.Conv1d / nn.Linear / F.conv_transpose1d forward and input-gradient at
 *   multimodal_context_net.py:12-23,48,60,88-93,100-104,213-226; tcn.py:19-31; embedding_net.py:46-65,188-206.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* A;  int lda;                 /* input rows, pitch in floats */
  int asc;                                  /* channel stride of A in floats (0 or 1 = contiguous channels) */
  long long a_bstride;                      /* floats between clips of A (0 = Tin*lda) */
  const float* W;  int ldw, wsj, wsc;
  float* Y;        int ldc;
  int B, Tout, Tin, N, Cin, taps, stride, dil, pad;
  int ToutFull, ostride, ooff;
  /* prologue on A (in-range elements only): a' = lrelu(a*pscale[c] + pshift[c], pslope); NULL = identity */
  const float* pscale; const float* pshift; float pslope;
  /* epilogue: v = acc*escale[n] + bias[n]; act1; v *= mask[orow, n]; v += residual[orow, n]; act2; (+= Y if accumulate) */
  const float* escale; const float* bias;
  int act1; float slope1;                   /* 0 none, 1 relu, 2 leaky-relu(slope1), 3 sigmoid */
  const float* mask;   int ldmask;
  const float* residual; int ldres;
  int act2;
  int accumulate;
} tg_conv_gemm_t;
int tg_conv_gemm_f32(const tg_conv_gemm_t* p, tg_stream stream);

/* Weight gradient of the same operator: dW(n,j,c) += sum_m G[m*ldg + n] * pro(A[row(m,j), c]); atomically accumulated
 * (grads accumulate across backward calls exactly like torch .grad does).  Replaces the autograd of the ops above. */
typedef struct {
  const float* A;  int lda;
  const float* G;  int ldg;
  float* dW;       int ldw, wsj, wsc;
  int B, Tout, Tin, N, Cin, taps, stride, dil, pad;
  const float* pscale; const float* pshift; float pslope;
  float* dbias;                             /* optional: dbias[n] += sum_m G[m,n] */
} tg_conv_wgrad_t;
int tg_conv_wgrad_f32(const tg_conv_wgrad_t* p, tg_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * "Fast" mode GEMM on the 5th-generation tensor cores: tcgen05.mma kind::tf32 fed by TMA, accumulators in TMEM.
 *   C[m, n] = epi( sum_{tap < taps} sum_k A[m + shift(tap), k] * Bw[tap*N + n, k] ),  shift(0) = shift0 if taps == 2 else 0
 * A [a_rows, K] and Bw [taps*N, K] are row-major fp32 (pitches lda / ldb floats, 16-byte aligned); operands are read
 * as TF32 (10-bit mantissa), accumulation is fp32.  taps == 2 is the TCN causal dilated convolution (tcn.py:19-31)
 * and its data gradient: the shifted tap only contributes to rows whose clip-local time t = m % T satisfies
 * 0 <= t + shift0 < T.  Epilogue as tg_conv_gemm_f32.  Same reference call sites as tg_conv_gemm_f32.
 * Clip mode (clip_rows > 0, taps == 1): the M = clips*clip_rows logical rows of A are addressed as
 * A + clip*a_clip_pitch + t*lda (+ k), where lda may be SMALLER than K - row t is then the window of a strided
 * convolution over a channels-last signal (lda = stride*Cin, K = k*Cin: WavEncoder conv2-4,
 * multimodal_context_net.py:16-22) read in place by the TMA unit: no im2col copy.  C / mask / residual stay flat [M, .].
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* A;  int lda; long long a_rows;
  const float* Bw; int ldb;
  float* C;        int ldc;
  int M, N, K, taps, shift0, T;
  const float* escale; const float* bias;
  int act1; float slope1;
  const float* mask;   int ldmask;
  const float* residual; int ldres;
  int act2;
  int accumulate;
  int clip_rows; long long a_clip_pitch;    /* clip mode, see above (0 = flat) */
} tg_gemm_tf32_t;
int tg_gemm_tf32(const tg_gemm_tf32_t* p, tg_stream stream);

/* Fast-mode weight gradient on the tensor cores:  dW[n*ldw + c] += sum_{b,t} G[(b*T+t)*ldg + n] * X[(b*T+t+shift)*ldx + c]
 * with rows t+shift outside [0,T) of clip b contributing zero (shift != 0 needs T <= 40: the TCN taps, tcn.py:19-31,
 * and the recurrent weight_hh gradient whose operand is the layer output one step earlier / later).
 * dbias (optional): dbias[n] += sum G[., n].  TF32 operands, fp32 accumulate, atomic accumulation into dW. */
typedef struct {
  const float* G; int ldg;
  const float* X; int ldx;
  float* dW;      int ldw;
  float* dbias;
  int B, T, N, Cin, shift;
  long long x_clip_pitch;   /* > 0: clip b of X starts at X + b*x_clip_pitch and its row t at + t*ldx, where ldx may be smaller
                               than Cin (row t = the window of a strided convolution, Cin = k*channels: WavEncoder conv2-4
                               weight gradients, multimodal_context_net.py:16-22); 0: X is the flat [B*T, Cin] matrix */
} tg_wgrad_tf32_t;
int tg_wgrad_tf32(const tg_wgrad_tf32_t* p, tg_stream stream);
int tg_col_sum_f32(const float* g, int ld, long long M, int N, float* out, tg_stream stream);

/* Direct strided convolution for a single input channel (WavEncoder conv1: Conv1d(1,16,15,stride 5,pad 1600),
 * multimodal_context_net.py:13).  HBM-bound: x [B,Tin] -> y [B,Tout,N] channels-last, N <= 32, taps <= 32. */
int tg_conv1_direct_f32(const float* x, const float* w, const float* bias, float* y,
                        int B, int Tin, int Tout, int N, int taps, int stride, int pad, tg_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Tensor-core WavEncoder path (fast mode; multimodal_context_net.py:12-23 forward and autograd).  conv2-4 run as
 * tg_gemm_tf32 (clip mode) / tg_wgrad_tf32 (x_clip_pitch) over the overlapping-window view of the channels-last
 * activation (BatchNorm + LeakyReLU(0.3) materialised by tg_affine_lrelu); these are the CUDA-core pieces around them.
 * --------------------------------------------------------------------------------------------------------- */
/* nn.Conv1d filter w [N,Cin,k] -> w2 [N, k*Cin] (tap-major: the element order of a window row) and, if non-NULL,
 * w2t [k*Cin, N] (operand of the data-gradient GEMM) */
int tg_window_weights(const float* w, float* w2, float* w2t, int N, int Cin, int k, tg_stream stream);
/* dw [N,Cin,k] += dw2 [N, k*Cin] */
int tg_window_wgrad_add(const float* dw2, float* dw, int N, int Cin, int k, tg_stream stream);
/* data gradient of a strided, unpadded Conv1d from the "column" matrix col [B*Tout, k*Cin] = dY @ w2:
 * da[b, s, c] = sum over (t, j) with t*stride + j == s of col[(b*Tout+t), j*Cin + c]; gather, no atomics; Cin % 4 == 0 */
int tg_col2im(const float* col, float* da, int B, int Tin, int Tout, int Cin, int k, int stride, tg_stream stream);
/* The same data gradient WITHOUT the column matrix (the column GEMM wrote and tg_col2im re-read 161 MB for conv2): with t = stride*q + r,
 *   da[b, t, c] = sum_{j < ceil(k/stride)} dy[b, q - j, :] . w[:, c, r + stride*j]      (dy rows outside [0,Tout) and taps >= k are zero)
 * i.e. one tensor-core GEMM over the rows q whose output row is the stride*Cin contiguous floats da[b, stride*q .. +stride-1, :], its
 * ceil(k/stride) taps accumulating in one TMEM accumulator; the shifted dy rows are read in place by TMA (zero fill outside the clip).
 * tg_window_dgrad_weights: nn.Conv1d filter w [N,Cin,k] -> wd [ceil(k/stride)][stride*Cin][N], wd[j][r*Cin+c][n] = w[n][c][r+stride*j].
 * dy [B,Tout,N], da [B,Tin,Cin] channels-last, pad 0, Cin % 4 == 0, N % 4 == 0, N >= 8, 16-byte aligned.
 * (autograd of F.conv1d in WavEncoder, multimodal_context_net.py:16-22) */
int tg_window_dgrad_weights(const float* w, float* wd, int N, int Cin, int k, int stride, tg_stream stream);
int tg_conv_dgrad_tf32(const float* dy, const float* wd, float* da, int B, int Tin, int Tout, int Cin, int N, int k, int stride,
                       tg_stream stream);
/* conv1 (Cin = 1, N = 16, taps <= 15) weight and bias gradient: dW[n,j] += sum dy[(b,t),n] * x[b, t*stride + j - pad];
 * dbias[n] += sum dy[(b,t),n] (dbias may be NULL) */
int tg_conv1_wgrad(const float* x, const float* dy, float* dW, float* dbias, int B, int Tin, int Tout, int N, int taps, int stride,
                   int pad, tg_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * BatchNorm1d, train mode (multimodal_context_net.py:14,17,20,214,217) over a [M,C] channels-last matrix.
 * --------------------------------------------------------------------------------------------------------- */
/* sums[0..C) += column sums, sums[C..2C) += column sums of squares (fp64 accumulators, caller zeroes them) */
int tg_col_stats_f64(const float* x, int ld, long long M, int C, double* sums, tg_stream stream);
/* from the fp64 sums: mean/rstd (biased var, eps), fused scale = gamma*rstd, shift = beta - mean*scale, and
 * `n_updates` momentum updates of running_mean/running_var (unbiased var) + num_batches_tracked += n_updates */
int tg_bn_finalize(const double* sums, long long M, int C, float eps, float momentum, int n_updates,
                   const float* gamma, const float* beta, float* running_mean, float* running_var,
                   long long* num_batches_tracked, float* mean, float* rstd, float* scale, float* shift, tg_stream stream);
/* eval-mode BatchNorm folded to an affine map: scale = gamma/sqrt(rv+eps), shift = beta + (conv_bias - rm)*scale
 * (conv_bias may be NULL = 0).  Used as conv prologue (WavEncoder) or epilogue (EmbeddingNet). */
int tg_bn_eval_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                    const float* conv_bias, float* scale, float* shift, int C, tg_stream stream);
/* y = lrelu(x*scale[c] + shift[c], slope) */
int tg_affine_lrelu(const float* x, float* y, long long M, int C, const float* scale, const float* shift, float slope,
                    tg_stream stream);
/* backward of y = lrelu(bn(x)): pass 1 accumulates sums[0..C) += sum dz, sums[C..2C) += sum dz*xhat (dz = dy*lrelu') */
int tg_bn_bwd_reduce(const float* dy, const float* x, long long M, int C, const float* mean, const float* rstd,
                     const float* scale, const float* shift, float slope, double* sums, tg_stream stream);
/* pass 2: dx = gamma*rstd*(dz - sum_dz/M - xhat*sum_dzxhat/M); dgamma += sum dz*xhat; dbeta += sum dz */
int tg_bn_bwd_apply(const float* dy, const float* x, float* dx, long long M, int C, const float* mean, const float* rstd,
                    const float* scale, const float* shift, float slope, const float* gamma, const double* sums,
                    float* dgamma, float* dbeta, tg_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Embedding (multimodal_context_net.py:40-43,58,88) with optional dropout mask; dense scatter-add backward.
 * idx_mod > 0 reads idx[m % idx_mod] (the three generator passes of one train_iter_gan share in_text).
 * --------------------------------------------------------------------------------------------------------- */
int tg_embedding_gather(const float* table, const long long* idx, int idx_mod, const float* mask, float* out,
                        long long M, int E, tg_stream stream);
int tg_embedding_scatter_add(const float* dout, const long long* idx, const float* mask, float* dtable,
                             long long M, int E, tg_stream stream);

/* torch.nn.utils.weight_norm, dim=0 (tcn.py:19-25): w = g*v/||v|| per output channel, v [N,Cin,taps] in nn.Conv1d
 * layout.  The effective weight is emitted TAP-MAJOR, w[tap][n][c] (each tap a K-contiguous [N,Cin] GEMM operand), plus
 * optionally its per-tap transpose wT[tap][c][n] (operand of the data gradient).  Backward reads dw tap-major. */
int tg_weight_norm_fwd(const float* v, const float* g, float* w, float* wT, float* inv_norm, int N, int Cin, int taps,
                       tg_stream stream);
int tg_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dv, float* dg,
                       int N, int Cin, int taps, tg_stream stream);

/* elementwise helpers: out = a*b ; out = a+b ; dx = dy*mask*(y>0) (ReLU+dropout backward, tcn.py:22-23,28-29,46) */
int tg_mul(const float* a, const float* b, float* out, long long n, tg_stream stream);
/* out = a + b, optionally followed by ReLU (TemporalBlock residual, tcn.py:46) */
int tg_add(const float* a, const float* b, float* out, long long n, int relu, tg_stream stream);
int tg_relu_mask_bwd(const float* dy, const float* y, const float* mask, float* dx, long long n, tg_stream stream);
/* TemporalBlock backward head (tcn.py:28-29,46) for the fast-mode forward that folds the residual add + final ReLU into conv2's GEMM epilogue
 * (xo = relu(y2 + x) with y2 = relu(conv2 + b) * mask never stored):  dpre = dxo * (xo > 0);  dc2 = dpre * mask * (xo - x > 0).
 * Where dpre != 0, xo = y2 + x, so xo - x > 0 is y2 > 0 (up to the rounding of that one addition); mask may be NULL (eval). */
int tg_tcn_res_bwd(const float* dxo, const float* xo, const float* x, const float* mask, float* dpre, float* dc2, long long n, tg_stream stream);
/* out[m, 0..H) = x[m, 0..H) + x[m, H..2H)  (sum of GRU directions, multimodal_context_net.py:156,243) and its backward */
int tg_sum_halves(const float* x, float* out, long long M, int H, tg_stream stream);
int tg_dup_halves(const float* d, float* dx, long long M, int H, tg_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Speaker style sampling (multimodal_context_net.py:125-131, embedding_net.py:10-13): z = mu + eps*exp(0.5*logvar)
 * and GRU-input assembly cat(pre_seq, audio, text, repeat(z)) (multimodal_context_net.py:139-153).
 * --------------------------------------------------------------------------------------------------------- */
int tg_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z, long long n, tg_stream stream);
/* dmu += dz ; dlogvar += dz*eps*0.5*exp(0.5*logvar) */
int tg_reparam_bwd(const float* dz, const float* logvar, const float* eps, float* dmu, float* dlogvar, long long n,
                   tg_stream stream);
/* seed-pose conditioning input of train_iter_gan (train_gan.py:20-22): pre [B,T,D+1] = target for t < n_pre with a
 * constraint bit 1 in the last channel, zero elsewhere */
int tg_make_pre_seq(const float* target, float* pre, int B, int T, int D, int n_pre, tg_stream stream);
/* pre and audio have Ba <= B clips and are read at b % Ba (the generator passes of one train_iter_gan share them).
 * Widths: pre Dp, audio Da, text Dt, z Dz (any may be 0). */
int tg_gru_input_concat(const float* pre, const float* audio, const float* text, const float* z, float* out,
                        int B, int Ba, int T, int Dp, int Da, int Dt, int Dz, tg_stream stream);
/* split d_in -> d_audio[B,T,Da], d_text[B,T,Dt], d_z[B,Dz] (= sum over T) */
int tg_gru_input_split_bwd(const float* din, float* daudio, float* dtext, float* dz,
                           int B, int T, int Dp, int Da, int Dt, int Dz, tg_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Bidirectional GRU layer recurrence (nn.GRU at multimodal_context_net.py:98-99,155,221-222,241; gate order r,z,n).
 * gi = x W_ih^T + b_ih for both directions is a tg_conv_gemm_f32 call ([B*T, 6H]: fwd r,z,n | rev r,z,n); the
 * sequential part runs as ONE persistent cooperative kernel per layer: W_hh slices stay resident in shared memory
 * across (unit-chunk x batch-tile x direction) CTAs which exchange h_t through L2 with release/acquire counters.
 *   out   [B,T,2H]  (fwd | rev)
 *   saved [4][B*T][2H] = r, z, n, hn (= W_hn h + b_hn) for the backward; may be NULL (inference)
 *   sync  int[tg_gru_sync_ints(B,H)] scratch, zeroed by the call
 * --------------------------------------------------------------------------------------------------------- */
/* whhT_* = transposed recurrent weights [H, 3H] (tg_transpose_f32 of weight_hh_l*), saved_qstride = floats between
 * the r/z/n/hn planes of `saved` (lets a backward run on a batch slice of a larger forward). */
int tg_gru_layer_fwd(const float* gi, const float* whhT_f, const float* whhT_r, const float* bhh_f, const float* bhh_r,
                     float* out, float* saved, long long saved_qstride, int* sync, int B, int T, int H, tg_stream stream);
/* BPTT: dout [B,T,2H] -> dgi [B*T,6H] (grad of gi = x W_ih^T + b_ih) and dgh [B*T,6H] (grad of gh = h W_hh^T + b_hh);
 * the weight grads are then tg_conv_wgrad_f32 calls (dW_hh uses A = out shifted by one step, pad = +1 / -1).
 * whh_* = weight_hh [3H, H] as stored.  partial: float scratch of tg_gru_bwd_scratch_floats(B,H). */
size_t tg_gru_bwd_scratch_floats(int B, int H);
int tg_gru_sync_ints(int B, int H);
int tg_gru_layer_bwd(const float* dout, const float* out, const float* saved, long long saved_qstride,
                     const float* whh_f, const float* whh_r, float* dgi, float* dgh, float* partial, int* sync,
                     int B, int T, int H, tg_stream stream);
int tg_transpose_f32(const float* in, float* out, int R, int C, tg_stream stream);

/* Fast-mode recurrence: the same persistent kernels with the per-step product on tcgen05 tensor cores (TF32 operands,
 * fp32 accumulate in TMEM, W_hh resident in shared memory via TMA).  NOTE the weight layouts are swapped with respect
 * to the fp32 kernels: forward takes weight_hh as stored [3H,H], backward takes its transpose [H,3H].
 * H % 4 == 0, 32 <= H <= 384.  sync: int[tg_gru_tf32_sync_ints(B,H)], partial: tg_gru_bwd_tf32_scratch_floats(B,H). */
int tg_gru_tf32_sync_ints(int B, int H);
/* development aid: when a device buffer of >= 16*T int64 is set, CTA (0,0,0) of the tensor-core GRU kernels stamps
 * %globaltimer at its per-step phases (NULL switches it off) */
/* ConvDiscriminator recurrent stack in one launch (multimodal_context_net.py:221-222,241-251): L-layer bidirectional GRU(I0 -> 64) over
 * x [B,T,I0], inter-layer dropout (masks[l]: already scaled keep-masks [B*T,2H] on the output of layer l < L-1, or NULL), sum of the
 * directions, Linear(64,1) per frame, Linear(T,1), sigmoid -> prob [B,1].  gru_params: the discriminator's GRU parameters as the flat
 * arena stores them (per layer: weight_ih | weight_ih_reverse | bias_ih | bias_ih_reverse | weight_hh | weight_hh_reverse | bias_hh |
 * bias_hh_reverse).  outs[l] / saved[l] (4 planes r,z,n,W_hn h + b_hn; plane stride saved_qstride) / drops[l] (masked outputs, l < L-1) /
 * hsum [B*T,64] / o1 [B*T] are what the per-layer backward kernels read.  H == 64, L <= 4, T <= 32, I0 <= 32, I0 % 4 == 0; x, gru_params
 * and the masks 16-byte aligned.  fast != 0 (the tensor-core arithmetic mode): sigmoid / tanh through ex2.approx / rcp.approx. */
int tg_dgru_stack_fwd(const float* x, const float* gru_params, const float* const* masks, float* const* outs, float* const* saved,
                      long long saved_qstride, float* const* drops, const float* w_out, const float* b_out, const float* w_out2,
                      const float* b_out2, float* hsum, float* o1, float* prob, int B, int T, int I0, int H, int L, int fast, tg_stream stream);

/* Backward of tg_dgru_stack_fwd in one launch: heads -> L x [recurrence backward + data gradient through W_ih (+ dropout mask of the layer
 * below)].  dlogit [B] = d loss / d (pre-sigmoid output).  Writes dgi[l] / dgh[l] [B*T,6H] for the recurrent layers' weight-gradient
 * GEMMs (which stay separate launches), dx0 [B*T,I0] (optional) = gradient w.r.t. the stack input, and ACCUMULATES the four head
 * gradients (out.weight [64], out.bias [1], out2.weight [T], out2.bias [1]) with atomics.  fast != 0: the per-clip GEMMs (data gradient
 * through W_ih) run on warp-level TF32 tensor-core tiles (mma.sync m16n8k8), as the forward's input projections do. */
int tg_dgru_stack_bwd(const float* dlogit, const float* gru_params, const float* const* masks, const float* const* outs,
                      const float* const* saved, long long saved_qstride, const float* hsum, const float* o1, const float* w_out,
                      const float* w_out2, float* const* dgi, float* const* dgh, float* dx0, float* g_w_out, float* g_b_out,
                      float* g_w_out2, float* g_b_out2, int B, int T, int I0, int H, int L, int fast, tg_stream stream);

/* ConvDiscriminator convolution stack in one launch (multimodal_context_net.py:212-220,233-236): Conv1d(27,16,3) -> BatchNorm1d(16) ->
 * identity -> Conv1d(16,8,3) -> BatchNorm1d(8) -> identity -> Conv1d(8,8,3) on channels-last poses x [B,34,27] (B <= 128, one 8-CTA cluster,
 * BatchNorm sums all-reduced through distributed shared memory).  Weights as stored ([N,Cin,3]).  training != 0: batch statistics, running
 * buffers updated once (momentum, unbiased variance), num_batches_tracked += 1; else running statistics.  Outputs: y0 [B*32,16], y1 [B*30,8]
 * (the PRE-BatchNorm conv outputs the backward needs), y2 [B*28,8], st1 [4*16] / st2 [4*8] = mean | rstd | scale | shift per BatchNorm. */
int tg_dconv_stack_fwd(const float* x, const float* w1, const float* b1, const float* g1, const float* be1, float* rm1, float* rv1,
                       long long* nbt1, const float* w2, const float* b2, const float* g2, const float* be2, float* rm2, float* rv2,
                       long long* nbt2, const float* w3, const float* b3, float* y0, float* y1, float* y2, float* st1, float* st2,
                       int B, int T, int D, int training, float eps, float momentum, tg_stream stream);
/* Backward of tg_dconv_stack_fwd (train mode) in one launch: dy2 [B*28,8] -> ACCUMULATES the three convolutions' weight / bias gradients
 * and both BatchNorms' gamma / beta gradients (atomics), optionally writes dx [B,34,27]. */
int tg_dconv_stack_bwd(const float* dy2, const float* x, const float* y0, const float* y1, const float* st1, const float* st2,
                       const float* w1, const float* w2, const float* w3, const float* g1, const float* g2, float* dw1, float* db1,
                       float* dw2, float* db2, float* dw3, float* db3, float* dg1, float* dbe1, float* dg2, float* dbe2, float* dx,
                       int B, int T, int D, tg_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Speech2Gesture baseline (scripts/model/speech2gesture.py, scripts/train_eval/train_speech2gesture.py): the non-GEMM pieces.
 * Conv2d_tf / Conv1d_tf (speech2gesture.py:9-101) = tg_im2col2d -> tg_gemm_tf32 / tg_conv_gemm_f32 (forward), tg_wgrad_tf32 on the
 * column matrix (weight gradient), column GEMM -> tg_col2im2d (data gradient).  Activations channels-last [B,H,W,C]; a sequence is H = 1.
 * --------------------------------------------------------------------------------------------------------- */
/* col[((b*Ho+ho)*Wo+wo), (i*kw+j)*C + c] = x[b, ho*sh+i-pt, wo*sw+j-pl, c], zero outside the image (TF "SAME": pt / pl = total_pad / 2) */
int tg_im2col2d(const float* x, float* col, int B, int H, int W, int C, int kh, int kw, int sh, int sw, int pt, int pl, int Ho, int Wo,
                tg_stream stream);
/* adjoint of tg_im2col2d: dx[b,h,w,c] = sum of the column entries that were gathered from it (no atomics) */
int tg_col2im2d(const float* col, float* dx, int B, int H, int W, int C, int kh, int kw, int sh, int sw, int pt, int pl, int Ho, int Wo,
                tg_stream stream);
/* torch.nn.Upsample(size=(Ho,Wo), mode='bilinear', align_corners=False) (speech2gesture.py:147,172) and its adjoint (dx is overwritten) */
int tg_resize_bilinear_fwd(const float* x, float* y, int B, int H, int W, int C, int Ho, int Wo, tg_stream stream);
int tg_resize_bilinear_bwd(const float* dy, float* dx, int B, int H, int W, int C, int Ho, int Wo, tg_stream stream);
/* UnetUp (speech2gesture.py:120-130): y[b,t,:] = x1[b,t/2,:] + x2[b,t,:], t < T2 <= 2*T1; backward w.r.t. x1 (w.r.t. x2 it is dy itself) */
int tg_upsample2_add_fwd(const float* x1, const float* x2, float* y, int B, int T1, int T2, int C, tg_stream stream);
int tg_upsample2_bwd(const float* dy, float* dx1, int B, int T1, int T2, int C, int accumulate, tg_stream stream);
/* x[:, 1:] - x[:, :-1] over [B,T,D] (train_speech2gesture.py:12-13) and its adjoint */
int tg_time_diff_fwd(const float* x, float* y, int B, int T, int D, tg_stream stream);
int tg_time_diff_bwd(const float* dy, float* dx, int B, int T, int D, int accumulate, tg_stream stream);
/* torch.cat((a, p.unsqueeze(2).repeat(1,1,T)), dim=1) in channels-last form (speech2gesture.py:219-221) and its adjoint */
int tg_concat_bcast_fwd(const float* a, const float* p, float* y, int B, int T, int Ca, int Cp, tg_stream stream);
int tg_concat_bcast_bwd(const float* d, float* da, float* dp, int B, int T, int Ca, int Cp, tg_stream stream);
/* LeakyReLU backward without a BatchNorm in front (speech2gesture.py:237): dx = dy * (x >= 0 ? 1 : slope) */
int tg_lrelu_bwd(const float* dy, const float* x, float* dx, long long n, float slope, tg_stream stream);
/* F.mse_loss(full_like(x, target), x) (train_speech2gesture.py:20,30): *scalar += mean (fp64); dx (optional) = w * d mean / dx */
int tg_mse_const(const float* x, long long n, float target, float w, double* scalar, float* dx, tg_stream stream);
/* torch.nn.L1Loss (train.py:62): *scalar += mean |x - y|; dx (optional) = w * sign(x - y) / n */
int tg_l1_loss(const float* x, const float* y, long long n, float w, double* scalar, float* dx, tg_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * seq2seq baseline (scripts/model/seq2seq_net.py, scripts/train_eval/train_seq2seq.py; config/seq2seq.yml).
 * The projections are GEMMs (tg_conv_gemm_f32 / tg_gemm_tf32); these are the per-step non-GEMM pieces.
 * --------------------------------------------------------------------------------------------------------- */
/* One GRU time step for B rows from gi = W_ih x + b_ih (row pitch ldgi) and gh = W_hh h + b_hh ([B,3H]); gate order r,z,n.
 * lengths (optional, device int64 [B]) restates pack_padded_sequence (seq2seq_net.py:52-56): rows with t >= length keep
 * hprev and emit 0.  out (optional, row pitch ldout) receives the step output; saved (optional) = planes r,z,n,hn. */
int tg_gru_gates_fwd(const float* gi, long long ldgi, const float* gh, const float* hprev, const long long* lengths, int t,
                     float* hnew, float* out, long long ldout, float* saved, long long saved_plane, int B, int H, tg_stream stream);
/* Backward of tg_gru_gates_fwd: g = dh (carry, optional) + dadd (this step's output gradient, optional, row pitch ldadd);
 * writes dgi (row pitch lddgi), dgh [B,3H] and dhprev = g*z (the caller adds dgh @ W_hh); masked rows: zeros / carry. */
int tg_gru_gates_bwd(const float* dh, const float* dadd, long long ldadd, const float* saved, long long saved_plane,
                     const float* hprev, const long long* lengths, int t, float* dgi, long long lddgi, float* dgh,
                     float* dhprev, int B, int H, tg_stream stream);
/* Attn.forward (seq2seq_net.py:72-94) + context (:172-174) of one decoder step: hq = W_a[:, :H] h + b_a [B,H],
 * eproj = W_a[:, H:] enc [B,Tm,H] (computed once per forward), enc [B,Tm,H], v [H] -> w [B,Tm], ctx [B,H]. */
int tg_attn_fwd(const float* hq, const float* eproj, const float* enc, const float* v, float* w, float* ctx, int B, int Tm, int H,
                tg_stream stream);
int tg_attn_bwd(const float* dctx, const float* w, const float* hq, const float* eproj, const float* enc, const float* v,
                float* denc, float* deproj, float* dv, float* dhq, int B, int Tm, int H, tg_stream stream);
/* custom_loss (train_seq2seq.py:6-36): *loss += w_mse*mse + w_cont*continuity + w_var*variance term; dy_tmajor [T,B,D] =
 * gradient w.r.t. out [B,T,D] (row t = 0 is zero: frame 0 is a copy of the input, seq2seq_net.py:244-245). */
int tg_s2s_loss(const float* out, const float* target, double* loss, float* dy_tmajor, int B, int T, int D, float w_mse,
                float w_cont, float w_var, tg_stream stream);
/* xin[t,b,:] = decoder input of step t >= 1 (seq2seq_net.py:244-252): poses[b,t-1] while t-1 < n_pre, else outputs[b,t-1] */
int tg_s2s_gather_inputs(const float* poses, const float* outputs, float* xin, int B, int T, int D, int n_pre, tg_stream stream);
/* torch.nn.utils.clip_grad_norm_ (train_seq2seq.py:48) over a flat gradient arena: *out += sum x^2; x *= min(1, max/(norm+1e-6)) */
int tg_sumsq_f64(const float* x, long long n, double* out, tg_stream stream);
int tg_clip_scale(float* x, long long n, const double* sumsq, float max_norm, tg_stream stream);

/* Validation metrics of evaluate_testset (scripts/train.py:283,293-310; convert_dir_vec_to_pose, scripts/utils/data_utils.py:77-98)
 * for out / target [B,T,27] direction vectors: acc[0] += sum |out - target|, acc[1] += sum over frames >= n_pre of |joint position
 * error| (10 joints x 3), acc[2] += sum over frames >= 2 of |second time difference of the joint position error| (fp64 accumulators). */
int tg_pose_eval_metrics(const float* out, const float* target, int B, int T, int D, int n_pre, double* acc, tg_stream stream);

/* Input staging (scripts/train.py:171-176 `.to(device)` of a collated batch): copies nbytes from src to dst with at most max_ctas CTAs.
 * src may be PINNED HOST memory (read over PCIe through its mapped address): unlike cudaMemcpyAsync this uses no copy engine, so
 * copy-engine operations of a concurrently running step never queue behind the transfer.  16-byte aligned pointers. */
int tg_copy_bytes(void* dst, const void* src, long long nbytes, int max_ctas, tg_stream stream);

int tg_debug_gru_trace(long long* device_buf);
/* development aid: how many 8-CTA clusters of the tensor-core recurrence (forward kernel, batch tile BT in {16,32,48}, hidden size H) the
 * device keeps resident at once (cudaOccupancyMaxActiveClusters); a B200 is expected to answer 16 = one wave for 8 tiles x 2 directions */
int tg_debug_gru_cluster_occupancy(int H, int BT);
/* development aid: the batch tiles the cluster recurrence picks for B clips, packed as forward_tile * 1000 + backward_tile (-1: no plan) */
int tg_debug_gru_cluster_tiles(int B, int H);
/* development aid: %globaltimer stamps of CTA (0,0) of the next tg_gemm_tf32 launches (7 slots; NULL disables) */
int tg_debug_gemm_trace(long long* device_buf);
size_t tg_gru_bwd_tf32_scratch_floats(int B, int H);
int tg_gru_layer_fwd_tf32(const float* gi, const float* whh_f, const float* whh_r, const float* bhh_f, const float* bhh_r,
                          float* out, float* saved, long long saved_qstride, int* sync, int B, int T, int H, tg_stream stream);
/* tg_gru_layer_fwd_tf32 + the inter-layer dropout of nn.GRU(dropout=p) (multimodal_context_net.py:98-99): drop [B*T,2H] = out * mask, stored
 * by the recurrence kernel beside `out` (mask: already scaled keep-mask [B*T,2H]; both 16-byte aligned) */
int tg_gru_layer_fwd_tf32_drop(const float* gi, const float* whh_f, const float* whh_r, const float* bhh_f, const float* bhh_r,
                               float* out, float* saved, long long saved_qstride, const float* mask, float* drop, int* sync, int B, int T,
                               int H, tg_stream stream);
int tg_gru_layer_bwd_tf32(const float* dout, const float* out, const float* saved, long long saved_qstride,
                          const float* whhT_f, const float* whhT_r, float* dgi, float* dgh, float* partial, int* sync,
                          int B, int T, int H, tg_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * Losses of train_iter_gan (train_gan.py:41,53-56,67-82), forward value + gradient in one pass.
 * scalars (fp64, caller zeroes): [0] sum huber(out,target;0.1)  [1] sum_i div_i  [2] sum (1+lv-mu^2-e^lv)
 * d_out = w_reg/(B*T*D) * huber' + w_div/B * d div_i ; dmu/dlogvar = kld gradient * w_kld.
 * --------------------------------------------------------------------------------------------------------- */
int tg_gen_losses(const float* out, const float* target, const float* out_rand, const float* z, const float* z_rand,
                  const float* mu, const float* logvar, int B, int TD, int Z, float w_reg, float w_div, float w_kld,
                  double* scalars, float* d_out, float* dmu, float* dlogvar, tg_stream stream);
/* NS-GAN terms on sigmoid outputs p[n]: loss += -sum log(s*p + o + 1e-8)/n (s=+1,o=0 for "real"/generator,
 * s=-1,o=1 for "fake"); dlogit = w * d loss/d p * p(1-p).  scalar[0] += loss (fp64). */
int tg_bce_sigmoid(const float* p, int n, float s, float o, float w, double* scalar, float* dlogit, tg_stream stream);

/* torch.optim.Adam (train.py:104-109; no weight decay / amsgrad) over a flat arena.  step_dev: device int64 holding
 * the 1-based step count of THIS update (host bumps it or a graph-resident kernel does). grad_scale multiplies g. */
int tg_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                 float grad_scale, const long long* step_dev, tg_stream stream);
int tg_increment_i64(long long* x, long long by, tg_stream stream);

/* Philox-4x32-10 counter RNG: standard normals (reparameterize noise) and scaled Bernoulli keep-masks (dropout).
 * (seed, *offset_dev + stream_id) select the stream; offset_dev is a device int64 so CUDA-graph replays advance. */
int tg_philox_normal(float* out, long long n, unsigned long long seed, const long long* offset_dev, int stream_id,
                     tg_stream stream);
int tg_philox_dropout_mask(float* out, long long n, float p, unsigned long long seed, const long long* offset_dev,
                           int stream_id, tg_stream stream);
/* random permutation of 0..n-1 (torch.randperm, train_gan.py:62), n <= 2048, one CTA bitonic sort of Philox keys */
int tg_philox_randperm(long long* out, int n, unsigned long long seed, const long long* offset_dev, int stream_id,
                       tg_stream stream);
int tg_gather_i64(const long long* src, const long long* idx, long long* out, int n, tg_stream stream);

/* FGD sufficient statistics (embedding_space_evaluator.py:79-80): acc[0]+=n, acc[1..1+F)+=sum x, acc[1+F..)+=sum x x^T */
int tg_feature_stats_f64(const float* feat, long long n, int F, double* acc, tg_stream stream);
/* sum_i sum_f |a-b| accumulated in fp64 (feat_dist, embedding_space_evaluator.py:95-99) */
int tg_l1_dist_f64(const float* a, const float* b, long long n, double* acc, tg_stream stream);

/* ---------------------------------------------------------------------------------------------------------
 * FGD auto-encoder trainer (train_feature_extractor.py:54-97 train_iter, train_joint_embed.py:5-65 train_iter_embed /
 * eval_embed on EmbeddingNet(mode='pose')).  The networks themselves run on the conv / BatchNorm / Adam entries above.
 * --------------------------------------------------------------------------------------------------------- */
/* L1 reconstruction loss, one CTA per clip (train_feature_extractor.py:64-72, train_joint_embed.py:21-29,59-61):
 *   l0_b = mean_{t,d} |recon - target|,  l1_b = mean_{t<T-1,d} |(recon[t+1]-recon[t]) - (target[t+1]-target[t])|
 *   acc[0] += sum_b (l0_b + use_diff*l1_b)  (the trainer's recon_loss),  acc[1] += sum_b l0_b  (eval_embed's numerator), fp64;
 *   d_recon (may be NULL) = weight * d acc[0] / d recon, with d|x|/dx = sign(x) (0 at 0) like torch. */
int tg_ae_recon_loss(const float* recon, const float* target, int B, int T, int D, int use_diff, float weight, double* acc,
                     float* d_recon, tg_stream stream);
/* out[b][c][r] = in[b][r][c]: channels-last [B,T,C] <-> the reference's channel-major flatten / view (embedding_net.py:71,213) */
int tg_transpose_batched_f32(const float* in, float* out, int B, int R, int C, tg_stream stream);

#ifdef __cplusplus
}
#endif
#endif
