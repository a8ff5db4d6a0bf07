/* x2i_b200 -- C ABI of the B200-native (sm_100a) X2I hot path.
 *
 * The reference (OPPO-Mente-Lab/X2I) is pure Python and has no FFI; its extension points for this path are
 * Python object protocols.  This header is the boundary the native library exports UNDER those
 * protocols: x2i_b200/ops.py binds every entry point below with ctypes (see INTEGRATION.md for the stub), and
 * x2i_b200/{flux,pipeline,proj,kd}.py mirror the reference interfaces on top.  Each entry cites the reference
 * code it replaces (paths relative to the reference repo; "[D031]" = diffusers==0.31.0, un-vendored).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; tensors are bf16 unless noted;
 *   - the caller owns all buffers (outputs, workspaces); nothing is allocated, nothing synchronises;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 on success, negative X2I_ERR_* otherwise; x2i_last_error() describes the last failure
 *     of the calling thread;
 *   - thread-safe; callable from any host thread (the reference does GPU work off-thread,
 *     core/data/dataloader.py:100-123).
 */
#ifndef X2I_B200_H_
#define X2I_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define X2I_OK 0
#define X2I_ERR_SHAPE (-1)  /* unsupported / inconsistent dimensions */
#define X2I_ERR_ALIGN (-2)  /* pointer or leading dimension not 16-byte aligned */
#define X2I_ERR_ARCH (-3)   /* device is not sm_100 */
#define X2I_ERR_LAUNCH (-4) /* CUDA runtime / driver error */

int x2i_version(void);
const char* x2i_last_error(void);
/* number of kernels launched through this library by this process (bench.py's gpu_launches) */
long long x2i_launch_count(void);

/* ---- dense contractions (tcgen05.mma, TMA, TMEM accumulators) -------------------------------------------------
 * C[M,N] = act(A[M,K] @ W[N,K]^T + bias[N]);  A, W row-major with leading dimensions lda, ldw (elements).
 * act: 0 none, 1 GELU(tanh), 2 GELU(erf).  Replaces nn.Linear (+ F.gelu) at:
 *   FeedForward.net[0] [D031] called lightcontrol/lightcontrol_flux.py:186,199; proj_mlp+act_mlp :90;
 *   x_embedder :445; context_embedder :457; proj_out :543; MLP3.projector / fc utils/proj.py:30-31.            */
int x2i_gemm_bias_act(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C, int64_t ldc,
                      int M, int N, int K, int act, void* stream);

/* C = A @ W^T + bias and C_gelu = GELU_erf(C) in one pass (two outputs).  MLP3: x2 = projector(x) is returned as the
 * prompt embedding while fc = Sequential(GELU, Linear) consumes GELU(x2) (utils/proj.py:22-25, :30-31).               */
int x2i_gemm_bias_dual(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C, int64_t ldc,
                       void* C_gelu, int64_t ldg, int M, int N, int K, void* stream);

/* C[M,N] = residual[M,N] + gate[m / rows_per_batch, :] * (A @ W^T + bias);  aux (nullable) receives the un-gated
 * A @ W^T + bias -- the tensor the reference's forward hooks capture (train/train_qwenvl.py:186-214).
 * Replaces to_out[0] / to_add_out / ff.net[2] / proj_out followed by `gate.unsqueeze(1) * y` and the residual add:
 * lightcontrol/lightcontrol_flux.py:97-100, :180-181, :186-189, :193-200.  C may alias residual.                  */
int x2i_gemm_gate_residual(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* gate,
                           int64_t gate_stride, int rows_per_batch, const void* residual, int64_t ldr, void* C,
                           int64_t ldc, void* aux, int64_t ldaux, int M, int N, int K, void* stream);

/* Fused QKV (+ single-block proj_mlp) projection.  W = [Wq; Wk; Wv; (Wmlp)] is [N, K] with N = 3*heads*128 (+ F).
 * (N may also stop after the q or k section; a q/k section whose rms weight is null is stored plain, like v -- used for
 * the Resampler's in-projections.)
 * Epilogue per 128-column head: q,k: +bias, RMSNorm(128, eps) * rms_{q,k}, RoPE (adjacent pairs, table
 * rope[L_total,64] of (cos,sin) from x2i_rope_table), stored head-major into q/k[B, heads, L_total, 128] at token
 * row_offset + (m % rows_per_batch); v: +bias, same layout; columns >= 3*heads*128: GELU(tanh) -> mlp[m*ldmlp + ..].
 * Replaces to_q/to_k/to_v/add_*_proj, norm_q/k, norm_added_q/k, the txt||img concat and apply_rotary_emb of
 * FluxAttnProcessor2_0 [D031] (call sites lightcontrol_flux.py:92-95, :173-177) and proj_mlp :90.   */
int x2i_gemm_qkv_rope(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* rms_q,
                      const void* rms_k, const void* rope /* float2, nullable */, void* q, void* k, void* v, void* mlp,
                      int64_t ldmlp, int M, int N, int K, int heads, int rows_per_batch, int row_offset, int L_total,
                      float eps, void* stream);

/* Grouped launch: up to two GEMMs of the SAME kind / N-tile shape in ONE persistent kernel (e.g. the image and the text
 * stream of a FluxTransformerBlock -- different weights, M = B*4096 and B*512 -- lightcontrol_flux.py:166-200), so the
 * small text problem fills the partial wave of the image problem instead of occupying the GPU alone.  Fields have the
 * meaning of the same-named arguments of x2i_gemm_bias_act / x2i_gemm_gate_residual / x2i_gemm_qkv_rope; unused ones
 * are zero.  Falls back to independent launches when a problem does not fit the CTA-pair kernel (N % 256, M <= 128).  */
#define X2I_GEMM_BIAS_ACT 0
#define X2I_GEMM_GATE_RESIDUAL 1
#define X2I_GEMM_QKV_ROPE 2
typedef struct x2i_gemm_desc {
  int kind, act;
  int M, N, K;
  int rows_per_batch, heads, row_offset, L_total;
  float eps;
  const void *A, *W, *bias;
  int64_t lda, ldw;
  void* C;
  int64_t ldc;
  const void *gate, *residual;
  int64_t gate_stride, ldr;
  void* aux;
  int64_t ldaux;
  const void *rms_q, *rms_k, *rope;
  void *q, *k, *v, *mlp;
  int64_t ldmlp;
  /* training-mode saves (all nullable; see x2i_gemm_qkv_rope_save / x2i_gemm_bias_act_save): */
  void* qk_pre;   /* QKV: pre-norm q|k, token-major [M, ldqk] */
  int64_t ldqk;
  void* mlp_pre;  /* QKV: pre-GELU proj_mlp values [M, ldmlp_pre] */
  int64_t ldmlp_pre;
  int aux_act;    /* BIAS_ACT with act = 0 and aux set: aux = aux_act(C), 1 GELU(tanh), 0/2 GELU(erf) */
} x2i_gemm_desc;
int x2i_gemm_grouped(const x2i_gemm_desc* descs, int n, void* stream);

/* C[M,N] = A[M,K] @ Bkn[K,N] (+bias): B given N-contiguous ("MN-major" tcgen05 operand, as V is in attention).     */
int x2i_gemm_kn(const void* A, int64_t lda, const void* Bkn, int64_t ldb, const void* bias, void* C, int64_t ldc,
                int M, int N, int K, void* stream);

/* ---- dense contractions of the backward pass (distillation training, train/train_qwenvl.py:625) -----------------
 * The FLUX weights are frozen (requires_grad_(False), train_qwenvl.py:417-429), so the student pass needs only
 * activation gradients (dgrad) through the 57 blocks; weight gradients (wgrad) exist for the projector alone.
 *
 * dgrad of y = x W^T:   dX[M, Kin] = addend + (dY[M, Nout] @ W[Nout, Kin]) * act'(pre[M, Kin - n_split])
 * W is consumed exactly as nn.Linear stores it (as the N-contiguous tcgen05 B operand; nothing is transposed in memory).
 * pre (nullable) holds the pre-activation of the Linear+GELU that PRODUCED x: columns >= n_split are multiplied by
 * GELU'(pre) (dact 1 tanh -- FeedForward / proj_mlp, 2 erf -- MLP3), columns < n_split pass through (the attention part
 * of a single block's [attn | mlp] concat, lightcontrol_flux.py:97).  addend (nullable) is added last; may alias dX.      */
int x2i_gemm_dgrad(const void* dY, int64_t lddy, const void* W, int64_t ldw, const void* pre, int64_t ldpre, int n_split,
                   int dact, const void* addend, int64_t ldadd, void* dX, int64_t lddx, int M, int Nout, int Kin, void* stream);

/* wgrad of y = x W^T:   dW[N, K] (+)= dY[M, N]^T @ X[M, K]   (contraction over the M token rows; both operands are read
 * as MN-major tcgen05 operands).  Projector linears: utils/proj.py:22-25 trained at train_qwenvl.py:453-459.            */
int x2i_gemm_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, void* dW, int64_t lddw, int M, int N, int K,
                   int accumulate, void* stream);
/* Same, with split-K when the output has only a handful of tiles (the convolution weight gradients of the LightControl trainer:
 * N = Cout, K = KH*KW*Cin, M = pixels): the k-blocks are dealt to several tile groups whose fp32 partials (workspace) are then
 * added in a fixed order -- deterministic.  workspace: x2i_gemm_wgrad_workspace_floats(M, N, K) floats (0 = no split needed).   */
int x2i_gemm_wgrad_splitk(const void* dY, int64_t lddy, const void* X, int64_t ldx, void* dW, int64_t lddw, int M, int N, int K,
                          int accumulate, float* workspace, int64_t workspace_floats, void* stream);
int64_t x2i_gemm_wgrad_workspace_floats(int M, int N, int K);

/* Forward Linear + GELU that keeps the pre-activation for the backward: C_pre = A W^T + bias, C_act = act(C_pre);
 * act 1 GELU(tanh), 2 GELU(erf).                                                                                    */
int x2i_gemm_bias_act_save(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C_pre, int64_t ldc,
                           void* C_act, int64_t ldg, int M, int N, int K, int act, void* stream);

/* ---- fused MMDiT attention ----------------------------------------------------------------------------------
 * O = softmax(Q K^T / sqrt(128)) V over q,k,v[B, heads, L, 128]; output token-major: rows with token < split go to
 * out0[(b*split + t) * ld0 + h*128], the rest to out1[(b*(L-split) + t-split) * ld1 + h*128].
 * Replaces F.scaled_dot_product_attention + transpose/reshape + the txt/img split in FluxAttnProcessor2_0 [D031].  */
int x2i_mmdit_attention(const void* q, const void* k, const void* v, void* out0, int64_t ld0, int split, void* out1,
                        int64_t ld1, int B, int heads, int L, void* stream);

/* Training forward: same as x2i_mmdit_attention, additionally lse fp32 [B*heads, Lpad] (Lpad = L rounded up to 128):
 * log2-domain log-sum-exp of every score row (+inf in the padding), consumed by x2i_mmdit_attention_bwd.             */
int x2i_mmdit_attention_lse(const void* q, const void* k, const void* v, void* out0, int64_t ld0, int split, void* out1,
                            int64_t ld1, float* lse, int B, int heads, int L, void* stream);

/* Backward of the fused attention (autograd of F.scaled_dot_product_attention in FluxAttnProcessor2_0 [D031]):
 * q,k,v,dout,dq,dk,dv head-major [B, heads, L, 128] bf16; lse from the forward, delta from x2i_attention_bwd_prep.
 * Two launches (dK/dV per key tile, dQ per query tile), no atomics, bit-reproducible.                               */
int x2i_mmdit_attention_bwd(const void* q, const void* k, const void* v, const void* dout, const float* lse, const float* delta,
                            void* dq, void* dk, void* dv, int B, int heads, int L, void* stream);

/* Prologue of the attention backward: token-major dO (+ optional addend = the KD-loss gradient of a hooked single-block
 * attention output, train_qwenvl.py:214) and token-major O -> head-major dO [B,heads,L,128] and
 * delta[b,h,t] = sum_d dO*O ([B*heads, Lpad] fp32).  Token-major tensors are split like the forward outputs: tokens
 * t < split in *0 (row (b*split + t) * ld), the rest in *1 (row (b*(L-split) + t-split) * ld).                      */
int x2i_attention_bwd_prep(const void* do0, int64_t lddo0, const void* do1, int64_t lddo1, const void* o0, int64_t ldo0,
                           const void* o1, int64_t ldo1, const void* add0, int64_t ldadd0, const void* add1, int64_t ldadd1,
                           void* do_hm, float* delta, int B, int heads, int L, int split, void* stream);

/* Backward of the QKV epilogue (per-head RMSNorm(128)*w then RoPE) + head-major -> token-major: dq,dk,dv
 * [B,heads,L_total,128] -> out[M, ldo] = [dq_pre | dk_pre | dv] for the M = B*rows_per_batch tokens at row_offset;
 * qk_pre = the pre-norm q|k saved by x2i_gemm_qkv_rope_save.  (norm_q/norm_k + apply_rotary_emb)   */
int x2i_qk_norm_rope_bwd(const void* dq, const void* dk, const void* dv, const void* qk_pre, int64_t ldqk, const void* rms_q,
                         const void* rms_k, const void* rope, void* out, int64_t ldo, int M, int heads, int rows_per_batch,
                         int row_offset, int L_total, float eps, void* stream);

/* x2i_gemm_qkv_rope that also stores what the backward needs: qk_pre [M, ldqk] (pre-norm q|k) and mlp_pre [M, ldmlp_pre]
 * (pre-GELU proj_mlp; nullable when N == 3*heads*128).                                                                */
int x2i_gemm_qkv_rope_save(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* rms_q,
                           const void* rms_k, const void* rope, void* q, void* k, void* v, void* mlp, int64_t ldmlp,
                           void* qk_pre, int64_t ldqk, void* mlp_pre, int64_t ldmlp_pre, int M, int N, int K, int heads,
                           int rows_per_batch, int row_offset, int L_total, float eps, void* stream);

/* Same kernel with different query / key lengths and an optional key-padding mask: q[B,heads,L,128],
 * k,v[B,heads,Lkv,128]; kv_len (device int32 [B], nullable) = number of valid keys of batch b.  Replaces the
 * cross-attention core of the MiniCPM-o Resampler (minicpm/resampler.py:170-176; MultiheadAttention :406-668 with
 * key_padding_mask, softmax(q k^T / sqrt(128) + mask) v).                                                          */
int x2i_cross_attention(const void* q, const void* k, const void* v, const int* kv_len, void* out0, int64_t ld0, int split,
                        void* out1, int64_t ld1, int B, int heads, int L, int Lkv, void* stream);

/* ---- row-wise HBM-bound kernels -----------------------------------------------------------------------------
 * y = LayerNorm(x; no affine, eps) * (1 + scale[b]) + shift[b], b = row / rows_per_batch.  AdaLayerNormZero /
 * ZeroSingle / Continuous [D031] and norm2 + modulate (lightcontrol_flux.py:89, :166-170, :183-184, :196-197, :542). */
int x2i_ln_modulate(const void* x, int64_t ldx, const void* scale, const void* shift, int64_t mod_stride, void* y,
                    int64_t ldy, int rows, int D, int rows_per_batch, float eps, void* stream);
/* The same for TWO row segments in one launch: the image and the text stream of a double block (FluxTransformerBlock.norm1 / norm1_context
 * and the two norm2 + modulate steps, lightcontrol_flux.py:140-146,:171-179) live in separate buffers with their own modulation rows. */
int x2i_ln_modulate2(const void* x0, int64_t ldx0, const void* scale0, const void* shift0, int64_t mod_stride0, void* y0, int64_t ldy0, int rows0,
                     int rows_per_batch0, const void* x1, int64_t ldx1, const void* scale1, const void* shift1, int64_t mod_stride1, void* y1,
                     int64_t ldy1, int rows1, int rows_per_batch1, int D, float eps, void* stream);

/* y = LayerNorm(x) * gamma + beta (nn.LayerNorm with affine): Resampler ln_q / ln_kv / ln_post (minicpm/resampler.py:
 * 116-119, :166-168, :184).                                                                                         */
int x2i_layernorm_affine(const void* x, int64_t ldx, const void* gamma, const void* beta, void* y, int64_t ldy, int rows, int D,
                         float eps, void* stream);

/* out[b,l,:] = x[b,l,:] + pos[l / w_b, l % w_b, :] for l < h_b*w_b else x[b,l,:]; tgt_sizes int32 [B,2] = (h_b, w_b), pos
 * bf16 [max_h,max_w,D]: the per-image slice / flatten / zero-pad of the Resampler's 2-D sincos table that is added to
 * the keys (minicpm/resampler.py:156-173).                                                                          */
int x2i_add_pos2d(const void* x, const void* pos, const int* tgt_sizes, void* out, int B, int L, int D, int max_h, int max_w,
                  void* stream);

/* x[r,:] += gate[r / rows_per_batch, :] * y[r,:] -- the un-fused `gate.unsqueeze(1) * attn_output` + residual of
 * lightcontrol_flux.py:180-181, :193-194, used only behind a plug-in attention processor.                           */
int x2i_gate_residual(void* x, int64_t ldx, const void* y, int64_t ldy, const void* gate, int64_t gate_stride, int rows,
                      int D, int rows_per_batch, void* stream);

/* ---- row-wise backward kernels (distillation training) --------------------------------------------------------
 * dy = gate[b] * dx (+ addend): backward of x' = x + gate * y towards y (lightcontrol_flux.py:100, :180-181, :187-189);
 * addend = KD-loss gradient arriving at the hooked y.                                                               */
int x2i_gate_bwd(const void* dx, int64_t lddx, const void* gate, int64_t gate_stride, const void* addend, int64_t ldadd, void* dy,
                 int64_t lddy, int rows, int D, int rows_per_batch, void* stream);

/* dx = dres + dLN: backward of x2i_ln_modulate (affine = 0, `scale` = the AdaLN scale) or x2i_layernorm_affine (affine = 1,
 * `scale` = gamma, mod_stride 0) towards x; stats (nullable) float2 [rows] receives (mean, rstd) for x2i_colsum.      */
int x2i_ln_modulate_bwd(const void* dn, int64_t lddn, const void* x, int64_t ldx, const void* scale, int64_t mod_stride,
                        const void* dres, int64_t ldr, void* dx, int64_t lddx, void* stats, int rows, int D, int rows_per_batch,
                        float eps, int affine, void* stream);

/* Per-batch column sums (deterministic two-stage): out0[b, :] (+)= sum_t A[b,t,:];  out1[b, :] (+)= sum_t A[b,t,:] * Bv[b,t,:]
 * with Bv = B or, when stats is given, (B - mean_t) * rstd_t.  fp32 outputs with row strides ldo0/ldo1; workspace of
 * x2i_colsum_workspace_floats() floats.  dshift/dscale/dgate of the AdaLN modulations, dgamma/dbeta, bias gradients.  */
int x2i_colsum(const void* A, int64_t lda, const void* Bm, int64_t ldb, const void* stats, float* out0, int64_t ldo0, float* out1,
               int64_t ldo1, float* workspace, int nbatch, int rows_per_batch, int D, int accumulate, void* stream);
int64_t x2i_colsum_workspace_floats(int nbatch, int rows_per_batch, int D);

/* out[b,k] (+)= act'(pre[b,k]) * sum_n g[b,n] W[n,k]  (g, out fp32; dact 0 none, 1 SiLU'): the backward of
 * x2i_skinny_linear -- all AdaLN modulation linears of a step in one pass over the concatenated weights, and the
 * time/text embedding MLPs.  workspace: x2i_skinny_linear_t_workspace_floats() floats.                              */
int x2i_skinny_linear_t(const float* g, int64_t ldg, const void* W, int64_t ldw, const void* pre, int64_t ldpre, float* out,
                        int64_t ldo, float* workspace, int B, int N, int K, int dact, int accumulate, void* stream);
int64_t x2i_skinny_linear_t_workspace_floats(int N, int K);

int x2i_f32_to_bf16(const float* in, void* out, int64_t n, void* stream);

/* out[b,n] (+)= bias[n] + sum_k act_in(x[b,k]) W[n,k];  B <= 64, act_in: 0 none, 1 SiLU.  All AdaLN modulation
 * linears of a step in one launch (weights concatenated), and the CombinedTimestep*TextProjEmbeddings MLPs [D031]. */
int x2i_skinny_linear(const void* x, int64_t ldx, const void* W, int64_t ldw, const void* bias, void* out, int64_t ldo,
                      int B, int N, int K, int act_in, int accumulate, void* stream);

/* Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0) [D031]: t fp32 [B] -> out bf16 [B, dim].            */
int x2i_timestep_sinusoid(const float* t, void* out, int B, int dim, void* stream);

/* FluxPosEmbed [D031] (lightcontrol_flux.py:247,:472): ids fp32 [L,3] -> cos,sin fp32 [L,a0+a1+a2] (nullable) and the
 * compact table rope float2 [L,(a0+a1+a2)/2] (nullable).                                                          */
int x2i_rope_table(const float* ids, int L, int a0, int a1, int a2, double theta, float* cos_out, float* sin_out,
                   void* rope, void* stream);

/* FlowMatchEulerDiscreteScheduler.step [D031]: x <- bf16(float(x) + dsigma * v), n elements (multiple of 8).        */
int x2i_euler_step(void* x, const void* v, float dsigma, int64_t n, void* stream);

/* ---- attention-distillation loss (train/train_qwenvl.py:58-61, :601-620) --------------------------------------
 * teacher/student: [rows, D] bf16.  Rows are grouped in SEGMENTS (contiguous runs belonging to one layer and one
 * batch element): segment s owns rows [seg_row_start[s], seg_row_start[s+1]) (device int64 [n_seg+1]) and belongs to
 * layer seg_layer[s] (device int32 [n_seg]); this covers both per-layer tensors and the reference's stacked
 * [B, n_layers, L, D] hook tensors (:590-592) without a copy.
 *   term_l = sum over the layer's rows of KL(softmax(norm(s)/T) || softmax(norm(t)/T)) / batch    (F.kl_div batchmean)
 *   loss = sum of finite term_l; valid[l] = isfinite(term_l)  (the reference's inf/nan guard, :606-609)
 * Workspaces: row_kl fp32 [rows], seg_sum fp64 [n_seg].  Outputs: layer_term fp32 [n_layers], loss fp32 [1], valid.  */
int x2i_kd_loss_fwd(const void* teacher, const void* student, int64_t rows, int D, float temperature,
                    const int64_t* seg_row_start, const int* seg_layer, int n_seg, int n_layers, int batch, float* row_kl,
                    double* seg_sum, float* layer_term, float* loss, int* valid, void* stream);
/* grad_student[rows, D] bf16 = dloss * d loss / d student (rows of invalid layers get 0); row_scale: fp32 [rows]
 * workspace; max_seg_rows = longest segment.  Recomputes the row statistics (no saved activations).                */
int x2i_kd_loss_bwd(const void* teacher, const void* student, int64_t rows, int D, float temperature,
                    const int64_t* seg_row_start, const int* seg_layer, int n_seg, int64_t max_seg_rows, int batch,
                    const int* valid, const float* dloss, float* row_scale, void* grad_student, void* stream);

/* ---- alignment projector front end (utils/proj.py:62-72, :29) --------------------------------------------------
 * y[b,s,:] = LayerNorm_H(mix_c x[b,c,s,:]) * gamma + beta;  mode 0: Conv2d(C->1, 5x5, pad 2) weights w fp32 [C,5,5]
 * + conv_bias; mode 1: mean_c(w[c] * x) (cha_scale); mode 2: mean_c(x).  x bf16 [B,C,S,H], y bf16 [B,S,H].         */
int x2i_proj_mix_ln(const void* x, int mode, const float* w, float conv_bias, const float* gamma, const float* beta,
                    float eps, void* y, int B, int C, int S, int H, void* stream);
/* pooled[b,n] = mean_s y[b,s,n]  (utils/proj.py:32)                                                                */
int x2i_mean_over_s(const void* y, void* out, int B, int S, int N, void* stream);

/* ---- projector backward (the only trained module: train/train_qwenvl.py:453-459) --------------------------------
 * x2i_proj_mix_ln that also stores xm bf16 [B,S,H], the mixed plane BEFORE the LayerNorm (needed by its backward).    */
int x2i_proj_mix_ln_save(const void* x, int mode, const float* w, float conv_bias, const float* gamma, const float* beta,
                         float eps, void* y, void* xm, int B, int C, int S, int H, void* stream);
/* Mode 0 (the 5x5 layer-mixing convolution, utils/proj.py:66-70) on the tensor pipe: the taps along H are a banded-Toeplitz B operand built
 * in shared memory, the five row shifts re-read one TMA box through descriptor row offsets, fp32 accumulation in TMEM; then LayerNorm over H.
 * Same results as x2i_proj_mix_ln(mode 0) up to fp32 summation order (bf16 taps: the reference's projector is bf16).  xm (nullable) receives
 * the pre-LayerNorm plane in bf16.  workspace: _workspace_floats() floats.  _supported(): S % 128 == 0, C <= 40, 512 <= H <= 4096, H % 8 == 0. */
int64_t x2i_proj_mix_ln_tc_supported(int B, int C, int S, int H);
int64_t x2i_proj_mix_ln_tc_workspace_floats(int B, int C, int S, int H);
int x2i_proj_mix_ln_tc(const void* x, const float* w, float conv_bias, const float* gamma, const float* beta, float eps, void* y, void* xm,
                       float* workspace, int B, int C, int S, int H, void* stream);
/* dy[b,s,:] = dpooled[b,:] / S : backward of the mean over S (utils/proj.py:32).                                    */
int x2i_mean_over_s_bwd(const void* dpooled, void* dy, int B, int S, int N, void* stream);
/* Weight gradient of the layer-mixing front end w.r.t. g = d loss / d mixed plane [B,S,H] (bf16):
 * mode 0: dw fp32 [C,5,5] of Conv2d(C->1, 5x5, pad 2) (utils/proj.py:69; the conv bias gradient is sum(g));
 * mode 1: dw fp32 [C] of cha_scale (utils/proj.py:67, includes the 1/C of the mean).
 * workspace: x2i_proj_mix_wgrad_workspace_floats() floats.  Deterministic two-stage reduction.                      */
int x2i_proj_mix_wgrad(const void* x, const void* g, int mode, float* dw, float* workspace, int B, int C, int S, int H, void* stream);
int64_t x2i_proj_mix_wgrad_workspace_floats(int B, int C, int S);

/* ---- ControlNeXt nets of the LightControl editing branch (lightcontrol/lightcontrol_flux.py:575-749) -------------------
 * Implicit-GEMM convolution on tcgen05, NHWC bf16: out[n,y,x,co] = relu?(conv(x, w)[..] + bias[co] + rowvec[n,co]) + residual.
 * x [N,H,W,Cin], w pre-packed [Cout, KH, KW, Cin], out / residual [N,Ho,Wo,Cout]; KH,KW <= 3, stride 1 or 2, pad 0 or 1,
 * Cin and Cout multiples of 64.  No im2col buffer: the A operand of each (tap, channel-chunk) k-block is one shifted
 * tensor-map box of the input, out-of-image rows zero-filled by TMA.  Replaces nn.Conv2d (cuDNN) in ControlNeXtModel.embedding
 * [3], [6] (:596-601), ResnetBlock2D.conv1/conv2/conv_shortcut and Downsample2D.conv [D031] (:605-624), mid_convs (:626-668);
 * rowvec = time_emb_proj(silu(temb)) of ResnetBlock2D; residual = shortcut / the FLUX hidden states for the final conv
 * (the injection hidden_states += out * scale of :505-507 fused into its epilogue; out may alias residual).             */
int x2i_conv2d_nhwc(const void* x, const void* w, const void* bias, const void* rowvec, int64_t rowvec_stride, const void* residual,
                    void* out, int Nimg, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int relu, void* stream);
/* Same with `groups` weight sets: image n uses w[n / (Nimg / groups)] (w [groups, Cout, KH*KW*Cin], bias [groups, Cout]) -- the 19
 * ControlNeXt nets of a LightControl step (same layer shapes, different weights) as ONE launch per layer -- and with separate
 * leading (top/left: `pad`) and trailing (bottom/right: `pad_end`) zero padding: the VAE encoder's Downsample2D pads (0,1,0,1)
 * before its stride-2 conv.                                                                                               */
int x2i_conv2d_nhwc_grouped(const void* x, const void* w, const void* bias, const void* rowvec, int64_t rowvec_stride, const void* residual,
                            void* out, int Nimg, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int pad_end, int relu,
                            int groups, void* stream);
/* ControlNeXtModel.embedding[0]: Conv2d(3 -> 64, 3x3, stride 2, pad 1) on the NCHW bf16 hint image -> NHWC bf16 [N,H/2,W/2,64];
 * w fp32 [64,3,3,3] (PyTorch layout), bias fp32 [64].                                                                  */
int x2i_conv_first(const void* x, const float* w, const float* bias, void* out, int Nimg, int H, int W, void* stream);
/* `groups` stems on the same hint: w [groups,64,3,3,3], bias [groups,64] -> out [groups, N, H/2, W/2, 64].                      */
int x2i_conv_first_grouped(const void* x, const float* w, const float* bias, void* out, int Nimg, int H, int W, int groups, void* stream);
/* nn.GroupNorm on NHWC bf16 + activation (0 none, 1 ReLU, 2 SiLU) + optional residual add: y = act(GN(x)) + residual.
 * gamma/beta bf16 [C]; C a power of two in [64, 2048], groups of 4 or a multiple of 8 channels; workspace of
 * x2i_groupnorm_workspace_floats() floats.  Deterministic.                                                                */
int x2i_groupnorm_nhwc(const void* x, const void* gamma, const void* beta, const void* residual, void* y, float* workspace, int Nimg,
                       int HW, int C, int G, float eps, int act, void* stream);
int64_t x2i_groupnorm_workspace_floats(int Nimg, int HW, int G);
/* Same with `param_sets` (gamma, beta) pairs [param_sets, C]: image n uses set n / (Nimg / param_sets).                        */
int x2i_groupnorm_nhwc_grouped(const void* x, const void* gamma, const void* beta, const void* residual, void* y, float* workspace, int Nimg,
                               int HW, int C, int G, float eps, int act, int param_sets, void* stream);

/* ---- LightControl trainer building blocks (lightcontrol/train_lightcontrol.py:672-775: the ControlNeXt nets are
 * the trainable part; the trainer is x2i_b200/train_lightcontrol.py) --------------------------------------------------------------
 * Backward of y = act(GroupNorm(x)): dx (bf16), dgamma / dbeta (fp32 [C], overwritten or accumulated); statistics are recomputed
 * from x.  Groups of a multiple of 8 channels.  Deterministic.  workspace: x2i_groupnorm_bwd_workspace_floats() floats.        */
int x2i_groupnorm_nhwc_bwd(const void* x, const void* dy, const void* gamma, const void* beta, void* dx, float* dgamma, float* dbeta,
                           float* workspace, int Nimg, int HW, int C, int G, float eps, int act, int accumulate, void* stream);
int64_t x2i_groupnorm_bwd_workspace_floats(int Nimg, int HW, int C, int G);
/* cols[(n, yo, xo), (ky, kx, ci)] = x[n, yo*stride + ky - pad, xo*stride + kx - pad, ci] (zero outside): the explicit operand of the
 * convolution weight gradient dW[Cout, KH*KW*Cin] = dY^T cols (x2i_gemm_wgrad).  cols: bf16 [Nimg*Ho*Wo, KH*KW*C].              */
/* dx = dy where y > 0 else 0 (backward of a ReLU fused into a conv epilogue, masked by the saved output); out = silu(x).      */
int x2i_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, void* stream);
int x2i_silu(const void* x, void* out, int64_t n, void* stream);
int x2i_im2col_nhwc(const void* x, void* cols, int Nimg, int H, int W, int C, int KH, int KW, int stride, int pad, int pad_end, void* stream);
/* Implicit convolution weight gradient (no im2col buffer): dW[Cout, KH*KW*Cin] (bf16, the packed layout above) (+)= sum over ALL images and
 * output pixels of dY[pix, Cout] x X[pix * stride + tap - pad, Cin], one launch (+ the fixed-order split-K reduction).  The tcgen05 GEMM's B
 * operand is loaded as 64-pixel x 64-channel boxes of x itself, shifted by the tap (stride 2: a 5-D parity view of x), zero padding = TMA
 * out-of-bounds fill.  Replaces torch autograd's conv weight gradient behind the reference's ControlNeXt nets
 * (lightcontrol/train_lightcontrol.py:760 `accelerator.backward(loss)`).  x [Nimg,H,W,Cin], dy [Nimg,Ho,Wo,Cout] contiguous NHWC bf16.
 * _supported(): 1 when the output rows tile into 64-pixel blocks (Wo % 64 == 0, or 64 % Wo == 0 with Ho*Wo % 64 == 0), Cin % 64 == 0,
 * kernel <= 3x3, stride 1 or 2; otherwise use x2i_im2col_nhwc + x2i_gemm_wgrad.  workspace: _workspace_floats() floats (may be 0).    */
int64_t x2i_conv2d_nhwc_wgrad_supported(int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int pad_end);
int64_t x2i_conv2d_nhwc_wgrad_workspace_floats(int Nimg, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, int pad_end);
int x2i_conv2d_nhwc_wgrad(const void* x, const void* dy, void* dw, float* workspace, int64_t workspace_floats, int Nimg, int H, int W, int Cin,
                          int Cout, int KH, int KW, int stride, int pad, int pad_end, int accumulate, void* stream);

/* ---- VAE decoder (reference call site infer/inference_qwenvl.py:209-216: vae.decode(latents)) ----------
 * The decoder's convolutions and GroupNorms run through x2i_conv2d_nhwc / x2i_groupnorm_nhwc (C up to 2048, groups of 4 or a
 * multiple of 8 channels); these three cover what is left of diffusers' AutoencoderKL decoder [D031].
 * C32[M, ldc] (fp32) = alpha * A[M,K] W[N,K]^T: the attention scores of the single-head d=512 mid-block attention.           */
int x2i_gemm_f32(const void* A, int64_t lda, const void* W, int64_t ldw, float* C32, int64_t ldc, int M, int N, int K, float alpha,
                 void* stream);
/* P[r, :] (bf16) = softmax(S[r, :]) over fp32 scores; cols % 4 == 0, cols <= 16384.                                           */
int x2i_softmax_rows(const float* S, int64_t lds, void* P, int64_t ldp, int rows, int cols, void* stream);
/* Upsample2D's F.interpolate(scale_factor=2, mode="nearest") on NHWC bf16: out [N, 2H, 2W, C].                              */
int x2i_upsample2x_nhwc(const void* x, void* out, int Nimg, int H, int W, int C, void* stream);

/* ---- MLLM prefill with all-layer hidden-state capture --------------------------------------------------
 * The producer of the projector's input: the reference runs the MLLM's decoder stack once over the padded prompt and keeps every
 * layer's hidden state -- `qwen_encoder.generate(**inputs, output_hidden_states=True, return_dict_in_generate=True)` then
 * `torch.cat(output_hidden_state["hidden_states"][0]).unsqueeze(0)` (infer/inference_qwenvl.py:176-179,:121-132) /
 * `torch.stack(generated_ids["hidden_states"][0], dim=1)` (train/train_qwenvl.py:773-775).  The model code is the third-party
 * `transformers` Qwen2.5-VL text decoder (Qwen2_5_VLTextModel: RMSNorm, grouped-query attention with rotate-half RoPE, SwiGLU MLP).
 * Here every layer writes its output straight into its slot of the projector's [B, C, S, H] input; the contractions run on
 * x2i_gemm_bias_act / x2i_gemm_swiglu / x2i_gemm_gate_residual (gate = NULL: plain residual).
 *
 * out[b, s, :] = table[ids[b, s], :] (nn.Embedding); rows addressed as out + b * out_batch_stride + s * ldo.  ids: device int64.  */
int x2i_gather_rows(const int64_t* ids, const void* table, int64_t ldt, int vocab, void* out, int64_t ldo, int64_t out_batch_stride,
                    int rows, int rows_per_batch, int D, void* stream);
/* y = weight * bf16(x * rsqrt(mean(x^2) + eps))   (Qwen2RMSNorm).  x and y rows addressed with a batch stride like above.         */
int x2i_rmsnorm(const void* x, int64_t ldx, int64_t x_batch_stride, const void* weight, void* y, int64_t ldy, int64_t y_batch_stride,
                int rows, int rows_per_batch, int D, float eps, void* stream);
/* Fused QKV rows [B*S, ld] = [q (heads*128) | k (heads_kv*128) | v (heads_kv*128)] -> head-major q [B,heads,S,128],
 * k, v [B,heads_kv,S,128] with the rotate-half RoPE of Qwen2 applied to q and k.  pos: device int32 [B*S] token positions
 * (cumsum(attention_mask) - 1, padded tokens 1: Qwen2_5_VLModel.get_rope_index, text-only); inv_freq: device fp32 [64].          */
int x2i_rope_half_split(const void* qkv, int64_t ld, const int* pos, const float* inv_freq, void* q, void* k, void* v, int B, int S,
                        int heads, int heads_kv, void* stream);
/* C[M, N/2] = act(A Wg^T + bg) * (A Wu^T + bu): a gated MLP's gate / up projections and activation in one GEMM.  act 0 = SiLU
 * (SwiGLU: Qwen2MLP), act 1 = tanh-GELU (GEGLU: T5's "gated-gelu" DenseGatedActDense with gelu_new, model_internvl/proj.py:143).
 * W [N, K] holds the gate and up rows interleaved in blocks of 128 (rows [256 t, 256 t + 128) = gate rows [128 t, +128), the next
 * 128 = the matching up rows); bias (nullable) is laid out the same way.  N % 256 == 0.                                        */
int x2i_gemm_swiglu(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C, int64_t ldc, int M, int N, int K,
                    int act, void* stream);
/* P[r, :] (bf16) = softmax(S[r, :] + bias[r % bias_rows, :]) over fp32 scores: T5Attention's `scores += position_bias` + softmax
 * (the T5Stack inside model_internvl/proj.py:139-211; transformers' modeling_t5.py).                                            */
int x2i_softmax_rows_bias(const float* S, int64_t lds, const float* bias, int64_t ldb, int bias_rows, void* P, int64_t ldp, int rows, int cols,
                          void* stream);
/* Causal grouped-query self-attention of a left-padded prompt: q [B,heads,L,128], k, v [B,heads_kv,L,128] -> out [B*L, ld] token-major
 * (column h*128 + d).  Key j is visible to query i iff kv_start[b] <= j <= i (kv_start: device int32 [B], nullable = 0); a query row
 * with no visible key (a padded position) outputs 0 -- the behaviour of transformers' sdpa / flash paths.  Same kernel as
 * x2i_mmdit_attention (softmax(q k^T / sqrt(128)) v), key tiles right of the diagonal and left of kv_start are skipped.         */
int x2i_causal_attention(const void* q, const void* k, const void* v, const int* kv_start, void* out, int64_t ld, int B, int heads,
                         int heads_kv, int L, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* X2I_B200_H_ */
