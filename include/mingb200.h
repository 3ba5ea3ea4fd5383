/*
 * libmingb200 — C ABI of the B200-native (sm_100a) operators behind the Ming-UniVision continuous-visual-token
 * hot path (MingTok encoder / semantic decoder / pixel decoder, Bailing-MoE AR step, rectified-flow head).
 *
 * The reference (inclusionAI/Ming-UniVision) is pure Python/PyTorch and has no FFI; the boundary it exposes is its
 * Python class API.  Each entry point below therefore names the reference call site whose
 * torch / flash-attn operator it replaces (paths relative to the reference repo).  The Python host modules in
 * ming_univision_b200/ keep the reference's class names, method signatures and state_dict keys and call these
 * functions through ctypes with raw device pointers.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  All data pointers are DEVICE pointers (cudaMalloc'ed / torch CUDA
 *     storage) unless a parameter is explicitly marked HOST.
 *   - Matrices are row-major.  `ld*` arguments are row strides in ELEMENTS.  bf16 = __nv_bfloat16 (uint16 storage).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous on that stream,
 *     never allocate, never synchronise, and are CUDA-graph capturable.
 *   - Return value: 0 (MB_OK) or a negative MB_ERR_* code; mb_last_error() returns a thread-local message.
 *   - There is no CPU fallback: on a machine without an sm_100 device every compute entry point returns MB_ERR_ARCH.
 */
#ifndef MINGB200_H_
#define MINGB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_OK 0
#define MB_ERR_SHAPE (-1) /* unsupported or inconsistent shape */
#define MB_ERR_ALIGN (-2) /* pointer / stride alignment (TMA needs 16-byte aligned bases and row strides) */
#define MB_ERR_ARCH (-3)  /* no sm_100 device */
#define MB_ERR_CUDA (-4)  /* CUDA runtime / driver error, see mb_last_error() */
#define MB_ERR_NCCL (-5)

/* GEMM epilogues (mb_gemm_bf16 `epi`) */
#define MB_EPI_BIAS 0     /* out = bf16(acc + bias)                                   nn.Linear */
#define MB_EPI_GELU 1     /* out = bf16(gelu_erf(bf16(acc + bias)))                   nn.Linear -> nn.GELU() */
#define MB_EPI_SWIGLU 2   /* packed w12: out = bf16(bf16(silu(x1)) * x2), N/2 columns F.silu(x1) * x2 */
#define MB_EPI_RESIDUAL 3 /* out = bf16(bf16(acc + bias) + residual)                  x + f(x) */
#define MB_EPI_SILU 4     /* (mb_gemv_bf16 only) out = bf16(silu(bf16(acc + bias)))  nn.Linear -> nn.SiLU() */
#define MB_EPI_GATED 5    /* (mb_gemv_bf16 only) out = bf16(res + bf16(gate * bf16(acc + bias)))  x + gate * h */

const char* mb_last_error(void);
/* ABI version (bumped whenever a signature changes) and the device check used by the loaders. */
int mb_abi_version(void);
int mb_device_ok(void); /* 1 if the current device is sm_100, else 0 */
int mb_num_sms(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Dense linear layers  (tcgen05.mma + TMEM accumulators, operands staged by TMA, 128B-swizzled shared memory)
 *
 * out[M, N'] = epilogue( A[M, K] @ W[N, K]^T + bias[N] )      A, W, bias, residual, out: bf16; accumulate fp32.
 * Replaces every nn.Linear on the path:
 *   mingtok/vision_transformer/layers/attention.py:49,63 (qkv), :51,72 (proj); layers/swiglu_ffn.py:28-34 (w12, w3);
 *   layers/mlp.py:27-39 (fc1 + GELU, fc2); layers/patch_embed.py:66,76 (32x32/s32 conv as GEMM after mb_patchify);
 *   vision_transformer.py:171,177 (out_proj), :282,379 (in_proj), :366,483 (head); modeling_mingtok.py:117,183
 *   (sem_to_pix); mingunivision/modeling_bailingmm.py:111-115 (linear_proj);
 *   mingunivision/modeling_bailing_moe.py:760,824 (query_key_value, dense), :479-484 (expert / shared-expert MLP),
 *   :1571-1574 (vis_head), :1619 (lm_head); mingunivision/diff_loss_rf_swiglu.py:27-34,215-219,311-313.
 *
 *   epi = MB_EPI_SWIGLU: W/bias must be packed by mb_pack_swiglu_rows (gate/up rows interleaved in blocks of 128),
 *         N is the packed row count (multiple of 256) and the output has N/2 columns.
 *   epi = MB_EPI_RESIDUAL: residual row for output row r is (res_row_mod > 0 ? r % res_row_mod : r).
 *   Output row remap (all epilogues): out_row = r + (r / out_row_group) * out_row_pad when out_row_group > 0
 *         (used to leave room for the cls token that the reference concatenates at the END of every image,
 *         vision_transformer.py:221).
 *   bias may be NULL.  K % 8 == 0, lda % 8 == 0, ldw % 8 == 0, ldo % 8 == 0; any M, N >= 1.
 * ------------------------------------------------------------------------------------------------------------- */
int mb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out, int64_t ldo,
                 int M, int N, int K, int epi, const void* residual, int64_t ldr, int res_row_mod, int out_row_group,
                 int out_row_pad, void* stream);

/* The same GEMM with a LayerNorm FOLDED into it (removes the LayerNorm kernel and its read + write of the activation
 * tensor between a residual GEMM and the qkv / w12 / fc1 GEMM that consumes the normalised rows;
 * block.py:80-105: x + attn(norm1(x)), x + mlp(norm2(x))):
 *   producer side  (epi = MB_EPI_RESIDUAL, ln_stats_out != NULL): while writing its rows the epilogue stores, per row
 *       and per 64-column box, (sum, sum of squares) of the bf16 values it writes into
 *       ln_stats_out[M][ceil(N / 64)][2] — every slot written exactly once, no atomics, bitwise reproducible;
 *   consumer side  (ln_stats_in != NULL): A holds the UN-normalised rows; W must be pre-scaled by gamma
 *       (W' = bf16(W * gamma)), ln_csum[n] = sum_k W'[n, k] and ln_bias_f32[n] = bias[n] + sum_k W[n, k] beta[k] (fp32,
 *       packed at load time); the epilogue computes  rstd_r * (acc - mean_r * csum[n]) + bias_f32[n]  with
 *       mean / rstd from the ln_slots_in partial sums of ln_stats_in[M][ln_slots_in][2] (added in slot order), which
 *       equals  LN(A) W^T + bias  with fp32 statistics and no intermediate bf16 rounding of the normalised
 *       activations.  `bias` is ignored on the consumer side.
 * mb_row_stats seeds the chain for the first block of a stage. */
int mb_gemm_bf16_ex(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out, int64_t ldo,
                    int M, int N, int K, int epi, const void* residual, int64_t ldr, int res_row_mod,
                    int out_row_group, int out_row_pad, const float* ln_stats_in, int ln_slots_in,
                    const float* ln_csum, const float* ln_bias_f32, float ln_eps, float* ln_stats_out, void* stream);
int mb_row_stats(const void* x, int64_t ldx, float* stats, int rows, int dim, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Weight-streaming skinny GEMM for the decode regime (M <= 8 rows; HBM-bound, every weight byte read once):
 *   out[M, N'] = epilogue(A[M, K] @ W[N, K]^T + bias)       A, W, bias, residual, gate, out: bf16; accumulate fp32.
 * Replaces the M <= 3 nn.Linear calls of the rectified-flow head (diff_loss_rf_swiglu.py:30-34, 215-219, 262-265,
 * 283-286, 311-313), of forward_for_image_generation_inner / the AR decode step (modeling_bailing_moe.py:760, 824,
 * 479-484, 1571-1574, 1619) and of the cached semantic-decoder step (layers/attention.py:215, swiglu_ffn.py:30-34).
 *   epi = MB_EPI_SWIGLU takes the REFERENCE layout of w12 ([2H, K]: x1 rows then x2 rows) and writes H columns.
 *   epi = MB_EPI_GATED : out = res + gate * (A W^T + b)   (ResBlock, diff_loss_rf_swiglu.py:272).
 *   out_f32 (optional, may be NULL): fp32 copy of the bf16-rounded output, dense [M, N'] (logits .float(),
 *   modeling_bailing_moe.py:1785).
 * ------------------------------------------------------------------------------------------------------------- */
int mb_gemv_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out, int64_t ldo,
                 int M, int N, int K, int epi, const void* residual, int64_t ldr, const void* gate, int64_t ldg,
                 void* out_f32, void* stream);
/* Same with the input normalisation FUSED into the staging of the activation rows (no separate row kernel in front
 * of the streaming GEMM):
 *   norm = 1  adaLN   : a' = (LN(a) * gamma + beta) * bf16(1 + scale[m]) + shift[m]   (gamma / beta may be NULL;
 *                       ResBlock / FinalLayer of the RF head, diff_loss_rf_swiglu.py:184-185, 270, 290)
 *   norm = 2  RMSNorm : a' = gamma * bf16(a * rsqrt(mean(a^2) + eps))                  (modeling_bailing_moe.py:131-136)
 * statistics in fp32, a' rounded to bf16 before the product, exactly as mb_adaln_modulate / mb_rmsnorm round. */
int mb_gemv_bf16_norm(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out, int64_t ldo,
                      int M, int N, int K, int epi, const void* residual, int64_t ldr, const void* gate, int64_t ldg,
                      void* out_f32, int norm, const void* gamma, const void* beta, const void* shift, int64_t ld_shift,
                      const void* scale, int64_t ld_scale, float eps, void* stream);

/* Tuning / test hook: pin the GEMM tile shape instead of the built-in heuristic.  cta_group: 1 = one CTA per
 * 128 x bn tile, 2 = CTA pair (tcgen05 cta_group::2) per 256 x bn tile, 0 = automatic; bn: 128, 256 or 0 = automatic.
 * Adding 16 to cta_group forces the direct register-store epilogue instead of the staged TMA-store epilogue.
 * Process-wide; also settable through the MB_GEMM_CG / MB_GEMM_BN / MB_GEMM_NO_TMA_EPI environment variables. */
int mb_gemm_force_tile(int cta_group, int bn);

/* Weight pre-pack for MB_EPI_SWIGLU: src is the reference's w12 [2*H, K] (x1 rows then x2 rows, swiglu_ffn.py:32),
 * dst is [2*Hp, K] with Hp = round_up(H, 128): block b holds rows x1[128b..128b+127] then x2[128b..128b+127]; rows
 * beyond H are zero.  The same call packs the bias with K = 1.  (One-time, at load.) */
int mb_pack_swiglu_rows(const void* src, void* dst, int H, int Hp, int K, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Normalisation (one warp per row, fp32 statistics, bf16 in/out)
 *   mb_layernorm: y = LN(x) * gamma + beta, eps; act = 0 none, 1 exact-erf GELU applied to bf16(LN) (encoder out
 *   layer, vision_transformer.py:173-178).  gamma/beta may be NULL (elementwise_affine=False,
 *   diff_loss_rf_swiglu.py:281).  Replaces nn.LayerNorm at layers/block.py:53,66,311,320,
 *   vision_transformer.py:169,363,431.
 *   Input row r lives at x + (r / rows_per_group) * group_stride_x + (r % rows_per_group) * ldx when
 *   rows_per_group > 0 (reads only the first n of every n+1 tokens: the decoder drops the trailing cls token after
 *   its final norm, vision_transformer.py:431-439); plain r * ldx when rows_per_group == 0.  Output rows are dense.
 * ------------------------------------------------------------------------------------------------------------- */
int mb_layernorm(const void* x, int64_t ldx, const void* gamma, const void* beta, void* y, int64_t ldy, int rows,
                 int dim, float eps, int act, int rows_per_group, int64_t group_stride_x, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Attention, head_dim 64, packed QKV as produced by the qkv Linear:  qkv[B, S, 3, H, 64] bf16 -> out[B, S, H*64].
 * softmax((q k^T) * scale) v with fp32 softmax; causal != 0 applies the lower-triangular mask.
 * Replaces flash_attn_func at layers/attention.py:101 (full) and :232-234 (causal) and the eager twins :61-74,
 * :138-163.
 * ------------------------------------------------------------------------------------------------------------- */
int mb_attn_hd64(const void* qkv, void* out, int B, int S, int H, float scale, int causal, void* stream);

/* General strided form (head_dim 64 or 128, GQA when Hq > Hkv): element strides (batch, token, head) per tensor;
 * causal masks are bottom-right aligned.  Used for the LLM prefill (flash_attn_func, modeling_bailing_moe.py:988-1005)
 * reading K/V straight from the [B, Hkv, Tmax, hd] cache. */
int mb_attn_fwd(const void* q, int64_t q_bs, int64_t q_ts, int64_t q_hs, const void* k, int64_t k_bs, int64_t k_ts,
                int64_t k_hs, const void* v, int64_t v_bs, int64_t v_ts, int64_t v_hs, void* out, int64_t o_bs,
                int64_t o_ts, int64_t o_hs, int B, int Sq, int Sk, int Hq, int Hkv, int hd, float scale, int causal,
                void* stream);

/* Attention backend of mb_attn_hd64 / mb_attn_fwd: 0 = auto (tcgen05 / TMEM kernel whenever eligible), 1 = same,
 * 2 = force the warp-level mma.sync kernel (A/B measurements and tests).  Env MB_ATTN_BACKEND sets the default. */
int mb_attn_set_backend(int backend);
/* Development aid: device buffer (>= 64 x 16 int64) receiving clock64() phase stamps of CTA 0 of the tcgen05 attention
 * kernel; NULL switches it off. */
int mb_attn_set_debug(void* dev_buf);
/* Same for the GEMM kernel: >= 16 x 8 int64, stamps of the first CTA's MMA-issue and epilogue roles per tile. */
int mb_gemm_set_debug(void* dev_buf);
/* Decode-step attention against a static KV cache (semantic decoder, q_len = 1; layers/attention.py:213-239 with
 * past_key_value).  qkv[B, 3, H, 64] holds the new token; its K/V are appended at position `t` of
 * kcache/vcache[B, H, Tmax, 64] (DynamicCache.update, vision_transformer.py:396) and q attends to positions 0..t.
 * The position is t + *t_dev when t_dev (an optional DEVICE int32 scalar) is given — see the AR-step note below. */
int mb_attn_hd64_decode(const void* qkv, void* kcache, void* vcache, void* out, int B, int H, int t, int Tmax,
                        float scale, const int32_t* t_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * MingTok data-movement operators
 * ------------------------------------------------------------------------------------------------------------- */
/* im2row for the 32x32 stride-32 patch conv (layers/patch_embed.py:76-78): img[B, C, Hh, Ww] (fp32 if
 * img_is_fp32 else bf16) -> rows[B*gh*gw, C*P*P] bf16, column order (c, py, px) = Conv2d weight.flatten(1). */
int mb_patchify(const void* img, int img_is_fp32, void* rows, int B, int C, int Hh, int Ww, int P, void* stream);
/* x[b, n, :] = bf16(cls + pos_cls) for every image (vision_transformer.py:221-222; cls token appended at the END). */
int mb_fill_cls_row(void* x, const void* cls, const void* pos_cls, int B, int n_plus_1, int dim, void* stream);
/* out[r, c] = mean_j x[r, c*g + j], g = dim / groups   (encoder shortcut, vision_transformer.py:174). */
int mb_group_mean(const void* x, int64_t ldx, void* out, int rows, int dim, int groups, void* stream);
/* y = bf16(x * scale + shift) elementwise; x is fp32 if x_is_fp32 else bf16 (latent (de)normalisation,
 * modeling_mingtok.py:162,168 — the RF sampler hands an fp32 latent to forward_feature_decoder). */
int mb_affine(const void* x, int x_is_fp32, void* y, int64_t n, float scale, float shift, void* stream);
/* Semantic-decoder input layer (vision_transformer.py:373-380):
 * out[r, :] = bf16(bf16(W[dim, in_dim] @ x[r] + b) + repeat_interleave(x[r], dim / in_dim)); in_dim <= 64. */
int mb_inproj_repeat(const void* x, const void* W, const void* b, void* out, int rows, int in_dim, int dim,
                     void* stream);
/* sem_to_pix rearrange "b (h w) (x y c) -> b (h x w y) c" (modeling_mingtok.py:184-188):
 * in[B, g*g, f*f*C] -> out[B, (g*f)*(g*f), C]. */
int mb_pixel_shuffle(const void* in, void* out, int B, int g, int f, int C, void* stream);
/* unpatchify + clamp(-1, 1) (vision_transformer.py:515-527, modeling_mingtok.py:192-194):
 * x[B, g*g, p*p*3] (channel-last inside the patch) -> img[B, 3, g*p, g*p]; out fp32 if out_is_fp32 else bf16. */
int mb_unpatchify_clamp(const void* x, void* img, int out_is_fp32, int B, int g, int p, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Rectified-flow head (mingunivision/diff_loss_rf_swiglu.py) row helpers
 * ------------------------------------------------------------------------------------------------------------- */
/* y[m] = bf16((LN(x[m]) * gamma + beta) * bf16(1 + scale[m]) + shift[m]); gamma/beta NULL for the affine-free final
 * norm.  modulate(), diff_loss_rf_swiglu.py:184-185 at :270 and :290. */
int mb_adaln_modulate(const void* x, int64_t ldx, const void* gamma, const void* beta, const void* shift,
                      int64_t ld_shift, const void* scale, int64_t ld_scale, void* y, int64_t ldy, int rows, int dim,
                      float eps, void* stream);
/* out[s*B + b, :] = bf16(silu(bf16(temb[s, :] + c[b, :]))): the SiLU(y = t + c) input of every adaLN_modulation
 * Linear (diff_loss_rf_swiglu.py:376, 262-265, 283-286) for all `steps` sampling steps at once, so the adaLN weights
 * (28 % of the head) are streamed once per token instead of once per Euler step. */
int mb_silu_add_rows(const void* temb, const void* c, void* out, int steps, int B, int dim, void* stream);
/* CFG combine + explicit Euler update of RectifiedFlowLoss.sample (diff_loss_rf_swiglu.py:145-179):
 * v bf16 [B, C]: B / cfg_rows independent samples of cfg_rows adjacent rows (cond, uncond[, text_uncond]);
 * cfg_rows == 3: v = v_u + image_cfg (v_tu - v_u) + text_cfg (v_c - v_tu); 2: v = v_u + text_cfg (v_c - v_u); 1: v.
 * x_f32[B, C] += bf16(v * dt) on every row of the sample; x_bf16 receives the bf16 copy that feeds input_proj on the
 * next step. */
int mb_rf_euler_step(void* x_f32, void* x_bf16, const void* v, int B, int cfg_rows, int C, float dt, float text_cfg,
                     float image_cfg, void* stream);
/* The whole sampler loop of RectifiedFlowLoss.sample (diff_loss_rf_swiglu.py:134-179: `steps` Euler steps x
 * [input_proj :371, `depth` x ResBlock :268-272, FinalLayer :288-292, CFG combine + Euler update]) as ONE persistent
 * weight-streaming kernel (csrc/rf_fused.cu): one CTA per SM, every CTA owns a fixed slice of the output rows of every
 * layer, the weights stream through a shared-memory ring by cp.async.bulk ahead of the grid barriers between layers.
 * mb_rf_pack_weights re-orders one weight matrix W [N, K] (nn.Linear layout; swiglu: N = 2H, gate rows then up rows)
 * ONCE into the per-CTA stage order for `n_cta` = mb_num_sms() CTAs (same byte count).  block_ptrs: DEVICE table
 * [depth][6] = {w12 packed, b12 [2H], w3 packed, b3 [W], in_ln weight [W], in_ln bias [W]}; mod [steps * B, ld_mod] =
 * the adaLN modulations of all steps (depth x (shift | scale | gate) then final (shift | scale)); x [B, C] fp32 is the
 * noise on entry and the sample on exit; scratch h [B, W], hid [B, H], v [B, C] bf16 and one u32 barrier word.
 * B rows = B / cfg_rows independent samples (images generated together) of cfg_rows adjacent CFG rows each (cond, uncond
 * [, text-uncond]): the weights stream ONCE for all of them.  Supported shapes: mb_rf_fused_supported(B, W, H, C)
 * (B <= 6 rows, W and H multiples of 1024, W <= 4096, C <= 32). */
int mb_rf_fused_supported(int B, int W, int H, int C);
int mb_rf_set_debug(void* buf16_u64); /* debug aid: per-phase wall time of the first / last CTA, NULL = off */
int mb_rf_pack_weights(const void* W, int N, int K, int swiglu, int n_cta, void* out, void* stream);
int mb_rf_sample_fused(const void* const* block_ptrs, const void* in_w, const void* in_b, const void* fin_w,
                       const void* fin_b, const void* mod, int64_t ld_mod, float* x, void* h_scratch, void* hid_scratch,
                       void* v_scratch, uint32_t* barrier, int B, int cfg_rows, int W, int H, int C, int depth,
                       int steps, float text_cfg, float image_cfg, int n_cta, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Bailing-MoE AR step (mingunivision/modeling_bailing_moe.py).  `t_dev` arguments are optional DEVICE int32 scalars
 * holding the current KV-cache length, so ONE captured CUDA graph can be replayed for every generated token; the
 * effective value is always  (t_dev ? *t_dev : 0) + t_host.
 * ------------------------------------------------------------------------------------------------------------- */
/* BailingMoeRMSNorm.forward (:131-136): y = bf16(w * (x * rsqrt(mean(x^2) + eps))), fp32 statistics. */
int mb_rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows, int dim, float eps,
               void* stream);
/* apply_rotary_pos_emb (:436-461) with the 1-D legacy tables (:213-237), fused with DynamicCache.update (:789):
 * qkv[B*S, (H + 2*Hkv) * hd] -> q_out[B*S, H*hd] (rotated), K (rotated) and V appended to kcache / vcache
 * [B, Hkv, Tmax, hd] at slots t .. t+S-1; position_ids int32 [B*S]. */
int mb_rope_kv_append(const void* qkv, const int32_t* position_ids, void* q_out, void* kcache, void* vcache, int B,
                      int S, int H, int Hkv, int hd, int Tmax, const int32_t* t_dev, int t_host, float rope_theta,
                      void* stream);
/* The config-gated 3-D multimodal variant (`rope_scaling.type == "3D"`): BailingMoe3DRotaryEmbedding.forward (:413-425)
 * + apply_multimodal_rotary_pos_emb (:463-469).  position_ids3: int32 [3, B, S] (temporal, height, width); frequency i
 * of a head takes its angle from component 0 / 1 / 2 for i < sec0 / < sec0 + sec1 / else (the reference's sections
 * [16, 24, 24] doubled over both halves).  cos / sin stay fp32 here, so q and k are rounded to bf16 once, after
 * q*cos + rotate_half(q)*sin in fp32 (the 1-D legacy path above rounds the tables and every product to bf16).
 * Layouts, cache slots and t_dev / t_host as mb_rope_kv_append. */
int mb_rope3d_kv_append(const void* qkv, const int32_t* position_ids3, void* q_out, void* kcache, void* vcache, int B,
                        int S, int H, int Hkv, int hd, int Tmax, const int32_t* t_dev, int t_host, float rope_theta,
                        int sec0, int sec1, int sec2, void* stream);
/* GQA attention for q_len == 1 (head_dim 128) over cache slots 0 .. T-1, skipping keys whose key_mask[b, j] == 0
 * (the 2-D padding mask of the CFG rows; _upad_input / flash_attn_varlen_func :1009-1045, eager :795-812).
 * key_mask may be NULL; mask_stride = elements between rows of key_mask.  n_splits > 1 spreads the keys of every head
 * over n_splits CTAs (flash-decoding: per-chunk partial max / sum / output in `workspace`, B * H * n_splits * 130
 * floats, merged by a second kernel) so that long contexts use the whole GPU; n_splits = 1 needs no workspace. */
int mb_attn_decode_gqa(const void* q, const void* kcache, const void* vcache, const int32_t* key_mask,
                       int64_t mask_stride, void* out, int B, int H, int Hkv, int hd, int Tmax, const int32_t* t_dev,
                       int t_host, float scale, float* workspace, int n_splits, void* stream);
/* Greedy next-token choice over fp32 logits [rows, V] (HF generate with do_sample = false, mingunivision/config.json:30;
 * first index on ties, as torch.argmax).  n_chunks > 1: two stages (per-chunk winners in ws_val / ws_idx
 * [rows * n_chunks], then the row winner) so that a 126 k-entry vocabulary is scanned by many SMs. */
int mb_argmax_f32(const float* x, int32_t* out, int rows, int V, float* ws_val, int32_t* ws_idx, int n_chunks,
                  void* stream);
/* BailingMoeGate.forward (:505-520) after the logits GEMM: fp32 softmax over E bf16 logits, top-k, renormalise.
 * With logits_img + image_mask (uint8 [T]) tokens flagged as image tokens use the image gate's logits (:574-580). */
int mb_router_topk(const void* logits, const void* logits_img, const uint8_t* image_mask, int32_t* idx,
                   float* weights, int T, int E, int k, int renorm, void* stream);
/* moe_infer (:608-639) for small token counts: counting sort of the (token, slot) pairs by expert, expert gate/up +
 * SwiGLU on the sorted pairs (Wgu[E][2I][D]: gate rows then up rows), expert down projection (Wd[E][D][I]) written
 * back in (token, slot) order, and the fp32 weighted combine + shared-expert add + layer residual.
 * E is the number of LOCAL experts and e_begin the first one (E = all, e_begin = 0 without expert parallelism). */
int mb_moe_sort(const int32_t* idx, int32_t* expert_offsets, int32_t* sorted_pair, int T, int k, int E, int e_begin,
                void* stream);
int mb_moe_gate_up(const void* x, const void* Wgu, const int32_t* expert_offsets, const int32_t* sorted_pair,
                   void* hid, int T, int k, int E, int D, int I, int mean_pairs, void* stream);
int mb_moe_down(const void* hid, const void* Wd, const int32_t* expert_offsets, const int32_t* sorted_pair,
                void* out_pairs, int T, int k, int E, int D, int I, int mean_pairs, void* stream);
/* (mean_pairs: expected pairs per expert of the call, ceil(T k / number of experts the pairs spread over) — selects the
 * 8- or the 32-rows-per-pass form of the streaming kernel.) */
/* pair_row == NULL: out_pairs is in (token, slot) order; else pair p = t*k+j lives in row pair_row[p] of out_pairs
 * (grouped layout of mb_moe_plan; negative = expert not local, contributes zero). */
int mb_moe_combine(const void* out_pairs, const float* weights, const void* shared, const void* residual, void* y,
                   float* y_partial, const int32_t* pair_row, int T, int k, int D, void* stream);
/* moe_infer (:608-639) for the PREFILL regime (hundreds of tokens per expert), on tcgen05 tensor cores:
 * mb_moe_plan lays the (token, slot) pairs out bucket-sorted, bucket = (expert - e_begin) / div, every bucket's
 * segment padded to `granule` rows (pair_row[T*k], row_token[max_rows], tile_expert[max_rows/128] or NULL,
 * meta = {128-row tiles, rows}, counts[E] or NULL; max_rows >= roundup(T*k + (granule-1) E)).  granule = 128, div = 1
 * is the tile plan of the grouped GEMMs; granule = 1, div = experts per rank orders an expert-parallel send buffer by
 * destination rank.  mb_moe_gather_rows copies rows x[row_token[r]] into that layout (row count from meta[1] on the
 * device, or max_rows when meta is NULL; negative indices give zero rows), and mb_moe_grouped_gemm runs
 * one persistent TMA/tcgen05 GEMM over all experts: swiglu = 1 -> out[r, 0:I] = silu(A W_gate^T) * (A W_up^T) with
 * W = [E][2I][K] (gate rows then up rows, no repack), swiglu = 0 -> out = A W^T with W = [E][N][K].  The number of
 * tiles is read from meta on the device: no host synchronisation anywhere (the reference syncs per layer, :616). */
int mb_moe_plan(const int32_t* idx, int32_t* pair_row, int32_t* row_token, int32_t* tile_expert, int32_t* meta,
                int32_t* counts, int T, int k, int E, int e_begin, int div, int granule, int max_rows, void* stream);
int mb_moe_gather_rows(const void* x, const int32_t* row_token, const int32_t* meta, void* out, int max_rows, int D,
                       void* stream);
int mb_moe_grouped_gemm(const void* A, const void* W, void* out, const int32_t* tile_expert,
                        const int32_t* num_m_tiles, int max_rows, int N, int K, int E, int swiglu, void* stream);
/* Whole-stage driver for decode-sized inputs: mb_moe_sort -> mb_moe_gate_up -> mb_moe_down -> mb_moe_combine chained on
 * the caller's stream inside a caller-provided workspace (mb_moe_ffn_workspace_bytes; 256-byte aligned): y [T, D] =
 * bf16(bf16(bf16(sum_j w[t,j] expert_{idx[t,j]}(x[t])) + shared[t]) + residual[t]) — one call per MoE layer for a host
 * that does not want to orchestrate the four kernels (BailingMoeSparseMoeBlock.moe_infer :608-639, :604-605, :1226). */
int mb_moe_ffn_workspace_bytes(int T, int k, int E, int D, int I, int64_t* bytes);
int mb_moe_ffn(const void* x, const int32_t* idx, const float* weights, const void* Wgu, const void* Wd,
               const void* shared, const void* residual, void* y, void* workspace, int64_t workspace_bytes, int T, int k,
               int E, int e_begin, int n_experts_total, int D, int I, void* stream);
/* Expert parallelism (no reference implementation — the reference keeps all experts on one device):
 * rank r owns experts [e_begin, e_begin + E) — mb_moe_sort lists only the pairs routed to them, mb_moe_gate_up /
 * mb_moe_down run on the local slabs, mb_moe_combine with y_partial != NULL writes this rank's fp32 share of the
 * weighted sum (non-local pairs contribute zeros), the ranks all-reduce the [T, D] fp32 partials over NCCL, and
 * mb_moe_finalize applies the reference's rounding chain: bf16(bf16(bf16(sum) + shared) + residual). */
int mb_moe_finalize(const float* y_sum, const void* shared, const void* residual, void* y, int T, int D, void* stream);
/* The same combine FUSED with its exchange over NVLink peer memory (decode-sized inputs; no NCCL, no host sync).
 * Every rank owns an exchange area of mb_moe_peer_area_bytes(G, Tmax, D) bytes, zero-initialised, in memory that all
 * ranks have mapped (torch symmetric memory / CUDA IPC); `peers` is a DEVICE array of the G base pointers as seen from
 * this process.  mb_moe_combine_push stores this rank's fp32 partial sums straight into every peer's area (st.global
 * on peer pointers) and raises a system-scope flag there; mb_moe_reduce_finalize waits for the G flags of the local
 * area, adds the G partials in rank order (bit-identical on every rank), applies the rounding chain of mb_moe_finalize
 * and advances the epoch.  `fin_done` is a zero-initialised local device word.  Both calls must be made by every rank,
 * in the same order, on its stream. */
int mb_moe_peer_area_bytes(int G, int Tmax, int D, int64_t* bytes);
int mb_moe_combine_push(const void* out_pairs, const float* weights, const int32_t* pair_row, float* const* peers,
                        int my_rank, int G, int T, int Tmax, int k, int D, void* stream);
int mb_moe_reduce_finalize(float* const* peers, int my_rank, int G, int T, int Tmax, int D, const void* shared,
                           const void* residual, void* y, uint32_t* fin_done, void* stream);
/* Expert parallelism with the DISPATCH on peer memory too ("data parallel x expert parallel"; csrc/ep.cu): every rank
 * works on its OWN T rows and owns experts [e_begin, e_begin + e_local).  One zero-initialised, peer-mapped exchange
 * area per rank (byte offsets of its x / idx / w / partial-sum / control regions and its size: mb_ep_area_layout ->
 * offsets6).  Per MoE layer, every rank, same order, same T, on its stream:
 *   mb_ep_dispatch_push   stores x [T, D] bf16, idx [T, k] i32 (GLOBAL expert ids), w [T, k] f32 into every peer's area
 *                         (row block my_rank of the packed [G*T, .] arrays), raises the dispatch flags;
 *   mb_ep_dispatch_wait   bounded wait for the G dispatch flags of the local area; afterwards the local area holds the
 *                         gathered rows of all ranks — run mb_moe_sort / gate_up / down (or mb_moe_plan / grouped GEMM)
 *                         on them with (e_begin, e_local);
 *   mb_ep_combine_push    fp32 partial sums of the local experts for all G*T rows (pair_row = grouped layout or NULL for
 *                         (token, slot) order), rows of source rank q stored into q's area, raises the combine flags;
 *   mb_ep_reduce_finalize bounded wait for the G combine flags, sum in rank order, bf16(bf16(bf16(sum) + shared) +
 *                         residual) -> y [T, D], advances the device-side epoch.
 * No NCCL, no host synchronisation, capturable in a CUDA graph.  A wait that exceeds MB_EP_TIMEOUT_MS (default 20000)
 * records a code in the control block (u32 words 2G+4: 1 = dispatch, 2 = combine; 2G+5: the missing rank) and returns;
 * the host checks it after synchronising. */
int mb_ep_area_layout(int G, int Tmax, int D, int k, int64_t* offsets6);
int mb_ep_dispatch_push(const void* x, const int32_t* idx, const float* w, void* const* peers, int my_rank, int G, int T,
                        int Tmax, int D, int k, void* stream);
int mb_ep_dispatch_wait(void* const* peers, int my_rank, int G, int Tmax, int D, int k, void* stream);
/* mb_ep_dispatch_wait + mb_moe_sort of the gathered pairs (expert_offsets [E + 1], sorted_pair [G*T*k]) in one launch. */
int mb_ep_wait_sort(void* const* peers, int my_rank, int G, int T, int Tmax, int D, int k, int32_t* expert_offsets,
                    int32_t* sorted_pair, int E, int e_begin, void* stream);
int mb_ep_combine_push(const void* out_pairs, const int32_t* pair_row, void* const* peers, int my_rank, int G, int T,
                       int Tmax, int D, int k, int e_begin, int e_local, void* stream);
int mb_ep_reduce_finalize(void* const* peers, int my_rank, int G, int T, int Tmax, int D, int k, const void* shared,
                          const void* residual, void* y, void* stream);

/* ---- Image pre- / post-processing either side of MingTok -------------------------------------------
 * Replaces, on the device, the PIL / torchvision CPU transforms of `CenterCropProcessor` (mingtok/utils/processor.py:17-27),
 * `MingTokUndProcessor` / `MingTokCenterCropProcessor` (mingunivision/processing_bailingmm.py:80-123):
 *   Resize(bicubic; PIL image => Pillow's ANTIALIASED resample in 8-bit fixed point, horizontal pass, u8 rounding,
 *   vertical pass) -> CenterCrop -> ToTensor (u8 / 255) -> Normalize ((x - mean) / std),
 * bit-exact with Pillow 12 + torchvision (u8 image identical, fp32 tensor identical; bf16 = its RN rounding).
 * src: [n, in_h, in_w, 3] u8 (RGB, HWC) in device memory.  res_h x res_w is the size after Resize (the caller applies
 * torchvision's rule: short edge -> size, long edge = int(size * long / short)); (crop_top, crop_left, out_h, out_w) the
 * kept window of the resized image (the whole image when there is no crop).  Only the kept columns / rows are computed.
 * out: [n, 3, out_h, out_w] bf16 (out_kind 0) or fp32 (out_kind 1), or — out_kind 2 — the resized + cropped image
 * itself, [n, out_h, out_w, 3] u8 (Pillow's `Image.resize` + crop on the device; mean / std unused).  workspace: >= mb_image_preprocess_workspace_bytes(...)
 * bytes, 16-byte aligned (coefficient tables + the u8 intermediate); no allocation, no host synchronisation.
 * MB_ERR_SHAPE for a crop window outside the resized image and for images more than 100 times taller than wide whose
 * height shrinks (Pillow >= 11 resizes those height-first; the pass order shows in the u8 result). */
int mb_image_preprocess_workspace_bytes(int n, int in_h, int in_w, int res_h, int res_w, int crop_top, int crop_left,
                                        int out_h, int out_w, int64_t* bytes);
int mb_image_preprocess_u8(const void* src, int n, int in_h, int in_w, int res_h, int res_w, int crop_top,
                           int crop_left, int out_h, int out_w, float mean0, float mean1, float mean2, float std0,
                           float std1, float std2, void* out, int out_kind, void* workspace,
                           int64_t workspace_bytes, void* stream);
/* `tensor_to_pil` (mingunivision/modeling_bailing_moe.py:84-90, test_infer_recon_image.py:24-28): img [n, 3, h, w]
 * (bf16, or fp32 with img_is_fp32) -> out [n, h, w, 3] u8 = trunc((x * std + mean) * 255), fp32 steps rounded
 * separately, ToPILImage's truncation toward zero (saturating outside [0, 255]). */
int mb_image_postprocess_u8(const void* img, int img_is_fp32, int n, int h, int w, float mean0, float mean1,
                            float mean2, float std0, float std1, float std2, void* out, void* stream);
/* The tail of forward_pixel_decoder (modeling_mingtok.py:190-196: unpatchify, clamp_(-1, 1)) fused with tensor_to_pil:
 * x [B, g*g, p*p*3] bf16 (rows of the head GEMM, channel-last inside the patch) -> out [B, g*p, g*p, 3] u8, identical
 * to mb_unpatchify_clamp followed by mb_image_postprocess_u8; the image never exists as a float tensor and 3 bytes per
 * pixel leave the device. */
int mb_unpatchify_to_u8(const void* x, void* out, int B, int g, int p, float mean0, float mean1, float mean2,
                        float std0, float std1, float std2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MINGB200_H_ */
