/*
 * stabletriton_b200 -- C ABI of the B200 (sm_100a) kernels behind `stabletriton_b200.compile(model)`.
 *
 * This is the drop-in boundary: each entry point is what the reference's
 * torch.fx-wrapped op wrappers bind for the SDXL UNet denoise step.  Conventions, all entry points:
 *   - plain device pointers + explicit sizes / strides (elements, not bytes); no torch types;
 *   - bf16 storage, fp32 statistics / accumulation;
 *   - activations are NHWC ("channels last"): a conv/GroupNorm tensor is [N, H*W, C] with C contiguous,
 *     a token tensor is [B*T, C] row-major -- the same memory, so the reference's NCHW<->(B,HW,C)
 *     permutes (optimizers/unet_pt.py:228-241) are free;
 *   - caller-owned outputs and workspaces: no allocation, no synchronisation, no host reads of device
 *     memory -> every call is CUDA-graph capturable on `stream`;
 *   - returns ST_OK (0) or a negative error code; never throws; st_last_error_string() explains.
 */
#ifndef STABLETRITON_B200_H_
#define STABLETRITON_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ST_OK 0
#define ST_ERR_INVALID_ARGUMENT (-1)
#define ST_ERR_CUDA (-2)
#define ST_ERR_UNSUPPORTED (-3)

#define ST_VERSION 100 /* 0.1.0 */

/* Epilogue flags for st_gemm_bf16 / st_conv3x3_nhwc_bf16 */
#define ST_EPI_SILU 1u  /* y = silu(acc + bias)                   (reference: kernels/linear.py:155-157) */
#define ST_EPI_GEGLU 2u /* y = (acc_s + b_s) * gelu_erf(acc_g + b_g), B = [state rows ; gate rows]
                           (reference: unet_pt.py:155-158 + kernels/geglu.py:11-26)                       */
#define ST_W_STATIC 4u  /* hint: the weight operand is not written by the kernel launched just before this one on
                           the stream, so its first tiles may be fetched before the programmatic (PDL) dependency on
                           that kernel resolves.  Results are identical with or without it.                       */

#define ST_EPI_F32OUT 8u /* st_gemm_bf16 only: D is fp32 [M, ldd] (ldd in fp32 elements, N % 4 == 0); not with GEGLU /
                            GroupNorm partials.  For results that feed a row softmax (the 512-wide VAE attention head). */

typedef void* st_stream_t; /* a cudaStream_t / CUstream */

/* ---- library ---------------------------------------------------------------------------------- */
int st_version(void);
const char* st_last_error_string(void);
/* Number of kernels this library has launched in this process (thread-safe enough for accounting). */
unsigned long long st_launch_count(void);
void st_reset_launch_count(void);

/* Optional fp32 scratch for stream-K GEMMs (small grids with a long K loop: the tiles x k-blocks space is cut into
 * one equal range per SM, partial tiles are merged through this buffer).  Caller-owned, one per device, shared by
 * all GEMM / conv launches on one stream; st_workspace_bytes() is the size to allocate.  Stream-K is
 * experimental and opt-in (environment ST_ENABLE_STREAMK=1): by default every tile is computed by one CTA. */
size_t st_workspace_bytes(void);
int st_set_workspace(void* ptr, size_t bytes);

/* ---- GroupNorm (+SiLU), NHWC ------------------------------------------------------------------
 * Replaces groupnorm_wrapper(input, num_groups, weight, bias, eps, activation)
 * (reference: kernels/groupnorm.py:128-161; wrapper optimizers/replace_groupnorm.py:18-19) with
 * torch.nn.GroupNorm semantics (biased variance) on 4-D input, which the reference kernel gets
 * wrong.  x, y: [N, HW, C] bf16; gamma, beta: [C] bf16; C % groups == 0, C % 8 == 0.
 * workspace: st_groupnorm_workspace_bytes() bytes of scratch, 16-byte aligned, owned by THIS call (per-image arrival
 * tickets, partial statistics, per-channel scale / shift all live there -- no library-global state, so concurrent calls
 * on different streams or in different captured graphs never interact).  Contents on entry are irrelevant. */
size_t st_groupnorm_workspace_bytes(int N, int HW, int C, int groups);
int st_groupnorm_nhwc_bf16(const void* x, void* y, const void* gamma, const void* beta, void* workspace, int N,
                           int HW, int C, int groups, float eps, int apply_silu, st_stream_t stream);

/* GroupNorm whose statistics were already emitted by the producer(s) of x: st_gemm_bf16 / st_conv3x3_nhwc_bf16 called
 * with `gn_partial` write, per 128-row tile and output column, (mean, M2) of the values they store; this entry point
 * merges them per (image, group) and normalises -- x is read ONCE (the reference's kernel, and the stand-alone entry
 * point above, read it twice: kernels/groupnorm.py:24-119).  Needs HW % 128 == 0.  part_a: [N*HW/128, C_a, 2] fp32;
 * part_b (may be NULL with C_b = 0): [N*HW/128, C_b, 2] for x = concat(a, b) along channels (unet_pt.py:356,385), whose
 * statistics are those of its two producers; C_a + C_b == C.  workspace: as above (only N*C*2 floats are used). */
int st_groupnorm_from_partials_nhwc_bf16(const void* x, void* y, const void* gamma, const void* beta, void* workspace,
                                         int N, int HW, int C, int groups, float eps, int apply_silu,
                                         const void* part_a, int C_a, const void* part_b, int C_b,
                                         st_stream_t stream);

/* ---- LayerNorm over the last dimension --------------------------------------------------------
 * Replaces layer_norm(x, weight, bias, eps) (reference: kernels/layer_norm.py:338-346, kernel
 * :114-205; wrapper optimizers/replace_layernorm.py:17-24).  x, y: [M, N] bf16 with row pitch ldx/ldy;
 * gamma/beta: [N] bf16 (beta may be NULL); N % 8 == 0, N <= 4096.
 * gamma / beta are treated as parameters (like W under ST_W_STATIC): they are fetched before the programmatic (PDL)
 * dependency on the kernel launched just before this one resolves, so they must not be written by that kernel. */
int st_layernorm_bf16(const void* x, int ldx, void* y, int ldy, const void* gamma, const void* beta, int M, int N,
                      float eps, st_stream_t stream);

/* ---- GEGLU, standalone elementwise -------------------------------------------------------------
 * Replaces geglu_wrapper(state, gate) (reference: kernels/geglu.py:28-35): out = state * gelu_erf(gate).
 * state/gate/out: [rows, cols] bf16 with row pitches; cols % 8 == 0.  (The hot path fuses this into
 * the projection GEMM via ST_EPI_GEGLU; this entry point exists for API parity and tests.) */
int st_geglu_bf16(const void* state, int ld_state, const void* gate, int ld_gate, void* out, int ld_out, int rows,
                  int cols, st_stream_t stream);

/* ---- Linear / GEMM on tcgen05 tensor cores -----------------------------------------------------
 * Replaces sdxl_forward(x, weight, bias, activation) (reference: kernels/linear.py:173-222):
 *   D[M, n_out] = epi(A[M, K] . W[N, K]^T + bias[N]) (+ residual[M, n_out])
 * A: activations, row pitch lda; W: nn.Linear weight layout (N rows of K), row pitch ldw; K % 64 == 0,
 * lda/ldw % 8 == 0, pointers 16-byte aligned.  With ST_EPI_GEGLU, N = 2*n_out and D has n_out columns;
 * otherwise n_out = N.  bias / residual may be NULL.  block_n: 0 = choose automatically.
 * gn_partial (may be NULL): fp32 [M/128, n_out, 2]; the epilogue also writes (mean, M2) of every output column over
 * the 128 rows of each row tile, computed from the bf16 values it stores -- the input of
 * st_groupnorm_from_partials_nhwc_bf16.  Needs M % 128 == 0, no GEGLU. */
int st_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* D, int ldd, int M, int N, int K,
                 const void* bias, const void* residual, int ldr, unsigned flags, int block_n, void* gn_partial,
                 st_stream_t stream);

/* Tiny-M Linear (time / added-condition embeddings, M <= 32): y = act_out(act_in(x) . W^T + b).
 * CUDA-core, weight-bandwidth bound.  silu_in applies SiLU to x on load (unet_pt.py:81-82).  flags: ST_W_STATIC
 * declares W a parameter (not written by the kernel launched just before this one): its rows are then requested
 * before the programmatic dependency on the preceding kernel resolves; without the flag W is read after it. */
int st_linear_small_m_bf16(const void* x, int ldx, const void* W, int ldw, const void* bias, void* y, int ldy, int M,
                           int N, int K, int silu_in, int silu_out, unsigned flags, st_stream_t stream);

/* ---- 3x3 convolution, pad 1, stride 1, NHWC, implicit GEMM on tcgen05 --------------------------
 * Replaces implicit_gemm_fprop(a NHWC, b KRSC) (reference: kernels/Conv_Kernels/conv_implicit_gemm.py:
 * 143-182) and torch.nn.Conv2d at the sites unet_pt.py:58-60,64-66,260-262.
 * x: [N, H, W, C]; w: [K, 3, 3, C] (KRSC == channels-last Conv2d weight); y: [N, H, W, K].
 * y = conv(x, w) + bias[K] (+ temb[N, K] broadcast over pixels, unet_pt.py:82-83) (+ residual[N,H,W,K],
 * unet_pt.py:93).  Needs C % 64 == 0, K % 8 == 0 and either (H*W) % 128 == 0 with W dividing 128 or a multiple of it, or
 * H*W dividing 128 (small feature maps: one tile covers several whole images).
 * gn_partial (may be NULL): fp32 [N*H*W/128, K, 2], as in st_gemm_bf16 (needs N*H*W % 128 == 0). */
int st_conv3x3_nhwc_bf16(const void* x, const void* w, const void* bias, void* y, int N, int H, int W, int C, int K,
                         const void* temb, int ld_temb, const void* residual, unsigned flags, int block_n,
                         void* gn_partial, st_stream_t stream);

/* Small-channel direct 3x3 conv (pad 1, stride 1) for conv_in (C=4 -> 320) and conv_out (320 -> 4)
 * (unet_pt.py:430,467).  CUDA-core.  Either C <= 8 (then K % 8 == 0, y dense NHWC, x addressed through
 * element strides xs_* so an NCHW latent is consumed in place) or K <= 8 (then C % 8 == 0, x dense NHWC,
 * y addressed through ys_* so the result lands directly in an NCHW tensor).  w: [K, 3, 3, C]. */
int st_conv3x3_direct_bf16(const void* x, long long xs_n, long long xs_h, long long xs_w, long long xs_c,
                           const void* w, const void* bias, void* y, long long ys_n, long long ys_h, long long ys_w,
                           long long ys_c, int N, int H, int W, int C, int K, st_stream_t stream);

/* im2col for 3x3 / pad 1 / stride s (used for the two stride-2 downsamplers, unet_pt.py:249-251):
 * col: [N*Ho*Wo, 9*C], tap-major (r, s, c) to match the KRSC weight; then st_gemm_bf16. */
int st_im2col3x3_nhwc_bf16(const void* x, void* col, int N, int H, int W, int C, int stride, st_stream_t stream);

/* conv_in on the tensor cores: im2col for tiny C (9*C <= 64) into [N*H*W, 64] rows (tap-major, zero padded),
 * then st_gemm_bf16 against the weight padded to (K, 64).  x through element strides, like st_conv3x3_direct_bf16. */
int st_im2col3x3_smallc_bf16(const void* x, long long xs_n, long long xs_h, long long xs_w, long long xs_c, void* col,
                             int N, int H, int W, int C, st_stream_t stream);

/* conv_out on the tensor cores: st_conv3x3_nhwc_bf16 with the 4 output channels padded to 8, then this
 * [N*HW, ld] -> dense NCHW (first C channels) transposer, so the scheduler sees a standard latent. */
int st_nhwc_to_nchw_bf16(const void* src, int ld, void* dst, int N, int HW, int C, st_stream_t stream);

/* Nearest-neighbour 2x upsample, NHWC (F.interpolate(scale_factor=2, mode="nearest"), unet_pt.py:265). */
int st_upsample_nearest2x_nhwc_bf16(const void* x, void* y, int N, int H, int W, int C, st_stream_t stream);

/* ---- Multi-head attention forward (flash, online softmax), head_dim 64 --------------------------
 * Implements the *pattern* of fuse_attention (reference: optimizers/replace_attention.py:76-86):
 * per head softmax(Q K^T * scale) V, heads = channels [64h, 64h+64) of (B, T, H*64) tensors, no mask.
 * Every tensor is addressed as [b][h][t][d] through element strides (sb, sh, st) with d contiguous:
 * (B, T, H*64) activations use sh = 64, st = row pitch, sb = T*pitch (so q/k/v may be column slices of one
 * fused QKV buffer); the reference's (B, H, T, D) layout (kernels/attention_fa2.py:113-140) uses
 * st = 64, sh = T*64.  Strides % 8 == 0.  Separate Tq / Tk and a masked K tail make cross-attention
 * (Tk = 77) work. */
int st_attention_bf16(const void* q, long long q_sb, long long q_sh, long long q_st, const void* k, long long k_sb,
                      long long k_sh, long long k_st, const void* v, long long v_sb, long long v_sh, long long v_st,
                      void* o, long long o_sb, long long o_sh, long long o_st, int B, int H, int Tq, int Tk,
                      float scale, st_stream_t stream);

/* ---- elementwise glue -------------------------------------------------------------------------
 * Sinusoidal timestep embedding (unet_pt.py:22-36): out[b, :] = cat(cos(t_b * f_i), sin(t_b * f_i)),
 * f_i = exp(-ln(10000) * i / half), t fp32 [B], out bf16 [B, 2*half] with row pitch ldo. */
int st_timestep_embedding_bf16(const float* t, void* out, int ldo, int B, int half, st_stream_t stream);

/* Channel concat of two NHWC tensors (torch.cat(dim=1) in NCHW terms, unet_pt.py:356,385):
 * y[p, :Ca] = a[p, :], y[p, Ca:] = b[p, :]; P pixels (N*H*W); Ca, Cb % 8 == 0. */
int st_concat_channels_bf16(const void* a, int Ca, const void* b, int Cb, void* y, long long P, st_stream_t stream);

/* ---- Euler-discrete scheduler + classifier-free guidance, device-resident loop state -------------
 * (the step on either side of the UNet, so a whole denoise step is one graph
 * replay.)  sigmas: fp32 [steps + 1] on the device; step: int32 on the device, advanced by
 * st_advance_step, so the captured graph is identical for every step.
 *   st_scale_model_input: model_in[c, :] = bf16(x / sqrt(sigma[step]^2 + 1)), c < copies (the CFG pair)
 *   st_euler_cfg_update : eps = eps_u + g (eps_c - eps_u); x += (sigma[step+1] - sigma[step]) * eps
 *                         (eps_cond may be NULL: no guidance)
 *   st_advance_step     : ++*step; if t_out: *t_out = timesteps[*step] */
int st_scale_model_input(const float* x, void* model_in, long long n, int copies, const float* sigmas,
                         const int* step, st_stream_t stream);
int st_euler_cfg_update(const void* eps_uncond, const void* eps_cond, float* x, long long n, float guidance,
                        const float* sigmas, const int* step, st_stream_t stream);
int st_advance_step(int* step, float* t_out, const float* timesteps, st_stream_t stream);

/* 1x1 convolution between tiny channel counts (Ci, Co <= 8), dense NCHW in and out, input pre-scaled by in_scale:
 * y[n, o, :] = b[o] + sum_i w[o, i] * in_scale * x[n, i, :].  The VAE's post_quant_conv on the latent (Diffusers
 * AutoencoderKL.decode: z / scaling_factor -> post_quant_conv); w is [Co, Ci] bf16. */
int st_pointwise_conv_small_bf16(const void* x, const void* w, const void* bias, void* y, int N, long long HW, int Ci,
                                 int Co, float in_scale, st_stream_t stream);

/* ---- row softmax (VAE mid-block attention) -------------------------------
 * P[i, :] = softmax(scale * S[i, :]) for fp32 scores S [M, lds] (from st_gemm_bf16 with ST_EPI_F32OUT) into bf16
 * probabilities P [M, ldp]; N % 4 == 0.  The SDXL VAE decoder's attention has ONE head of width 512 over H*W tokens
 * (Diffusers AutoencoderKL mid block): Q K^T and P V run on st_gemm_bf16, this kernel sits between them. */
int st_softmax_rows_f32_bf16(const float* S, long long lds, void* P, long long ldp, int M, int N, float scale,
                             st_stream_t stream);

/* ---- 2-GPU CFG split: eps exchange over NVLink peer memory, fused with the Euler update -----------
 * (one prompt on two GPUs, rank 0 = uncond row, rank 1 = cond row; the reference has no
 * multi-GPU path.)  Each rank owns an exchange slab of st_peer_slab_bytes(n) bytes obtained from st_peer_alloc
 * (cudaMalloc, zeroed), exports it with st_peer_export (a 64-byte CUDA IPC handle the host code hands to the other
 * process, e.g. through torch.distributed) and maps the peer's slab with st_peer_import.
 *   st_cfg_exchange_euler_update: ONE kernel that stores this rank's eps row (n bf16) into the peer's slab with P2P
 *   stores, publishes a sequence number there (release.sys), waits for the peer's row in its own slab (acquire.sys)
 *   and applies x += (sigma[step+1] - sigma[step]) * (eps_u + g (eps_c - eps_u)) to its replica of the latents.
 *   Capturable; both ranks must launch it the same number of times (slab parity = a device-resident epoch counter).
 *   A peer that never publishes is given 2 s, then the launch finishes and st_peer_error reports the exchange. */
size_t st_peer_slab_bytes(long long n);
int st_peer_alloc(size_t bytes, void** ptr);
int st_peer_free(void* ptr);
int st_peer_export(void* ptr, unsigned char* handle64);
int st_peer_import(const unsigned char* handle64, void** ptr);
int st_peer_close(void* ptr);
int st_cfg_exchange_euler_update(const void* eps_local, int row, void* slab_local, void* slab_peer, float* x, long long n,
                                 float guidance, const float* sigmas, const int* step, st_stream_t stream);
int st_peer_error(const void* slab_local, unsigned* out);

#ifdef __cplusplus
}
#endif
#endif /* STABLETRITON_B200_H_ */
