/* tclight.h — C ABI of libtclight.so, the sm_100a implementation of TC-Light's two hot paths.
 *
 * The reference (Linketic/TC-Light) has no FFI layer: its "operator API" for these paths is the
 * set of Python callables listed in SURVEY.md §8(b).  Each entry point below replaces the
 * PyTorch/diffusers op sequence behind one of those callables; the citation names the reference
 * lines it stands in for (paths relative to the reference root).  The Python mirror in
 * tclight_b200/ binds these with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative TCL_ERR_* code; tcl_last_error() gives
 *     the message of the last failure on the calling thread;
 *   - all pointers are DEVICE pointers unless a parameter name ends in _host;
 *   - functions enqueue work on `stream` and never synchronise, allocate or take ownership;
 *   - "16-bit" activations are fp16 or bf16, selected by a TCL_DTYPE_* argument;
 *   - image activations are NHWC, token activations are [tokens, channels] row-major.
 */
#ifndef TCLIGHT_H_
#define TCLIGHT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* tcl_stream_t; /* == cudaStream_t */

#define TCL_DTYPE_FP16 0
#define TCL_DTYPE_BF16 1

/* ---- library ---------------------------------------------------------------------------- */
const char* tcl_last_error(void);
int tcl_version(void);
/* number of kernel launches issued through this library since the last reset (bench.py's
 * gpu_launches counter) */
long long tcl_launch_count(void);
void tcl_launch_count_reset(void);

/* ---- implicit GEMM (tcgen05) ------------------------------------------------------------
 * Replaces cuDNN/cuBLAS behind diffusers' ResnetBlock2D convs, Downsample2D/Upsample2D convs,
 * Transformer2DModel proj_in/proj_out, Attention.to_q/k/v/out and GEGLU/FF linears
 * (in-repo restatements utils/VidToMe/pnp_utils.py:40-97,
 * 110-164).  out[pixel, n] = sum_k A[pixel, k] * W[n, k] with K described as up to four
 * segments (NHWC tensor, taps): taps==9 is a 3x3 window with zero padding 1, taps==1 a 1x1.
 */
#define TCL_IGEMM_MAX_SRC 4
#define TCL_EPI_NHWC 0   /* out[pixel*out_pitch + n] = (acc + bias[n] + residual) * out_scale   */
#define TCL_EPI_GEGLU 1  /* weights interleaved per 256 rows (128 value | 128 gate);
                            out[pixel, N/2] = value * gelu(gate)                                 */
#define TCL_EPI_HEADS 2  /* N = nsec * sec_cols; section i written to sec_ptr[i] either as
                            [b, head, tok_pitch, d_pad] (sec_vt[i]==0) or transposed
                            [b, head, d_pad, tok_pitch] (sec_vt[i]==1)                           */

typedef struct {
  const void* ptr; /* NHWC 16-bit, already offset to the first channel of this segment */
  int64_t n, h, w; /* INPUT image grid */
  int64_t c;       /* channels in this segment (multiple of 64) */
  int64_t pitch;   /* elements between consecutive pixels (>= c, multiple of 8) */
  int32_t taps;    /* 1 or 9 */
  int32_t stride;  /* 1 or 2: input coordinate = stride * output coordinate + tap - pad */
  int32_t no_lead_pad; /* 3x3 only: 0 = zero padding 1 on every side; 1 = no padding before the first row/column,
                          zero fill after the last (the F.pad(x, (0,1,0,1)) + stride-2 conv of the VAE encoder's
                          Downsample2D) */
  int32_t reserved_;
} tcl_igemm_src;

typedef struct {
  int32_t dtype;
  int32_t num_src;
  tcl_igemm_src src[TCL_IGEMM_MAX_SRC];
  int32_t n_img, out_h, out_w; /* OUTPUT pixel grid; a token matrix [M,K] is n_img=1,out_h=1,out_w=M */
  int32_t N;                   /* output columns */
  int64_t K;                   /* sum over segments of taps*c; weight is [N, K] row-major 16-bit */
  const void* weight;
  const float* bias; /* [N] fp32 or NULL */
  int32_t mode;      /* TCL_EPI_* */
  /* TCL_EPI_NHWC / TCL_EPI_GEGLU */
  void* out;
  int64_t out_pitch;
  const void* residual; /* NHWC 16-bit or NULL (TCL_EPI_NHWC only) */
  int64_t res_pitch;
  float out_scale; /* 0 is treated as 1 */
  /* TCL_EPI_HEADS */
  void* sec_ptr[3];
  int32_t sec_vt[3];
  int32_t sec_cols, heads, d, d_pad;
  int64_t tok_per_batch, tok_pitch;
} tcl_igemm_desc;

int tcl_igemm(const tcl_igemm_desc* desc, tcl_stream_t stream);

/* ---- attention (tcgen05) ----------------------------------------------------------------
 * Replaces diffusers Attention + AttnProcessor2_0 -> F.scaled_dot_product_attention
 * (utils/model_utils.py:66; restated utils/VidToMe/pnp_utils.py:40-97; called from
 * utils/VidToMe/vidtome/patch.py:157, 178).  Operands are the head-split outputs of tcl_igemm
 * (TCL_EPI_HEADS): q [batch*heads, tq_pitch, d_pad], k [kv_batch*heads, tk_pitch, d_pad],
 * vt [kv_batch*heads, d_pad, tk_pitch], zero padded beyond d.  kv_batch = batch/kv_batch_div
 * (cross-attention: every frame of a CFG half shares one text embedding, generate.py:295).
 * out is token-major [batch, tq, heads*d].  Single-pass online softmax (fp32 statistics, P rounded to the
 * 16-bit type before P V as in flash attention), scale 1/sqrt(d).
 */
typedef struct {
  int32_t dtype;
  int32_t batch, heads;
  int32_t tq, tk;
  int32_t d, d_pad; /* d_pad in {64, 128, 192} */
  int32_t kv_batch_div;
  int64_t tq_pitch, tk_pitch;
  const void* q;
  const void* k;
  const void* vt;
  void* out;
  void* workspace;        /* optional (may be NULL): >= tcl_attention_workspace_bytes() of device scratch.  With it, long
                             self-attention launches cut the work items of their last, partial wave of CTAs into slices of
                             the key range that run side by side, and merge the partial (O, max, denominator) results   */
  size_t workspace_bytes;
} tcl_attn_desc;

size_t tcl_attention_workspace_bytes(void);
int tcl_attention(const tcl_attn_desc* desc, tcl_stream_t stream);
/* ---- normalisation / staging (HBM-bound) --------------------------------------------------
 * GroupNorm(+SiLU) of diffusers ResnetBlock2D / Transformer2DModel / conv_norm_out over NHWC,
 * reading the channel concat [x1 | x2] of an up-block skip connection on the fly (x2 may be
 * NULL with c2 == 0).  stats_ws: 2*groups*n_img floats of scratch.  (SURVEY.md §8a row A5;
 * utils/VidToMe/pnp_utils.py:110-164.)
 */
int tcl_groupnorm(int dtype, const void* x1, int c1, const void* x2, int c2, int n_img,
                  long long pix_per_img, int groups, const float* gamma, const float* beta, float eps,
                  int silu, float* stats_ws, void* out, tcl_stream_t stream);

/* LayerNorm over the last dim of [rows, C] (BasicTransformerBlock.norm1/2/3,
 * utils/VidToMe/vidtome/patch.py:146, 170, 187). */
int tcl_layernorm(int dtype, const void* x, long long rows, int C, const float* gamma, const float* beta,
                  float eps, void* out, tcl_stream_t stream);

/* F.interpolate(mode="nearest") to (oh, ow) over NHWC (diffusers Upsample2D). */
int tcl_upsample_nearest(const void* x, int n, int h, int w, int c, int oh, int ow, void* out,
                         tcl_stream_t stream);

#define TCL_LATENT_FP32 0
#define TCL_LATENT_FP16 1
#define TCL_LATENT_BF16 2
/* pred_noise input staging (generate.py:298 cat([x, x]); utils/model_utils.py:35-40 concat of the
 * condition latent): x / cond are strided 4-channel latent views, strides xs/cs = HOST arrays
 * {image, channel, row, col} in elements; out = NHWC [2F, H, W, 64] 16-bit (channels 8..63 zero);
 * duplicate == 0 writes only the first F images. */
int tcl_stage_latent(int dtype, int latent_dtype, const void* x, const long long* xs_host, const void* cond,
                     const long long* cs_host, int F, int H, int W, int duplicate, void* out, tcl_stream_t stream);

/* CFG combine (generate.py:349-350) of the UNet output eps NHWC [2F, H, W, pitch] into a strided
 * 4-channel latent view (os_host as above). */
int tcl_cfg_store(int dtype, int latent_dtype, const void* eps, int pitch, float guidance_scale, int F, int H,
                  int W, void* out, const long long* os_host, tcl_stream_t stream);

/* y[N] = W[N,K] (16-bit) * f(x[K]) + bias, one vector: TimestepEmbedding linear_1/linear_2 and the
 * per-resnet time_emb_proj(SiLU(temb)) (utils/VidToMe/pnp_utils.py:135-137).  x, bias, y
 * are fp32; silu_in applies SiLU to x; round16 rounds like a 16-bit nn.Linear would. */
int tcl_gemv(int dtype, const void* W, const float* x, const float* bias, int N, int K, int silu_in, int round16,
             float* y, tcl_stream_t stream);

/* ---- sampler tail (HBM-bound, latent dtype = TCL_LATENT_*) ----------------------------------
 * adain_blend: noises_t <- AdaIN(noises_t, noises); noises <- sqrt(alpha) noises_t + sqrt(1-alpha) noises
 *   over `planes` = N*4 planes of `plane_elems` = h*w (generate.py:281-282; calc_mean_std /
 *   adaptive_instance_normalization, utils/general_utils.py:137-156: unbiased var + 1e-5).
 * scale_inplace: x *= s  (overlap rescale of later temporal windows, generate.py:276-278).
 * dpm_step: diffusers DPMSolverMultistepScheduler.step (sde-dpmsolver++, midpoint) for one step:
 *   x0 = (x - sigma_c_hat*eps)/alpha_c_hat;  x' = A x + B x0 [+ 0.5 B (x0 - x0_prev) inv_r0] + Cn z
 *   with host-computed fp32 coefficients.  Writes x0 (history) and x'.
 */
int tcl_adain_blend(int latent_dtype, void* noises_t, void* noises, int planes, int plane_elems, float alpha,
                    tcl_stream_t stream);
int tcl_scale_inplace(int latent_dtype, void* x, long long n, float s, tcl_stream_t stream);
int tcl_dpm_step(int latent_dtype, const void* eps, const void* x, const void* x0_prev, const float* z,
                 void* x0_out, void* x_out, long long n, float sigma_c_hat, float alpha_c_hat, float A, float B,
                 float Cn, float inv_r0, int second_order, tcl_stream_t stream);

/* DDIM step in either direction (invert.py:215-244 Inverter.pred_next_x):
 *   x_out = mu_out * ((x - sig_in * eps) / mu_in) + sig_out * eps
 * inversion: (mu_in, sig_in) = (sqrt(a_prev), sqrt(1-a_prev)), (mu_out, sig_out) = (sqrt(a_t), sqrt(1-a_t));
 * sampling: the two pairs swapped.  Each tensor op rounds to the latent dtype like the reference expression. */
int tcl_ddim_next(int latent_dtype, const void* eps, const void* x, void* x_out, long long n, float mu_in, float sig_in,
                  float mu_out, float sig_out, tcl_stream_t stream);

/* ---- VAE support (HBM-bound; utils/VidToMe/generate_utils.py:140-172) -------------------
 * softmax_rows : in-place softmax over the first `cols` entries of every row of a 16-bit [rows, pitch] matrix
 *                (entries cols..pitch-1 are set to 0): the key axis of AutoencoderKL's single-head mid-block attention,
 *                whose scores come from tcl_igemm.
 * image_to_nhwc: out[b, y, x, c] = src[b, c, y, x] * scale + shift for c < C, 0 for C <= c < c_pad
 *                (src_dtype / out_dtype are TCL_LATENT_* codes); nhwc_to_image is the inverse with an optional clamp.
 */
int tcl_softmax_rows(int dtype, void* x, long long rows, int cols, int pitch, tcl_stream_t stream);
int tcl_image_to_nhwc(int dtype, int src_dtype, const void* src, int B, int C, int H, int W, int c_pad, float scale,
                      float shift, void* out, tcl_stream_t stream);
int tcl_nhwc_to_image(int dtype, const void* in, int B, int C, int H, int W, int pitch, float scale, float shift,
                      int do_clamp, float lo, float hi, int out_dtype, void* out, tcl_stream_t stream);

/* ---- VidToMe token merging ------------------------------------------------------------------
 * bipartite_soft_matching_randframe (utils/VidToMe/vidtome/merge.py:20-159) and
 * bipartite_soft_matching_2s (:343-463) as called by compute_merge (patch.py:14-91).
 *
 * normalize_split: merge.py:84-85 — metric / metric.norm(dim=-1), then split into src (a) and
 *   dst (b).  Tokens are [x0 | x1] along the token axis (x1 may be NULL: n1 == 0); dst is the
 *   contiguous token range [d0, d1), src every other token in order.  a_out [batch, n_src, C],
 *   b_out [batch, n_dst, C].
 * match: merge.py:87-97 — scores = a b^T rounded to the 16-bit type, node_max/node_idx = max over
 *   dst (lowest index among ties).  align_batch != 0 concatenates the dst axes of the batch
 *   samples (node_idx in [0, batch*n_dst), outputs [n_src]); otherwise outputs are [batch, n_src].
 *   The score matrix is never written to memory.
 * plan: merge.py:98-108 + the index algebra of merge()/unmerge() (:119-155) turned into two gather
 *   maps shared by all batch samples: merge_map[n_src - r + n_dst] (merged row -> token) and
 *   unmerge_map[n_src + n_dst] (token -> merged row).  `edge` is argsort(node_max, descending).
 * gather_rows: out[b, i, :] = [x0 | x1][b, map[i], :] (+ add[b, i, :]) — merge, unmerge and the
 *   residual add of patch.py:168-169 in one pass.
 */
size_t tcl_vidtome_match_workspace_bytes(int batch, int n_src);
int tcl_vidtome_normalize_split(int dtype, const void* x0, long long n0, const void* x1, long long n1, int batch,
                                int C, long long d0, long long d1, void* a_out, void* b_out, tcl_stream_t stream);
int tcl_vidtome_match(int dtype, const void* a, const void* b, int batch, int n_src, int n_dst, int C,
                      int align_batch, float* node_max, long long* node_idx, void* workspace,
                      size_t workspace_bytes, tcl_stream_t stream);
int tcl_vidtome_plan(const long long* edge, const long long* node_idx, int n_src, int n_dst, int r, long long d0,
                     int* merge_map, int* unmerge_map, tcl_stream_t stream);
int tcl_gather_rows(int dtype, const void* x0, long long n0, const void* x1, long long n1, const int* map,
                    int map_per_batch, long long n_out, int batch, int C, const void* add, void* out,
                    tcl_stream_t stream);

/* ---- two-stage temporal-consistency optimiser (HBM-bound, fp32) ------------------------------
 * One call = one optimiser iteration of the reference loops, no host synchronisation:
 *   tcl_exposure_iteration  Generator.exposure_align body, generate.py:392-436 (stage 1)
 *   tcl_uvt_iteration       Generator.unique_tensor_optimization body, generate.py:494-517 (stage 2)
 * covering OptDataset.__getitem__ (utils/dataloader.py:29-36), warp_flow (utils/flow_utils.py:5-16),
 * l1_loss / relaxed_ms_ssim(start_level=1, data_range=1) / TVLoss (utils/loss_utils.py:25, 73-211,
 * 324-339), SH2RGB (utils/sh_utils.py:117-118), index_select + its index_add backward, and
 * torch.optim.Adam.step + zero_grad.
 *   grad     : stage 2: [U,4] fp32, ZERO-FILLED once by the caller (a row is {dR, dG, dB, padding} so that every
 *              scatter is one 16-byte reduction); stage 1: [N,12].  The calls leave it zeroed.
 *   idx_host : HOST array of the batch's frame indices (what the DataLoader's sampler drew)
 *   ids      : unq_inv as int32 [N*H*W] (data_parser.unq_inv, video_dataparser.py:59)
 *   loss_out : device float[3] = {loss, loss_flow, loss_photometric} of this iteration (or NULL)
 *   step     : 1-based Adam step count
 */
#define TCL_POSTOPT_MAX_BATCH 32
typedef struct {
  int32_t N, H, W;
  const float* edited;     /* [N,3,H,W] OptDataset.edited_images                            */
  const float* past_flows; /* [N,2,H,W]                                                      */
  const float* mask_bwd;   /* [N,1,H,W] soft masks                                           */
  const float* ypyr;       /* target side from tcl_postopt_build_pyramid (pyramid levels 1..4 and the
                              constant SSIM maps G*y, G*y^2): [N,3,tcl_postopt_target_elems(H,W)] */
  float lambda_dssim, lambda_flow, lambda_tv;
  int32_t max_batch;       /* largest n_batch that will be used with this workspace          */
  int32_t norm_batch;      /* data parallel: GLOBAL batch size the loss means are taken over
                              (0 = the call's own batch)                                     */
  int32_t norm_valid;      /* data parallel: GLOBAL number of batch frames with index > 0    */
  void* workspace;         /* >= tcl_postopt_workspace_bytes(H, W, max_batch), ZERO-FILLED once
                              by the caller before the first iteration                        */
  size_t workspace_bytes;
} tcl_postopt_ctx;

size_t tcl_postopt_workspace_bytes(int H, int W, int max_batch);
long long tcl_postopt_pyramid_elems(int H, int W);
long long tcl_postopt_target_elems(int H, int W);
int tcl_postopt_build_pyramid(const float* edited, int N, int H, int W, float* ypyr, tcl_stream_t stream);
int tcl_uvt_iteration(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, const int* ids, long long U,
                      float* fdc, float* grad, float* m, float* v, float lr, float beta1, float beta2, float eps,
                      int step, float* loss_out, tcl_stream_t stream);
int tcl_exposure_iteration(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, float* exposure, float* grad,
                           float* m, float* v, float lr, float beta1, float beta2, float eps, int step,
                           float* loss_out, tcl_stream_t stream);
/* Data-parallel split (each rank takes a slice of the batch; loss_out = this rank's additive share
 * of {loss, flow, photometric}; ctx.norm_batch / norm_valid carry the GLOBAL normalisers).
 *
 * Stage 1: tcl_exposure_gradient -> all-reduce of the [N,12] gradient (1.2 KB/frame) -> tcl_adam_step on every rank.
 *
 * Stage 2: the UVT rows and their gradient are SHARDED by row range over the ranks of one NVSwitch box, and every
 * shard is mapped into every rank's address space (CUDA IPC, tcl_ipc_open).  tcl_uvt_gradient_sharded gathers rows
 * with plain (peer) loads and scatters gradients with 16-byte (peer) reductions from inside the gather / level-0
 * kernels, so only the rows a batch touches cross NVLink; after a cross-rank barrier each rank runs
 * tcl_adam_step_uvt on ITS shard (dense Adam, 1/world of the rows), and a second barrier orders it before the next
 * gather.  Row id -> (owner = id / rows_per_rank, local = id % rows_per_rank).  No dense all-reduce, no replicated
 * Adam.  Replaces the autograd index_select / index_add_ + optimizer.step of generate.py:496-517. */
#define TCL_MAX_RANKS 8
typedef struct {
  int32_t world, rank;
  int64_t rows_per_rank;        /* ceil(U / world) rounded up to a multiple of 4 (>= 256 when world > 1)        */
  float* fdc[TCL_MAX_RANKS];    /* shard r: [rows_per_rank,3]; entry `rank` is local, the others peer mappings   */
  float* grad[TCL_MAX_RANKS];   /* shard r: [rows_per_rank,4], zero-filled once by its owner                     */
} tcl_uvt_shards;
int tcl_uvt_gradient_sharded(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, const int* ids,
                             const tcl_uvt_shards* shards, float* loss_out, tcl_stream_t stream);
int tcl_exposure_gradient(const tcl_postopt_ctx* ctx, const int* idx_host, int n_batch, const float* exposure, float* grad,
                          float* loss_out, tcl_stream_t stream);
int tcl_adam_step(float* p, float* grad, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                  int step, tcl_stream_t stream);
/* Adam over a UVT row shard: fdc, m, v are [rows,3], grad4 the [rows,4] gradient; leaves the gradient zeroed */
int tcl_adam_step_uvt(float* fdc, float* grad4, float* m, float* v, long long rows, float lr, float beta1, float beta2,
                      float eps, int step, tcl_stream_t stream);
/* Shard storage: tcl_peer_alloc = cudaMalloc of its own (outside any caching allocator), zero-filled, plus the
 * 64-byte cudaIpcMemHandle_t of its base in handle_out; the owner frees it with tcl_peer_free after every peer
 * has unmapped it.  tcl_ipc_open maps a peer's handle (same box) into this process: *base receives the mapped
 * address; handles are cached per process (opening one twice returns the same mapping); tcl_ipc_close_all unmaps
 * them all, tcl_ipc_close one. */
int tcl_peer_alloc(size_t bytes, void** ptr, void* handle_out);
int tcl_peer_free(void* ptr);
int tcl_ipc_open(const void* handle, void** base);
int tcl_ipc_close(void* base);
int tcl_ipc_close_all(void);
/* Cross-rank barrier on the stream, over peer-mapped flags: flags[r] = rank r's int32[TCL_MAX_RANKS] slot array
 * (zero-filled once).  Rank `rank` adds 1 to slot `rank` of every peer (release, system scope) and waits until all
 * its own slots reach `epoch` (acquire); `epoch` is 1 for the first barrier and grows by 1 per call.  Work enqueued
 * after the call sees every peer's writes enqueued before their matching call.  Gives up (and flags the error in
 * tcl_peer_barrier_timeouts) after ~2 s instead of hanging the device. */
int tcl_peer_barrier(int32_t* const* flags, int world, int rank, int epoch, tcl_stream_t stream);
long long tcl_peer_barrier_timeouts(void);
/* generate.py:477-479: fdc = RGB2SH(scatter_mean(edited, unq_inv)); cnt_ws = U floats of scratch */
int tcl_uvt_init(const float* edited, const int* ids, int N, int H, int W, long long U, float* fdc, float* cnt_ws,
                 tcl_stream_t stream);
/* generate.py:529-531: out[N,3,H,W] = clamp(SH2RGB(fdc)[unq_inv], 0, 1) */
int tcl_uvt_render(const float* fdc, const int* ids, int N, int H, int W, float* out, tcl_stream_t stream);
/* OptDataset.exposure_align (utils/dataloader.py:38-42): edited <- clamp(edited x E[:3,:3] + E[:,3]) in place */
int tcl_exposure_bake(float* edited, const float* exposure, int N, int H, int W, tcl_stream_t stream);

/* ---- stage-2 producers (HBM-bound, fp32 / int32) -------------------------
 * warp_bicubic   : warp_flow (utils/flow_utils.py:5-16): out[n,c,y,x] = bicubic sample of frames[n,c] at
 *                  (x + flows[n,0,y,x], y + flows[n,1,y,x]), zeros padding, align_corners=True.
 * max_f32        : out[0] = max(x) on the device (feeds the thresholds below without a host sync;
 *                  flow_utils.py:52, 71 call .max().item()).
 * soft_mask_bwd  : get_soft_mask_bwds (flow_utils.py:40-54): out [N,1,H,W]; frame 0 = 1, frame i =
 *                  sigmoid(-beta*(|past_i + W(flow_{i-1})| - (|past_i| + |W(flow_{i-1})| + 1)*alpha))
 *                  * sigmoid(-beta*(max_c|W(img_{i-1}) - img_i| - images_max*diff_threshold)), W = warp by past_i.
 * flow_ids       : get_flowid (flow_utils.py:56-92): ids [N,H,W] int32.  A pixel of frame i inherits the id of
 *                  the frame i-1 pixel whose rounded forward flow lands on it (valid if in bounds,
 *                  mask_bwds[i] > 0.5 at the source, max_c|frames_i(target) - frames_{i-1}(source)| <
 *                  frames_max*rgb_threshold; several candidates: the largest source index wins = the
 *                  reference's CPU result); all other pixels get fresh consecutive ids in frame-major /
 *                  row-major order.  num_ids (device int64, may be NULL) receives the id count U.
 * unique_inverse : voxelization(voxel_size=None) (utils/general_utils.py:223-233) =
 *                  torch.unique(ids[:,None], dim=0, return_inverse=True)[1] for ids in [0, id_range):
 *                  inverse[i] = rank of ids[i] among the distinct ids (int64 like torch), num_unique = U.
 */
int tcl_warp_bicubic(const float* frames, const float* flows, int N, int C, int H, int W, float* out,
                     tcl_stream_t stream);
int tcl_max_f32(const float* x, long long n, float* out, tcl_stream_t stream);
int tcl_soft_mask_bwd(const float* images, const float* flows, const float* past_flows, int N, int H, int W, float alpha,
                      float beta, double diff_threshold, const float* images_max, float* out, tcl_stream_t stream);
size_t tcl_flow_ids_workspace_bytes(int N, int H, int W);
int tcl_flow_ids(const float* frames, const float* flows, const float* mask_bwds, int N, int H, int W,
                 double rgb_threshold, const float* frames_max, int* ids, long long* num_ids, void* workspace,
                 size_t workspace_bytes, tcl_stream_t stream);
size_t tcl_unique_inverse_workspace_bytes(long long id_range);
int tcl_unique_inverse(const int* ids, long long n, long long id_range, long long* inverse, long long* num_unique,
                       void* workspace, size_t workspace_bytes, tcl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TCLIGHT_H_ */
