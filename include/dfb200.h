/* dfb200.h -- C ABI of libdformer_b200.so: hand-written sm_100a kernels for the DFormer RGB-D
 * forward/backward hot path (encoder Block/attention, LightHamHead/NMF2D, upsample + CE loss).
 *
 * The reference (Originofamonia/DFormer) is pure PyTorch and has NO native interface; every entry
 * point below replaces a group of ATen/cuDNN/cuBLAS calls issued by the reference file:line cited at
 * the declaration.  Conventions:
 *   - plain pointers + sizes, no torch types; the CALLER owns all memory (outputs, workspaces);
 *   - every launcher is asynchronous on the given `stream` (a cudaStream_t passed as void*),
 *     never synchronises, never allocates device memory;
 *   - returns 0 on success, a negative code otherwise; dfb200_last_error() gives the message;
 *   - `dtype` codes: 0 = float32, 1 = bfloat16.  Activations are channels-last: [B, H, W, C]
 *     (= row-major [M, C] with M = B*H*W).  Statistics, residual stream and parameters are float32.
 */
#ifndef DFB200_H_
#define DFB200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB200_F32 0
#define DFB200_BF16 1
#define DFB200_ACT_NONE 0
#define DFB200_ACT_GELU 1
#define DFB200_ACT_RELU 2
#define DFB200_BACKEND_SIMT 0    /* CUDA-core fp32-accumulate GEMM (exact fp32 path)         */
#define DFB200_BACKEND_TCGEN05 1 /* tcgen05/TMEM/TMA GEMM (bf16 operands, fp32 accumulation) */
#define DFB200_BACKEND_AUTO 2

const char* dfb200_last_error(void);
int dfb200_version(void);

/* ---- GEMM: nn.Linear / 1x1 conv / im2col conv forward, dgrad and wgrad --------------------------
 * Replaces addmm/convolution/bmm of DFormer.py:60,65,110-121,125,133,141-143, ham_head.py:48,89,122-141,
 * 174-177,234,238, decode_head.py:230 and their autograd backward.
 *   C[M,N] = act( opA(A) * opB(B) + bias ),   opA(A) is M x K, opB(B) is K x N
 *   transA = 0: A stored [M,K] (lda)    transA = 1: A stored [K,M]
 *   transB = 0: B stored [K,N] (ldb)    transB = 1: B stored [N,K]   (nn.Linear weight layout)
 * `act` is applied to columns >= act_col_start only (fused q|q_cut|l projection, DFormer.py:110-113).
 * `accumulate` adds into an fp32 C.  `batch` > 1 runs strided-batched problems (SIMT backend only).
 * The SIMT backend accepts any mix of fp32/bf16 A and B; tcgen05 needs both bf16. */
typedef struct dfb200_gemm_args {
  const void* A; const void* B; void* C;
  const float* bias;
  long lda, ldb, ldc;
  long strideA, strideB, strideC;
  int M, N, K, batch;
  int batch_inner;       /* >1: two-level batch, z = zo*batch_inner + zi, offsets zo*stride + zi*stride_in */
  long strideA_in, strideB_in, strideC_in;
  int transA, transB;
  int a_dtype, b_dtype, out_dtype;
  int act, act_col_start;
  int accumulate;
  int backend;
  int splitk;            /* 0 = auto, 1 = off, >1 = forced number of reduction splits */
  float alpha;           /* scales the product before bias/activation (0 is treated as 1) */
  /* Fused epilogue of the tcgen05 backend (bf16 C through the TMA-store path, batch 1; any other configuration is rejected):
   *   epi_mode 1 "gate": C = (A.B + bias) * aux   (aux: bf16 [M, N], leading dimension ld_aux; DFormer.py:134-135 `q * a`,
   *                      `cutted_x * x_e` computed where the second factor is produced and written straight into a column slice of
   *                      the concat buffer, :137-140); out2 (optional, bf16 [M, N], ld_out2) receives A.B + bias, which the backward
   *                      pass needs.
   * `ls`, `scale_b`, `rows_per_sample` are reserved (NULL / 0).  Measured and removed: a layer-scale-residual epilogue and a
   * gelu'-multiply epilogue (round 1), a "dual activation" epilogue emitting GELU(l) beside l (round 2, +0.3 ms/step: eight
   * epilogue warps evaluate erf slower than a full-occupancy streaming kernel). */
  int epi_mode;
  const void* aux; long ld_aux;
  void* out2; long ld_out2;
  const float* ls; const float* scale_b; int rows_per_sample;
} dfb200_gemm_args;
int dfb200_gemm(const dfb200_gemm_args* args, void* stream);
/* SMs the persistent tcgen05 GEMM grids may cover: `main_sms` for forward / dgrad problems, `split_sms` for split-K (weight-gradient)
 * problems; 0 = the defaults (5/6 and 49 % of the chip, measured best inside the four-stream training step); the SM count = the whole
 * chip (best for a GEMM running alone).  Process-wide, takes effect for the following launches. */
int dfb200_gemm_sm_budget(int main_sms, int split_sms);

/* column sums: out[n] (+)= sum_m X[m,n]  (bias gradients).  X is [M,N] with leading dimension ldx. */
int dfb200_colsum(const void* X, int dtype, long ldx, int M, int N, float* out, int accumulate, void* stream);

/* ---- parameter packing: fp32 parameters -> compute-dtype GEMM operands ---------------------------
 * One launch converts a table of parameter tensors.  kind 0: row-major [rows, cols] copy into
 * dst (leading dimension dst_ld, zero-padding cols..dst_ld-1 is the caller's job via memset);
 * kind 1: conv weight [Cout, Cin, 3, 3] -> [Cout, 9*Cin (+pad)] with k = (ky*3+kx)*Cin + ci.
 * The table lives in DEVICE memory (uploaded once by the caller). */
typedef struct dfb200_pack_entry {
  const float* src; void* dst;
  int rows, cols, dst_ld, kind;   /* kind 1: cols = Cin */
} dfb200_pack_entry;
int dfb200_pack_params(const dfb200_pack_entry* table_dev, int n_entries, int max_elems, int dst_dtype, void* stream);
/* inverse of kind 1 for gradients: dW[Cout,Cin,3,3] = gather(dWp[Cout, ld]) */
int dfb200_unpack_conv_grad(const float* dWp, int ld, int Cout, int Cin, float* dW, void* stream);

/* ---- LayerNorm over channels (DFormer.py:37-39, eps 1e-6) ---------------------------------------- */
int dfb200_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int M, int C,
                         void* y, int y_dtype, float* mean, float* rstd, void* stream);
/* dx is fp32 [M,C]; dx = (dx_in ? dx_in : 0) + LN-gradient (dx_in = gradient of the residual branch; may alias dx).
 * dgamma/dbeta are accumulated with atomics and must be zero-initialised by the caller. */
/* Fused layer-scale residual + LayerNorm (DFormer.py:173-179 then :59): x_out = res + scale_b[sample] * ls * branch (fp32,
 * written once), y = LN(x_out) in `dtype` (= dtype of branch), mean / rstd kept for backward. */
int dfb200_scale_residual_layernorm_fwd(const float* res, const void* branch, long ld_branch, int dtype, const float* ls,
                                        const float* scale_b, int rows_per_sample, int M, int C, float* x_out,
                                        const float* gamma, const float* beta, float eps, void* y, float* mean, float* rstd,
                                        void* stream);
int dfb200_layernorm_bwd(const void* dy, const void* dy2 /* optional second gradient, added to dy (fan-in) */, int dy_dtype, const float* x, const float* gamma, const float* mean,
                         const float* rstd, int M, int C, const float* dx_in, float* dx, float* dgamma, float* dbeta,
                         void* stream);

/* ---- depthwise k x k conv, stride 1, 'same' padding, channels-last (DFormer.py:54,62,80-81,115,133)
 * y = act( dw(x) + bias [+ x if add_input] ).  weight is the nn.Conv2d tensor [C,1,k,k] (fp32), k in {3,7}. */
int dfb200_dwconv_fwd(const void* x, int dtype, const float* weight, const float* bias, int B, int H, int W, int C,
                      int k, int add_input, int act, void* y, void* z_out /* optional: pre-activation, kept for backward */,
                      void* stream);
/* Backward.  Given dy and the forward input x:
 *   dz = dy * act'(z) with z taken from `z` when given, else recomputed from x (act != 0);  dx = dw^T(dz) [+ dz];
 *   dweight[C,1,k,k], dbias[C] accumulated with atomics (zero-initialised by the caller).
 * `dz_buf` is a caller-provided scratch of the same shape/dtype as dy (only used when act != 0). */
int dfb200_dwconv_bwd(const void* dy, const void* x, const void* z /* optional saved pre-activation */, int dtype,
                      const float* weight, const float* bias, int B, int H, int W, int C, int k, int add_input, int act,
                      void* dz_buf, void* dx, float* dweight, float* dbias, void* stream);

/* ---- fused middle of the MLP (DFormer.py:62-64), bf16 channels-last activations only (dtype must be 1) ------
 * forward:  u = GELU(dw3x3(h) + bias + h); the pre-activation is not stored.
 * backward: dz = du * GELU'(dw3x3(h) + bias + h) (recomputed, kept on chip), dh = dz + dw3x3^T(dz),
 *           dweight[C,1,3,3] += dz (*) h, dbias[C] += colsum(dz), dh_colsum[C] += colsum(dh) (= fc1 bias gradient;
 *           optional).  The three accumulators are zero-initialised by the caller (fp32, atomics).
 * TMA-fed persistent kernels: one pass over h (+ du) and one over the output. */
int dfb200_mlp_dw_fwd(const void* h, int dtype, const float* weight, const float* bias, int B, int H, int W, int C,
                      void* u, void* gp /* optional: GELU'(pre-activation) in bf16, kept for the backward pass */, void* stream);
/* gp given: dz = du * gp (pure stream, 4 passes); gp NULL: GELU'(dw3x3(h) + bias + h) is recomputed (3 passes, more math) */
int dfb200_mlp_dw_bwd(const void* du, const void* gp, const void* h, int dtype, const float* weight, const float* bias,
                      int B, int H, int W, int C, void* dh, float* dweight, float* dbias, float* dh_colsum, void* stream);

/* ---- elementwise glue of Block/Attention ---------------------------------------------------------
 * mul:  out[m, n] = a[m, n] * b[m, n]  with independent leading dimensions (q*a, cut*e: DFormer.py:134-135;
 *       `out` is a column slice of the concat buffer of :137-140). */
int dfb200_mul_fwd(const void* a, long lda, const void* b, long ldb, void* out, long ldo, int dtype, int M, int N, void* stream);
int dfb200_mul_bwd(const void* dout, long ldo, const void* a, long lda, const void* b, long ldb, void* da, long ldda,
                   void* db, long lddb, int dtype, int M, int N,
                   float* da_colsum, float* db_colsum /* optional pair: += column sums of da / db (bias gradients) */, void* stream);
/* layer-scale residual (DFormer.py:173-179): out[m,c] = res[m,c] + scale_b[b] * ls[c] * y[m,c]
 * (scale_b = DropPath mask / keep_prob per sample, NULL = 1).  res/out fp32, y in `dtype`. rows_per_sample = H*W. */
int dfb200_scale_residual_fwd(const float* res, const void* y, long ldy, int dtype, const float* ls, const float* scale_b,
                              int M, int C, int rows_per_sample, float* out, void* stream);
/* dy = dout * ls * scale_b;  dls[c] += sum_m dout*y*scale_b (atomics; zero-init by caller). d(res) = dout (alias). */
int dfb200_scale_residual_bwd(const float* dout, const void* y, long ldy, int dtype, const float* ls, const float* scale_b,
                              int M, int C, int rows_per_sample, void* dy, long lddy, float* dls,
                              float* dy_colsum /* optional: += column sums of dy (bias gradient of the layer that produced y) */,
                              void* stream);
/* stand-alone activation on a column slice: out = act(in);  din = dout * act'(z) (GELU: z = pre-activation;
 * ReLU: z may be the forward output).  All operands have independent leading dimensions. */
int dfb200_act_fwd(const void* in, long ldi, void* out, long ldo, int dtype, int act, int M, int N, void* stream);
int dfb200_act_bwd(const void* dout, long lddo, const void* dout2 /* optional second gradient added to dout */, long lddo2,
                   const void* z, long ldz, void* din, long lddi, int dtype, int act, int M, int N,
                   float* din_colsum /* optional: += column sums of din */, void* stream);

/* ---- Global Awareness Attention pieces (DFormer.py:107-108,120-131) ------------------------------
 * pool: AdaptiveAvgPool2d(7,7) of cat[xn (C1 ch), en (C2 ch)] -> out [B,49,C1+C2] (compute dtype). */
int dfb200_pool7_fwd(const void* xn, int C1, const void* en, int C2, int dtype, int B, int H, int W, void* out, void* stream);
int dfb200_pool7_bwd(const void* dout, int C1, int C2, int dtype, int B, int H, int W, void* dxn, void* den, void* stream);
/* attention: queries m [B,49,heads*d], keys/values kv [B,HW,2*heads*d] (k = cols [0,heads*d), v = rest,
 * head-major), softmax over the HW pixels, scale d^-0.5.  out [B,49,heads*d].
 * out is fp32 (it is tiny and feeds the 7x7 -> HxW resize). */
int dfb200_gaa_fwd(const void* m, const void* kv, int dtype, int B, int HW, int heads, int d, float* out,
                   float* probs /* [B,heads,49,HW] fp32, kept for backward */, void* stream);
/* dout fp32 [B,49,heads*d]; dm fp32 [B,49,heads*d]; dkv [B,HW,2*heads*d] in `dtype`;
 * scratch: 2 * B*heads*49*HW floats (dP / dS). */
int dfb200_gaa_bwd(const float* dout, const void* m, const void* kv, const float* probs, int dtype, int B, int HW,
                   int heads, int d, float* dm, void* dkv, float* scratch, void* stream);

/* Fused form of the same attention core (one launch per direction, nothing of size 49 x HW in HBM; head dim d in
 * {16, 32, 36, 48}).  forward: out [B*49, heads*d] fp32, lse [B*heads*49] (row log-sum-exp kept for backward);
 * scratch: B*heads*ceil(HW/128)*49*(d+4) floats of partials; counters: B*heads ints, zero before the first call
 * (self-resetting).  backward: recomputes the probabilities from (m, kv, lse); dm is zero-filled by the launcher.
 * dtype bf16 runs on the tensor cores (mma.sync) with thread-block clusters: partials are merged through distributed shared
 * memory, so `scratch` and `counters` are unused there and may be NULL. */
int dfb200_gaa_fused_fwd(const void* m, const void* kv, int dtype, int B, int HW, int heads, int d, float* out, float* lse,
                         float* scratch, int* counters, void* stream);
int dfb200_gaa_fused_bwd(const float* dout, const float* out, const float* lse, const void* m, const void* kv, int dtype,
                         int B, int HW, int heads, int d, float* dm, void* dkv, void* stream);
/* bf16 only: the same backward pass, also emitting (each optional, NULL = skip) dkv_colsum[2*heads*d] += column sums of dkv (bias
 * gradient of `kv`, DFormer.py:121), dm_colsum[heads*d] += column sums of dm (bias gradient of `short_cut_linear`, :108) and
 * dm_lo = dm in bf16 (the operand of that layer's gradient GEMMs): three launches less per Block. */
int dfb200_gaa_fused_bwd_ex(const float* dout, const float* out, const float* lse, const void* m, const void* kv, int B, int HW,
                            int heads, int d, float* dm, void* dkv, float* dkv_colsum, float* dm_colsum, void* dm_lo, void* stream);
/* bilinear (align_corners=False) resize of a channels-last map into a column slice of a wider buffer:
 * out[b, y, x, col0 + c] = interp(in[b, :, :, c]).  Used for 7x7 -> HxW (DFormer.py:131), the head's
 * resize+concat (ham_head.py:226-233) with fp32 inputs. */
int dfb200_resize_fwd(const void* in, int in_dtype, int B, int Hi, int Wi, int C, void* out, int out_dtype, int Ho,
                      int Wo, long ldo, int col0, void* stream);
/* din (+)= resize^T(dout slice); din has the input's dtype; accumulate only for fp32 din. */
int dfb200_resize_bwd(const void* dout, int out_dtype, long ldo, int col0, int B, int Hi, int Wi, int C, int Ho, int Wo,
                      void* din, int in_dtype, int accumulate, void* stream);

/* ---- dense 3x3 stride-2 convs of the stems / downsample layers (DFormer.py:194-228,295-296) --------
 * im2col gather into [B*Ho*Wo, ld] (k = (ky*3+kx)*Cin + ci, zero padded to ld) followed by dfb200_gemm.
 * Input addressing is generic: element (b, y, x, c) at in[b*sb + y*sy + x*sx + c*sc] so that the NCHW
 * network inputs (and the channel-0 slice of modal_x, DFormer.py:286) are read in place. */
int dfb200_im2col3x3s2_fwd(const void* in, int in_dtype, long sb, long sy, long sx, long sc, int B, int H, int W,
                           int Cin, void* out, int out_dtype, int ld, void* stream);
int dfb200_im2col3x3s2_bwd(const void* dcol, int col_dtype, int ld, int B, int H, int W, int Cin, void* din,
                           int in_dtype, void* stream);

/* ---- BatchNorm2d on channels-last activations (DFormer.py:196,199,207,210,219,225; ham_head.py:170,209,219)
 * stats: sum[c], sumsq[c] over M rows (fp32, written not accumulated).  Cross-rank SyncBN all-reduces them. */
int dfb200_bn_stats(const void* x, int dtype, int M, int C, double* sum, double* sumsq, void* stream);
/* finalize: mean/invstd from (sum, sumsq, count); updates running stats (momentum, unbiased var) if non-NULL. */
int dfb200_bn_finalize(const double* sum, const double* sumsq, double count, float eps, float momentum, int C, float* mean,
                       float* invstd, float* running_mean, float* running_var, void* stream);
/* eval: mean/invstd from running statistics */
int dfb200_bn_eval_stats(const float* running_mean, const float* running_var, float eps, int C, float* mean, float* invstd, void* stream);
/* y = act( (x-mean)*invstd*gamma + beta [+ residual] ) [* chan_scale[b,c]] ; act: 0 none, 1 GELU, 2 ReLU */
int dfb200_bn_apply(const void* x, int x_dtype, const float* mean, const float* invstd, const float* gamma,
                    const float* beta, const void* residual, int act, const float* chan_scale, int rows_per_sample,
                    int M, int C, void* y, int y_dtype, void* stream);
/* backward, pass 1: g = (dy [+ dy2]) [* chan_scale] * act'(.) ; writes g (dtype of y) into gbuf, accumulates
 * sum_g[c], sum_gx[c] = sum g*xhat (zero-init by caller).  dy2 (optional, same dtype / shape as dy) is a second incoming
 * gradient (fan-in of a residual branch) summed in the same pass.  pass 2 (bn_bwd_apply):
 *   dx = gamma*invstd*( g - sum_g/n - xhat*sum_gx/n )   (training)  or gamma*invstd*g (eval)          */
int dfb200_bn_bwd_reduce(const void* dy, const void* dy2, int y_dtype, const void* x, int x_dtype, const float* mean, const float* invstd,
                         const float* gamma, const float* beta, const void* residual, int act, const float* chan_scale,
                         int rows_per_sample, int M, int C, void* gbuf, float* sum_g, float* sum_gx,
                         float* dbeta /* optional: += sum_g (the bias gradient) */, float* dgamma /* optional: += sum_gx */, void* stream);
int dfb200_bn_bwd_apply(const void* gbuf, int y_dtype, const void* x, int x_dtype, const float* mean, const float* invstd,
                        const float* gamma, const float* sum_g, const float* sum_gx, float count, int training, int M,
                        int C, void* dx, int dx_dtype, void* stream);

/* ---- NMF2D element-wise steps (ham_head.py:49,115,126,133,143); the bmm's go through dfb200_gemm ---- */
int dfb200_normalize_cols(const float* in, int B, int D, int R, float* out, float* norms, void* stream); /* F.normalize(dim=1) */
int dfb200_softmax_rows(const float* in, int rows, int cols, float* out, void* stream);
int dfb200_softmax_rows_bwd(const float* dout, const float* out, int rows, int cols, float* din, void* stream);
/* out = a * num / (den + eps) */
int dfb200_mu_update(const float* a, const float* num, const float* den, float eps, long n, float* out,
                     void* out_lo /* optional copy of out in lo_dtype (operand of the next GEMM) */, int lo_dtype, void* stream);
/* da (fp32, optionally accumulated); dnum [n/cols rows x cols, leading dimension ld_dnum] and dden [n] in lo_dtype */
int dfb200_mu_update_bwd(const float* dout, const float* a, const float* num, const float* den, float eps, long n,
                         float* da, int accumulate_da, void* dnum, long ld_dnum, int cols, void* dden, int lo_dtype,
                         void* stream);
int dfb200_cast(const void* in, int in_dtype, void* out, int out_dtype, long n, void* stream);
/* strided 2-D convert/copy: out[r*ld_out + c] = in[r*ld_in + c], r < rows, c < cols (column slices of concatenated operands) */
/* out[b,i,j] = in[b,i,j] + in[b,j,i] for `batch` contiguous R x R matrices (ham_head.py:122-141 backward: d(B^T B) enters both
 * factors, so one product with the symmetrised gradient replaces two). in fp32/bf16, out fp32/bf16 (not bf16 -> fp32). */
int dfb200_sym_cast(const void* in, int in_dtype, void* out, int out_dtype, int batch, int R, void* stream);
int dfb200_cast2d(const void* in, int in_dtype, long ld_in, void* out, int out_dtype, long ld_out, long rows, int cols,
                  void* stream);
int dfb200_axpy(const void* x, int x_dtype, float alpha, void* y, int y_dtype, long n, void* stream); /* y += alpha*x */

/* ---- fused x8 bilinear upsample + cross entropy (builder.py:203,230) ------------------------------
 * logits_small: channels-last [B, h, w, ncls] (compute dtype).  Writes (optionally) the NCHW fp32
 * up-sampled logits `out` [B, ncls, H, W], and accumulates loss_sum / valid_count (2 floats, zero-init).
 * label: int64 [B,H,W], `ignore` = 255.  loss = loss_sum / valid_count is finished by the caller or by
 * dfb200_ce_finalize (device scalar, no host sync). */
int dfb200_upsample_ce_fwd(const void* logits_small, int dtype, int B, int h, int w, int ncls, int H, int W,
                           const int64_t* label, int ignore, float* out_nchw, float* lse /* [B,H,W] */, float* loss_acc,
                           void* up_lowp /* optional [B,ncls,H,W] copy of the up-sampled logits in `dtype` (kept for backward) */,
                           void* stream);
int dfb200_ce_finalize(const float* loss_acc, float* loss, void* stream);
/* Training step in one pass (builder.py:203,230 forward AND backward): loss_acc[0..1] += (sum of NLL, number of valid pixels), lse (optional)
 * = per-pixel log-sum-exp, dgrad (fp32 [B*h*w, ncls], zero-initialised by the caller) += d(sum of NLL)/d(logits_small); nothing hi-res
 * is materialised.  dfb200_ce_grad_finalize turns it into the gradient of the mean: out = dgrad * dloss / loss_acc[1] (dtype of choice).
 * Returns DFB_ERR_UNSUPPORTED for down-sampling geometries or more than 12 hi-res rows per source row (use fwd + bwd_fused there). */
int dfb200_upsample_ce_train(const void* logits_small, int dtype, int B, int h, int w, int ncls, int H, int W, const int64_t* label,
                             int ignore, float* lse, float* loss_acc, float* dgrad, void* stream);
int dfb200_ce_grad_finalize(const float* dgrad, long n, const float* loss_acc, const float* dloss, void* out, int out_dtype, void* stream);
/* dlogits_small [B,h,w,ncls] = resize^T( (softmax(up) - onehot) * dloss / valid )  (gather form, deterministic) */
/* Separable (rows, then columns) form of the same adjoint, fed by the up-sampled logits `up` [B,ncls,H,W] (dtype up_dtype)
 * that the forward kernel can emit: each hi-res probability is evaluated once instead of once per overlapping footprint.
 * scratch: B*ncls*h*W floats. */
int dfb200_upsample_ce_bwd_sep(const void* up, int up_dtype, int B, int h, int w, int ncls, int H, int W, const int64_t* label,
                               int ignore, const float* lse, const float* loss_acc, const float* dloss, float* scratch,
                               void* dlogits_small, int dl_dtype, void* stream);
/* Separable adjoint that recomputes the hi-res logits from logits_small (nothing but lse is kept by the forward pass):
 * rows pass (one softmax evaluation per hi-res pixel and class, split between its two source rows) then columns pass.
 * scratch: 2*B*h*W*ncls floats. */
int dfb200_upsample_ce_bwd_fused(const void* logits_small, int dtype, int B, int h, int w, int ncls, int H, int W,
                                 const int64_t* label, int ignore, const float* lse, const float* loss_acc,
                                 const float* dloss, float* scratch, void* dlogits_small, int dl_dtype, void* stream);
int dfb200_upsample_ce_bwd(const void* logits_small, int dtype, int B, int h, int w, int ncls, int H, int W,
                           const int64_t* label, int ignore, const float* lse, const float* loss_acc, const float* dloss,
                           void* dlogits_small, int dl_dtype, void* stream);

/* ---- training input pipeline (utils/dataloader/dataloader.py:20-73 TrainPre; row N2) --------------------------------------
 * rgb / modal [B,H,W,3] uint8, label [B,H,W] uint8 on the device; params [B][5] int32 = {mirror, scaled_h, scaled_w, crop_y,
 * crop_x} (the reference's random draws, made on the host in its order); lut_* [3][256] fp32 = ((v/255 - mean)/std) built in
 * float64.  Outputs: rgb / modal [B,3,crop_h,crop_w] fp32 (padding 0), label [B,crop_h,crop_w] int64 (padding 255).
 * Bit-exact with cv2.flip / cv2.resize(INTER_LINEAR | INTER_NEAREST) / copyMakeBorder on uint8 + numpy normalisation. */
int dfb200_train_pre(const void* rgb, const void* modal, const void* label, int B, int H, int W, const int* params,
                     const float* lut_rgb, const float* lut_modal, int crop_h, int crop_w, float* out_rgb, float* out_modal,
                     int64_t* out_label, void* stream);

/* ---- inference-time BatchNorm folding for conv -> BN pairs (row N4): scales the packed GEMM weight [rows, ld] (dtype) row-wise in
 * place by gamma * rsqrt(running_var + eps) and writes bias_out[rows] = (conv_bias - running_mean) * scale + beta. */
int dfb200_bn_fold(void* w_packed, int dtype, int rows, int cols, long ld, const float* conv_bias, const float* running_mean,
                   const float* running_var, float eps, const float* gamma, const float* beta, float* bias_out, void* stream);

/* ---- multi-scale + flip evaluation and the mIoU confusion matrix (utils/val_mm.py:257-470, utils/metrics_new.py:6-47) ----
 * resize_nchw_ac:   out[B,C,Ho,Wo] = bilinear(in[B,C,Hi,Wi], align_corners=True), optionally mirrored along W afterwards
 *                   (val_mm.py:366-368,378-380: F.interpolate(..., align_corners=True) then torch.flip(dims=(3,))).
 * ms_softmax_accum: acc[B,C,H,W] += softmax_C( bilinear_ac( flip ? mirror_W(logits) : logits ) )   (val_mm.py:385-399).
 * argmax_confusion: pred = argmax_C score; hist[target * C + pred] += 1 where target != ignore (metrics_new.py:18-22);
 *                   hist [C*C] fp32 accumulated, pred [B*HW] optional, target may be NULL (prediction only). */
int dfb200_resize_nchw_ac(const float* in, int B, int C, int Hi, int Wi, float* out, int Ho, int Wo, int flip, void* stream);
int dfb200_ms_softmax_accum(const float* logits, int B, int C, int h, int w, float* acc, int H, int W, int flip, void* stream);
int dfb200_argmax_confusion(const float* score, const int64_t* target, int B, int C, long HW, int ignore, float* hist,
                            int64_t* pred, void* stream);

/* ---- fused multi-tensor AdamW on flat fp32 buffers (utils/train.py:211,336; next-row N1) ------------ */
/* per-element `wd_arr` (weight decay, NULL = weight_decay everywhere) and `lr_arr` (lr multiplier, 0 freezes an
 * element; NULL = 1) reproduce the reference's parameter groups (utils/init_func.py:26-70) on one flat buffer. */
int dfb200_adamw(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, float bias_c1, float bias_c2, float grad_scale, const float* wd_arr,
                 const float* lr_arr, const float* dyn /* optional device {lr, step}: overrides lr and the bias corrections
                 so a captured CUDA graph follows the schedule */, void* stream);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches evidence) */
long dfb200_launch_count(void);

/* ---- SyncBatchNorm statistics exchange over NVLink peer memory (torch SyncBatchNorm of utils/train.py:182-194) ------------
 * Small-message all-reduce among the ranks of ONE box (one process per GPU).  Every rank allocates an exchange buffer
 * (dfb200_peer_alloc: one cudaMalloc of ~0.5 MB, zero-filled), exports it as a 64-byte CUDA IPC handle, opens the
 * handles of its peers (the caller ships the handles between processes, e.g. torch.distributed.all_gather_object) and
 * uploads the table of `world` base pointers -- entry `rank` = its own buffer -- to device memory.
 * dfb200_peer_allreduce: out[i] = sum over ranks of in[i] (n elements, dtype 0 = float32, 2 = float64, at most 16 KB),
 * summed in rank order (bitwise identical on every rank); one single-CTA kernel, asynchronous on `stream`, graph-capturable
 * (the call sequence number lives in the buffer and is advanced on the device).  Every rank must issue the same sequence
 * of calls; a rank that waits ~20 s for a peer traps (launch failure) instead of hanging. */
int dfb200_peer_alloc(void** ptr);
int dfb200_peer_free(void* ptr);
int dfb200_peer_export(void* ptr, unsigned char* handle64);
int dfb200_peer_open(const unsigned char* handle64, void** ptr);
int dfb200_peer_close(void* ptr);
int dfb200_peer_allreduce(const void* in, void* out, int dtype, int n, void* const* bases_dev, int rank, int world, void* stream);

/* ---- NCCL thin wrappers: the gradient all-reduce of DistributedDataParallel (utils/train.py:238-243) ------
 * For hosts that do not bring torch.distributed (the Python mirror uses torch.distributed's NCCL backend, dformer_b200/parallel.py).
 * libnccl is resolved at run time: $DFB200_NCCL_LIB, else "libnccl.so.2" (inside a PyTorch process: the already loaded
 * torch-bundled library), else "libnccl.so"; every call returns -3 with a message when it cannot be found.
 *   dfb200_nccl_unique_id: rank 0 fills 128 bytes (ncclUniqueId) and ships them to the other ranks out of band;
 *   dfb200_nccl_comm_init: collective over the `nranks` processes (one per GPU; the calling thread's current device is the rank's GPU);
 *   dfb200_nccl_all_reduce: in place on `buf` (count elements, dtype 0 = float32, 1 = bfloat16, 2 = float64), sum or -- `average`
 *                           != 0 -- mean over the ranks (ncclAvg, what DDP applies to gradients), asynchronous on `stream`:
 *                           call it per finished bucket of the flat gradient arena on a communication stream. */
int dfb200_nccl_version(int* version);
int dfb200_nccl_unique_id(void* id128);
int dfb200_nccl_comm_init(const void* id128, int nranks, int rank, void** comm);
int dfb200_nccl_all_reduce(void* comm, void* buf, long count, int dtype, int average, void* stream);
int dfb200_nccl_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* DFB200_H_ */
