/*
 * s2f.h -- C ABI of libs2f.so, the sm_100a kernels of the Spike2Former spiking hot path.
 *
 * Every entry point takes raw device pointers, plain sizes and a cudaStream_t (as void*),
 * returns 0 on success and a non-zero code on failure (message via s2f_last_error()).
 * Nothing allocates, nothing synchronises with the host: every call is CUDA-graph capturable.
 * There is no CPU implementation behind any of these symbols.
 *
 * Layouts: activations are channels-last.  "spikes" are int8 levels 0..D (the reference's
 * float spike value is level/norm); real tensors are fp32.  n = T*B images.
 *
 * Each function cites the reference code it replaces (paths relative to
 * /root/reference/Segmentation).
 */
#ifndef S2F_H_
#define S2F_H_

#include <stdint.h>

#if defined(__GNUC__)
#define S2F_API __attribute__((visibility("default")))
#else
#define S2F_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define S2F_OK 0
#define S2F_ERR_ARG 1
#define S2F_ERR_CUDA 2
#define S2F_ERR_UNSUPPORTED 3

/* Thread-local description of the last failure. */
S2F_API const char* s2f_last_error(void);
/* ABI version of the library (bumped when a signature changes). */
S2F_API int s2f_abi_version(void);   /* currently 10 */
/* Number of kernel launches issued through this library since load (for bench.py's gpu_launches). */
S2F_API uint64_t s2f_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * (a) NI-LIF neuron.  Replaces Q_IFNode.forward / BaseNode.forward + quant.forward:
 *     Qtrick_architecture/clock_driven/neuron.py:459-460,115-131,133-153,166-197 and
 *     surrogate.py:522-529.
 *   for t in [0,T): u = x[t,i]*scale[c] + shift[c] + residual[t,i]   (each term optional, c = i % C)
 *                   v += u ; s = rint(clamp(v, 0, d_max)) ; v -= s
 *   levels[t,i] = s (int8)   ;  y_norm[t,i] = s / norm (fp32, optional)
 *   v starts at v_in[i] (or 0) and is kept in a register across the T loop; v_out optional.
 *   residual_period: the residual index is (t*N + i) % residual_period (0 = no wrap), which lets a
 *   [N_tokens, C] positional table be broadcast over the batch.
 *   transpose_rows/cols: when both > 0 the output of element f (within an image of rows*cols
 *   elements) is written at (f % rows) * cols + f / rows, i.e. the [cols, rows] matrix read in the
 *   reference's reinterpreting reshape (mmcv_spike/transformer.py:777) is stored transposed.
 *   ties (optional, device uint64): incremented by the number of exact .5 ties inside (0, d_max).
 */
S2F_API int s2f_nilif_fwd(const float* x, const float* scale, const float* shift, const float* residual,
                  int64_t residual_period, const float* v_in, float* v_out, int8_t* levels, float* y_norm,
                  int T, int64_t N, int C, float d_max, float norm, int transpose_rows, int transpose_cols,
                  unsigned long long* ties, void* stream);

/* Two stateless neurons on one read of x (T = 1):
 *   levels_with_res    = NI-LIF(x*scale + shift + residual[i % residual_period])
 *   levels_without_res = NI-LIF(x*scale + shift)
 * -- the key / value inputs of one pyramid level of the transformer decoder, LIF(y + level_embed + pos) and
 * LIF(y + level_embed) (dense_heads/maskformer_head.py:535-549 feeding mmcv_spike/transformer.py:318-361).
 * N % 16 == 0, C % 4 == 0, residual_period % 4 == 0 (0: residual has N elements), 16-byte aligned pointers. */
S2F_API int s2f_nilif_pair(const float* x, const float* scale, const float* shift, const float* residual,
                   int64_t residual_period, int8_t* levels_with_res, int8_t* levels_without_res, int64_t N, int C,
                   float d_max, void* stream);

/* Surrogate-gradient backward of the neuron for T=1, v0=0 (quant.backward, surrogate.py:531-538,
 * then the /norm of neuron.py:197):  gx = gy / norm * 1[0 <= u <= d_max],  u as above. */
S2F_API int s2f_nilif_bwd(const float* x, const float* scale, const float* shift, const float* residual,
                  const float* gy, float* gx, int64_t N, int C, float d_max, float norm, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (b) Convolution / linear layers with a spike (or real) activation operand and the folded
 *     BatchNorm affine + optional residual + optional NI-LIF in the epilogue.
 *     Replaces nn.Conv2d/Conv1d/Linear + BatchNorm (+ Q_IFNode) chains, e.g. sdtv2.py:207-219
 *     (MS_ConvBlock), :242-255 (MS_MLP), :412-421 (MS_DownSampling), SNN_core.py:47-63, pixel_decoder.py:451-470.
 *
 *   A: activations [n, H, W, Cin] channels-last; a_is_spike: int8 levels (value = level * a_scale)
 *      else fp32.  W: fp32 [Cout, KH, KW, Cin] with each row zero-padded to a multiple of 4 floats
 *      (row stride = (KH*KW*Cin + 3)/4*4), 16-byte aligned.  Output pixel grid Ho x Wo from stride/pad.
 *   acc = sum A*W ;  y = acc * scale[co] + shift[co] + residual[n,ho,wo,co]   (residual optional)
 *   out_f32 (optional): y.   out_spike (optional): rint(clamp(y,0,d_max)) as int8.
 *   out_transposed: outputs are written channel-major per image ([n, Cout, Ho*Wo]) -- the layout the
 *   reference then reinterprets (dcnv3.py:214-215, mmcv_spike/transformer.py:781,829).
 *   Generic A strides: a_stride_m / a_stride_k (in elements) are used when KH=KW=1 and they are
 *   non-zero: A[m,k] = base + img*a_img_stride + m*a_stride_m + k*a_stride_k (covers transposed operands).
 *   The fp32 7x7 stem (Cin = 3) honours them too, with m = input pixel index and k = channel: a_stride_m = 1,
 *   a_stride_k = H*W reads the planar NCHW image the reference's backbone receives without an NHWC copy.
 */
typedef struct {
  const void* a; int a_is_spike; float a_scale;
  const float* w; const float* scale; const float* shift; const float* residual;
  float* out_f32; int8_t* out_spike; int out_transposed;
  int n, H, W, Cin, Cout, KH, KW, stride, pad;
  int64_t a_img_stride, a_stride_m, a_stride_k;
  int64_t w_img_stride;          /* != 0: a different weight matrix per image (batched matmul) */
  float d_max;
} s2f_conv_args;

S2F_API int s2f_conv_simt(const s2f_conv_args* args, void* stream);

/* tcgen05 / TMEM / TMA spike GEMM (sm_100a).  A: int8 levels [n,H,W,Cin] channels-last (Cin >= 32, Cin % 16 == 0).
 * w_packed: the layer's fp32 weights split by s2f_pack_weights_i8 into `pieces` signed base-128 int8 digit planes;
 * one tcgen05.mma.kind::i8 (M=128, N=64*pieces, K=32) accumulates all planes exactly in int32 (TMEM) and the epilogue
 * recombines them: acc = sum_p plane_p * 128^(pieces-1-p).  Then, as in s2f_conv_simt,
 *   y = acc * scale[co] + shift[co] + residual ; out_f32 = y ; out_spike = rint(clamp(y, 0, d_max)).
 * `scale[co]` must already contain the packer's w_rowscale[co] and the spike normaliser (1/8), both powers of two,
 * besides the folded BatchNorm scale.  KH=KW in {1,3}; stride in {1,2}; out_transposed as in s2f_conv_simt. */
typedef struct {
  const int8_t* a; const int8_t* w_packed;
  const float* scale; const float* shift; const float* residual;
  float* out_f32; int8_t* out_spike; int out_transposed;
  int n, H, W, Cin, Cout, KH, KW, stride, pad, pieces;
  float d_max;
  int per_image_weights;   /* 1: w_packed / scale / shift hold one matrix per image (1x1, Ho*Wo % 128 == 0);
                            * 2: w_packed per image, scale / shift shared by all images */
  /* Top-down FPN merge fused into the epilogue (pixel_decoder.py:455-462): when up_prev != NULL,
   *   y += bilinear_upsample(up_prev [n, up_H, up_W, Cout] fp32 -> [Ho, Wo], align_corners=False)[row, co]
   * is added after the affine (same arithmetic as s2f_upsample_add_lif), so lateral conv + upsample + add + NI-LIF
   * is one launch and the fp32 lateral map never reaches HBM.  Not combinable with out_transposed. */
  const float* up_prev; int up_H, up_W;
  /* 1x1 layers: bytes between consecutive rows of A (0 = Cin), so that A may be a column slice of a wider buffer
   * (the q block of a fused [n, N, 3C] q|k|v projection).  Multiple of 16. */
  int64_t a_ld;
} s2f_gemm_tc_args;

S2F_API int s2f_gemm_i8_tc(const s2f_gemm_tc_args* args, void* stream);

/* Host-side helper (no GPU work): split fp32 weights [Cout, taps*Cin] (Cin fastest) into `pieces` (1..3) digit planes
 * in the tile layout the kernel's TMA expects (rows = ceil(Cout/64) * pieces * 64, row length taps * cin_pad bytes)
 * and write w_rowscale[Cout] (powers of two).  Returns the packed size in bytes (also when w_packed == NULL), -1 on
 * bad arguments. */
S2F_API int64_t s2f_pack_weights_i8(const float* w, int Cout, int taps, int Cin, int pieces, int8_t* w_packed,
                            float* w_rowscale);

/* Device-side packer for weights computed on the GPU (e.g. the per-image mask-embedding x mask_feature product):
 * w fp32 [n_img*rows_per_img, ld] (first K columns are weights), one matrix of rows_per_img rows per image ->
 * digit planes in the kernel's per-image tile layout (buffer must be zero-initialised: padding rows are not written),
 * scale_out[row] = rowscale * post_scale, shift_out[row] = bias_col[row*ld] (or 0). */
S2F_API int s2f_pack_rows_i8_device(const float* w, int ld, int n_img, int rows_per_img, int K, int pieces,
                            int8_t* packed, float* scale_out, float* shift_out, const float* bias_col,
                            float post_scale, void* stream);

/* Depthwise k x k convolution (k in {3,5,7}), pad (k-1)/2, stride 1, channels-last, with the same
 * affine / NI-LIF epilogue.  Replaces the depthwise nn.Conv2d of sdtv2.py:154-162, SNN_core.py:36-40,
 * dcnv3.py:150-158, pixel_decoder.py:373-378.  w: fp32 tap-major [k*k, C] (the reference's
 * [C,1,k,k] transposed by the host).  no_pad: 'valid' convolution (output shrinks by k-1).
 * pad_value is reserved and must be NULL: the BNAndPadLayer border of sdtv2.py:48-89 never reaches
 * the device because RepConv is re-parameterised into one dense 3x3 convolution on the host.  */
S2F_API int s2f_dwconv(const void* a, int a_is_spike, float a_scale, const float* w, const float* scale, const float* shift,
               const float* pad_value, float* out_f32, int8_t* out_spike, int n, int H, int W, int C, int k,
               int no_pad, float d_max, void* stream);

/* SepConv tail in one launch (sdtv2.py:176-179: `x = self.dwconv(x); x = self.bn2(self.pwconv2(x))`; the stencil output
 * is real-valued, so the 1x1 is a real x real product): depthwise k x k (pad (k-1)/2) over int8 levels a [n,H,W,Cm]
 * (value = level * a_scale), tap-major fp32 weights w_dw [k*k, Cm]; then
 *   y = (W_pw x2) * scale[co] + shift[co] + residual ; out_f32 = y ; out_spike = rint(clamp(y, 0, d_max)).
 * The stencil runs on CUDA cores in the reference tap order (bit-identical to s2f_dwconv), its output stays in shared
 * memory as fp16 hi + lo (scaled by the power of two a_pre; choose it so that |x2| * a_pre < 32768) and the 1x1 runs on
 * tcgen05.mma.kind::f16 (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM).  w_pw_packed: the K-major SWIZZLE_128B fp16
 * hi / lo image of W_pw [Cout, Cm] with one power-of-two scale per row (spike2former_b200/ops.py::pack_pw_f16,
 * s2f_sepconv_bpack_bytes bytes); `scale` must contain rowscale[co] / a_pre besides the folded BatchNorm scale.
 * Cm % 64 == 0, 16 <= Cout <= 128, Cout % 4 == 0. */
/* Top-down FPN merge of the two finest levels in one launch (pixel_decoder.py:451-462: `cur = lateral_conv(x);
 * y = cur + F.interpolate(y, cur.shape[-2:], mode='bilinear', align_corners=False); y = output_conv(spike(y))`):
 *   out_spike = NI-LIF( (W a) * scale[co] + shift[co] + bilinear_x2(prev)[pixel, co] )
 * a: int8 levels [n,H,W,Cin] (Cin in {16,32,48,64}), prev: fp32 [n,H/2,W/2,256], Cout = 256.  The lateral conv runs on
 * tcgen05.mma.kind::f16 (levels are exact in fp16; w_packed = fp16 hi / lo image of W [256, 64] with Cin zero-padded to
 * 64, ops.pack_pw_f16), one fp32 accumulator per output; `scale` holds rowscale[co] * a_scale * BN scale.
 * Same arithmetic for the upsample as s2f_upsample_add_lif. */
S2F_API int s2f_fpn_merge_f16(const int8_t* a, const void* w_packed, const float* scale, const float* shift,
                      const float* prev, int8_t* out_spike, int n, int H, int W, int Cin, int Cout, int Hp, int Wp,
                      float d_max, void* stream);

S2F_API int64_t s2f_sepconv_bpack_bytes(int Cm, int Cout);
S2F_API int s2f_sepconv_dwpw(const int8_t* a, float a_scale, const float* w_dw, const void* w_pw_packed, float a_pre,
                     const float* scale, const float* shift, const float* residual, float* out_f32, int8_t* out_spike,
                     int n, int H, int W, int Cm, int Cout, int k, float d_max, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Spike-driven (linear) attention.  q,k,v: int8 levels [n, Ntok, heads*d] (value = level/norm).
 *   kv[h] = K_h^T V_h (exact integers) ; out = (Q_h kv[h]) * out_scale ; spikes = NI-LIF(out).
 * With out_scale = d^-0.5 / norm^3 this is MS_Attention_RepConv_qkv_id (sdtv2.py:335-336); with
 * 1/(sqrt(C) * norm^3) it is {Cross,}MultiHeadAttentionBlock without mask, where
 * (Q K^T / sqrt(C)) V == Q (K^T V) / sqrt(C) (mmcv_spike/transformer.py:262-270, 345-353).
 * q has Nq tokens, k/v have Nk tokens.  kv_ws: workspace of s2f_linear_attn_ws_bytes(n, heads, d) bytes (16-byte aligned);
 * it starts with the int32 [n, heads, d, d] products K^T V.  On aligned operands both contractions run on tcgen05
 * (csrc/attn_tc.cu): K^T V with MN-major operands, Q (K^T V) as a spike GEMM whose per-image weight matrix is the
 * block-diagonal K^T V in three 7-bit digit planes (exact while Nk * 64 < 2^21).
 * out_f32 (optional) receives the pre-activation.
 * q_ld / kv_ld: elements between consecutive token rows of q and of k,v (>= heads*d), so the three
 * operands may be column slices of one fused [n, N, 3C] projection output.  out_ld: elements per output row;
 * columns [heads*d, out_ld) are written as zeros (channel padding to the 16-byte rows TMA needs). */
S2F_API int64_t s2f_linear_attn_ws_bytes(int n, int heads, int d);
S2F_API int s2f_linear_attn(const int8_t* q, const int8_t* k, const int8_t* v, int32_t* kv_ws, int8_t* out_spike,
                    float* out_f32, int n, int Nq, int Nk, int heads, int d, int q_ld, int kv_ld, int out_ld,
                    float out_scale, float d_max, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (c) DCNv3 sampling core.  Replaces dcnv3_core_pytorch (ops_dcnv3/functions/dcnv3_func.py:147-189,
 * F.grid_sample bilinear / zeros / align_corners=False on the 1-pixel zero-padded map).
 *   x [n,H,W,G*Cg] fp32; offset [n,H,W,G*K*K*2] fp32 (x,y interleaved); mask int8 levels [n,H,W,G*K*K]
 *   (value = level*mask_scale); out [n,H,W,G*Cg] fp32.  stride 1, dilation 1, pad = (K-1)/2. */
S2F_API int s2f_dcnv3_gather(const float* x, const float* offset, const int8_t* mask, float mask_scale, float* out, int n,
                     int H, int W, int G, int Cg, int K, float offset_scale, void* stream);

/* Top-down FPN merge: spikes = NI-LIF(cur + bilinear_up(prev)), align_corners=False
 * (pixel_decoder.py:455-462).  cur [n,H,W,C], prev [n,Hp,Wp,C] fp32. */
S2F_API int s2f_upsample_add_lif(const float* cur, const float* prev, int8_t* out_spike, float* out_f32, int n, int H, int W,
                         int Hp, int Wp, int C, float d_max, void* stream);

/* Element-wise residual merge in the layout the reference reinterprets (mmcv_spike/transformer.py:781,829;
 * detr_layers.py:337-338, 553-555):  y[i] = x[i]*scale[i % C] + residual[i]  (scale optional);
 * out_f32 = y, out_spike = NI-LIF(y) (either optional).  N % 4 == 0, C % 4 == 0. */
S2F_API int s2f_affine_add_lif(const float* x, const float* scale, const float* residual, float* out_f32,
                       int8_t* out_spike, int64_t N, int C, float d_max, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Callers either side of the path.
 *
 * s2f_preprocess_u8 replaces SegDataPreProcessor.forward (mmseg/models/data_preprocessor.py:109-152) and the
 * right/bottom padding of stack_batch (mmseg/utils/misc.py:77-93) for n equally sized uint8 images:
 *   img  [n,3,H,W] (chw = 1: what PackSegInputs hands the preprocessor) or [n,H,W,3] (chw = 0: decoder output order)
 *   out  fp32 [n,Hp,Wp,3] channels-last -- the layout the stem reads --
 *        out[y,x,c] = (float(img[y,x,swap_rb ? 2-c : c]) - mean[c]) / std[c]   for y < H, x < W, else pad_val.
 * mean / std are HOST pointers to 3 floats (NULL, NULL: no normalisation, data_preprocessor.py:125).  Bit-exact with
 * the reference's fp32 sub + div. */
S2F_API int s2f_preprocess_u8(const uint8_t* img, int chw, float* out, int n, int H, int W, int Hp, int Wp,
                      const float* mean, const float* std, int swap_rb, float pad_val, void* stream);

/* SegDataPreProcessor fused INTO the stem (data_preprocessor.py:121-126 + sdtv2.py:412-421, first MS_DownSampling):
 * uint8 image -> conv 7x7 stride 2 pad 3 over the normalised image + bias + BatchNorm -> fp32 stream and int8 levels,
 * on tcgen05.mma.kind::i8 with an unsigned A operand (pixels are exact integers).  The fp32 image is never formed.
 *   img        uint8 [n,H,W,3] (chw = 0) or [n,3,H,W] (chw = 1)
 *   w_packed   three int8 digit planes of W[co, cm, kh, kw] / std[cm] in the s2f_pack_weights_i8 layout for a
 *              [Cout, 192] matrix (taps = 1, Cin = 192), K index = kh*24 + j with j running over the 21 bytes of one
 *              kernel row in the image's own memory order ((kw, c) for HWC, (c, kw) for CHW; channel flip folded in),
 *              zeros elsewhere; w_ld = row length in bytes (256)
 *   scale      [Cout]  bn_scale * rowscale
 *   shift_tab  [4][4][4][4][Cout]  folded shift for every (top, bottom, left, right) count of kernel rows / columns cut
 *              off by the image border: bn_shift + bn_scale * (bias - sum over the in-bounds taps of Wq * mean)
 * Cout in {16, 32, 48, 64}; outputs channels-last [n, H/2, W/2, Cout], 16-byte aligned. */
S2F_API int s2f_stem_u8(const uint8_t* img, int chw, const int8_t* w_packed, int w_ld, const float* scale,
                const float* shift_tab, float* out_f32, int8_t* out_spike, int n, int H, int W, int Cout, float d_max,
                void* stream);

/* Training-mode neuron (T = 1, v0 = 0): forward writes the normalised spikes y = rint(clamp(x,0,d_max)) / norm (fp32) and
 * ONE tag byte per neuron = level | 0x80 when x lies outside [0, d_max]; backward is `quant.backward`
 * (surrogate.py:531-538) followed by the "/ norm": gx = gy / norm where the range bit is clear, else 0.  The saved state of
 * a neuron is 1 B instead of the 4 B pre-activation.  All pointers 16-byte aligned. */
S2F_API int s2f_nilif_train_fwd(const float* x, float* y_norm, uint8_t* tag, int64_t N, float d_max, float norm, void* stream);
S2F_API int s2f_nilif_train_bwd(const uint8_t* tag, const float* gy, float* gx, int64_t N, float norm, void* stream);

/* Histogram of spike levels: hist16[l] += #{i : levels[i] == l}, l = 0..15 (uint64, accumulated: zero it first).
 * The firing-rate census of tools/cal_firing_num.py:140-171 (firing rate = 1 - hist[0]/N, mean level = sum l*hist[l]/N)
 * taken from the int8 levels the kernels emit.  levels must be 16-byte aligned. */
S2F_API int s2f_level_hist(const int8_t* levels, int64_t N, unsigned long long* hist16, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Head tail.  sigmoid -> NI-LIF of the stacked decoder states (maskformer_head.py:572-573):
 *   levels = rint(clamp(sigmoid(x), 0, d_max)). */
S2F_API int s2f_sigmoid_lif(const float* x, int8_t* levels, int64_t N, float d_max, void* stream);

/* Semantic inference (decode_heads/maskformer_head.py:163-177):
 *   up = bilinear(mask_pred [n,Q,h,w] -> [H,W], align_corners=False); prob = softmax(cls)[..., :-1]
 *   logits[n,c,H,W] = sum_q prob[n,q,c] * sigmoid(up[n,q,H,W]).
 * mask_pred is given pixel-major [n, h*w, Q] (as the mask GEMM writes it); cls [n,Q,K+1]. */
S2F_API int s2f_semantic_tail(const float* mask_pred, const float* cls, float* logits, float* prob_ws, int n, int Q, int K,
                      int h, int w, int H, int W, void* stream);

/* The same on the tensor cores (tcgen05 kind::f16, bf16 hi+lo split operands, fp32 accumulation in TMEM):
 * Q <= 128, K <= 256.  logits (fp32 [n,K,H,W]) and/or labels (uint8 [n,H,W] = argmax over classes, first maximum,
 * i.e. BaseSegmentor.postprocess_result's argmax, segmentors/base.py:177-188) may be requested.
 * ws: device workspace of s2f_semantic_tail_ws_bytes(n, K) bytes. */
S2F_API int64_t s2f_semantic_tail_ws_bytes(int n, int K);
S2F_API int s2f_semantic_tail_tc(const float* mask_pred, const float* cls, float* logits, uint8_t* labels, void* ws,
                         int n, int Q, int K, int h, int w, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Masked decoder attention in the reference's association order ({Cross,}MultiHeadAttentionBlock.forward,
 * mmcv_spike/transformer.py:262-270 and 345-353): scores = Q K^T * (out_scale folds 1/sqrt(embed_dim) and the spike
 * normalisers); scores.masked_fill(mask, 0); out = scores V; levels = NI-LIF(out).  mask: uint8 [n*heads, Nq, Nk],
 * non-zero = masked (the reference's bool attn_mask reshaped at :266-267 / :350-351), or NULL.  q [n,Nq,q_ld],
 * k / v [n,Nk,kv_ld] int8 levels with head h at columns h*d..h*d+d-1; exact int64 accumulation over the keys. */
S2F_API int s2f_dec_attn(const int8_t* q, const int8_t* k, const int8_t* v, const uint8_t* mask, int8_t* out_spike,
                 float* out_f32, int n, int Nq, int Nk, int heads, int d, int q_ld, int kv_ld, int out_ld,
                 float out_scale, float d_max, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Measurement aid ("int8 peak not measured yet -> measure it before quoting utilisation").
 * Issues `iters` x 4 back-to-back tcgen05.mma (M=128, N=256; kind 0 = kind::i8, K=32; kind 1 = kind::f16 on bf16,
 * K=16) per SM on operands resident in shared memory: the tensor-pipe ceiling for that MMA kind.
 * Returns the operations (2 x MACs) one launch performs, or -1. */
S2F_API int64_t s2f_peak_mma(int kind, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2F_H_ */
