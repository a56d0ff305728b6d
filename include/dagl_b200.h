/* dagl_b200 — C-ABI of the B200-native dynamic attentive graph block.
 *
 * The reference (jianzhangcs/DAGL) has no FFI: the hot path is the Python
 * nn.Module `CE` (DN_Gray/model/dagl.py:174-277) called from `CES.forward`
 * (dagl.py:112-119).  This header is the boundary a maintainer would bind
 * instead of the body of `CE.forward` (dagl.py:207-275); each entry point
 * names the reference lines it replaces.  Plain pointers and sizes only; no
 * torch / C++ types; no exceptions cross it; nothing is allocated or kept by
 * the library between calls (caller owns every buffer incl. the workspace).
 *
 * All device pointers are fp32, contiguous, on the current CUDA device.
 * All calls are asynchronous on `stream` (a cudaStream_t passed as void*).
 * Return value: 0 on success, <0 on error (message via dagl_last_error()).
 */
#ifndef DAGL_B200_H_
#define DAGL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAGL_ABI_VERSION 5

enum {
  DAGL_OK = 0,
  DAGL_ERR_INVALID_ARG = -1,   /* null pointer, non-positive size */
  DAGL_ERR_UNSUPPORTED = -2,   /* ksize/stride/channels the kernels are not built for */
  DAGL_ERR_WORKSPACE = -3,     /* workspace too small */
  DAGL_ERR_CUDA = -4           /* a CUDA runtime call failed; see dagl_last_error() */
};

/* Which fused graph kernel to run. */
enum {
  DAGL_IMPL_AUTO = 0,   /* the tensor-core kernel (currently always DAGL_IMPL_TC4) */
  DAGL_IMPL_SIMT = 1,   /* fp32 CUDA-core kernel (bit-faithful neighbour mask) */
  DAGL_IMPL_TC = 2,     /* tcgen05 tensor-core kernel (split-fp16 scores, fp16 P.V), 2-CTA clusters sharing P; A/B aid */
  DAGL_IMPL_TC4 = 4     /* 4-CTA clusters, query tile resident in TMEM (A operand of the score MMAs) */
};

/* Borrowed device pointers to one CE head's parameters, in the reference's
 * state_dict layout (CE.__init__, dagl.py:175-205):
 *   g      Conv2d(C,16,3,pad 1)   weight [16][C][3][3]   bias [16]
 *   theta  Conv2d(C,16,1)         weight [16][C]         bias [16]
 *   fc1    Linear(784,196)        weight [196][784]      bias [196]   (query embedding)
 *   fc2    Linear(784,196)        weight [196][784]      bias [196]   (key embedding)
 *   thr    Conv2d(C,1,7,stride 4) weight [C][7][7]       bias [1]
 *   bias   Conv2d(C,1,7,stride 4) weight [C][7][7]       bias [1]
 * `W` (dagl.py:193) is dead in forward and is not passed.                   */
typedef struct DaglCEWeights {
  const float* g_w;     const float* g_b;
  const float* theta_w; const float* theta_b;
  const float* fc1_w;   const float* fc1_b;
  const float* fc2_w;   const float* fc2_b;
  const float* thr_w;   const float* thr_b;
  const float* bias_w;  const float* bias_b;
  int32_t in_channels;      /* C; kernels are built for C % 4 == 0, C <= 256 */
  int32_t inter_channels;   /* must be 16 */
  int32_t ksize;            /* must be 7  */
  int32_t stride_q;         /* must be 4  (stride_1)  */
  int32_t stride_k;         /* must be 1  (stride_2)  */
  float softmax_scale;      /* reference default 10   */
  const void* packed_fc;    /* optional (may be NULL): fc1/fc2 and g/theta pre-packed by dagl_ce_pack_weights_f32
                               for the tensor-core kernels; saves the per-call weight packing at inference   */
  int32_t legacy_topk;      /* 0: the shipping CE (adaptive per-query threshold, dagl.py:256-257; thr/bias are used).
                               k in [1, 64]: the neighbour selection of the reference's legacy fixed-top-k variant
                               (DN_Gray/model/.ipynb_checkpoints/GReccR2b_3mh_1-checkpoint.py:243-250): the k keys with
                               the largest scores, logits scale * S on them (thr / bias are ignored; ties at the k-th
                               score are all kept).  Tensor-core impls only.                                     */
} DaglCEWeights;

int32_t dagl_abi_version(void);

/* Thread-local description of the last error returned on this thread. */
const char* dagl_last_error(void);

/* Bytes of device workspace dagl_ce_forward_* needs for a [B,C,H,W] input (sized for the launch of the whole image:
 * a query-sharded rows call has its own query, dagl_ce_rows_workspace_bytes). */
size_t dagl_ce_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W);
/* The same for one given impl: the tensor-core forward keeps no fp32 intermediates (G, theta, Q, K), so without the debug
 * outputs it needs ~35 % less (`debug` != 0: the size dagl_ce_forward_debug_f32 / dagl_ce_workspace_view need). */
size_t dagl_ce_workspace_bytes_ex(int32_t B, int32_t C, int32_t H, int32_t W, int32_t impl, int32_t debug);

/* y[B,16,H,W] = CE.forward(b[B,C,H,W])          — replaces dagl.py:207-275. */
int32_t dagl_ce_forward_f32(const DaglCEWeights* w, const float* b, float* y,
                            int32_t B, int32_t H, int32_t W,
                            void* workspace, size_t workspace_bytes,
                            int32_t impl, void* stream);

/* Optional weight pre-packing.  The library keeps no state between calls, so by default every forward re-packs
 * fc1 / fc2 (fp32 [196][784]) and g / theta into the fp16 hi/lo tensor-core operand images (~40 us).  A caller
 * whose weights are constant (inference) can do it once: pack into a caller-owned device buffer of
 * dagl_ce_packed_weights_bytes() and pass it as DaglCEWeights.packed_fc; it must be re-packed whenever fc1_w,
 * fc2_w, g_w or theta_w change (w->in_channels must be set).  Only the tensor-core implementations read it.   */
size_t dagl_ce_packed_weights_bytes(void);
int32_t dagl_ce_pack_weights_f32(const DaglCEWeights* w, void* packed, size_t packed_bytes, void* stream);

/* The heads of one CES stage (reference CES.forward, dagl.py:114-118: `torch.cat([c_1(x), .., c_4(x)], dim=1)`):
 * head h of `heads[0..n_heads)` runs on the shared input b[B,C,H,W] and writes its 16 channels straight into
 * channels [16h, 16h+16) of ycat[B, 16*n_heads, H, W] — the concatenation is never a separate copy.  The 1x1 merge
 * conv and the residual of the stage stay with the caller.
 * With dagl_ces_workspace_bytes() of workspace (and a tensor-core impl, 64 input channels, one softmax scale) the heads
 * are a GRID DIMENSION: every kernel of the forward runs once over B x n_heads "virtual images" (the shared input is
 * repacked once; ~15 launches per stage instead of ~15 per head).  With only dagl_ce_workspace_bytes() the heads run one
 * after the other.  The two routes agree up to the fp32 summation order of the key-split partial sums.              */
size_t dagl_ces_workspace_bytes(int32_t n_heads, int32_t B, int32_t C, int32_t H, int32_t W);
int32_t dagl_ces_heads_forward_f32(const DaglCEWeights* const* heads, int32_t n_heads, const float* b, float* ycat,
                                   int32_t B, int32_t H, int32_t W,
                                   void* workspace, size_t workspace_bytes,
                                   int32_t impl, void* stream);

/* Same, and also reports the neighbour selection of dagl.py:256-257:
 *   mask_bits [B][Nq][ceil(Nk/32)] : bit j of word w set <=> key 32w+j is a
 *                                    neighbour of the query (mask_b != 0)
 *   nnz       [B][Nq]              : neighbours per query
 * either may be NULL.  Nq = ceil(H/4)*ceil(W/4), Nk = H*W (row-major).       */
int32_t dagl_ce_forward_debug_f32(const DaglCEWeights* w, const float* b, float* y,
                                  int32_t B, int32_t H, int32_t W,
                                  void* workspace, size_t workspace_bytes,
                                  int32_t impl, void* stream,
                                  uint32_t* mask_bits, int32_t* nnz);

/* Host-buffer entry: b_host / y_host are HOST pointers (pinned for async
 * copies); the H2D copy of b, the forward and the D2H copy of y are all
 * enqueued on `stream`.  Needs dagl_ce_workspace_bytes_ex(.., impl, 0) +
 * dagl_ce_host_staging_bytes() of device workspace (dagl_ce_workspace_bytes() + staging is always enough).   */
size_t dagl_ce_host_staging_bytes(int32_t B, int32_t C, int32_t H, int32_t W);
int32_t dagl_ce_forward_host_f32(const DaglCEWeights* w, const float* b_host, float* y_host,
                                 int32_t B, int32_t H, int32_t W,
                                 void* workspace, size_t workspace_bytes,
                                 int32_t impl, void* stream);

/* Query sharding across GPUs (one image, several ranks; SURVEY §8e scheme 3).  Rows of the score matrix are
 * independent given all keys, so each rank runs the (cheap) prologue redundantly and the fused graph stage only
 * for the 128-query tiles [q_tile_begin, q_tile_end); it writes the merged, normalised aggregation rows
 *   rows [B][Nq][49 shifts (dy*7+dx)][16 channels]      (dagl.py:263-264, before the fold)
 * for its queries (other rows untouched).  After the ranks exchange their rows (one all-gather),
 * dagl_ce_fold_rows_f32 performs the fold + coverage normalisation (dagl.py:265-272).                         */
int32_t dagl_ce_num_query_tiles(int32_t H, int32_t W);
/* workspace for a rows call over [q_tile_begin, q_tile_end) (the key-split factor depends on the range) */
size_t dagl_ce_rows_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W, int32_t q_tile_begin, int32_t q_tile_end);
int32_t dagl_ce_forward_rows_f32(const DaglCEWeights* w, const float* b, float* rows,
                                 int32_t B, int32_t H, int32_t W,
                                 int32_t q_tile_begin, int32_t q_tile_end,
                                 void* workspace, size_t workspace_bytes, int32_t impl /* auto, tc or tc4 */, void* stream);
int32_t dagl_ce_fold_rows_f32(const float* rows, float* y, int32_t B, int32_t H, int32_t W, void* stream);

/* Split entry for the fused graph stage alone (dagl.py:250-272), taking the
 * embeddings as inputs:
 *   Q [B][Nq][196], K [B][Nk][196] (post-ReLU), Kbar [B][196] = mean_k K,
 *   gamma, beta [B][Nq], theta [B][16][H][W]  ->  y [B][16][H][W].           */
size_t dagl_graph_attend_workspace_bytes(int32_t B, int32_t H, int32_t W);
int32_t dagl_graph_attend_f32(const float* Q, const float* K, const float* Kbar,
                              const float* gamma, const float* beta, const float* theta,
                              float* y, int32_t B, int32_t H, int32_t W, float softmax_scale,
                              void* workspace, size_t workspace_bytes,
                              int32_t impl, void* stream,
                              uint32_t* mask_bits, int32_t* nnz);

/* Backward of the fused graph stage (the reference trains through plain autograd, DN_Gray/trainer.py:51-57, i.e. through
 * dagl.py:250-272).  Inputs: the embeddings Q [B][Nq][196], K [B][Nk][196] (post-ReLU), theta [B][16][H][W], gamma, beta
 * [B][Nq] and dy [B][16][H][W]; outputs the gradients with respect to the first five (same shapes), with the reference's
 * gradient semantics: through S, the row mean, relu(S - mu*gamma + beta) where it multiplies the logits and the softmax;
 * none through the 0/1 neighbour indicator.  fp32 CUDA-core kernels, deterministic.  The convolutions / linears in front
 * of the graph stage are differentiated by the caller (dagl_b200/autograd.py uses PyTorch for them).               */
size_t dagl_graph_attend_backward_workspace_bytes(int32_t B, int32_t H, int32_t W);
int32_t dagl_graph_attend_backward_f32(const float* Q, const float* K, const float* theta, const float* gamma,
                                       const float* beta, const float* dy, float* dQ, float* dK, float* dtheta,
                                       float* dgamma, float* dbeta, int32_t B, int32_t H, int32_t W, float softmax_scale,
                                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- ResBlock chains: the callers either side of the graph blocks -----------------------------------------------
 * common.ResBlock (DN_Gray/model/common.py:59-79): y = conv2(PReLU(conv1(x))) * res_scale + x with two
 * Conv2d(64,64,3,padding 1) (common.default_conv, common.py:8-11).  CES.RBS1 / RBS2 (dagl.py:86-101, 115, 117) are chains
 * of four, RR.body (dagl.py:27-34) holds two chains of eight around the CES module.
 * dagl_resblocks_forward_f32 runs a whole chain: every convolution is one tcgen05 kernel (split-fp16 x3, fp32-accurate)
 * whose epilogue applies bias / PReLU / residual and writes the next convolution's operand image, so no elementwise
 * kernel and no fp32 re-pack runs inside the chain.  Borrowed device pointers in the reference's state_dict layout:
 *   body.0.weight [64][64][3][3]  body.0.bias [64] (NULL: no bias)
 *   body.1.weight [1] or [64]     (PReLU)
 *   body.2.weight [64][64][3][3]  body.2.bias [64] (NULL: no bias)                                                     */
typedef struct DaglResBlockWeights {
  const float* conv1_w; const float* conv1_b;
  const float* prelu_w; int32_t prelu_n;
  const float* conv2_w; const float* conv2_b;
  float res_scale;          /* ResBlock.res_scale (common.py:73,76) */
  const void* packed;       /* optional (may be NULL): dagl_resblock_pack_weights_f32 image; saves the per-call packing */
} DaglResBlockWeights;

size_t dagl_resblock_packed_weights_bytes(void);
int32_t dagl_resblock_pack_weights_f32(const DaglResBlockWeights* w, void* packed, size_t packed_bytes, void* stream);
size_t dagl_resblocks_workspace_bytes(int32_t n_blocks, int32_t B, int32_t C, int32_t H, int32_t W);
/* x, y: fp32 [B][C][H][W] (C must be 64; y may alias x).  mode 0: CTA-pair kernel (cta_group::2), 1: single-CTA kernel,
 * 2: chosen by size (the default of the Python binding). */
int32_t dagl_resblocks_forward_f32(const DaglResBlockWeights* blocks, int32_t n_blocks, const float* x, float* y,
                                   int32_t B, int32_t C, int32_t H, int32_t W,
                                   void* workspace, size_t workspace_bytes, int32_t mode, void* stream);

/* Intermediate views inside the workspace after dagl_ce_forward_* (device
 * pointers, valid until the workspace is reused): which = 0 G [B,16,H,W],
 * 1 theta [B,16,H,W], 2 gamma [B,Nq], 3 beta [B,Nq], 4 Q [B,Nq,196],
 * 5 K [B,Nk,196], 6 Kbar [B,196].  Used by the parity tests.                 */
const float* dagl_ce_workspace_view(void* workspace, int32_t which,
                                    int32_t B, int32_t C, int32_t H, int32_t W);

/* Name of the kernel family actually used by the last forward on this thread
 * ("simt", "tc" or "tc4"); lets callers assert that no fallback happened.    */
const char* dagl_last_impl(void);

/* Number of kernel launches issued by the last forward on this thread.      */
int32_t dagl_last_launch_count(void);

/* Profiling aid for bench.py (not used on the product path).  When enabled on
 * this thread the library brackets the dominant fused graph kernel of every
 * forward with a pair of CUDA events recorded on the caller's stream (a ring
 * of 256 pairs).  dagl_profile_read() synchronises on the recorded events and
 * returns up to `max` kernel durations in milliseconds, oldest first, and
 * resets the ring.  Returns the number written, <0 on error.
 * on = 2: development aid — one event after EVERY kernel launch of a forward; read() then returns the time between
 * consecutive launches' completions (entry 0 is ~0), i.e. a warm, in-pipeline per-launch timeline.             */
int32_t dagl_profile_enable(int32_t on);
int32_t dagl_profile_read(float* ms, int32_t max);

#ifdef __cplusplus
}
#endif
#endif /* DAGL_B200_H_ */
