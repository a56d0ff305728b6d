// Shared definitions for the dagl_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace dagl {

constexpr int KS = 7;            // patch size        (CE ksize,   dagl.py:175)
constexpr int KK = KS * KS;      // 49 taps
constexpr int CI = 16;           // inter_channels    (dagl.py:176)
constexpr int ED = 196;          // embedding width   = KK*CI/4  (dagl.py:196-203)
constexpr int VD = KK * CI;      // 784 value-patch width
constexpr int SQ = 4;            // query stride      (stride_1)
constexpr int PADK = 3;          // SAME pad of the stride-1 unfold, also the fold padding (dagl.py:243,267)

// absmax[B][AMAX_STRIDE]: float bits of max Q, max K, max|Theta|, max|G| per image
constexpr int AMAX_STRIDE = 4;
enum { AMAX_Q = 0, AMAX_K = 1, AMAX_THETA = 2, AMAX_G = 3 };

// Heads of one CES stage as a grid dimension (dagl.py:114-118: four CE heads run on the SAME input): the kernels see
// B = (real images) x NH "virtual images", virtual image v = img * NH + head.  Only three things know about heads: the
// kernels that read the shared input b (real image v / NH), the kernels that read weights (head v % NH, through HeadPtrs)
// and the fold, which writes head h's 16 channels at channel offset 16 h of the concatenated output.
constexpr int MAX_HEADS = 4;
struct HeadPtrs { const void* p[MAX_HEADS]; };

struct Geom {
  int B, C, H, W;                // B counts virtual images (real images x NH)
  int nqy, nqx, Nq, Nk;
  int qpad_top, qpad_left;       // SAME pad of the stride-4 unfold (dagl.py:126-136)
  long long y_img_stride;        // elements between consecutive REAL images of the OUTPUT (NH*CI*Nk, or the caller's stage buffer)
  int NH;                        // heads per real image (1: plain CE.forward)
  __host__ __device__ int real_img(int v) const { return NH == 1 ? v : v / NH; }
  __host__ __device__ int head(int v) const { return NH == 1 ? 0 : v % NH; }
  // start of the [16][H][W] result of virtual image v inside the output buffer
  __host__ __device__ size_t y_offset(int v) const { return (size_t)real_img(v) * y_img_stride + (size_t)head(v) * CI * Nk; }
};

__host__ __device__ inline int same_pad_before(int n, int k, int s) {
  int out = (n + s - 1) / s;
  int total = (out - 1) * s + k - n;
  if (total < 0) total = 0;
  return total / 2;
}

inline Geom make_geom(int B, int C, int H, int W, int NH = 1) {
  Geom g;
  g.NH = NH;
  g.B = B * NH; g.C = C; g.H = H; g.W = W;
  g.nqy = (H + SQ - 1) / SQ; g.nqx = (W + SQ - 1) / SQ;
  g.Nq = g.nqy * g.nqx; g.Nk = H * W;
  g.qpad_top = same_pad_before(H, KS, SQ);
  g.qpad_left = same_pad_before(W, KS, SQ);
  g.y_img_stride = (long long)NH * CI * H * W;
  return g;
}

// thread-local call state (error text, launch counter, impl name)
constexpr int PROF_RING = 256;
struct CallState {
  std::string err;
  int launches = 0;
  const char* impl = "none";
  // optional event ring around the dominant kernel (dagl_profile_*)
  int prof_on = 0;                    // 0 off, 1 dominant kernel only, 2 one mark per launch
  int prof_n = 0;
  cudaEvent_t prof_start[PROF_RING];
  cudaEvent_t prof_stop[PROF_RING];
  bool prof_created = false;
};
// record helpers: no-ops unless profiling is enabled on this thread
int prof_begin(cudaStream_t st);
int prof_end(cudaStream_t st);
void prof_mark(cudaStream_t st);      // prof_on == 2: one event after every launch (per-launch timeline of a forward)
CallState& call_state();

#define DAGL_CUDA_OK(expr)                                                           \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::dagl::call_state().err = std::string(#expr) + ": " + cudaGetErrorString(_e); \
      return -4;                                                                     \
    }                                                                                \
  } while (0)

// after every kernel launch (a stream variable `st` is in scope in every launcher)
#define DAGL_LAUNCH_CHECK()                   \
  do {                                        \
    ::dagl::call_state().launches++;          \
    DAGL_CUDA_OK(cudaGetLastError());         \
    ::dagl::prof_mark(st);                    \
  } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// The kernels of a forward are short (14 launches in ~0.9 ms).  Every hot-path kernel starts with pdl_prologue(): it lets the
// NEXT kernel of the stream be launched as soon as all CTAs of this one have started (its CTAs then sit in
// griddepcontrol.wait until this grid has completed and flushed), which hides the launch latency and the ramp-up between
// kernels.  The dependent kernel must be launched with launch_pdl(); without the attribute the two instructions are no-ops.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr = {};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- launchers implemented in the .cu files ------------------------------
int launch_feature_maps(const Geom& g, const float* b, const float* g_w, const float* g_b,
                        const float* th_w, const float* th_b, float* G, float* Th, unsigned* absmax, cudaStream_t st);
int launch_gamma_beta(const Geom& g, const float* b, const float* thr_w, const float* thr_b,
                      const float* bias_w, const float* bias_b, float* gamma, float* beta, cudaStream_t st);
// positions = ny*nx outputs at stride s, window origin (oy*s - off_y, ox*s - off_x)
int launch_embed(const Geom& g, const float* G, const float* fc_w, const float* fc_b, float* out,
                 int ny, int nx, int s, int off_y, int off_x,
                 float* colsum_partial /*nullable [B][nblk][196]*/,
                 unsigned* absmax /*nullable: [B][AMAX_STRIDE]*/, int absmax_slot, cudaStream_t st);
int embed_num_blocks(int npos);
int launch_kbar(const Geom& g, const float* colsum_partial, int nblk, float* Kbar, cudaStream_t st);

// tensor-core embeddings (embed_tc.cu): Q, K fp32
size_t embed_tc_workspace_bytes(const Geom& g);
int embed_tc_num_tiles(const Geom& g);
bool feature_maps_tc_supported(const Geom& g);
size_t feature_maps_tc_workspace_bytes(const Geom& g);
size_t feature_maps_tc_packed_weights_bytes();
int launch_pack_feat_weights(int C, const float* g_w, const float* g_b, const float* th_w, const float* th_b, void* packed,
                             size_t packed_bytes, cudaStream_t st);
// where the feature-map epilogue writes the fp16 images of G (embed_tc.cu's input) and theta (attend_tc.cu's value operand)
struct FeatTargets { uint8_t* ghi; uint8_t* glo; int npg; uint8_t* thp; int np_t; };
const float* embed_tc_fc_meta(const void* packed);    // [l1 fc1, max|b1|, l1 fc2, max|b2|] inside a packed-weights image
// per-head parameter pointers of the (up to MAX_HEADS) heads that share the input
struct HeadWeights {
  const float* g_w[MAX_HEADS]; const float* g_b[MAX_HEADS]; const float* th_w[MAX_HEADS]; const float* th_b[MAX_HEADS];
  const float* fc1_w[MAX_HEADS]; const float* fc1_b[MAX_HEADS]; const float* fc2_w[MAX_HEADS]; const float* fc2_b[MAX_HEADS];
  const float* thr_w[MAX_HEADS]; const float* thr_b[MAX_HEADS]; const float* bias_w[MAX_HEADS]; const float* bias_b[MAX_HEADS];
  const void* packed[MAX_HEADS];       // nullable: dagl_ce_pack_weights_f32 image (fc1 | fc2 | meta, then g/theta)
};
int launch_feature_maps_tc(const Geom& g, const float* b, const HeadWeights& hw, float* G, float* Th, float* gamma,
                           float* beta, unsigned* absmax, void* ws, size_t ws_bytes, bool reuse_b, const FeatTargets& out,
                           cudaStream_t st);
size_t embed_tc_packed_weights_bytes();
int launch_pack_fc_weights(const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b, void* packed,
                           size_t packed_bytes, cudaStream_t st);
// query-side outputs of the embedding epilogue: mu partials [B][nqt*128][2 output halves], gamma / beta padded to whole tiles
struct EmbQOut { float* thr4; const float* kbar; const float* gamma; const float* beta; };
// where the embedding epilogues write the graph kernel's operands (inside its workspace, attend_tc_buffers)
struct EmbTargets {
  uint8_t* ktiles; float* colsum; float* kbar; uint8_t* qtiles; float* thr4; unsigned long long* tilemask;
};
void embed_tc_g_buffers(const Geom& g, void* ws, uint8_t** ghi, uint8_t** glo, int* npg);
int launch_embed_tc(const Geom& g, const HeadWeights& hw, float* Q, float* K, const unsigned* absmax, void* ws, size_t ws_bytes,
                    const float* gamma, const float* beta, const EmbTargets& out, cudaStream_t st);

struct AttendArgs {
  const float* Q; const float* K; const float* Kbar; const float* gamma; const float* beta;
  const float* theta; float* y; float scale;
  uint32_t* mask_bits; int32_t* nnz;
  void* ws; size_t ws_bytes;
  float* kbar_out;     // optional: where a launcher that forms Kbar itself (Kbar == nullptr) also stores it
  // query sharding (multi-GPU): only the 128-query tiles [qt_begin, qt_end) are computed (0 / <=0: all); when
  // rows_out is set the merged rows [B][Nq][49][16] are written there and the fold is left to the caller
  int qt_begin = 0, qt_end = 0;
  float* rows_out = nullptr;
  // full forward: every operand of the graph kernel (key / query tiles, theta image, tile masks, threshold terms, Kbar) was
  // already written into its workspace by the epilogues of the prologue kernels; Q, K, Kbar, gamma, beta, theta are unused
  bool k_packed = false;
  int topk = 0;        // > 0: the legacy fixed-top-k neighbour selection (k edges per query) instead of the adaptive threshold
};
// operand buffers inside the tensor-core graph kernel's workspace that the prologue epilogues write directly
struct AttendBuffers {
  uint8_t* ktiles; float* colsum; uint8_t* qtiles; uint8_t* thp; int np_t; unsigned long long* tilemask;
  float* thr4;      // [B][nqt*128][4]: mu partial (fc1 outputs 0..111), mu partial (112..195), gamma, beta
  float* kbar;
};
AttendBuffers attend_tc_buffers(const Geom& g, void* attend_ws);
size_t merge_fold_scratch_bytes(const Geom& g);      // Omerged [B][Nq][784]
int launch_rows_fold(const Geom& g, int nsplit, const float* Opart, const float* coef, float* Omerged, float* y,
                     int shift_major, cudaStream_t st);
int launch_merge_rows(const Geom& g, int nsplit, int q_begin, int q_end, const float* Opart, const float* coef,
                      float* Omerged, cudaStream_t st);
int launch_fold_rows(const Geom& g, const float* Omerged, float* y, int shift_major, cudaStream_t st);
int launch_fold_partials(const Geom& g, int nsplit, int nparts, const float* Opart, const float* lpart, float* y, cudaStream_t st,
                         bool serialize = false);
int launch_merge_fold(const Geom& g, int nsplit, const float* Opart, const float* mpart, const float* lpart,
                      float* coef, float* Omerged, float* y, int log2_units, int shift_major, float out_scale,
                      cudaStream_t st);
size_t attend_simt_workspace_bytes(const Geom& g);
int launch_attend_simt(const Geom& g, const AttendArgs& a, cudaStream_t st);

// tensor-core path (attend_tc.cu).  `absmax` [B][AMAX_STRIDE] holds max Q, max K, max|theta| as float bits;
// when null the launcher computes it with a reduction kernel (split entry).
size_t attend_tc_workspace_bytes(const Geom& g, int nqt_range = 0);   // nqt_range: query tiles a ranged (sharded) launch covers; 0 = all
// variant 2: 2-CTA clusters sharing P (DSMEM); variant 4: 4-CTA clusters, query tile resident in TMEM
int launch_attend_tc(const Geom& g, const AttendArgs& a, const unsigned* absmax, int variant, cudaStream_t st);

// ResBlock chains (resblock_tc.cu): common.ResBlock, common.py:59-79
struct ResBlockParams {
  const float *w1, *b1;        // body.0  Conv2d(64,64,3,pad 1)
  const float* slope;          // body.1  PReLU weight
  int slope_n;                 // 1 or 64
  const float *w2, *b2;        // body.2
  float res_scale;
  const void* packed;          // nullable: launch_pack_resblock_weights image
};
size_t resblock_packed_weights_bytes();
int launch_pack_resblock_weights(const float* w1, const float* b1, const float* w2, const float* b2, void* packed,
                                 size_t packed_bytes, cudaStream_t st);
size_t resblocks_workspace_bytes(int B, int H, int W, int nblocks);
int launch_resblocks(int B, int H, int W, const float* x, float* y, int nblocks, const ResBlockParams* blocks, void* ws,
                     size_t ws_bytes, int mode, cudaStream_t st);

}  // namespace dagl
