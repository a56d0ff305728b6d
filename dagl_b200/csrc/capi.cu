// C-ABI of dagl_b200 (see include/dagl_b200.h).  Thin dispatch only: validates
// arguments, carves the caller-owned workspace and enqueues the kernels on the
// caller's stream.  No allocation, no synchronisation, no CPU fallback.
#include "../../include/dagl_b200.h"
#include "common.cuh"

namespace dagl {
CallState& call_state() {
  static thread_local CallState s;
  return s;
}

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

void prof_mark(cudaStream_t st) {
  CallState& s = call_state();
  if (s.prof_on != 2 || s.prof_n >= PROF_RING) return;
  if (s.prof_n == 0) cudaEventRecord(s.prof_start[0], st);       // origin: right after the first launch
  cudaEventRecord(s.prof_stop[s.prof_n++], st);
}

int prof_begin(cudaStream_t st) {
  CallState& s = call_state();
  if (s.prof_on != 1 || s.prof_n >= PROF_RING) return 0;
  DAGL_CUDA_OK(cudaEventRecord(s.prof_start[s.prof_n], st));
  return 0;
}
int prof_end(cudaStream_t st) {
  CallState& s = call_state();
  if (s.prof_on != 1 || s.prof_n >= PROF_RING) return 0;
  DAGL_CUDA_OK(cudaEventRecord(s.prof_stop[s.prof_n], st));
  s.prof_n++;
  return 0;
}

// Workspace carve-up shared by workspace_bytes / forward / workspace_view.
struct WsLayout {
  size_t gamma, beta, Kbar, absmax, packw, feat, embed, attend, lean_total;   // what the tensor-core forward needs
  size_t G, Th, Q, K, kpart, total;                                            // + fp32 intermediates (CUDA-core prologue, debug entry)
  int kblocks_simt;
};

// The fp32 copies of G, theta, Q, K come LAST: the tensor-core forward never touches them (every kernel writes the next
// one's fp16 operands from its epilogue), so a caller that only runs the product path can hand in `lean_total` bytes
// (dagl_ce_workspace_bytes_ex); the debug entry and the CUDA-core prologue need `total`.  All other offsets are the same
// in both cases.
static WsLayout ws_layout(const Geom& g, int nqt_range = 0) {
  WsLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  const size_t f = sizeof(float);
  L.kblocks_simt = embed_num_blocks(g.Nk);
  L.gamma = take((size_t)g.B * g.Nq * f);
  L.beta = take((size_t)g.B * g.Nq * f);
  L.Kbar = take((size_t)g.B * ED * f);
  L.absmax = take((size_t)g.B * AMAX_STRIDE * sizeof(unsigned));
  L.packw = take((size_t)g.NH * (embed_tc_packed_weights_bytes() + feature_maps_tc_packed_weights_bytes()));   // per-call weight packing
  L.feat = take(feature_maps_tc_workspace_bytes(g));
  L.embed = take(embed_tc_workspace_bytes(g));
  L.attend = off;
  const size_t a_simt = attend_simt_workspace_bytes(g), a_tc = attend_tc_workspace_bytes(g, nqt_range);
  off += align_up(a_simt > a_tc ? a_simt : a_tc);
  L.lean_total = off;
  L.G = take((size_t)g.B * CI * g.Nk * f);
  L.Th = take((size_t)g.B * CI * g.Nk * f);
  L.Q = take((size_t)g.B * g.Nq * ED * f);
  L.K = take((size_t)g.B * g.Nk * ED * f);
  L.kpart = take((size_t)g.B * L.kblocks_simt * ED * f);
  L.total = off;
  return L;
}

static int check_weights(const DaglCEWeights* w) {
  if (!w || !w->g_w || !w->g_b || !w->theta_w || !w->theta_b || !w->fc1_w || !w->fc1_b || !w->fc2_w ||
      !w->fc2_b || !w->thr_w || !w->thr_b || !w->bias_w || !w->bias_b) {
    call_state().err = "null weight pointer";
    return DAGL_ERR_INVALID_ARG;
  }
  if (w->inter_channels != CI || w->ksize != KS || w->stride_q != SQ || w->stride_k != 1) {
    call_state().err = "unsupported CE configuration: kernels are built for ksize=7, stride_1=4, stride_2=1, inter_channels=16";
    return DAGL_ERR_UNSUPPORTED;
  }
  if (w->in_channels <= 0 || w->in_channels > 256 || (w->in_channels % 4) != 0) {
    call_state().err = "unsupported in_channels (need a multiple of 4 in [4,256])";
    return DAGL_ERR_UNSUPPORTED;
  }
  return 0;
}

static int check_shape(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) {
    call_state().err = "non-positive shape";
    return DAGL_ERR_INVALID_ARG;
  }
  if ((long long)H * W > (1 << 22) || (long long)B * H * W * CI > (1ll << 30)) {
    call_state().err = "shape too large";
    return DAGL_ERR_UNSUPPORTED;
  }
  return 0;
}

static int run_attend(const Geom& g, const AttendArgs& a, int impl, const unsigned* absmax, cudaStream_t st) {
  // measured (tools/big_shapes.py, round 1): the 4-CTA-cluster kernel is 10-18 % ahead of the 2-CTA one at every size tried
  // (64^2 .. 512^2 keys), so it is the default at every size; the 2-CTA kernel stays selectable for A/B checks
  if (impl == DAGL_IMPL_AUTO) impl = DAGL_IMPL_TC4;
  if (impl == DAGL_IMPL_TC) {
    call_state().impl = "tc";
    return launch_attend_tc(g, a, absmax, 2, st);
  }
  if (impl == DAGL_IMPL_TC4) {
    call_state().impl = "tc4";
    return launch_attend_tc(g, a, absmax, 4, st);
  }
  if (impl != DAGL_IMPL_SIMT) {
    call_state().err = "unknown impl";
    return DAGL_ERR_INVALID_ARG;
  }
  call_state().impl = "simt";
  return launch_attend_simt(g, a, st);
}

// One forward of `nh` heads (1 <= nh <= MAX_HEADS) that share the input b.  nh == 1 is CE.forward; nh > 1 is the head
// batch of a CES stage (heads as a grid dimension; tensor-core path only): every kernel runs once over B x nh virtual
// images and head h's result lands at channel offset 16 h of y (y_img_stride between real images).
static int forward_impl(const DaglCEWeights* const* heads, int nh, const float* b, float* y, int B, int H, int W, void* ws,
                        size_t ws_bytes, int impl, cudaStream_t st, uint32_t* mask_bits, int32_t* nnz,
                        float* rows_out = nullptr, int qt_begin = 0, int qt_end = 0, long long y_img_stride = 0,
                        bool reuse_b = false) {
  call_state().launches = 0;
  call_state().impl = "none";
  if (!heads || nh < 1 || nh > MAX_HEADS) {
    call_state().err = "bad head list";
    return DAGL_ERR_INVALID_ARG;
  }
  int rc;
  for (int h = 0; h < nh; ++h) {
    if ((rc = check_weights(heads[h]))) return rc;
    if (heads[h]->in_channels != heads[0]->in_channels || heads[h]->softmax_scale != heads[0]->softmax_scale ||
        heads[h]->legacy_topk != heads[0]->legacy_topk) {
      call_state().err = "heads of one batched stage call must share in_channels, softmax_scale and legacy_topk";
      return DAGL_ERR_INVALID_ARG;
    }
  }
  const DaglCEWeights* w = heads[0];
  if (w->legacy_topk < 0 || w->legacy_topk > 64 || (w->legacy_topk > 0 && impl == DAGL_IMPL_SIMT)) {
    call_state().err = "legacy_topk must be in [0, 64] and needs a tensor-core impl";
    return DAGL_ERR_UNSUPPORTED;
  }
  rc = check_shape(B * nh, H, W);
  if (rc) return rc;
  if (!b || (!y && !rows_out) || !ws) {
    call_state().err = "null buffer";
    return DAGL_ERR_INVALID_ARG;
  }
  Geom g = make_geom(B, w->in_channels, H, W, nh);
  if (nh > 1 && (impl == DAGL_IMPL_SIMT || !feature_maps_tc_supported(g) || rows_out || mask_bits || nnz)) {
    call_state().err = "batched heads need the tensor-core path (in_channels 64), without debug / row outputs";
    return DAGL_ERR_UNSUPPORTED;
  }
  if (y_img_stride != 0) {
    if (y_img_stride < g.y_img_stride) {
      call_state().err = "output image stride smaller than one [16*heads,H,W] result";
      return DAGL_ERR_INVALID_ARG;
    }
    g.y_img_stride = y_img_stride;
  }
  const WsLayout L = ws_layout(g, rows_out ? qt_end - qt_begin : 0);
  const bool debug = (mask_bits != nullptr) || (nnz != nullptr);
  const bool tc_prologue = impl != DAGL_IMPL_SIMT && feature_maps_tc_supported(g);
  if (ws_bytes < ((tc_prologue && !debug) ? L.lean_total : L.total)) {
    call_state().err = "workspace too small";
    return DAGL_ERR_WORKSPACE;
  }
  char* base = static_cast<char*>(ws);
  float* G = reinterpret_cast<float*>(base + L.G);
  float* Th = reinterpret_cast<float*>(base + L.Th);
  float* gamma = reinterpret_cast<float*>(base + L.gamma);
  float* beta = reinterpret_cast<float*>(base + L.beta);
  float* Q = reinterpret_cast<float*>(base + L.Q);
  float* K = reinterpret_cast<float*>(base + L.K);
  float* kpart = reinterpret_cast<float*>(base + L.kpart);
  float* Kbar = reinterpret_cast<float*>(base + L.Kbar);
  unsigned* absmax = reinterpret_cast<unsigned*>(base + L.absmax);

  HeadWeights hw{};
  for (int h = 0; h < nh; ++h) {
    const DaglCEWeights* x = heads[h];
    hw.g_w[h] = x->g_w; hw.g_b[h] = x->g_b; hw.th_w[h] = x->theta_w; hw.th_b[h] = x->theta_b;
    hw.fc1_w[h] = x->fc1_w; hw.fc1_b[h] = x->fc1_b; hw.fc2_w[h] = x->fc2_w; hw.fc2_b[h] = x->fc2_b;
    hw.thr_w[h] = x->thr_w; hw.thr_b[h] = x->thr_b; hw.bias_w[h] = x->bias_w; hw.bias_b[h] = x->bias_b;
    hw.packed[h] = x->packed_fc;
  }

  bool k_packed = false;
  DAGL_CUDA_OK(cudaMemsetAsync(absmax, 0, (size_t)g.B * AMAX_STRIDE * sizeof(unsigned), st));
  if (tc_prologue) {
    // Tensor-core prologue: every kernel writes the next one's operands from its epilogue.  feature maps -> fp16 images of G
    // (embedding input) and theta (graph kernel's value operand) + all fp16 scales; key embedding -> key tiles + column sums;
    // Kbar; query embedding -> query tiles + threshold terms.  The fp32 intermediates exist only for the debug entry
    // (parity tests read them through dagl_ce_workspace_view).
    const size_t pack_bytes = embed_tc_packed_weights_bytes() + feature_maps_tc_packed_weights_bytes();
    for (int h = 0; h < nh; ++h)
      if (hw.packed[h] == nullptr) {                    // the caller did not pre-pack (dagl_ce_pack_weights_f32): do it per call
        char* slot = base + L.packw + (size_t)h * pack_bytes;
        if ((rc = launch_pack_fc_weights(hw.fc1_w[h], hw.fc1_b[h], hw.fc2_w[h], hw.fc2_b[h], slot, embed_tc_packed_weights_bytes(), st))) return rc;
        if ((rc = launch_pack_feat_weights(g.C, hw.g_w[h], hw.g_b[h], hw.th_w[h], hw.th_b[h], slot + embed_tc_packed_weights_bytes(),
                                           feature_maps_tc_packed_weights_bytes(), st))) return rc;
        hw.packed[h] = slot;
      }
    const AttendBuffers ab = attend_tc_buffers(g, base + L.attend);
    uint8_t *ghi, *glo;
    int npg;
    embed_tc_g_buffers(g, base + L.embed, &ghi, &glo, &npg);
    const FeatTargets ft{ghi, glo, npg, ab.thp, ab.np_t};
    if ((rc = launch_feature_maps_tc(g, b, hw, debug ? G : nullptr, debug ? Th : nullptr, gamma, beta, absmax, base + L.feat,
                                     L.embed - L.feat, reuse_b, ft, st))) return rc;
    const EmbTargets et{ab.ktiles, ab.colsum, Kbar, ab.qtiles, ab.thr4, ab.tilemask};
    if ((rc = launch_embed_tc(g, hw, debug ? Q : nullptr, debug ? K : nullptr, absmax, base + L.embed, L.attend - L.embed, gamma,
                              beta, et, st))) return rc;
    k_packed = true;
  } else {
    // fp32 CUDA-core prologue (impl simt, or an input channel count the tensor-core feature kernel is not built for): the
    // graph kernel packs its operands itself from the fp32 embeddings
    if ((rc = launch_feature_maps(g, b, w->g_w, w->g_b, w->theta_w, w->theta_b, G, Th, absmax, st))) return rc;
    if ((rc = launch_gamma_beta(g, b, w->thr_w, w->thr_b, w->bias_w, w->bias_b, gamma, beta, st))) return rc;
    if ((rc = launch_embed(g, G, w->fc1_w, w->fc1_b, Q, g.nqy, g.nqx, SQ, g.qpad_top, g.qpad_left, nullptr, absmax, AMAX_Q, st))) return rc;
    if ((rc = launch_embed(g, G, w->fc2_w, w->fc2_b, K, g.H, g.W, 1, PADK, PADK, kpart, absmax, AMAX_K, st))) return rc;
    if ((rc = launch_kbar(g, kpart, L.kblocks_simt, Kbar, st))) return rc;
  }

  AttendArgs a;
  a.Q = Q; a.K = K; a.Kbar = Kbar; a.gamma = gamma; a.beta = beta; a.theta = Th; a.y = y;
  a.scale = w->softmax_scale; a.mask_bits = mask_bits; a.nnz = nnz;
  a.ws = base + L.attend; a.ws_bytes = L.lean_total - L.attend;
  a.k_packed = k_packed;
  a.topk = w->legacy_topk;
  a.kbar_out = reinterpret_cast<float*>(base + L.Kbar);     // keeps dagl_ce_workspace_view(…, 6) valid on every path
  a.rows_out = rows_out; a.qt_begin = qt_begin; a.qt_end = qt_end;
  return run_attend(g, a, impl, absmax, st);
}
}  // namespace dagl

using namespace dagl;

extern "C" {

int32_t dagl_abi_version(void) { return DAGL_ABI_VERSION; }

const char* dagl_last_error(void) { return call_state().err.c_str(); }
const char* dagl_last_impl(void) { return call_state().impl; }
int32_t dagl_last_launch_count(void) { return call_state().launches; }

size_t dagl_ce_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return ws_layout(make_geom(B, C, H, W)).total;
}

size_t dagl_ce_workspace_bytes_ex(int32_t B, int32_t C, int32_t H, int32_t W, int32_t impl, int32_t debug) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  const Geom g = make_geom(B, C, H, W);
  const WsLayout L = ws_layout(g);
  return (impl != DAGL_IMPL_SIMT && feature_maps_tc_supported(g) && !debug) ? L.lean_total : L.total;
}

size_t dagl_ce_rows_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W, int32_t q_tile_begin, int32_t q_tile_end) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || q_tile_end <= q_tile_begin) return 0;
  const Geom g = make_geom(B, C, H, W);
  const WsLayout L = ws_layout(g, q_tile_end - q_tile_begin);
  return feature_maps_tc_supported(g) ? L.lean_total : L.total;
}

size_t dagl_ces_workspace_bytes(int32_t n_heads, int32_t B, int32_t C, int32_t H, int32_t W) {
  if (n_heads <= 0 || B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  const int nh = n_heads < MAX_HEADS ? n_heads : MAX_HEADS;
  // the batched route only exists on the tensor-core path (lean layout); the serial per-head route may be the CUDA-core one
  const size_t batched = ws_layout(make_geom(B, C, H, W, nh)).lean_total, single = ws_layout(make_geom(B, C, H, W)).total;
  return batched > single ? batched : single;
}

int32_t dagl_ce_forward_f32(const DaglCEWeights* w, const float* b, float* y, int32_t B, int32_t H, int32_t W,
                            void* workspace, size_t workspace_bytes, int32_t impl, void* stream) {
  return forward_impl(&w, 1, b, y, B, H, W, workspace, workspace_bytes, impl, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

int32_t dagl_ce_forward_debug_f32(const DaglCEWeights* w, const float* b, float* y, int32_t B, int32_t H, int32_t W,
                                  void* workspace, size_t workspace_bytes, int32_t impl, void* stream,
                                  uint32_t* mask_bits, int32_t* nnz) {
  return forward_impl(&w, 1, b, y, B, H, W, workspace, workspace_bytes, impl, static_cast<cudaStream_t>(stream), mask_bits, nnz);
}

int32_t dagl_ces_heads_forward_f32(const DaglCEWeights* const* heads, int32_t n_heads, const float* b, float* ycat,
                                   int32_t B, int32_t H, int32_t W, void* workspace, size_t workspace_bytes,
                                   int32_t impl, void* stream) {
  if (!heads || n_heads <= 0 || n_heads > 16) {
    call_state().err = "bad head list";
    return DAGL_ERR_INVALID_ARG;
  }
  for (int h = 0; h < n_heads; ++h)
    if (!heads[h]) { call_state().err = "null head"; return DAGL_ERR_INVALID_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long stride = (long long)n_heads * CI * H * W;
  int total_launches = 0;
  // Heads as a grid dimension: groups of up to MAX_HEADS heads run as ONE set of launches over B x heads virtual images
  // (tensor-core path, 64 input channels, equal softmax scale).  Anything else takes the serial per-head route.
  bool batch_ok = impl != DAGL_IMPL_SIMT && n_heads > 1 && heads[0]->in_channels == 64 &&
                  workspace_bytes >= dagl_ces_workspace_bytes(n_heads, B, heads[0]->in_channels, H, W);
  for (int h = 1; h < n_heads && batch_ok; ++h)
    batch_ok = heads[h]->in_channels == heads[0]->in_channels && heads[h]->softmax_scale == heads[0]->softmax_scale &&
               heads[h]->legacy_topk == heads[0]->legacy_topk;
  if (batch_ok) {
    for (int h0 = 0; h0 < n_heads; h0 += MAX_HEADS) {
      const int nh = n_heads - h0 < MAX_HEADS ? n_heads - h0 : MAX_HEADS;
      float* yh = ycat ? ycat + (size_t)h0 * CI * H * W : nullptr;
      const int rc = forward_impl(heads + h0, nh, b, yh, B, H, W, workspace, workspace_bytes, impl, st, nullptr, nullptr, nullptr,
                                  0, 0, stride, /*reuse_b=*/h0 > 0 && nh == MAX_HEADS);   // same workspace layout as the first group
      if (rc) return rc;
      total_launches += call_state().launches;
    }
    call_state().launches = total_launches;
    return 0;
  }
  for (int h = 0; h < n_heads; ++h) {
    float* yh = ycat ? ycat + (size_t)h * CI * H * W : nullptr;
    // the heads share the input: its fp16 repack (and maximum) from head 0 stays valid in the workspace for the others
    const bool reuse_b = h > 0 && heads[h]->in_channels == heads[0]->in_channels;
    const int rc = forward_impl(heads + h, 1, b, yh, B, H, W, workspace, workspace_bytes, impl, st, nullptr, nullptr, nullptr, 0, 0,
                                stride, reuse_b);
    if (rc) return rc;
    total_launches += call_state().launches;
  }
  call_state().launches = total_launches;
  return 0;
}

size_t dagl_ce_packed_weights_bytes(void) { return embed_tc_packed_weights_bytes() + feature_maps_tc_packed_weights_bytes(); }

int32_t dagl_ce_pack_weights_f32(const DaglCEWeights* w, void* packed, size_t packed_bytes, void* stream) {
  call_state().launches = 0;
  if (!w || !w->fc1_w || !w->fc2_w || !w->fc1_b || !w->fc2_b || !w->g_w || !w->g_b || !w->theta_w || !w->theta_b || !packed) {
    call_state().err = "null pointer";
    return DAGL_ERR_INVALID_ARG;
  }
  if (w->inter_channels != CI || w->ksize != KS) {
    call_state().err = "unsupported CE configuration";
    return DAGL_ERR_UNSUPPORTED;
  }
  if (packed_bytes < dagl_ce_packed_weights_bytes()) {
    call_state().err = "packed-weights buffer too small";
    return DAGL_ERR_WORKSPACE;
  }
  int rc = launch_pack_fc_weights(w->fc1_w, w->fc1_b, w->fc2_w, w->fc2_b, packed, embed_tc_packed_weights_bytes(),
                                  static_cast<cudaStream_t>(stream));
  if (rc == 0)
    rc = launch_pack_feat_weights(w->in_channels, w->g_w, w->g_b, w->theta_w, w->theta_b, static_cast<char*>(packed) + embed_tc_packed_weights_bytes(),
                                  feature_maps_tc_packed_weights_bytes(), static_cast<cudaStream_t>(stream));
  return rc == -3 ? DAGL_ERR_WORKSPACE : rc;
}

// ---- ResBlock chains (common.py:59-79) ----------------------------------------------------------------------
size_t dagl_resblock_packed_weights_bytes(void) { return resblock_packed_weights_bytes(); }

int32_t dagl_resblock_pack_weights_f32(const DaglResBlockWeights* w, void* packed, size_t packed_bytes, void* stream) {
  call_state().launches = 0;
  if (!w || !w->conv1_w || !w->conv2_w || !packed) {
    call_state().err = "null pointer";
    return DAGL_ERR_INVALID_ARG;
  }
  const int rc = launch_pack_resblock_weights(w->conv1_w, w->conv1_b, w->conv2_w, w->conv2_b, packed, packed_bytes,
                                              static_cast<cudaStream_t>(stream));
  return rc == -3 ? DAGL_ERR_WORKSPACE : rc;
}

size_t dagl_resblocks_workspace_bytes(int32_t n_blocks, int32_t B, int32_t C, int32_t H, int32_t W) {
  if (n_blocks <= 0 || B <= 0 || C != 64 || H <= 0 || W <= 0) return 0;
  return resblocks_workspace_bytes(B, H, W, n_blocks);
}

int32_t dagl_resblocks_forward_f32(const DaglResBlockWeights* blocks, int32_t n_blocks, const float* x, float* y,
                                   int32_t B, int32_t C, int32_t H, int32_t W,
                                   void* workspace, size_t workspace_bytes, int32_t mode, void* stream) {
  call_state().launches = 0;
  call_state().impl = "none";
  if (!blocks || !x || !y || !workspace || n_blocks <= 0 || B <= 0 || H <= 0 || W <= 0) {
    call_state().err = "null pointer or non-positive size";
    return DAGL_ERR_INVALID_ARG;
  }
  if (C != 64 || n_blocks > 64 || mode < 0 || mode > 2) {
    call_state().err = "ResBlock chain: built for 64 channels, at most 64 blocks, mode 0 / 1 / 2";
    return DAGL_ERR_UNSUPPORTED;
  }
  if ((size_t)B * H * W >= ((size_t)1 << 31) / 64) {
    call_state().err = "ResBlock chain: tensor too large";
    return DAGL_ERR_UNSUPPORTED;
  }
  ResBlockParams rb[64];
  for (int i = 0; i < n_blocks; ++i) {
    const DaglResBlockWeights& w = blocks[i];
    if (!w.conv1_w || !w.conv2_w || !w.prelu_w || (w.prelu_n != 1 && w.prelu_n != 64)) {
      call_state().err = "ResBlock chain: null weight or PReLU parameter count not 1 / 64";
      return DAGL_ERR_INVALID_ARG;
    }
    rb[i] = ResBlockParams{w.conv1_w, w.conv1_b, w.prelu_w, w.prelu_n, w.conv2_w, w.conv2_b, w.res_scale, w.packed};
  }
  const int rc = launch_resblocks(B, H, W, x, y, n_blocks, rb, workspace, workspace_bytes, mode, static_cast<cudaStream_t>(stream));
  if (rc == 0) call_state().impl = mode == 0 ? "resblock_tc_pair" : mode == 1 ? "resblock_tc" : "resblock_tc_auto";
  return rc == -3 ? DAGL_ERR_WORKSPACE : rc;
}

size_t dagl_ce_host_staging_bytes(int32_t B, int32_t C, int32_t H, int32_t W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return align_up((size_t)B * C * H * W * sizeof(float)) + align_up((size_t)B * CI * H * W * sizeof(float));
}

int32_t dagl_ce_forward_host_f32(const DaglCEWeights* w, const float* b_host, float* y_host, int32_t B, int32_t H,
                                 int32_t W, void* workspace, size_t workspace_bytes, int32_t impl, void* stream) {
  int rc = check_weights(w);
  if (rc) return rc;
  rc = check_shape(B, H, W);
  if (rc) return rc;
  if (!b_host || !y_host || !workspace) {
    call_state().err = "null buffer";
    return DAGL_ERR_INVALID_ARG;
  }
  const int C = w->in_channels;
  const size_t need = dagl_ce_workspace_bytes_ex(B, C, H, W, impl, 0);     // staging buffers follow the forward's own workspace
  const size_t stage = dagl_ce_host_staging_bytes(B, C, H, W);
  if (workspace_bytes < need + stage) {
    call_state().err = "workspace too small for host entry (need workspace_bytes + host_staging_bytes)";
    return DAGL_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(workspace);
  float* b_dev = reinterpret_cast<float*>(base + need);
  const size_t b_bytes = (size_t)B * C * H * W * sizeof(float);
  const size_t y_bytes = (size_t)B * CI * H * W * sizeof(float);
  float* y_dev = reinterpret_cast<float*>(base + need + align_up(b_bytes));
  DAGL_CUDA_OK(cudaMemcpyAsync(b_dev, b_host, b_bytes, cudaMemcpyHostToDevice, st));
  rc = forward_impl(&w, 1, b_dev, y_dev, B, H, W, workspace, need, impl, st, nullptr, nullptr);
  if (rc) return rc;
  DAGL_CUDA_OK(cudaMemcpyAsync(y_host, y_dev, y_bytes, cudaMemcpyDeviceToHost, st));
  return 0;
}

int32_t dagl_ce_num_query_tiles(int32_t H, int32_t W) {
  if (H <= 0 || W <= 0) return 0;
  const Geom g = make_geom(1, 64, H, W);
  return (g.Nq + 127) / 128;
}

int32_t dagl_ce_forward_rows_f32(const DaglCEWeights* w, const float* b, float* rows, int32_t B, int32_t H, int32_t W,
                                 int32_t q_tile_begin, int32_t q_tile_end, void* workspace, size_t workspace_bytes,
                                 int32_t impl, void* stream) {
  if (q_tile_begin < 0 || q_tile_end <= q_tile_begin || q_tile_end > dagl_ce_num_query_tiles(H, W)) {
    call_state().err = "bad query-tile range";
    return DAGL_ERR_INVALID_ARG;
  }
  if (impl == DAGL_IMPL_SIMT) {
    call_state().err = "query-tile ranges / row output need a tensor-core impl (auto, tc, tc4)";
    return DAGL_ERR_UNSUPPORTED;
  }
  return forward_impl(&w, 1, b, nullptr, B, H, W, workspace, workspace_bytes, impl, static_cast<cudaStream_t>(stream),
                      nullptr, nullptr, rows, q_tile_begin, q_tile_end);
}

int32_t dagl_ce_fold_rows_f32(const float* rows, float* y, int32_t B, int32_t H, int32_t W, void* stream) {
  call_state().launches = 0;
  int rc = check_shape(B, H, W);
  if (rc) return rc;
  if (!rows || !y) {
    call_state().err = "null buffer";
    return DAGL_ERR_INVALID_ARG;
  }
  return launch_fold_rows(make_geom(B, 64, H, W), rows, y, /*shift_major=*/1, static_cast<cudaStream_t>(stream));
}

size_t dagl_graph_attend_workspace_bytes(int32_t B, int32_t H, int32_t W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  const Geom g = make_geom(B, 64, H, W);
  const size_t a_simt = attend_simt_workspace_bytes(g), a_tc = attend_tc_workspace_bytes(g);
  return a_simt > a_tc ? a_simt : a_tc;
}

int32_t dagl_graph_attend_f32(const float* Q, const float* K, const float* Kbar, const float* gamma,
                              const float* beta, const float* theta, float* y, int32_t B, int32_t H, int32_t W,
                              float softmax_scale, void* workspace, size_t workspace_bytes, int32_t impl,
                              void* stream, uint32_t* mask_bits, int32_t* nnz) {
  call_state().launches = 0;
  call_state().impl = "none";
  int rc = check_shape(B, H, W);
  if (rc) return rc;
  if (!Q || !K || !Kbar || !gamma || !beta || !theta || !y || !workspace) {
    call_state().err = "null buffer";
    return DAGL_ERR_INVALID_ARG;
  }
  const Geom g = make_geom(B, 64, H, W);
  AttendArgs a;
  a.Q = Q; a.K = K; a.Kbar = Kbar; a.gamma = gamma; a.beta = beta; a.theta = theta; a.y = y;
  a.scale = softmax_scale; a.mask_bits = mask_bits; a.nnz = nnz;
  a.ws = workspace; a.ws_bytes = workspace_bytes; a.kbar_out = nullptr;
  return run_attend(g, a, impl, nullptr, static_cast<cudaStream_t>(stream));
}

int32_t dagl_profile_enable(int32_t on) {
  CallState& s = call_state();
  if (on && !s.prof_created) {
    for (int i = 0; i < PROF_RING; ++i) {
      DAGL_CUDA_OK(cudaEventCreate(&s.prof_start[i]));
      DAGL_CUDA_OK(cudaEventCreate(&s.prof_stop[i]));
    }
    s.prof_created = true;
  }
  s.prof_on = on;
  s.prof_n = 0;
  return 0;
}

int32_t dagl_profile_read(float* ms, int32_t max) {
  CallState& s = call_state();
  if (!ms || max < 0) { s.err = "bad profile buffer"; return DAGL_ERR_INVALID_ARG; }
  int n = s.prof_n < max ? s.prof_n : max;
  for (int i = 0; i < n; ++i) {
    DAGL_CUDA_OK(cudaEventSynchronize(s.prof_stop[i]));
    if (s.prof_on == 2)       // time since the previous mark (mark 0: since the origin recorded with it, ~0)
      DAGL_CUDA_OK(cudaEventElapsedTime(&ms[i], i == 0 ? s.prof_start[0] : s.prof_stop[i - 1], s.prof_stop[i]));
    else
      DAGL_CUDA_OK(cudaEventElapsedTime(&ms[i], s.prof_start[i], s.prof_stop[i]));
  }
  s.prof_n = 0;
  return n;
}

const float* dagl_ce_workspace_view(void* workspace, int32_t which, int32_t B, int32_t C, int32_t H, int32_t W) {
  if (!workspace || B <= 0 || C <= 0 || H <= 0 || W <= 0) return nullptr;
  const WsLayout L = ws_layout(make_geom(B, C, H, W));
  const char* base = static_cast<const char*>(workspace);
  switch (which) {
    case 0: return reinterpret_cast<const float*>(base + L.G);
    case 1: return reinterpret_cast<const float*>(base + L.Th);
    case 2: return reinterpret_cast<const float*>(base + L.gamma);
    case 3: return reinterpret_cast<const float*>(base + L.beta);
    case 4: return reinterpret_cast<const float*>(base + L.Q);
    case 5: return reinterpret_cast<const float*>(base + L.K);
    case 6: return reinterpret_cast<const float*>(base + L.Kbar);
    default: return nullptr;
  }
}

}  // extern "C"
