// Expert graph: the small GroupNorm ResBlock UNet.
// reference: mnist/models/unet_small.py:47-92, shapes/models/unet_small.py:53-120.
//
// Parameters arrive by state_dict key; cdm_unet_finalize() packs them once for both precisions:
//   conv weights   OIHW fp32 -> [Ktot][Cout] fp32 (CUDA-core path) and [Cout][Ktot] fp16 (tcgen05 path),
//                  K ordered tap-major / channel-minor, the 1x1 res_conv appended as extra K rows
//   time MLPs      transposed to [in][out]; the five per-block Linear(256, Cout) are concatenated into
//                  one [256][640] matrix whose bias also carries each block's conv1 bias
// The forward is a fixed launch sequence per micro-batch (see forward_chunk); activations are NHWC
// in a caller-provided workspace.
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "group.cuh"
#include "layers.cuh"

namespace cdm {

struct ParamSpec { std::string key; std::vector<int64_t> shape; int64_t numel; };

struct BlockW {
  int cin, cout;
  bool has_res;
  float *g1, *b1, *g2, *b2;           // GroupNorm affine
  float *w1_f32, *w2_f32;             // [Ktot][Cout]
  h16 *w1_h16, *w2_h16;   // [Cout][Ktot], tap-major K (conv_tc.cu)
  h16 *w1_halo, *w2_halo;   // [Cout][Ktot], chunk-major K (conv_tc2.cu)
  h16 *w1_stack, *w2_stack; // [192][3*Cin (+Cres)], dx taps stacked along N (conv_tc3.cu; Cout = 64 blocks only)
  h16 *w1_x3h, *w1_x3l, *w2_x3h, *w2_x3l;   // [Cout][Ktot] hi / lo planes of w * 2^8 (conv_x3.cu, CDM_PREC_F16X3)
  float* bias2;                        // [Cout] conv2 bias (+ res_conv bias)
  int bias_off;                        // column of this block in block_bias
};

}  // namespace cdm

using namespace cdm;

struct cdm_unet {
  cdm_unet_config cfg;
  int device = 0;
  int num_sms = 148;
  std::vector<ParamSpec> specs;
  std::map<std::string, std::vector<float>> host;
  bool finalized = false;
  std::vector<void*> allocs;
  // packed
  TembWeights temb{};
  float *init_w = nullptr, *init_b = nullptr, *out_w = nullptr, *out_b = nullptr, *zeros = nullptr;
  BlockW blk[5];
  int nb_total = 0;
  // last-forward bookkeeping for debug reads
  void* last_ws = nullptr;
  int last_prec = -1, last_B = 0, last_S = 0;
};

namespace cdm {

static void add_spec(cdm_unet* m, const std::string& key, std::vector<int64_t> shape) {
  int64_t n = 1;
  for (auto s : shape) n *= s;
  m->specs.push_back({key, shape, n});
}
static void add_block_spec(cdm_unet* m, const std::string& p, int cin, int cout, int td) {
  add_spec(m, p + ".block1.0.weight", {cin});
  add_spec(m, p + ".block1.0.bias", {cin});
  add_spec(m, p + ".block1.2.weight", {cout, cin, 3, 3});
  add_spec(m, p + ".block1.2.bias", {cout});
  add_spec(m, p + ".time_mlp.1.weight", {cout, td});
  add_spec(m, p + ".time_mlp.1.bias", {cout});
  add_spec(m, p + ".block2.0.weight", {cout});
  add_spec(m, p + ".block2.0.bias", {cout});
  add_spec(m, p + ".block2.3.weight", {cout, cout, 3, 3});
  add_spec(m, p + ".block2.3.bias", {cout});
  if (cin != cout) {
    add_spec(m, p + ".res_conv.weight", {cout, cin, 1, 1});
    add_spec(m, p + ".res_conv.bias", {cout});
  }
}

static const char* BLOCK_NAMES[5] = {"down1", "down2", "bot1", "up1", "up2"};

template <typename T> static int upload(cdm_unet* m, const std::vector<T>& h, T** dptr) {
  void* d = nullptr;
  CDM_CUDA_OK(cudaMalloc(&d, h.size() * sizeof(T) + 16));
  m->allocs.push_back(d);
  CDM_CUDA_OK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dptr = (T*)d;
  return CDM_OK;
}

// OIHW (+ optional [O][Cres] 1x1) -> K-major packs.  k = tap*Cin + ci, then 9*Cin + cr.
void pack_conv(const std::vector<float>& w, int cout, int cin, int taps, const std::vector<float>* wres, int cres,
                      std::vector<float>& kn, std::vector<h16>& nk) {
  const int ktot = taps * cin + (wres ? cres : 0);
  kn.assign((size_t)ktot * cout, 0.f);
  nk.assign((size_t)cout * ktot, f_to_h16(0.f));
  for (int o = 0; o < cout; ++o) {
    for (int ci = 0; ci < cin; ++ci)
      for (int tap = 0; tap < taps; ++tap) {
        const float v = w[((size_t)o * cin + ci) * taps + tap];
        const int k = tap * cin + ci;
        kn[(size_t)k * cout + o] = v;
        nk[(size_t)o * ktot + k] = f_to_h16(v);
      }
    if (wres)
      for (int cr = 0; cr < cres; ++cr) {
        const float v = (*wres)[(size_t)o * cres + cr];
        const int k = taps * cin + cr;
        kn[(size_t)k * cout + o] = v;
        nk[(size_t)o * ktot + k] = f_to_h16(v);
      }
  }
}

static std::vector<float> transpose(const std::vector<float>& w, int rows, int cols) {  // [rows][cols] -> [cols][rows]
  std::vector<float> t((size_t)rows * cols);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = w[(size_t)r * cols + c];
  return t;
}

// ---- workspace plan ------------------------------------------------------------------------------
struct Plan {
  size_t block_bias, stats, x0, h, y, d1, p1, d2, p2, b1, cat1, u1, cat2, u2, xs, total;
  int chunk;
};
static int g_microbatch = -1;
static int microbatch() {
  if (g_microbatch < 0) {
    const char* e = getenv("CDM_MICROBATCH");
    g_microbatch = e ? atoi(e) : 4096;
    if (g_microbatch < 1) g_microbatch = 4096;
  }
  return g_microbatch;
}
static Plan make_plan(const cdm_unet* m, int B, int S, int prec) {
  Plan p{};
  const size_t es = (prec == CDM_PREC_F16) ? 2 : 4;
  const int d = m->cfg.base_dim;
  p.chunk = B < microbatch() ? B : microbatch();
  const size_t n = (size_t)p.chunk, s2 = (size_t)S * S, s4 = s2 / 4, s16 = s2 / 16;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  p.block_bias = take(((size_t)B * m->nb_total + m->cfg.time_emb_dim) * 4);   // + [TD] scratch of launch_temb_row
  p.stats = take(12 * n * GN_GROUPS * 2 * sizeof(stat_t));
  p.x0 = take(n * s2 * d * es);
  p.h = take(n * s2 * 3 * d * es);
  p.y = take(n * s2 * d * es);
  p.d1 = take(n * s2 * d * es);
  p.p1 = take(n * s4 * d * es);
  p.d2 = take(n * s4 * 2 * d * es);
  p.p2 = take(n * s16 * 2 * d * es);
  p.b1 = take(n * s16 * 4 * d * es);
  p.cat1 = take(n * s4 * 6 * d * es);
  p.u1 = take(n * s4 * 2 * d * es);
  p.cat2 = take(n * s2 * 3 * d * es);
  p.u2 = take(n * s2 * d * es);
  p.xs = (prec == CDM_PREC_F16X3) ? take(n * s2 * 3 * d * 4) : 0;   // hi / lo planes of a block input (the folded res_conv's operand)
  p.total = off;
  return p;
}

#ifdef CDM_INSTRUMENT
extern int g_conv_timing;
#endif
static int g_conv_halo = -1, g_fuse_gn = -1, g_conv_stack = -1;
static int g_fuse_proj = -1;
static int g_grouped = -1;
static bool grouped_enabled() {
  if (g_grouped < 0) { const char* e = getenv("CDM_GROUPED"); g_grouped = e ? atoi(e) : 1; }
  return g_grouped != 0;
}
static bool fuse_proj_enabled() {
  if (g_fuse_proj < 0) { const char* e = getenv("CDM_FUSE_PROJ"); g_fuse_proj = e ? atoi(e) : 1; }
  return g_fuse_proj != 0;
}
static int stack_mode() {
  if (g_conv_stack < 0) { const char* e = getenv("CDM_CONV_STACK"); g_conv_stack = e ? atoi(e) : 2; }
  return g_conv_stack;
}
static bool halo_enabled() {
  if (g_conv_halo < 0) { const char* e = getenv("CDM_CONV_HALO"); g_conv_halo = e ? atoi(e) : 1; }
  return g_conv_halo != 0;
}
static bool fuse_gn_enabled() {
  if (g_fuse_gn < 0) { const char* e = getenv("CDM_FUSE_GN"); g_fuse_gn = e ? atoi(e) : 1; }
  return g_fuse_gn != 0;
}
template <typename T> struct PrecTraits;
template <> struct PrecTraits<float> {
  static bool can_fuse_gn(int, int, int, int, int) { return false; }
  static bool can_virtual_concat(int, int, int, int, int) { return false; }
  static bool can_fuse_proj(const ConvArgs<float>&, const BlockW&) { return false; }
  static int conv(const cdm_unet*, const ConvArgs<float>& c, const BlockW& b, int which, cudaStream_t st) {
    return launch_conv_fp32(c, which == 1 ? b.w1_f32 : b.w2_f32, st);
  }
};
template <> struct PrecTraits<h16> {
  static bool can_fuse_gn(int H, int W, int Cin, int Cres, int Cout) {
    return halo_enabled() && fuse_gn_enabled() && conv_halo_supported(H, W, Cin, Cres, Cout, 9);
  }
  // up blocks: conv1 (K = Ca + Cs) and conv2's folded res_conv read cat([upsampled, skip]) from the two tensors in place
  static bool can_virtual_concat(int H, int W, int Ca, int Cs, int Cout) {
    static const int env = [] { const char* e = getenv("CDM_VIRTUAL_CONCAT"); return e ? atoi(e) : 1; }();
    return env && can_fuse_gn(H, W, Ca + Cs, 0, Cout) && can_fuse_gn(H, W, Cout, Ca + Cs, Cout) && upcat_virtual_supported(Ca, Cs);
  }
  // the fused out_conv lives in the epilogues of the stacked kernel and of the halo kernel's PROJ instance (BN = 64)
  static bool can_fuse_proj(const ConvArgs<h16>& c, const BlockW& b) {
    const int sm = stack_mode();
    if (!fuse_proj_enabled() || !halo_enabled()) return false;
    if ((sm == 1 || (sm == 2 && (c.r || c.Cin > 64))) && b.w2_stack && conv_stack3_supported(c.H, c.W, c.Cin, c.r ? c.Cres : 0, c.Cout, c.taps))
      return true;                                                    // the stacked kernel takes the layer
    return c.Cout == 64 && conv_halo_supported(c.H, c.W, c.Cin, c.r ? c.Cres : 0, c.Cout, c.taps);   // else the halo kernel's PROJ instance
  }
  static int conv(const cdm_unet* m, const ConvArgs<h16>& c, const BlockW& b, int which, cudaStream_t st) {
    const h16* ws = which == 1 ? b.w1_stack : b.w2_stack;
    // conv_stack: 0 = never, 1 = every supported layer, 2 (default) = layers with a folded res_conv or more than one
    // 64-channel K chunk (28x28 64+192->64: 0.56 ms vs 0.71 ms; 192->64: 0.89 vs 0.95 ms).  Single-chunk 64->64 layers
    // are bound by the CUDA-core work of prologue + epilogue in either kernel, and the halo kernel has less of it (no
    // +-1 row shuffles): 0.43 vs 0.47 ms
    const int sm = stack_mode();
    if (halo_enabled() && (sm == 1 || (sm == 2 && (c.r || c.Cin > 64))) && ws && conv_stack3_supported(c.H, c.W, c.Cin, c.r ? c.Cres : 0, c.Cout, c.taps))
      return launch_conv_stack3(c, ws, m->num_sms, st);
    if (halo_enabled() && conv_halo_supported(c.H, c.W, c.Cin, c.r ? c.Cres : 0, c.Cout, c.taps))
      return launch_conv_halo(c, which == 1 ? b.w1_halo : b.w2_halo, m->num_sms, st);
    return launch_conv_tc(c, which == 1 ? b.w1_h16 : b.w2_h16, m->num_sms, st);
  }
};

// ResBlock: GN -> SiLU -> conv3x3 (+temb) -> GN -> SiLU -> conv3x3 + (res_conv(x) | x)
// reference: mnist/models/unet_small.py:39-44
// `proj` (optional): the UNet's out_conv; *proj_done reports whether conv2's epilogue computed it (then `out` is NOT written)
struct OutProj { const float* w; const float* b; float* out; int c; };
// `xin2` (optional): the block input is the virtual concat cat([xin (cin1 channels), xin2]) -- see ConvArgs::a2.
// `st_out` (optional): GroupNorm {sum, sumsq} of the block output (a later virtual concat needs them for its skip part).
template <typename T>
static int resblock(const cdm_unet* m, const BlockW& bw, const T* xin, const stat_t* st_in, stat_t* st_mid, T* h, T* y,
                    T* out, const float* block_bias, int bias_stride, int n, int H, int W, cudaStream_t st,
                    const OutProj* proj = nullptr, bool* proj_done = nullptr, const T* xin2 = nullptr, int cin1 = 0,
                    stat_t* st_out = nullptr) {
  using P = PrecTraits<T>;
  // GroupNorm+SiLU runs inside the conv (on the halo tile in shared memory) when the halo kernel takes the layer
  const bool fuse1 = P::can_fuse_gn(H, W, bw.cin, 0, bw.cout);
  const bool fuse2 = P::can_fuse_gn(H, W, bw.cout, bw.has_res ? bw.cin : 0, bw.cout);
  ConvArgs<T> c1{};
  if (xin2 && !(fuse1 && bw.has_res)) return fail(CDM_ERR_INVALID, "resblock: a virtual concat needs the fused GroupNorm prologue and a res_conv");
  if (fuse1) { c1.a = xin; c1.a2 = xin2; c1.a_split = cin1; c1.gn_stats = st_in; c1.gn_gamma = bw.g1; c1.gn_beta = bw.b1; }
  else { CDM_TRY(launch_gn_silu<T>(xin, st_in, bw.g1, bw.b1, h, n, H * W, bw.cin, st)); c1.a = h; }
  c1.out = y; c1.bias = block_bias + bw.bias_off; c1.bias_stride = bias_stride; c1.stats = st_mid;
  c1.B = n; c1.H = H; c1.W = W; c1.Cin = bw.cin; c1.Cout = bw.cout; c1.taps = 9;
  CDM_TRY(P::conv(m, c1, bw, 1, st));
  ConvArgs<T> c2{};
  if (fuse2) { c2.a = y; c2.gn_stats = st_mid; c2.gn_gamma = bw.g2; c2.gn_beta = bw.b2; }
  else { CDM_TRY(launch_gn_silu<T>(y, st_mid, bw.g2, bw.b2, h, n, H * W, bw.cout, st)); c2.a = h; }
  c2.out = out; c2.bias = bw.bias2; c2.bias_stride = 0; c2.stats = st_out;
  c2.B = n; c2.H = H; c2.W = W; c2.Cin = bw.cout; c2.Cout = bw.cout; c2.taps = 9;
  if (bw.has_res) { c2.r = xin; c2.Cres = bw.cin; c2.r2 = xin2; c2.r_split = cin1; } else { c2.identity = xin; }
  if (proj_done) *proj_done = false;
  if (proj && P::can_fuse_proj(c2, bw)) {
    c2.proj_w = proj->w; c2.proj_b = proj->b; c2.proj_out = proj->out; c2.proj_c = proj->c;
    *proj_done = true;
  }
  CDM_TRY(P::conv(m, c2, bw, 2, st));
  return CDM_OK;
}

// The same ResBlock on the fp32-class tensor-core path (CDM_PREC_F16X3): activations stay fp32 in HBM; the GroupNorm+SiLU
// pass writes its result as hi / lo fp16 planes (into `hbuf`, and the raw block input into `xsbuf` for a folded res_conv),
// and both convs run on tcgen05 with three MMAs per K step (conv_x3.cu).
static int resblock_x3(const cdm_unet* m, const BlockW& bw, const float* xin, const stat_t* st_in, stat_t* st_mid, float* hbuf,
                       float* xsbuf, float* y, float* out, const float* block_bias, int bias_stride, int n, int H, int W,
                       cudaStream_t st) {
  const size_t ne_in = (size_t)n * H * W * bw.cin, ne_mid = (size_t)n * H * W * bw.cout;
  const X3Planes none{nullptr, nullptr};
  X3Planes hp{reinterpret_cast<h16*>(hbuf), reinterpret_cast<h16*>(hbuf) + ne_in};
  X3Planes rp = bw.has_res ? X3Planes{reinterpret_cast<h16*>(xsbuf), reinterpret_cast<h16*>(xsbuf) + ne_in} : none;
  CDM_TRY(launch_gn_silu_split(xin, st_in, bw.g1, bw.b1, hp, rp, n, H * W, bw.cin, st));
  ConvArgs<float> c1{};
  c1.out = y; c1.bias = block_bias + bw.bias_off; c1.bias_stride = bias_stride; c1.stats = st_mid;
  c1.B = n; c1.H = H; c1.W = W; c1.Cin = bw.cin; c1.Cout = bw.cout; c1.taps = 9;
  CDM_TRY(launch_conv_x3(c1, hp, none, bw.w1_x3h, bw.w1_x3l, m->num_sms, st));
  X3Planes hp2{reinterpret_cast<h16*>(hbuf), reinterpret_cast<h16*>(hbuf) + ne_mid};
  CDM_TRY(launch_gn_silu_split(y, st_mid, bw.g2, bw.b2, hp2, none, n, H * W, bw.cout, st));
  ConvArgs<float> c2{};
  c2.out = out; c2.bias = bw.bias2; c2.bias_stride = 0; c2.stats = nullptr;
  c2.B = n; c2.H = H; c2.W = W; c2.Cin = bw.cout; c2.Cout = bw.cout; c2.taps = 9;
  if (bw.has_res) { c2.r = xin; c2.Cres = bw.cin; } else { c2.identity = xin; }
  CDM_TRY(launch_conv_x3(c2, hp2, rp, bw.w2_x3h, bw.w2_x3l, m->num_sms, st));
  return CDM_OK;
}

template <typename T>
static int forward_chunk(const cdm_unet* m, const Plan& pl, uint8_t* ws, const float* x, float* eps, const float* bias,
                         int bias_stride, int n, int S, cudaStream_t st, bool x3 = false) {
  const int d = m->cfg.base_dim, cin = m->cfg.in_channels;
  stat_t* stats = reinterpret_cast<stat_t*>(ws + pl.stats);
  const size_t ss = (size_t)n * GN_GROUPS * 2;
  auto stat = [&](int i) { return stats + ss * i; };
  auto buf = [&](size_t off) { return reinterpret_cast<T*>(ws + off); };
  T *x0 = buf(pl.x0), *h = buf(pl.h), *y = buf(pl.y), *d1 = buf(pl.d1), *p1 = buf(pl.p1), *d2 = buf(pl.d2),
    *p2 = buf(pl.p2), *b1 = buf(pl.b1), *cat1 = buf(pl.cat1), *u1 = buf(pl.u1), *cat2 = buf(pl.cat2), *u2 = buf(pl.u2);
  const int S2 = S / 2, S4 = S / 4;

  using P = PrecTraits<T>;
  // virtual concat (fp16 halo / stacked kernels): the skip tensors d2 / d1 are NOT copied next to the upsampled tensor;
  // cat1 / cat2 then hold only the upsampled channels and the convs of the up blocks read both tensors in place.  The
  // skip's share of the concat's GroupNorm statistics comes from the statistics of d2 / d1 (slots 11 / 10), which the
  // max-pool kernels accumulate while they read those tensors anyway.
  const bool v1 = P::can_virtual_concat(S2, S2, 4 * d, 2 * d, 2 * d), v2 = P::can_virtual_concat(S, S, 2 * d, d, d);
  CDM_CUDA_OK(cudaMemsetAsync(stats, 0, ss * 12 * sizeof(stat_t), st));
  CDM_TRY(launch_init_conv<T>(x, m->init_w, m->init_b, x0, stat(0), n, cin, S, S, d, st));
  // one ResBlock: the fp32 workspace layout serves both the CUDA-core path and the fp32-class tensor-core path (x3)
  auto RB = [&](int i, const T* xin, const stat_t* st_in, stat_t* st_mid, T* out, int H, const OutProj* proj = nullptr,
                bool* pdone = nullptr, const T* xin2 = nullptr, int cin1 = 0) -> int {
    if constexpr (std::is_same<T, float>::value) {
      if (x3) {
        if (pdone) *pdone = false;
        return resblock_x3(m, m->blk[i], xin, st_in, st_mid, h, reinterpret_cast<float*>(ws + pl.xs), y, out, bias, bias_stride, n, H, H, st);
      }
    }
    return resblock<T>(m, m->blk[i], xin, st_in, st_mid, h, y, out, bias, bias_stride, n, H, H, st, proj, pdone, xin2, cin1);
  };
  CDM_TRY(RB(0, x0, stat(0), stat(1), d1, S));
  CDM_TRY(launch_maxpool_stats<T>(d1, p1, stat(2), n, S, S, d, st, v2 ? stat(10) : nullptr));       // + statistics of d1
  CDM_TRY(RB(1, p1, stat(2), stat(3), d2, S2));
  CDM_TRY(launch_maxpool_stats<T>(d2, p2, stat(4), n, S2, S2, 2 * d, st, v1 ? stat(11) : nullptr));   // + statistics of d2
  CDM_TRY(RB(2, p2, stat(4), stat(5), b1, S4));
  CDM_TRY(launch_upcat_stats<T>(b1, d2, cat1, stat(6), n, S4, S4, 4 * d, 2 * d, st, v1 ? stat(11) : nullptr));
  CDM_TRY(RB(3, cat1, stat(6), stat(7), u1, S2, nullptr, nullptr, v1 ? d2 : nullptr, v1 ? 4 * d : 0));
  CDM_TRY(launch_upcat_stats<T>(u1, d1, cat2, stat(8), n, S2, S2, 2 * d, d, st, v2 ? stat(10) : nullptr));
  const OutProj proj{m->out_w, m->out_b, eps, cin};
  bool proj_done = false;
  CDM_TRY(RB(4, cat2, stat(8), stat(9), u2, S, &proj, &proj_done, v2 ? d1 : nullptr, v2 ? 2 * d : 0));
  if (!proj_done) CDM_TRY(launch_out_conv<T>(u2, m->out_w, m->out_b, eps, n, S * S, d, cin, st));
  return CDM_OK;
}

// ---- grouped forward: the same layer of K experts in ONE launch per convolution (group.cuh) ------------------------------
// The K experts advance in lockstep, each in its own workspace.  Every convolution of the fp16 graph (halo / stacked kernels)
// is recorded per expert and issued as one grouped launch; the elementwise layers and the shifted-box kernel of the smallest
// maps run per expert.  Results are bit-identical to K separate forwards: the same kernels body, parameters and tile order.
static int resblock_group(cdm_unet* const* ms, int K, int bi, const h16* const* xin, const stat_t* const* st_in, stat_t* const* st_mid,
                          h16* const* h, h16* const* y, h16* const* out, const float* const* block_bias, int bias_stride, int n, int H,
                          cudaStream_t st, const OutProj* proj, bool* proj_done, const h16* const* xin2, int cin1) {
  using P = PrecTraits<h16>;
  const BlockW& b0 = ms[0]->blk[bi];
  const bool fuse1 = P::can_fuse_gn(H, H, b0.cin, 0, b0.cout);
  const bool fuse2 = P::can_fuse_gn(H, H, b0.cout, b0.has_res ? b0.cin : 0, b0.cout);
  if (xin2 && !(fuse1 && b0.has_res)) return fail(CDM_ERR_INVALID, "resblock: a virtual concat needs the fused GroupNorm prologue and a res_conv");
  const int sms = ms[0]->num_sms;
  if (!fuse1)
    for (int k = 0; k < K; ++k) CDM_TRY(launch_gn_silu<h16>(xin[k], st_in[k], ms[k]->blk[bi].g1, ms[k]->blk[bi].b1, h[k], n, H * H, b0.cin, st));
  group_begin();
  for (int k = 0; k < K; ++k) {
    const BlockW& bw = ms[k]->blk[bi];
    ConvArgs<h16> c1{};
    if (fuse1) { c1.a = xin[k]; c1.a2 = xin2 ? xin2[k] : nullptr; c1.a_split = cin1; c1.gn_stats = st_in[k]; c1.gn_gamma = bw.g1; c1.gn_beta = bw.b1; }
    else c1.a = h[k];
    c1.out = y[k]; c1.bias = block_bias[k] + bw.bias_off; c1.bias_stride = bias_stride; c1.stats = st_mid[k];
    c1.B = n; c1.H = H; c1.W = H; c1.Cin = bw.cin; c1.Cout = bw.cout; c1.taps = 9;
    const int rc = P::conv(ms[k], c1, bw, 1, st);
    if (rc != CDM_OK) { group_state().recording = false; return rc; }
  }
  CDM_TRY(group_flush(sms, st));
  if (!fuse2)
    for (int k = 0; k < K; ++k) CDM_TRY(launch_gn_silu<h16>(y[k], st_mid[k], ms[k]->blk[bi].g2, ms[k]->blk[bi].b2, h[k], n, H * H, b0.cout, st));
  if (proj_done) *proj_done = false;
  bool pd = false;
  group_begin();
  for (int k = 0; k < K; ++k) {
    const BlockW& bw = ms[k]->blk[bi];
    ConvArgs<h16> c2{};
    if (fuse2) { c2.a = y[k]; c2.gn_stats = st_mid[k]; c2.gn_gamma = bw.g2; c2.gn_beta = bw.b2; }
    else c2.a = h[k];
    c2.out = out[k]; c2.bias = bw.bias2; c2.bias_stride = 0; c2.stats = nullptr;
    c2.B = n; c2.H = H; c2.W = H; c2.Cin = bw.cout; c2.Cout = bw.cout; c2.taps = 9;
    if (bw.has_res) { c2.r = xin[k]; c2.Cres = bw.cin; c2.r2 = xin2 ? xin2[k] : nullptr; c2.r_split = cin1; } else { c2.identity = xin[k]; }
    if (proj && P::can_fuse_proj(c2, bw)) {
      c2.proj_w = proj[k].w; c2.proj_b = proj[k].b; c2.proj_out = proj[k].out; c2.proj_c = proj[k].c;
      pd = true;
    }
    const int rc = P::conv(ms[k], c2, bw, 2, st);
    if (rc != CDM_OK) { group_state().recording = false; return rc; }
  }
  CDM_TRY(group_flush(sms, st));
  if (proj_done) *proj_done = pd;
  return CDM_OK;
}

static int forward_chunk_group(cdm_unet* const* ms, int K, const Plan& pl, uint8_t* const* wss, const float* const* xs, float* const* epss,
                               const float* const* biases, int bias_stride, int n, int S, cudaStream_t st) {
  using P = PrecTraits<h16>;
  const int d = ms[0]->cfg.base_dim, S2 = S / 2, S4 = S / 4;
  const size_t ss = (size_t)n * GN_GROUPS * 2;
  const bool v1 = P::can_virtual_concat(S2, S2, 4 * d, 2 * d, 2 * d), v2 = P::can_virtual_concat(S, S, 2 * d, d, d);
  // per-expert views of the workspace
  stat_t* stats[GROUP_MAX];
  h16 *x0[GROUP_MAX], *h[GROUP_MAX], *y[GROUP_MAX], *d1[GROUP_MAX], *p1[GROUP_MAX], *d2[GROUP_MAX], *p2[GROUP_MAX], *b1[GROUP_MAX],
      *cat1[GROUP_MAX], *u1[GROUP_MAX], *cat2[GROUP_MAX], *u2[GROUP_MAX];
  for (int k = 0; k < K; ++k) {
    uint8_t* ws = wss[k];
    auto buf = [&](size_t off) { return reinterpret_cast<h16*>(ws + off); };
    stats[k] = reinterpret_cast<stat_t*>(ws + pl.stats);
    x0[k] = buf(pl.x0); h[k] = buf(pl.h); y[k] = buf(pl.y); d1[k] = buf(pl.d1); p1[k] = buf(pl.p1); d2[k] = buf(pl.d2); p2[k] = buf(pl.p2);
    b1[k] = buf(pl.b1); cat1[k] = buf(pl.cat1); u1[k] = buf(pl.u1); cat2[k] = buf(pl.cat2); u2[k] = buf(pl.u2);
    CDM_CUDA_OK(cudaMemsetAsync(stats[k], 0, ss * 12 * sizeof(stat_t), st));
    CDM_TRY(launch_init_conv<h16>(xs[k], ms[k]->init_w, ms[k]->init_b, x0[k], stats[k], n, ms[k]->cfg.in_channels, S, S, d, st));
  }
  auto stat = [&](int i, const stat_t* (&in)[GROUP_MAX]) { for (int k = 0; k < K; ++k) in[k] = stats[k] + ss * i; };
  auto statm = [&](int i, stat_t* (&out)[GROUP_MAX]) { for (int k = 0; k < K; ++k) out[k] = stats[k] + ss * i; };
  const stat_t* sin[GROUP_MAX];
  stat_t* smid[GROUP_MAX];
  auto RB = [&](int bi, h16* const* xin, int si, h16* const* out, int H, const OutProj* proj = nullptr, bool* pdone = nullptr,
                h16* const* xin2 = nullptr, int cin1 = 0) -> int {
    stat(si, sin); statm(si + 1, smid);
    return resblock_group(ms, K, bi, xin, sin, smid, h, y, out, biases, bias_stride, n, H, st, proj, pdone, xin2, cin1);
  };
  CDM_TRY(RB(0, x0, 0, d1, S));
  for (int k = 0; k < K; ++k) CDM_TRY(launch_maxpool_stats<h16>(d1[k], p1[k], stats[k] + ss * 2, n, S, S, d, st, v2 ? stats[k] + ss * 10 : nullptr));
  CDM_TRY(RB(1, p1, 2, d2, S2));
  for (int k = 0; k < K; ++k) CDM_TRY(launch_maxpool_stats<h16>(d2[k], p2[k], stats[k] + ss * 4, n, S2, S2, 2 * d, st, v1 ? stats[k] + ss * 11 : nullptr));
  CDM_TRY(RB(2, p2, 4, b1, S4));
  for (int k = 0; k < K; ++k) CDM_TRY(launch_upcat_stats<h16>(b1[k], d2[k], cat1[k], stats[k] + ss * 6, n, S4, S4, 4 * d, 2 * d, st, v1 ? stats[k] + ss * 11 : nullptr));
  CDM_TRY(RB(3, cat1, 6, u1, S2, nullptr, nullptr, v1 ? d2 : nullptr, v1 ? 4 * d : 0));
  for (int k = 0; k < K; ++k) CDM_TRY(launch_upcat_stats<h16>(u1[k], d1[k], cat2[k], stats[k] + ss * 8, n, S2, S2, 2 * d, d, st, v2 ? stats[k] + ss * 10 : nullptr));
  OutProj proj[GROUP_MAX];
  for (int k = 0; k < K; ++k) proj[k] = OutProj{ms[k]->out_w, ms[k]->out_b, epss[k], ms[k]->cfg.in_channels};
  bool proj_done = false;
  CDM_TRY(RB(4, cat2, 8, u2, S, proj, &proj_done, v2 ? d1 : nullptr, v2 ? 2 * d : 0));
  if (!proj_done)
    for (int k = 0; k < K; ++k) CDM_TRY(launch_out_conv<h16>(u2[k], ms[k]->out_w, ms[k]->out_b, epss[k], n, S * S, d, ms[k]->cfg.in_channels, st));
  return CDM_OK;
}

// ---- forward-mode (primal + tangent) graph -----------------------------------------------------------
// Tangent twins of every activation live in a second workspace region at byte offset pl.total.  T = float: CUDA-core
// convs (parity path).  T = h16: every conv (primal and tangent) on the tensor cores; GroupNorm+SiLU and its tangent
// are one elementwise pass (the fused conv prologue cannot carry a tangent), statistics stay fp32.
template <typename T>
static int resblock_jvp(const cdm_unet* m, const BlockW& bw, const T* xin, const T* dxin, const stat_t* st_in,
                        stat_t* st_mid, stat_t* stt_in, stat_t* stt_mid, T* h, T* dh, T* y, T* dy,
                        T* out, T* dout, const float* block_bias, int n, int H, int W, cudaStream_t st) {
  using P = PrecTraits<T>;
  const int HW = H * W;
  CDM_TRY(launch_pair_stats<T>(xin, dxin, stt_in, n, HW, bw.cin, st));
  CDM_TRY(launch_gn_silu_jvp<T>(xin, dxin, st_in, stt_in, bw.g1, bw.b1, h, dh, n, HW, bw.cin, st));
  ConvArgs<T> c1{};
  c1.a = h; c1.out = y; c1.bias = block_bias + bw.bias_off; c1.bias_stride = m->nb_total; c1.stats = st_mid;
  c1.B = n; c1.H = H; c1.W = W; c1.Cin = bw.cin; c1.Cout = bw.cout; c1.taps = 9;
  CDM_TRY(P::conv(m, c1, bw, 1, st));
  ConvArgs<T> t1 = c1;
  t1.a = dh; t1.out = dy; t1.bias = m->zeros; t1.bias_stride = 0; t1.stats = nullptr;
  CDM_TRY(P::conv(m, t1, bw, 1, st));
  CDM_TRY(launch_pair_stats<T>(y, dy, stt_mid, n, HW, bw.cout, st));
  CDM_TRY(launch_gn_silu_jvp<T>(y, dy, st_mid, stt_mid, bw.g2, bw.b2, h, dh, n, HW, bw.cout, st));
  ConvArgs<T> c2{};
  c2.a = h; c2.out = out; c2.bias = bw.bias2; c2.bias_stride = 0; c2.stats = nullptr;
  c2.B = n; c2.H = H; c2.W = W; c2.Cin = bw.cout; c2.Cout = bw.cout; c2.taps = 9;
  ConvArgs<T> t2 = c2;
  t2.a = dh; t2.out = dout; t2.bias = m->zeros;
  if (bw.has_res) { c2.r = xin; c2.Cres = bw.cin; t2.r = dxin; t2.Cres = bw.cin; }
  else { c2.identity = xin; t2.identity = dxin; }
  CDM_TRY(P::conv(m, c2, bw, 2, st));
  CDM_TRY(P::conv(m, t2, bw, 2, st));
  return CDM_OK;
}

template <typename T>
static int forward_jvp_chunk(const cdm_unet* m, const Plan& pl, uint8_t* ws, const float* x, const float* v_in,
                             const float* v_out, float* eps, float* deps, float* vjv, const float* bias, int n, int S,
                             cudaStream_t st) {
  const int d = m->cfg.base_dim, cin = m->cfg.in_channels;
  uint8_t* wt = ws + pl.total;   // tangent region
  stat_t* stats = reinterpret_cast<stat_t*>(ws + pl.stats);
  stat_t* stats_t = reinterpret_cast<stat_t*>(wt + pl.stats);
  const size_t ss = (size_t)n * GN_GROUPS * 2;
  auto stat = [&](int i) { return stats + ss * i; };
  auto statt = [&](int i) { return stats_t + ss * i; };
  auto P = [&](size_t off) { return reinterpret_cast<T*>(ws + off); };
  auto D = [&](size_t off) { return reinterpret_cast<T*>(wt + off); };
  const int S2 = S / 2, S4 = S / 4;
  CDM_CUDA_OK(cudaMemsetAsync(stats, 0, ss * 10 * sizeof(stat_t), st));
  CDM_CUDA_OK(cudaMemsetAsync(stats_t, 0, ss * 10 * sizeof(stat_t), st));
  CDM_TRY(launch_init_conv<T>(x, m->init_w, m->init_b, P(pl.x0), stat(0), n, cin, S, S, d, st));
  CDM_TRY(launch_init_conv<T>(v_in, m->init_w, m->zeros, D(pl.x0), nullptr, n, cin, S, S, d, st));
  CDM_TRY(resblock_jvp<T>(m, m->blk[0], P(pl.x0), D(pl.x0), stat(0), stat(1), statt(0), statt(1), P(pl.h), D(pl.h), P(pl.y),
                          D(pl.y), P(pl.d1), D(pl.d1), bias, n, S, S, st));
  CDM_TRY(launch_maxpool_jvp<T>(P(pl.d1), D(pl.d1), P(pl.p1), D(pl.p1), stat(2), n, S, S, d, st));
  CDM_TRY(resblock_jvp<T>(m, m->blk[1], P(pl.p1), D(pl.p1), stat(2), stat(3), statt(2), statt(3), P(pl.h), D(pl.h), P(pl.y),
                          D(pl.y), P(pl.d2), D(pl.d2), bias, n, S2, S2, st));
  CDM_TRY(launch_maxpool_jvp<T>(P(pl.d2), D(pl.d2), P(pl.p2), D(pl.p2), stat(4), n, S2, S2, 2 * d, st));
  CDM_TRY(resblock_jvp<T>(m, m->blk[2], P(pl.p2), D(pl.p2), stat(4), stat(5), statt(4), statt(5), P(pl.h), D(pl.h), P(pl.y),
                          D(pl.y), P(pl.b1), D(pl.b1), bias, n, S4, S4, st));
  CDM_TRY(launch_upcat_stats<T>(P(pl.b1), P(pl.d2), P(pl.cat1), stat(6), n, S4, S4, 4 * d, 2 * d, st));
  CDM_TRY(launch_upcat_stats<T>(D(pl.b1), D(pl.d2), D(pl.cat1), nullptr, n, S4, S4, 4 * d, 2 * d, st));
  CDM_TRY(resblock_jvp<T>(m, m->blk[3], P(pl.cat1), D(pl.cat1), stat(6), stat(7), statt(6), statt(7), P(pl.h), D(pl.h),
                          P(pl.y), D(pl.y), P(pl.u1), D(pl.u1), bias, n, S2, S2, st));
  CDM_TRY(launch_upcat_stats<T>(P(pl.u1), P(pl.d1), P(pl.cat2), stat(8), n, S2, S2, 2 * d, d, st));
  CDM_TRY(launch_upcat_stats<T>(D(pl.u1), D(pl.d1), D(pl.cat2), nullptr, n, S2, S2, 2 * d, d, st));
  CDM_TRY(resblock_jvp<T>(m, m->blk[4], P(pl.cat2), D(pl.cat2), stat(8), stat(9), statt(8), statt(9), P(pl.h), D(pl.h),
                          P(pl.y), D(pl.y), P(pl.u2), D(pl.u2), bias, n, S, S, st));
  CDM_TRY(launch_out_conv<T>(P(pl.u2), m->out_w, m->out_b, eps, n, S * S, d, cin, st));
  CDM_TRY(launch_out_conv<T>(D(pl.u2), m->out_w, m->zeros, deps, n, S * S, d, cin, st));
  CDM_TRY(launch_rowdot(deps, v_out, vjv, n, cin * S * S, st));
  return CDM_OK;
}

}  // namespace cdm

extern "C" {

int cdm_set_microbatch(int samples) {
  g_microbatch = samples > 0 ? samples : -1;
  return CDM_OK;
}

int cdm_set_option(const char* name, int value) {
  if (!name) return fail(CDM_ERR_INVALID, "cdm_set_option: null name");
  std::string n(name);
  if (n == "microbatch") return cdm_set_microbatch(value);
  if (n == "conv_halo") { g_conv_halo = value; return CDM_OK; }
  if (n == "fuse_gn") { g_fuse_gn = value; return CDM_OK; }
  if (n == "conv_stack") { g_conv_stack = value; return CDM_OK; }
  if (n == "fuse_proj") { g_fuse_proj = value; return CDM_OK; }
  if (n == "grouped") { g_grouped = value; return CDM_OK; }
  if (n == "conv_pair") { set_conv_pair(value); return CDM_OK; }
  if (n == "pdl") { pdl_flag() = value; return CDM_OK; }
  if (n == "conv_pair64") { set_conv_pair64(value); return CDM_OK; }
  if (n == "stack_pair") { set_stack_pair(value); return CDM_OK; }
  if (n == "conv_scheme_c") { set_conv_scheme_c(value); return CDM_OK; }
  if (n == "init_conv_tc") { set_init_conv_tc(value); return CDM_OK; }
#ifdef CDM_INSTRUMENT
  if (n == "conv_timing") { g_conv_timing = value; return CDM_OK; }
#endif
  return fail(CDM_ERR_KEY, "cdm_set_option: unknown option %s", name);
}

int cdm_unet_create(const cdm_unet_config* cfg, int device, cdm_unet** out) {
  if (!cfg || !out) return fail(CDM_ERR_INVALID, "cdm_unet_create: null argument");
  if (cfg->base_dim != 64 || cfg->time_emb_dim % 4 || cfg->time_emb_dim > 1024)
    return fail(CDM_ERR_UNSUPPORTED, "cdm_unet_create: base_dim=%d time_emb_dim=%d (this build: base_dim 64)", cfg->base_dim, cfg->time_emb_dim);
  if (cfg->in_channels < 1 || cfg->in_channels > 4) return fail(CDM_ERR_UNSUPPORTED, "cdm_unet_create: in_channels=%d", cfg->in_channels);
  cdm_unet* m = new cdm_unet();
  m->cfg = *cfg;
  m->device = device;
  const int d = cfg->base_dim, td = cfg->time_emb_dim;
  add_spec(m, "time_mlp.1.weight", {td, d});
  add_spec(m, "time_mlp.1.bias", {td});
  add_spec(m, "time_mlp.3.weight", {td, td});
  add_spec(m, "time_mlp.3.bias", {td});
  if (cfg->num_classes > 0) add_spec(m, "label_emb.weight", {cfg->num_classes, td});
  add_spec(m, "init_conv.weight", {d, cfg->in_channels, 3, 3});
  add_spec(m, "init_conv.bias", {d});
  const int cins[5] = {d, d, 2 * d, 6 * d, 3 * d}, couts[5] = {d, 2 * d, 4 * d, 2 * d, d};
  for (int i = 0; i < 5; ++i) add_block_spec(m, BLOCK_NAMES[i], cins[i], couts[i], td);
  add_spec(m, "out_conv.weight", {cfg->in_channels, d, 1, 1});
  add_spec(m, "out_conv.bias", {cfg->in_channels});
  *out = m;
  return CDM_OK;
}

void cdm_unet_destroy(cdm_unet* m) {
  if (!m) return;
  for (void* p : m->allocs) cudaFree(p);
  delete m;
}

int cdm_unet_num_params(const cdm_unet* m) { return m ? (int)m->specs.size() : 0; }

const char* cdm_unet_param_key(const cdm_unet* m, int i, int64_t* numel) {
  if (!m || i < 0 || i >= (int)m->specs.size()) return nullptr;
  if (numel) *numel = m->specs[i].numel;
  return m->specs[i].key.c_str();
}

int cdm_unet_set_param(cdm_unet* m, const char* key, const float* host_data, int64_t numel) {
  if (!m || !key || !host_data) return fail(CDM_ERR_INVALID, "cdm_unet_set_param: null argument");
  std::string k(key);
  if (k == "@sin_freq") {   // optional: the sinusoidal frequency table exactly as torch evaluates it
    if (numel != m->cfg.base_dim / 2) return fail(CDM_ERR_KEY, "@sin_freq: expected %d values, got %lld", m->cfg.base_dim / 2, (long long)numel);
    m->host[k].assign(host_data, host_data + numel);
    m->finalized = false;
    return CDM_OK;
  }
  for (auto& s : m->specs)
    if (s.key == k) {
      if (s.numel != numel) return fail(CDM_ERR_KEY, "size mismatch for %s: expected %lld elements, got %lld", key, (long long)s.numel, (long long)numel);
      m->host[k].assign(host_data, host_data + numel);
      m->finalized = false;
      return CDM_OK;
    }
  return fail(CDM_ERR_KEY, "unexpected key %s", key);
}

int cdm_unet_finalize(cdm_unet* m) {
  if (!m) return fail(CDM_ERR_INVALID, "cdm_unet_finalize: null model");
  for (auto& s : m->specs)
    if (!m->host.count(s.key)) return fail(CDM_ERR_KEY, "missing key %s", s.key.c_str());
  CDM_CUDA_OK(cudaSetDevice(m->device));
  int sms = 0;
  CDM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device));
  m->num_sms = sms;
  for (void* p : m->allocs) cudaFree(p);
  m->allocs.clear();
  auto& H = m->host;
  const int d = m->cfg.base_dim, td = m->cfg.time_emb_dim, half = d / 2;

  std::vector<float> freq(half);
  if (H.count("@sin_freq")) freq = H["@sin_freq"];
  else {
    const float k = (float)(-(logf(10000.f)) / (float)(half - 1));   // torch: exp(arange(half) * -(log(10000)/(half-1)))
    const double kd = -std::log(10000.0) / (half - 1);
    (void)k;
    for (int i = 0; i < half; ++i) freq[i] = expf((float)i * (float)kd);
  }
  float* tmp = nullptr;
  CDM_TRY(upload(m, freq, &tmp)); m->temb.freq = tmp;
  CDM_TRY(upload(m, transpose(H["time_mlp.1.weight"], td, d), &tmp)); m->temb.w1t = tmp;
  CDM_TRY(upload(m, H["time_mlp.1.bias"], &tmp)); m->temb.b1 = tmp;
  CDM_TRY(upload(m, transpose(H["time_mlp.3.weight"], td, td), &tmp)); m->temb.w3t = tmp;
  CDM_TRY(upload(m, H["time_mlp.3.bias"], &tmp)); m->temb.b3 = tmp;
  m->temb.label = nullptr;
  if (m->cfg.num_classes > 0) { CDM_TRY(upload(m, H["label_emb.weight"], &tmp)); m->temb.label = tmp; }
  m->temb.D = d; m->temb.TD = td; m->temb.num_classes = m->cfg.num_classes;

  const int cins[5] = {d, d, 2 * d, 6 * d, 3 * d}, couts[5] = {d, 2 * d, 4 * d, 2 * d, d};
  int nb = 0;
  for (int i = 0; i < 5; ++i) { m->blk[i].bias_off = nb; nb += couts[i]; }
  m->nb_total = nb;
  m->temb.NB = nb;
  std::vector<float> wcat_t((size_t)td * nb), bcat(nb);
  for (int i = 0; i < 5; ++i) {
    BlockW& b = m->blk[i];
    const std::string p = BLOCK_NAMES[i];
    b.cin = cins[i]; b.cout = couts[i]; b.has_res = cins[i] != couts[i];
    const auto& tw = H[p + ".time_mlp.1.weight"];   // [cout][td]
    const auto& tb = H[p + ".time_mlp.1.bias"];
    const auto& cb = H[p + ".block1.2.bias"];
    for (int o = 0; o < b.cout; ++o) {
      for (int k = 0; k < td; ++k) wcat_t[(size_t)k * nb + b.bias_off + o] = tw[(size_t)o * td + k];
      bcat[b.bias_off + o] = tb[o] + cb[o];
    }
    CDM_TRY(upload(m, H[p + ".block1.0.weight"], &b.g1));
    CDM_TRY(upload(m, H[p + ".block1.0.bias"], &b.b1));
    CDM_TRY(upload(m, H[p + ".block2.0.weight"], &b.g2));
    CDM_TRY(upload(m, H[p + ".block2.0.bias"], &b.b2));
    std::vector<float> kn;
    std::vector<h16> nk;
    pack_conv(H[p + ".block1.2.weight"], b.cout, b.cin, 9, nullptr, 0, kn, nk);
    CDM_TRY(upload(m, kn, &b.w1_f32));
    CDM_TRY(upload(m, nk, &b.w1_h16));
    {
      std::vector<h16> xh, xl;
      pack_conv_x3(H[p + ".block1.2.weight"], b.cout, b.cin, 9, nullptr, 0, xh, xl);
      CDM_TRY(upload(m, xh, &b.w1_x3h));
      CDM_TRY(upload(m, xl, &b.w1_x3l));
      pack_conv_x3(H[p + ".block2.3.weight"], b.cout, b.cout, 9, b.has_res ? &H[p + ".res_conv.weight"] : nullptr, b.cin, xh, xl);
      CDM_TRY(upload(m, xh, &b.w2_x3h));
      CDM_TRY(upload(m, xl, &b.w2_x3l));
    }
    pack_conv_halo(H[p + ".block1.2.weight"], b.cout, b.cin, nullptr, 0, nk);
    CDM_TRY(upload(m, nk, &b.w1_halo));
    b.w1_stack = b.w2_stack = nullptr;
    if (b.cout == 64) {
      pack_conv_stack3(H[p + ".block1.2.weight"], b.cin, nullptr, 0, nk);
      CDM_TRY(upload(m, nk, &b.w1_stack));
      pack_conv_stack3(H[p + ".block2.3.weight"], b.cout, b.has_res ? &H[p + ".res_conv.weight"] : nullptr, b.cin, nk);
      CDM_TRY(upload(m, nk, &b.w2_stack));
    }
    std::vector<float> bias2 = H[p + ".block2.3.bias"];
    if (b.has_res) {
      pack_conv(H[p + ".block2.3.weight"], b.cout, b.cout, 9, &H[p + ".res_conv.weight"], b.cin, kn, nk);
      const auto& rb = H[p + ".res_conv.bias"];
      for (int o = 0; o < b.cout; ++o) bias2[o] += rb[o];
    } else {
      pack_conv(H[p + ".block2.3.weight"], b.cout, b.cout, 9, nullptr, 0, kn, nk);
    }
    CDM_TRY(upload(m, kn, &b.w2_f32));
    CDM_TRY(upload(m, nk, &b.w2_h16));
    pack_conv_halo(H[p + ".block2.3.weight"], b.cout, b.cout, b.has_res ? &H[p + ".res_conv.weight"] : nullptr, b.cin, nk);
    CDM_TRY(upload(m, nk, &b.w2_halo));
    CDM_TRY(upload(m, bias2, &b.bias2));
  }
  CDM_TRY(upload(m, wcat_t, &tmp)); m->temb.wcat_t = tmp;
  CDM_TRY(upload(m, bcat, &tmp)); m->temb.bcat = tmp;
  CDM_TRY(upload(m, H["init_conv.weight"], &m->init_w));
  CDM_TRY(upload(m, H["init_conv.bias"], &m->init_b));
  CDM_TRY(upload(m, H["out_conv.weight"], &m->out_w));
  CDM_TRY(upload(m, H["out_conv.bias"], &m->out_b));
  CDM_TRY(upload(m, std::vector<float>(1024, 0.f), &m->zeros));
  m->finalized = true;
  return CDM_OK;
}

size_t cdm_unet_workspace_bytes(const cdm_unet* m, int B, int img_size, int precision) {
  if (!m || B <= 0 || img_size <= 0 || !m->nb_total) return 0;
  return make_plan(m, B, img_size, precision).total;
}

}  // extern "C"

// One expert forward.  `uniform`: the caller guarantees that every sample carries the same (t, y) -- always true inside a
// sampler loop -- so the time/label embedding is ONE row (temb kernel with a single CTA) that every conv reads with
// bias_stride 0, instead of B identical rows ("should be 1 row / step / expert").
static int unet_forward_impl(cdm_unet* m, const float* x, const float* t, const int64_t* y, float* eps, int B, int img_size,
                             int precision, void* workspace, size_t workspace_bytes, cudaStream_t st, bool uniform) {
  if (!m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_unet_forward: parameters not finalized");
  if (m->cfg.num_classes > 0 && !y) return fail(CDM_ERR_INVALID, "Class labels `y` must be provided for a conditional UNet.");
  if (precision != CDM_PREC_FP32 && precision != CDM_PREC_F16 && precision != CDM_PREC_F16X3)
    return fail(CDM_ERR_INVALID, "cdm_unet_forward: precision %d", precision);
  if (img_size % 4) return fail(CDM_ERR_UNSUPPORTED, "cdm_unet_forward: img_size=%d must be a multiple of 4", img_size);
  const Plan pl = make_plan(m, B, img_size, precision);
  if (!workspace || workspace_bytes < pl.total)
    return fail(CDM_ERR_WORKSPACE, "cdm_unet_forward: workspace %zu bytes < required %zu", workspace_bytes, pl.total);
  uint8_t* ws = (uint8_t*)workspace;
  float* bias = reinterpret_cast<float*>(ws + pl.block_bias);
  if (uniform) CDM_TRY(launch_temb_row(m->temb, t, y, nullptr, bias, bias + (size_t)B * m->nb_total, st));
  else CDM_TRY(launch_temb(m->temb, t, y, nullptr, bias, B, st));
  const int bias_stride = uniform ? 0 : m->nb_total;
  const size_t img = (size_t)m->cfg.in_channels * img_size * img_size;
  for (int b0 = 0; b0 < B; b0 += pl.chunk) {
    const int n = (B - b0 < pl.chunk) ? B - b0 : pl.chunk;
    const float* bias_c = bias + (size_t)b0 * bias_stride;
    if (precision != CDM_PREC_F16)
      CDM_TRY(forward_chunk<float>(m, pl, ws, x + b0 * img, eps + b0 * img, bias_c, bias_stride, n, img_size, st, precision == CDM_PREC_F16X3));
    else
      CDM_TRY(forward_chunk<h16>(m, pl, ws, x + b0 * img, eps + b0 * img, bias_c, bias_stride, n, img_size, st));
  }
  m->last_ws = workspace; m->last_prec = precision; m->last_B = B; m->last_S = img_size;
  return CDM_OK;
}

// K experts of the same architecture, one grouped launch per convolution (forward_chunk_group).  Each expert works in its own
// slice of the workspace (K x the single-expert size).  `uniform`: as unet_forward_impl.
static bool group_compatible(cdm_unet* const* ms, int K, int precision) {
  if (K < 2 || K > GROUP_MAX || precision != CDM_PREC_F16 || !grouped_enabled()) return false;
  for (int k = 0; k < K; ++k)
    if (!ms[k] || !ms[k]->finalized || ms[k]->cfg.base_dim != ms[0]->cfg.base_dim || ms[k]->cfg.time_emb_dim != ms[0]->cfg.time_emb_dim ||
        ms[k]->nb_total != ms[0]->nb_total || ms[k]->device != ms[0]->device)
      return false;
  return true;
}
static size_t group_slice_bytes(const cdm_unet* m, int B, int S) { return (make_plan(m, B, S, CDM_PREC_F16).total + 255) & ~(size_t)255; }

static int unet_forward_group_impl(cdm_unet* const* ms, int K, const float* const* xs, const float* t, const int64_t* const* ys,
                                   float* const* epss, int B, int img_size, void* workspace, size_t workspace_bytes, cudaStream_t st,
                                   bool uniform_t, int y_uniform) {
  if (img_size % 4) return fail(CDM_ERR_UNSUPPORTED, "cdm_unet_forward_grouped: img_size=%d must be a multiple of 4", img_size);
  const Plan pl = make_plan(ms[0], B, img_size, CDM_PREC_F16);
  const size_t slice = group_slice_bytes(ms[0], B, img_size);
  if (!workspace || workspace_bytes < slice * K)
    return fail(CDM_ERR_WORKSPACE, "cdm_unet_forward_grouped: workspace %zu bytes < required %zu", workspace_bytes, slice * K);
  uint8_t* wss[GROUP_MAX];
  const float* bias[GROUP_MAX];
  int stride = 0;
  bool all_uniform = true;
  for (int k = 0; k < K; ++k) {
    if (ms[k]->cfg.num_classes > 0 && !(ys && ys[k])) return fail(CDM_ERR_INVALID, "Class labels `y` must be provided for a conditional UNet.");
    all_uniform = all_uniform && uniform_t && (!(ys && ys[k]) || y_uniform);
  }
  for (int k = 0; k < K; ++k) {
    wss[k] = (uint8_t*)workspace + k * slice;
    float* bk = reinterpret_cast<float*>(wss[k] + pl.block_bias);
    const int64_t* yk = ys ? ys[k] : nullptr;
    // one row per expert when every sample shares (t, y); all experts must then agree on the row stride
    if (all_uniform) CDM_TRY(launch_temb_row(ms[k]->temb, t, yk, nullptr, bk, bk + (size_t)B * ms[k]->nb_total, st));
    else CDM_TRY(launch_temb(ms[k]->temb, t, yk, nullptr, bk, B, st));
    bias[k] = bk;
  }
  stride = all_uniform ? 0 : ms[0]->nb_total;
  for (int b0 = 0; b0 < B; b0 += pl.chunk) {
    const int n = (B - b0 < pl.chunk) ? B - b0 : pl.chunk;
    const float* xc[GROUP_MAX];
    float* ec[GROUP_MAX];
    const float* bc[GROUP_MAX];
    for (int k = 0; k < K; ++k) {
      const size_t img = (size_t)ms[k]->cfg.in_channels * img_size * img_size;
      xc[k] = xs[k] + b0 * img; ec[k] = epss[k] + b0 * img; bc[k] = bias[k] + (size_t)b0 * stride;
    }
    CDM_TRY(forward_chunk_group(ms, K, pl, wss, xc, ec, bc, stride, n, img_size, st));
  }
  for (int k = 0; k < K; ++k) ms[k]->last_ws = nullptr;
  return CDM_OK;
}

__global__ void fill_f32_kernel(float* p, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

extern "C" {

int cdm_unet_forward(cdm_unet* m, const float* x, const float* t, const int64_t* y, float* eps, int B, int img_size,
                     int precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (!m || !x || !t || !eps) return fail(CDM_ERR_INVALID, "cdm_unet_forward: null argument");
  return unet_forward_impl(m, x, t, y, eps, B, img_size, precision, workspace, workspace_bytes, (cudaStream_t)stream, false);
}

size_t cdm_unet_forward_grouped_workspace_bytes(cdm_unet* const* experts, int K, int B, int img_size, int precision) {
  if (!experts || K < 1 || B <= 0 || img_size <= 0) return 0;
  for (int k = 0; k < K; ++k)
    if (!experts[k] || !experts[k]->nb_total) return 0;
  if (group_compatible(experts, K, precision)) return group_slice_bytes(experts[0], B, img_size) * K;
  size_t w = 0;
  for (int k = 0; k < K; ++k) { const size_t wk = cdm_unet_workspace_bytes(experts[k], B, img_size, precision); if (wk > w) w = wk; }
  return w;
}

int cdm_unet_forward_grouped(cdm_unet* const* experts, int K, const float* const* x, const float* t, const int64_t* const* y,
                             float* const* eps, int B, int img_size, int precision, void* workspace, size_t workspace_bytes,
                             void* stream) {
  if (B <= 0) return CDM_OK;
  if (!experts || !x || !t || !eps) return fail(CDM_ERR_INVALID, "cdm_unet_forward_grouped: null argument");
  if (K < 1 || K > CDM_MAX_EXPERTS) return fail(CDM_ERR_INVALID, "cdm_unet_forward_grouped: K=%d out of range 1..%d", K, CDM_MAX_EXPERTS);
  for (int k = 0; k < K; ++k)
    if (!experts[k] || !x[k] || !eps[k]) return fail(CDM_ERR_INVALID, "cdm_unet_forward_grouped: null expert / x / eps %d", k);
  cudaStream_t st = (cudaStream_t)stream;
  if (group_compatible(experts, K, precision))
    return unet_forward_group_impl(experts, K, x, t, y, eps, B, img_size, workspace, workspace_bytes, st, false, 0);
  for (int k = 0; k < K; ++k)       // experts that cannot share a launch (other precision / more than GROUP_MAX / one expert): back to back
    CDM_TRY(unet_forward_impl(experts[k], x[k], t, y ? y[k] : nullptr, eps[k], B, img_size, precision, workspace, workspace_bytes, st, false));
  return CDM_OK;
}

// ---- whole reverse-SDE chain for K UNet experts (mnist/compose_scores.py:26-46) in one host call --------------------
static size_t sample_ws_layout(cdm_unet* const* experts, int K, int B, int S, int precision, size_t* eps_off, size_t* t_off) {
  size_t expert_ws = 0, img = 0;
  for (int k = 0; k < K; ++k) {
    const size_t w = cdm_unet_workspace_bytes(experts[k], B, S, precision);
    if (w > expert_ws) expert_ws = w;
    img = (size_t)experts[k]->cfg.in_channels * S * S * sizeof(float);
  }
  if (group_compatible(experts, K, precision)) expert_ws = group_slice_bytes(experts[0], B, S) * K;    // one slice per expert
  expert_ws = (expert_ws + 255) & ~(size_t)255;
  const size_t eps_bytes = ((size_t)B * img + 255) & ~(size_t)255;
  if (eps_off) *eps_off = expert_ws;
  if (t_off) *t_off = expert_ws + (size_t)K * eps_bytes;
  return expert_ws + (size_t)K * eps_bytes + (((size_t)B * 4 + 255) & ~(size_t)255);
}

size_t cdm_unet_sample_workspace_bytes(cdm_unet* const* experts, int K, int B, int img_size, int precision) {
  if (!experts || K < 1 || K > CDM_MAX_EXPERTS || B <= 0 || img_size <= 0) return 0;
  for (int k = 0; k < K; ++k)
    if (!experts[k] || !experts[k]->nb_total) return 0;
  return sample_ws_layout(experts, K, B, img_size, precision, nullptr, nullptr);
}

int cdm_unet_sample_sde(cdm_unet* const* experts, const float* w, int K, float* x, const int64_t* const* y, int y_uniform,
                        const float* z, const cdm_rng* rng, const float* step_coef_host, int n_steps, float dt, int B,
                        int img_size, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || n_steps <= 0) return CDM_OK;
  if (!experts || !w || !x || !step_coef_host) return fail(CDM_ERR_INVALID, "cdm_unet_sample_sde: null argument");
  if (K < 1 || K > CDM_MAX_EXPERTS) return fail(CDM_ERR_INVALID, "cdm_unet_sample_sde: K=%d out of range 1..%d", K, CDM_MAX_EXPERTS);
  if (!z && !rng) return fail(CDM_ERR_INVALID, "cdm_unet_sample_sde: neither injected noise nor an rng");
  for (int k = 0; k < K; ++k) {
    if (!experts[k]) return fail(CDM_ERR_INVALID, "cdm_unet_sample_sde: null expert %d", k);
    if (experts[k]->cfg.in_channels != experts[0]->cfg.in_channels)
      return fail(CDM_ERR_UNSUPPORTED, "cdm_unet_sample_sde: experts disagree on the number of image channels");
  }
  size_t eps_off, t_off;
  const size_t need = sample_ws_layout(experts, K, B, img_size, precision, &eps_off, &t_off);
  if (!workspace || workspace_bytes < need)
    return fail(CDM_ERR_WORKSPACE, "cdm_unet_sample_sde: workspace %zu bytes < required %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  const int C = experts[0]->cfg.in_channels, HW = img_size * img_size;
  const size_t eps_bytes = (t_off - eps_off) / K;
  float* tbuf = reinterpret_cast<float*>(ws + t_off);
  const float* eps[CDM_MAX_EXPERTS];
  int ech[CDM_MAX_EXPERTS];
  for (int k = 0; k < K; ++k) { eps[k] = reinterpret_cast<const float*>(ws + eps_off + k * eps_bytes); ech[k] = C; }
  for (int i = 0; i < n_steps; ++i) {
    const float* cf = step_coef_host + 4 * (size_t)i;       // {t, a, c, g}
    fill_f32_kernel<<<ceil_div(B, 256), 256, 0, st>>>(tbuf, cf[0], B);
    CDM_LAUNCH_OK("fill_f32_kernel");
    if (group_compatible(experts, K, precision)) {          // every convolution of the K experts in ONE grouped launch
      const float* xs[GROUP_MAX];
      float* es[GROUP_MAX];
      for (int k = 0; k < K; ++k) { xs[k] = x; es[k] = const_cast<float*>(eps[k]); }
      CDM_TRY(unet_forward_group_impl(experts, K, xs, tbuf, y, es, B, img_size, ws, eps_off, st, true, y_uniform));
    } else
    for (int k = 0; k < K; ++k) {
      const int64_t* yk = y ? y[k] : nullptr;
      const bool uniform = !yk || y_uniform;
      CDM_TRY(unet_forward_impl(experts[k], x, tbuf, yk, const_cast<float*>(eps[k]), B, img_size, precision, ws, eps_off, st, uniform));
    }
    cdm_rng r{};
    if (rng) { r = *rng; r.step += (uint64_t)i; }
    CDM_TRY(cdm_step_sde(x, eps, ech, w, K, z ? z + (size_t)i * B * C * HW : nullptr, (rng && !z) ? &r : nullptr, cf[1], cf[2], dt,
                         cf[3], x, B, C, HW, stream));
  }
  return CDM_OK;
}

// ---- whole DDIM chain for K UNet experts (shapes/compose_images_ddim.py:39-68) ---------------------------------------
// workspace: [expert forward workspace | K eps buffers of B*C*HW floats | gray B*HW | t B]
static size_t ddim_ws_layout(cdm_unet* const* experts, int K, int B, int C, int S, int precision, size_t* eps_off, size_t* gray_off,
                             size_t* t_off) {
  size_t expert_ws = 0;
  for (int k = 0; k < K; ++k) {
    const size_t w = cdm_unet_workspace_bytes(experts[k], B, S, precision);
    if (w > expert_ws) expert_ws = w;
  }
  if (group_compatible(experts, K, precision)) expert_ws = group_slice_bytes(experts[0], B, S) * K;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  expert_ws = up(expert_ws);
  const size_t eps_bytes = up((size_t)B * C * S * S * sizeof(float)), gray_bytes = up((size_t)B * S * S * sizeof(float));
  if (eps_off) *eps_off = expert_ws;
  if (gray_off) *gray_off = expert_ws + (size_t)K * eps_bytes;
  if (t_off) *t_off = expert_ws + (size_t)K * eps_bytes + gray_bytes;
  return expert_ws + (size_t)K * eps_bytes + gray_bytes + up((size_t)B * 4);
}

static int check_chain_experts(const char* who, cdm_unet* const* experts, int K, int C) {
  if (!experts) return fail(CDM_ERR_INVALID, "%s: null experts", who);
  if (K < 1 || K > CDM_MAX_EXPERTS) return fail(CDM_ERR_INVALID, "%s: K=%d out of range 1..%d", who, K, CDM_MAX_EXPERTS);
  for (int k = 0; k < K; ++k) {
    if (!experts[k]) return fail(CDM_ERR_INVALID, "%s: null expert %d", who, k);
    if (!experts[k]->finalized) return fail(CDM_ERR_NOT_READY, "%s: expert %d not finalized", who, k);
    const int ic = experts[k]->cfg.in_channels;
    if (ic != C && !(ic == 1 && C == 3))
      return fail(CDM_ERR_UNSUPPORTED, "%s: expert %d has %d input channels; the state has %d (1-channel experts read Grayscale(x) of an RGB state)", who, k, ic, C);
  }
  return CDM_OK;
}

size_t cdm_unet_sample_ddim_workspace_bytes(cdm_unet* const* experts, int K, int B, int C, int img_size, int precision) {
  if (!experts || K < 1 || K > CDM_MAX_EXPERTS || B <= 0 || img_size <= 0 || C <= 0) return 0;
  for (int k = 0; k < K; ++k)
    if (!experts[k] || !experts[k]->nb_total) return 0;
  return ddim_ws_layout(experts, K, B, C, img_size, precision, nullptr, nullptr, nullptr);
}

int cdm_unet_sample_ddim(cdm_unet* const* experts, const float* w, int K, float wsum, float* x, const int64_t* const* y,
                         int y_uniform, const float* step_coef_host, int n_steps, int B, int C, int img_size, int precision,
                         void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || n_steps <= 0) return CDM_OK;
  if (!w || !x || !step_coef_host) return fail(CDM_ERR_INVALID, "cdm_unet_sample_ddim: null argument");
  CDM_TRY(check_chain_experts("cdm_unet_sample_ddim", experts, K, C));
  size_t eps_off, gray_off, t_off;
  const size_t need = ddim_ws_layout(experts, K, B, C, img_size, precision, &eps_off, &gray_off, &t_off);
  if (!workspace || workspace_bytes < need)
    return fail(CDM_ERR_WORKSPACE, "cdm_unet_sample_ddim: workspace %zu bytes < required %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  const int HW = img_size * img_size;
  const size_t eps_bytes = (gray_off - eps_off) / K;
  float* gray = reinterpret_cast<float*>(ws + gray_off);
  float* tbuf = reinterpret_cast<float*>(ws + t_off);
  const float* eps[CDM_MAX_EXPERTS];
  int ech[CDM_MAX_EXPERTS];
  bool need_gray = false;
  for (int k = 0; k < K; ++k) {
    eps[k] = reinterpret_cast<const float*>(ws + eps_off + k * eps_bytes);
    ech[k] = experts[k]->cfg.in_channels;
    need_gray = need_gray || (ech[k] == 1 && C == 3);
  }
  if (need_gray) CDM_TRY(cdm_grayscale(x, gray, B, HW, stream));          // afterwards the step kernel emits Grayscale(x') itself
  for (int i = 0; i < n_steps; ++i) {
    const float* c0 = step_coef_host + 3 * (size_t)i;       // {t, alpha, sigma} at grid points i and i + 1
    const float* c1 = c0 + 3;
    fill_f32_kernel<<<ceil_div(B, 256), 256, 0, st>>>(tbuf, c0[0], B);
    CDM_LAUNCH_OK("fill_f32_kernel");
    if (group_compatible(experts, K, precision)) {
      const float* xs[GROUP_MAX];
      float* es[GROUP_MAX];
      for (int k = 0; k < K; ++k) { xs[k] = (ech[k] == 1 && C == 3) ? gray : x; es[k] = const_cast<float*>(eps[k]); }
      CDM_TRY(unet_forward_group_impl(experts, K, xs, tbuf, y, es, B, img_size, ws, eps_off, st, true, y_uniform));
    } else
    for (int k = 0; k < K; ++k) {
      const int64_t* yk = y ? y[k] : nullptr;
      const bool uniform = !yk || y_uniform;
      const float* xin = (ech[k] == 1 && C == 3) ? gray : x;
      CDM_TRY(unet_forward_impl(experts[k], xin, tbuf, yk, const_cast<float*>(eps[k]), B, img_size, precision, ws, eps_off, st, uniform));
    }
    CDM_TRY(cdm_step_ddim(x, eps, ech, w, K, wsum, c0[1], c0[2], c1[1], c1[2], x, need_gray ? gray : nullptr, B, C, HW, stream));
  }
  return CDM_OK;
}

size_t cdm_unet_jvp_workspace_bytes(const cdm_unet* m, int B, int img_size, int precision) {
  if (!m || B <= 0 || img_size <= 0 || !m->nb_total) return 0;
  if (precision == CDM_PREC_F16X3) precision = CDM_PREC_FP32;
  const Plan pl = make_plan(m, B, img_size, precision);
  const size_t chunk_img = (size_t)pl.chunk * m->cfg.in_channels * img_size * img_size * sizeof(float);
  return 2 * pl.total + ((chunk_img + 255) & ~(size_t)255);
}

int cdm_unet_forward_jvp(cdm_unet* m, const float* x, const float* t, const int64_t* y, const float* v_in,
                         const float* v_out, float* eps, float* vjv, int B, int img_size, int precision, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (B <= 0) return CDM_OK;   // an empty batch is a no-op (checked before the pointers: empty tensors have none)
  if (!m || !x || !t || !eps || !v_in || !vjv) return fail(CDM_ERR_INVALID, "cdm_unet_forward_jvp: null argument");
  if (!m->finalized) return fail(CDM_ERR_NOT_READY, "cdm_unet_forward_jvp: parameters not finalized");
  if (m->cfg.num_classes > 0 && !y) return fail(CDM_ERR_INVALID, "Class labels `y` must be provided for a conditional UNet.");
  if (img_size % 4) return fail(CDM_ERR_UNSUPPORTED, "cdm_unet_forward_jvp: img_size=%d must be a multiple of 4", img_size);
  if (precision == CDM_PREC_F16X3) precision = CDM_PREC_FP32;   // the tangent graph has no three-term variant: fp32-class = the CUDA-core path
  if (precision != CDM_PREC_FP32 && precision != CDM_PREC_F16) return fail(CDM_ERR_INVALID, "cdm_unet_forward_jvp: precision %d", precision);
  if (B <= 0) return CDM_OK;
  if (!v_out) v_out = v_in;
  const Plan pl = make_plan(m, B, img_size, precision);
  const size_t need = cdm_unet_jvp_workspace_bytes(m, B, img_size, precision);
  if (!workspace || workspace_bytes < need)
    return fail(CDM_ERR_WORKSPACE, "cdm_unet_forward_jvp: workspace %zu bytes < required %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  float* bias = reinterpret_cast<float*>(ws + pl.block_bias);
  float* deps = reinterpret_cast<float*>(ws + 2 * pl.total);
  CDM_TRY(launch_temb(m->temb, t, y, nullptr, bias, B, st));
  const size_t img = (size_t)m->cfg.in_channels * img_size * img_size;
  for (int b0 = 0; b0 < B; b0 += pl.chunk) {
    const int n = (B - b0 < pl.chunk) ? B - b0 : pl.chunk;
    if (precision == CDM_PREC_FP32)
      CDM_TRY(forward_jvp_chunk<float>(m, pl, ws, x + b0 * img, v_in + b0 * img, v_out + b0 * img, eps + b0 * img, deps, vjv + b0,
                                       bias + (size_t)b0 * m->nb_total, n, img_size, st));
    else
      CDM_TRY(forward_jvp_chunk<h16>(m, pl, ws, x + b0 * img, v_in + b0 * img, v_out + b0 * img, eps + b0 * img, deps, vjv + b0,
                                     bias + (size_t)b0 * m->nb_total, n, img_size, st));
  }
  m->last_ws = nullptr;
  return CDM_OK;
}

// ---- whole Ito / kappa probability-flow chain for K = 2 .. 4 UNet experts (shapes/compose_images_ito.py:88-137, _2.py) -----
// workspace: [JVP workspace | K eps (B*3*HW) | K div (B) | gray B*HW | probe B*3*HW | v_in B*HW | v_out B*HW | t B]
struct ItoLayout { size_t eps, div, gray, probe, vin, vout, t, total, jvp; };
static ItoLayout ito_ws_layout(cdm_unet* const* experts, int K, int B, int S, int precision) {
  ItoLayout L{};
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t jvp = 0;
  for (int k = 0; k < K; ++k) {
    const size_t w = cdm_unet_jvp_workspace_bytes(experts[k], B, S, precision);
    if (w > jvp) jvp = w;
  }
  const size_t plane = up((size_t)B * S * S * sizeof(float)), img = up((size_t)B * 3 * S * S * sizeof(float));
  size_t off = L.jvp = up(jvp);
  L.eps = off; off += (size_t)K * img;
  L.div = off; off += (size_t)K * up((size_t)B * 4);
  L.gray = off; off += plane;
  L.probe = off; off += img;
  L.vin = off; off += plane;
  L.vout = off; off += plane;
  L.t = off; off += up((size_t)B * 4);
  L.total = off;
  return L;
}

__global__ void channel_sum3_kernel(const float* __restrict__ v, float* __restrict__ out, int64_t n, int HW) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t b = i / HW, p = i % HW;
  const float* vb = v + b * 3 * HW + p;
  out[i] = fadd(fadd(vb[0], vb[HW]), vb[2 * (int64_t)HW]);      // torch: sum over dim 1 in channel order
}

size_t cdm_unet_sample_ito_workspace_bytes(cdm_unet* const* experts, int K, int B, int img_size, int precision) {
  if (!experts || K < 2 || K > 4 || B <= 0 || img_size <= 0) return 0;
  for (int k = 0; k < K; ++k)
    if (!experts[k] || !experts[k]->nb_total) return 0;
  return ito_ws_layout(experts, K, B, img_size, precision).total;
}

int cdm_unet_sample_ito(cdm_unet* const* experts, int K, float* x, const int64_t* const* y, int variant,
                        const float* const* probes, const cdm_rng* rng, const float* step_coef_host, int n_steps, float dt, int B,
                        int img_size, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || n_steps <= 0) return CDM_OK;
  if (!x || !step_coef_host) return fail(CDM_ERR_INVALID, "cdm_unet_sample_ito: null argument");
  if (K < 2 || K > 4) return fail(CDM_ERR_UNSUPPORTED, "cdm_unet_sample_ito: K=%d (2..4)", K);
  if (variant != 0 && variant != 1) return fail(CDM_ERR_INVALID, "cdm_unet_sample_ito: variant %d", variant);
  if (!probes && !rng) return fail(CDM_ERR_INVALID, "cdm_unet_sample_ito: neither injected probes nor an rng");
  CDM_TRY(check_chain_experts("cdm_unet_sample_ito", experts, K, 3));
  if (K == 2 && experts[1]->cfg.in_channels != 3)
    return fail(CDM_ERR_UNSUPPORTED, "cdm_unet_sample_ito: with two experts the last one must be the 3-channel (colour) expert");
  const ItoLayout L = ito_ws_layout(experts, K, B, img_size, precision);
  if (!workspace || workspace_bytes < L.total)
    return fail(CDM_ERR_WORKSPACE, "cdm_unet_sample_ito: workspace %zu bytes < required %zu", workspace_bytes, L.total);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  const int HW = img_size * img_size;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t img = up((size_t)B * 3 * HW * sizeof(float)), dvb = up((size_t)B * 4);
  float* gray = reinterpret_cast<float*>(ws + L.gray);
  float* probe_buf = reinterpret_cast<float*>(ws + L.probe);
  float* vin = reinterpret_cast<float*>(ws + L.vin);
  float* vout = reinterpret_cast<float*>(ws + L.vout);
  float* tbuf = reinterpret_cast<float*>(ws + L.t);
  const float* eps[4];
  const float* div[4];
  int ech[4];
  float dscale[4];
  bool any_gray = false;
  for (int k = 0; k < K; ++k) {
    eps[k] = reinterpret_cast<const float*>(ws + L.eps + k * img);
    div[k] = reinterpret_cast<const float*>(ws + L.div + k * dvb);
    ech[k] = experts[k]->cfg.in_channels;
    // "beta" variant: the divergence of a 1-channel expert is taken w.r.t. its grayscale input and scaled by 3 (:113)
    dscale[k] = (ech[k] == 1 && variant == 0) ? 3.f : 1.f;
    any_gray = any_gray || ech[k] == 1;
  }
  for (int i = 0; i < n_steps; ++i) {
    const float* cf = step_coef_host + 4 * (size_t)i;       // {t, sigma, a, coef}
    fill_f32_kernel<<<ceil_div(B, 256), 256, 0, st>>>(tbuf, cf[0], B);
    CDM_LAUNCH_OK("fill_f32_kernel");
    if (any_gray) CDM_TRY(cdm_grayscale(x, gray, B, HW, stream));
    for (int k = 0; k < K; ++k) {
      const int pc = (ech[k] == 1 && variant == 0) ? 1 : 3;          // channels of expert k's Hutchinson probe
      const float* pk;
      if (probes && probes[k]) {
        pk = probes[k] + (size_t)i * B * pc * HW;
      } else {
        if (!rng) return fail(CDM_ERR_INVALID, "cdm_unet_sample_ito: no probes for expert %d and no rng", k);
        cdm_rng r = *rng;
        r.step += (uint64_t)i * K + k;
        CDM_TRY(cdm_fill_normal(probe_buf, (int64_t)B * pc * HW, &r, stream));
        pk = probe_buf;
      }
      const int64_t* yk = y ? y[k] : nullptr;
      float* ek = const_cast<float*>(eps[k]);
      float* dk = const_cast<float*>(div[k]);
      if (ech[k] == 3) {
        CDM_TRY(cdm_unet_forward_jvp(experts[k], x, tbuf, yk, pk, nullptr, ek, dk, B, img_size, precision, ws, L.jvp, stream));
      } else if (variant == 0) {
        CDM_TRY(cdm_unet_forward_jvp(experts[k], gray, tbuf, yk, pk, nullptr, ek, dk, B, img_size, precision, ws, L.jvp, stream));
      } else {   // divergence through Grayscale w.r.t. the RGB input: v_in = Grayscale(v), v_out = sum over channels of v
        CDM_TRY(cdm_grayscale(pk, vin, B, HW, stream));
        const int64_t n = (int64_t)B * HW;
        channel_sum3_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(pk, vout, n, HW);
        CDM_LAUNCH_OK("channel_sum3_kernel");
        CDM_TRY(cdm_unet_forward_jvp(experts[k], gray, tbuf, yk, vin, vout, ek, dk, B, img_size, precision, ws, L.jvp, stream));
      }
    }
    CDM_TRY(cdm_step_ode_kappa_k(x, eps, ech, div, dscale, K, cf[1], cf[2], cf[3], dt, 1e-9f, x, nullptr, B, 3, HW, stream));
  }
  return CDM_OK;
}

int cdm_unet_debug_read(cdm_unet* m, const char* name, float* out, int B, int img_size, void* stream) {
  if (!m || !name || !out) return fail(CDM_ERR_INVALID, "cdm_unet_debug_read: null argument");
  if (!m->last_ws || m->last_B != B || m->last_S != img_size) return fail(CDM_ERR_NOT_READY, "cdm_unet_debug_read: no matching forward");
  const Plan pl = make_plan(m, B, img_size, m->last_prec);
  if (pl.chunk < B) return fail(CDM_ERR_UNSUPPORTED, "cdm_unet_debug_read: batch was micro-batched");
  const int d = m->cfg.base_dim, S = img_size;
  size_t off; int hw, c;
  std::string n(name);
  if (n == "x0") { off = pl.x0; hw = S * S; c = d; }
  else if (n == "d1") { off = pl.d1; hw = S * S; c = d; }
  else if (n == "d2") { off = pl.d2; hw = S * S / 4; c = 2 * d; }
  else if (n == "b1") { off = pl.b1; hw = S * S / 16; c = 4 * d; }
  else if (n == "u1") { off = pl.u1; hw = S * S / 4; c = 2 * d; }
  else if (n == "u2") { off = pl.u2; hw = S * S; c = d; }
  else return fail(CDM_ERR_KEY, "cdm_unet_debug_read: unknown intermediate %s", name);
  uint8_t* ws = (uint8_t*)m->last_ws;
  if (m->last_prec != CDM_PREC_F16) return launch_nhwc_to_nchw<float>((const float*)(ws + off), out, B, hw, c, (cudaStream_t)stream);
  return launch_nhwc_to_nchw<h16>((const h16*)(ws + off), out, B, hw, c, (cudaStream_t)stream);
}

}  // extern "C"
