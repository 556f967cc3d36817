// Library-level C ABI: version, error text, device check, and the single-layer test hook.
#include <string.h>

#include <vector>

#include "group.cuh"
#include "layers.cuh"

using namespace cdm;

namespace cdm {

template <typename T>
static int debug_conv_t(const float* x, const float* w_host, const float* bias, int bias_rows, const float* res,
                        const float* wres_host, const float* identity, float* out, float* stats_out, int B, int Cin,
                        int Cres, int Cout, int H, int W, int taps, cudaStream_t st, int variant = 0) {
  const bool halo = variant == 1, stack = variant == 2;
  const int HW = H * W;
  std::vector<float> w(w_host, w_host + (size_t)Cout * Cin * taps), wres, kn;
  std::vector<h16> nk;
  if (res) wres.assign(wres_host, wres_host + (size_t)Cout * Cres);
  pack_conv(w, Cout, Cin, taps, res ? &wres : nullptr, Cres, kn, nk);
  T *a = nullptr, *r = nullptr, *idn = nullptr, *o = nullptr;
  void* wd = nullptr;
  stat_t* stats = nullptr;
  int rc = CDM_OK;
  auto cleanup = [&]() { cudaFree(a); cudaFree(r); cudaFree(idn); cudaFree(o); cudaFree(wd); cudaFree(stats); };
#define DBG_OK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { cleanup(); return fail(CDM_ERR_CUDA, "%s: %s", #e, cudaGetErrorString(_e)); } } while (0)
#define DBG_TRY(e) do { rc = (e); if (rc != CDM_OK) { cleanup(); return rc; } } while (0)
  DBG_OK(cudaMalloc(&a, (size_t)B * HW * Cin * sizeof(T)));
  DBG_OK(cudaMalloc(&o, (size_t)B * HW * Cout * sizeof(T)));
  DBG_OK(cudaMalloc(&stats, (size_t)B * GN_GROUPS * 2 * sizeof(stat_t)));
  DBG_OK(cudaMemsetAsync(stats, 0, (size_t)B * GN_GROUPS * 2 * sizeof(stat_t), st));
  DBG_TRY(launch_nchw_to_nhwc<T>(x, a, B, HW, Cin, st));
  if (res) {
    DBG_OK(cudaMalloc(&r, (size_t)B * HW * Cres * sizeof(T)));
    DBG_TRY(launch_nchw_to_nhwc<T>(res, r, B, HW, Cres, st));
  }
  if (identity) {
    DBG_OK(cudaMalloc(&idn, (size_t)B * HW * Cout * sizeof(T)));
    DBG_TRY(launch_nchw_to_nhwc<T>(identity, idn, B, HW, Cout, st));
  }
  ConvArgs<T> c{};
  c.a = a; c.r = r; c.identity = idn; c.out = o; c.bias = bias; c.bias_stride = bias_rows > 1 ? Cout : 0;
  c.stats = stats_out ? stats : nullptr;
  c.B = B; c.H = H; c.W = W; c.Cin = Cin; c.Cres = Cres; c.Cout = Cout; c.taps = taps;
  if constexpr (sizeof(T) == 4) {
    if (variant == 3) {   // fp32-class tensor-core path: split the inputs into hi / lo planes, three MMAs per K step
      std::vector<h16> xh, xl;
      pack_conv_x3(w, Cout, Cin, taps, res ? &wres : nullptr, Cres, xh, xl);
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      h16* planes = nullptr;
      const size_t na = (size_t)B * HW * Cin, nr = res ? (size_t)B * HW * Cres : 0;
      DBG_OK(cudaMalloc(&wd, (2 * xh.size() + 2 * na + 2 * nr) * sizeof(h16)));
      h16* wh = (h16*)wd; h16* wl = wh + xh.size();
      planes = wl + xl.size();
      DBG_OK(cudaMemcpyAsync(wh, xh.data(), xh.size() * sizeof(h16), cudaMemcpyHostToDevice, st));
      DBG_OK(cudaMemcpyAsync(wl, xl.data(), xl.size() * sizeof(h16), cudaMemcpyHostToDevice, st));
      const X3Planes none{nullptr, nullptr};
      X3Planes ap{planes, planes + na}, rp = res ? X3Planes{planes + 2 * na, planes + 2 * na + nr} : none;
      DBG_TRY(launch_gn_silu_split((const float*)a, nullptr, nullptr, nullptr, ap, none, B, HW, Cin, st));
      if (res) DBG_TRY(launch_gn_silu_split((const float*)r, nullptr, nullptr, nullptr, rp, none, B, HW, Cres, st));
      DBG_TRY(launch_conv_x3(reinterpret_cast<const ConvArgs<float>&>(c), ap, rp, wh, wl, sms, st));
    } else {
    DBG_OK(cudaMalloc(&wd, kn.size() * sizeof(float)));
    DBG_OK(cudaMemcpyAsync(wd, kn.data(), kn.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    DBG_TRY(launch_conv_fp32(reinterpret_cast<const ConvArgs<float>&>(c), (const float*)wd, st));
    }
  } else {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (halo) pack_conv_halo(w, Cout, Cin, res ? &wres : nullptr, Cres, nk);
    if (stack) {
      if (!conv_stack3_supported(H, W, Cin, res ? Cres : 0, Cout, taps)) { cleanup(); return fail(CDM_ERR_UNSUPPORTED, "conv_stack3: unsupported shape"); }
      pack_conv_stack3(w, Cin, res ? &wres : nullptr, Cres, nk);
    }
    DBG_OK(cudaMalloc(&wd, nk.size() * sizeof(h16)));
    DBG_OK(cudaMemcpyAsync(wd, nk.data(), nk.size() * sizeof(h16), cudaMemcpyHostToDevice, st));
    if (stack) DBG_TRY(launch_conv_stack3(c, (const h16*)wd, sms, st));
    else if (halo) DBG_TRY(launch_conv_halo(c, (const h16*)wd, sms, st));
    else DBG_TRY(launch_conv_tc(c, (const h16*)wd, sms, st));
  }
  DBG_TRY(launch_nhwc_to_nchw<T>(o, out, B, HW, Cout, st));
  if (stats_out) DBG_TRY(launch_stats_to_float(stats, stats_out, B * GN_GROUPS * 2, st));
  DBG_OK(cudaStreamSynchronize(st));
  cleanup();
  return CDM_OK;
#undef DBG_OK
#undef DBG_TRY
}

// ---- grouped launches (group.cuh) ----------------------------------------------------------------------------------
// Launch what was recorded since group_begin(): ONE grouped launch when the K records name the same kernel instance and
// shared-memory size, else each record through a group of one (same kernels, blockIdx.y = 0).
int group_flush(int num_sms, cudaStream_t st) {
  GroupState& g = group_state();
  g.recording = false;
  const int K = g.n;
  g.n = 0;
  if (K == 0) return CDM_OK;
  bool same = true;
  for (int k = 1; k < K; ++k) same = same && g.rec[k].kind == g.rec[0].kind && g.rec[k].inst == g.rec[0].inst && g.rec[k].smem == g.rec[0].smem;
  auto launch = [&](const GroupRec* r, int n) -> int {
    if (r[0].kind == GK_HALO) return launch_halo_group(r, n, num_sms, st);
    if (r[0].kind == GK_STACK3) return launch_stack3_group(r, n, num_sms, st);
    return fail(CDM_ERR_INVALID, "group_flush: unknown kernel kind %d", r[0].kind);
  };
  if (same) return launch(g.rec, K);
  for (int k = 0; k < K; ++k) CDM_TRY(launch(&g.rec[k], 1));
  return CDM_OK;
}

// One general fp16 tensor-core convolution (conv_x3.cu, TERMS = 1) with torch layouts in and out
static int debug_conv_t16(const float* x1, const float* x2, const float* w_host, const float* bias, float* out, int B, int C1, int C2,
                          int Cout, int H, int W, int kind, int relu, cudaStream_t st) {
  auto pad64 = [](int c) { return (c + 63) / 64 * 64; };
  const int c1p = pad64(C1), c2p = C2 ? pad64(C2) : 0, cop = pad64(Cout);
  const int k = kind == CT16_K3 ? 3 : (kind == CT16_K1 ? 1 : 4);
  const int Ho = kind == CT16_K4S2 ? H / 2 : (kind == CT16_T4S2 ? 2 * H : H), Wo = kind == CT16_K4S2 ? W / 2 : (kind == CT16_T4S2 ? 2 * W : W);
  std::vector<float> w(w_host, w_host + (size_t)Cout * (C1 + C2) * k * k);
  std::vector<h16> pk;
  pack_conv_t16(w, Cout, C1, C2, kind, cop, c1p, c2p, pk);
  h16 *a1 = nullptr, *a2 = nullptr, *o = nullptr, *wd = nullptr;
  float *t1 = nullptr, *bp = nullptr;
  auto cleanup = [&]() { cudaFree(a1); cudaFree(a2); cudaFree(o); cudaFree(wd); cudaFree(t1); cudaFree(bp); };
#define T16_OK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { cleanup(); return fail(CDM_ERR_CUDA, "%s: %s", #e, cudaGetErrorString(_e)); } } while (0)
#define T16_TRY(e) do { int _rc = (e); if (_rc != CDM_OK) { cleanup(); return _rc; } } while (0)
  const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo;
  // zero-padded NHWC fp16 copies of the inputs: NCHW fp32 -> padded NCHW fp32 (memset + strided copy) -> NHWC fp16
  auto stage = [&](const float* src, int C, int Cp, h16** dst) -> int {
    T16_OK(cudaMalloc(dst, (size_t)B * HW * Cp * sizeof(h16)));
    cudaFree(t1); t1 = nullptr;
    T16_OK(cudaMalloc(&t1, (size_t)B * Cp * HW * sizeof(float)));
    T16_OK(cudaMemsetAsync(t1, 0, (size_t)B * Cp * HW * sizeof(float), st));
    T16_OK(cudaMemcpy2DAsync(t1, (size_t)Cp * HW * sizeof(float), src, (size_t)C * HW * sizeof(float), (size_t)C * HW * sizeof(float), B,
                             cudaMemcpyDeviceToDevice, st));
    T16_TRY(launch_nchw_to_nhwc<h16>(t1, *dst, B, (int)HW, Cp, st));
    return CDM_OK;
  };
  T16_TRY(stage(x1, C1, c1p, &a1));
  if (C2) T16_TRY(stage(x2, C2, c2p, &a2));
  T16_OK(cudaMalloc(&o, (size_t)B * HWo * cop * sizeof(h16)));
  T16_OK(cudaMalloc(&wd, pk.size() * sizeof(h16)));
  T16_OK(cudaMemcpyAsync(wd, pk.data(), pk.size() * sizeof(h16), cudaMemcpyHostToDevice, st));
  T16_OK(cudaMalloc(&bp, (size_t)cop * sizeof(float)));
  T16_OK(cudaMemsetAsync(bp, 0, (size_t)cop * sizeof(float), st));
  if (bias) T16_OK(cudaMemcpyAsync(bp, bias, (size_t)Cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  ConvT16 c{};
  c.a1 = a1; c.C1 = c1p; c.a2 = a2; c.C2 = c2p; c.out = o; c.B = B; c.H = H; c.W = W; c.Cout = cop; c.kind = kind; c.w = wd; c.bias = bp;
  c.relu = relu;
  T16_TRY(launch_conv_t16(c, sms, st));
  cudaFree(t1); t1 = nullptr;
  T16_OK(cudaMalloc(&t1, (size_t)B * cop * HWo * sizeof(float)));
  T16_TRY(launch_nhwc_to_nchw<h16>(o, t1, B, (int)HWo, cop, st));
  T16_OK(cudaMemcpy2DAsync(out, (size_t)Cout * HWo * sizeof(float), t1, (size_t)cop * HWo * sizeof(float), (size_t)Cout * HWo * sizeof(float), B,
                           cudaMemcpyDeviceToDevice, st));
  T16_OK(cudaStreamSynchronize(st));
  cleanup();
  return CDM_OK;
#undef T16_OK
#undef T16_TRY
}

// Test hooks for the two elementwise producers of the UNet graphs (torch layouts in and out): 2x2 max pool and
// bilinear x2 upsample (align_corners = True) + channel concat, each with the GroupNorm statistics it accumulates.
template <typename T>
static int debug_maxpool_t(const float* x, float* out, float* stats_out, float* stats_in_out, int B, int C, int H, int W, cudaStream_t st) {
  const size_t HW = (size_t)H * W, HWo = HW / 4, ns = (size_t)B * GN_GROUPS * 2;
  T *a = nullptr, *o = nullptr; stat_t* sd = nullptr;
  auto cleanup = [&]() { cudaFree(a); cudaFree(o); cudaFree(sd); };
#define EW_OK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { cleanup(); return fail(CDM_ERR_CUDA, "%s: %s", #e, cudaGetErrorString(_e)); } } while (0)
#define EW_TRY(e) do { int _rc = (e); if (_rc != CDM_OK) { cleanup(); return _rc; } } while (0)
  EW_OK(cudaMalloc(&a, (size_t)B * HW * C * sizeof(T)));
  EW_OK(cudaMalloc(&o, (size_t)B * HWo * C * sizeof(T)));
  EW_OK(cudaMalloc(&sd, 2 * ns * sizeof(stat_t)));
  EW_OK(cudaMemsetAsync(sd, 0, 2 * ns * sizeof(stat_t), st));
  EW_TRY(launch_nchw_to_nhwc<T>(x, a, B, (int)HW, C, st));
  EW_TRY(launch_maxpool_stats<T>(a, o, stats_out ? sd : nullptr, B, H, W, C, st, stats_in_out ? sd + ns : nullptr));
  EW_TRY(launch_nhwc_to_nchw<T>(o, out, B, (int)HWo, C, st));
  if (stats_out) EW_TRY(launch_stats_to_float(sd, stats_out, (int)ns, st));
  if (stats_in_out) EW_TRY(launch_stats_to_float(sd + ns, stats_in_out, (int)ns, st));
  EW_OK(cudaStreamSynchronize(st));
  cleanup();
  return CDM_OK;
}

template <typename T>
static int debug_upcat_t(const float* low, const float* skip, float* out, float* stats_out, int B, int Ca, int Cs, int h, int w,
                         int virt, cudaStream_t st) {
  const int H = 2 * h, W = 2 * w, Co = virt ? Ca : Ca + Cs;
  const size_t hw = (size_t)h * w, HW = (size_t)H * W, ns = (size_t)B * GN_GROUPS * 2;
  T *l = nullptr, *s = nullptr, *o = nullptr, *junk = nullptr; stat_t* sd = nullptr;
  auto cleanup = [&]() { cudaFree(l); cudaFree(s); cudaFree(o); cudaFree(junk); cudaFree(sd); };
  if (virt && !upcat_virtual_supported(Ca, Cs)) return fail(CDM_ERR_UNSUPPORTED, "cdm_debug_upcat: no virtual concat for Ca=%d Cs=%d", Ca, Cs);
  EW_OK(cudaMalloc(&l, (size_t)B * hw * Ca * sizeof(T)));
  EW_OK(cudaMalloc(&s, (size_t)B * HW * Cs * sizeof(T)));
  EW_OK(cudaMalloc(&o, (size_t)B * HW * Co * sizeof(T)));
  EW_OK(cudaMalloc(&sd, 2 * ns * sizeof(stat_t)));
  EW_OK(cudaMemsetAsync(sd, 0, 2 * ns * sizeof(stat_t), st));
  EW_TRY(launch_nchw_to_nhwc<T>(low, l, B, (int)hw, Ca, st));
  EW_TRY(launch_nchw_to_nhwc<T>(skip, s, B, (int)HW, Cs, st));
  if (virt) {
    // the skip tensor's own {sum, sumsq} per Cs/8-channel group, produced the way the graph does: by the max pool that reads it
    EW_OK(cudaMalloc(&junk, (size_t)B * (HW / 4) * Cs * sizeof(T)));
    EW_TRY(launch_maxpool_stats<T>(s, junk, nullptr, B, H, W, Cs, st, sd + ns));
  }
  EW_TRY(launch_upcat_stats<T>(l, s, o, stats_out ? sd : nullptr, B, h, w, Ca, Cs, st, virt ? sd + ns : nullptr));
  EW_TRY(launch_nhwc_to_nchw<T>(o, out, B, (int)HW, Co, st));
  if (stats_out) EW_TRY(launch_stats_to_float(sd, stats_out, (int)ns, st));
  EW_OK(cudaStreamSynchronize(st));
  cleanup();
  return CDM_OK;
#undef EW_OK
#undef EW_TRY
}

}  // namespace cdm

extern "C" {

int cdm_debug_maxpool(const float* x, float* out, float* stats_out, float* stats_in_out, int B, int C, int H, int W, int precision,
                      void* stream) {
  if (!x || !out) return fail(CDM_ERR_INVALID, "cdm_debug_maxpool: null argument");
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(CDM_ERR_INVALID, "cdm_debug_maxpool: empty tensor");
  if (precision == CDM_PREC_FP32) return debug_maxpool_t<float>(x, out, stats_out, stats_in_out, B, C, H, W, (cudaStream_t)stream);
  if (precision == CDM_PREC_F16) return debug_maxpool_t<h16>(x, out, stats_out, stats_in_out, B, C, H, W, (cudaStream_t)stream);
  return fail(CDM_ERR_INVALID, "cdm_debug_maxpool: precision %d", precision);
}

int cdm_debug_upcat(const float* low, const float* skip, float* out, float* stats_out, int B, int Ca, int Cs, int h, int w,
                    int precision, int virtual_concat, void* stream) {
  if (!low || !skip || !out) return fail(CDM_ERR_INVALID, "cdm_debug_upcat: null argument");
  if (B <= 0 || Ca <= 0 || Cs <= 0 || h <= 0 || w <= 0) return fail(CDM_ERR_INVALID, "cdm_debug_upcat: empty tensor");
  if (precision == CDM_PREC_FP32) return debug_upcat_t<float>(low, skip, out, stats_out, B, Ca, Cs, h, w, virtual_concat, (cudaStream_t)stream);
  if (precision == CDM_PREC_F16) return debug_upcat_t<h16>(low, skip, out, stats_out, B, Ca, Cs, h, w, virtual_concat, (cudaStream_t)stream);
  return fail(CDM_ERR_INVALID, "cdm_debug_upcat: precision %d", precision);
}

int cdm_debug_conv_t16(const float* x1, const float* x2, const float* w_host, const float* bias, float* out, int B, int C1, int C2,
                       int Cout, int H, int W, int kind, int relu, void* stream) {
  if (!x1 || !w_host || !out) return fail(CDM_ERR_INVALID, "cdm_debug_conv_t16: null argument");
  if (kind < 0 || kind > 3) return fail(CDM_ERR_INVALID, "cdm_debug_conv_t16: kind %d", kind);
  if ((C2 != 0) != (x2 != nullptr)) return fail(CDM_ERR_INVALID, "cdm_debug_conv_t16: x2 and C2 go together");
  return debug_conv_t16(x1, x2, w_host, bias, out, B, C1, C2, Cout, H, W, kind, relu, (cudaStream_t)stream);
}

int cdm_debug_init_conv(const float* x, const float* w, const float* bias, float* out, float* stats_out, int B, int Cin, int H, int W,
                        int tensor_core, void* stream) {
  if (!x || !w || !out) return fail(CDM_ERR_INVALID, "cdm_debug_init_conv: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t HW = (size_t)H * W;
  h16* o = nullptr; stat_t* sd = nullptr;
  auto cleanup = [&]() { cudaFree(o); cudaFree(sd); };
#define IC_OK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { cleanup(); return fail(CDM_ERR_CUDA, "%s: %s", #e, cudaGetErrorString(_e)); } } while (0)
#define IC_TRY(e) do { int _rc = (e); if (_rc != CDM_OK) { cleanup(); return _rc; } } while (0)
  IC_OK(cudaMalloc(&o, (size_t)B * HW * 64 * sizeof(h16)));
  IC_OK(cudaMalloc(&sd, (size_t)B * GN_GROUPS * 2 * sizeof(stat_t)));
  IC_OK(cudaMemsetAsync(sd, 0, (size_t)B * GN_GROUPS * 2 * sizeof(stat_t), st));
  set_init_conv_tc(tensor_core ? 1 : 0);
  int rc = CDM_OK;
  if (tensor_core && !init_conv_tc_supported(Cin, H, W, 64, nullptr)) rc = fail(CDM_ERR_UNSUPPORTED, "cdm_debug_init_conv: no tcgen05 instance for Cin=%d %dx%d", Cin, H, W);
  else rc = launch_init_conv<h16>(x, w, bias, o, stats_out ? sd : nullptr, B, Cin, H, W, 64, st);
  set_init_conv_tc(-1);
  IC_TRY(rc);
  IC_TRY(launch_nhwc_to_nchw<h16>(o, out, B, (int)HW, 64, st));
  if (stats_out) IC_TRY(launch_stats_to_float(sd, stats_out, B * GN_GROUPS * 2, st));
  IC_OK(cudaStreamSynchronize(st));
  cleanup();
  return CDM_OK;
#undef IC_OK
#undef IC_TRY
}

int cdm_abi_version(void) { return CDM_ABI_VERSION; }

#ifndef CDM_ABI_STAMP
#define CDM_ABI_STAMP 0u
#endif
unsigned cdm_abi_stamp(void) { return CDM_ABI_STAMP; }

long long cdm_launch_count(void) { return prof_state().launches.load(); }

int cdm_prof_enable(int on) {
  ProfState& p = prof_state();
  for (auto& r : p.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  p.recs.clear();
  p.enabled = on != 0;
  return CDM_OK;
}

int cdm_prof_summary(cdm_prof_entry* out, int max_entries) {
  static const char* names[KC_COUNT] = {"step", "temb", "init_conv", "gn_silu", "maxpool", "upcat", "out_conv",
                                        "conv_fp32", "conv_tc", "mlp", "misc"};
  if (!out || max_entries < KC_COUNT) return fail(CDM_ERR_INVALID, "cdm_prof_summary: need room for %d entries", (int)KC_COUNT);
  ProfState& p = prof_state();
  for (int i = 0; i < KC_COUNT; ++i) {
    memset(&out[i], 0, sizeof(out[i]));
    strncpy(out[i].name, names[i], sizeof(out[i].name) - 1);
  }
  for (auto& r : p.recs) {
    CDM_CUDA_OK(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    CDM_CUDA_OK(cudaEventElapsedTime(&ms, r.e0, r.e1));
    out[r.kc].launches += 1; out[r.kc].ms += ms; out[r.kc].flops += r.flops; out[r.kc].bytes += r.bytes;
  }
  return KC_COUNT;
}

int cdm_prof_dump(void) {
  static const char* names[KC_COUNT] = {"step", "temb", "init_conv", "gn_silu", "maxpool", "upcat", "out_conv",
                                        "conv_fp32", "conv_tc", "mlp", "misc"};
  ProfState& p = prof_state();
  for (auto& r : p.recs) {
    CDM_CUDA_OK(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    CDM_CUDA_OK(cudaEventElapsedTime(&ms, r.e0, r.e1));
    fprintf(stderr, "[prof] %-10s %-44s %8.4f ms %8.1f TFLOP/s %8.1f GB/s\n", names[r.kc], r.tag, ms,
            ms > 0 ? r.flops / ms * 1e-9 : 0.0, ms > 0 ? r.bytes / ms * 1e-6 : 0.0);
  }
  return (int)p.recs.size();
}

const char* cdm_last_error(void) { return last_error_ref().c_str(); }

int cdm_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(CDM_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(CDM_ERR_INVALID, "device %d out of range (have %d)", device, n);
  cudaDeviceProp prop;
  CDM_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(CDM_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  return CDM_OK;
}

int cdm_debug_conv(const float* x, const float* w_host, const float* bias, int bias_rows, const float* res,
                   const float* wres_host, const float* identity, float* out, float* stats_out, int B, int Cin,
                   int Cres, int Cout, int H, int W, int taps, int precision, void* stream) {
  if (!x || !w_host || !bias || !out) return fail(CDM_ERR_INVALID, "cdm_debug_conv: null argument");
  if ((res == nullptr) != (wres_host == nullptr)) return fail(CDM_ERR_INVALID, "cdm_debug_conv: res and wres_host go together");
  if (taps != 9 && taps != 1) return fail(CDM_ERR_INVALID, "cdm_debug_conv: taps=%d", taps);
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == CDM_PREC_FP32)
    return debug_conv_t<float>(x, w_host, bias, bias_rows, res, wres_host, identity, out, stats_out, B, Cin, Cres, Cout, H, W, taps, st);
  if (precision == CDM_PREC_F16)
    return debug_conv_t<h16>(x, w_host, bias, bias_rows, res, wres_host, identity, out, stats_out, B, Cin, Cres, Cout, H, W, taps, st);
  if (precision == 2)   // fp16, halo-tile kernel (conv_tc2.cu)
    return debug_conv_t<h16>(x, w_host, bias, bias_rows, res, wres_host, identity, out, stats_out, B, Cin, Cres, Cout, H, W, taps, st, 1);
  if (precision == CDM_PREC_F16X3)   // three-term split-fp16 kernel (conv_x3.cu), fp32 activations
    return debug_conv_t<float>(x, w_host, bias, bias_rows, res, wres_host, identity, out, stats_out, B, Cin, Cres, Cout, H, W, taps, st, 3);
  if (precision == 3)   // fp16, stacked halo-tile kernel (conv_tc3.cu)
    return debug_conv_t<h16>(x, w_host, bias, bias_rows, res, wres_host, identity, out, stats_out, B, Cin, Cres, Cout, H, W, taps, st, 2);
  return fail(CDM_ERR_INVALID, "cdm_debug_conv: precision %d", precision);
}

}  // extern "C"
