// pnb_engine.cu -- host side of the C-ABI (include/percepnet_b200.h): engine lifetime, weight ingest
// from the reference's RNNModel layout, constant tables, and the per-call kernel schedule.
//
// Schedule of one pnb_process_* call with F hops on S streams:
//   stage_in -> analysis (all F hops; one warp per stream) -> F x network step -> synthesis -> slide history
// The network step is the only part that is sequential in time across the batch; analysis and synthesis
// need no network state, so hops are processed F at a time.
#include <cuda.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stddef.h>
#include <string>
#include <vector>

#include "../../include/percepnet_b200.h"
#include "pnb_engine.h"

using namespace pnb;

static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) return fail(PNB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

int tc_fail(int code, const char *msg) { return fail(code, "%s", msg); }

extern "C" const char *pnb_last_error(void) { return g_err.c_str(); }
extern "C" const char *pnb_version(void) { return "percepnet_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------
// constant tables, with the reference's own expressions so that they round identically
// ------------------------------------------------------------------------------------------
static void build_tables(Tables &t) {
  memset(&t, 0, sizeof t);
  for (int i = 0; i < kFrame; i++) {  // denoise.cpp:191-192
    double a = .5 * M_PI * (i + .5) / kFrame;
    t.half_window[i] = (float)sin(.5 * M_PI * sin(a) * sin(a));
  }
  {  // denoise.cpp:200-206: float running sum
    float acc = 0;
    for (int i = 1; i < 8; i++) {
      t.comb_w[i - 1] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / 8));
      acc += t.comb_w[i - 1];
    }
    for (int i = 0; i < 7; i++) t.comb_w[i] /= acc;
  }
  for (int i = 0; i < kWin; i++) {  // kiss_fft.cpp:415-419
    const double pi = 3.14159265358979323846264338327;
    double ph = (-2 * pi / kWin) * i;
    t.tw[i].x = (float)cos(ph);
    t.tw[i].y = (float)sin(ph);
  }
  {  // erbband.h:34-99 with (960, 32, 0, 20000), denoise.cpp:87
    auto hz2erb = [](float hz) { return (float)(9.265 * log(1 + hz / (24.7 * 9.265))); };
    auto erb2hz = [](float e) { return (float)(24.7 * 9.265 * (exp(e / 9.265) - 1)); };
    float lo = hz2erb(0.f), hi = hz2erb(20000.f), step = (hi - lo) / (34.f - 1);
    int border[kBands];
    for (int i = 0; i < kBands; i++) {
      float cut = erb2hz(i < kBands - 1 ? lo + step * i : hi);
      border[i] = (int)((cut + 25) / 50.f);
    }
    for (int i = 0; i < kBands - 2; i++)
      if (border[i + 1] - border[i] < 2) border[i + 1] += 2 - (border[i + 1] - border[i]);
    for (int i = 0; i < kBands; i++) t.border[i] = (short)border[i];
    t.border[kBands] = t.border[kBands + 1] = (short)border[kBands - 1];
    for (int b = 0; b < kBands - 1; b++) {
      int width = border[b + 1] - border[b];
      for (int j = 0; j < width; j++) {
        int bin = border[b] + j;
        if (bin >= kBins) continue;
        float frac = (float)j / width;  // denoise.cpp:99
        t.frac[bin] = frac;
        t.omf[bin] = 1 - frac;
        t.band_of[bin] = (short)b;
      }
    }
  }
  // tansig_table.h: tanh(0.04 i) to six decimals; the three entries where the shipped table is not the
  // correctly rounded value carry the shipped value
  for (int i = 0; i <= 200; i++) t.tansig[i] = (float)(floor(tanh(0.04 * i) * 1e6 + 0.5) / 1e6);
  t.tansig[70] = 0.992631f;
  t.tansig[170] = 0.999997f;
  t.tansig[190] = 1.000000f;
}

// ------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t dalloc(T **p, size_t n, bool zero = true) {
  cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
  if (e != cudaSuccess) return e;
  if (zero) return cudaMemset(*p, 0, n * sizeof(T));
  return cudaSuccess;
}
static cudaError_t upload(float **dst, const float *src, size_t n) {
  cudaError_t e = cudaMalloc((void **)dst, n * sizeof(float));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice);
}

static int check_model(const pnb_model *m) {
  if (!m || !m->fc || !m->conv1 || !m->conv2 || !m->gru1 || !m->gru2 || !m->gru3 || !m->gru_gb || !m->gru_rb ||
      !m->fc_gb || !m->fc_rb)
    return fail(PNB_ERR_ARG, "model or one of its ten layers is NULL");
  // the architecture the hot path is built for (rnn_train.py:111-121, rnn.cpp:42-81)
  bool ok = m->fc->nb_inputs == 70 && m->fc->nb_neurons == 128 && m->conv1->nb_inputs == 128 &&
            m->conv1->kernel_size == 5 && m->conv1->nb_neurons == 512 && m->conv2->nb_inputs == 512 &&
            m->conv2->kernel_size == 3 && m->conv2->nb_neurons == 512 && m->gru_rb->nb_inputs == 1024 &&
            m->gru_rb->nb_neurons == 128 && m->fc_gb->nb_inputs == 2560 && m->fc_gb->nb_neurons == 34 &&
            m->fc_rb->nb_inputs == 128 && m->fc_rb->nb_neurons == 34;
  const pnb_gru_layer *g4[4] = {m->gru1, m->gru2, m->gru3, m->gru_gb};
  for (int i = 0; i < 4; i++) ok = ok && g4[i]->nb_inputs == 512 && g4[i]->nb_neurons == 512;
  const pnb_gru_layer *g5[5] = {m->gru1, m->gru2, m->gru3, m->gru_gb, m->gru_rb};
  for (int i = 0; i < 5; i++) ok = ok && g5[i]->reset_after == 1 && g5[i]->activation == PNB_ACT_TANH;
  if (!ok) return fail(PNB_ERR_ARG, "model dimensions are not PercepNet's (70-128-conv5x512-conv3x512-4xGRU512-GRU128-34/34)");
  // the kernels fuse each layer's activation (rnn_train.py:111-121 / dump_percepnet.py); a model that asks for
  // anything else would silently compute something the reference does not
  if (m->fc->activation != PNB_ACT_RELU || m->conv1->activation != PNB_ACT_RELU || m->conv2->activation != PNB_ACT_TANH ||
      m->fc_gb->activation != PNB_ACT_SIGMOID || m->fc_rb->activation != PNB_ACT_SIGMOID)
    return fail(PNB_ERR_ARG, "model activations are not PercepNet's (fc relu, conv1 relu, conv2 tanh, fc_gb / fc_rb sigmoid)");
  return PNB_OK;
}

constexpr int kMaxChunks = 512;  // chunks one call may be cut into (chunked overlap schedule)

// ------------------------------------------------------------------------------------------
// SM partition (driver green contexts, reached through the runtime's entry-point query: no libcuda link)
// ------------------------------------------------------------------------------------------
namespace {
struct GreenApi {
  CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
  CUresult (*DevSmResourceSplitByCount)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
  CUresult (*GreenCtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*GreenCtxStreamCreate)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
  CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
  bool ok = false;
};
const GreenApi &green_api() {
  static GreenApi g = [] {
    GreenApi a;
    auto get = [](const char *name, void **fn) {
      cudaDriverEntryPointQueryResult qr;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &qr) == cudaSuccess && *fn != nullptr;
    };
    a.ok = get("cuDeviceGet", (void **)&a.DeviceGet) && get("cuDeviceGetDevResource", (void **)&a.DeviceGetDevResource) &&
           get("cuDevSmResourceSplitByCount", (void **)&a.DevSmResourceSplitByCount) &&
           get("cuDevResourceGenerateDesc", (void **)&a.DevResourceGenerateDesc) && get("cuGreenCtxCreate", (void **)&a.GreenCtxCreate) &&
           get("cuGreenCtxStreamCreate", (void **)&a.GreenCtxStreamCreate) && get("cuGreenCtxDestroy", (void **)&a.GreenCtxDestroy);
    return a;
  }();
  return g;
}
}  // namespace

// Splits the device's SMs into a share for the network kernels and the rest for the DSP kernels and creates one stream
// in each.  Best effort: on any failure the engine simply keeps the serial schedule (net_sms stays 0).
static void setup_overlap(pnb_engine *e, int want_net_sms) {
  const GreenApi &G = green_api();
  if (!G.ok) return;
  CUdevice dev;
  CUdevResource all, grp, rem;
  unsigned int ngrp = 1;
  CUdevResourceDesc d_net = nullptr, d_dsp = nullptr;
  CUgreenCtx g_net = nullptr, g_dsp = nullptr;
  CUstream s_net = nullptr, s_dsp = nullptr, s_syn = nullptr;
  if (G.DeviceGet(&dev, e->device) != CUDA_SUCCESS) return;
  if (G.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return;
  // default granularity: 8 SMs (keeps whole GPCs' cluster scheduling); PNB_SPLIT_FINE=1 asks for 2-SM (TPC) granularity,
  // enough for the CTA pairs of the network kernels
  const char *fine = getenv("PNB_SPLIT_FINE");
  const unsigned split_flags = (fine && atoi(fine) != 0) ? CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING : 0;
  if (G.DevSmResourceSplitByCount(&grp, &ngrp, &all, &rem, split_flags, (unsigned)want_net_sms) != CUDA_SUCCESS || ngrp != 1) return;
  if (grp.sm.smCount < 8 || rem.sm.smCount < 8) return;
  if (G.DevResourceGenerateDesc(&d_net, &grp, 1) != CUDA_SUCCESS || G.DevResourceGenerateDesc(&d_dsp, &rem, 1) != CUDA_SUCCESS) return;
  if (G.GreenCtxCreate(&g_net, d_net, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return;
  if (G.GreenCtxCreate(&g_dsp, d_dsp, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) { G.GreenCtxDestroy(g_net); return; }
  if (G.GreenCtxStreamCreate(&s_net, g_net, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS ||
      G.GreenCtxStreamCreate(&s_dsp, g_dsp, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS ||
      G.GreenCtxStreamCreate(&s_syn, g_dsp, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) {
    if (s_net) cudaStreamDestroy((cudaStream_t)s_net);
    if (s_dsp) cudaStreamDestroy((cudaStream_t)s_dsp);
    G.GreenCtxDestroy(g_net); G.GreenCtxDestroy(g_dsp);
    return;
  }
  bool ev_ok = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&e->ev_join_net, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&e->ev_join_dsp, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&e->ev_join_syn, cudaEventDisableTiming) == cudaSuccess;
  const int nchunks = (e->Fmax + e->chunk - 1) / e->chunk + 5;  // + the short chunks at both ends of a call
  e->ev_ana.assign(nchunks, nullptr);
  e->ev_net.assign(nchunks, nullptr);
  e->ev_syn[0].assign(nchunks, nullptr);
  e->ev_syn[1].assign(nchunks, nullptr);
  for (int k = 0; k < nchunks && ev_ok; k++)
    ev_ok = cudaEventCreateWithFlags(&e->ev_ana[k], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_net[k], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_syn[0][k], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_syn[1][k], cudaEventDisableTiming) == cudaSuccess;
  e->green_net = g_net; e->green_dsp = g_dsp;
  e->s_net = (cudaStream_t)s_net; e->s_dsp = (cudaStream_t)s_dsp; e->s_syn = (cudaStream_t)s_syn;
  if (!ev_ok) return;  // net_sms stays 0: the serial schedule; pnb_destroy releases what was created
  e->net_sms = (int)grp.sm.smCount;
  e->dsp_sms = (int)rem.sm.smCount;
}

extern "C" int pnb_create(pnb_engine **out, int n_streams, int max_frames, const pnb_model *model, unsigned flags,
                          int device) {
  if (!out) return fail(PNB_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (n_streams < 1 || max_frames < 1) return fail(PNB_ERR_ARG, "n_streams and max_frames_per_call must be >= 1");
  if ((long long)n_streams * max_frames > (1ll << 24))
    return fail(PNB_ERR_ARG, "n_streams x max_frames_per_call above 2^24 rows per call: split the batch across engines");
  const bool train_mode = (flags & PNB_TRAIN_DATA) != 0;
  if (train_mode && (n_streams & 1)) return fail(PNB_ERR_ARG, "PNB_TRAIN_DATA needs n_streams = 2 x pairs");
  if ((flags & PNB_CONV_WIDE) && !(flags & PNB_NN_TENSOR)) return fail(PNB_ERR_ARG, "PNB_CONV_WIDE applies to PNB_NN_TENSOR engines");
  if (train_mode && (flags & (PNB_NN_TENSOR | PNB_POSTFILTER)))
    return fail(PNB_ERR_ARG, "PNB_TRAIN_DATA runs no network: combine it only with PNB_KEEP_TAPS");
  int rc = train_mode ? 0 : check_model(model);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(PNB_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU path",
                ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (device == -1) {  // the calling thread's current device
    ce = cudaGetDevice(&device);
    if (ce != cudaSuccess) return fail(PNB_ERR_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(ce));
  }
  if (device < 0 || device >= ndev) return fail(PNB_ERR_ARG, "device %d out of range (have %d)", device, ndev);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(PNB_ERR_NO_DEVICE, "device %d is sm_%d%d; this build carries sm_100a code only", device, prop.major,
                prop.minor);

  pnb_engine *e = new pnb_engine();
  e->S = n_streams;
  e->Fmax = max_frames;
  e->device = device;
  e->flags = flags;
  e->sm_count = prop.multiProcessorCount;
  const size_t S = n_streams, F = max_frames;
#define CKD(call)                                                                                   \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      int _rc = fail(_e == cudaErrorMemoryAllocation ? PNB_ERR_ALLOC : PNB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); \
      pnb_destroy(e);                                                                               \
      return _rc;                                                                                   \
    }                                                                                               \
  } while (0)

  CKD(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CKD(dsp_configure());
  {
    Tables *t = new Tables();
    build_tables(*t);
    cudaError_t r = cudaMalloc((void **)&e->d_tab, sizeof(Tables));
    if (r == cudaSuccess) r = cudaMemcpy(e->d_tab, t, sizeof(Tables), cudaMemcpyHostToDevice);
    delete t;
    CKD(r);
  }
  // weights, in the reference's layout
  if (!train_mode) {
  CKD(upload(&e->fc.W, model->fc->input_weights, 70 * 128));
  CKD(upload(&e->fc.b, model->fc->bias, 128));
  CKD(upload(&e->conv1.W, model->conv1->input_weights, 5 * 128 * 512));
  CKD(upload(&e->conv1.b, model->conv1->bias, 512));
  CKD(upload(&e->conv2.W, model->conv2->input_weights, 3 * 512 * 512));
  CKD(upload(&e->conv2.b, model->conv2->bias, 512));
  const pnb_gru_layer *g5[5] = {model->gru1, model->gru2, model->gru3, model->gru_gb, model->gru_rb};
  for (int i = 0; i < 5; i++) {
    int M = g5[i]->nb_inputs, H = g5[i]->nb_neurons;
    e->gru[i].M = M;
    e->gru[i].H = H;
    CKD(upload(&e->gru[i].W, g5[i]->input_weights, (size_t)M * 3 * H));
    CKD(upload(&e->gru[i].U, g5[i]->recurrent_weights, (size_t)H * 3 * H));
    CKD(upload(&e->gru[i].b, g5[i]->bias, 6 * (size_t)H));
  }
  CKD(upload(&e->fc_gb.W, model->fc_gb->input_weights, 2560 * 34));
  CKD(upload(&e->fc_gb.b, model->fc_gb->bias, 34));
  CKD(upload(&e->fc_rb.W, model->fc_rb->input_weights, 128 * 34));
  CKD(upload(&e->fc_rb.b, model->fc_rb->bias, 34));
  e->act_fc = model->fc->activation;
  e->act_conv1 = model->conv1->activation;
  e->act_conv2 = model->conv2->activation;
  e->act_gb = model->fc_gb->activation;
  e->act_rb = model->fc_rb->activation;
  }

  // stream state
  // room for several calls' hops behind the history (advance_line): eight calls when calls are short, fewer
  // when one call already carries many hops (the move is then amortised over those hops anyway)
  const size_t line_calls = F >= 128 ? 1 : (128 / F > (size_t)kLineCalls ? (size_t)kLineCalls : 128 / F);
  e->pcm_stride = kKeep + line_calls * F * kFrame;
  CKD(dalloc(&e->d_pcm, S * e->pcm_stride));
  CKD(dalloc(&e->d_synth, S * kFrame));
  CKD(dalloc(&e->d_status, 1));
  CKD(cudaHostAlloc((void **)&e->h_status, sizeof(int), cudaHostAllocDefault));
  *e->h_status = 0;
  CKD(dalloc(&e->d_last_period, S));
  CKD(dalloc(&e->d_last_gain, S));
  // per-call buffers
  CKD(dalloc(&e->d_feat, F * S * kFeat));
  e->ring = (int)F + 5;
  CKD(dalloc(&e->d_zring, (size_t)e->ring * S * kBins));
  CKD(dalloc(&e->d_ering, (size_t)e->ring * S * kBands));
  if (!train_mode) CKD(dalloc(&e->d_P, F * S * kBins));
  CKD(dalloc(&e->d_Ex, F * S * kBands));
  CKD(dalloc(&e->d_sil, F * S));
  if (!train_mode) CKD(dalloc(&e->d_gr, F * S * 68));
  if (train_mode) CKD(dalloc(&e->d_raw, F * S * 68));
  if (flags & PNB_KEEP_TAPS) {
    CKD(dalloc(&e->d_tap_pitch, F * S * 4));
    CKD(dalloc(&e->d_tap_pitchf, F * S * 2));
    if (!train_mode) CKD(dalloc(&e->d_tap_g, F * S * kBands));
  }
  if (train_mode) {
    CKD(cudaDeviceSynchronize());
    *out = e;
    return PNB_OK;
  }
  // network state; the conv rings and gate sums are scratch of the fp32 path only
  CKD(dalloc(&e->c2, S * 512));
  if (flags & PNB_NN_TENSOR) {
    for (int i = 0; i < 5; i++)
      for (int p = 0; p < 2; p++) CKD(dalloc(&e->h[i][p], S * e->gru[i].H));
  } else {
    // fp32 path: hop-slot buffers of one chunk of at most f32_chunk = min(max_frames, 32) hops.  Slot 0.. of the fc / conv1 buffers hold
    // the conv histories (oldest first), slot 0 of every state buffer the state before the chunk's first hop; the
    // chunk's carry moves the last slots back there, so between calls the state lives at the front (par stays 0).
    e->f32_chunk = max_frames < kF32ChainMaxHops ? max_frames : kF32ChainMaxHops;
    const size_t C = e->f32_chunk;
    e->f32_rt = S <= 64 ? 1 : 8;
    e->f32_rb = (int)((S + 16 * e->f32_rt - 1) / (16 * e->f32_rt));
    CKD(dalloc(&e->ring_fc, (C + 4) * S * 128));
    CKD(dalloc(&e->ring_c1, (C + 2) * S * 512));
    CKD(dalloc(&e->c2_all, C * S * 512));
    for (int i = 0; i < 5; i++) CKD(dalloc(&e->h[i][0], (C + 1) * S * e->gru[i].H));
    CKD(dalloc(&e->f32_cnt, (size_t)5 * e->f32_rb));
  }
  e->tc_sms = e->sm_count;
  if (flags & PNB_NN_TENSOR) {
    int trc = tc_prepare(e, model);
    if (trc) { pnb_destroy(e); return trc; }
    // Calls of at least two chunks overlap the DSP kernels with the network on disjoint SMs (pnb_engine.h).
    // PNB_OVERLAP=0 keeps the serial schedule; PNB_NET_SMS / PNB_CHUNK tune the split and the chunk length.
    const char *ov = getenv("PNB_OVERLAP"), *ns = getenv("PNB_NET_SMS"), *ch = getenv("PNB_CHUNK");
    if (ch && atoi(ch) >= 1) e->chunk = atoi(ch);
    const char *rp = getenv("PNB_RAMP");  // 0: all chunks equal (submitted calls overlap each other, so a call's ends need not be short)
    if (rp) e->ramp = atoi(rp) != 0;
    const bool worth = (long long)n_streams >= 2048 && max_frames >= 2 * e->chunk && (max_frames + e->chunk - 1) / e->chunk + 5 <= kMaxChunks;
    if (!(ov && atoi(ov) == 0) && (worth || (ov && atoi(ov) == 2)) && max_frames >= 2 * e->chunk)
      setup_overlap(e, ns && atoi(ns) >= 8 ? atoi(ns) : (e->sm_count * 7 / 16 / 8) * 8);
  }
  CKD(cudaDeviceSynchronize());
  *out = e;
  return PNB_OK;
}

extern "C" void pnb_destroy(pnb_engine *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  tc_release(e);
  float *fl[] = {e->fc.W, e->fc.b, e->conv1.W, e->conv1.b, e->conv2.W, e->conv2.b, e->fc_gb.W, e->fc_gb.b,
                 e->fc_rb.W, e->fc_rb.b, e->d_pcm, e->d_synth, e->d_last_gain, e->d_feat, e->d_Ex, e->d_gr,
                 e->d_tap_pitchf, e->d_tap_g, e->ring_fc, e->ring_c1, e->c2, e->c2_all, e->d_hin, e->d_hout,
                 e->d_raw, e->d_records};
  for (float *p : fl) if (p) cudaFree(p);
  for (int i = 0; i < 5; i++) {
    if (e->gru[i].W) cudaFree(e->gru[i].W);
    if (e->gru[i].U) cudaFree(e->gru[i].U);
    if (e->gru[i].b) cudaFree(e->gru[i].b);
    for (int p = 0; p < 2; p++) if (e->h[i][p]) cudaFree(e->h[i][p]);
  }
  if (e->f32_cnt) cudaFree(e->f32_cnt);
  if (e->d_zring) cudaFree(e->d_zring);
  if (e->d_ering) cudaFree(e->d_ering);
  if (e->d_P) cudaFree(e->d_P);
  if (e->d_sil) cudaFree(e->d_sil);
  if (e->d_last_period) cudaFree(e->d_last_period);
  if (e->d_status) cudaFree(e->d_status);
  if (e->h_status) cudaFreeHost(e->h_status);
  if (e->d_tap_pitch) cudaFree(e->d_tap_pitch);
  if (e->d_tab) cudaFree(e->d_tab);
  if (e->d_hin16) cudaFree(e->d_hin16);
  if (e->d_hout16) cudaFree(e->d_hout16);
  for (int k = 0; k < 2; k++) {
    if (e->pipe_in[k]) cudaFree(e->pipe_in[k]);
    if (e->pipe_out[k]) cudaFree(e->pipe_out[k]);
    if (e->ev_in[k]) cudaEventDestroy(e->ev_in[k]);
    if (e->ev_cmp[k]) cudaEventDestroy(e->ev_cmp[k]);
    if (e->ev_out[k]) cudaEventDestroy(e->ev_out[k]);
  }
  if (e->s_in) cudaStreamDestroy(e->s_in);
  if (e->s_out) cudaStreamDestroy(e->s_out);
  for (cudaEvent_t ev : e->ev_ana) if (ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : e->ev_net) if (ev) cudaEventDestroy(ev);
  for (int p = 0; p < 2; p++)
    for (cudaEvent_t ev : e->ev_syn[p]) if (ev) cudaEventDestroy(ev);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join_net) cudaEventDestroy(e->ev_join_net);
  if (e->ev_join_dsp) cudaEventDestroy(e->ev_join_dsp);
  if (e->ev_join_syn) cudaEventDestroy(e->ev_join_syn);
  if (e->s_net) cudaStreamDestroy(e->s_net);
  if (e->s_dsp) cudaStreamDestroy(e->s_dsp);
  if (e->s_syn) cudaStreamDestroy(e->s_syn);
  if (e->green_net) green_api().GreenCtxDestroy((CUgreenCtx)e->green_net);
  if (e->green_dsp) green_api().GreenCtxDestroy((CUgreenCtx)e->green_dsp);
  for (auto &r : e->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto ev : e->prof_pool) cudaEventDestroy(ev);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

extern "C" int pnb_reset(pnb_engine *e) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  CK(cudaSetDevice(e->device));
  cudaDeviceSynchronize();  // a poisoned engine may carry a sticky launch error of its own; the memsets below report a dead context
  cudaGetLastError();
  const size_t S = e->S;
  CK(cudaMemset(e->d_pcm, 0, S * e->pcm_stride * sizeof(float)));
  e->line_off = 0;
  CK(cudaMemset(e->d_synth, 0, S * kFrame * sizeof(float)));
  CK(cudaMemset(e->d_zring, 0, (size_t)e->ring * S * kBins * sizeof(float2)));
  CK(cudaMemset(e->d_ering, 0, (size_t)e->ring * S * kBands * sizeof(float)));
  CK(cudaMemset(e->d_last_period, 0, S * sizeof(int)));
  CK(cudaMemset(e->d_last_gain, 0, S * sizeof(float)));
  CK(cudaMemset(e->d_status, 0, sizeof(int)));
  *e->h_status = 0;
  e->hop = 0;
  e->poisoned = false;
  e->unjoined = false;      // the device was synchronised above (green-context streams included)
  e->prev_start.clear();
  if (e->flags & PNB_TRAIN_DATA) {
    CK(cudaDeviceSynchronize());
    return PNB_OK;
  }
  if (e->ring_fc) CK(cudaMemset(e->ring_fc, 0, (size_t)(e->f32_chunk + 4) * S * 128 * sizeof(float)));
  if (e->ring_c1) CK(cudaMemset(e->ring_c1, 0, (size_t)(e->f32_chunk + 2) * S * 512 * sizeof(float)));
  if (e->f32_cnt) CK(cudaMemset(e->f32_cnt, 0, (size_t)5 * e->f32_rb * sizeof(unsigned)));
  CK(cudaMemset(e->c2, 0, S * 512 * sizeof(float)));
  for (int i = 0; i < 5; i++)
    for (int p = 0; p < 2; p++)
      if (e->h[i][p]) CK(cudaMemset(e->h[i][p], 0, (size_t)((e->flags & PNB_NN_TENSOR) ? 1 : e->f32_chunk + 1) * S * e->gru[i].H * sizeof(float)));
  for (int i = 0; i < 5; i++) e->par[i] = 0;
  int trc = tc_reset(e);
  if (trc) return trc;
  CK(cudaDeviceSynchronize());
  return PNB_OK;
}

// ------------------------------------------------------------------------------------------
// per-kernel-class timing
// ------------------------------------------------------------------------------------------
static cudaEvent_t prof_event(pnb_engine *e) {
  if (!e->prof_pool.empty()) { cudaEvent_t ev = e->prof_pool.back(); e->prof_pool.pop_back(); return ev; }
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  return ev;
}
ProfScope::ProfScope(pnb_engine *eng, int cls, cudaStream_t s) : e(eng), idx(-1), st(s) {
  if (!e->profiling) return;
  pnb_engine::ProfRec r;
  r.cls = cls; r.a = prof_event(e); r.b = prof_event(e);
  cudaEventRecord(r.a, st);
  e->prof_pending.push_back(r);
  idx = (int)e->prof_pending.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(e->prof_pending[idx].b, st);
}
extern "C" int pnb_profile_enable(pnb_engine *e, int on) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  e->profiling = on != 0;
  return PNB_OK;
}
extern "C" int pnb_profile_read(pnb_engine *e, double *ms, long long *counts) {
  if (!e || !ms || !counts) return fail(PNB_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(e->device));
  CK(cudaDeviceSynchronize());
  for (auto &r : e->prof_pending) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, r.a, r.b));
    e->prof_ms[r.cls] += t;
    e->prof_n[r.cls] += 1;
    e->prof_pool.push_back(r.a);
    e->prof_pool.push_back(r.b);
  }
  e->prof_pending.clear();
  for (int i = 0; i < PNB_NUM_KERNEL_CLASSES; i++) {
    ms[i] = e->prof_ms[i]; counts[i] = e->prof_n[i];
    e->prof_ms[i] = 0; e->prof_n[i] = 0;
  }
  return PNB_OK;
}
// Every launch of the profiled calls as (class, start, end) in ms relative to the first launch, in launch order;
// consumes the pending records like pnb_profile_read.  Returns the number of records written (<= cap) or a negative error.
extern "C" int pnb_profile_timeline(pnb_engine *e, int *cls, double *t0_ms, double *t1_ms, int cap) {
  if (!e || !cls || !t0_ms || !t1_ms) return fail(PNB_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(e->device));
  CK(cudaDeviceSynchronize());
  int n = 0;
  for (auto &r : e->prof_pending) {
    if (n < cap) {
      float a = 0.f, b = 0.f;
      CK(cudaEventElapsedTime(&a, e->prof_pending[0].a, r.a));
      CK(cudaEventElapsedTime(&b, e->prof_pending[0].a, r.b));
      cls[n] = r.cls; t0_ms[n] = a; t1_ms[n] = b;
      n++;
    }
    e->prof_pool.push_back(r.a);
    e->prof_pool.push_back(r.b);
  }
  e->prof_pending.clear();
  return n;
}
extern "C" const char *pnb_kernel_class_name(int cls) {
  static const char *names[PNB_NUM_KERNEL_CLASSES] = {"stage_in_kernel", "analysis_kernel", "fc_f32_kernel",
      "gemm_f32_kernel", "f32_carry_kernel", "synthesis_kernel", "slide_history_kernel", "tc_gemm_kernel", "tc_aux_kernel", "train_labels_kernel"};
  return (cls >= 0 && cls < PNB_NUM_KERNEL_CLASSES) ? names[cls] : "?";
}

// ------------------------------------------------------------------------------------------
// one network step in fp32 (rnn.cpp:42-81), hop index `c` counted since reset
// ------------------------------------------------------------------------------------------
static GemmSeg seg(const float *A, int lda, const float *B, int ldb, int K) {
  GemmSeg s;
  s.A = A; s.lda = lda; s.B = B; s.ldb = ldb; s.K = K;
  return s;
}

// The network of hops [h0, h0 + n) of a call (n <= e->f32_chunk) in fp32 FMA (rnn.cpp:42-81): the non-recurrent layers
// once over all n S rows, the five GRUs of all n hops in one persistent launch, the output layers over n S rows, and the
// carry that moves the conv histories and the states back to the front of their slot buffers.
static int nn_chunk_f32(pnb_engine *e, int h0, int n, cudaStream_t st) {
  const int S = e->S, rows = n * S;
  const float *tbl = e->tansig();
  int nl = 0;
  // fc -> slots 4.. (slots 0..3 hold the last four hops' outputs)
  { ProfScope ps(e, PNB_K_FC, st); nl += launch_fc_f32(e->d_feat + (size_t)h0 * S * kFeat, e->fc.W, e->fc.b, e->ring_fc + (size_t)4 * S * 128, rows, 70, 128, st); }
  GemmArgs g;
  memset(&g, 0, sizeof g);
  g.M = rows; g.tansig = tbl;
  // conv1: tap q of hop t reads fc slot t + q (oldest first, nnet.cpp:182-200; weights [tap][c][n]) -> c1 slots 2..
  for (int q = 0; q < 5; q++) g.seg[q] = seg(e->ring_fc + (size_t)q * S * 128, 128, e->conv1.W + (size_t)q * 128 * 512, 512, 128);
  g.n_seg = 5; g.N = 512; g.C = e->ring_c1 + (size_t)2 * S * 512; g.ldc = 512; g.bias = e->conv1.b; g.act = e->act_conv1;
  { ProfScope ps(e, PNB_K_GEMM_F32, st); nl += launch_gemm_f32(g, st); }
  for (int q = 0; q < 3; q++) g.seg[q] = seg(e->ring_c1 + (size_t)q * S * 512, 512, e->conv2.W + (size_t)q * 512 * 512, 512, 512);
  g.n_seg = 3; g.N = 512; g.C = e->c2_all; g.ldc = 512; g.bias = e->conv2.b; g.act = e->act_conv2;
  { ProfScope ps(e, PNB_K_GEMM_F32, st); nl += launch_gemm_f32(g, st); }

  // stacked GRUs: each consumes the freshly updated state of the one below
  F32ChainArgs a;
  memset(&a, 0, sizeof a);
  a.S = S; a.n_rb = e->f32_rb; a.cnt = e->f32_cnt; a.tansig = tbl;
  for (int li = 0; li < 5; li++) {
    F32ChainLayer &L = a.L[li];
    const int H = e->gru[li].H;
    L.H = H; L.W = e->gru[li].W; L.U = e->gru[li].U; L.bias = e->gru[li].b; L.ldw = 3 * H; L.h = e->h[li][0];
    if (li == 0) {          // gru1 <- conv2 out
      L.n_x = 1; L.dep = -1; L.x[0] = e->c2_all; L.x_ld[0] = 512; L.x_K[0] = 512; L.x_slot1[0] = 0; L.x_slot_stride[0] = (size_t)S * 512;
    } else if (li < 4) {    // gru2 <- gru1, gru3 <- gru2, gru_gb <- gru3
      L.n_x = 1; L.dep = li - 1; L.x[0] = e->h[li - 1][0]; L.x_ld[0] = 512; L.x_K[0] = 512; L.x_slot1[0] = 1; L.x_slot_stride[0] = (size_t)S * 512;
    } else {                // gru_rb <- [gru3 state, conv2 out] (rnn.cpp:69-71)
      L.n_x = 2; L.dep = 2;
      L.x[0] = e->h[2][0]; L.x_ld[0] = 512; L.x_K[0] = 512; L.x_slot1[0] = 1; L.x_slot_stride[0] = (size_t)S * 512; L.w_row0[0] = 0;
      L.x[1] = e->c2_all;  L.x_ld[1] = 512; L.x_K[1] = 512; L.x_slot1[1] = 0; L.x_slot_stride[1] = (size_t)S * 512; L.w_row0[1] = 512;
    }
  }
  {
    // anti-diagonal order: (t, l) sits on diagonal t + depth(l); everything it needs is on an earlier one
    static const int depth[5] = {0, 1, 2, 3, 3};
    int k = 0, unit = 0;
    for (int d = 0; d < n + 3; d++)
      for (int li = 0; li < 5; li++) {
        const int t = d - depth[li];
        if (t < 0 || t >= n) continue;
        a.lh[k] = (unsigned char)((t << 3) | li);
        a.lh_unit0[k] = unit;
        unit += e->f32_rb * (e->gru[li].H / 32);
        k++;
      }
    a.lh_unit0[k] = unit;
    a.n_lh = k; a.n_units = unit;
  }
  {
    ProfScope ps(e, PNB_K_GEMM_F32, st);
    if (launch_gru_chain_f32(a, e->f32_rt, e->sm_count, st) < 0) return fail(PNB_ERR_CUDA, "cannot configure the fp32 GRU chain kernel");
    nl++;
  }
  // fc_gb on [conv2, gru1, gru2, gru3, gru_gb] (rnn.cpp:73-78), fc_rb on gru_rb (rnn.cpp:80); state after hop t = slot t+1
  float *gr = e->d_gr + (size_t)h0 * S * 68;
  g.seg[0] = seg(e->c2_all, 512, e->fc_gb.W, 34, 512);
  for (int q = 1; q < 5; q++) g.seg[q] = seg(e->h[q - 1][0] + (size_t)S * 512, 512, e->fc_gb.W + (size_t)q * 512 * 34, 34, 512);
  g.n_seg = 5; g.N = 34; g.C = gr; g.ldc = 68; g.bias = e->fc_gb.b; g.act = e->act_gb;
  { ProfScope ps(e, PNB_K_GEMM_F32, st); nl += launch_gemm_f32(g, st); }
  g.seg[0] = seg(e->h[4][0] + (size_t)S * 128, 128, e->fc_rb.W, 34, 128);
  g.n_seg = 1; g.N = 34; g.C = gr + 34; g.ldc = 68; g.bias = e->fc_rb.b; g.act = e->act_rb;
  { ProfScope ps(e, PNB_K_GEMM_F32, st); nl += launch_gemm_f32(g, st); }
  // carry
  F32CarryArgs c;
  memset(&c, 0, sizeof c);
  int ns = 0;
  auto cseg = [&](float *base, size_t slot_floats, int n_slots, int from) {
    c.seg[ns++] = F32CarrySeg{reinterpret_cast<float4 *>(base), reinterpret_cast<const float4 *>(base + (size_t)from * slot_floats), slot_floats / 4, n_slots};
  };
  cseg(e->ring_fc, (size_t)S * 128, 4, n);
  cseg(e->ring_c1, (size_t)S * 512, 2, n);
  for (int li = 0; li < 5; li++) cseg(e->h[li][0], (size_t)S * e->gru[li].H, 1, n);
  c.seg[ns++] = F32CarrySeg{reinterpret_cast<float4 *>(e->c2), reinterpret_cast<const float4 *>(e->c2_all + (size_t)(n - 1) * S * 512), (size_t)S * 512 / 4, 1};
  c.n_seg = ns; c.cnt = e->f32_cnt; c.n_cnt = 5 * e->f32_rb;
  { ProfScope ps(e, PNB_K_F32_CARRY, st); nl += launch_f32_carry(c, st); }
  return nl;
}

// ------------------------------------------------------------------------------------------
// The next call's window starts n_frames*480 samples further on; when it would not fit any more, the last 5280
// samples are moved back to the start of the rows.  Returns the number of launches.
static int advance_line(pnb_engine *e, int F, cudaStream_t st) {
  e->line_off += (size_t)F * kFrame;
  if (e->line_off + kKeep + (size_t)e->Fmax * kFrame <= e->pcm_stride) return 0;
  ProfScope ps(e, PNB_K_SLIDE, st);
  int n = launch_slide_history(e->d_pcm, e->pcm_stride, e->S, (int)e->line_off, st);
  e->line_off = 0;
  return n;
}

// Enqueues one call.  Host-side stream state (hop counter, line offset, GRU buffer parity) is committed only after
// every launch was accepted; a failure in between leaves work half-enqueued, so the engine is poisoned until pnb_reset.
static AnalysisArgs analysis_args(pnb_engine *e, int h0, int n) {
  const size_t S = e->S, o = (size_t)h0 * S;
  AnalysisArgs a;
  a.pcm = e->d_pcm + e->line_off + (size_t)h0 * kFrame;  // this chunk's [history | new hops] window of every row
  a.pcm_stride = e->pcm_stride; a.n_streams = e->S; a.n_frames = n; a.tab = e->d_tab;
  a.feat = e->d_feat + o * kFeat; a.zring = e->d_zring; a.ering = e->d_ering; a.ring = e->ring; a.hop0 = e->hop + h0;
  a.P = e->d_P ? e->d_P + o * kBins : nullptr; a.Ex = e->d_Ex + o * kBands; a.raw = nullptr; a.silence = e->d_sil + o;
  a.last_period = e->d_last_period; a.last_gain = e->d_last_gain;
  a.tap_pitch = e->d_tap_pitch ? e->d_tap_pitch + o * 4 : nullptr;
  a.tap_pitchf = e->d_tap_pitchf ? e->d_tap_pitchf + o * 2 : nullptr;
  return a;
}
static SynthesisArgs synthesis_args(pnb_engine *e, int h0, int n, float *d_out, short *d_out16, size_t out_stride) {
  const size_t S = e->S, o = (size_t)h0 * S;
  SynthesisArgs s;
  s.zring = e->d_zring; s.ring = e->ring; s.hop0 = e->hop + h0; s.P = e->d_P + o * kBins; s.gr = e->d_gr + o * 68;
  s.Ex = e->d_Ex + o * kBands; s.silence = e->d_sil + o; s.n_streams = e->S; s.n_frames = n;
  s.tab = e->d_tab; s.synth_mem = e->d_synth;
  s.out = d_out ? d_out + (size_t)h0 * kFrame : nullptr;
  s.out16 = d_out16 ? d_out16 + (size_t)h0 * kFrame : nullptr;
  s.out_stride = out_stride;
  s.postfilter = (e->flags & PNB_POSTFILTER) ? 1 : 0;
  s.tap_g = e->d_tap_g ? e->d_tap_g + o * kBands : nullptr;
  return s;
}

// How a call is tied to the caller: a plain call forks from the caller's stream and joins back into it; a submitted
// call (pnb_submit_*) starts when `ready` fires (or at the current end of the caller's stream when it is null), records
// `done` behind its last synthesis and is NOT joined into any stream, so that the next call's analysis can start
// while this call's network and synthesis are still running.
struct CallLink {
  bool submitted = false;
  cudaEvent_t ready = nullptr, done = nullptr;
};

// The caller's stream (or the host, st == nullptr ... not used) catches up with work left in flight by submitted calls.
static int join_engine_streams(pnb_engine *e, cudaStream_t st) {
  if (!e->unjoined) return PNB_OK;
  CK(cudaEventRecord(e->ev_join_net, e->s_net));
  CK(cudaEventRecord(e->ev_join_dsp, e->s_dsp));
  CK(cudaEventRecord(e->ev_join_syn, e->s_syn));
  CK(cudaStreamWaitEvent(st, e->ev_join_net, 0));
  CK(cudaStreamWaitEvent(st, e->ev_join_dsp, 0));
  CK(cudaStreamWaitEvent(st, e->ev_join_syn, 0));
  e->unjoined = false;
  return PNB_OK;
}
static int drain_engine_streams(pnb_engine *e) {  // host-side
  if (!e->unjoined) return PNB_OK;
  CK(cudaStreamSynchronize(e->s_dsp));
  CK(cudaStreamSynchronize(e->s_net));
  CK(cudaStreamSynchronize(e->s_syn));
  e->unjoined = false;
  return PNB_OK;
}

// Enqueues one call.  Host-side stream state (hop counter, line offset, GRU buffer parity) is committed only after
// every launch was accepted; a failure in between leaves work half-enqueued, so the engine is poisoned until pnb_reset.
static int process_device_enqueue(pnb_engine *e, const float *d_in, const short *d_in16, size_t in_stride, float *d_out,
                                  short *d_out16, size_t out_stride, int F, float *d_gr, cudaStream_t st, const CallLink &lk,
                                  long long *n_out) {
  const int S = e->S;
  long long n = 0;
  float *line = e->d_pcm + e->line_off;
  const bool tensor = (e->flags & PNB_NN_TENSOR) != 0;
  if (tensor && e->net_sms > 0 && F >= 2 * e->chunk) {
    // ---- chunked schedule: s_dsp runs stage-in + analysis + fc of chunk k+1 and s_syn the synthesis of chunk k-1 (both
    // on the DSP partition) while s_net runs the network of chunk k on its own SMs
    cudaStream_t sd = e->s_dsp, sn = e->s_net, ss = e->s_syn;
    // chunk lengths: short at both ends (the first analysis and the last network chunk run with the other partition
    // idle unless calls overlap), e->chunk in between
    int len[kMaxChunks], start[kMaxChunks + 1], C = 0;
    {
      int left = F, head[2] = {e->chunk / 4 > 0 ? e->chunk / 4 : 1, e->chunk / 2 > 0 ? e->chunk / 2 : 1};
      if (!e->ramp) {
        while (left > 0) { len[C] = left < e->chunk ? left : e->chunk; left -= len[C++]; }
      }
      const int tail_total = head[0] + head[1];
      for (int i = 0; i < 2 && left > tail_total + e->chunk; i++) { len[C++] = head[i]; left -= head[i]; }
      while (left > tail_total + e->chunk) { len[C++] = e->chunk; left -= e->chunk; }
      if (left > tail_total) { len[C++] = left - tail_total; left = tail_total; }
      if (left > head[0]) { len[C++] = left - head[0]; left = head[0]; }
      if (left > 0) len[C++] = left;
      start[0] = 0;
      for (int k = 0; k < C; k++) start[k + 1] = start[k] + len[k];
    }
    // start of the call: the caller's stream unless a ready event was given; when the previous call did not run on the
    // engine's streams (first call, serial schedule) the caller's stream is what orders this call behind it
    if (!lk.ready || e->prev_start.empty()) {
      CK(cudaEventRecord(e->ev_fork, st));
      CK(cudaStreamWaitEvent(sd, e->ev_fork, 0));
      CK(cudaStreamWaitEvent(sn, e->ev_fork, 0));
      CK(cudaStreamWaitEvent(ss, e->ev_fork, 0));
    }
    if (lk.ready) CK(cudaStreamWaitEvent(sd, lk.ready, 0));
    const int cur = e->syn_par ^ 1, prv = e->syn_par;
    const int pC = (int)e->prev_start.size();
    e->tc_sms = e->net_sms;
    for (int k = 0; k < C; k++) {
      const int h0 = start[k], nh = len[k], h1 = h0 + nh;
      if (pC) {
        // What this chunk overwrites was last read by the previous call: hop slots [h0, h1) of the per-call buffers and
        // the ring slots behind them by its synthesis of the same hops, fc slots [h0+4, h1+4) by its conv1 of hops
        // < h1+4 (and by its carry, which runs before the last chunk's network-done event).  The synthesis of the last
        // previous chunk that starts before hop h1+4 is behind all of them.
        int j = 0;
        while (j + 1 < pC && e->prev_start[j + 1] < h1 + 4) j++;
        CK(cudaStreamWaitEvent(sd, e->ev_syn[prv][j], 0));
      }
      {
        ProfScope ps(e, PNB_K_STAGE_IN, sd);
        n += launch_stage_in(line + (size_t)h0 * kFrame, e->pcm_stride, d_in ? d_in + (size_t)h0 * kFrame : nullptr,
                             d_in16 ? d_in16 + (size_t)h0 * kFrame : nullptr, in_stride, S, nh * kFrame, sd);
      }
      { ProfScope ps(e, PNB_K_ANALYSIS, sd); n += launch_analysis(analysis_args(e, h0, nh), sd); }
      n += tc_fc(e, h0, nh, sd);  // CUDA-core work: with the DSP kernels
      CK(cudaEventRecord(e->ev_ana[k], sd));
      CK(cudaStreamWaitEvent(sn, e->ev_ana[k], 0));
      int q = tc_front(e, h0, nh, F, sn);
      if (q < 0) { e->tc_sms = e->sm_count; return q; }
      n += q;
      if ((q = tc_gru_chain(e, h0, nh, sn)) < 0) { e->tc_sms = e->sm_count; return q; }
      n += q;
      if ((q = tc_out(e, h0, nh, sn)) < 0) { e->tc_sms = e->sm_count; return q; }
      n += q;
      if (k == C - 1) {  // the carry reads the call's last fc / conv1 / state slots: before the event the next call waits for
        if ((q = tc_carry(e, F, sn)) < 0) { e->tc_sms = e->sm_count; return q; }
        n += q;
      }
      CK(cudaEventRecord(e->ev_net[k], sn));
      CK(cudaStreamWaitEvent(ss, e->ev_net[k], 0));
      {
        ProfScope ps(e, PNB_K_SYNTHESIS, ss);
        n += launch_synthesis(synthesis_args(e, h0, nh, d_out, d_out16, out_stride), ss);
      }
      CK(cudaEventRecord(e->ev_syn[cur][k], ss));
    }
    e->tc_sms = e->sm_count;
    n += advance_line(e, F, sd);
    e->prev_start.assign(start, start + C);
    e->syn_par = cur;
    e->unjoined = true;
    if (lk.submitted) {
      if (lk.done) CK(cudaEventRecord(lk.done, ss));
    } else {
      int rc = join_engine_streams(e, st);
      if (rc) return rc;
    }
  } else {
    // ---- serial schedule on the caller's stream
    if (lk.ready) CK(cudaStreamWaitEvent(st, lk.ready, 0));
    { int rc = join_engine_streams(e, st); if (rc) return rc; }
    e->prev_start.clear();
    { ProfScope ps(e, PNB_K_STAGE_IN, st); n += launch_stage_in(line, e->pcm_stride, d_in, d_in16, in_stride, S, F * kFrame, st); }
    { ProfScope ps(e, PNB_K_ANALYSIS, st); n += launch_analysis(analysis_args(e, 0, F), st); }
    CK(cudaGetLastError());
    if (tensor) {
      n += tc_fc(e, 0, F, st);
      int k = tc_front(e, 0, F, F, st);
      if (k < 0) return k;
      n += k;
      if ((k = tc_gru_chain(e, 0, F, st)) < 0) return k;
      n += k;
      if ((k = tc_out(e, 0, F, st)) < 0) return k;
      n += k;
      if ((k = tc_carry(e, F, st)) < 0) return k;
      n += k;
    } else {
      for (int h0 = 0; h0 < F; h0 += e->f32_chunk) {
        const int q = nn_chunk_f32(e, h0, F - h0 < e->f32_chunk ? F - h0 : e->f32_chunk, st);
        if (q < 0) return q;
        n += q;
      }
    }
    { ProfScope ps(e, PNB_K_SYNTHESIS, st); n += launch_synthesis(synthesis_args(e, 0, F, d_out, d_out16, out_stride), st); }
    n += advance_line(e, F, st);
    if (lk.submitted && lk.done) CK(cudaEventRecord(lk.done, st));
  }
  if (d_gr) CK(cudaMemcpyAsync(d_gr, e->d_gr, (size_t)F * S * 68 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CK(cudaGetLastError());
  *n_out = n;
  return PNB_OK;
}

static int process_device(pnb_engine *e, const float *d_in, const short *d_in16, size_t in_stride, float *d_out,
                          short *d_out16, size_t out_stride, int F, float *d_gr, cudaStream_t st,
                          const CallLink &lk = CallLink()) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  if (e->flags & PNB_TRAIN_DATA) return fail(PNB_ERR_ARG, "engine was created with PNB_TRAIN_DATA: use pnb_train_records_*");
  if (e->poisoned) return fail(PNB_ERR_CUDA, "an earlier call failed while enqueuing work; the streams' state is undefined until pnb_reset");
  if (F < 1 || F > e->Fmax) return fail(PNB_ERR_ARG, "n_frames %d outside [1, %d]", F, e->Fmax);
  if ((!d_in && !d_in16) || (!d_out && !d_out16)) return fail(PNB_ERR_ARG, "input/output pointer is NULL");
  if (in_stride < (size_t)F * kFrame || out_stride < (size_t)F * kFrame)
    return fail(PNB_ERR_ARG, "row stride smaller than n_frames*480");
  CK(cudaSetDevice(e->device));
  long long n = 0;
  if (lk.submitted && d_gr) return fail(PNB_ERR_ARG, "submitted calls do not copy g/r out");
  const int rc = process_device_enqueue(e, d_in, d_in16, in_stride, d_out, d_out16, out_stride, F, d_gr, st, lk, &n);
  if (rc) {
    e->poisoned = true;
    return rc;
  }
  e->hop += F;
  e->last_frames = F;
  e->launches += n;
  return PNB_OK;
}

extern "C" int pnb_process_device_f32(pnb_engine *e, const float *d_in, size_t in_stride, float *d_out,
                                      size_t out_stride, int n_frames, float *d_gr, void *cuda_stream) {
  return process_device(e, d_in, nullptr, in_stride, d_out, nullptr, out_stride, n_frames, d_gr,
                        (cudaStream_t)cuda_stream);
}
extern "C" int pnb_process_device_i16(pnb_engine *e, const short *d_in, size_t in_stride, short *d_out,
                                      size_t out_stride, int n_frames, float *d_gr, void *cuda_stream) {
  return process_device(e, nullptr, d_in, in_stride, nullptr, d_out, out_stride, n_frames, d_gr,
                        (cudaStream_t)cuda_stream);
}

// ------------------------------------------------------------------------------------------
// training-data generator (row f1): stage both files of every pair, analysis over 2N streams, labels
// ------------------------------------------------------------------------------------------
extern "C" int pnb_train_records_device(pnb_engine *e, const short *d_speech, size_t speech_stride,
                                        const short *d_noisy, size_t noisy_stride, int F, float *d_records,
                                        size_t records_stride, void *cuda_stream) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  if (!(e->flags & PNB_TRAIN_DATA)) return fail(PNB_ERR_ARG, "engine was not created with PNB_TRAIN_DATA");
  if (F < 1 || F > e->Fmax) return fail(PNB_ERR_ARG, "n_frames %d outside [1, %d]", F, e->Fmax);
  if (!d_speech || !d_noisy || !d_records) return fail(PNB_ERR_ARG, "speech/noisy/records pointer is NULL");
  if (speech_stride < (size_t)F * kFrame || noisy_stride < (size_t)F * kFrame)
    return fail(PNB_ERR_ARG, "row stride smaller than n_frames*480");
  if (records_stride < (size_t)F * PNB_RECORD_FLOATS) return fail(PNB_ERR_ARG, "records_stride smaller than n_frames*138");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int S = e->S, N = S / 2;
  long long n = 0;
  float *line = e->d_pcm + e->line_off;
  {
    ProfScope ps(e, PNB_K_STAGE_IN, st);
    n += launch_stage_in(line, e->pcm_stride, nullptr, d_noisy, noisy_stride, N, F * kFrame, st, 1.f);
    n += launch_stage_in(line + (size_t)N * e->pcm_stride, e->pcm_stride, nullptr, d_speech, speech_stride, N, F * kFrame, st, 1.f);
  }
  AnalysisArgs a;
  a.pcm = line; a.pcm_stride = e->pcm_stride; a.n_streams = S; a.n_frames = F; a.tab = e->d_tab;
  a.feat = e->d_feat; a.zring = e->d_zring; a.ering = e->d_ering; a.ring = e->ring; a.hop0 = e->hop;
  a.P = nullptr; a.Ex = e->d_Ex; a.raw = e->d_raw; a.silence = e->d_sil;
  a.last_period = e->d_last_period; a.last_gain = e->d_last_gain;
  a.tap_pitch = e->d_tap_pitch; a.tap_pitchf = e->d_tap_pitchf;
  { ProfScope ps(e, PNB_K_ANALYSIS, st); n += launch_analysis(a, st); }
  LabelArgs l;
  l.feat = e->d_feat; l.raw = e->d_raw; l.Ex = e->d_Ex; l.tab = e->d_tab; l.n_pairs = N; l.n_frames = F;
  l.records = d_records; l.pair_stride = records_stride;
  { ProfScope ps(e, PNB_K_LABELS, st); n += launch_train_labels(l, st); }
  n += advance_line(e, F, st);
  CK(cudaGetLastError());
  e->hop += F;
  e->last_frames = F;
  e->launches += n;
  return PNB_OK;
}

extern "C" int pnb_train_records_host(pnb_engine *e, const short *speech, size_t speech_stride, const short *noisy,
                                      size_t noisy_stride, int F, float *records, size_t records_stride) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  if (!(e->flags & PNB_TRAIN_DATA)) return fail(PNB_ERR_ARG, "engine was not created with PNB_TRAIN_DATA");
  if (!speech || !noisy || !records) return fail(PNB_ERR_ARG, "speech/noisy/records pointer is NULL");
  if (F < 1 || F > e->Fmax) return fail(PNB_ERR_ARG, "n_frames %d outside [1, %d]", F, e->Fmax);
  if (speech_stride < (size_t)F * kFrame || noisy_stride < (size_t)F * kFrame)
    return fail(PNB_ERR_ARG, "row stride smaller than n_frames*480");
  if (records_stride < (size_t)F * PNB_RECORD_FLOATS) return fail(PNB_ERR_ARG, "records_stride smaller than n_frames*138");
  CK(cudaSetDevice(e->device));
  const size_t N = e->S / 2, row = (size_t)e->Fmax * kFrame, rrow = (size_t)e->Fmax * PNB_RECORD_FLOATS;
  if (!e->d_hin16) CK(cudaMalloc((void **)&e->d_hin16, 2 * N * row * sizeof(short)));
  if (!e->d_records) CK(cudaMalloc((void **)&e->d_records, N * rrow * sizeof(float)));
  const size_t w = (size_t)F * kFrame * sizeof(short);
  short *d_noisy = e->d_hin16, *d_speech = e->d_hin16 + N * row;
  CK(cudaMemcpy2DAsync(d_noisy, row * 2, noisy, noisy_stride * 2, w, N, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpy2DAsync(d_speech, row * 2, speech, speech_stride * 2, w, N, cudaMemcpyHostToDevice, e->stream));
  int rc = pnb_train_records_device(e, d_speech, row, d_noisy, row, F, e->d_records, rrow, e->stream);
  if (rc) return rc;
  CK(cudaMemcpy2DAsync(records, records_stride * 4, e->d_records, rrow * 4, (size_t)F * PNB_RECORD_FLOATS * 4, N,
                       cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return PNB_OK;
}

// after the stream that carried the status copy was synchronised
static int report_status(pnb_engine *e) {
  if (*e->h_status & 1)
    return fail(PNB_ERR_DOMAIN, "tensor path: a pre-activation left the domain of the reference's tansig_approx (|x| >= 8.5e7, "
                                "src/vec.h:63 is undefined there); results since the last pnb_reset are not the reference's -- "
                                "use PNB_NN_FP32 for such input");
  return PNB_OK;
}

template <typename T>
static int process_host(pnb_engine *e, const T *in, size_t in_stride, T *out, size_t out_stride, int F, float *gr,
                        T **d_in_p, T **d_out_p) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  if (e->flags & PNB_TRAIN_DATA) return fail(PNB_ERR_ARG, "engine was created with PNB_TRAIN_DATA: use pnb_train_records_*");
  if (!in || !out) return fail(PNB_ERR_ARG, "input/output pointer is NULL");
  if (F < 1 || F > e->Fmax) return fail(PNB_ERR_ARG, "n_frames %d outside [1, %d]", F, e->Fmax);
  if (in_stride < (size_t)F * kFrame || out_stride < (size_t)F * kFrame)
    return fail(PNB_ERR_ARG, "row stride smaller than n_frames*480");
  CK(cudaSetDevice(e->device));
  const size_t S = e->S, row = (size_t)e->Fmax * kFrame;
  if (!*d_in_p) CK(cudaMalloc((void **)d_in_p, S * row * sizeof(T)));
  if (!*d_out_p) CK(cudaMalloc((void **)d_out_p, S * row * sizeof(T)));
  const size_t w = (size_t)F * kFrame * sizeof(T);
  CK(cudaMemcpy2DAsync(*d_in_p, row * sizeof(T), in, in_stride * sizeof(T), w, S, cudaMemcpyHostToDevice, e->stream));
  int rc;
  if (sizeof(T) == 4)
    rc = process_device(e, (const float *)*d_in_p, nullptr, row, (float *)*d_out_p, nullptr, row, F, nullptr, e->stream);
  else
    rc = process_device(e, nullptr, (const short *)*d_in_p, row, nullptr, (short *)*d_out_p, row, F, nullptr, e->stream);
  if (rc) return rc;
  CK(cudaMemcpy2DAsync(out, out_stride * sizeof(T), *d_out_p, row * sizeof(T), w, S, cudaMemcpyDeviceToHost, e->stream));
  if (gr) CK(cudaMemcpyAsync(gr, e->d_gr, (size_t)F * S * 68 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(e->h_status, e->d_status, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return report_status(e);
}

extern "C" int pnb_process_host_f32(pnb_engine *e, const float *in, size_t in_stride, float *out, size_t out_stride,
                                    int n_frames, float *gr) {
  return process_host<float>(e, in, in_stride, out, out_stride, n_frames, gr, &e->d_hin, &e->d_hout);
}
extern "C" int pnb_process_host_i16(pnb_engine *e, const short *in, size_t in_stride, short *out, size_t out_stride,
                                    int n_frames, float *gr) {
  return process_host<short>(e, in, in_stride, out, out_stride, n_frames, gr, &e->d_hin16, &e->d_hout16);
}

// ------------------------------------------------------------------------------------------
// pipelined host entry: H2D of call i+1 and D2H of call i-1 overlap the kernels of call i
// ------------------------------------------------------------------------------------------
template <typename T>
static int submit_host(pnb_engine *e, const T *in, size_t in_stride, T *out, size_t out_stride, int F) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  if (e->flags & PNB_TRAIN_DATA) return fail(PNB_ERR_ARG, "engine was created with PNB_TRAIN_DATA: use pnb_submit_train_records");
  if (!in || !out) return fail(PNB_ERR_ARG, "input/output pointer is NULL");
  if (F < 1 || F > e->Fmax) return fail(PNB_ERR_ARG, "n_frames %d outside [1, %d]", F, e->Fmax);
  if (in_stride < (size_t)F * kFrame || out_stride < (size_t)F * kFrame)
    return fail(PNB_ERR_ARG, "row stride smaller than n_frames*480");
  CK(cudaSetDevice(e->device));
  const size_t S = e->S, row = (size_t)e->Fmax * kFrame;
  if (!e->s_in) {
    CK(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      CK(cudaEventCreateWithFlags(&e->ev_in[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&e->ev_cmp[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&e->ev_out[k], cudaEventDisableTiming));
    }
  }
  if (e->pipe_elem != sizeof(T)) {
    CK(cudaDeviceSynchronize());
    for (int k = 0; k < 2; k++) {
      if (e->pipe_in[k]) cudaFree(e->pipe_in[k]);
      if (e->pipe_out[k]) cudaFree(e->pipe_out[k]);
      CK(cudaMalloc(&e->pipe_in[k], S * row * sizeof(T)));
      CK(cudaMalloc(&e->pipe_out[k], S * row * sizeof(T)));
    }
    e->pipe_elem = sizeof(T);
    e->submitted = 0;
  }
  const int k = (int)(e->submitted & 1);
  if (e->submitted >= 2) CK(cudaEventSynchronize(e->ev_out[k]));  // the slot's previous round trip is complete
  const size_t w = (size_t)F * kFrame * sizeof(T);
  CK(cudaMemcpy2DAsync(e->pipe_in[k], row * sizeof(T), in, in_stride * sizeof(T), w, S, cudaMemcpyHostToDevice, e->s_in));
  CK(cudaEventRecord(e->ev_in[k], e->s_in));
  CallLink lk;  // starts when the copy has landed, not joined into any stream: the next call's analysis may overlap this call's tail
  lk.submitted = true; lk.ready = e->ev_in[k]; lk.done = e->ev_cmp[k];
  int rc;
  if (sizeof(T) == 4)
    rc = process_device(e, (const float *)e->pipe_in[k], nullptr, row, (float *)e->pipe_out[k], nullptr, row, F, nullptr, e->stream, lk);
  else
    rc = process_device(e, nullptr, (const short *)e->pipe_in[k], row, nullptr, (short *)e->pipe_out[k], row, F, nullptr, e->stream, lk);
  if (rc) return rc;
  CK(cudaStreamWaitEvent(e->s_out, e->ev_cmp[k], 0));
  CK(cudaMemcpy2DAsync(out, out_stride * sizeof(T), e->pipe_out[k], row * sizeof(T), w, S, cudaMemcpyDeviceToHost, e->s_out));
  CK(cudaEventRecord(e->ev_out[k], e->s_out));
  e->submitted++;
  return PNB_OK;
}
extern "C" int pnb_submit_host_f32(pnb_engine *e, const float *in, size_t in_stride, float *out, size_t out_stride,
                                   int n_frames) {
  return submit_host<float>(e, in, in_stride, out, out_stride, n_frames);
}
extern "C" int pnb_submit_host_i16(pnb_engine *e, const short *in, size_t in_stride, short *out, size_t out_stride,
                                   int n_frames) {
  return submit_host<short>(e, in, in_stride, out, out_stride, n_frames);
}
// pipelined twin of pnb_train_records_host: same slots, streams and events as submit_host
extern "C" int pnb_submit_train_records(pnb_engine *e, const short *speech, size_t speech_stride, const short *noisy,
                                        size_t noisy_stride, int F, float *records, size_t records_stride) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  if (!(e->flags & PNB_TRAIN_DATA)) return fail(PNB_ERR_ARG, "engine was not created with PNB_TRAIN_DATA");
  if (!speech || !noisy || !records) return fail(PNB_ERR_ARG, "speech/noisy/records pointer is NULL");
  if (F < 1 || F > e->Fmax) return fail(PNB_ERR_ARG, "n_frames %d outside [1, %d]", F, e->Fmax);
  if (speech_stride < (size_t)F * kFrame || noisy_stride < (size_t)F * kFrame)
    return fail(PNB_ERR_ARG, "row stride smaller than n_frames*480");
  if (records_stride < (size_t)F * PNB_RECORD_FLOATS) return fail(PNB_ERR_ARG, "records_stride smaller than n_frames*138");
  CK(cudaSetDevice(e->device));
  const size_t N = e->S / 2, row = (size_t)e->Fmax * kFrame, rrow = (size_t)e->Fmax * PNB_RECORD_FLOATS;
  if (!e->s_in) {
    CK(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      CK(cudaEventCreateWithFlags(&e->ev_in[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&e->ev_cmp[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&e->ev_out[k], cudaEventDisableTiming));
    }
  }
  if (!e->pipe_in[0]) {
    for (int k = 0; k < 2; k++) {
      CK(cudaMalloc(&e->pipe_in[k], 2 * N * row * sizeof(short)));
      CK(cudaMalloc(&e->pipe_out[k], N * rrow * sizeof(float)));
    }
    e->submitted = 0;
  }
  const int k = (int)(e->submitted & 1);
  if (e->submitted >= 2) CK(cudaEventSynchronize(e->ev_out[k]));  // the slot's previous round trip is complete
  short *d_noisy = (short *)e->pipe_in[k], *d_speech = d_noisy + N * row;
  const size_t w = (size_t)F * kFrame * sizeof(short);
  CK(cudaMemcpy2DAsync(d_noisy, row * 2, noisy, noisy_stride * 2, w, N, cudaMemcpyHostToDevice, e->s_in));
  CK(cudaMemcpy2DAsync(d_speech, row * 2, speech, speech_stride * 2, w, N, cudaMemcpyHostToDevice, e->s_in));
  CK(cudaEventRecord(e->ev_in[k], e->s_in));
  CK(cudaStreamWaitEvent(e->stream, e->ev_in[k], 0));
  int rc = pnb_train_records_device(e, d_speech, row, d_noisy, row, F, (float *)e->pipe_out[k], rrow, e->stream);
  if (rc) return rc;
  CK(cudaEventRecord(e->ev_cmp[k], e->stream));
  CK(cudaStreamWaitEvent(e->s_out, e->ev_cmp[k], 0));
  CK(cudaMemcpy2DAsync(records, records_stride * 4, e->pipe_out[k], rrow * 4, (size_t)F * PNB_RECORD_FLOATS * 4, N,
                       cudaMemcpyDeviceToHost, e->s_out));
  CK(cudaEventRecord(e->ev_out[k], e->s_out));
  e->submitted++;
  return PNB_OK;
}

// Device-buffer twins of pnb_submit_host_*: the call starts when `cuda_stream` reaches this point (the inputs are ready in
// its order) and is not joined back; pnb_flush makes a stream wait for everything submitted, pnb_wait the host.
extern "C" int pnb_submit_device_f32(pnb_engine *e, const float *d_in, size_t in_stride, float *d_out, size_t out_stride,
                                     int n_frames, void *cuda_stream) {
  CallLink lk;
  lk.submitted = true;
  return process_device(e, d_in, nullptr, in_stride, d_out, nullptr, out_stride, n_frames, nullptr, (cudaStream_t)cuda_stream, lk);
}
extern "C" int pnb_submit_device_i16(pnb_engine *e, const short *d_in, size_t in_stride, short *d_out, size_t out_stride,
                                     int n_frames, void *cuda_stream) {
  CallLink lk;
  lk.submitted = true;
  return process_device(e, nullptr, d_in, in_stride, nullptr, d_out, out_stride, n_frames, nullptr, (cudaStream_t)cuda_stream, lk);
}
extern "C" int pnb_flush(pnb_engine *e, void *cuda_stream) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  CK(cudaSetDevice(e->device));
  return join_engine_streams(e, (cudaStream_t)cuda_stream);
}

extern "C" int pnb_wait(pnb_engine *e) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  CK(cudaSetDevice(e->device));
  if (e->s_in) CK(cudaStreamSynchronize(e->s_in));
  { int rc = drain_engine_streams(e); if (rc) return rc; }
  CK(cudaMemcpyAsync(e->h_status, e->d_status, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (e->s_out) CK(cudaStreamSynchronize(e->s_out));
  return report_status(e);
}
extern "C" int pnb_check(pnb_engine *e, void *cuda_stream) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  int rc = pnb_wait(e);
  if (rc) return rc;
  if (e->poisoned) return fail(PNB_ERR_CUDA, "an earlier call failed while enqueuing work; pnb_reset the engine");
  return PNB_OK;
}

extern "C" int pnb_read_tap(pnb_engine *e, int what, void *dst, size_t dst_bytes) {
  if (!e || !dst) return fail(PNB_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(e->device));
  const size_t n = (size_t)e->last_frames * e->S;
  const void *src = nullptr;
  size_t bytes = 0;
  switch (what) {
    case PNB_TAP_FEATURES: src = e->d_feat; bytes = n * kFeat * 4; break;
    case PNB_TAP_PITCH: src = e->d_tap_pitch; bytes = n * 16; break;
    case PNB_TAP_PITCHF: src = e->d_tap_pitchf; bytes = n * 8; break;
    case PNB_TAP_X: {  // gather from the ring: the analysis spectrum of hop c is the slot of hop c-5
      bytes = n * kBins * 8;
      if (dst_bytes < bytes) return fail(PNB_ERR_ARG, "tap %d needs %zu bytes, got %zu", what, bytes, dst_bytes);
      CK(cudaDeviceSynchronize());
      const size_t per = (size_t)e->S * kBins * 8;
      for (int k = 0; k < e->last_frames; k++) {
        long c = e->hop - e->last_frames + k;
        int slot = (int)(((c - 5) % e->ring + e->ring) % e->ring);
        CK(cudaMemcpy((char *)dst + k * per, (const char *)e->d_zring + slot * per, per, cudaMemcpyDeviceToHost));
      }
      return PNB_OK;
    }
    case PNB_TAP_P: src = e->d_P; bytes = n * kBins * 8; break;
    case PNB_TAP_EX: src = e->d_Ex; bytes = n * kBands * 4; break;
    case PNB_TAP_GR: src = e->d_gr; bytes = n * 68 * 4; break;
    case PNB_TAP_G_USED: src = e->d_tap_g; bytes = n * kBands * 4; break;
    case PNB_TAP_NN_C2: src = e->c2; bytes = (size_t)e->S * 512 * 4; break;
    case PNB_TAP_NN_H0: case PNB_TAP_NN_H0 + 1: case PNB_TAP_NN_H0 + 2: case PNB_TAP_NN_H0 + 3: case PNB_TAP_NN_H0 + 4: {
      int li = what - PNB_TAP_NN_H0;
      src = e->h[li][e->par[li]]; bytes = (size_t)e->S * e->gru[li].H * 4; break;
    }
    default: return fail(PNB_ERR_ARG, "unknown tap %d", what);
  }
  if (!src) return fail(PNB_ERR_ARG, "tap %d needs PNB_KEEP_TAPS at pnb_create", what);
  if (dst_bytes < bytes) return fail(PNB_ERR_ARG, "tap %d needs %zu bytes, got %zu", what, bytes, dst_bytes);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return PNB_OK;
}

// ------------------------------------------------------------------------------------------
// per-stream state export / import (the reference's DenoiseState + RNNState, src/denoise.cpp:71-85, nnet_data.h:28-38,
// reduced to what is live: SURVEY.md App. A.2) -- for moving a stream between engines / GPUs at a call boundary
// ------------------------------------------------------------------------------------------
namespace {
struct StreamState {
  unsigned magic, version;
  float pcm[kKeep];          // the last 5280 input samples (comb_buf without its newest hop)
  float synth[kFrame];       // overlap-add memory
  int last_period;
  float last_gain;
  float spec[5][2 * kBins];  // spectra of the last five windowed blocks (analysis spectrum of the next five hops), oldest first
  float eband[5][kBands];    // their band energies
  float fc_hist[4][128];     // conv1 input history (last four fc outputs), oldest first
  float c1_hist[2][512];     // conv2 input history (last two conv1 outputs)
  float h[4 * 512 + 128];    // gru1, gru2, gru3, gru_gb, gru_rb
  // tensor-mode engines keep the conv histories as three bf16 terms; the sum above does not always split back into the
  // same terms (a low term of exactly half an ulp), so the raw terms travel too and a tensor-mode engine restores those
  unsigned has_terms;
  unsigned short fc_terms[3][4][128];
  unsigned short c1_terms[3][2][512];
};
constexpr unsigned kStateMagic = 0x53424E50u;  // "PNBS"
}  // namespace

extern "C" size_t pnb_state_size(void) { return sizeof(StreamState); }

static int state_args(pnb_engine *e, int stream, const void *buf, size_t bytes) {
  if (!e || !buf) return fail(PNB_ERR_ARG, "NULL argument");
  if (stream < 0 || stream >= e->S) return fail(PNB_ERR_ARG, "stream %d outside [0, %d)", stream, e->S);
  if (bytes < sizeof(StreamState)) return fail(PNB_ERR_ARG, "state buffer of %zu bytes, need pnb_state_size() = %zu", bytes, sizeof(StreamState));
  if (e->poisoned) return fail(PNB_ERR_CUDA, "engine state is undefined after a failed call; pnb_reset it");
  return PNB_OK;
}

extern "C" int pnb_get_state(pnb_engine *e, int stream, void *dst, size_t bytes) {
  int rc = state_args(e, stream, dst, bytes);
  if (rc) return rc;
  CK(cudaSetDevice(e->device));
  CK(cudaDeviceSynchronize());
  StreamState *st = static_cast<StreamState *>(dst);
  memset(st, 0, sizeof *st);
  st->magic = kStateMagic; st->version = 1;
  const size_t s = stream, S = e->S;
  CK(cudaMemcpy(st->pcm, e->d_pcm + s * e->pcm_stride + e->line_off, kKeep * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(st->synth, e->d_synth + s * kFrame, kFrame * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st->last_period, e->d_last_period + s, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st->last_gain, e->d_last_gain + s, 4, cudaMemcpyDeviceToHost));
  for (int k = 0; k < 5; k++) {
    const long c = e->hop - 5 + k;
    const size_t slot = (size_t)((c % e->ring + e->ring) % e->ring);
    CK(cudaMemcpy(st->spec[k], e->d_zring + (slot * S + s) * kBins, kBins * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st->eband[k], e->d_ering + (slot * S + s) * kBands, kBands * 4, cudaMemcpyDeviceToHost));
  }
  if (e->flags & PNB_TRAIN_DATA) return PNB_OK;
  size_t off = 0;
  for (int li = 0; li < 5; li++) {
    const size_t H = e->gru[li].H;
    CK(cudaMemcpy(st->h + off, e->h[li][e->par[li]] + s * H, H * 4, cudaMemcpyDeviceToHost));
    off += H;
  }
  if (e->flags & PNB_NN_TENSOR) {
    st->has_terms = 1;
    return tc_get_stream_hist(e, stream, &st->fc_hist[0][0], &st->c1_hist[0][0], &st->fc_terms[0][0][0], &st->c1_terms[0][0][0]);
  }
  for (int k = 0; k < 4; k++)   // slots 0..3 / 0..1: the histories, oldest first (nn_chunk_f32's carry)
    CK(cudaMemcpy(st->fc_hist[k], e->ring_fc + ((size_t)k * S + s) * 128, 128 * 4, cudaMemcpyDeviceToHost));
  for (int k = 0; k < 2; k++)
    CK(cudaMemcpy(st->c1_hist[k], e->ring_c1 + ((size_t)k * S + s) * 512, 512 * 4, cudaMemcpyDeviceToHost));
  return PNB_OK;
}

extern "C" int pnb_set_state(pnb_engine *e, int stream, const void *src, size_t bytes) {
  int rc = state_args(e, stream, src, bytes);
  if (rc) return rc;
  const StreamState *st = static_cast<const StreamState *>(src);
  if (st->magic != kStateMagic || st->version != 1) return fail(PNB_ERR_ARG, "not a pnb_get_state blob (magic/version)");
  if (e->flags & PNB_NN_TENSOR)
    for (size_t i = 0; i < sizeof st->h / sizeof(float); i++)
      if (!(fabsf(st->h[i]) <= 1.0001f))
        return fail(PNB_ERR_ARG, "GRU state value %g outside [-1, 1]: not a state the network can reach", (double)st->h[i]);
  CK(cudaSetDevice(e->device));
  CK(cudaDeviceSynchronize());
  const size_t s = stream, S = e->S;
  CK(cudaMemcpy(e->d_pcm + s * e->pcm_stride + e->line_off, st->pcm, kKeep * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->d_synth + s * kFrame, st->synth, kFrame * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->d_last_period + s, &st->last_period, 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->d_last_gain + s, &st->last_gain, 4, cudaMemcpyHostToDevice));
  for (int k = 0; k < 5; k++) {
    const long c = e->hop - 5 + k;
    const size_t slot = (size_t)((c % e->ring + e->ring) % e->ring);
    CK(cudaMemcpy(e->d_zring + (slot * S + s) * kBins, st->spec[k], kBins * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_ering + (slot * S + s) * kBands, st->eband[k], kBands * 4, cudaMemcpyHostToDevice));
  }
  if (e->flags & PNB_TRAIN_DATA) return PNB_OK;
  size_t off = 0;
  for (int li = 0; li < 5; li++) {
    const size_t H = e->gru[li].H;
    CK(cudaMemcpy(e->h[li][e->par[li]] + s * H, st->h + off, H * 4, cudaMemcpyHostToDevice));
    off += H;
  }
  if (e->flags & PNB_NN_TENSOR)
    return tc_set_stream_hist(e, stream, &st->fc_hist[0][0], &st->c1_hist[0][0], st->h,
                              st->has_terms ? &st->fc_terms[0][0][0] : nullptr, st->has_terms ? &st->c1_terms[0][0][0] : nullptr);
  for (int k = 0; k < 4; k++)
    CK(cudaMemcpy(e->ring_fc + ((size_t)k * S + s) * 128, st->fc_hist[k], 128 * 4, cudaMemcpyHostToDevice));
  for (int k = 0; k < 2; k++)
    CK(cudaMemcpy(e->ring_c1 + ((size_t)k * S + s) * 512, st->c1_hist[k], 512 * 4, cudaMemcpyHostToDevice));
  return PNB_OK;
}

// ------------------------------------------------------------------------------------------
// binary weight file -> pnb_model (layout of the generated nnet_data.cpp)
// ------------------------------------------------------------------------------------------
namespace {
struct BlobModel {
  pnb_model m;
  pnb_dense_layer fc, fc_gb, fc_rb;
  pnb_conv1d_layer conv1, conv2;
  pnb_gru_layer gru[5];
  std::vector<std::vector<float>> arrays;
};
bool read_array(FILE *f, size_t expect, std::vector<float> &dst) {
  unsigned long long n = 0;
  if (fread(&n, 8, 1, f) != 1 || n != expect) return false;
  dst.resize(n);
  return fread(dst.data(), 4, n, f) == n;
}
}  // namespace

extern "C" int pnb_model_load_blob(const char *path, pnb_model **out) {
  if (!path || !out) return fail(PNB_ERR_ARG, "NULL argument");
  *out = nullptr;
  FILE *f = fopen(path, "rb");
  if (!f) return fail(PNB_ERR_ARG, "cannot open %s", path);
  int rc = pnb_model_load_stream(f, out);
  fclose(f);
  return rc;
}

extern "C" int pnb_model_load_stream(void *file, pnb_model **out) {
  if (!file || !out) return fail(PNB_ERR_ARG, "NULL argument");
  *out = nullptr;
  FILE *f = static_cast<FILE *>(file);
  const char *path = "weight stream";
  char magic[8];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "PNBW0001", 8) != 0) return fail(PNB_ERR_ARG, "%s is not a PNBW0001 weight file", path);
  BlobModel *b = new BlobModel();
  b->arrays.resize(25);
  int ai = 0;
  bool ok = true;
  auto dense = [&](pnb_dense_layer &l, int M, int N, int act) {
    ok = ok && read_array(f, (size_t)M * N, b->arrays[ai]) && read_array(f, N, b->arrays[ai + 1]);
    l.input_weights = b->arrays[ai].data(); l.bias = b->arrays[ai + 1].data(); l.nb_inputs = M; l.nb_neurons = N; l.activation = act;
    ai += 2;
  };
  auto conv = [&](pnb_conv1d_layer &l, int C, int K, int N, int act) {
    ok = ok && read_array(f, (size_t)C * K * N, b->arrays[ai]) && read_array(f, N, b->arrays[ai + 1]);
    l.input_weights = b->arrays[ai].data(); l.bias = b->arrays[ai + 1].data(); l.nb_inputs = C; l.kernel_size = K; l.nb_neurons = N; l.activation = act;
    ai += 2;
  };
  auto gru = [&](pnb_gru_layer &l, int M, int H) {
    ok = ok && read_array(f, (size_t)M * 3 * H, b->arrays[ai]) && read_array(f, (size_t)H * 3 * H, b->arrays[ai + 1]) &&
         read_array(f, 6 * (size_t)H, b->arrays[ai + 2]);
    l.input_weights = b->arrays[ai].data(); l.recurrent_weights = b->arrays[ai + 1].data(); l.bias = b->arrays[ai + 2].data();
    l.nb_inputs = M; l.nb_neurons = H; l.activation = PNB_ACT_TANH; l.reset_after = 1;
    ai += 3;
  };
  dense(b->fc, 70, 128, PNB_ACT_RELU);
  conv(b->conv1, 128, 5, 512, PNB_ACT_RELU);
  conv(b->conv2, 512, 3, 512, PNB_ACT_TANH);
  for (int i = 0; i < 4; i++) gru(b->gru[i], 512, 512);
  gru(b->gru[4], 1024, 128);
  dense(b->fc_gb, 2560, 34, PNB_ACT_SIGMOID);
  dense(b->fc_rb, 128, 34, PNB_ACT_SIGMOID);
  if (!ok) { delete b; return fail(PNB_ERR_ARG, "%s is truncated or has unexpected layer sizes", path); }
  b->m.fc = &b->fc; b->m.conv1 = &b->conv1; b->m.conv2 = &b->conv2;
  b->m.gru1 = &b->gru[0]; b->m.gru2 = &b->gru[1]; b->m.gru3 = &b->gru[2]; b->m.gru_gb = &b->gru[3]; b->m.gru_rb = &b->gru[4];
  b->m.fc_gb = &b->fc_gb; b->m.fc_rb = &b->fc_rb;
  *out = &b->m;  // first member: the BlobModel is recovered from the pointer in pnb_model_free
  return PNB_OK;
}
extern "C" void pnb_model_free(pnb_model *m) {
  if (m) delete reinterpret_cast<BlobModel *>(m);
}

// ------------------------------------------------------------------------------------------
// the pitch analysis alone (BASELINE.json config 5)
// ------------------------------------------------------------------------------------------
extern "C" int pnb_pitch_only_device(const float *d_pitch_buf, size_t stride, long long n_units, const int *d_prev_period,
                                     const float *d_prev_gain, int *d_period, float *d_corr, float *d_gain, int *d_lag,
                                     void *cuda_stream) {
  if (!d_pitch_buf || !d_period || !d_corr || !d_gain) return fail(PNB_ERR_ARG, "NULL argument");
  if (n_units < 1 || stride < 1728) return fail(PNB_ERR_ARG, "n_units must be >= 1 and stride >= 1728");
  if (n_units > (1ll << 31) - 8) return fail(PNB_ERR_ARG, "n_units above 2^31");
  PitchOnlyArgs a;
  a.buf = d_pitch_buf; a.stride = stride; a.n_units = (long)n_units; a.prev_period = d_prev_period; a.prev_gain = d_prev_gain;
  a.T = d_period; a.corr = d_corr; a.gain = d_gain; a.lag = d_lag;
  if (launch_pitch_only(a, (cudaStream_t)cuda_stream) < 0) return fail(PNB_ERR_CUDA, "pitch_only_kernel could not be configured");
  CK(cudaGetLastError());
  return PNB_OK;
}

// host buffers: staging grows on demand and is kept per host thread (there is no engine to own it)
extern "C" int pnb_pitch_only_host(const float *pitch_buf, size_t stride, long long n_units, const int *prev_period,
                                   const float *prev_gain, int *period, float *corr, float *gain, int *lag) {
  if (!pitch_buf || !period || !corr || !gain) return fail(PNB_ERR_ARG, "NULL argument");
  if (n_units < 1 || stride < 1728) return fail(PNB_ERR_ARG, "n_units must be >= 1 and stride >= 1728");
  struct Staging { float *buf = nullptr; int *i4 = nullptr; float *f3 = nullptr; long long cap = 0; int dev = -1; cudaStream_t st = nullptr; };
  static thread_local Staging g;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (g.cap < n_units || g.dev != dev) {
    if (g.buf) { cudaFree(g.buf); cudaFree(g.i4); cudaFree(g.f3); g.buf = nullptr; }
    if (!g.st || g.dev != dev) CK(cudaStreamCreateWithFlags(&g.st, cudaStreamNonBlocking));
    CK(cudaMalloc((void **)&g.buf, (size_t)n_units * 1728 * sizeof(float)));
    CK(cudaMalloc((void **)&g.i4, (size_t)n_units * 3 * sizeof(int)));
    CK(cudaMalloc((void **)&g.f3, (size_t)n_units * 3 * sizeof(float)));
    g.cap = n_units; g.dev = dev;
  }
  const size_t n = (size_t)n_units;
  int *d_T = g.i4, *d_lag = g.i4 + n, *d_pT = g.i4 + 2 * n;
  float *d_corr = g.f3, *d_gain = g.f3 + n, *d_pg = g.f3 + 2 * n;
  CK(cudaMemcpy2DAsync(g.buf, 1728 * sizeof(float), pitch_buf, stride * sizeof(float), 1728 * sizeof(float), n, cudaMemcpyHostToDevice, g.st));
  if (prev_period) CK(cudaMemcpyAsync(d_pT, prev_period, n * sizeof(int), cudaMemcpyHostToDevice, g.st));
  if (prev_gain) CK(cudaMemcpyAsync(d_pg, prev_gain, n * sizeof(float), cudaMemcpyHostToDevice, g.st));
  int rc = pnb_pitch_only_device(g.buf, 1728, n_units, prev_period ? d_pT : nullptr, prev_gain ? d_pg : nullptr, d_T, d_corr, d_gain,
                                 lag ? d_lag : nullptr, g.st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(period, d_T, n * sizeof(int), cudaMemcpyDeviceToHost, g.st));
  CK(cudaMemcpyAsync(corr, d_corr, n * sizeof(float), cudaMemcpyDeviceToHost, g.st));
  CK(cudaMemcpyAsync(gain, d_gain, n * sizeof(float), cudaMemcpyDeviceToHost, g.st));
  if (lag) CK(cudaMemcpyAsync(lag, d_lag, n * sizeof(int), cudaMemcpyDeviceToHost, g.st));
  CK(cudaStreamSynchronize(g.st));
  return PNB_OK;
}

extern "C" long long pnb_launch_count(const pnb_engine *e) { return e ? e->launches : 0; }
extern "C" int pnb_launches_per_call(const pnb_engine *e, int n_frames) {
  if (!e) return 0;
  // without the history move of advance_line (at most one more launch per call)
  if (e->flags & PNB_NN_TENSOR) {
    const int C = (e->net_sms > 0 && n_frames >= 2 * e->chunk) ? (n_frames + e->chunk - 1) / e->chunk + 4 : 1;  // upper bound
    return 1 + C * (2 + tc_launches_per_chunk(e)) + 1;  // stage_in + per chunk (analysis, network, synthesis) + carry
  }
  return 3 + 7 * ((n_frames + e->f32_chunk - 1) / e->f32_chunk);  // stage_in, analysis, synthesis + per chunk: fc, conv1, conv2, GRU chain, fc_gb, fc_rb, carry
}
// Runtime switch for the chunked overlap schedule (an engine created without it cannot turn it on): profiling a
// kernel class alone on all SMs needs the serial schedule.
extern "C" int pnb_set_overlap(pnb_engine *e, int on) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  if (!e->s_net) return on ? fail(PNB_ERR_ARG, "this engine was created without an SM partition") : PNB_OK;
  CK(cudaSetDevice(e->device));
  CK(cudaDeviceSynchronize());
  if (on) {
    if (e->net_sms == 0) { e->net_sms = e->saved_net_sms; }
  } else if (e->net_sms) {
    e->saved_net_sms = e->net_sms;
    e->net_sms = 0;
  }
  return PNB_OK;
}
extern "C" int pnb_overlap_info(const pnb_engine *e, int *net_sms, int *dsp_sms, int *chunk_hops) {
  if (!e) return fail(PNB_ERR_ARG, "engine is NULL");
  if (net_sms) *net_sms = e->net_sms;
  if (dsp_sms) *dsp_sms = e->dsp_sms;
  if (chunk_hops) *chunk_hops = e->chunk;
  return PNB_OK;
}
extern "C" int pnb_n_streams(const pnb_engine *e) { return e ? e->S : 0; }
extern "C" int pnb_max_frames(const pnb_engine *e) { return e ? e->Fmax : 0; }
