// pnb_engine.h -- the engine object behind the opaque pnb_engine handle.
#pragma once
#include <stddef.h>
#include <vector>
#include "pnb_kernels.h"
#include "../../include/percepnet_b200.h"
#include "../../include/pnb_nnet_layout.h"

struct pnb_tc_state;  // tensor-core path (pnb_nn_tc.cu)

struct pnb_engine {
  int S = 0, Fmax = 0, device = 0, sm_count = 0;
  unsigned flags = 0;
  cudaStream_t stream = nullptr;  // used by the host-buffer entry points
  pnb::Tables *d_tab = nullptr;

  // weights in the reference's layout
  struct { float *W = nullptr, *b = nullptr; } fc, conv1, conv2, fc_gb, fc_rb;
  struct { float *W = nullptr, *U = nullptr, *b = nullptr; int M = 0, H = 0; } gru[5];  // gru1..3, gru_gb, gru_rb
  int act_fc = 0, act_conv1 = 0, act_conv2 = 0, act_gb = 0, act_rb = 0;

  // per-stream signal state
  // [S][5280 + c*Fmax*480], c <= kLineCalls: each stream's history line.  A call appends its hops behind the history and
  // the next call simply starts F*480 samples further on (line_off); only when the row is used up are the last
  // 5280 samples moved back to its start, once every c calls instead of after every call.
  float *d_pcm = nullptr;
  size_t pcm_stride = 0;
  size_t line_off = 0;
  float *d_synth = nullptr;
  int *d_last_period = nullptr;
  float *d_last_gain = nullptr;

  // per-call intermediates [F][S][...]
  float *d_feat = nullptr;
  float2 *d_zring = nullptr;  // [ring][S][400] carried spectra (see pnb_dsp.cu), slot = hop % ring
  float *d_ering = nullptr;   // [ring][S][34]
  int ring = 0;
  float2 *d_P = nullptr;
  float *d_Ex = nullptr;
  float *d_raw = nullptr;     // [F][S][68] training-data mode only
  float *d_records = nullptr; // [S/2][Fmax][138] staging of pnb_train_records_host
  unsigned char *d_sil = nullptr;
  float *d_gr = nullptr;
  int *d_tap_pitch = nullptr;
  float *d_tap_pitchf = nullptr;
  float *d_tap_g = nullptr;   // [F][S][34] gains as applied (PNB_KEEP_TAPS)

  // network state (fp32 path): conv rings, GRU states (ping-pong), scratch sums
  float *ring_fc = nullptr;  // [chunk+4][S][128] fc outputs: slots 0..3 the last four hops (oldest first), 4.. the chunk's
  float *ring_c1 = nullptr;  // [chunk+2][S][512] conv1 outputs, likewise
  float *c2 = nullptr;       // [S][512]
  float *h[5][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  int par[5] = {0, 0, 0, 0, 0};
  float *c2_all = nullptr;   // fp32 path: [chunk][S][512] conv2 outputs of a chunk of hops
  unsigned *f32_cnt = nullptr;  // fp32 path: [5][f32_rb] dependency counters of the persistent GRU chain
  int f32_chunk = 8;            // hops per chunk of the fp32 network = min(max_frames, kF32ChainMaxHops)
  int f32_rt = 8, f32_rb = 0;   // rows per thread of the chain kernel (8: 128-stream blocks, 1: 16-stream blocks); stream blocks
  long hop = 0;  // hops processed since reset

  // staging for the host-buffer entry points
  float *d_hin = nullptr, *d_hout = nullptr;
  short *d_hin16 = nullptr, *d_hout16 = nullptr;

  // pipelined host entry (pnb_submit_host_*): two slots of device staging, three streams
  cudaStream_t s_in = nullptr, s_out = nullptr;
  void *pipe_in[2] = {nullptr, nullptr}, *pipe_out[2] = {nullptr, nullptr};
  size_t pipe_elem = 0;  // element size the staging was allocated for
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  long long submitted = 0;

  pnb_tc_state *tc = nullptr;
  int tc_sms = 0;  // SMs the persistent network kernels may occupy (all of them, or the network's share of a partition)
  // Overlap of the DSP kernels with the network inside long calls (tensor mode): the call is cut into chunks of hops,
  // analysis / synthesis of neighbouring chunks run on s_dsp while the network of a chunk runs on s_net, each stream
  // in its own green context = its own disjoint set of SMs.  The network phase alone is capped by the board's power
  // limit (it clocks down to ~1.4 GHz) while the DSP phase leaves that budget unused; side by side they even out.
  void *green_net = nullptr, *green_dsp = nullptr;  // CUgreenCtx
  cudaStream_t s_net = nullptr, s_dsp = nullptr, s_syn = nullptr;  // s_syn: synthesis, in the DSP partition
  int net_sms = 0, dsp_sms = 0, chunk = 8, saved_net_sms = 0;
  bool ramp = true;              // short chunks at both ends of a call (PNB_RAMP=0: all chunks equal)
  cudaEvent_t ev_fork = nullptr, ev_join_net = nullptr, ev_join_dsp = nullptr, ev_join_syn = nullptr;
  std::vector<cudaEvent_t> ev_ana, ev_net;
  // Calls overlap each other as well (pnb_submit_*): the analysis of call i+1 starts while the network and synthesis
  // of call i are still running.  What call i+1 overwrites (hop-slot buffers, spectrum ring, fc slots) is guarded by
  // the synthesis-done events of call i, kept per chunk and double-buffered by call parity.
  std::vector<cudaEvent_t> ev_syn[2];
  std::vector<int> prev_start;   // first hop of every chunk of the previous chunked call (empty: it was not chunked)
  int syn_par = 0;               // which ev_syn[] the previous chunked call recorded into
  bool unjoined = false;         // submitted work is in flight on the engine's own streams, not joined into any caller stream
  int last_frames = 0;
  long long launches = 0;
  // status word on the device: bit 0 = an activation left the domain in which the reference's tansig_approx is
  // defined (tensor path only; reported as PNB_ERR_DOMAIN by the synchronising entry points, cleared by pnb_reset)
  int *d_status = nullptr;
  int *h_status = nullptr;  // pinned host copy, refreshed by the synchronising entry points
  // a call failed after it had started to enqueue work: the streams' state is undefined until pnb_reset
  bool poisoned = false;

  // optional per-kernel-class timing (pnb_profile_enable): CUDA events around every launch
  bool profiling = false;
  struct ProfRec { int cls; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[PNB_NUM_KERNEL_CLASSES] = {0};
  long long prof_n[PNB_NUM_KERNEL_CLASSES] = {0};

  const float *tansig() const {
    return reinterpret_cast<const float *>(reinterpret_cast<const char *>(d_tab) + offsetof(pnb::Tables, tansig));
  }
};

// tensor-core path hooks (pnb_nn_tc.cu); return PNB_OK / negative error; the call phases return their launch count
int tc_prepare(pnb_engine *e, const pnb_model *model);
void tc_release(pnb_engine *e);
int tc_reset(pnb_engine *e);
// the network of hops [h0, h0+n) of a call of F hops: hop-parallel front, the GRU chain, hop-parallel output layers
int tc_fc(pnb_engine *e, int h0, int n, cudaStream_t st);
int tc_front(pnb_engine *e, int h0, int n, int F, cudaStream_t st);
int tc_gru_chain(pnb_engine *e, int h0, int n, cudaStream_t st);
int tc_out(pnb_engine *e, int h0, int n, cudaStream_t st);
int tc_carry(pnb_engine *e, int F, cudaStream_t st);       // once per call, after its last tc_out
int tc_launches_per_chunk(const pnb_engine *e);
// pnb_get_state / pnb_set_state: conv histories as fp32 sums and as the raw bf16 term pairs [2 terms][slots][width]
int tc_get_stream_hist(pnb_engine *e, int s, float *fc_hist, float *c1_hist, unsigned short *fc_terms, unsigned short *c1_terms);
int tc_set_stream_hist(pnb_engine *e, int s, const float *fc_hist, const float *c1_hist, const float *h,
                       const unsigned short *fc_terms, const unsigned short *c1_terms);

// RAII marker used by the launch schedule: records an event pair around a launch when profiling
struct ProfScope {
  pnb_engine *e; int idx;
  ProfScope(pnb_engine *eng, int cls, cudaStream_t st);
  ~ProfScope();
  cudaStream_t st;
};
