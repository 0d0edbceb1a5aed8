/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin extern "C" veneer over the *unmodified* reference sources, compiled
 * from where they lie under /root/reference/src by oracle/Makefile into
 * oracle/_ref/libpercepnet_ref.so.  It exists so that tests, the golden-vector
 * generator and bench.py's CPU-baseline leg can drive the real reference:
 *
 *   - it supplies the symbol `percepnet_model_orig` that the reference expects
 *     from the generated src/nnet_data.cpp (src/denoise.cpp:50,267), but backed
 *     by weight arrays handed in at run time (same RNNModel layout,
 *     src/nnet_data.h:6-26), so the 180 MB generated file need not be compiled
 *     for every weight set;
 *   - it forwards to the reference's public API (src/rnnoise.h:52-68) and to the
 *     non-static stage functions (src/pitch.h:41-48, src/celt_lpc.h, src/nnet.h,
 *     src/kiss_fft.h, src/denoise.cpp:89-182,436) for per-stage taps.
 *
 * Nothing in here restates an algorithm; every number it returns is computed
 * by reference code.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <vector>

#include "rnnoise.h"   /* reference: pulls nnet_data.h / nnet.h */
#include "pitch.h"
#include "celt_lpc.h"
#include "kiss_fft.h"
#include "erbband.h"

#include "../include/pnb_nnet_layout.h"

/* reference globals / non-static functions that have no header declaration */
extern ERBBand *erb_band;                                            /* src/denoise.cpp:87  */
void compute_band_energy(float *bandE, const kiss_fft_cpx *X);      /* src/denoise.cpp:89  */
void compute_band_corr(float *bandE, const kiss_fft_cpx *X, const kiss_fft_cpx *P); /* :125 */
void interp_band_gain(float *g, const float *bandE);                /* src/denoise.cpp:162 */
void pitch_filter(kiss_fft_cpx *X, const kiss_fft_cpx *P, const float *Ex, const float *Ep,
                  const float *Exp, const float *g, const float *r); /* src/denoise.cpp:436 */

/* ---- run-time backed replacement for the generated nnet_data.cpp ---------- */
static DenseLayer  L_fc, L_fc_gb, L_fc_rb;
static Conv1DLayer L_conv1, L_conv2;
static GRULayer    L_gru1, L_gru2, L_gru3, L_gru_gb, L_gru_rb;

extern const RNNModel percepnet_model_orig = {
  &L_fc, &L_conv1, &L_conv2, &L_gru1, &L_gru2, &L_gru3, &L_gru_gb, &L_gru_rb, &L_fc_gb, &L_fc_rb
};

static void put_dense(DenseLayer &d, const pnb_dense_layer *s) {
  d.bias = s->bias; d.input_weights = s->input_weights;
  d.nb_inputs = s->nb_inputs; d.nb_neurons = s->nb_neurons; d.activation = s->activation;
}
static void put_conv(Conv1DLayer &d, const pnb_conv1d_layer *s) {
  d.bias = s->bias; d.input_weights = s->input_weights; d.nb_inputs = s->nb_inputs;
  d.kernel_size = s->kernel_size; d.nb_neurons = s->nb_neurons; d.activation = s->activation;
}
static void put_gru(GRULayer &d, const pnb_gru_layer *s) {
  d.bias = s->bias; d.input_weights = s->input_weights; d.recurrent_weights = s->recurrent_weights;
  d.nb_inputs = s->nb_inputs; d.nb_neurons = s->nb_neurons; d.activation = s->activation;
  d.reset_after = s->reset_after;
}

int train(int argc, char **argv); /* denoise.cpp:600 */

extern "C" {

/* The arrays stay owned by the caller and must outlive every state created afterwards. */
void ref_set_model(const pnb_model *m) {
  put_dense(L_fc, m->fc);       put_conv(L_conv1, m->conv1);  put_conv(L_conv2, m->conv2);
  put_gru(L_gru1, m->gru1);     put_gru(L_gru2, m->gru2);     put_gru(L_gru3, m->gru3);
  put_gru(L_gru_gb, m->gru_gb); put_gru(L_gru_rb, m->gru_rb);
  put_dense(L_fc_gb, m->fc_gb); put_dense(L_fc_rb, m->fc_rb);
}

void *ref_create(void) { return rnnoise_create(NULL); }            /* src/denoise.cpp:252 */
void ref_destroy(void *st) { rnnoise_destroy((DenoiseState *)st); } /* src/denoise.cpp:326 */
int ref_state_size(void) { return rnnoise_get_size(); }

/* One frame through src/denoise.cpp:508.  gr (may be NULL) receives the 34 g + 34 r
 * floats the reference fwrite()s to f_feature (src/denoise.cpp:533-534). */
void ref_process_frame(void *st, float *out, const float *in, float *gr) {
  static FILE *devnull = NULL;
  if (gr) {
    float tmp[72]; /* a little slack: fmemopen("w") wants room for a trailing NUL */
    FILE *f = fmemopen(tmp, sizeof tmp, "wb");
    rnnoise_process_frame((DenoiseState *)st, out, in, f);
    fclose(f);
    memcpy(gr, tmp, 68 * sizeof(float));
  } else {
    if (!devnull) devnull = fopen("/dev/null", "wb");
    rnnoise_process_frame((DenoiseState *)st, out, in, devnull);
  }
}

/* n_frames through the float C API on one stream; in/out hold n_frames*480 floats;
 * gr (may be NULL) n_frames*68. */
void ref_process_stream(void *st, float *out, const float *in, int n_frames, float *gr) {
  for (int t = 0; t < n_frames; t++) {
    float x[480];
    memcpy(x, in + 480 * t, sizeof x);
    ref_process_frame(st, x, x, gr ? gr + 68 * t : NULL);   /* aliased like src/main.cpp:35 */
    memcpy(out + 480 * t, x, sizeof x);
  }
}

/* The I/O conversions of src/main.cpp:30-39 around the C API: int16 in, /32768,
 * process, *32768 with C truncation, first output frame dropped.
 * out16 receives (n_frames-1)*480 samples. */
void ref_run_pcm16(const short *in16, int n_frames, short *out16, float *gr) {
  DenoiseState *st = rnnoise_create(NULL);
  int first = 1;
  for (int t = 0; t < n_frames; t++) {
    float x[480];
    short tmp[480];
    for (int i = 0; i < 480; i++) x[i] = ((float)in16[480 * t + i]) / 32768.f;
    ref_process_frame(st, x, x, gr ? gr + 68 * t : NULL);
    for (int i = 0; i < 480; i++) tmp[i] = x[i] * 32768;
    if (!first) memcpy(out16 + 480 * (t - 1), tmp, sizeof tmp);
    first = 0;
  }
  rnnoise_destroy(st);
}

/* Multi-stream CPU timing leg: streams independent, one state each,
 * OpenMP over streams when compiled with -fopenmp.  in/out: [n_streams][n_frames*480]. */
void ref_process_streams_omp(int n_streams, int n_frames, const float *in, float *out, int n_threads) {
  std::vector<void *> st(n_streams);
  for (int s = 0; s < n_streams; s++) st[s] = ref_create();
  { float z[480] = {0}, o[480]; void *w = ref_create(); ref_process_frame(w, o, z, NULL); ref_destroy(w); }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
  for (int s = 0; s < n_streams; s++)
    ref_process_stream(st[s], out + (size_t)s * n_frames * 480, in + (size_t)s * n_frames * 480, n_frames, NULL);
  for (int s = 0; s < n_streams; s++) ref_destroy(st[s]);
  (void)n_threads;
}

/* ------------------------------ stage taps -------------------------------- */
void ref_erb_borders(int *out34) { for (int i = 0; i < 34; i++) out34[i] = erb_band->nfftborder[i]; }

void ref_fft960(const float *in_ri, float *out_ri) {   /* src/kiss_fft.cpp:566 */
  static kiss_fft_state *k = NULL;
  if (!k) k = opus_fft_alloc_twiddles(960, NULL, NULL, NULL, 0);
  opus_fft_c(k, (const kiss_fft_cpx *)in_ri, (kiss_fft_cpx *)out_ri);
}
void ref_band_energy(float *bandE, const float *X_ri) { compute_band_energy(bandE, (const kiss_fft_cpx *)X_ri); }
void ref_band_corr(float *bandE, const float *X_ri, const float *P_ri) {
  compute_band_corr(bandE, (const kiss_fft_cpx *)X_ri, (const kiss_fft_cpx *)P_ri);
}
/* g must hold 481 floats and is used exactly as the reference uses it (only the first
 * 481 BYTES are cleared, src/denoise.cpp:164) */
void ref_interp_band_gain(float *g481, const float *bandE) { interp_band_gain(g481, bandE); }
void ref_pitch_filter(float *X_ri, const float *P_ri, const float *g, const float *r) {
  float Ex[34] = {0}, Ep[34] = {0}, Exp[34] = {0};
  pitch_filter((kiss_fft_cpx *)X_ri, (const kiss_fft_cpx *)P_ri, Ex, Ep, Exp, g, r);
}
void ref_pitch_downsample(const float *pitch_buf1728, float *lp864) {   /* src/pitch.cpp:148 */
  float *pre[1];
  pre[0] = (float *)pitch_buf1728;
  pitch_downsample(pre, lp864, 1728, 1);
}
void ref_pitch_search(float *lp864, int *pitch, float *corr) {          /* src/pitch.cpp:283 */
  pitch_search(lp864 + 384, lp864, 960, 588, pitch, corr);
}
float ref_remove_doubling(float *lp864, int *T0, int prev_period, float prev_gain) { /* :424 */
  return remove_doubling(lp864, 768, 60, 960, T0, prev_period, prev_gain);
}
void ref_autocorr_lpc(const float *x, int n, float *ac5, float *lpc4) {  /* src/celt_lpc.cpp:198,37 */
  _celt_autocorr(x, ac5, NULL, 0, 4, n);
  float ac[5];
  memcpy(ac, ac5, sizeof ac);
  _celt_lpc(lpc4, ac, 4);
}
void ref_pitch_xcorr(const float *x, const float *y, float *xcorr, int len, int max_pitch) {
  celt_pitch_xcorr(x, y, xcorr, len, max_pitch);                        /* src/pitch.cpp:218 */
}

/* NN taps on the currently installed model (src/nnet.cpp:105,120,182; src/rnn.cpp:42) */
void ref_dense(int which, float *out, const float *in) {
  const DenseLayer *l = which == 0 ? &L_fc : which == 1 ? &L_fc_gb : &L_fc_rb;
  compute_dense(l, out, in);
}
void ref_conv1d(int which, float *out, float *mem, const float *in) {
  compute_conv1d(which == 0 ? &L_conv1 : &L_conv2, out, mem, in);
}
void ref_gru(int which, float *state, const float *in) {
  const GRULayer *l = which == 0 ? &L_gru1 : which == 1 ? &L_gru2 : which == 2 ? &L_gru3
                    : which == 3 ? &L_gru_gb : &L_gru_rb;
  compute_gru(l, state, in);
}
/* generic-layer variants for the reference's own toy known-answer vectors
 * (tests/nnet_data_test.h, tests/testnnet.cpp:19-66) */
void ref_dense_layer(const pnb_dense_layer *s, float *out, const float *in) {
  DenseLayer d; put_dense(d, s); compute_dense(&d, out, in);
}
void ref_conv1d_layer(const pnb_conv1d_layer *s, float *out, float *mem, const float *in) {
  Conv1DLayer d; put_conv(d, s); compute_conv1d(&d, out, mem, in);
}
void ref_gru_layer(const pnb_gru_layer *s, float *state, const float *in) {
  GRULayer d; put_gru(d, s); compute_gru(&d, state, in);
}

/* compute_rnn on a caller-held state block: 512+1024+4*512+128 floats laid out
 * [conv1 mem 512][conv2 mem 1024][gru1][gru2][gru3][gru_gb][gru_rb 128] */
void ref_compute_rnn(float *state, float *gains, float *strengths, const float *features) {
  RNNState r;
  memset(&r, 0, sizeof r);
  r.model = &percepnet_model_orig;
  r.first_conv1d_state = state;
  r.second_conv1d_state = state + 512;
  r.gru1_state = state + 512 + 1024;
  r.gru2_state = r.gru1_state + 512;
  r.gru3_state = r.gru2_state + 512;
  r.gb_gru_state = r.gru3_state + 512;
  r.rb_gru_state = r.gb_gru_state + 512;
  compute_rnn(&r, gains, strengths, features);
}

/* The reference's training-data generator, denoise.cpp:600 (non-static, compiled into every build of
 * denoise.cpp): <speech.pcm> <noisy.pcm> <count> <out.f32>. */
int ref_train_files(const char *speech, const char *noisy, int count, const char *out) {
  char cnt[32];
  snprintf(cnt, sizeof cnt, "%d", count);
  char *argv[5] = {(char *)"percepNet", (char *)speech, (char *)noisy, cnt, (char *)out};
  /* train() also drops test_input.pcm / test_output.pcm into the working directory (TEST is defined at
   * denoise.cpp:45-46): run it from the directory of the output file (all paths must be absolute). */
  char cwd[4096], dir[4096];
  if (!getcwd(cwd, sizeof cwd)) return -1;
  snprintf(dir, sizeof dir, "%s", out);
  char *slash = strrchr(dir, '/');
  if (!slash) return -1;
  *slash = 0;
  if (chdir(dir) != 0) return -1;
  int rc = train(5, argv);
  if (chdir(cwd) != 0) return -1;
  return rc;
}

} /* extern "C" */
