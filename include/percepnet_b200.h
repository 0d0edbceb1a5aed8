/*
 * percepnet_b200.h -- C-ABI of the B200-native PercepNet enhancement hot path.
 *
 * Plain pointers and sizes only (no torch / C++ types).  Two layers of entry points:
 *
 *  1. the batched-streams engine (pnb_*): S independent 48 kHz streams advance together,
 *     F hops of 480 samples per call.  This is what a host integrates for throughput.
 *  2. the reference's own single-stream API (rnnoise_create / rnnoise_process_frame /
 *     rnnoise_destroy, C++ linkage like /root/reference/src/rnnoise.h:52-60) lives in
 *     librnnoise_b200.so (percepnet_b200/csrc/rnnoise_shim.cpp) on top of this ABI, so the
 *     unmodified /root/reference/src/main.cpp links against it.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   pnb_create          rnnoise_create + rnnoise_init            src/denoise.cpp:252-280
 *                       + check_init table construction           src/denoise.cpp:186-214
 *                       + weight ingest of `const RNNModel*`      src/nnet_data.h:6-26
 *   pnb_process_*       rnnoise_process_frame for S streams x F hops   src/denoise.cpp:508-547
 *                       (compute_frame_features :372, compute_rnn src/rnn.cpp:42,
 *                        pitch_filter :436, interp_band_gain :162, frame_synthesis :352)
 *   pnb_process_*_i16   the int16 I/O conversions of the CLI      src/main.cpp:30-39
 *   pnb_destroy         rnnoise_destroy                           src/denoise.cpp:326-331
 *   pnb_reset           the memset in rnnoise_init                src/denoise.cpp:260
 *
 * Frame contract (src/rnnoise.h:60): hops of exactly 480 float samples at
 * 48 kHz, any amplitude scale; output hop t carries input hop t-6.  Streams are laid out
 * stream-major: sample n of stream s is at base[s*stride + n].
 *
 * All functions return PNB_OK (0) or a negative error code; pnb_last_error() gives text.
 * There is no CPU fallback: without a usable CUDA device pnb_create fails with
 * PNB_ERR_NO_DEVICE.
 */
#ifndef PERCEPNET_B200_H
#define PERCEPNET_B200_H

#include <stddef.h>
#include "pnb_nnet_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PNB_FRAME 480        /* FRAME_SIZE, src/denoise.cpp:19 */
#define PNB_BANDS 34         /* NB_BANDS,   src/denoise.cpp:35 */
#define PNB_FEATURES 70      /* NB_FEATURES, src/denoise.cpp:40 */
#define PNB_LATENCY_FRAMES 6 /* 5 hops look-ahead + 1 hop overlap-add */

enum {
  PNB_OK = 0,
  PNB_ERR_ARG = -1,
  PNB_ERR_CUDA = -2,
  PNB_ERR_NO_DEVICE = -3,
  PNB_ERR_ALLOC = -4,
  PNB_ERR_DOMAIN = -5  /* PNB_NN_TENSOR only: a pre-activation reached |x| >= 8.5e7, where the reference's tansig_approx
                          (src/vec.h:63) overflows its float->int conversion (undefined behaviour; on x86 "tanh" returns
                          its argument).  The tensor path does not reproduce that: outputs since the last pnb_reset are not
                          the reference's.  Reported by the synchronising entry points (pnb_process_host_*, pnb_wait,
                          pnb_check); sticky until pnb_reset.  The fp32 path (PNB_NN_FP32) follows the reference there too. */
};

/* pnb_create flags */
enum {
  PNB_NN_FP32 = 0,       /* network contraction in fp32 FMA (BASELINE.json config 2)                  */
  PNB_NN_TENSOR = 1,     /* network contraction on tcgen05 tensor cores, split-fp16 operands, fp32 accumulate */
  PNB_POSTFILTER = 2,    /* apply the envelope post-filter (src/denoise.cpp:216-250) to g; off in the reference */
  PNB_KEEP_TAPS = 4,     /* keep per-frame intermediates of the last call readable through pnb_read_tap    */
  PNB_CONV_WIDE = 16,    /* PNB_NN_TENSOR only: conv1 / conv2 with three bf16 terms per operand (six tensor-core products)
                            instead of two (three).  Two terms reproduce the double-precision network to 7e-6 on g/r for
                            input at the CLI's amplitude scale (PCM / 32768, |x| <= 1) and to 4e-4 at worst for float
                            input hundreds to thousands of times louder (the reference's own fp32 arithmetic: 3e-5 there);
                            three terms stay within 2.1e-5 at every scale and cost 8 % of the throughput.  The mode is
                            per engine, never per input, so a stream's result does not depend on its neighbours.          */
  PNB_TRAIN_DATA = 8     /* training-data generator (pnb_train_records_*): n_streams = 2 x pairs, no network,
                            no synthesis; model may be NULL; the pnb_process_* entry points are refused      */
};

typedef struct pnb_engine pnb_engine;

/* model: the reference's weight layout (a `const RNNModel*` from a compiled nnet_data.cpp may be
 * passed as is); the weights are copied to the device, the caller's arrays are not retained. */
int pnb_create(pnb_engine **out, int n_streams, int max_frames_per_call, const pnb_model *model,
               unsigned flags, int device);   /* device: CUDA ordinal, or -1 = the calling thread's current device */
void pnb_destroy(pnb_engine *e);
int pnb_reset(pnb_engine *e);

/* Binary weight file (percepnet_b200.weights.PackedModel.save_blob): the arrays of the generated nnet_data.cpp
 * as raw float32, 32 MB instead of 180 MB of C source and no 46 s compile.  The returned model owns its arrays;
 * release it with pnb_model_free (after pnb_create has copied it, at any time). */
int pnb_model_load_blob(const char *path, pnb_model **out);
/* the same from an open stream (a FILE* positioned at the start of the blob; passed as void* to keep <stdio.h> out
 * of this header); the stream is left open, positioned behind the blob */
int pnb_model_load_stream(void *file, pnb_model **out);
void pnb_model_free(pnb_model *m);

/* Host buffers (pageable or pinned).  in/out: n_streams rows of n_frames*480 samples, row strides
 * in elements.  in may equal out.  gr (NULL ok) receives the raw network outputs the reference
 * fwrite()s per frame (src/denoise.cpp:533-534) as [n_frames][n_streams][68] = 34 g then 34 r.
 * Blocking: returns when out (and gr) are complete. */
int pnb_process_host_f32(pnb_engine *e, const float *in, size_t in_stride, float *out, size_t out_stride,
                         int n_frames, float *gr);

/* int16 PCM wire format with the CLI's conversions (src/main.cpp:34,36): x = s/32768.f in,
 * (short)(y*32768) out (C truncation).  The first-hop drop of main.cpp:37-38 is the caller's. */
int pnb_process_host_i16(pnb_engine *e, const short *in, size_t in_stride, short *out, size_t out_stride,
                         int n_frames, float *gr);

/* Pipelined host entry for sustained throughput: enqueues host->device copy, processing and device->host copy
 * on internal streams and returns; the copies of neighbouring calls overlap the kernels.  At most two calls are
 * in flight (a third blocks until the oldest completed).  Buffers must stay valid -- and should be pinned --
 * until pnb_wait() returns.  Calls are processed in submission order (the streams' state advances in order). */
int pnb_submit_host_f32(pnb_engine *e, const float *in, size_t in_stride, float *out, size_t out_stride, int n_frames);
int pnb_submit_host_i16(pnb_engine *e, const short *in, size_t in_stride, short *out, size_t out_stride, int n_frames);
int pnb_wait(pnb_engine *e);
/* Waits for everything enqueued on the engine's internal streams and on `cuda_stream` (the stream the caller passed to
 * pnb_process_device_*), then reports the engine's status: PNB_OK, PNB_ERR_DOMAIN, or PNB_ERR_CUDA. */
int pnb_check(pnb_engine *e, void *cuda_stream);

/* Device buffers on the engine's device; asynchronous on cuda_stream (a cudaStream_t, NULL = default
 * stream).  d_gr (NULL ok) as above, in device memory. */
int pnb_process_device_f32(pnb_engine *e, const float *d_in, size_t in_stride, float *d_out, size_t out_stride,
                           int n_frames, float *d_gr, void *cuda_stream);
int pnb_process_device_i16(pnb_engine *e, const short *d_in, size_t in_stride, short *d_out, size_t out_stride,
                           int n_frames, float *d_gr, void *cuda_stream);

/* Submitted twins of pnb_process_device_*: the call starts when cuda_stream reaches this point (its inputs are ready in
 * that stream's order) but is NOT joined back into it, so a following call's analysis overlaps this call's network and
 * synthesis (the engine orders what they share).  d_out is complete once pnb_flush(e, stream) has made `stream` wait for
 * everything submitted so far, or after pnb_wait / pnb_check on the host.  Calls of one engine are issued in order. */
int pnb_submit_device_f32(pnb_engine *e, const float *d_in, size_t in_stride, float *d_out, size_t out_stride,
                          int n_frames, void *cuda_stream);
int pnb_submit_device_i16(pnb_engine *e, const short *d_in, size_t in_stride, short *d_out, size_t out_stride,
                          int n_frames, void *cuda_stream);
int pnb_flush(pnb_engine *e, void *cuda_stream);

/* Training-data generator: the per-frame loop of the reference's train() (src/denoise.cpp:600-787, as shipped:
 * gains fixed at 1, no biquads, the second file is the finished noisy mixture, TEST defined so g is post-filtered).
 * The engine must have been created with PNB_TRAIN_DATA and n_streams = 2 x n_pairs.  speech/noisy hold n_pairs
 * rows of int16 PCM (row strides in samples); for pair p and frame t of the call the 138 floats
 *   Ey_lookahead[34] Ephaty[34] T pitchcorr g[34] r[34]                      (src/denoise.cpp:761-773)
 * are written to records + p*records_stride + t*138 -- a row is what train() writes to its <output> file for that
 * pair of files.  State carries over between calls, so a long pair of files can be fed in chunks of at most
 * max_frames_per_call frames.  The test_input.pcm / test_output.pcm debug audio of train() is not produced. */
#define PNB_RECORD_FLOATS 138
int pnb_train_records_host(pnb_engine *e, const short *speech, size_t speech_stride, const short *noisy,
                           size_t noisy_stride, int n_frames, float *records, size_t records_stride);
/* Pipelined form (pair it with pnb_wait, like pnb_submit_host_*): the copies of neighbouring calls overlap the
 * kernels; the buffers of a call must stay untouched until pnb_wait returns or two further calls were submitted. */
int pnb_submit_train_records(pnb_engine *e, const short *speech, size_t speech_stride, const short *noisy,
                             size_t noisy_stride, int n_frames, float *records, size_t records_stride);
int pnb_train_records_device(pnb_engine *e, const short *d_speech, size_t speech_stride, const short *d_noisy,
                             size_t noisy_stride, int n_frames, float *d_records, size_t records_stride,
                             void *cuda_stream);

/* Per-stream state for migration between engines / GPUs (the live part of the reference's DenoiseState + RNNState,
 * src/denoise.cpp:71-85, src/nnet_data.h:28-38): input history, overlap-add memory, pitch continuity, the spectra the
 * next five hops will analyse, conv histories and GRU states -- pnb_state_size() bytes, a flat host blob.
 * Both calls wait for the engine to be idle.  A state taken from stream a of one engine and set on stream b of another
 * (any batch size, either network mode, same model) continues the stream exactly where it left off: bit for bit between
 * engines of the same mode.  PNB_TRAIN_DATA engines carry the signal part only. */
size_t pnb_state_size(void);
int pnb_get_state(pnb_engine *e, int stream, void *dst, size_t dst_bytes);
int pnb_set_state(pnb_engine *e, int stream, const void *src, size_t src_bytes);

/* The pitch analysis alone (BASELINE.json config 5; needs no engine): for each of n_units pitch buffers of 1728
 * float samples (unit u at d_pitch_buf + u*stride; the reference's st->pitch_buf, src/denoise.cpp:80,405-411) runs
 * pitch_downsample, pitch_search and remove_doubling (src/pitch.cpp:148-216, 283-386, 423-527) exactly as the hot path
 * does, and writes the final period T in [60, 767], the raw pitch correlation (feature 69) and the pitch gain.
 * d_prev_period / d_prev_gain (NULL = 0) are the previous frame's values remove_doubling's continuity test uses;
 * d_lag (NULL ok) receives pitch_search's lag before the doubling check.  Asynchronous on cuda_stream, current device. */
int pnb_pitch_only_device(const float *d_pitch_buf, size_t stride, long long n_units, const int *d_prev_period,
                          const float *d_prev_gain, int *d_period, float *d_corr, float *d_gain, int *d_lag,
                          void *cuda_stream);

/* Host-buffer form (blocking): copies the buffers to the current device, runs the kernel, copies the results back. */
int pnb_pitch_only_host(const float *pitch_buf, size_t stride, long long n_units, const int *prev_period,
                        const float *prev_gain, int *period, float *corr, float *gain, int *lag);

/* Per-frame intermediates of the LAST call (requires PNB_KEEP_TAPS); copies to host memory.
 * Layouts are [n_frames][n_streams][...]:                                                   */
enum {
  PNB_TAP_FEATURES = 0, /* float[70]  network input (src/denoise.cpp:487-496)                */
  PNB_TAP_PITCH = 1,    /* int[4]     {pitch_search lag, final period T, silence, 0}          */
  PNB_TAP_PITCHF = 2,   /* float[2]   {pitch_corr, pitch_gain}                                */
  PNB_TAP_X = 3,        /* float[800] analysis spectrum bins 0..399 (re,im)                   */
  PNB_TAP_P = 4,        /* float[800] comb-filtered spectrum bins 0..399 (re,im)              */
  PNB_TAP_EX = 5,       /* float[34]  band energy of X                                        */
  PNB_TAP_GR = 6,       /* float[68]  g, r                                                    */
  /* network state after the last hop of the call, layout [n_streams][width] (no PNB_KEEP_TAPS needed): */
  PNB_TAP_NN_C2 = 7,    /* float[512] conv2 output                                            */
  PNB_TAP_NN_H0 = 8,    /* +0..4: float[512|128] states of gru1, gru2, gru3, gru_gb, gru_rb   */
  PNB_TAP_G_USED = 13   /* float[34]  band gains as applied: g, post-filtered under PNB_POSTFILTER (PNB_KEEP_TAPS) */
};
int pnb_read_tap(pnb_engine *e, int what, void *dst, size_t dst_bytes);

/* Per-kernel-class device timing.  While enabled, every launch of a pnb_process_* call is bracketed by
 * CUDA events on the launching stream; pnb_profile_read waits for them and returns the accumulated
 * milliseconds and launch counts per class since the last read (arrays of PNB_NUM_KERNEL_CLASSES). */
enum {
  PNB_K_STAGE_IN = 0, PNB_K_ANALYSIS = 1, PNB_K_FC = 2, PNB_K_GEMM_F32 = 3, PNB_K_F32_CARRY = 4,  /* fp32 path: the per-chunk carry (was the GRU gate kernel before the gates were fused) */
 
  PNB_K_SYNTHESIS = 5, PNB_K_SLIDE = 6, PNB_K_TC_GEMM = 7, PNB_K_TC_AUX = 8, PNB_K_LABELS = 9,
  PNB_NUM_KERNEL_CLASSES = 10
};
int pnb_profile_enable(pnb_engine *e, int on);
int pnb_profile_read(pnb_engine *e, double *ms, long long *counts);
/* The same records as a timeline: launch i of the profiled calls as (class, start ms, end ms) relative to the first
 * launch, in launch order (launches on different internal streams overlap).  Returns the count written (<= cap). */
int pnb_profile_timeline(pnb_engine *e, int *cls, double *t0_ms, double *t1_ms, int cap);
const char *pnb_kernel_class_name(int cls);

/* Number of kernels this library has launched on behalf of e since creation. */
long long pnb_launch_count(const pnb_engine *e);
/* Kernel launches one pnb_process_* call with n_frames hops issues (one more on the calls that also move the
 * streams' history back to the start of their rows, at most every call, typically every eighth). */
int pnb_launches_per_call(const pnb_engine *e, int n_frames);

/* How long calls are scheduled on this engine (PNB_NN_TENSOR engines with many streams): calls of at least two
 * chunks of `chunk_hops` hops run the network of chunk k on `net_sms` SMs while analysis of chunk k+1 and synthesis
 * of chunk k-1 run on the other `dsp_sms` SMs (two green contexts).  net_sms = 0: every call runs its kernels one
 * after the other on all SMs.  Environment: PNB_OVERLAP=0 disables, =2 forces it for small batches; PNB_NET_SMS,
 * PNB_CHUNK tune it (read by pnb_create).  Results do not depend on the schedule. */
int pnb_overlap_info(const pnb_engine *e, int *net_sms, int *dsp_sms, int *chunk_hops);
/* on = 0: run every call serially on all SMs (e.g. to time one kernel class alone); on = 1: back to the partition the
 * engine was created with.  Waits for the engine to be idle. */
int pnb_set_overlap(pnb_engine *e, int on);
int pnb_n_streams(const pnb_engine *e);
int pnb_max_frames(const pnb_engine *e);
const char *pnb_last_error(void);
const char *pnb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PERCEPNET_B200_H */
