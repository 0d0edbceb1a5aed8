// rnnoise_shim.cpp -- the reference's own single-stream API on top of the batched engine.
//
// Exports, with C++ linkage exactly like /root/reference/src/rnnoise.h:52-60 (the reference sources are
// all .cpp and the header has no extern "C"), the symbols an unmodified
// /root/reference/src/main.cpp imports:
//   int rnnoise_get_size();                                   src/denoise.cpp:348
//   int rnnoise_init(DenoiseState*, RNNModel*);               src/denoise.cpp:259
//   DenoiseState* rnnoise_create(RNNModel*);                  src/denoise.cpp:252
//   void rnnoise_destroy(DenoiseState*);                      src/denoise.cpp:326
//   float rnnoise_process_frame(DenoiseState*, float*, const float*, FILE*);   src/denoise.cpp:508
// A stream handle is a batch of one; it exists for drop-in correctness (percepNet_run), throughput
// comes from the pnb_* batch API.  model == NULL selects `percepnet_model_orig`, the symbol the
// generated src/nnet_data.cpp defines (dump_percepnet.py:149); it is referenced weakly so that hosts
// which always pass a model do not need to link one.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/percepnet_b200.h"

struct RNNModel;      // layout == pnb_model (include/pnb_nnet_layout.h)
struct DenoiseState { // opaque to callers (rnnoise.h:49)
  pnb_engine *engine;
  float gr[68];
};

extern const RNNModel percepnet_model_orig __attribute__((weak));

// Declared by the reference (src/rnnoise.h:62-64) but defined nowhere in it: here they read the
// binary weight file of pnb_model_load_blob, so a host can run without compiling nnet_data.cpp.
RNNModel *rnnoise_model_from_file(FILE *f) {
  if (!f) return NULL;
  pnb_model *m = NULL;
  if (pnb_model_load_stream(f, &m) != PNB_OK) {
    fprintf(stderr, "rnnoise_model_from_file: %s\n", pnb_last_error());
    return NULL;
  }
  return reinterpret_cast<RNNModel *>(m);
}
void rnnoise_model_free(RNNModel *model) { pnb_model_free(reinterpret_cast<pnb_model *>(model)); }

int rnnoise_get_size() { return (int)sizeof(DenoiseState); }

int rnnoise_init(DenoiseState *st, RNNModel *model) {
  memset(st, 0, sizeof *st);
  const RNNModel *m = model ? model : &percepnet_model_orig;
  if (!m) {
    fprintf(stderr, "rnnoise_init: no model given and percepnet_model_orig is not linked\n");
    return -1;
  }
  // PNB_SHIM_NN=tensor runs the network on the tensor cores (2.8x lower latency per frame; g/r within 1e-5 of
  // the reference instead of 5e-7); the default keeps the fp32 network, which also reproduces the reference
  // outside tansig_approx's defined range
  const char *nn = getenv("PNB_SHIM_NN");
  const unsigned flags = (nn && strcmp(nn, "tensor") == 0) ? PNB_NN_TENSOR : PNB_NN_FP32;
  // PNB_DEVICE=<ordinal> picks the GPU; otherwise the calling thread's current CUDA device (device 0 in a fresh process)
  const char *dv = getenv("PNB_DEVICE");
  int rc = pnb_create(&st->engine, 1, 1, reinterpret_cast<const pnb_model *>(m), flags, dv ? atoi(dv) : -1);
  if (rc != PNB_OK) {
    fprintf(stderr, "rnnoise_init: %s\n", pnb_last_error());
    return rc;
  }
  return 0;
}

DenoiseState *rnnoise_create(RNNModel *model) {
  DenoiseState *st = (DenoiseState *)malloc(rnnoise_get_size());
  if (!st) return NULL;
  if (rnnoise_init(st, model) != 0) {  // no CPU fallback: the caller gets NULL and the reason on stderr
    free(st);
    fprintf(stderr, "rnnoise_create: cannot create the CUDA engine (there is no CPU path)\n");
    return NULL;
  }
  return st;
}

void rnnoise_destroy(DenoiseState *st) {
  if (!st) return;
  pnb_destroy(st->engine);
  free(st);
}

float rnnoise_process_frame(DenoiseState *st, float *out, const float *in, FILE *f_feature) {
  if (!st || !st->engine) {
    fprintf(stderr, "rnnoise_process_frame: no engine (rnnoise_create failed?)\n");
    abort();  // the reference would dereference NULL here; a frame cannot be skipped silently
  }
  int rc = pnb_process_host_f32(st->engine, in, PNB_FRAME, out, PNB_FRAME, 1, st->gr);
  if (rc != PNB_OK) {
    fprintf(stderr, "rnnoise_process_frame: %s\n", pnb_last_error());
    abort();
  }
  if (f_feature) {  // src/denoise.cpp:533-534
    fwrite(st->gr, sizeof(float), 34, f_feature);
    fwrite(st->gr + 34, sizeof(float), 34, f_feature);
  }
  return 0;  // src/denoise.cpp:546
}
