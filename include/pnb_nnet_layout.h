/*
 * pnb_nnet_layout.h -- the weight-layout contract of the hot path, as plain C.
 *
 * These structs are layout-compatible (same field order, same types) with the
 * layer structs and the RNNModel aggregate that the reference declares in
 *   /root/reference/src/nnet.h:44-89      (DenseLayer, GRULayer, Conv1DLayer)
 *   /root/reference/src/nnet_data.h:6-26  (RNNModel: 10 layer pointers)
 * and that dump_percepnet.py emits into the generated src/nnet_data.cpp
 * (dump_percepnet.py:56-126, 146-152).  A `const RNNModel*` obtained from an
 * unmodified, compiled nnet_data.cpp can therefore be passed wherever a
 * `const pnb_model*` is expected (reinterpret the pointer); no conversion step.
 *
 * Array indexing:
 *   dense  : W[j*N + i]                j = input, i = output   (dump_percepnet.py:62)
 *   conv1d : W[(t*C + c)*N + i]        t = tap (0 = oldest)    (dump_percepnet.py:113)
 *   gru    : W[j*3H + g*H + i], U same g = 0:z 1:r 2:n          (dump_percepnet.py:68-85)
 *            bias[6H] = b_iz b_ir b_in b_hz b_hr b_hn          (dump_percepnet.py:78-87)
 */
#ifndef PNB_NNET_LAYOUT_H
#define PNB_NNET_LAYOUT_H

#ifdef __cplusplus
extern "C" {
#endif

/* activation codes, numerically equal to nnet.h:35-39 */
enum {
  PNB_ACT_LINEAR = 0,
  PNB_ACT_SIGMOID = 1,
  PNB_ACT_TANH = 2,
  PNB_ACT_RELU = 3
};

typedef struct pnb_dense_layer {
  const float *bias;          /* [nb_neurons] */
  const float *input_weights; /* [nb_inputs][nb_neurons] */
  int nb_inputs;
  int nb_neurons;
  int activation;
} pnb_dense_layer;

typedef struct pnb_gru_layer {
  const float *bias;              /* [6*nb_neurons] */
  const float *input_weights;     /* [nb_inputs][3*nb_neurons] */
  const float *recurrent_weights; /* [nb_neurons][3*nb_neurons] */
  int nb_inputs;
  int nb_neurons;
  int activation;
  int reset_after;
} pnb_gru_layer;

typedef struct pnb_conv1d_layer {
  const float *bias;          /* [nb_neurons] */
  const float *input_weights; /* [kernel_size][nb_inputs][nb_neurons] */
  int nb_inputs;
  int kernel_size;
  int nb_neurons;
  int activation;
} pnb_conv1d_layer;

typedef struct pnb_model {
  const pnb_dense_layer *fc;      /* 70   -> 128, relu            */
  const pnb_conv1d_layer *conv1;  /* 128x5 -> 512, relu           */
  const pnb_conv1d_layer *conv2;  /* 512x3 -> 512, tanh           */
  const pnb_gru_layer *gru1;      /* 512  -> 512                  */
  const pnb_gru_layer *gru2;      /* 512  -> 512                  */
  const pnb_gru_layer *gru3;      /* 512  -> 512                  */
  const pnb_gru_layer *gru_gb;    /* 512  -> 512                  */
  const pnb_gru_layer *gru_rb;    /* 1024 -> 128                  */
  const pnb_dense_layer *fc_gb;   /* 2560 -> 34, sigmoid          */
  const pnb_dense_layer *fc_rb;   /* 128  -> 34, sigmoid          */
} pnb_model;

#ifdef __cplusplus
}
#endif
#endif /* PNB_NNET_LAYOUT_H */
