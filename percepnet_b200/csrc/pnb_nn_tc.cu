// pnb_nn_tc.cu -- tensor-core (tcgen05) network path.  Placeholder until the UMMA kernels land:
// requesting PNB_NN_TENSOR fails loudly instead of silently running something else.
#include "../../include/percepnet_b200.h"
#include "pnb_engine.h"

int tc_prepare(pnb_engine *, const pnb_model *) { return PNB_ERR_ARG; }
void tc_release(pnb_engine *) {}
int tc_reset(pnb_engine *) { return PNB_OK; }
int tc_step(pnb_engine *, int, cudaStream_t) { return PNB_ERR_ARG; }
int tc_launches_per_step(const pnb_engine *) { return 0; }
