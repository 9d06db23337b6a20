// Tensor-core (tcgen05, TF32) attention forward -- placeholder until the kernel lands.
#include "common.cuh"

int pa_attn_fwd_tc(const pa_attn_fwd_args* a, void* stream) {
  (void)a; (void)stream;
  pa_set_error("pa_attn_fwd: tensor-core path not built yet");
  return PA_ERR_UNSUPPORTED;
}
