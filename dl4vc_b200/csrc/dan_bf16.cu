// dan_bf16.cu — tcgen05 (bf16 operands, fp32 TMEM accumulators) implementation of the DAN forward.
#include "dan_kernels_common.cuh"

int dan_bf16_supported(const dan_model* m) { (void)m; return 0; }
size_t dan_bf16_workspace_bytes(const dan_model* m, int batch) { (void)m; (void)batch; return 0; }
int dan_bf16_forward(dan_model*, const DevInputs&, int, float*, void*, size_t, cudaStream_t) {
  dan_set_error("bf16 path not built");
  return DAN_E_UNSUPPORTED;
}
int dan_bf16_debug_fc_input(dan_model*, int, const void*, float*, cudaStream_t) { return DAN_E_UNSUPPORTED; }
int dan_bf16_pack(dan_model*, const dan_weights*, cudaStream_t) { return DAN_OK; }
void dan_bf16_free(dan_model*) {}
