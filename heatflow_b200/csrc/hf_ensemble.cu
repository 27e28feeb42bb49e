// heatflow_b200 - batched multi-RHS ensemble (parameter_sweep) - placeholder until the batched kernels land.
#include "hf_ctx.cuh"

struct EnsState {};
void hf_ens_free(hf_ctx* c) {
  delete c->ens;
  c->ens = nullptr;
}
extern "C" int hf_ens_create(hf_ctx*, int32_t, const double*, const double*, int32_t) { return hf_fail(HF_ERR_STATE, "ensemble not built"); }
extern "C" int hf_ens_run(hf_ctx*, int32_t, const double*, double, int32_t, const int32_t*, double*, int32_t*) { return hf_fail(HF_ERR_STATE, "ensemble not built"); }
extern "C" int hf_ens_get_state(hf_ctx*, double*) { return hf_fail(HF_ERR_STATE, "ensemble not built"); }
extern "C" int hf_ens_destroy(hf_ctx*) { return HF_OK; }
