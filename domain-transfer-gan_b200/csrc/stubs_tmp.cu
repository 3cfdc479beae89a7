// TEMPORARY stubs while kernels are being brought up (removed once implemented).
#include "common.cuh"
#define STUB(name, ...) extern "C" int name(__VA_ARGS__) { dtg::set_error(#name ": not implemented yet"); return DTG_ERR_UNSUPPORTED; }
extern "C" size_t dtg_norm_workspace_bytes(const dtg_plane*) { return 0; }
STUB(dtg_norm_fwd, const dtg_norm_args*, const dtg_plane*, const dtg_plane*, const float*, const float*, float*, float*, float*, float*, const dtg_plane*, void*)
STUB(dtg_norm_bwd, const dtg_norm_args*, const dtg_plane*, const dtg_plane*, const dtg_plane*, const dtg_plane*, const float*, const float*, float*, float*, float*, float*, const dtg_plane*, const dtg_plane*, void*)
STUB(dtg_cin_affine_fwd, const float*, const float*, const float*, const float*, const float*, int, int, int, float*, float*, void*)
STUB(dtg_cin_affine_bwd, const float*, const float*, const float*, const float*, const float*, const float*, int, int, int, float*, float*, float*, float*, float*, void*)
STUB(dtg_loss_lsgan, const float*, int, int, int, float, float, float*, int, int, const dtg_plane*, void*, void*)
STUB(dtg_loss_l1, const float*, const float*, int, int, int, int, float, int, float*, int, int, const dtg_plane*, void*, void*)
STUB(dtg_grad_sumsq, const float*, size_t, float, float*, void*, void*)
STUB(dtg_adam_clip, float*, float*, float*, float*, size_t, const float*, const float*, const int32_t*, float, void*)
STUB(dtg_step_increment, int32_t*, void*)
