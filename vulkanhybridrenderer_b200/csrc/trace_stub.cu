// temporary stubs until bvh_build.cu / trace_kernels.cu land
#include "vhr_internal.h"
namespace vhr {
int launch_trace_rays(vhr_context *, uint32_t, uint32_t) { return fail(VHR_ERR_STATE, "ray pass not built yet"); }
int launch_gbuffer(vhr_context *, uint32_t, uint32_t) { return fail(VHR_ERR_STATE, "gbuffer pass not built yet"); }
int launch_trace_explicit(vhr_context *, const float *, uint32_t, int, float *, uint32_t *, float *) { return fail(VHR_ERR_STATE, "ray pass not built yet"); }
int build_bvh(vhr_context *) { return VHR_OK; }
void free_bvh(vhr_context *) {}
}
