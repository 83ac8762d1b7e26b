#include "slab_common.cuh"
#include "slab_ctx.cuh"
extern "C" int slab_encode(SlabCtx* ctx, SlabEncodeJob* job) { (void)ctx; (void)job; slab_set_error("encoder not built yet"); return -1; }
