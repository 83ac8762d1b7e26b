/* slab_ctx.cu - device context: stream, arenas, error text. No codec logic. */
#include "slab_ctx.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

/* per thread: the pipeline workers of one handle (and of different handles) fail independently */
static thread_local char g_error[512] = "";

void slab_set_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

extern "C" const char* slab_last_error(void) { return g_error; }
extern "C" void slab_set_error_text(const char* text) { slab_set_error("%s", text); }

extern "C" int slab_is_hostsim(void)
{
#ifdef SLAB_EMUL
  return 1;
#else
  return 0;
#endif
}

extern "C" SlabCtx* slab_ctx_create(void)
{
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    slab_set_error("sla_b200: no CUDA device available (%s); this library has no CPU path",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return NULL;
  }
  SlabCtx* ctx = (SlabCtx*)calloc(1, sizeof(SlabCtx));
  if (!ctx) return NULL;
  if (cudaGetDevice(&ctx->device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    slab_set_error("sla_b200: cannot create a CUDA stream");
    free(ctx);
    return NULL;
  }
  for (int i = 0; i < 4; i++) cudaEventCreate(&ctx->ev[i]);
  for (int i = 0; i < 2; i++) cudaEventCreate(&ctx->ev_span[i]);
  return ctx;
}

extern "C" void slab_ctx_destroy(SlabCtx* ctx)
{
  if (!ctx) return;
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < SLAB_NUM_ARENAS; i++) if (ctx->arena[i]) cudaFree(ctx->arena[i]);
  for (uint32_t i = 0; i < ctx->num_windows; i++) cudaFree(ctx->windows[i].dev);
  free(ctx->windows);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  free(ctx->host_scratch);
  for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 2; i++) cudaEventDestroy(ctx->ev_span[i]);
  for (int i = 0; i < SLAB_MAX_PROF; i++)
    if (ctx->prof_ev[i][0]) { cudaEventDestroy(ctx->prof_ev[i][0]); cudaEventDestroy(ctx->prof_ev[i][1]); }
  cudaStreamDestroy(ctx->stream);
  free(ctx);
}

void* slab_arena(SlabCtx* ctx, int slot, size_t bytes)
{
  if (bytes == 0) bytes = 16;
  if (ctx->arena_bytes[slot] >= bytes) return ctx->arena[slot];
  if (ctx->arena[slot]) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->arena[slot]);
    ctx->arena[slot] = NULL; ctx->arena_bytes[slot] = 0;
  }
  size_t want = bytes + bytes / 8 + 256;       /* slack so that similar-sized calls do not realloc */
  void* p = NULL;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    slab_set_error("sla_b200: cudaMalloc(%zu) failed for arena %d", want, slot);
    return NULL;
  }
  ctx->arena[slot] = p; ctx->arena_bytes[slot] = want;
  return p;
}

void* slab_pinned(SlabCtx* ctx, size_t bytes)
{
  if (ctx->pinned_bytes >= bytes) return ctx->pinned;
  if (ctx->pinned) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->pinned); ctx->pinned = NULL; ctx->pinned_bytes = 0; }
  if (cudaMallocHost(&ctx->pinned, bytes + 4096) != cudaSuccess) {
    slab_set_error("sla_b200: cudaMallocHost(%zu) failed", bytes);
    ctx->pinned = NULL;
    return NULL;
  }
  ctx->pinned_bytes = bytes + 4096;
  return ctx->pinned;
}

void* slab_host_scratch(SlabCtx* ctx, size_t bytes)
{
  if (ctx->host_scratch_bytes >= bytes) return ctx->host_scratch;
  free(ctx->host_scratch);
  ctx->host_scratch = malloc(bytes + 4096);
  ctx->host_scratch_bytes = ctx->host_scratch ? bytes + 4096 : 0;
  return ctx->host_scratch;
}

extern "C" void slab_last_timing(const SlabCtx* ctx, float ms[SLAB_T_COUNT])
{
  for (int i = 0; i < SLAB_T_COUNT; i++) ms[i] = ctx->last_ms[i];
}

extern "C" uint32_t slab_last_launches(const SlabCtx* ctx) { return ctx->launches; }

extern "C" int slab_copy_to_device(SlabCtx* ctx, void* dst_device, const void* src_host, size_t bytes)
{
  SLAB_CUDA_TRY(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  SLAB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int slab_copy_from_device(SlabCtx* ctx, void* dst_host, const void* src_device, size_t bytes)
{
  SLAB_CUDA_TRY(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SLAB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" void slab_ctx_bind(SlabCtx* ctx) { cudaSetDevice(ctx->device); }

extern "C" void* slab_user_buffer(SlabCtx* ctx, int which, size_t bytes)
{
  if (which < 0 || which > 3) return NULL;
  return slab_arena(ctx, SLAB_NUM_ARENAS - 4 + which, bytes);
}

extern "C" int slab_upload_async(SlabCtx* ctx, void* dst_device, const void* src_host, size_t bytes)
{
  SLAB_CUDA_TRY(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

extern "C" int slab_download_async(SlabCtx* ctx, void* dst_host, const void* src_device, size_t bytes)
{
  SLAB_CUDA_TRY(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

extern "C" int slab_copy_d2d_async(SlabCtx* ctx, void* dst_device, const void* src_device, size_t bytes)
{
  SLAB_CUDA_TRY(cudaMemcpyAsync(dst_device, src_device, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

extern "C" int slab_profile_enabled(const SlabCtx* ctx) { return ctx->profile; }

/* device-time span of a call that runs on several contexts: begin before the workers start, end after
 * they have all synchronised; the span becomes the call's "kernels" time */
extern "C" int slab_span_begin(SlabCtx* ctx)
{
  SLAB_CUDA_TRY(cudaEventRecord(ctx->ev_span[0], ctx->stream));
  return 0;
}
extern "C" int slab_span_end(SlabCtx* ctx, uint32_t launches)
{
  SLAB_CUDA_TRY(cudaEventRecord(ctx->ev_span[1], ctx->stream));
  SLAB_CUDA_TRY(cudaEventSynchronize(ctx->ev_span[1]));
  ctx->last_ms[SLAB_T_H2D] = 0.f; ctx->last_ms[SLAB_T_D2H] = 0.f;
  cudaEventElapsedTime(&ctx->last_ms[SLAB_T_KERNELS], ctx->ev_span[0], ctx->ev_span[1]);
  ctx->launches = launches;
  return 0;
}

extern "C" int slab_stream_sync(SlabCtx* ctx)
{
  SLAB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

/* ---- per-kernel timing ---- */
void slab_prof_reset(SlabCtx* ctx) { ctx->prof_count = 0; }

void slab_prof_begin(SlabCtx* ctx, const char* name)
{
  if (!ctx->profile || ctx->prof_count >= SLAB_MAX_PROF) return;
  const uint32_t i = ctx->prof_count;
  if (!ctx->prof_ev[i][0]) { cudaEventCreate(&ctx->prof_ev[i][0]); cudaEventCreate(&ctx->prof_ev[i][1]); }
  ctx->prof_name[i] = name;
  cudaEventRecord(ctx->prof_ev[i][0], ctx->stream);
}

void slab_prof_end(SlabCtx* ctx)
{
  if (!ctx->profile || ctx->prof_count >= SLAB_MAX_PROF) return;
  cudaEventRecord(ctx->prof_ev[ctx->prof_count][1], ctx->stream);
  ctx->prof_count++;
}

void slab_prof_collect(SlabCtx* ctx)
{
  if (!ctx->profile) return;
  for (uint32_t i = 0; i < ctx->prof_count; i++) {
    ctx->prof_ms[i] = 0.f;
    cudaEventElapsedTime(&ctx->prof_ms[i], ctx->prof_ev[i][0], ctx->prof_ev[i][1]);
  }
}

extern "C" void slab_set_profile(SlabCtx* ctx, int on) { ctx->profile = on; ctx->prof_count = 0; }

extern "C" uint32_t slab_get_profile(const SlabCtx* ctx, const char** names, float* ms, uint32_t max_entries)
{
  uint32_t n = ctx->prof_count < max_entries ? ctx->prof_count : max_entries;
  for (uint32_t i = 0; i < n; i++) { names[i] = ctx->prof_name[i]; ms[i] = ctx->prof_ms[i]; }
  return n;
}
