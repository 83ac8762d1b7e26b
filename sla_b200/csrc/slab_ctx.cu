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

/* kernels already opted in to the full dynamic shared memory, per device */
#include <mutex>
static std::mutex g_optin_mu;
static struct { const void* fn; int device; size_t limit; } g_optin[512];
static int g_optin_count = 0;

int slab_optin_lookup(const void* fn, size_t* limit)
{
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_optin_mu);
  for (int i = 0; i < g_optin_count; i++)
    if (g_optin[i].fn == fn && g_optin[i].device == dev) { *limit = g_optin[i].limit; return 1; }
  return 0;
}

void slab_optin_store(const void* fn, size_t limit)
{
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_optin_mu);
  if (g_optin_count < 512) { g_optin[g_optin_count].fn = fn; g_optin[g_optin_count].device = dev; g_optin[g_optin_count].limit = limit; g_optin_count++; }
}

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
#ifndef SLAB_EMUL
  /* A pipelined call keeps up to eight contexts plus a copy stream busy; with the default of 8 hardware
   * work queues their streams alias and serialise (measured on B200: 34.5 -> 32.4 ms encode, 40 -> 34 ms
   * decode end to end on C2 with 32).  Only effective when this is the first CUDA call of the process;
   * otherwise the host program sets the variable itself. */
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
#endif
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
  ctx->stream_main = ctx->stream;
  {
    /* a second stream at the highest priority; without priorities it is simply another stream */
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) { lo = hi = 0; cudaGetLastError(); }
    if (cudaStreamCreateWithPriority(&ctx->stream_hi, cudaStreamNonBlocking, hi) != cudaSuccess) { ctx->stream_hi = NULL; cudaGetLastError(); }
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  }
  for (int i = 0; i < 4; i++) cudaEventCreate(&ctx->ev[i]);
  for (int i = 0; i < 2; i++) cudaEventCreate(&ctx->ev_span[i]);
  return ctx;
}

extern "C" void slab_ctx_destroy(SlabCtx* ctx)
{
  if (!ctx) return;
  ctx->stream = ctx->stream_main;
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < SLAB_NUM_ARENAS; i++) if (ctx->arena[i]) cudaFree(ctx->arena[i]);
  for (uint32_t i = 0; i < ctx->num_windows; i++) if (ctx->windows[i].owns) cudaFree(ctx->windows[i].dev);
  free(ctx->windows);
  free(ctx->win_lut);
  for (int i = 0; i < 4; i++) if (ctx->fft_tab[i]) cudaFree(ctx->fft_tab[i]);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  free(ctx->host_scratch);
  for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 2; i++) cudaEventDestroy(ctx->ev_span[i]);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  for (int i = 0; i < SLAB_XFER_EVENTS; i++) if (ctx->xfer_ev[i]) cudaEventDestroy(ctx->xfer_ev[i]);
  for (int i = 0; i < SLAB_BOUNCE_SLOTS; i++) {
    if (ctx->bounce[i]) cudaFreeHost(ctx->bounce[i]);
    if (ctx->bounce_ev[i]) cudaEventDestroy(ctx->bounce_ev[i]);
  }
  for (int i = 0; i < 2; i++) {
    if (ctx->dl_bounce[i]) cudaFreeHost(ctx->dl_bounce[i]);
    if (ctx->dl_ev[i]) cudaEventDestroy(ctx->dl_ev[i]);
  }
  for (int i = 0; i < SLAB_MAX_PROF; i++)
    if (ctx->prof_ev[i][0]) { cudaEventDestroy(ctx->prof_ev[i][0]); cudaEventDestroy(ctx->prof_ev[i][1]); }
  if (ctx->stream_hi) cudaStreamDestroy(ctx->stream_hi);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
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
  if (which < 0 || which >= SLAB_USER_BUFFERS) return NULL;
  return slab_arena(ctx, SLAB_NUM_ARENAS - SLAB_USER_BUFFERS + which, bytes);
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

/* ---- ordered transfers for the pipelined whole-file calls ----
 * Uploads issued by different contexts on their own streams share PCIe piece by piece, so every chunk
 * of a file would arrive at about the same time - the end.  One copy stream per call serves the chunks
 * in order instead; a mark (event) after each chunk lets the context that encodes it wait on the device.
 * Caller memory that is not page-locked (what a drop-in caller passes: malloc) moves through a small
 * ring of pinned staging pieces filled by the calling thread: a plain cudaMemcpyAsync from pageable
 * memory runs at a fifth of the PCIe rate on this platform (11 against 55 GB/s). */
extern "C" int slab_host_is_pinned(const void* p)
{
#ifdef SLAB_EMUL
  (void)p;
  return 1;
#else
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  return at.type == cudaMemoryTypeHost;
#endif
}

static int xfer_ready(SlabCtx* ctx)
{
  if (ctx->copy_stream == NULL) SLAB_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  return 0;
}

extern "C" int slab_xfer_prepare(SlabCtx* ctx) { return xfer_ready(ctx); }

extern "C" uint32_t slab_xfer_piece_bytes(void) { return SLAB_BOUNCE_BYTES; }

/* page-locked source: one asynchronous copy on the copy stream */
extern "C" int slab_xfer_upload(SlabCtx* ctx, void* dst_device, const void* src_host, size_t bytes)
{
  if (xfer_ready(ctx) != 0) return -1;
  SLAB_CUDA_TRY(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  return 0;
}

/* pageable source: at most one staging piece, through pinned slot `slot`.  Several host threads feed the
 * copy stream this way, each with its own slots (a single memcpy thread moves 11 GB/s here, PCIe 55). */
extern "C" int slab_xfer_upload_staged(SlabCtx* ctx, uint32_t slot, void* dst_device, const void* src_host, size_t bytes)
{
  if (slot >= SLAB_BOUNCE_SLOTS || bytes > SLAB_BOUNCE_BYTES) return -1;
  if (ctx->bounce[slot] == NULL) {
    SLAB_CUDA_TRY(cudaMallocHost(&ctx->bounce[slot], SLAB_BOUNCE_BYTES));
    SLAB_CUDA_TRY(cudaEventCreateWithFlags(&ctx->bounce_ev[slot], cudaEventDisableTiming));
  }
  if (ctx->bounce_busy[slot]) SLAB_CUDA_TRY(cudaEventSynchronize(ctx->bounce_ev[slot]));
  memcpy(ctx->bounce[slot], src_host, bytes);
  SLAB_CUDA_TRY(cudaMemcpyAsync(dst_device, ctx->bounce[slot], bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  SLAB_CUDA_TRY(cudaEventRecord(ctx->bounce_ev[slot], ctx->copy_stream));
  ctx->bounce_busy[slot] = 1;
  return 0;
}

extern "C" int slab_xfer_mark(SlabCtx* ctx, uint32_t index)
{
  if (index >= SLAB_XFER_EVENTS || xfer_ready(ctx) != 0) return -1;
  if (ctx->xfer_ev[index] == NULL) SLAB_CUDA_TRY(cudaEventCreateWithFlags(&ctx->xfer_ev[index], cudaEventDisableTiming));
  SLAB_CUDA_TRY(cudaEventRecord(ctx->xfer_ev[index], ctx->copy_stream));
  return 0;
}

/* the waiter's stream does not run past this point before mark `index` of the owner's copy stream;
 * the mark must have been issued (host side) before this call */
extern "C" int slab_xfer_wait(SlabCtx* waiter, SlabCtx* owner, uint32_t index)
{
  if (index >= SLAB_XFER_EVENTS || owner->xfer_ev[index] == NULL) return -1;
  SLAB_CUDA_TRY(cudaStreamWaitEvent(waiter->stream, owner->xfer_ev[index], 0));
  if (waiter->stream_hi) SLAB_CUDA_TRY(cudaStreamWaitEvent(waiter->stream_hi, owner->xfer_ev[index], 0));
  return 0;
}

/* work queued on the context's high-priority stream from now on starts after everything queued on its
 * main stream so far */
extern "C" int slab_join_hi(SlabCtx* ctx)
{
  if (ctx->stream_hi == NULL) return 0;
  SLAB_CUDA_TRY(cudaEventRecord(ctx->ev_join, ctx->stream));
  SLAB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream_hi, ctx->ev_join, 0));
  return 0;
}

/* continue the context's launch sequence on stream `to`, ordered after everything launched so far */
int slab_hop(SlabCtx* ctx, cudaStream_t to)
{
  if (to == ctx->stream) return 0;
  SLAB_CUDA_TRY(cudaEventRecord(ctx->ev_join, ctx->stream));
  SLAB_CUDA_TRY(cudaStreamWaitEvent(to, ctx->ev_join, 0));
  ctx->stream = to;
  return 0;
}

extern "C" int slab_xfer_sync(SlabCtx* ctx)
{
  if (ctx->copy_stream) SLAB_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
  return 0;
}

/* device -> caller memory on the context's stream.  Page-locked destination: asynchronous (the caller
 * synchronises the stream).  Pageable destination: two pinned staging pieces, the copy of one piece
 * overlapping the memcpy of the other; the data has landed when the function returns. */
extern "C" int slab_download(SlabCtx* ctx, void* dst_host, const void* src_device, size_t bytes, int dst_pinned)
{
  if (bytes == 0) return 0;
  if (dst_pinned) {
    SLAB_CUDA_TRY(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
  }
  for (int i = 0; i < 2; i++)
    if (ctx->dl_bounce[i] == NULL) {
      SLAB_CUDA_TRY(cudaMallocHost(&ctx->dl_bounce[i], SLAB_BOUNCE_BYTES));
      SLAB_CUDA_TRY(cudaEventCreateWithFlags(&ctx->dl_ev[i], cudaEventDisableTiming));
    }
  unsigned char* dst = (unsigned char*)dst_host;
  const unsigned char* src = (const unsigned char*)src_device;
  const size_t pieces = (bytes + SLAB_BOUNCE_BYTES - 1) / SLAB_BOUNCE_BYTES;
  for (size_t k = 0; k <= pieces; k++) {
    if (k < pieces) {
      const size_t off = k * SLAB_BOUNCE_BYTES, take = bytes - off < SLAB_BOUNCE_BYTES ? bytes - off : SLAB_BOUNCE_BYTES;
      SLAB_CUDA_TRY(cudaMemcpyAsync(ctx->dl_bounce[k & 1], src + off, take, cudaMemcpyDeviceToHost, ctx->stream));
      SLAB_CUDA_TRY(cudaEventRecord(ctx->dl_ev[k & 1], ctx->stream));
    }
    if (k > 0) {
      const size_t off = (k - 1) * SLAB_BOUNCE_BYTES, take = bytes - off < SLAB_BOUNCE_BYTES ? bytes - off : SLAB_BOUNCE_BYTES;
      SLAB_CUDA_TRY(cudaEventSynchronize(ctx->dl_ev[(k - 1) & 1]));
      memcpy(dst + off, ctx->dl_bounce[(k - 1) & 1], take);
    }
  }
  return 0;
}

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
