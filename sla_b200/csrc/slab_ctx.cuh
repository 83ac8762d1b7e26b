/* slab_ctx.cuh - per-handle device context shared by the CUDA translation units. */
#ifndef SLAB_CTX_CUH
#define SLAB_CTX_CUH

#include "slab_cuda.h"
#include "slab_device.h"

#include <stdio.h>

#define SLAB_NUM_ARENAS 56
#define SLAB_USER_BUFFERS 8             /* the last arenas: slab_user_buffer(ctx, 0..7) */
#define SLAB_XFER_EVENTS 65             /* chunk marks on the copy stream of a pipelined call */
#define SLAB_BOUNCE_SLOTS 16
#define SLAB_BOUNCE_BYTES (4u << 20)     /* pinned staging piece for pageable caller memory */
#define SLAB_MAX_OPTIN_SMEM 232448u      /* 227 KB: opt-in shared memory per CTA on sm_100 */
#define SLAB_MAX_PROF 48

struct SlabCtx {
  int device;
  cudaStream_t stream;      /* where launches go: stream_main, or stream_hi for parts of a chunk-mode encode */
  cudaStream_t stream_main;
  cudaStream_t stream_hi;   /* highest priority: the short kernels other chunks of a pipelined call wait for */
  cudaEvent_t  ev_join;
  void*  arena[SLAB_NUM_ARENAS];
  size_t arena_bytes[SLAB_NUM_ARENAS];
  void*  host_scratch;      /* plain host scratch that lives as long as the handle */
  size_t host_scratch_bytes;
  void*  pinned;            /* small pinned scratch for result read-back */
  size_t pinned_bytes;
  cudaEvent_t ev[4];
  cudaEvent_t ev_span[2];   /* wall-clock span of a multi-context call, recorded on this context's stream */
  /* ordered transfers of the pipelined whole-file calls (created on first use) */
  cudaStream_t copy_stream;                 /* uploads of one call complete in issue order here */
  cudaEvent_t  xfer_ev[SLAB_XFER_EVENTS];
  void*        bounce[SLAB_BOUNCE_SLOTS];   /* pinned staging for pageable sources (upload side) */
  cudaEvent_t  bounce_ev[SLAB_BOUNCE_SLOTS];
  int          bounce_busy[SLAB_BOUNCE_SLOTS];
  void*        dl_bounce[2];                /* pinned staging for pageable destinations (download side) */
  cudaEvent_t  dl_ev[2];
  float  last_ms[SLAB_T_COUNT];
  uint32_t launches;
  /* optional per-kernel timing (CUDA events around every launch) */
  int      profile;
  uint32_t prof_count;
  const char* prof_name[SLAB_MAX_PROF];
  cudaEvent_t prof_ev[SLAB_MAX_PROF][2];
  float    prof_ms[SLAB_MAX_PROF];
  /* encoder: host-computed analysis windows, cached per distinct block length */
  struct WindowEntry { uint32_t type, length; double* dev; uint32_t owns; }* windows;   /* owns: dev is the base of an allocation */
  uint32_t num_windows, cap_windows;
  double** win_lut;         /* [16385]: table of window type win_lut_type by length (NULL: not cached) */
  uint32_t win_lut_type;
  /* encoder: trigonometric factors of the reference's FFT (long-term fallback), per transform size */
  uint32_t fft_tab_size;
  double*  fft_tab[4];
};

void slab_set_error(const char* fmt, ...);

#define SLAB_CUDA_TRY(expr)                                                                   \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      slab_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));   \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)

/* grow-only device buffer; contents are not preserved across growth */
void* slab_arena(SlabCtx* ctx, int slot, size_t bytes);
void* slab_pinned(SlabCtx* ctx, size_t bytes);
void* slab_host_scratch(SlabCtx* ctx, size_t bytes);

template <typename T> static inline T* slab_arena_as(SlabCtx* ctx, int slot, size_t count)
{
  return reinterpret_cast<T*>(slab_arena(ctx, slot, count * sizeof(T)));
}

int slab_hop(SlabCtx* ctx, cudaStream_t to);
void slab_prof_reset(SlabCtx* ctx);
void slab_prof_begin(SlabCtx* ctx, const char* name);
void slab_prof_end(SlabCtx* ctx);
void slab_prof_collect(SlabCtx* ctx);      /* after the stream has been synchronised */

/* launch + count + optional timing; wrap template kernels in parentheses */
#define SLAB_RUN(ctx, name, kexpr, grid, block, smem, ...)                         \
  do {                                                                             \
    auto kp_ = kexpr;                                                              \
    slab_prof_begin((ctx), (name));                                                \
    SLAB_LAUNCH(kp_, grid, block, smem, (ctx)->stream, __VA_ARGS__);               \
    {                                                                              \
      cudaError_t le_ = cudaPeekAtLastError();                                     \
      if (le_ != cudaSuccess) {                                                    \
        slab_set_error("%s:%d: launch of %s failed: %s", __FILE__, __LINE__, (name), cudaGetErrorString(le_)); \
        return -1;                                                                 \
      }                                                                            \
    }                                                                              \
    slab_prof_end((ctx));                                                          \
    (ctx)->launches++;                                                             \
  } while (0)

/* the same for call sites that must restore state before returning: sets rc to -1 instead of returning */
#define SLAB_RUN_RC(rc, ctx, name, kexpr, grid, block, smem, ...)                  \
  do {                                                                             \
    auto kp_ = kexpr;                                                              \
    slab_prof_begin((ctx), (name));                                                \
    SLAB_LAUNCH(kp_, grid, block, smem, (ctx)->stream, __VA_ARGS__);               \
    if (cudaPeekAtLastError() != cudaSuccess) (rc) = -1;                           \
    slab_prof_end((ctx));                                                          \
    (ctx)->launches++;                                                             \
  } while (0)

/* Opt a kernel in to as much dynamic shared memory as the device allows next to the kernel's static
 * allocation.  The attribute is per function and process-wide: a per-launch value could be lowered by
 * another host thread (several contexts work on chunks of one file concurrently) between this call
 * and the launch, so it is always set to the same maximum.  The opt-in limit does not affect
 * occupancy; the dynamic size given at launch does. */
/* The attribute calls are made once per kernel and device and remembered (slab_ctx.cu): repeated from
 * every chunk of a pipelined call they are driver round trips on the launch path of eight threads. */
int  slab_optin_lookup(const void* fn, size_t* limit);
void slab_optin_store(const void* fn, size_t limit);
template <typename K> static inline int slab_opt_in_smem(K kernel, size_t bytes)
{
  size_t limit = 0;
  if (!slab_optin_lookup(reinterpret_cast<const void*>(kernel), &limit)) {
    cudaFuncAttributes fa;
    SLAB_CUDA_TRY(cudaFuncGetAttributes(&fa, kernel));
    limit = SLAB_MAX_OPTIN_SMEM > fa.sharedSizeBytes ? SLAB_MAX_OPTIN_SMEM - fa.sharedSizeBytes : 0;
    SLAB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    slab_optin_store(reinterpret_cast<const void*>(kernel), limit);
  }
  if (bytes > limit) { slab_set_error("sla_b200: kernel needs %zu bytes of dynamic shared memory, %zu available", bytes, limit); return -1; }
  return 0;
}

static inline unsigned slab_div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

#endif
