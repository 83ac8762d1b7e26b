/*
 * slab_host.c - host side of libsla_b200.so in plain C: the SLA public API (handles, capacity and
 * argument checks, the 43-byte container header, the block chain walk) on top of the CUDA layer
 * behind slab_device.h.  No sample ever passes through a CPU codec here: every encode/decode call
 * ends in slab_encode()/slab_decode(), and handle creation fails when there is no CUDA device.
 *
 * Mirrors, function by function, src/SLAEncoder.c:56-292,804-932 and src/SLADecoder.c:68-305,660-732
 * of the reference (same status codes in the same situations; see tests/test_api_errors.py).
 */
#define _POSIX_C_SOURCE 200809L      /* clock_gettime */
#include "sla_b200.h"
#include "slab_device.h"

#include <pthread.h>
#include <time.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define FLAG_WAVE_FORMAT   1u
#define FLAG_ENCODE_PARAM  2u
#define MIN_BLOCK_SAMPLES  2048u     /* SLAInternal.h:15 */
#define MIN_BLOCK_HEADER   11u       /* SLAInternal.h:35 */

/* Pipelined whole-file calls (host buffers only): the file is cut into chunks, each chunk runs the
 * full device pipeline on one of several contexts (own stream, own arenas, own host thread), so that
 * the PCIe copies of one chunk overlap the kernels of the others and the latency-bound sequential
 * kernels of different chunks overlap each other.  Results are identical to the single-pass path. */
#define PIPE_MAX_WORKERS   16
#define PIPE_MAX_CHUNKS    64
#define PIPE_ENC_MIN_SAMPLES (4u << 20)    /* per channel; below this a single pass is as fast */
#define PIPE_DEC_MIN_BLOCKS  256u

struct SLAEncoder {
  struct SLAEncoderConfig   config;
  struct SLAWaveFormat      wave_format;
  struct SLAEncodeParameter encode_param;
  uint32_t                  status;
  SlabCtx*                  ctx;
  SlabCtx*                  pipe_ctx[PIPE_MAX_WORKERS];   /* [0] = ctx; the others are created on first use */
  uint32_t                  last_fallbacks[3];
  struct SLAB200BlockRecord* dbg_records;
  uint32_t                  dbg_max_records;
  int32_t* const*           dbg_residual;
};

struct SLADecoder {
  struct SLADecoderConfig   config;
  struct SLAWaveFormat      wave_format;
  struct SLAEncodeParameter encode_param;
  uint32_t                  status;
  SlabCtx*                  ctx;
  SlabCtx*                  pipe_ctx[PIPE_MAX_WORKERS];
  uint32_t*                 chain;        /* host block table: off | smp | n, grown on demand */
  uint32_t                  chain_cap;
  SlabCtx*                  dl_ctx[8];         /* download threads of a pipelined decode into pageable memory */
  float                     batch_kernel_ms;   /* last DecodeBatchPCM: kernel time summed over the groups */
  uint32_t                  batch_launches;
};

/* ---------------------------------------------------------------- small helpers ---- */
static uint16_t host_crc16(const uint8_t* p, size_t n)   /* CRC-16/IBM; container header only */
{
  uint16_t crc = 0;
  while (n--) {
    int b;
    crc ^= *p++;
    for (b = 0; b < 8; b++) crc = (uint16_t)((crc & 1u) ? (crc >> 1) ^ 0xA001u : (crc >> 1));
  }
  return crc;
}

static void put_be(uint8_t** p, uint32_t v, int bytes)
{
  while (bytes--) *(*p)++ = (uint8_t)(v >> (8 * bytes));
}

static uint32_t get_be(const uint8_t** p, int bytes)
{
  uint32_t v = 0;
  while (bytes--) v = (v << 8) | *(*p)++;
  return v;
}

static uint32_t roundup_pow2(uint32_t x)
{
  uint32_t p = 1;
  while (p < x && p < 0x80000000u) p <<= 1;
  return p;
}


/* ---------------------------------------------------------------- pipeline plumbing ---- */
static uint32_t env_u32(const char* name, uint32_t dflt)
{
  const char* v = getenv(name);
  if (v == NULL || *v == 0) return dflt;
  return (uint32_t)strtoul(v, NULL, 10);
}

/* contexts [0, want) of a handle; returns how many exist */
static uint32_t pipe_contexts(SlabCtx** slots, SlabCtx* primary, uint32_t want)
{
  uint32_t w;
  slots[0] = primary;
  if (want > PIPE_MAX_WORKERS) want = PIPE_MAX_WORKERS;
  for (w = 1; w < want; w++) {
    if (slots[w] == NULL) slots[w] = slab_ctx_create();
    if (slots[w] == NULL) return w;
  }
  return want;
}

static uint32_t pipe_default_workers(void)
{
  /* the host simulator keeps its CUDA-thread state in globals: one worker, run inline */
  if (slab_is_hostsim()) return 1;
  return env_u32("SLAB200_PIPE_WORKERS", 4);
}

/* SLAB200_PIPE_TRACE=1: one line per chunk with its milestones (milliseconds since the call began) */
static double pipe_now_ms(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

/* `failed` is written under the pipe's mutex; the workers poll it without the lock */
#define PIPE_FAILED(p) (__atomic_load_n(&(p)->failed, __ATOMIC_ACQUIRE) != 0)

typedef void* (*pipe_fn)(void*);
/* run fn(arg[w]) for w in [0, n): n - 1 threads plus the caller */
static void pipe_run(pipe_fn fn, void** args, uint32_t n)
{
  pthread_t th[PIPE_MAX_WORKERS];
  int started[PIPE_MAX_WORKERS];
  uint32_t w;
  for (w = 1; w < n; w++) started[w] = (pthread_create(&th[w], NULL, fn, args[w]) == 0);
  fn(args[0]);
  for (w = 1; w < n; w++) {
    if (started[w]) pthread_join(th[w], NULL);
    else fn(args[w]);                       /* could not start a thread: do its share here */
  }
}

/* the same with a leader: the caller runs lead(lead_arg) - the ordered uploads of a pipelined call -
 * while n worker threads run fn(arg[w]).  With one worker, or in the host simulator (whose CUDA-thread
 * state lives in globals), everything runs inline: the leader first, then the workers in turn. */
static void pipe_run_led(pipe_fn fn, void** args, uint32_t n, pipe_fn lead, void* lead_arg)
{
  pthread_t th[PIPE_MAX_WORKERS];
  int started[PIPE_MAX_WORKERS];
  uint32_t w;
  if (lead == NULL) { pipe_run(fn, args, n); return; }
  if (n <= 1u || slab_is_hostsim()) {
    lead(lead_arg);
    for (w = 0; w < n; w++) fn(args[w]);
    return;
  }
  for (w = 0; w < n; w++) started[w] = (pthread_create(&th[w], NULL, fn, args[w]) == 0);
  lead(lead_arg);
  for (w = 0; w < n; w++) {
    if (started[w]) pthread_join(th[w], NULL);
    else fn(args[w]);
  }
}

const char* SLAB200_LastError(void) { return slab_last_error(); }

/* ---------------------------------------------------------------- ordered uploads ---- */
/* A pipelined call sends its input up once, in file order, on the primary context's copy stream, and
 * records a mark after the part each chunk needs.  Page-locked sources: one thread issues asynchronous
 * copies.  Pageable sources (what a drop-in caller passes: malloc): the ranges are cut into staging
 * pieces and several threads copy them into pinned slots and issue them - a single memcpy thread moves
 * about 11 GB/s on the GPU box, a fifth of the link.  Marks are recorded in chunk order once every piece
 * of the chunk (and of all earlier chunks) has been issued. */
#define UP_MAX_THREADS 8u
struct UpPiece { void* dst; const void* src; size_t bytes; uint32_t chunk; };
struct UpPlan {
  SlabCtx* ctx;
  struct UpPiece* piece; uint32_t npieces, cap;
  uint32_t nchunks;
  uint32_t* last_piece;                /* per chunk: index one past its last piece */
  int pinned;
  void (*publish)(void* user, uint32_t chunk);   /* mark `chunk` has been recorded */
  void (*fail)(void* user);
  int (*cancelled)(void* user);
  void* user;
  pthread_mutex_t mu;
  uint8_t* done; uint32_t prefix, next_mark, next_piece;
  uint32_t threads;
};
struct UpThread { struct UpPlan* plan; uint32_t index; };

static int up_add(struct UpPlan* u, void* dst, const void* src, size_t bytes, uint32_t chunk, size_t piece_bytes)
{
  const uint8_t* s = (const uint8_t*)src; uint8_t* d = (uint8_t*)dst;
  while (bytes > 0) {
    const size_t take = (piece_bytes == 0 || bytes < piece_bytes) ? bytes : piece_bytes;
    if (u->npieces == u->cap) {
      const uint32_t ncap = u->cap ? u->cap * 2u : 256u;
      struct UpPiece* g = (struct UpPiece*)realloc(u->piece, sizeof(*g) * ncap);
      if (g == NULL) return -1;
      u->piece = g; u->cap = ncap;
    }
    u->piece[u->npieces].dst = d; u->piece[u->npieces].src = s; u->piece[u->npieces].bytes = take; u->piece[u->npieces].chunk = chunk;
    u->npieces++;
    s += take; d += take; bytes -= take;
  }
  return 0;
}

/* after piece j has been issued: advance the issued prefix and record the marks it completes */
static int up_issued(struct UpPlan* u, uint32_t j)
{
  int rc = 0;
  pthread_mutex_lock(&u->mu);
  u->done[j] = 1;
  while (u->prefix < u->npieces && u->done[u->prefix]) u->prefix++;
  while (u->next_mark < u->nchunks && u->prefix >= u->last_piece[u->next_mark]) {
    if (slab_xfer_mark(u->ctx, u->next_mark) != 0) { rc = -1; break; }
    u->publish(u->user, u->next_mark);
    u->next_mark++;
  }
  pthread_mutex_unlock(&u->mu);
  return rc;
}

static void* up_thread(void* arg)
{
  struct UpThread* t = (struct UpThread*)arg;
  struct UpPlan* u = t->plan;
  uint32_t turn = 0;
  slab_ctx_bind(u->ctx);
  for (;;) {
    uint32_t j;
    int bad;
    pthread_mutex_lock(&u->mu);
    j = u->next_piece++;
    pthread_mutex_unlock(&u->mu);
    if (j >= u->npieces || u->cancelled(u->user)) break;
    if (u->pinned) bad = slab_xfer_upload(u->ctx, u->piece[j].dst, u->piece[j].src, u->piece[j].bytes);
    else bad = slab_xfer_upload_staged(u->ctx, 2u * t->index + (turn++ & 1u), u->piece[j].dst, u->piece[j].src, u->piece[j].bytes);
    if (bad || up_issued(u, j) != 0) { u->fail(u->user); break; }
  }
  return NULL;
}

/* runs the plan on the calling thread (plus helpers for pageable sources); chunks without pieces still
 * get their mark */
static void up_run(struct UpPlan* u)
{
  struct UpThread th[UP_MAX_THREADS];
  pthread_t tid[UP_MAX_THREADS];
  int started[UP_MAX_THREADS];
  uint32_t t, nthreads = 1;
  u->done = (uint8_t*)calloc(u->npieces + 1u, 1);
  if (u->done == NULL || slab_xfer_prepare(u->ctx) != 0) { u->fail(u->user); free(u->done); return; }
  pthread_mutex_init(&u->mu, NULL);
  if (!u->pinned && !slab_is_hostsim()) {
    nthreads = env_u32("SLAB200_BOUNCE_THREADS", 6);
    if (nthreads < 1u) nthreads = 1u;
    if (nthreads > UP_MAX_THREADS) nthreads = UP_MAX_THREADS;
  }
  u->threads = nthreads;
  /* marks of leading chunks that need nothing */
  pthread_mutex_lock(&u->mu);
  while (u->next_mark < u->nchunks && u->last_piece[u->next_mark] == 0) {
    if (slab_xfer_mark(u->ctx, u->next_mark) != 0) { u->fail(u->user); break; }
    u->publish(u->user, u->next_mark);
    u->next_mark++;
  }
  pthread_mutex_unlock(&u->mu);
  for (t = 0; t < nthreads; t++) { th[t].plan = u; th[t].index = t; }
  for (t = 1; t < nthreads; t++) started[t] = (pthread_create(&tid[t], NULL, up_thread, &th[t]) == 0);
  up_thread(&th[0]);
  for (t = 1; t < nthreads; t++) if (started[t]) pthread_join(tid[t], NULL);
  pthread_mutex_destroy(&u->mu);
  free(u->done);
}

static void up_free(struct UpPlan* u) { free(u->piece); free(u->last_piece); }

/* ================================================================ encoder ==== */
struct SLAEncoder* SLAEncoder_Create(const struct SLAEncoderConfig* config)
{
  struct SLAEncoder* enc;
  if (config == NULL) return NULL;
  enc = (struct SLAEncoder*)calloc(1, sizeof(*enc));
  if (enc == NULL) return NULL;
  enc->config = *config;
  enc->ctx = slab_ctx_create();
  if (enc->ctx == NULL) {
    fprintf(stderr, "SLAEncoder_Create: %s\n", slab_last_error());
    free(enc);
    return NULL;
  }
  return enc;
}

void SLAEncoder_Destroy(struct SLAEncoder* encoder)
{
  int w;
  if (encoder == NULL) return;
  for (w = 1; w < PIPE_MAX_WORKERS; w++) if (encoder->pipe_ctx[w]) slab_ctx_destroy(encoder->pipe_ctx[w]);
  slab_ctx_destroy(encoder->ctx);
  free(encoder);
}

SLAApiResult SLAEncoder_SetWaveFormat(struct SLAEncoder* encoder, const struct SLAWaveFormat* wave_format)
{
  if (encoder == NULL || wave_format == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (wave_format->num_channels > encoder->config.max_num_channels || wave_format->bit_per_sample > 32)
    return SLA_APIRESULT_EXCEED_HANDLE_CAPACITY;
  encoder->wave_format = *wave_format;
  encoder->status |= FLAG_WAVE_FORMAT;
  return SLA_APIRESULT_OK;
}

SLAApiResult SLAEncoder_SetEncodeParameter(struct SLAEncoder* encoder, const struct SLAEncodeParameter* p)
{
  if (encoder == NULL || p == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (p->parcor_order > encoder->config.max_parcor_order
      || p->longterm_order > encoder->config.max_longterm_order
      || p->lms_order_per_filter > encoder->config.max_lms_order_per_filter
      || p->max_num_block_samples > encoder->config.max_num_block_samples
      || p->max_num_block_samples < MIN_BLOCK_SAMPLES)
    return SLA_APIRESULT_EXCEED_HANDLE_CAPACITY;
  encoder->encode_param = *p;
  encoder->status |= FLAG_ENCODE_PARAM;
  return SLA_APIRESULT_OK;
}

SLAApiResult SLAEncoder_EncodeHeader(const struct SLAHeaderInfo* h, uint8_t* data, uint32_t data_size)
{
  uint8_t* q = data;
  uint16_t crc;
  if (h == NULL || data == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (data_size < SLA_HEADER_SIZE) return SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;
  put_be(&q, 'S', 1); put_be(&q, 'L', 1); put_be(&q, '*', 1); put_be(&q, 1, 1);
  put_be(&q, SLA_HEADER_SIZE - 8, 4);
  put_be(&q, 0, 2);
  put_be(&q, SLA_FORMAT_VERSION, 4);
  put_be(&q, h->wave_format.num_channels, 1);
  put_be(&q, h->num_samples, 4);
  put_be(&q, h->wave_format.sampling_rate, 4);
  put_be(&q, h->wave_format.bit_per_sample, 1);
  put_be(&q, h->wave_format.offset_lshift, 1);
  put_be(&q, h->encode_param.parcor_order, 1);
  put_be(&q, h->encode_param.longterm_order, 1);
  put_be(&q, h->encode_param.lms_order_per_filter, 1);
  put_be(&q, (uint32_t)h->encode_param.ch_process_method, 1);
  put_be(&q, h->num_blocks, 4);
  put_be(&q, h->encode_param.max_num_block_samples, 2);
  put_be(&q, h->max_block_size, 4);
  put_be(&q, h->max_bit_per_second, 4);
  crc = host_crc16(data + 10, SLA_HEADER_SIZE - 10);
  data[8] = (uint8_t)(crc >> 8); data[9] = (uint8_t)crc;
  return SLA_APIRESULT_OK;
}

/* parameter combinations the block pipeline refuses, in the order the reference trips over them */
static SLAApiResult encoder_precheck(const struct SLAEncoder* e)
{
  const struct SLAEncodeParameter* p = &e->encode_param;
  if ((e->status & (FLAG_WAVE_FORMAT | FLAG_ENCODE_PARAM)) != (FLAG_WAVE_FORMAT | FLAG_ENCODE_PARAM))
    return SLA_APIRESULT_PARAMETER_NOT_SET;
  if (p->ch_process_method == SLA_CHPROCESSMETHOD_STEREO_MS && e->wave_format.num_channels != 2)
    return SLA_APIRESULT_INVAILD_CHPROCESSMETHOD;                      /* SLAEncoder.c:331-337 */
  if ((uint32_t)p->window_function_type > (uint32_t)SLA_WINDOWFUNCTIONTYPE_VORBIS)
    return SLA_APIRESULT_INVALID_WINDOWFUNCTION_TYPE;                  /* SLAEncoder.c:316-318 */
  if (e->wave_format.num_channels == 0 || e->wave_format.num_channels > 8
      || e->wave_format.bit_per_sample == 0)
    return SLA_APIRESULT_EXCEED_HANDLE_CAPACITY;
  if ((p->longterm_order & 1u) == 0)                                   /* SLAPredictor.c:808 */
    return SLA_APIRESULT_FAILED_TO_CALCULATE_COEF;
  if (p->lms_order_per_filter < 4 || (p->lms_order_per_filter & (p->lms_order_per_filter - 1)) != 0
      || p->lms_order_per_filter > 32 || p->parcor_order == 0 || p->parcor_order > 64
      || p->longterm_order > 7)
    return SLA_APIRESULT_FAILED_TO_PREDICT;                            /* SLAPredictor.c:1223-1224 */
  return SLA_APIRESULT_OK;
}

static void fill_job(const struct SLAEncoder* e, SlabEncodeJob* job)
{
  memset(job, 0, sizeof(*job));
  job->num_channels = e->wave_format.num_channels;
  job->bits_per_sample = e->wave_format.bit_per_sample;
  job->sampling_rate = e->wave_format.sampling_rate;
  job->parcor_order = e->encode_param.parcor_order;
  job->longterm_order = e->encode_param.longterm_order;
  job->lms_order = e->encode_param.lms_order_per_filter;
  job->ch_process = (uint32_t)e->encode_param.ch_process_method;
  job->window_type = (uint32_t)e->encode_param.window_function_type;
  job->max_block_samples = e->encode_param.max_num_block_samples;
  job->fft_size = roundup_pow2(2u * e->config.max_num_block_samples);   /* SLAEncoder.c:110 */
  job->forced_lshift = -1;
}


/* ---------------------------------------------------------------- pipelined encode ---- */
/* Chunk i covers the segment chain from where chunk i - 1 stopped up to the first segment boundary
 * at or after its nominal end (a multiple of max_num_block_samples), so the chain - including the
 * re-basing done by the leading-silence rule, SLAEncoder.c:393-408 - is exactly the single-pass one.
 *
 * Data movement: the caller's samples go up ONCE, in file order, on the primary context's copy stream
 * into a device image of the whole file (planes, or raw PCM in PCM mode); after the part chunk i needs
 * (its range plus one block: its last segment may run past the nominal end) a mark is recorded.  The
 * context that encodes chunk i waits for mark i on the device, so chunk 0 starts as soon as its own
 * samples have arrived while the rest of the file is still crossing PCIe.  (When every context uploaded
 * its own chunk the copies shared the link piece by piece and all chunks arrived together, late.)
 * A chunk's chain start arrives from the previous chunk as soon as that chunk's chain is computed.
 * Output offsets are handed over in chunk order; every chunk brings its own bytes down. */
struct EncPipe {
  struct SLAEncoder* enc;
  const int32_t* const* input;       /* planar int32 planes (host, or device when dev), or NULL in PCM mode */
  int dev;                           /* input planes and output stream are device memory */
  uint32_t launches;
  const uint8_t* pcm;                /* interleaved little-endian PCM in host memory (PCM mode) */
  uint32_t pcm_bytes;                /* bytes per sample in PCM mode */
  uint32_t N, nchunks, lshift, data_size;
  uint32_t bound[PIPE_MAX_CHUNKS + 1];   /* nominal chunk boundaries: multiples of the block size, bound[nchunks] = N */
  int lshift_known;          /* 0: a single chunk covers the file and works the shift out itself */
  uint8_t* data;
  /* the file's device image (host-input modes) */
  SlabCtx* xfer;             /* primary context: owns the copy stream, the marks and the image */
  int32_t* d_full; size_t full_plane;    /* planar mode */
  uint8_t* d_pcm_full;                   /* PCM mode */
  int src_pinned, dst_pinned;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  uint32_t marks_issued;     /* marks [0, marks_issued) have been recorded on the copy stream */
  uint32_t starts_known;     /* start[i] is valid for i < starts_known */
  uint32_t* start;           /* absolute first sample of each chunk's chain */
  uint32_t out_turn;         /* chunk whose output goes next */
  uint64_t out_off;
  uint32_t num_blocks, max_block_size, max_bps, or_mask;
  int failed;                /* 1 = device failure, 2 = output buffer too small */
  int trace; double t0;
};
struct EncPipeWorker { struct EncPipe* p; SlabCtx* ctx; uint32_t index, stride; };
struct EncPipeCb { struct EncPipe* p; uint32_t chunk, base; };

static void enc_pipe_publish_start(struct EncPipe* p, uint32_t chunk, uint32_t start)
{
  pthread_mutex_lock(&p->mu);
  if (chunk < p->nchunks) p->start[chunk] = start;
  if (p->starts_known < chunk + 1u) p->starts_known = chunk + 1u;
  pthread_cond_broadcast(&p->cv);
  pthread_mutex_unlock(&p->mu);
}

static void enc_pipe_on_consumed(void* user, uint32_t consumed)
{
  struct EncPipeCb* cb = (struct EncPipeCb*)user;
  enc_pipe_publish_start(cb->p, cb->chunk + 1u, cb->base + consumed);
}

static void enc_pipe_fail(struct EncPipe* p, uint32_t chunk, int why)
{
  pthread_mutex_lock(&p->mu);
  if (p->failed == 0) __atomic_store_n(&p->failed, why, __ATOMIC_RELEASE);
  p->starts_known = p->nchunks + 1u;       /* release everybody */
  p->marks_issued = p->nchunks + 1u;
  if (p->out_turn <= chunk) p->out_turn = p->nchunks;
  pthread_cond_broadcast(&p->cv);
  pthread_mutex_unlock(&p->mu);
}

/* last sample (exclusive) chunk i reads: one block beyond its nominal end */
static uint32_t enc_pipe_up_end(const struct EncPipe* p, uint32_t i)
{
  const uint32_t maxblk = p->enc->encode_param.max_num_block_samples;
  const uint32_t nominal_end = p->bound[i + 1u];
  return (p->N - nominal_end > maxblk) ? nominal_end + maxblk : p->N;
}

/* the calling thread: the whole file goes up in order, one mark per chunk */
static void enc_pipe_publish_mark(void* user, uint32_t chunk)
{
  struct EncPipe* p = (struct EncPipe*)user;
  pthread_mutex_lock(&p->mu);
  if (p->marks_issued < chunk + 1u) p->marks_issued = chunk + 1u;
  pthread_cond_broadcast(&p->cv);
  pthread_mutex_unlock(&p->mu);
}
static void enc_pipe_upload_failed(void* user) { enc_pipe_fail((struct EncPipe*)user, 0, 1); }
static int enc_pipe_cancelled(void* user) { return PIPE_FAILED((struct EncPipe*)user); }

static void* enc_pipe_uploader(void* arg)
{
  struct EncPipe* p = (struct EncPipe*)arg;
  const uint32_t nch = p->enc->wave_format.num_channels;
  const size_t piece = p->src_pinned ? 0 : slab_xfer_piece_bytes();
  struct UpPlan u;
  uint32_t i, c, lo = 0;
  int bad = 0;
  if (p->dev) return NULL;
  memset(&u, 0, sizeof(u));
  u.ctx = p->xfer; u.nchunks = p->nchunks; u.pinned = p->src_pinned;
  u.publish = enc_pipe_publish_mark; u.fail = enc_pipe_upload_failed; u.cancelled = enc_pipe_cancelled; u.user = p;
  u.last_piece = (uint32_t*)calloc(p->nchunks + 1u, sizeof(uint32_t));
  if (u.last_piece == NULL) { enc_pipe_fail(p, 0, 1); return NULL; }
  for (i = 0; i < p->nchunks && !bad; i++) {
    const uint32_t hi = enc_pipe_up_end(p, i);
    if (hi > lo) {
      if (p->pcm != NULL) {
        const size_t fb = (size_t)nch * p->pcm_bytes;
        bad = up_add(&u, p->d_pcm_full + (size_t)lo * fb, p->pcm + (size_t)lo * fb, (size_t)(hi - lo) * fb, i, piece);
      } else {
        for (c = 0; c < nch && !bad; c++)
          bad = up_add(&u, p->d_full + p->full_plane * c + lo, p->input[c] + lo, (size_t)(hi - lo) * 4u, i, piece);
      }
      lo = hi;
    }
    u.last_piece[i] = u.npieces;
  }
  if (bad) enc_pipe_fail(p, 0, 1);
  else up_run(&u);
  up_free(&u);
  return NULL;
}

static void* enc_pipe_worker(void* arg)
{
  struct EncPipeWorker* wk = (struct EncPipeWorker*)arg;
  struct EncPipe* p = wk->p;
  const struct SLAEncoder* e = p->enc;
  const uint32_t nch = e->wave_format.num_channels;
  uint32_t turn = 0;
  slab_ctx_bind(wk->ctx);
  for (;;) {
    uint32_t i, base, nominal_end, up_end, len, c, start;
    size_t plane, cap;
    int32_t* d_in = NULL;
    uint8_t* d_out;
    const int32_t* planes[8];
    SlabEncodeJob job;
    struct EncPipeCb cb;

    double t_take, t_start = 0, t_enc = 0;
    /* chunk i always goes to context i mod workers: the chunks differ in size, and a context whose
     * arenas had to grow for a bigger chunk than last time would stall the whole device in cudaMalloc */
    i = wk->index + wk->stride * turn++;
    if (i >= p->nchunks || PIPE_FAILED(p)) break;
    t_take = pipe_now_ms() - p->t0;

    base = p->bound[i];
    nominal_end = p->bound[i + 1u];
    up_end = enc_pipe_up_end(p, i);
    len = up_end - base;
    plane = ((size_t)len + 3u) & ~(size_t)3u;
    cap = 2u * (size_t)nch * len * ((e->wave_format.bit_per_sample + 7u) / 8u) + (size_t)(len / 1024u + 16u) * 1024u + 65536u;
    d_out = (uint8_t*)slab_user_buffer(wk->ctx, 1, cap + 64u);
    if (d_out == NULL) { enc_pipe_fail(p, i, 1); break; }
    if (p->dev) {
      for (c = 0; c < nch; c++) planes[c] = p->input[c] + base;      /* already resident: chunks only share the GPU */
    } else {
      /* my samples are on their way: wait for the mark on the device, not on the host */
      pthread_mutex_lock(&p->mu);
      while (p->marks_issued <= i && !p->failed) pthread_cond_wait(&p->cv, &p->mu);
      pthread_mutex_unlock(&p->mu);
      if (PIPE_FAILED(p)) break;
      if (slab_xfer_wait(wk->ctx, p->xfer, i) != 0) { enc_pipe_fail(p, i, 1); break; }
      if (p->pcm != NULL) {
        /* raw PCM went up as it is (half the bytes of the int32 planes for 16-bit audio) and is
         * de-interleaved on the device */
        const size_t fb = (size_t)nch * p->pcm_bytes;
        d_in = (int32_t*)slab_user_buffer(wk->ctx, 0, plane * nch * sizeof(int32_t));
        if (d_in == NULL || slab_pcm_to_planar(wk->ctx, d_in, plane, p->d_pcm_full + (size_t)base * fb, nch, p->pcm_bytes, len) != 0) {
          enc_pipe_fail(p, i, 1); return NULL;
        }
        for (c = 0; c < nch; c++) planes[c] = d_in + plane * c;
      } else {
        for (c = 0; c < nch; c++) planes[c] = p->d_full + p->full_plane * c + base;
      }
    }
    /* where does this chunk's chain start? */
    pthread_mutex_lock(&p->mu);
    while (p->starts_known <= i && !p->failed) pthread_cond_wait(&p->cv, &p->mu);
    start = p->start[i];
    pthread_mutex_unlock(&p->mu);
    if (PIPE_FAILED(p)) break;
    t_start = pipe_now_ms() - p->t0;

    if (start >= nominal_end) {
      /* the previous chunk's last segment reached the end of this chunk (the tail of the file, when it
       * is shorter than a block): nothing to do but pass the start on */
      enc_pipe_publish_start(p, i + 1u, start);
      pthread_mutex_lock(&p->mu);
      while (p->out_turn != i && !p->failed) pthread_cond_wait(&p->cv, &p->mu);
      p->out_turn = i + 1u;
      pthread_cond_broadcast(&p->cv);
      pthread_mutex_unlock(&p->mu);
      continue;
    }

    fill_job(e, &job);
    job.input = planes; job.input_on_device = 1; job.num_samples = len;
    job.first_sample = start - base;
    job.soft_end = (i + 1u == p->nchunks) ? 0u : nominal_end - base;
    job.forced_lshift = p->lshift_known ? (int32_t)p->lshift : -1;
    job.out = d_out; job.out_on_device = 1; job.out_offset = 0;
    job.out_capacity = cap > 0xFFFFFFFFu ? 0xFFFFFFFFu : (uint32_t)cap;
    cb.p = p; cb.chunk = i; cb.base = base;
    job.on_consumed = enc_pipe_on_consumed; job.user = &cb;
    /* the last chunks are what is left after the last byte has arrived: ahead of everything still in flight */
    job.high_priority = (!p->dev && p->nchunks > 2u && i + 2u >= p->nchunks && env_u32("SLAB200_PIPE_TAIL_PRIORITY", 1) != 0);
    if (p->trace > 1) slab_set_profile(wk->ctx, 1);
    if (slab_encode(wk->ctx, &job) != 0 || job.overflow) { enc_pipe_fail(p, i, 1); break; }
    t_enc = pipe_now_ms() - p->t0;
    if (p->trace > 1) {
      /* SLAB200_PIPE_TRACE=2: the kernels of this chunk as they ran next to the other chunks */
      const char* names[64]; float ms[64]; char line[2048]; int at = 0;
      uint32_t k, cnt = slab_get_profile(wk->ctx, names, ms, 64);
      for (k = 0; k < cnt && at < 1900; k++) at += snprintf(line + at, sizeof(line) - (size_t)at, " %s=%.2f", strrchr(names[k], ' ') ? strrchr(names[k], ' ') + 1 : names[k], ms[k]);
      fprintf(stderr, "enc chunk %2u kernels:%s\n", i, line);
      slab_set_profile(wk->ctx, 0);
    }

    /* my turn to place the bytes */
    pthread_mutex_lock(&p->mu);
    while (p->out_turn != i && !p->failed) pthread_cond_wait(&p->cv, &p->mu);
    if (!p->failed) {
      if (p->out_off + job.total_bytes > p->data_size) {
        __atomic_store_n(&p->failed, 2, __ATOMIC_RELEASE);
        p->starts_known = p->nchunks + 1u; p->marks_issued = p->nchunks + 1u;
      } else {
        uint8_t* dst = p->data + p->out_off;
        p->out_off += job.total_bytes;
        p->num_blocks += job.num_blocks;
        if (job.max_block_size > p->max_block_size) p->max_block_size = job.max_block_size;
        if (job.max_bit_per_second > p->max_bps) p->max_bps = job.max_bit_per_second;
        p->or_mask |= job.input_or_mask;
        p->launches += slab_last_launches(wk->ctx);
        if (!p->lshift_known) p->lshift = job.offset_lshift;
        p->out_turn = i + 1u;
        pthread_cond_broadcast(&p->cv);
        pthread_mutex_unlock(&p->mu);
        if ((p->dev ? slab_copy_d2d_async(wk->ctx, dst, d_out, job.total_bytes)
                    : slab_download(wk->ctx, dst, d_out, job.total_bytes, p->dst_pinned)) != 0
            || slab_stream_sync(wk->ctx) != 0) {
          enc_pipe_fail(p, i, 1);
          break;
        }
        if (p->trace)
          fprintf(stderr, "enc chunk %2u: taken %7.2f  start known %7.2f  encoded %7.2f  out %7.2f ms  (%u blocks)\n",
                  i, t_take, t_start, t_enc, pipe_now_ms() - p->t0, job.num_blocks);
        continue;
      }
    }
    pthread_cond_broadcast(&p->cv);
    pthread_mutex_unlock(&p->mu);
    break;
  }
  return NULL;
}

/* returns 1 when the pipelined path produced the result (rc, job summary filled in), 0 when the
 * caller should take the single-pass path */
static int encode_whole_pipelined(struct SLAEncoder* encoder, const int32_t* const* input, int dev, const uint8_t* pcm,
    uint32_t pcm_bytes, int force, uint32_t num_samples, uint8_t* data, uint32_t data_size, SlabEncodeJob* summary,
    SLAApiResult* rc)
{
  const uint32_t maxblk = encoder->encode_param.max_num_block_samples;
  const uint32_t bits = encoder->wave_format.bit_per_sample;
  const uint32_t nch = encoder->wave_format.num_channels;
  uint32_t workers = pipe_default_workers();
  uint32_t chunk = env_u32("SLAB200_PIPE_CHUNK_SAMPLES", 0), nchunks, w, probe;
  uint32_t bound_keep[PIPE_MAX_CHUNKS + 1];
  int lshift_known = 1;
  struct EncPipe p;
  struct EncPipeWorker wk[PIPE_MAX_WORKERS];
  void* args[PIPE_MAX_WORKERS];
  SlabEncodeJob mask_job;

  if (encoder->dbg_records != NULL || encoder->dbg_residual != NULL) return 0;
  /* input already resident: chunking buys nothing measurable on B200 (27.4 ms against 25.9 ms for the
   * single pass on C2 - the streams' big kernels simply queue behind each other), so it is opt-in */
  if (dev && (env_u32("SLAB200_PIPE_DEVICE", 0) == 0 || slab_profile_enabled(encoder->ctx))) return 0;
  /* Nominal chunk boundaries.  Default: SLAB200_PIPE_CHUNKS equal chunks - a chunk costs about the
   * same few milliseconds of dependent kernels whatever its size, so the GPU is kept busy by having
   * several chunks in flight (one per context), and the smaller the last chunk the less is left to do
   * after the last byte arrived.  SLAB200_PIPE_CHUNK_SAMPLES forces chunks of that length (tests use one
   * block per chunk). */
  {
    uint32_t bound_tmp[PIPE_MAX_CHUNKS + 1];
    uint32_t k;
    if (chunk == 0) {
      const int small = (workers < 2 || num_samples < PIPE_ENC_MIN_SAMPLES);
      if (small && !force) return 0;
      nchunks = small ? 1u : env_u32("SLAB200_PIPE_CHUNKS", 8);
      if (nchunks > PIPE_MAX_CHUNKS) nchunks = PIPE_MAX_CHUNKS;
      if (nchunks < 1u) nchunks = 1u;
      /* no chunk below 1 Mi samples per channel */
      while (nchunks > 1u && num_samples / nchunks < (1u << 20)) nchunks--;
      {
        /* equal chunks, except that the last two are a half and a quarter of the regular size: what is
         * left to do after the last byte has arrived is one small chunk */
        const int taper = nchunks >= 4u && env_u32("SLAB200_PIPE_TAPER", 1) != 0;
        const uint64_t wsum = taper ? 4ull * (nchunks - 2u) + 3ull : (uint64_t)nchunks;
        uint64_t acc = 0;
        bound_tmp[0] = 0;
        for (k = 0; k < nchunks; k++) {
          acc += !taper ? 1u : (k + 2u < nchunks ? 4u : (k + 2u == nchunks ? 2u : 1u));
          bound_tmp[k + 1u] = (uint32_t)(((uint64_t)num_samples * acc / wsum + maxblk - 1u) / maxblk * maxblk);
        }
      }
    } else {
      chunk = ((chunk + maxblk - 1u) / maxblk) * maxblk;
      if (chunk < maxblk) chunk = maxblk;
      nchunks = (num_samples + chunk - 1u) / chunk;
      if (nchunks > PIPE_MAX_CHUNKS) { nchunks = PIPE_MAX_CHUNKS; chunk = ((num_samples / nchunks + maxblk) / maxblk) * maxblk; nchunks = (num_samples + chunk - 1u) / chunk; }
      for (k = 0; k <= nchunks; k++) bound_tmp[k] = (uint32_t)((uint64_t)k * chunk > num_samples ? num_samples : k * chunk);
    }
    /* drop empty chunks, close the list at the end of the file */
    {
      uint32_t m = 0;
      bound_keep[0] = 0;
      for (k = 1; k <= nchunks; k++) {
        uint32_t bnd = bound_tmp[k] > num_samples ? num_samples : bound_tmp[k];
        if (k == nchunks) bnd = num_samples;
        if (bnd > bound_keep[m]) bound_keep[++m] = bnd;
      }
      nchunks = m;
    }
    if (nchunks < 2 && !force) return 0;
    if (nchunks < 1) return 0;
  }

  /* offset_lshift is a whole-file property (SLAEncoder.c:425-455).  When the first stretch of the
   * file already has the lowest bit of its declared width set, the shift is 0 whatever follows and
   * the chunks can start before the rest of the file has been seen; otherwise one pass over the whole
   * file works it out (the planar entry point falls back to the caller's single pass, the PCM entry
   * point runs the file as one chunk). */
  probe = num_samples < (1u << 20) ? num_samples : (1u << 20);
  fill_job(encoder, &mask_job);
  mask_job.num_samples = probe; mask_job.mask_only = 1;
  if (pcm != NULL) {
    const size_t fb = (size_t)nch * pcm_bytes, plane = ((size_t)probe + 3u) & ~(size_t)3u;
    const int32_t* planes[8];
    int32_t* d_in = (int32_t*)slab_user_buffer(encoder->ctx, 0, plane * nch * sizeof(int32_t));
    void* d_pcm = slab_user_buffer(encoder->ctx, 2, (size_t)probe * fb + 64u);
    uint32_t c;
    if (d_in == NULL || d_pcm == NULL || slab_upload_async(encoder->ctx, d_pcm, pcm, (size_t)probe * fb) != 0
        || slab_pcm_to_planar(encoder->ctx, d_in, plane, d_pcm, nch, pcm_bytes, probe) != 0) { *rc = SLA_APIRESULT_NG; return 1; }
    for (c = 0; c < nch; c++) planes[c] = d_in + plane * c;
    mask_job.input = planes; mask_job.input_on_device = 1;
    if (slab_encode(encoder->ctx, &mask_job) != 0) { *rc = SLA_APIRESULT_NG; return 1; }
  } else {
    mask_job.input = input; mask_job.input_on_device = dev;
    if (slab_encode(encoder->ctx, &mask_job) != 0) return 0;
  }
  {
    int certain = (bits >= 32u) ? (mask_job.input_or_mask & 1u) != 0 : ((mask_job.input_or_mask >> (32u - bits)) & 1u) != 0;
    if (bits < 32u && (mask_job.input_or_mask & ((1u << (32u - bits)) - 1u)) != 0) certain = 0;
    if (!certain) {
      if (!force) return 0;
      nchunks = 1; bound_keep[0] = 0; bound_keep[1] = num_samples;   /* the whole file as one chunk */
    }
    lshift_known = certain;
  }
  if (nchunks > 1u && !slab_is_hostsim()) workers = env_u32("SLAB200_PIPE_ENC_WORKERS", workers > 8u ? workers : 8u);
  workers = pipe_contexts(encoder->pipe_ctx, encoder->ctx, workers);
  if (workers > nchunks) workers = nchunks;
  memset(&p, 0, sizeof(p));
  p.lshift = 0; p.lshift_known = lshift_known;
  p.enc = encoder; p.input = input; p.dev = dev; p.pcm = pcm; p.pcm_bytes = pcm_bytes;
  p.N = num_samples; p.nchunks = nchunks;
  memcpy(p.bound, bound_keep, sizeof(uint32_t) * (nchunks + 1u));
  p.data = data; p.data_size = data_size; p.out_off = SLA_HEADER_SIZE;
  p.xfer = encoder->ctx;
  if (!dev) {
    /* device image of the file */
    if (pcm != NULL) {
      p.d_pcm_full = (uint8_t*)slab_user_buffer(encoder->ctx, 4, (size_t)num_samples * nch * pcm_bytes + 64u);
      if (p.d_pcm_full == NULL) { *rc = SLA_APIRESULT_NG; return 1; }
      p.src_pinned = slab_host_is_pinned(pcm);
    } else {
      p.full_plane = ((size_t)num_samples + 3u) & ~(size_t)3u;
      p.d_full = (int32_t*)slab_user_buffer(encoder->ctx, 5, p.full_plane * nch * sizeof(int32_t));
      if (p.d_full == NULL) { *rc = SLA_APIRESULT_NG; return 1; }
      p.src_pinned = slab_host_is_pinned(input[0]);
    }
    p.dst_pinned = slab_host_is_pinned(data);
  }
  p.start = (uint32_t*)calloc(nchunks + 1u, sizeof(uint32_t));
  if (p.start == NULL) return 0;
  p.starts_known = 1;                      /* start[0] = 0 */
  p.trace = (int)env_u32("SLAB200_PIPE_TRACE", 0); p.t0 = pipe_now_ms();
  pthread_mutex_init(&p.mu, NULL);
  pthread_cond_init(&p.cv, NULL);
  for (w = 0; w < workers; w++) { wk[w].p = &p; wk[w].ctx = encoder->pipe_ctx[w]; wk[w].index = w; wk[w].stride = workers; args[w] = &wk[w]; }
  if (dev) slab_span_begin(encoder->ctx);
  pipe_run_led(enc_pipe_worker, args, workers, dev ? NULL : enc_pipe_uploader, &p);
  if (dev) slab_span_end(encoder->ctx, p.launches);
  else slab_xfer_sync(encoder->ctx);
  pthread_cond_destroy(&p.cv);
  pthread_mutex_destroy(&p.mu);
  free(p.start);

  if (p.failed == 1) { *rc = SLA_APIRESULT_NG; return 1; }
  if (p.failed == 2) { *rc = SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE; return 1; }
  /* garbage below the declared width anywhere in the file: let the single pass report it */
  if (!force && bits < 32u && (p.or_mask & ((1u << (32u - bits)) - 1u)) != 0) return 0;
  summary->offset_lshift = p.lshift;
  summary->num_blocks = p.num_blocks;
  summary->total_bytes = (uint32_t)(p.out_off - SLA_HEADER_SIZE);
  summary->max_block_size = p.max_block_size;
  summary->max_bit_per_second = p.max_bps;
  *rc = SLA_APIRESULT_OK;
  return 1;
}

static SLAApiResult encode_whole_common(struct SLAEncoder* encoder, const int32_t* const* input,
    int on_device, uint32_t num_samples, uint8_t* data, uint32_t data_size, uint32_t* output_size)
{
  struct SLAHeaderInfo header;
  uint8_t head[SLA_HEADER_SIZE];
  SlabEncodeJob job;
  SLAApiResult rc;

  if (encoder == NULL || input == NULL || data == NULL || output_size == NULL)
    return SLA_APIRESULT_INVALID_ARGUMENT;
  if (data_size < SLA_HEADER_SIZE) return SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;
  slab_ctx_bind(encoder->ctx);      /* the calling thread's current device may differ from the handle's */
  if ((rc = encoder_precheck(encoder)) != SLA_APIRESULT_OK) return rc;

  fill_job(encoder, &job);
  job.input = input; job.input_on_device = on_device; job.num_samples = num_samples;
  job.out = data; job.out_on_device = on_device; job.out_capacity = data_size;
  job.out_offset = SLA_HEADER_SIZE;
  job.records = (struct SlabBlockRecord*)encoder->dbg_records;
  job.max_records = encoder->dbg_max_records;
  job.residual_out = encoder->dbg_residual;
  if (num_samples > 0) {
    SLAApiResult prc = SLA_APIRESULT_OK;
    if (encode_whole_pipelined(encoder, input, on_device, NULL, 0, 0, num_samples, data, data_size, &job, &prc)) {
      if (prc == SLA_APIRESULT_NG) fprintf(stderr, "SLAEncoder_EncodeWhole: %s\n", slab_last_error());
      if (prc != SLA_APIRESULT_OK) return prc;
    } else {
      if (slab_encode(encoder->ctx, &job) != 0) {
        fprintf(stderr, "SLAEncoder_EncodeWhole: %s\n", slab_last_error());
        return SLA_APIRESULT_NG;
      }
      if (job.overflow) return SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;   /* SLAEncoder.c:848,915 */
      encoder->last_fallbacks[0] = job.fallback_ltfft; encoder->last_fallbacks[1] = job.fallback_exact_autocorr;
      encoder->last_fallbacks[2] = job.fallback_scalar_ltcorr;
    }
  }
  /* the reference leaves the analysed shift in the handle, SLAEncoder.c:835-837 */
  encoder->wave_format.offset_lshift = (uint8_t)job.offset_lshift;

  header.wave_format = encoder->wave_format;
  header.encode_param = encoder->encode_param;
  header.num_samples = num_samples;
  header.num_blocks = job.num_blocks;
  header.max_block_size = job.max_block_size;
  header.max_bit_per_second = job.max_bit_per_second;
  SLAEncoder_EncodeHeader(&header, head, sizeof(head));
  if (on_device) {
    if (slab_copy_to_device(encoder->ctx, data, head, sizeof(head)) != 0) return SLA_APIRESULT_NG;
  } else {
    memcpy(data, head, sizeof(head));
  }
  *output_size = SLA_HEADER_SIZE + job.total_bytes;
  if (encoder->config.verpose_flag != 0 && num_samples > 0) {
    double raw = (double)num_samples * encoder->wave_format.num_channels * encoder->wave_format.bit_per_sample / 8.0;
    printf("progress:100%% (compress ratio:%3.1f %%)\r", 100.0 * (*output_size) / (raw > 0 ? raw : 1));
    fflush(stdout);
  }
  return SLA_APIRESULT_OK;
}

SLAApiResult SLAEncoder_EncodeWhole(struct SLAEncoder* encoder, const int32_t* const* input,
    uint32_t num_samples, uint8_t* data, uint32_t data_size, uint32_t* output_size)
{
  return encode_whole_common(encoder, input, 0, num_samples, data, data_size, output_size);
}

SLAApiResult SLAB200_Encoder_EncodeWholeDevice(struct SLAEncoder* encoder, const int32_t* const* d_input,
    uint32_t num_samples, uint8_t* d_data, uint32_t data_size, uint32_t* output_size)
{
  return encode_whole_common(encoder, d_input, 1, num_samples, d_data, data_size, output_size);
}

/* Whole-file encode from interleaved little-endian PCM in host memory (the data chunk of a WAV file):
 * what the reference CLI does with src/wav.c:208-252 + SLAEncoder_EncodeWhole, minus the int32 staging
 * on the host - the raw bytes are uploaded and de-interleaved on the device. */
SLAApiResult SLAB200_Encoder_EncodePCM(struct SLAEncoder* encoder, const void* pcm, uint32_t num_samples,
    uint8_t* data, uint32_t data_size, uint32_t* output_size)
{
  struct SLAHeaderInfo header;
  SlabEncodeJob job;
  SLAApiResult rc, prc = SLA_APIRESULT_OK;
  uint32_t bits;
  if (encoder == NULL || pcm == NULL || data == NULL || output_size == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (data_size < SLA_HEADER_SIZE) return SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;
  slab_ctx_bind(encoder->ctx);      /* the calling thread's current device may differ from the handle's */
  if ((rc = encoder_precheck(encoder)) != SLA_APIRESULT_OK) return rc;
  bits = encoder->wave_format.bit_per_sample;
  if (bits != 8 && bits != 16 && bits != 24 && bits != 32) return SLA_APIRESULT_INVALID_ARGUMENT;   /* src/wav.c:224-240 */
  fill_job(encoder, &job);
  if (num_samples > 0) {
    if (!encode_whole_pipelined(encoder, NULL, 0, (const uint8_t*)pcm, bits / 8u, 1, num_samples, data, data_size, &job, &prc))
      prc = SLA_APIRESULT_NG;
    if (prc == SLA_APIRESULT_NG) fprintf(stderr, "SLAB200_Encoder_EncodePCM: %s\n", slab_last_error());
    if (prc != SLA_APIRESULT_OK) return prc;
  }
  encoder->wave_format.offset_lshift = (uint8_t)job.offset_lshift;
  header.wave_format = encoder->wave_format;
  header.encode_param = encoder->encode_param;
  header.num_samples = num_samples;
  header.num_blocks = job.num_blocks;
  header.max_block_size = job.max_block_size;
  header.max_bit_per_second = job.max_bit_per_second;
  SLAEncoder_EncodeHeader(&header, data, SLA_HEADER_SIZE);
  *output_size = SLA_HEADER_SIZE + job.total_bytes;
  return SLA_APIRESULT_OK;
}

/* ---------------------------------------------------------------- batch encode ---- */
/* Many files of one wave format and parameter set (the handle's).  A short file cannot fill the GPU and pays
 * the latency of its ~50 dependent launches (the sequential kernels take as long for one block as for ten
 * thousand); here the files are laid back to back in one set of planes - every file on a multiple of 1024
 * samples, zero-filled in between - and a whole group goes through ONE launch sequence: a segment chain,
 * offset_lshift and statistics per file, everything else per segment or block as in a single file
 * (SlabEncodeJob.num_files).  Groups are cut at ENC_BATCH_MAX_FRAMES / ENC_BATCH_MAX_CHSAMPLES so that the arenas
 * stay bounded, and spread over the pipeline contexts (SLAB200_BATCH_ENC_WORKERS, default 6) so that the copies
 * of one group overlap the kernels of another.  Every stream is the one SLAB200_Encoder_EncodePCM produces for
 * that file. */
#define ENC_BATCH_MAX_FRAMES     (48u << 20)
#define ENC_BATCH_MAX_CHSAMPLES  (192u << 20)
#define ENC_BATCH_LONG_FILE      (4u << 20)           /* frames: a file of this size alone in its group is a plain job */
struct EncGroup { uint32_t first, count; };           /* range of order[] */
struct EncBatch {
  struct SLAEncoder* enc;
  struct SLAB200EncodeItem* items;
  const uint32_t* order;                               /* the non-empty, well-formed items */
  const struct EncGroup* groups;
  uint32_t ngroups;
  uint32_t last_item, last_lshift;
  uint32_t reserve_frames;                             /* plane length of the largest group of the call */
  uint32_t long_frames;                                /* a file of at least this many frames alone in its group: plain job */
  int failed;
};
struct EncBatchWorker { struct EncBatch* b; SlabCtx* ctx; uint32_t index, stride; };

static void enc_batch_header(const struct SLAEncoder* e, struct SLAB200EncodeItem* it, uint32_t lshift, uint32_t num_blocks,
                             uint32_t max_block_size, uint32_t max_bps, uint32_t bytes)
{
  struct SLAHeaderInfo header;
  header.wave_format = e->wave_format;
  header.wave_format.offset_lshift = (uint8_t)lshift;
  header.encode_param = e->encode_param;
  header.num_samples = it->num_samples;
  header.num_blocks = num_blocks;
  header.max_block_size = max_block_size;
  header.max_bit_per_second = max_bps;
  SLAEncoder_EncodeHeader(&header, it->data, SLA_HEADER_SIZE);
  it->output_size = SLA_HEADER_SIZE + bytes;
  it->result = SLA_APIRESULT_OK;
}

/* a group of one long file: nothing to merge, the plain single-file job (no file tables, no per-block look-ups) */
static SLAApiResult enc_batch_one(const struct SLAEncoder* e, SlabCtx* ctx, struct SLAB200EncodeItem* it, uint32_t* lshift)
{
  const uint32_t nch = e->wave_format.num_channels, pb = e->wave_format.bit_per_sample / 8u;
  const uint32_t n = it->num_samples;
  const size_t fb = (size_t)nch * pb, plane = ((size_t)n + 3u) & ~(size_t)3u;
  struct SLAHeaderInfo header;
  SlabEncodeJob job;
  const int32_t* planes[8];
  uint32_t c;
  it->output_size = 0;
  if (it->data == NULL || (it->pcm == NULL && n > 0)) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (it->data_size < SLA_HEADER_SIZE) return SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;
  fill_job(e, &job);
  if (n > 0) {
    size_t cap = (size_t)it->data_size - SLA_HEADER_SIZE;
    int32_t* d_in = (int32_t*)slab_user_buffer(ctx, 0, plane * nch * sizeof(int32_t));
    uint8_t* d_out = (uint8_t*)slab_user_buffer(ctx, 1, cap + 64u);
    void* d_pcm = slab_user_buffer(ctx, 2, (size_t)n * fb + 64u);
    if (d_in == NULL || d_out == NULL || d_pcm == NULL) return SLA_APIRESULT_NG;
    if (slab_upload_async(ctx, d_pcm, it->pcm, (size_t)n * fb) != 0
        || slab_pcm_to_planar(ctx, d_in, plane, d_pcm, nch, pb, n) != 0) return SLA_APIRESULT_NG;
    for (c = 0; c < nch; c++) planes[c] = d_in + plane * c;
    job.input = planes; job.input_on_device = 1; job.num_samples = n;
    job.out = d_out; job.out_on_device = 1; job.out_offset = 0;
    job.out_capacity = cap > 0xFFFFFFFFu ? 0xFFFFFFFFu : (uint32_t)cap;
    if (slab_encode(ctx, &job) != 0) return SLA_APIRESULT_NG;
    if (job.overflow) return SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;
    if (slab_download_async(ctx, it->data + SLA_HEADER_SIZE, d_out, job.total_bytes) != 0 || slab_stream_sync(ctx) != 0)
      return SLA_APIRESULT_NG;
  }
  header.wave_format = e->wave_format;
  header.wave_format.offset_lshift = (uint8_t)job.offset_lshift;
  header.encode_param = e->encode_param;
  header.num_samples = n;
  header.num_blocks = job.num_blocks;
  header.max_block_size = job.max_block_size;
  header.max_bit_per_second = job.max_bit_per_second;
  SLAEncoder_EncodeHeader(&header, it->data, SLA_HEADER_SIZE);
  it->output_size = SLA_HEADER_SIZE + job.total_bytes;
  *lshift = job.offset_lshift;
  return SLA_APIRESULT_OK;
}

static int enc_batch_group(struct EncBatch* b, SlabCtx* ctx, const struct EncGroup* grp)
{
  const struct SLAEncoder* e = b->enc;
  const uint32_t nch = e->wave_format.num_channels, pb = e->wave_format.bit_per_sample / 8u, nf = grp->count;
  const size_t fb = (size_t)nch * pb;
  uint32_t* start = (uint32_t*)malloc(sizeof(uint32_t) * 2u * nf);
  uint64_t* pcm_off = (uint64_t*)malloc(sizeof(uint64_t) * nf);
  struct SlabFileResult* res = (struct SlabFileResult*)malloc(sizeof(struct SlabFileResult) * nf);
  uint32_t* len = start ? start + nf : NULL;
  uint64_t plane_len = 0, pcm_total = 0, bound;
  const int32_t* planes[8];
  SlabEncodeJob job;
  int32_t* d_in; uint8_t* d_out; uint8_t* d_pcm;
  uint32_t f, c;
  int rc = -1;
  if (start == NULL || pcm_off == NULL || res == NULL) goto done;
  for (f = 0; f < nf; f++) {
    const uint32_t n = b->items[b->order[grp->first + f]].num_samples;
    start[f] = (uint32_t)plane_len; len[f] = n; pcm_off[f] = pcm_total;
    plane_len += ((uint64_t)n + 1023u) & ~(uint64_t)1023u;
    pcm_total += ((uint64_t)n * fb + 15u) & ~(uint64_t)15u;
  }
  if (plane_len > 0xFFFFF000ull) { slab_set_error_text("sla_b200: batch group beyond 2^32 frames"); goto done; }
  bound = 2ull * nch * plane_len * pb + (plane_len / MIN_BLOCK_SAMPLES + 2ull * nf) * 1024ull + 65536ull;
  if (bound > 0xF0000000ull) bound = 0xF0000000ull;
  {
    /* sized for the largest group of the call: a context allocates once, not once per growing group */
    const uint64_t rl = b->reserve_frames > plane_len ? b->reserve_frames : plane_len;
    uint64_t rb = 2ull * nch * rl * pb + (rl / MIN_BLOCK_SAMPLES + 2ull * nf) * 1024ull + 65536ull;
    if (rb > 0xF0000000ull) rb = 0xF0000000ull;
    if (rb < bound) rb = bound;
    d_in = (int32_t*)slab_user_buffer(ctx, 0, (size_t)rl * nch * sizeof(int32_t));
    d_out = (uint8_t*)slab_user_buffer(ctx, 1, (size_t)rb + 64u);
    d_pcm = (uint8_t*)slab_user_buffer(ctx, 2, (size_t)(rl * fb > pcm_total ? rl * fb : pcm_total) + 16u * nf + 64u);
  }
  if (d_in == NULL || d_out == NULL || d_pcm == NULL) goto done;
  for (f = 0; f < nf; f++)
    if (slab_upload_async(ctx, d_pcm + pcm_off[f], b->items[b->order[grp->first + f]].pcm, (size_t)len[f] * fb) != 0) goto done;
  if (slab_pcm_to_planar_files(ctx, d_in, (size_t)plane_len, (uint32_t)plane_len, d_pcm, nch, pb, nf, start, len, pcm_off) != 0) goto done;
  fill_job(e, &job);
  for (c = 0; c < nch; c++) planes[c] = d_in + (size_t)plane_len * c;
  job.input = planes; job.input_on_device = 1;
  job.num_samples = start[nf - 1u] + len[nf - 1u];
  job.num_files = nf; job.file_start = start; job.file_len = len; job.files = res;
  job.reserve_samples = b->reserve_frames;
  job.out = d_out; job.out_on_device = 1; job.out_offset = 0; job.out_capacity = (uint32_t)bound;
  if (slab_encode(ctx, &job) != 0) goto done;
  if (job.overflow) { slab_set_error_text("sla_b200: batch group outgrew its output bound"); goto done; }
  for (f = 0; f < nf; f++) {
    const uint32_t idx = b->order[grp->first + f];
    struct SLAB200EncodeItem* it = &b->items[idx];
    if ((uint64_t)res[f].num_bytes + SLA_HEADER_SIZE > it->data_size) { it->result = SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE; continue; }
    if (slab_download_async(ctx, it->data + SLA_HEADER_SIZE, d_out + res[f].byte_offset, res[f].num_bytes) != 0) goto done;
    enc_batch_header(e, it, res[f].offset_lshift, res[f].num_blocks, res[f].max_block_size, res[f].max_bit_per_second, res[f].num_bytes);
    if (idx == b->last_item) b->last_lshift = res[f].offset_lshift;
  }
  if (slab_stream_sync(ctx) != 0) goto done;
  rc = 0;
done:
  free(start); free(pcm_off); free(res);
  return rc;
}

static void* enc_batch_worker(void* arg)
{
  struct EncBatchWorker* wk = (struct EncBatchWorker*)arg;
  struct EncBatch* b = wk->b;
  uint32_t g;
  slab_ctx_bind(wk->ctx);
  /* group g always runs on context g mod workers: repeated calls on a similar corpus find their arenas large enough */
  for (g = wk->index; g < b->ngroups && !PIPE_FAILED(b); g += wk->stride) {
    const struct EncGroup* grp = &b->groups[g];
    int rc;
    if (grp->count == 1u && b->items[b->order[grp->first]].num_samples >= b->long_frames) {
      const uint32_t idx = b->order[grp->first];
      uint32_t lshift = 0;
      b->items[idx].result = enc_batch_one(b->enc, wk->ctx, &b->items[idx], &lshift);
      if (idx == b->last_item && b->items[idx].result == SLA_APIRESULT_OK) b->last_lshift = lshift;
      rc = (b->items[idx].result == SLA_APIRESULT_NG) ? -1 : 0;
    } else rc = enc_batch_group(b, wk->ctx, grp);
    if (rc != 0) { __atomic_store_n(&b->failed, 1, __ATOMIC_RELEASE); break; }
  }
  return NULL;
}

SLAApiResult SLAB200_Encoder_EncodeBatchPCM(struct SLAEncoder* encoder, struct SLAB200EncodeItem* items, uint32_t num_items)
{
  struct EncBatch b;
  struct EncBatchWorker wk[PIPE_MAX_WORKERS];
  void* args[PIPE_MAX_WORKERS];
  struct EncGroup* groups;
  uint32_t* order;
  uint32_t workers, w, i, bits, nch, norder = 0, ng = 0;
  uint64_t frames = 0;
  const uint64_t max_frames = env_u32("SLAB200_BATCH_ENC_FRAMES", ENC_BATCH_MAX_FRAMES);     /* tests cut small groups */
  SLAApiResult rc;
  if (encoder == NULL || (items == NULL && num_items > 0)) return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(encoder->ctx);      /* the calling thread's current device may differ from the handle's */
  if ((rc = encoder_precheck(encoder)) != SLA_APIRESULT_OK) return rc;
  bits = encoder->wave_format.bit_per_sample; nch = encoder->wave_format.num_channels;
  if (bits != 8 && bits != 16 && bits != 24 && bits != 32) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (num_items == 0) return SLA_APIRESULT_OK;
  order = (uint32_t*)malloc(sizeof(uint32_t) * num_items);
  groups = (struct EncGroup*)malloc(sizeof(struct EncGroup) * num_items);
  if (order == NULL || groups == NULL) { free(order); free(groups); return SLA_APIRESULT_NG; }
  memset(&b, 0, sizeof(b));
  b.last_item = 0xFFFFFFFFu; b.last_lshift = encoder->wave_format.offset_lshift;
  for (i = 0; i < num_items; i++) {
    struct SLAB200EncodeItem* it = &items[i];
    uint64_t padded;
    it->output_size = 0;
    if (it->data == NULL || (it->pcm == NULL && it->num_samples > 0)) { it->result = SLA_APIRESULT_INVALID_ARGUMENT; continue; }
    if (it->data_size < SLA_HEADER_SIZE) { it->result = SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE; continue; }
    if (it->num_samples == 0) {                        /* a header and no block, as EncodeWhole of nothing */
      enc_batch_header(encoder, it, 0, 0, 0, 0, 0);
      if (i + 1u == num_items) b.last_lshift = 0;
      continue;
    }
    padded = ((uint64_t)it->num_samples + 1023u) & ~(uint64_t)1023u;
    if (ng == 0 || frames + padded > max_frames || (frames + padded) * nch > ENC_BATCH_MAX_CHSAMPLES) {
      groups[ng].first = norder; groups[ng].count = 0; ng++; frames = 0;
    }
    groups[ng - 1u].count++; frames += padded;
    if (frames > b.reserve_frames && frames < 0xFFFFF000ull) b.reserve_frames = (uint32_t)frames;
    order[norder++] = i;
    if (i + 1u == num_items) b.last_item = i;
    it->result = SLA_APIRESULT_NG;                      /* until its group has run */
  }
  if (ng > 0) {
    workers = slab_is_hostsim() ? 1u : env_u32("SLAB200_BATCH_ENC_WORKERS", 6);
    if (workers > ng) workers = ng;
    if (workers < 1) workers = 1;
    workers = pipe_contexts(encoder->pipe_ctx, encoder->ctx, workers);
    b.enc = encoder; b.items = items; b.order = order; b.groups = groups; b.ngroups = ng;
    b.long_frames = env_u32("SLAB200_BATCH_LONG_FRAMES", ENC_BATCH_LONG_FILE);              /* tests lower it */
    for (w = 0; w < workers; w++) {
      wk[w].b = &b; wk[w].ctx = encoder->pipe_ctx[w]; wk[w].index = w; wk[w].stride = workers;
      args[w] = &wk[w];
    }
    pipe_run(enc_batch_worker, args, workers);
    slab_ctx_bind(encoder->ctx);
  }
  free(order); free(groups);
  encoder->wave_format.offset_lshift = (uint8_t)b.last_lshift;       /* as after EncodeWhole of the last file */
  if (b.failed) {
    fprintf(stderr, "SLAB200_Encoder_EncodeBatchPCM: %s\n", slab_last_error());
    return SLA_APIRESULT_NG;
  }
  return SLA_APIRESULT_OK;
}

/* One block with the handle's current offset_lshift and no partition search, SLAEncoder.c:458-801 */
SLAApiResult SLAEncoder_EncodeBlock(struct SLAEncoder* encoder, const int32_t* const* input,
    uint32_t num_samples, uint8_t* data, uint32_t data_size, uint32_t* output_size)
{
  SlabEncodeJob job;
  SLAApiResult rc;
  if (encoder == NULL || input == NULL || data == NULL || output_size == NULL)
    return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(encoder->ctx);      /* the calling thread's current device may differ from the handle's */
  if ((encoder->status & (FLAG_WAVE_FORMAT | FLAG_ENCODE_PARAM)) != (FLAG_WAVE_FORMAT | FLAG_ENCODE_PARAM))
    return SLA_APIRESULT_PARAMETER_NOT_SET;
  if (num_samples > encoder->config.max_num_block_samples) return SLA_APIRESULT_EXCEED_HANDLE_CAPACITY;
  if (data_size <= SLA_BLOCK_HEADER_SIZE) return SLA_APIRESULT_INSUFFICIENT_DATA_SIZE;
  if ((rc = encoder_precheck(encoder)) != SLA_APIRESULT_OK) return rc;
  if (num_samples == 0) return SLA_APIRESULT_INVALID_ARGUMENT;
  fill_job(encoder, &job);
  job.input = input; job.num_samples = num_samples;
  job.out = data; job.out_capacity = data_size; job.out_offset = 0;
  job.forced_lshift = encoder->wave_format.offset_lshift;
  job.single_block = 1;
  if (slab_encode(encoder->ctx, &job) != 0) {
    fprintf(stderr, "SLAEncoder_EncodeBlock: %s\n", slab_last_error());
    return SLA_APIRESULT_NG;
  }
  if (job.overflow) return SLA_APIRESULT_INSUFFICIENT_DATA_SIZE;
  *output_size = job.total_bytes;
  return SLA_APIRESULT_OK;
}

SLAApiResult SLAB200_Encoder_InputOrMask(struct SLAEncoder* encoder, const int32_t* const* input,
    uint32_t num_samples, uint32_t* or_mask)
{
  SlabEncodeJob job;
  SLAApiResult rc;
  if (encoder == NULL || input == NULL || or_mask == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(encoder->ctx);      /* the calling thread's current device may differ from the handle's */
  if ((rc = encoder_precheck(encoder)) != SLA_APIRESULT_OK) return rc;
  fill_job(encoder, &job);
  job.input = input; job.num_samples = num_samples; job.mask_only = 1;
  if (num_samples > 0 && slab_encode(encoder->ctx, &job) != 0) return SLA_APIRESULT_NG;
  *or_mask = job.input_or_mask;
  return SLA_APIRESULT_OK;
}

SLAApiResult SLAB200_Encoder_EncodeRange(struct SLAEncoder* encoder, const int32_t* const* input,
    uint32_t num_samples, uint32_t offset_lshift, uint8_t* data, uint32_t data_size,
    struct SLAB200RangeResult* result)
{
  SlabEncodeJob job;
  SLAApiResult rc;
  if (encoder == NULL || input == NULL || data == NULL || result == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(encoder->ctx);      /* the calling thread's current device may differ from the handle's */
  if ((rc = encoder_precheck(encoder)) != SLA_APIRESULT_OK) return rc;
  fill_job(encoder, &job);
  job.input = input; job.num_samples = num_samples;
  job.out = data; job.out_capacity = data_size; job.out_offset = 0;
  job.forced_lshift = (int32_t)offset_lshift;
  job.records = (struct SlabBlockRecord*)encoder->dbg_records;
  job.max_records = encoder->dbg_max_records;
  job.residual_out = encoder->dbg_residual;
  if (num_samples > 0 && slab_encode(encoder->ctx, &job) != 0) {
    fprintf(stderr, "SLAB200_Encoder_EncodeRange: %s\n", slab_last_error());
    return SLA_APIRESULT_NG;
  }
  if (job.overflow) return SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;
  result->num_blocks = job.num_blocks; result->total_bytes = job.total_bytes;
  result->max_block_size = job.max_block_size; result->max_bit_per_second = job.max_bit_per_second;
  result->input_or_mask = job.input_or_mask;
  return SLA_APIRESULT_OK;
}

SLAApiResult SLAB200_Encoder_InputOrMaskDevice(struct SLAEncoder* encoder, const int32_t* const* d_input,
    uint32_t num_samples, uint32_t* or_mask)
{
  SlabEncodeJob job;
  SLAApiResult rc;
  if (encoder == NULL || d_input == NULL || or_mask == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(encoder->ctx);
  if ((rc = encoder_precheck(encoder)) != SLA_APIRESULT_OK) return rc;
  fill_job(encoder, &job);
  job.input = d_input; job.input_on_device = 1; job.num_samples = num_samples; job.mask_only = 1;
  if (num_samples > 0 && slab_encode(encoder->ctx, &job) != 0) return SLA_APIRESULT_NG;
  *or_mask = job.input_or_mask;
  return SLA_APIRESULT_OK;
}

/* One shard of a multi-GPU encode: the chunk mode of the pipelined whole-file call (chain start in,
 * chain end out through the callback), driven by the caller's ranks instead of this handle's contexts. */
SLAApiResult SLAB200_Encoder_EncodeShard(struct SLAEncoder* encoder, const int32_t* const* input, int input_on_device,
    uint32_t num_samples, uint32_t chain_start, uint32_t soft_end, uint32_t offset_lshift,
    uint8_t* data, int data_on_device, uint32_t data_size,
    SLAB200ChainCallback on_chain, void* user, struct SLAB200ShardResult* result)
{
  SlabEncodeJob job;
  SLAApiResult rc;
  if (encoder == NULL || input == NULL || data == NULL || result == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(encoder->ctx);
  if ((rc = encoder_precheck(encoder)) != SLA_APIRESULT_OK) return rc;
  memset(result, 0, sizeof(*result));
  if (offset_lshift >= encoder->wave_format.bit_per_sample) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (soft_end != 0 && soft_end > num_samples) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (chain_start >= (soft_end != 0 ? soft_end : num_samples)) {
    /* the previous shard's last segment covered this shard's whole range (or the file ended): no blocks */
    result->next_start = chain_start;
    if (on_chain) on_chain(user, chain_start);
    return SLA_APIRESULT_OK;
  }
  fill_job(encoder, &job);
  job.input = input; job.input_on_device = input_on_device; job.num_samples = num_samples;
  job.first_sample = chain_start; job.soft_end = soft_end;
  job.forced_lshift = (int32_t)offset_lshift;
  job.out = data; job.out_on_device = data_on_device; job.out_capacity = data_size; job.out_offset = 0;
  job.on_consumed = on_chain; job.user = user;
  if (slab_encode(encoder->ctx, &job) != 0) {
    fprintf(stderr, "SLAB200_Encoder_EncodeShard: %s\n", slab_last_error());
    return SLA_APIRESULT_NG;
  }
  if (job.overflow) return SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;
  result->num_blocks = job.num_blocks; result->total_bytes = job.total_bytes;
  result->max_block_size = job.max_block_size; result->max_bit_per_second = job.max_bit_per_second;
  result->next_start = job.consumed_samples;
  return SLA_APIRESULT_OK;
}

SLAApiResult SLAB200_Encoder_Download(struct SLAEncoder* encoder, void* dst_host, const void* src_device, uint32_t bytes)
{
  if (encoder == NULL || (bytes > 0 && (dst_host == NULL || src_device == NULL))) return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(encoder->ctx);
  if (slab_download(encoder->ctx, dst_host, src_device, bytes, slab_host_is_pinned(dst_host)) != 0
      || slab_stream_sync(encoder->ctx) != 0) return SLA_APIRESULT_NG;
  return SLA_APIRESULT_OK;
}

/* Test hook: SLALongTermCalculator_CalculateCoef as this encoder handle would run it on a block's PARCOR
 * residual (transform size from the handle's capacity, SLAEncoder.c:110). */
SLAApiResult SLAB200_Debug_LongTerm(struct SLAEncoder* encoder, const int32_t* residual, uint32_t num_samples,
    uint32_t num_taps, uint32_t* pitch_period, double* coef)
{
  if (encoder == NULL || residual == NULL || pitch_period == NULL || coef == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(encoder->ctx);
  if (slab_debug_longterm(encoder->ctx, residual, num_samples, num_taps, roundup_pow2(2u * encoder->config.max_num_block_samples),
                          pitch_period, coef) != 0) return SLA_APIRESULT_NG;
  return SLA_APIRESULT_OK;
}

void SLAB200_Encoder_SetDebugExport(struct SLAEncoder* encoder, struct SLAB200BlockRecord* records,
    uint32_t max_records, int32_t* const* residual_out)
{
  if (encoder == NULL) return;
  encoder->dbg_records = records;
  encoder->dbg_max_records = records ? max_records : 0;
  encoder->dbg_residual = residual_out;
}

void SLAB200_Encoder_LastTiming(const struct SLAEncoder* encoder, float ms[3], uint32_t* launches)
{
  if (encoder == NULL) return;
  slab_last_timing(encoder->ctx, ms);
  if (launches) *launches = slab_last_launches(encoder->ctx);
}

void SLAB200_Encoder_LastFallbacks(const struct SLAEncoder* encoder, uint32_t counts[3])
{
  if (encoder == NULL || counts == NULL) return;
  counts[0] = encoder->last_fallbacks[0]; counts[1] = encoder->last_fallbacks[1]; counts[2] = encoder->last_fallbacks[2];
}

void SLAB200_Encoder_EnableProfile(struct SLAEncoder* encoder, int on) { if (encoder) slab_set_profile(encoder->ctx, on); }
uint32_t SLAB200_Encoder_GetProfile(const struct SLAEncoder* encoder, const char** names, float* ms, uint32_t max_entries)
{
  return encoder ? slab_get_profile(encoder->ctx, names, ms, max_entries) : 0;
}

/* ================================================================ decoder ==== */
SLAApiResult SLADecoder_DecodeHeader(const uint8_t* data, uint32_t data_size, struct SLAHeaderInfo* out)
{
  const uint8_t* q = data;
  struct SLAHeaderInfo h;
  SLAApiResult rc = SLA_APIRESULT_OK;
  if (data == NULL || out == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (data_size < SLA_HEADER_SIZE) return SLA_APIRESULT_INSUFFICIENT_DATA_SIZE;
  if (q[0] != 'S' || q[1] != 'L' || q[2] != '*' || q[3] != 1) return SLA_APIRESULT_INVALID_HEADER_FORMAT;
  q += 8;
  /* a CRC mismatch is reported but the fields are still returned, SLADecoder.c:202-206,251-253 */
  if (get_be(&q, 2) != host_crc16(data + 10, SLA_HEADER_SIZE - 10)) rc = SLA_APIRESULT_DETECT_DATA_CORRUPTION;
  if (get_be(&q, 4) != SLA_FORMAT_VERSION) return SLA_APIRESULT_INVALID_HEADER_FORMAT;
  memset(&h, 0, sizeof(h));
  h.wave_format.num_channels = get_be(&q, 1);
  h.num_samples = get_be(&q, 4);
  h.wave_format.sampling_rate = get_be(&q, 4);
  h.wave_format.bit_per_sample = get_be(&q, 1);
  h.wave_format.offset_lshift = (uint8_t)get_be(&q, 1);
  h.encode_param.parcor_order = get_be(&q, 1);
  h.encode_param.longterm_order = get_be(&q, 1);
  h.encode_param.lms_order_per_filter = get_be(&q, 1);
  h.encode_param.ch_process_method = (SLAChannelProcessMethod)get_be(&q, 1);
  h.num_blocks = get_be(&q, 4);
  h.encode_param.max_num_block_samples = get_be(&q, 2);
  h.max_block_size = get_be(&q, 4);
  h.max_bit_per_second = get_be(&q, 4);
  *out = h;
  return rc;
}

struct SLADecoder* SLADecoder_Create(const struct SLADecoderConfig* config)
{
  struct SLADecoder* dec;
  if (config == NULL) return NULL;
  dec = (struct SLADecoder*)calloc(1, sizeof(*dec));
  if (dec == NULL) return NULL;
  dec->config = *config;
  dec->ctx = slab_ctx_create();
  if (dec->ctx == NULL) {
    fprintf(stderr, "SLADecoder_Create: %s\n", slab_last_error());
    free(dec);
    return NULL;
  }
  return dec;
}

void SLADecoder_Destroy(struct SLADecoder* decoder)
{
  int w;
  if (decoder == NULL) return;
  for (w = 1; w < PIPE_MAX_WORKERS; w++) if (decoder->pipe_ctx[w]) slab_ctx_destroy(decoder->pipe_ctx[w]);
  for (w = 0; w < 8; w++) if (decoder->dl_ctx[w]) slab_ctx_destroy(decoder->dl_ctx[w]);
  slab_ctx_destroy(decoder->ctx);
  free(decoder->chain);
  free(decoder);
}

SLAApiResult SLADecoder_SetWaveFormat(struct SLADecoder* decoder, const struct SLAWaveFormat* wave_format)
{
  if (decoder == NULL || wave_format == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (wave_format->num_channels > decoder->config.max_num_channels || wave_format->bit_per_sample > 32)
    return SLA_APIRESULT_EXCEED_HANDLE_CAPACITY;
  decoder->wave_format = *wave_format;
  decoder->status |= FLAG_WAVE_FORMAT;
  return SLA_APIRESULT_OK;
}

SLAApiResult SLADecoder_SetEncodeParameter(struct SLADecoder* decoder, const struct SLAEncodeParameter* p)
{
  if (decoder == NULL || p == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (p->parcor_order > decoder->config.max_parcor_order
      || p->longterm_order > decoder->config.max_longterm_order
      || p->lms_order_per_filter > decoder->config.max_lms_order_per_filter
      || p->max_num_block_samples > decoder->config.max_num_block_samples
      || p->max_num_block_samples < MIN_BLOCK_SAMPLES)
    return SLA_APIRESULT_EXCEED_HANDLE_CAPACITY;
  decoder->encode_param = *p;
  decoder->status |= FLAG_ENCODE_PARAM;
  return SLA_APIRESULT_OK;
}

static void fill_decode_job(const struct SLADecoder* d, SlabDecodeJob* job)
{
  memset(job, 0, sizeof(*job));
  job->num_channels = d->wave_format.num_channels;
  job->bits_per_sample = d->wave_format.bit_per_sample;
  job->offset_lshift = d->wave_format.offset_lshift;
  job->parcor_order = d->encode_param.parcor_order;
  job->longterm_order = d->encode_param.longterm_order;
  job->lms_order = d->encode_param.lms_order_per_filter;
  job->ch_process = (uint32_t)d->encode_param.ch_process_method;
  job->check_crc = (d->config.enable_crc_check == 1);
  job->max_block_samples = d->encode_param.max_num_block_samples;
}

static SLAApiResult decoder_header_setup(struct SLADecoder* decoder, const struct SLAHeaderInfo* header)
{
  SLAApiResult rc;
  if ((rc = SLADecoder_SetWaveFormat(decoder, &header->wave_format)) != SLA_APIRESULT_OK) return rc;
  if ((rc = SLADecoder_SetEncodeParameter(decoder, &header->encode_param)) != SLA_APIRESULT_OK) return rc;
  if (header->num_samples == 0) return SLA_APIRESULT_OK;
  if (header->encode_param.ch_process_method == SLA_CHPROCESSMETHOD_STEREO_MS
      && header->wave_format.num_channels != 2)
    return SLA_APIRESULT_INVAILD_CHPROCESSMETHOD;                      /* SLADecoder.c:607-615 */
  if (header->wave_format.num_channels == 0 || header->wave_format.num_channels > 8
      || header->wave_format.bit_per_sample == 0
      || header->wave_format.offset_lshift >= header->wave_format.bit_per_sample
      || header->encode_param.parcor_order > 64 || header->encode_param.longterm_order > 7
      || header->encode_param.lms_order_per_filter < 4 || header->encode_param.lms_order_per_filter > 32
      || (header->encode_param.lms_order_per_filter & (header->encode_param.lms_order_per_filter - 1)) != 0)
    return SLA_APIRESULT_INVALID_HEADER_FORMAT;
  return SLA_APIRESULT_OK;
}


/* ---------------------------------------------------------------- downloads into pageable memory ---- */
/* A pageable destination is filled by memcpy out of pinned staging slots, about 11 GB/s per thread on the
 * GPU box.  The decode contexts hand the pieces of a finished chunk to this pool, so that several threads -
 * each with its own stream and slots - copy at once and the chunk comes down at the PCIe rate. */
#define DL_MAX_THREADS 8u
#define DL_QUEUE       8192u
struct DlPiece { void* dst; const void* src; size_t bytes; uint32_t* remaining; };
struct DlPool {
  SlabCtx* ctx[DL_MAX_THREADS];
  pthread_t th[DL_MAX_THREADS];
  int started[DL_MAX_THREADS];
  uint32_t nthreads;
  pthread_mutex_t mu;
  pthread_cond_t cv_work, cv_done;
  struct DlPiece* q;
  uint32_t head, tail;
  int stop, failed;
};
struct DlThread { struct DlPool* pool; uint32_t index; };

static void* dl_thread(void* arg)
{
  struct DlThread* t = (struct DlThread*)arg;
  struct DlPool* p = t->pool;
  slab_ctx_bind(p->ctx[t->index]);
  for (;;) {
    struct DlPiece pc;
    int bad;
    pthread_mutex_lock(&p->mu);
    while (p->head == p->tail && !p->stop) pthread_cond_wait(&p->cv_work, &p->mu);
    if (p->head == p->tail) { pthread_mutex_unlock(&p->mu); break; }
    pc = p->q[p->head % DL_QUEUE]; p->head++;
    pthread_mutex_unlock(&p->mu);
    bad = slab_download(p->ctx[t->index], pc.dst, pc.src, pc.bytes, 0);
    pthread_mutex_lock(&p->mu);
    if (bad) p->failed = 1;
    (*pc.remaining)--;
    pthread_cond_broadcast(&p->cv_done);
    pthread_mutex_unlock(&p->mu);
  }
  return NULL;
}

/* a chunk's plane (or PCM run) in pieces; returns when all of them have landed */
static int dl_pool_copy(struct DlPool* p, void* dst, const void* src, size_t bytes, uint32_t* remaining)
{
  const size_t piece = slab_xfer_piece_bytes();
  uint8_t* d = (uint8_t*)dst; const uint8_t* s = (const uint8_t*)src;
  pthread_mutex_lock(&p->mu);
  while (bytes > 0) {
    const size_t take = bytes < piece ? bytes : piece;
    while (p->tail - p->head >= DL_QUEUE) pthread_cond_wait(&p->cv_done, &p->mu);
    p->q[p->tail % DL_QUEUE].dst = d; p->q[p->tail % DL_QUEUE].src = s; p->q[p->tail % DL_QUEUE].bytes = take;
    p->q[p->tail % DL_QUEUE].remaining = remaining;
    p->tail++; (*remaining)++;
    d += take; s += take; bytes -= take;
  }
  pthread_cond_broadcast(&p->cv_work);
  pthread_mutex_unlock(&p->mu);
  return 0;
}

static int dl_pool_wait(struct DlPool* p, uint32_t* remaining)
{
  int failed;
  pthread_mutex_lock(&p->mu);
  while (*remaining != 0) pthread_cond_wait(&p->cv_done, &p->mu);
  failed = p->failed;
  pthread_mutex_unlock(&p->mu);
  return failed ? -1 : 0;
}

static int dl_pool_start(struct DlPool* p, struct SLADecoder* dec, struct DlThread* args)
{
  uint32_t t, n = env_u32("SLAB200_DOWNLOAD_THREADS", 6);
  memset(p, 0, sizeof(*p));
  if (n < 1u) n = 1u;
  if (n > DL_MAX_THREADS) n = DL_MAX_THREADS;
  p->q = (struct DlPiece*)malloc(sizeof(struct DlPiece) * DL_QUEUE);
  if (p->q == NULL) return -1;
  for (t = 0; t < n; t++) {
    if (dec->dl_ctx[t] == NULL) dec->dl_ctx[t] = slab_ctx_create();
    if (dec->dl_ctx[t] == NULL) break;
    p->ctx[t] = dec->dl_ctx[t];
  }
  if (t == 0) { free(p->q); return -1; }
  p->nthreads = t;
  pthread_mutex_init(&p->mu, NULL);
  pthread_cond_init(&p->cv_work, NULL);
  pthread_cond_init(&p->cv_done, NULL);
  for (t = 0; t < p->nthreads; t++) {
    args[t].pool = p; args[t].index = t;
    p->started[t] = (pthread_create(&p->th[t], NULL, dl_thread, &args[t]) == 0);
  }
  return 0;
}

static void dl_pool_stop(struct DlPool* p)
{
  uint32_t t;
  pthread_mutex_lock(&p->mu);
  p->stop = 1;
  pthread_cond_broadcast(&p->cv_work);
  pthread_mutex_unlock(&p->mu);
  for (t = 0; t < p->nthreads; t++) if (p->started[t]) pthread_join(p->th[t], NULL);
  pthread_cond_destroy(&p->cv_done);
  pthread_cond_destroy(&p->cv_work);
  pthread_mutex_destroy(&p->mu);
  free(p->q);
}

/* ---------------------------------------------------------------- pipelined decode ---- */
/* The host walk has produced the block table, so chunks are simply ranges of blocks with about the
 * same number of samples.  The stream bytes go up once, in order, on the primary context's copy stream
 * (one mark per chunk, as in the encoder); the context that decodes chunk i waits for its mark on the
 * device, runs the kernels into device planes and brings its samples down.  A block costs the same few
 * milliseconds of sequential entropy decoding whether it is decoded alone or next to ten thousand
 * others, so many small chunks in flight put the first samples on the way down early and keep the
 * device->host link busy from then on. */
struct DecPipe {
  struct SLADecoder* dec;
  const uint8_t* data;
  int32_t** buffer;          /* planar int32 host planes, or NULL in PCM mode */
  uint8_t* pcm;              /* interleaved little-endian PCM out (PCM mode) */
  uint32_t pcm_bytes;
  uint32_t nb, nchunks, end_off;
  const uint32_t* off; const uint32_t* smp; const uint32_t* n;
  uint32_t* first_block;     /* nchunks + 1 entries */
  SlabCtx* xfer;             /* primary context: copy stream, marks, device image of the stream */
  uint8_t* d_image;          /* image[k] = data[off[0] + k] */
  int src_pinned, dst_pinned;
  struct DlPool* pool;       /* download threads (pageable destination only) */
  pthread_mutex_t mu;
  pthread_cond_t cv;
  uint32_t marks_issued;
  int failed;
  uint32_t bad_block, bad_code;
  int trace; double t0;
};
struct DecPipeWorker { struct DecPipe* p; SlabCtx* ctx; uint32_t index, stride; };

static void dec_pipe_fail(struct DecPipe* p)
{
  pthread_mutex_lock(&p->mu);
  __atomic_store_n(&p->failed, 1, __ATOMIC_RELEASE);
  p->marks_issued = p->nchunks + 1u;
  pthread_cond_broadcast(&p->cv);
  pthread_mutex_unlock(&p->mu);
}

static void dec_pipe_publish_mark(void* user, uint32_t chunk)
{
  struct DecPipe* p = (struct DecPipe*)user;
  pthread_mutex_lock(&p->mu);
  if (p->marks_issued < chunk + 1u) p->marks_issued = chunk + 1u;
  pthread_cond_broadcast(&p->cv);
  pthread_mutex_unlock(&p->mu);
}
static void dec_pipe_upload_failed(void* user) { dec_pipe_fail((struct DecPipe*)user); }
static int dec_pipe_cancelled(void* user) { return PIPE_FAILED((struct DecPipe*)user); }

static void* dec_pipe_uploader(void* arg)
{
  struct DecPipe* p = (struct DecPipe*)arg;
  const size_t piece = p->src_pinned ? 0 : slab_xfer_piece_bytes();
  struct UpPlan u;
  uint32_t i;
  int bad = 0;
  memset(&u, 0, sizeof(u));
  u.ctx = p->xfer; u.nchunks = p->nchunks; u.pinned = p->src_pinned;
  u.publish = dec_pipe_publish_mark; u.fail = dec_pipe_upload_failed; u.cancelled = dec_pipe_cancelled; u.user = p;
  u.last_piece = (uint32_t*)calloc(p->nchunks + 1u, sizeof(uint32_t));
  if (u.last_piece == NULL) { dec_pipe_fail(p); return NULL; }
  for (i = 0; i < p->nchunks && !bad; i++) {
    const uint32_t b0 = p->first_block[i], b1 = p->first_block[i + 1u];
    const uint32_t lo = (b0 < p->nb) ? p->off[b0] : p->end_off, hi = (b1 < p->nb) ? p->off[b1] : p->end_off;
    if (hi > lo) bad = up_add(&u, p->d_image + (lo - p->off[0]), p->data + lo, hi - lo, i, piece);
    u.last_piece[i] = u.npieces;
  }
  if (bad) dec_pipe_fail(p);
  else up_run(&u);
  up_free(&u);
  return NULL;
}

static void* dec_pipe_worker(void* arg)
{
  struct DecPipeWorker* wk = (struct DecPipeWorker*)arg;
  struct DecPipe* p = wk->p;
  const uint32_t nch = p->dec->wave_format.num_channels;
  uint32_t* tab = NULL;
  uint32_t tab_cap = 0, turn = 0;
  slab_ctx_bind(wk->ctx);
  for (;;) {
    uint32_t i, b0, b1, nbk, k, c, total;
    size_t plane;
    int32_t* outs[8];
    int32_t* d_planes;
    SlabDecodeJob job;
    double t_take, t_dec;
    /* chunk i always on context i mod workers (stable arena sizes from call to call) */
    i = wk->index + wk->stride * turn++;
    if (i >= p->nchunks || PIPE_FAILED(p)) break;
    t_take = pipe_now_ms() - p->t0;
    b0 = p->first_block[i]; b1 = p->first_block[i + 1u]; nbk = b1 - b0;
    if (nbk == 0) continue;
    if (tab_cap < nbk) {
      free(tab);
      tab = (uint32_t*)malloc(sizeof(uint32_t) * 3u * nbk);
      tab_cap = tab ? nbk : 0;
      if (tab == NULL) { dec_pipe_fail(p); break; }
    }
    for (k = 0; k < nbk; k++) {
      tab[k] = p->off[b0 + k] - p->off[b0];
      tab[nbk + k] = p->smp[b0 + k] - p->smp[b0];
      tab[2u * nbk + k] = p->n[b0 + k];
    }
    total = (p->smp[b1 - 1u] - p->smp[b0]) + p->n[b1 - 1u];
    plane = ((size_t)total + 3u) & ~(size_t)3u;
    d_planes = (int32_t*)slab_user_buffer(wk->ctx, 0, plane * nch * sizeof(int32_t));
    if (d_planes == NULL) { dec_pipe_fail(p); break; }
    for (c = 0; c < nch; c++) outs[c] = d_planes + plane * c;
    pthread_mutex_lock(&p->mu);
    while (p->marks_issued <= i && !p->failed) pthread_cond_wait(&p->cv, &p->mu);
    pthread_mutex_unlock(&p->mu);
    if (PIPE_FAILED(p)) break;
    if (slab_xfer_wait(wk->ctx, p->xfer, i) != 0) { dec_pipe_fail(p); break; }
    fill_decode_job(p->dec, &job);
    job.stream = p->d_image + (p->off[b0] - p->off[0]);
    job.stream_size = ((b1 < p->nb) ? p->off[b1] : p->end_off) - p->off[b0];
    job.stream_on_device = 1;
    job.num_blocks = nbk;
    job.blk_byte_off = tab; job.blk_smp_off = tab + nbk; job.blk_nsmp = tab + 2u * nbk;
    job.total_samples = total;
    job.max_samples = total;
    job.out = outs; job.out_on_device = 1;
    if (slab_decode(wk->ctx, &job) != 0) { dec_pipe_fail(p); break; }
    t_dec = pipe_now_ms() - p->t0;
    if (p->pcm != NULL) {
      /* interleave on the device, bring the bytes down */
      const size_t fb = (size_t)nch * p->pcm_bytes;
      void* d_pcm = slab_user_buffer(wk->ctx, 2, (size_t)total * fb + 64u);
      if (d_pcm == NULL || slab_planar_to_pcm(wk->ctx, d_pcm, d_planes, plane, nch, p->pcm_bytes, total) != 0) { dec_pipe_fail(p); break; }
      if (p->pool != NULL) {
        uint32_t remaining = 0;
        if (slab_stream_sync(wk->ctx) != 0
            || dl_pool_copy(p->pool, p->pcm + (size_t)p->smp[b0] * fb, d_pcm, (size_t)total * fb, &remaining) != 0
            || dl_pool_wait(p->pool, &remaining) != 0) { dec_pipe_fail(p); break; }
      } else if (slab_download(wk->ctx, p->pcm + (size_t)p->smp[b0] * fb, d_pcm, (size_t)total * fb, p->dst_pinned) != 0
                 || slab_stream_sync(wk->ctx) != 0) { dec_pipe_fail(p); break; }
    } else if (p->pool != NULL) {
      uint32_t remaining = 0;
      int bad = 0;
      for (c = 0; c < nch && !bad; c++)
        bad = dl_pool_copy(p->pool, p->buffer[c] + p->smp[b0], outs[c], (size_t)total * 4u, &remaining);
      if (dl_pool_wait(p->pool, &remaining) != 0 || bad) { dec_pipe_fail(p); break; }
    } else {
      int bad = 0;
      for (c = 0; c < nch && !bad; c++)
        bad = slab_download(wk->ctx, p->buffer[c] + p->smp[b0], outs[c], (size_t)total * 4u, p->dst_pinned);
      if (bad || slab_stream_sync(wk->ctx) != 0) { dec_pipe_fail(p); break; }
    }
    if (p->trace)
      fprintf(stderr, "dec chunk %2u: taken %7.2f  decoded %7.2f  out %7.2f ms  (%u blocks)\n",
              i, t_take, t_dec, pipe_now_ms() - p->t0, nbk);
    if (job.first_bad_block != 0xFFFFFFFFu) {
      pthread_mutex_lock(&p->mu);
      if (b0 + job.first_bad_block < p->bad_block) { p->bad_block = b0 + job.first_bad_block; p->bad_code = job.first_bad_code; }
      pthread_mutex_unlock(&p->mu);
    }
  }
  free(tab);
  return NULL;
}

/* 1 = handled (rc set), 0 = take the single-pass path */
static int decode_whole_pipelined(struct SLADecoder* decoder, const uint8_t* data, uint32_t end_off,
    int32_t** buffer, uint8_t* pcm, uint32_t pcm_bytes, int force, uint32_t nb, uint32_t total_samples, SLAApiResult* rc)
{
  uint32_t workers = pipe_default_workers();
  uint32_t nchunks = env_u32("SLAB200_PIPE_DEC_CHUNKS", 0), w, b, i;
  struct DecPipe p;
  struct DecPipeWorker wk[PIPE_MAX_WORKERS];
  void* args[PIPE_MAX_WORKERS];
  if (nchunks == 0) {
    if ((workers < 2 || nb < PIPE_DEC_MIN_BLOCKS) && !force) return 0;
    nchunks = (workers < 2 || nb < PIPE_DEC_MIN_BLOCKS) ? 1u : 8u;
  }
  if (nchunks > PIPE_MAX_CHUNKS) nchunks = PIPE_MAX_CHUNKS;
  if (nchunks > nb) nchunks = nb;
  if (nchunks < 2 && !force) return 0;
  if (nchunks < 1) nchunks = 1;
  memset(&p, 0, sizeof(p));
  p.dec = decoder; p.data = data; p.buffer = buffer; p.pcm = pcm; p.pcm_bytes = pcm_bytes;
  p.nb = nb; p.nchunks = nchunks; p.end_off = end_off;
  p.off = decoder->chain; p.smp = decoder->chain + decoder->chain_cap; p.n = decoder->chain + 2u * decoder->chain_cap;
  p.bad_block = 0xFFFFFFFFu;
  p.xfer = decoder->ctx;
  p.d_image = (uint8_t*)slab_user_buffer(decoder->ctx, 4, (size_t)(end_off - p.off[0]) + 64u);
  if (p.d_image == NULL) { *rc = SLA_APIRESULT_NG; return 1; }
  p.src_pinned = slab_host_is_pinned(data);
  p.dst_pinned = slab_host_is_pinned(pcm != NULL ? (const void*)pcm : (const void*)buffer[0]);
  p.first_block = (uint32_t*)malloc(sizeof(uint32_t) * (nchunks + 1u));
  if (p.first_block == NULL) return 0;
  /* cut where the running sample count crosses i / nchunks of the total */
  for (i = 0, b = 0; i < nchunks; i++) {
    const uint64_t target = (uint64_t)total_samples * i / nchunks;
    while (b < nb && p.smp[b] < target) b++;
    p.first_block[i] = b;
  }
  p.first_block[nchunks] = nb;
  if (nchunks > 1u && !slab_is_hostsim()) workers = env_u32("SLAB200_PIPE_DEC_WORKERS", 8u);
  workers = pipe_contexts(decoder->pipe_ctx, decoder->ctx, workers);
  if (workers > nchunks) workers = nchunks;
  p.trace = (int)env_u32("SLAB200_PIPE_TRACE", 0); p.t0 = pipe_now_ms();
  pthread_mutex_init(&p.mu, NULL);
  pthread_cond_init(&p.cv, NULL);
  for (w = 0; w < workers; w++) { wk[w].p = &p; wk[w].ctx = decoder->pipe_ctx[w]; wk[w].index = w; wk[w].stride = workers; args[w] = &wk[w]; }
  {
    struct DlPool pool;
    struct DlThread dl_args[DL_MAX_THREADS];
    const int use_pool = !p.dst_pinned && workers > 1u && !slab_is_hostsim() && dl_pool_start(&pool, decoder, dl_args) == 0;
    if (use_pool) p.pool = &pool;
    pipe_run_led(dec_pipe_worker, args, workers, dec_pipe_uploader, &p);
    if (use_pool) dl_pool_stop(&pool);
  }
  slab_xfer_sync(decoder->ctx);
  pthread_cond_destroy(&p.cv);
  pthread_mutex_destroy(&p.mu);
  free(p.first_block);
  if (p.failed) { *rc = SLA_APIRESULT_NG; return 1; }
  *rc = (p.bad_block != 0xFFFFFFFFu) ? (SLAApiResult)p.bad_code : SLA_APIRESULT_OK;
  return 1;
}

static SLAApiResult decode_whole_common(struct SLADecoder* decoder, const uint8_t* data, uint32_t data_size,
    int32_t** buffer, uint8_t* pcm, uint32_t buffer_num_samples, uint32_t* output_num_samples)
{
  struct SLAHeaderInfo header;
  SlabDecodeJob job;
  SLAApiResult rc, walk_rc = SLA_APIRESULT_OK;
  uint32_t off = SLA_HEADER_SIZE, smp = 0, nb = 0, cap;

  if (decoder == NULL || (buffer == NULL && pcm == NULL) || data == NULL || output_num_samples == NULL)
    return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(decoder->ctx);      /* the calling thread's current device may differ from the handle's */
  if ((rc = SLADecoder_DecodeHeader(data, data_size, &header)) != SLA_APIRESULT_OK) return rc;
  if ((rc = decoder_header_setup(decoder, &header)) != SLA_APIRESULT_OK) return rc;
  if (pcm != NULL && header.wave_format.bit_per_sample != 8 && header.wave_format.bit_per_sample != 16
      && header.wave_format.bit_per_sample != 24 && header.wave_format.bit_per_sample != 32)
    return SLA_APIRESULT_INVALID_HEADER_FORMAT;                      /* no WAV sample format for it, src/wav.c:630-668 */

  /* D0: walk the block chain on the host copy (the decoder has no index; SLADecoder.c:697-719) */
  cap = header.num_samples / 1024u + 64u;
  if (decoder->chain_cap < cap) {
    free(decoder->chain);
    decoder->chain = (uint32_t*)malloc(sizeof(uint32_t) * 3u * cap);
    decoder->chain_cap = decoder->chain ? cap : 0;
    if (decoder->chain == NULL) return SLA_APIRESULT_NG;
  }
  while (smp < header.num_samples) {
    uint32_t avail, bsize, n;
    const uint8_t* b;
    if (off > data_size) { walk_rc = SLA_APIRESULT_INSUFFICIENT_DATA_SIZE; break; }
    avail = data_size - off;
    b = data + off;
    if (avail < MIN_BLOCK_HEADER) { walk_rc = SLA_APIRESULT_INSUFFICIENT_DATA_SIZE; break; }
    if (b[0] != 0xFF || b[1] != 0xFF) { walk_rc = SLA_APIRESULT_FAILED_TO_FIND_SYNC_CODE; break; }
    bsize = (((uint32_t)b[2] << 24) | ((uint32_t)b[3] << 16) | ((uint32_t)b[4] << 8) | b[5]) + 6u;
    n = ((uint32_t)b[8] << 8) | b[9];
    if (bsize > avail || bsize < 10u) {
      walk_rc = SLA_APIRESULT_INSUFFICIENT_DATA_SIZE; break;            /* SLADecoder.c:628 */
    }
    if (n > buffer_num_samples - smp) {
      /* the reference verifies the CRC before it notices the short buffer, SLADecoder.c:346,633 */
      uint16_t stored = (uint16_t)(((uint32_t)b[6] << 8) | b[7]);
      walk_rc = (decoder->config.enable_crc_check == 1 && host_crc16(b + 8, bsize - 8) != stored)
              ? SLA_APIRESULT_DETECT_DATA_CORRUPTION : SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;
      break;
    }
    if (nb == decoder->chain_cap) {
      uint32_t ncap = decoder->chain_cap * 2u, i;
      uint32_t* grown = (uint32_t*)malloc(sizeof(uint32_t) * 3u * ncap);
      if (grown == NULL) return SLA_APIRESULT_NG;
      for (i = 0; i < nb; i++) {
        grown[i] = decoder->chain[i];
        grown[ncap + i] = decoder->chain[decoder->chain_cap + i];
        grown[2u * ncap + i] = decoder->chain[2u * decoder->chain_cap + i];
      }
      free(decoder->chain);
      decoder->chain = grown; decoder->chain_cap = ncap;
    }
    decoder->chain[nb] = off;
    decoder->chain[decoder->chain_cap + nb] = smp;
    decoder->chain[2u * decoder->chain_cap + nb] = n;
    nb++; off += bsize; smp += n;
  }

  fill_decode_job(decoder, &job);
  job.stream = data; job.stream_size = data_size; job.stream_on_device = 0;
  job.num_blocks = nb;
  job.blk_byte_off = decoder->chain;
  job.blk_smp_off = decoder->chain + decoder->chain_cap;
  job.blk_nsmp = decoder->chain + 2u * decoder->chain_cap;
  job.total_samples = smp; job.max_samples = header.num_samples;
  job.out = buffer; job.out_on_device = 0;
  if (nb > 0) {
    SLAApiResult prc = SLA_APIRESULT_OK;
    if (decode_whole_pipelined(decoder, data, off, buffer, pcm, header.wave_format.bit_per_sample / 8u, pcm != NULL, nb, smp, &prc)) {
      if (prc == SLA_APIRESULT_NG) fprintf(stderr, "SLADecoder_DecodeWhole: %s\n", slab_last_error());
      if (prc != SLA_APIRESULT_OK) return prc;
    } else {
      if (slab_decode(decoder->ctx, &job) != 0) {
        fprintf(stderr, "SLADecoder_DecodeWhole: %s\n", slab_last_error());
        return SLA_APIRESULT_NG;
      }
      if (job.first_bad_block != 0xFFFFFFFFu) return (SLAApiResult)job.first_bad_code;
    }
  }
  if (walk_rc != SLA_APIRESULT_OK) return walk_rc;
  *output_num_samples = smp;
  if (decoder->config.verpose_flag != 0) { printf("progress:100%% \r"); fflush(stdout); }
  return SLA_APIRESULT_OK;
}

SLAApiResult SLADecoder_DecodeWhole(struct SLADecoder* decoder, const uint8_t* data, uint32_t data_size,
    int32_t** buffer, uint32_t buffer_num_samples, uint32_t* output_num_samples)
{
  if (buffer == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  return decode_whole_common(decoder, data, data_size, buffer, NULL, buffer_num_samples, output_num_samples);
}

/* Whole-file decode to interleaved little-endian PCM in host memory (the data chunk of a WAV file): the
 * reference CLI's SLADecoder_DecodeWhole + src/wav.c:630-668, with the interleave done on the device and
 * only the PCM bytes crossing PCIe.  `pcm` has room for buffer_num_samples frames. */
SLAApiResult SLAB200_Decoder_DecodePCM(struct SLADecoder* decoder, const uint8_t* data, uint32_t data_size,
    void* pcm, uint32_t buffer_num_samples, uint32_t* output_num_samples)
{
  if (pcm == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  return decode_whole_common(decoder, data, data_size, NULL, (uint8_t*)pcm, buffer_num_samples, output_num_samples);
}

/* ---------------------------------------------------------------- batch decode ---- */
/* Files with the same stream parameters are laid end to end in one device image (64-byte aligned), their
 * block tables merged, and decoded by one launch of each kernel; groups are cut when they reach
 * BATCH_MAX_BYTES / BATCH_MAX_FRAMES so that the arenas stay bounded, and the groups are spread over the
 * pipeline contexts. */
#define BATCH_MAX_BYTES   (512u << 20)
#define BATCH_MAX_FRAMES  (96u << 20)

struct BatchFile {
  uint32_t item;                 /* index into items[] */
  uint32_t first_block, num_blocks, frames, end_off;
  SLAApiResult walk_rc;
  struct SLAHeaderInfo header;
};
struct BatchGroup { uint32_t first_file, num_files; };
struct BatchPipe {
  struct SLADecoder* dec;
  struct SLAB200BatchItem* items;
  struct BatchFile* files;       /* sorted so that a group's files are consecutive */
  struct BatchGroup* groups;
  uint32_t ngroups;
  const uint32_t* blk_off; const uint32_t* blk_smp; const uint32_t* blk_n;   /* per file, relative to the file */
  pthread_mutex_t mu;
  uint32_t next_group;
  int failed;
  double kernel_ms; uint32_t launches;
};
struct BatchWorker { struct BatchPipe* p; SlabCtx* ctx; uint32_t index, stride; };

static int batch_same_params(const struct SLAHeaderInfo* a, const struct SLAHeaderInfo* b)
{
  return a->wave_format.num_channels == b->wave_format.num_channels
      && a->wave_format.bit_per_sample == b->wave_format.bit_per_sample
      && a->wave_format.offset_lshift == b->wave_format.offset_lshift
      && a->encode_param.parcor_order == b->encode_param.parcor_order
      && a->encode_param.longterm_order == b->encode_param.longterm_order
      && a->encode_param.lms_order_per_filter == b->encode_param.lms_order_per_filter
      && a->encode_param.ch_process_method == b->encode_param.ch_process_method;
}

static int batch_file_cmp(const void* pa, const void* pb)
{
  const struct BatchFile* a = (const struct BatchFile*)pa;
  const struct BatchFile* b = (const struct BatchFile*)pb;
  const struct SLAHeaderInfo* x = &a->header; const struct SLAHeaderInfo* y = &b->header;
#define CMP(f) if (x->f != y->f) return x->f < y->f ? -1 : 1
  CMP(wave_format.num_channels); CMP(wave_format.bit_per_sample); CMP(wave_format.offset_lshift);
  CMP(encode_param.parcor_order); CMP(encode_param.longterm_order); CMP(encode_param.lms_order_per_filter);
  CMP(encode_param.ch_process_method);
#undef CMP
  return a->item < b->item ? -1 : (a->item > b->item);
}

static void* batch_worker(void* arg)
{
  struct BatchWorker* wk = (struct BatchWorker*)arg;
  struct BatchPipe* p = wk->p;
  uint32_t* tab = NULL; uint32_t tab_cap = 0, turn = 0;
  slab_ctx_bind(wk->ctx);
  for (;;) {
    uint32_t g, f, k, c, nblocks = 0, frames = 0, nch, bytes, maxblk = 0;
    size_t img_bytes = 0, plane, fb;
    const struct BatchGroup* grp;
    const struct BatchFile* f0;
    uint8_t* d_img; int32_t* d_planes; uint8_t* d_pcm;
    int32_t* outs[8];
    SlabDecodeJob job;
    struct SLADecoder local;

    /* group g always runs on context g mod workers: repeated calls on the same corpus then find their
     * arenas already large enough (a growing arena stalls the whole device in cudaMalloc) */
    g = wk->index + wk->stride * turn++;
    if (g >= p->ngroups || PIPE_FAILED(p)) break;
    grp = &p->groups[g];
    f0 = &p->files[grp->first_file];
    nch = f0->header.wave_format.num_channels; bytes = f0->header.wave_format.bit_per_sample / 8u; fb = (size_t)nch * bytes;
    for (f = 0; f < grp->num_files; f++) {
      const struct BatchFile* bf = &f0[f];
      nblocks += bf->num_blocks; frames += bf->frames;
      img_bytes += ((size_t)bf->end_off + 63u) & ~(size_t)63u;
      if (bf->header.encode_param.max_num_block_samples > maxblk) maxblk = bf->header.encode_param.max_num_block_samples;
    }
    if (nblocks == 0) continue;
    if (tab_cap < nblocks) {
      free(tab);
      tab = (uint32_t*)malloc(sizeof(uint32_t) * 4u * nblocks);
      tab_cap = tab ? nblocks : 0;
      if (tab == NULL) goto fail;
    }
    plane = ((size_t)frames + 3u) & ~(size_t)3u;
    d_img = (uint8_t*)slab_user_buffer(wk->ctx, 3, img_bytes + 64u);
    d_planes = (int32_t*)slab_user_buffer(wk->ctx, 0, plane * nch * sizeof(int32_t));
    d_pcm = (uint8_t*)slab_user_buffer(wk->ctx, 2, (size_t)frames * fb + 64u);
    if (d_img == NULL || d_planes == NULL || d_pcm == NULL) goto fail;
    {
      size_t img_off = 0; uint32_t smp_base = 0, b = 0;
      for (f = 0; f < grp->num_files; f++) {
        const struct BatchFile* bf = &f0[f];
        if (bf->num_blocks == 0) continue;
        if (slab_upload_async(wk->ctx, d_img + img_off, p->items[bf->item].data, bf->end_off) != 0) goto fail;
        for (k = 0; k < bf->num_blocks; k++, b++) {
          tab[b] = (uint32_t)img_off + p->blk_off[bf->first_block + k];
          tab[nblocks + b] = smp_base + p->blk_smp[bf->first_block + k];
          tab[2u * nblocks + b] = p->blk_n[bf->first_block + k];
        }
        img_off += ((size_t)bf->end_off + 63u) & ~(size_t)63u;
        smp_base += bf->frames;
      }
    }
    local = *p->dec;                                   /* parameters of this group, without touching the handle */
    local.wave_format = f0->header.wave_format;
    local.encode_param = f0->header.encode_param;
    local.encode_param.max_num_block_samples = maxblk;
    fill_decode_job(&local, &job);
    job.stream = d_img; job.stream_size = (uint32_t)img_bytes; job.stream_on_device = 1;
    job.num_blocks = nblocks;
    job.blk_byte_off = tab; job.blk_smp_off = tab + nblocks; job.blk_nsmp = tab + 2u * nblocks;
    job.total_samples = frames; job.max_samples = frames;
    for (c = 0; c < nch; c++) outs[c] = d_planes + plane * c;
    job.out = outs; job.out_on_device = 1;
    job.blk_err_out = tab + 3u * nblocks;
    if (slab_decode(wk->ctx, &job) != 0) goto fail;
    {
      float ms[SLAB_T_COUNT];
      slab_last_timing(wk->ctx, ms);
      pthread_mutex_lock(&p->mu);
      p->kernel_ms += ms[SLAB_T_KERNELS]; p->launches += slab_last_launches(wk->ctx) + 1u;
      pthread_mutex_unlock(&p->mu);
    }
    if (slab_planar_to_pcm(wk->ctx, d_pcm, d_planes, plane, nch, bytes, frames) != 0) goto fail;
    {
      uint32_t smp_base = 0, b = 0;
      for (f = 0; f < grp->num_files; f++) {
        const struct BatchFile* bf = &f0[f];
        struct SLAB200BatchItem* it = &p->items[bf->item];
        SLAApiResult res = SLA_APIRESULT_OK;
        if (bf->num_blocks == 0) continue;
        for (k = 0; k < bf->num_blocks; k++)
          if (tab[3u * nblocks + b + k] != 0) { res = (SLAApiResult)tab[3u * nblocks + b + k]; break; }
        b += bf->num_blocks;
        if (res == SLA_APIRESULT_OK) res = bf->walk_rc;
        it->result = res;
        if (res == SLA_APIRESULT_OK) {
          it->output_num_samples = bf->frames;
          if (slab_download_async(wk->ctx, it->pcm, d_pcm + (size_t)smp_base * fb, (size_t)bf->frames * fb) != 0) goto fail;
        }
        smp_base += bf->frames;
      }
    }
    if (slab_stream_sync(wk->ctx) != 0) goto fail;
    continue;
fail:
    pthread_mutex_lock(&p->mu); p->failed = 1; pthread_mutex_unlock(&p->mu);
    break;
  }
  free(tab);
  return NULL;
}

SLAApiResult SLAB200_Decoder_DecodeBatchPCM(struct SLADecoder* decoder, struct SLAB200BatchItem* items, uint32_t num_items)
{
  struct BatchFile* files;
  struct BatchGroup* groups;
  struct BatchPipe p;
  struct BatchWorker wk[PIPE_MAX_WORKERS];
  void* args[PIPE_MAX_WORKERS];
  uint32_t* blk = NULL;          /* off | smp | n, blk_cap entries each */
  uint32_t blk_cap = 0, nblk = 0, nfiles = 0, i, w, workers;
  if (decoder == NULL || (items == NULL && num_items > 0)) return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(decoder->ctx);      /* the calling thread's current device may differ from the handle's */
  if (num_items == 0) return SLA_APIRESULT_OK;
  files = (struct BatchFile*)calloc(num_items, sizeof(*files));
  groups = (struct BatchGroup*)calloc(num_items, sizeof(*groups));
  if (files == NULL || groups == NULL) { free(files); free(groups); return SLA_APIRESULT_NG; }

  /* headers and block chains on the host (SLADecoder.c:684-719), one file after the other */
  for (i = 0; i < num_items; i++) {
    struct SLAB200BatchItem* it = &items[i];
    struct BatchFile* bf = &files[nfiles];
    struct SLADecoder probe = *decoder;
    uint32_t off = SLA_HEADER_SIZE, smp = 0, bits;
    SLAApiResult rc;
    it->output_num_samples = 0;
    if (it->data == NULL || it->pcm == NULL) { it->result = SLA_APIRESULT_INVALID_ARGUMENT; continue; }
    if ((rc = SLADecoder_DecodeHeader(it->data, it->data_size, &bf->header)) != SLA_APIRESULT_OK) { it->result = rc; continue; }
    if ((rc = decoder_header_setup(&probe, &bf->header)) != SLA_APIRESULT_OK) { it->result = rc; continue; }
    bits = bf->header.wave_format.bit_per_sample;
    if (bits != 8 && bits != 16 && bits != 24 && bits != 32) { it->result = SLA_APIRESULT_INVALID_HEADER_FORMAT; continue; }
    bf->item = i; bf->first_block = nblk; bf->walk_rc = SLA_APIRESULT_OK;
    while (smp < bf->header.num_samples) {
      uint32_t avail, bsize, n;
      const uint8_t* b;
      if (off > it->data_size) { bf->walk_rc = SLA_APIRESULT_INSUFFICIENT_DATA_SIZE; break; }
      avail = it->data_size - off; b = it->data + off;
      if (avail < MIN_BLOCK_HEADER) { bf->walk_rc = SLA_APIRESULT_INSUFFICIENT_DATA_SIZE; break; }
      if (b[0] != 0xFF || b[1] != 0xFF) { bf->walk_rc = SLA_APIRESULT_FAILED_TO_FIND_SYNC_CODE; break; }
      bsize = (((uint32_t)b[2] << 24) | ((uint32_t)b[3] << 16) | ((uint32_t)b[4] << 8) | b[5]) + 6u;
      n = ((uint32_t)b[8] << 8) | b[9];
      if (bsize > avail || bsize < 10u) { bf->walk_rc = SLA_APIRESULT_INSUFFICIENT_DATA_SIZE; break; }
      if (n > it->capacity_samples - smp) {
        uint16_t stored = (uint16_t)(((uint32_t)b[6] << 8) | b[7]);
        bf->walk_rc = (decoder->config.enable_crc_check == 1 && host_crc16(b + 8, bsize - 8) != stored)
                    ? SLA_APIRESULT_DETECT_DATA_CORRUPTION : SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;
        break;
      }
      if (nblk == blk_cap) {
        const uint32_t ncap = blk_cap ? blk_cap * 2u : 4096u;
        uint32_t* grown = (uint32_t*)malloc(sizeof(uint32_t) * 3u * ncap);
        if (grown == NULL) { free(blk); free(files); free(groups); return SLA_APIRESULT_NG; }
        if (blk) {
          memcpy(grown, blk, sizeof(uint32_t) * nblk);
          memcpy(grown + ncap, blk + blk_cap, sizeof(uint32_t) * nblk);
          memcpy(grown + 2u * ncap, blk + 2u * blk_cap, sizeof(uint32_t) * nblk);
          free(blk);
        }
        blk = grown; blk_cap = ncap;
      }
      blk[nblk] = off; blk[blk_cap + nblk] = smp; blk[2u * blk_cap + nblk] = n;
      nblk++; off += bsize; smp += n;
    }
    bf->num_blocks = nblk - bf->first_block; bf->frames = smp; bf->end_off = off;
    if (bf->num_blocks == 0) { it->result = bf->walk_rc; continue; }      /* nothing decodable (or an empty file) */
    nfiles++;
  }

  memset(&p, 0, sizeof(p));
  if (nfiles > 0) {
    /* files of one parameter set side by side, then cut into bounded groups */
    uint32_t ng = 0;
    size_t bytes = 0, frames = 0;
    qsort(files, nfiles, sizeof(*files), batch_file_cmp);
    for (i = 0; i < nfiles; i++) {
      const size_t fbytes = ((size_t)files[i].end_off + 63u) & ~(size_t)63u;
      const int fresh = (i == 0) || !batch_same_params(&files[i].header, &files[i - 1u].header)
                     || bytes + fbytes > BATCH_MAX_BYTES || frames + files[i].frames > BATCH_MAX_FRAMES;
      if (fresh) { groups[ng].first_file = i; groups[ng].num_files = 0; ng++; bytes = 0; frames = 0; }
      groups[ng - 1u].num_files++;
      bytes += fbytes; frames += files[i].frames;
    }
    p.dec = decoder; p.items = items; p.files = files; p.groups = groups; p.ngroups = ng;
    p.blk_off = blk; p.blk_smp = blk + blk_cap; p.blk_n = blk + 2u * blk_cap;
    workers = pipe_contexts(decoder->pipe_ctx, decoder->ctx, pipe_default_workers());
    if (workers > ng) workers = ng;
    pthread_mutex_init(&p.mu, NULL);
    for (w = 0; w < workers; w++) { wk[w].p = &p; wk[w].ctx = decoder->pipe_ctx[w]; wk[w].index = w; wk[w].stride = workers; args[w] = &wk[w]; }
    pipe_run(batch_worker, args, workers);
    pthread_mutex_destroy(&p.mu);
  }
  free(blk); free(files); free(groups);
  decoder->batch_kernel_ms = (float)p.kernel_ms; decoder->batch_launches = p.launches;
  if (p.failed) { fprintf(stderr, "SLAB200_Decoder_DecodeBatchPCM: %s\n", slab_last_error()); return SLA_APIRESULT_NG; }
  return SLA_APIRESULT_OK;
}

SLAApiResult SLAB200_Decoder_DecodeWholeDevice(struct SLADecoder* decoder, const uint8_t* d_data,
    uint32_t data_size, int32_t** d_buffer, uint32_t buffer_num_samples, uint32_t* output_num_samples)
{
  struct SLAHeaderInfo header;
  uint8_t head[SLA_HEADER_SIZE];
  SlabDecodeJob job;
  SLAApiResult rc;
  if (decoder == NULL || d_buffer == NULL || d_data == NULL || output_num_samples == NULL)
    return SLA_APIRESULT_INVALID_ARGUMENT;
  slab_ctx_bind(decoder->ctx);      /* the calling thread's current device may differ from the handle's */
  if (data_size < SLA_HEADER_SIZE) return SLA_APIRESULT_INSUFFICIENT_DATA_SIZE;
  if (slab_copy_from_device(decoder->ctx, head, d_data, sizeof(head)) != 0) return SLA_APIRESULT_NG;
  if ((rc = SLADecoder_DecodeHeader(head, sizeof(head), &header)) != SLA_APIRESULT_OK) return rc;
  if ((rc = decoder_header_setup(decoder, &header)) != SLA_APIRESULT_OK) return rc;
  fill_decode_job(decoder, &job);
  job.stream = d_data; job.stream_size = data_size; job.stream_on_device = 1;
  job.blk_byte_off = NULL;                    /* chain is walked on the device */
  job.max_samples = header.num_samples < buffer_num_samples ? header.num_samples : buffer_num_samples;
  job.out = d_buffer; job.out_on_device = 1;
  if (header.num_samples > 0) {
    if (slab_decode(decoder->ctx, &job) != 0) {
      fprintf(stderr, "SLAB200_Decoder_DecodeWholeDevice: %s\n", slab_last_error());
      return SLA_APIRESULT_NG;
    }
    if (job.first_bad_block != 0xFFFFFFFFu) return (SLAApiResult)job.first_bad_code;
    if (job.decoded_samples < header.num_samples) return SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;
  }
  *output_num_samples = job.decoded_samples;
  return SLA_APIRESULT_OK;
}

void SLAB200_Decoder_LastTiming(const struct SLADecoder* decoder, float ms[3], uint32_t* launches)
{
  if (decoder == NULL) return;
  slab_last_timing(decoder->ctx, ms);
  if (launches) *launches = slab_last_launches(decoder->ctx);
}

void SLAB200_Decoder_LastBatchTiming(const struct SLADecoder* decoder, float* kernel_ms, uint32_t* launches)
{
  if (decoder == NULL) return;
  if (kernel_ms) *kernel_ms = decoder->batch_kernel_ms;
  if (launches) *launches = decoder->batch_launches;
}

void SLAB200_Decoder_EnableProfile(struct SLADecoder* decoder, int on) { if (decoder) slab_set_profile(decoder->ctx, on); }
uint32_t SLAB200_Decoder_GetProfile(const struct SLADecoder* decoder, const char** names, float* ms, uint32_t max_entries)
{
  return decoder ? slab_get_profile(decoder->ctx, names, ms, max_entries) : 0;
}

/* ================================================================ streaming decoder ==== */
/* SLAStreamingDecoder_* (src/SLADecoder.c:734-1123, SLADataPacketQueue src/SLAUtility.c:699-900) as a thin
 * host layer over the GPU block decoder.  The data side is the reference's: fragments are queued by
 * pointer (8 packets), drained into a contiguous buffer of two worst-case blocks, handed back through
 * CollectDataFragment once consumed, and the byte-rate estimate follows the block being played.  The
 * decode side differs in granularity: the reference pulls num_output_samples_per_decode samples out of a
 * half-read block through its bit reader; here every block that is complete in the buffer is decoded
 * whole on the device (all of them in one launch sequence) into a sample cache, and Decode() serves its
 * quota from the cache.  A block that has not fully arrived is therefore never touched: Decode returns
 * fewer samples (possibly none) with SLA_APIRESULT_OK while the caller is still appending data, and
 * SLA_APIRESULT_INSUFFICIENT_DATA_SIZE when it is called again without new data and nothing to play. */
#define STRM_PACKETS      8u          /* SLA_STREAMING_DECODE_MAX_NUM_PACKETS, SLAInternal.h:22 */
#define STRM_MARGIN       1.05f       /* SLA_STREAMING_DECODE_NUM_SAMPLES_MARGIN, SLAInternal.h:21 */
#define STRM_MAX_BLOCKS   64u         /* blocks decoded per refill */
#define STRM_MAX_CH       8u          /* SLA_MAX_CHANNELS */

struct StrmPacket { const uint8_t* data; uint32_t size, used; };

struct SLAStreamingDecoder {
  struct SLADecoder*        core;
  struct SLAWaveFormat      wave_format;
  struct SLAEncodeParameter encode_param;
  int                       have_format, have_param;
  float                     decode_interval_hz;
  uint32_t                  max_bit_per_sample;
  uint32_t                  per_decode;               /* num_output_samples_per_decode */
  float                     bytes_per_sample;         /* estimated_bytes_per_sample */
  uint8_t*                  data;                     /* contiguous buffer: blocks not yet decoded */
  uint32_t                  data_size, provided;
  uint8_t*                  image;                    /* 43-byte header + the blocks of one refill */
  struct StrmPacket         packet[STRM_PACKETS];
  uint32_t                  write_pos, read_pos, collect_pos, free_packets;
  int32_t*                  cache[STRM_MAX_CH];       /* decoded, not yet served samples */
  uint32_t                  cache_cap, cache_n, cache_pos;
  uint32_t                  blk_n[STRM_MAX_BLOCKS], blk_bytes[STRM_MAX_BLOCKS];
  uint32_t                  blk_count, blk_at, blk_pos;  /* block being served and the offset inside it */
  int                       fed_since_stall;
};

static uint32_t strm_queue_remain(const struct SLAStreamingDecoder* d)
{
  uint32_t i, total = 0;
  for (i = 0; i < STRM_PACKETS; i++) total += d->packet[i].size - d->packet[i].used;
  return total;
}

/* src/SLADecoder.c:976-984: move as much queued data as fits into the contiguous buffer */
static void strm_drain_queue(struct SLAStreamingDecoder* d)
{
  while (d->free_packets < STRM_PACKETS && d->provided < d->data_size) {
    struct StrmPacket* pk = &d->packet[d->read_pos];
    uint32_t take = pk->size - pk->used;
    if (take == 0) break;                               /* read position has caught up with the writer */
    if (take > d->data_size - d->provided) take = d->data_size - d->provided;
    memcpy(d->data + d->provided, pk->data + pk->used, take);
    d->provided += take; pk->used += take;
    if (pk->used == pk->size) d->read_pos = (d->read_pos + 1u) % STRM_PACKETS;
    else break;
  }
}

struct SLAStreamingDecoder* SLAStreamingDecoder_Create(const struct SLAStreamingDecoderConfig* config)
{
  struct SLAStreamingDecoder* d;
  struct SLADecoderConfig core;
  uint32_t c;
  if (config == NULL || !(config->decode_interval_hz > 0.0f)) return NULL;       /* SLADecoder.c:757-764 */
  if ((d = (struct SLAStreamingDecoder*)calloc(1, sizeof(*d))) == NULL) return NULL;
  core = config->core_config;
  core.verpose_flag = 0;
  if ((d->core = SLADecoder_Create(&core)) == NULL) { free(d); return NULL; }
  d->decode_interval_hz = config->decode_interval_hz;
  d->max_bit_per_sample = config->max_bit_per_sample;
  d->data_size = 2u * SLA_CalculateSufficientBlockSize(core.max_num_channels, core.max_num_block_samples,
                                                       config->max_bit_per_sample);     /* SLADecoder.c:788-791 */
  d->bytes_per_sample = (float)((double)core.max_num_channels * (config->max_bit_per_sample / 8u));
  d->cache_cap = STRM_MAX_BLOCKS * core.max_num_block_samples;
  d->data = (uint8_t*)calloc(d->data_size ? d->data_size : 1u, 1);
  d->image = (uint8_t*)malloc((size_t)SLA_HEADER_SIZE + d->data_size);
  for (c = 0; c < core.max_num_channels && c < STRM_MAX_CH; c++)
    d->cache[c] = (int32_t*)malloc(sizeof(int32_t) * (size_t)(d->cache_cap ? d->cache_cap : 1u));
  d->free_packets = STRM_PACKETS;
  if (d->data == NULL || d->image == NULL) { SLAStreamingDecoder_Destroy(d); return NULL; }
  for (c = 0; c < core.max_num_channels && c < STRM_MAX_CH; c++)
    if (d->cache[c] == NULL) { SLAStreamingDecoder_Destroy(d); return NULL; }
  return d;
}

void SLAStreamingDecoder_Destroy(struct SLAStreamingDecoder* d)
{
  uint32_t c;
  if (d == NULL) return;
  SLADecoder_Destroy(d->core);
  for (c = 0; c < STRM_MAX_CH; c++) free(d->cache[c]);
  free(d->data); free(d->image); free(d);
}

SLAApiResult SLAStreamingDecoder_SetWaveFormat(struct SLAStreamingDecoder* d, const struct SLAWaveFormat* w)
{
  SLAApiResult rc;
  if (d == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if ((rc = SLADecoder_SetWaveFormat(d->core, w)) != SLA_APIRESULT_OK) return rc;
  if (w->bit_per_sample > d->max_bit_per_sample) return SLA_APIRESULT_EXCEED_HANDLE_CAPACITY;   /* SLADecoder.c:839-841 */
  d->wave_format = *w; d->have_format = 1;
  d->per_decode = (uint32_t)ceil(STRM_MARGIN * (float)w->sampling_rate / d->decode_interval_hz);   /* :844-845 */
  return SLA_APIRESULT_OK;
}

SLAApiResult SLAStreamingDecoder_SetEncodeParameter(struct SLAStreamingDecoder* d, const struct SLAEncodeParameter* p)
{
  SLAApiResult rc;
  if (d == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if ((rc = SLADecoder_SetEncodeParameter(d->core, p)) != SLA_APIRESULT_OK) return rc;
  d->encode_param = *p; d->have_param = 1;
  return SLA_APIRESULT_OK;
}

/* SLADecoder.c:862-884 */
SLAApiResult SLAStreamingDecoder_EstimateMinimumNessesaryDataSize(struct SLAStreamingDecoder* d, uint32_t* v)
{
  uint32_t est;
  if (d == NULL || v == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  est = (uint32_t)ceil((double)d->bytes_per_sample * d->per_decode);
  *v = est > MIN_BLOCK_HEADER ? est : MIN_BLOCK_HEADER;
  return SLA_APIRESULT_OK;
}

/* SLADecoder.c:887-913 */
SLAApiResult SLAStreamingDecoder_EstimateDecodableNumSamples(struct SLAStreamingDecoder* d, uint32_t* v)
{
  uint32_t remain;
  SLAApiResult rc;
  if (d == NULL || v == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if ((rc = SLAStreamingDecoder_GetRemainDataSize(d, &remain)) != SLA_APIRESULT_OK) return rc;
  *v = (uint32_t)floor((float)remain / d->bytes_per_sample);
  return SLA_APIRESULT_OK;
}

SLAApiResult SLAStreamingDecoder_GetOutputNumSamplesPerDecode(struct SLAStreamingDecoder* d, uint32_t* v)
{
  if (d == NULL || v == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  *v = d->per_decode;
  return SLA_APIRESULT_OK;
}

/* SLADecoder.c:932-958: queued bytes + buffered bytes not yet played.  The reference subtracts its bit
 * reader's position inside the current block; a block decoded whole counts in proportion to the samples
 * of it that were served. */
SLAApiResult SLAStreamingDecoder_GetRemainDataSize(struct SLAStreamingDecoder* d, uint32_t* v)
{
  uint32_t b, pending = 0;
  if (d == NULL || v == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  for (b = d->blk_at; b < d->blk_count; b++) {
    uint32_t bytes = d->blk_bytes[b];
    if (b == d->blk_at && d->blk_n[b] > 0)
      bytes -= (uint32_t)((uint64_t)bytes * d->blk_pos / d->blk_n[b]);
    pending += bytes;
  }
  *v = strm_queue_remain(d) + d->provided + pending;
  return SLA_APIRESULT_OK;
}

/* SLADecoder.c:960-986 with SLAUtility.c:733-766 */
SLAApiResult SLAStreamingDecoder_AppendDataFragment(struct SLAStreamingDecoder* d, const uint8_t* p, uint32_t n)
{
  if (d == NULL || p == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (d->free_packets == 0) return SLA_APIRESULT_EXCEED_HANDLE_CAPACITY;
  if (n > 0) {
    struct StrmPacket* pk = &d->packet[d->write_pos];
    pk->data = p; pk->size = n; pk->used = 0;
    d->write_pos = (d->write_pos + 1u) % STRM_PACKETS;
    d->free_packets--;
    d->fed_since_stall = 1;
  }
  strm_drain_queue(d);
  return SLA_APIRESULT_OK;
}

/* SLADecoder.c:989-1005 with SLAUtility.c:822-870: the consumed front of the oldest packet */
SLAApiResult SLAStreamingDecoder_CollectDataFragment(struct SLAStreamingDecoder* d, const uint8_t** p, uint32_t* n)
{
  struct StrmPacket* pk;
  if (d == NULL || p == NULL || n == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (d->free_packets == STRM_PACKETS) return SLA_APIRESULT_NO_DATA_FRAGMENTS;
  pk = &d->packet[d->collect_pos];
  if (pk->used == 0) return SLA_APIRESULT_NO_DATA_FRAGMENTS;
  *p = pk->data; *n = pk->used;
  pk->data += pk->used; pk->size -= pk->used; pk->used = 0;
  if (pk->size == 0) {
    d->collect_pos = (d->collect_pos + 1u) % STRM_PACKETS;
    d->free_packets++;
  }
  return SLA_APIRESULT_OK;
}

/* decode every block that is complete in the buffer into the sample cache; 0 blocks is not an error */
static SLAApiResult strm_refill(struct SLAStreamingDecoder* d)
{
  struct SLAHeaderInfo h;
  uint32_t off = 0, nb = 0, total = 0, got = 0;
  SLAApiResult rc;
  strm_drain_queue(d);
  d->cache_n = d->cache_pos = 0;
  d->blk_count = d->blk_at = d->blk_pos = 0;
  while (nb < STRM_MAX_BLOCKS && d->provided - off >= MIN_BLOCK_HEADER) {
    const uint8_t* b = d->data + off;
    uint32_t bsize, n;
    if (b[0] != 0xFF || b[1] != 0xFF) { if (nb > 0) break; return SLA_APIRESULT_FAILED_TO_FIND_SYNC_CODE; }   /* SLADecoder.c:334-337 */
    bsize = (((uint32_t)b[2] << 24) | ((uint32_t)b[3] << 16) | ((uint32_t)b[4] << 8) | b[5]) + 6u;
    n = ((uint32_t)b[8] << 8) | b[9];
    if (bsize < 10u || bsize > d->data_size) { if (nb > 0) break; return SLA_APIRESULT_INSUFFICIENT_DATA_SIZE; }
    if (bsize > d->provided - off || total + n > d->cache_cap) break;
    d->blk_n[nb] = n; d->blk_bytes[nb] = bsize;
    nb++; off += bsize; total += n;
  }
  if (nb == 0) return SLA_APIRESULT_OK;
  memset(&h, 0, sizeof(h));
  h.wave_format = d->wave_format; h.encode_param = d->encode_param;
  h.num_samples = total; h.num_blocks = nb;
  h.max_block_size = SLA_MAX_BLOCK_SIZE_INVAILD; h.max_bit_per_second = 0;
  if ((rc = SLAEncoder_EncodeHeader(&h, d->image, SLA_HEADER_SIZE)) != SLA_APIRESULT_OK) return rc;
  memcpy(d->image + SLA_HEADER_SIZE, d->data, off);
  rc = SLADecoder_DecodeWhole(d->core, d->image, SLA_HEADER_SIZE + off, d->cache, d->cache_cap, &got);
  if (rc != SLA_APIRESULT_OK) return rc;
  memmove(d->data, d->data + off, d->provided - off);                 /* SLADecoder.c:1083-1086 */
  d->provided -= off;
  d->cache_n = got; d->blk_count = nb;
  strm_drain_queue(d);
  return SLA_APIRESULT_OK;
}

/* SLADecoder.c:1008-1123 */
SLAApiResult SLAStreamingDecoder_Decode(struct SLAStreamingDecoder* d, int32_t** buffer, uint32_t buffer_num_samples,
    uint32_t* num_output_samples)
{
  uint32_t goal, progress = 0, c;
  if (d == NULL || buffer == NULL || num_output_samples == NULL) return SLA_APIRESULT_INVALID_ARGUMENT;
  if (!d->have_format || !d->have_param) return SLA_APIRESULT_PARAMETER_NOT_SET;
  goal = buffer_num_samples < d->per_decode ? buffer_num_samples : d->per_decode;
  while (progress < goal) {
    uint32_t take;
    if (d->cache_pos == d->cache_n) {
      SLAApiResult rc = strm_refill(d);
      if (rc != SLA_APIRESULT_OK) return rc;
      if (d->cache_n == 0) break;                       /* the next block has not fully arrived */
    }
    take = d->cache_n - d->cache_pos;
    if (take > goal - progress) take = goal - progress;
    for (c = 0; c < d->wave_format.num_channels; c++)
      memcpy(buffer[c] + progress, d->cache[c] + d->cache_pos, sizeof(int32_t) * take);
    d->cache_pos += take; progress += take;
    /* follow the block being played: its byte rate is the estimate (SLADecoder.c:1048-1050) */
    d->blk_pos += take;
    while (d->blk_at < d->blk_count && d->blk_pos >= d->blk_n[d->blk_at]) {
      d->blk_pos -= d->blk_n[d->blk_at];
      d->blk_at++;
    }
    if (d->blk_at < d->blk_count && d->blk_n[d->blk_at] > 0)
      d->bytes_per_sample = (float)((double)d->blk_bytes[d->blk_at] / d->blk_n[d->blk_at]);
    else if (d->blk_count > 0 && d->blk_n[d->blk_count - 1u] > 0)
      d->bytes_per_sample = (float)((double)d->blk_bytes[d->blk_count - 1u] / d->blk_n[d->blk_count - 1u]);
  }
  *num_output_samples = progress;
  if (progress == 0 && goal > 0) {
    if (!d->fed_since_stall) return SLA_APIRESULT_INSUFFICIENT_DATA_SIZE;
    d->fed_since_stall = 0;
  }
  return SLA_APIRESULT_OK;
}
