/*
 * slab_device.h - the thin C-ABI between the host-side C code (slab_host.c: handles, argument
 * checks, container header, block chain) and the CUDA translation units (slab_decode.cu,
 * slab_encode.cu).  Plain pointers and sizes only.
 */
#ifndef SLAB_DEVICE_H
#define SLAB_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SlabCtx SlabCtx;   /* one per handle: device ordinal, stream, grow-only arenas */

/* NULL when no usable CUDA device exists (the public Create functions then fail: no CPU path). */
SlabCtx* slab_ctx_create(void);
void     slab_ctx_destroy(SlabCtx* ctx);
const char* slab_last_error(void);
void     slab_set_error_text(const char* text);
/* Non-zero when this library was built for the host simulator (tests only). */
int      slab_is_hostsim(void);

/* ---- timing probes: device milliseconds of the stages of the last call (CUDA events) ---- */
enum { SLAB_T_H2D = 0, SLAB_T_KERNELS = 1, SLAB_T_D2H = 2, SLAB_T_COUNT = 3 };
void     slab_last_timing(const SlabCtx* ctx, float ms[SLAB_T_COUNT]);
uint32_t slab_last_launches(const SlabCtx* ctx);

/* per-kernel device timing of the next calls (CUDA events around every launch) */
void     slab_set_profile(SlabCtx* ctx, int on);
uint32_t slab_get_profile(const SlabCtx* ctx, const char** names, float* ms, uint32_t max_entries);

/* ---- support for the pipelined whole-file calls in slab_host.c (several contexts, one host thread
 * each, chunks of the file in flight on different streams) ---- */
void  slab_ctx_bind(SlabCtx* ctx);                                   /* make ctx's device current in this thread */
void* slab_user_buffer(SlabCtx* ctx, int which, size_t bytes);      /* grow-only device buffer, which = 0..7 */
int   slab_upload_async(SlabCtx* ctx, void* dst_device, const void* src_host, size_t bytes);
int   slab_download_async(SlabCtx* ctx, void* dst_host, const void* src_device, size_t bytes);
int   slab_stream_sync(SlabCtx* ctx);
int   slab_copy_d2d_async(SlabCtx* ctx, void* dst_device, const void* src_device, size_t bytes);
int   slab_profile_enabled(const SlabCtx* ctx);
/* ordered transfers (see slab_ctx.cu): uploads on the context's copy stream complete in issue order;
 * marks are events on that stream other contexts' streams can wait for */
int   slab_host_is_pinned(const void* p);
int   slab_xfer_prepare(SlabCtx* ctx);
uint32_t slab_xfer_piece_bytes(void);
int   slab_xfer_upload(SlabCtx* ctx, void* dst_device, const void* src_host, size_t bytes);
int   slab_xfer_upload_staged(SlabCtx* ctx, uint32_t slot, void* dst_device, const void* src_host, size_t bytes);
int   slab_xfer_mark(SlabCtx* ctx, uint32_t index);
int   slab_xfer_wait(SlabCtx* waiter, SlabCtx* owner, uint32_t index);
int   slab_xfer_sync(SlabCtx* ctx);
int   slab_join_hi(SlabCtx* ctx);
int   slab_download(SlabCtx* ctx, void* dst_host, const void* src_device, size_t bytes, int dst_pinned);
int   slab_span_begin(SlabCtx* ctx);
int   slab_span_end(SlabCtx* ctx, uint32_t launches);

/* interleaved little-endian PCM (WAV data-chunk layout) <-> planar left-justified int32, both in
 * device memory, asynchronous on the context's stream (slab_pcm.cu) */
int   slab_pcm_to_planar(SlabCtx* ctx, int32_t* d_planes, size_t plane_stride, const void* d_pcm,
                         uint32_t num_channels, uint32_t bytes_per_sample, uint32_t num_frames);
/* merged jobs: file f's frames start at byte pcm_off[f] (a multiple of 16) of d_pcm and go to plane offset
 * file_start[f] (a multiple of 1024); the planes are zero-filled from the end of a file to the next start
 * (plane_len = end of the last file rounded up to 1024).  file_start / file_len / pcm_off: host arrays. */
int   slab_pcm_to_planar_files(SlabCtx* ctx, int32_t* d_planes, size_t plane_stride, uint32_t plane_len, const void* d_pcm,
                               uint32_t num_channels, uint32_t bytes_per_sample, uint32_t num_files,
                               const uint32_t* file_start, const uint32_t* file_len, const uint64_t* pcm_off);
int   slab_planar_to_pcm(SlabCtx* ctx, void* d_pcm, const int32_t* d_planes, size_t plane_stride,
                         uint32_t num_channels, uint32_t bytes_per_sample, uint32_t num_frames);

/* small synchronous copies (container header) */
int slab_copy_to_device(SlabCtx* ctx, void* dst_device, const void* src_host, size_t bytes);
int slab_copy_from_device(SlabCtx* ctx, void* dst_host, const void* src_device, size_t bytes);

/* ------------------------------------------------------------------ decode ---- */
typedef struct SlabDecodeJob {
  /* from the file header */
  uint32_t num_channels, bits_per_sample, offset_lshift;
  uint32_t parcor_order, longterm_order, lms_order, ch_process;
  uint32_t check_crc;
  /* stream */
  const uint8_t* stream;       /* whole .sla image (host or device memory) */
  uint32_t stream_size;
  int      stream_on_device;
  /* block chain produced by the caller's walk (host arrays, num_blocks entries each);
   * may be NULL with stream_on_device: the chain is then walked on the device */
  uint32_t num_blocks;
  const uint32_t* blk_byte_off;
  const uint32_t* blk_smp_off;
  const uint32_t* blk_nsmp;
  uint32_t total_samples;      /* samples covered by the chain */
  uint32_t max_samples;        /* header num_samples: upper bound for a device-side walk */
  uint32_t max_block_samples;  /* header field; 0 = unknown */
  /* output: num_channels planar pointers (host or device) with room for total_samples each */
  int32_t* const* out;
  int      out_on_device;
  /* results */
  uint32_t decoded_blocks;     /* device walk: blocks found */
  uint32_t decoded_samples;
  uint32_t first_bad_block;    /* 0xFFFFFFFF when none */
  uint32_t first_bad_code;     /* SLAApiResult value */
  uint32_t* blk_err_out;       /* optional host array, num_blocks entries: SLAApiResult per block (0 = fine) */
} SlabDecodeJob;

/* 0 on success (per-block stream errors are reported in the job), -1 on a CUDA failure. */
int slab_decode(SlabCtx* ctx, SlabDecodeJob* job);

/* ------------------------------------------------------------------ encode ---- */
/* per file of a merged job: what SLAEncoder_EncodeWhole would have returned for that file alone */
typedef struct SlabFileResult {
  uint32_t offset_lshift, num_blocks, byte_offset, num_bytes, max_block_size, max_bit_per_second;
} SlabFileResult;

typedef struct SlabEncodeJob {
  uint32_t num_channels, bits_per_sample, sampling_rate;
  uint32_t parcor_order, longterm_order, lms_order, ch_process, window_type;
  uint32_t max_block_samples;
  uint32_t fft_size;           /* reference handle property; only its scale enters thresholds */
  /* input: num_channels planar pointers, left-justified int32 (host or device) */
  const int32_t* const* input;
  int      input_on_device;
  uint32_t num_samples;
  /* range mode (multi-GPU sharding): encode only [first_sample, first_sample + num_samples) of a
   * longer file whose offset_lshift was agreed beforehand; <0 = compute it from this input */
  int32_t  forced_lshift;
  /* chunk mode (pipelined whole-file encode): the segment chain starts at first_sample and stops
   * starting segments at soft_end (0 = none); the last segment may run past soft_end, which is why
   * num_samples extends one block beyond it.  consumed_samples = where the next chunk starts; it is
   * also handed to on_consumed as soon as the chain is known, long before the chunk is encoded. */
  uint32_t first_sample, soft_end;
  void   (*on_consumed)(void* user, uint32_t consumed_samples);
  void*    user;
  uint32_t consumed_samples;
  int      high_priority;      /* chunk mode: every kernel of this job on the context's high-priority stream (the last
                                * chunk of a pipelined call: what is left to do after the last byte has arrived) */
  /* merged mode (many files of one format and parameter set in one launch sequence): the planes hold
   * num_files files back to back, file f at plane offset file_start[f] - ascending multiples of 1024, the gap
   * up to the next start zero-filled - with file_len[f] > 0 samples; num_samples covers the last file.  Every
   * file gets its own segment chain, offset_lshift and statistics (files[f]); its blocks are consecutive in the
   * output, in file order.  Host arrays.  Excludes range, chunk and single-block mode. */
  uint32_t num_files;
  uint32_t reserve_samples;    /* size the per-sample arenas for at least this many samples per channel: the groups of
                                * one batch call then allocate once per context instead of growing from group to group */
  const uint32_t* file_start;
  const uint32_t* file_len;
  struct SlabFileResult* files;
  int      single_block;       /* SLAEncoder_EncodeBlock: exactly one block, no partition search */
  int      mask_only;          /* only compute input_or_mask */
  /* output: block bytes are written from out + out_offset; capacity in bytes */
  uint8_t* out;
  int      out_on_device;
  uint32_t out_capacity;
  uint32_t out_offset;
  /* results */
  uint32_t offset_lshift;
  uint32_t num_blocks;
  uint32_t total_bytes;        /* bytes of all blocks (excludes the 43-byte file header) */
  uint32_t max_block_size;
  uint32_t max_bit_per_second;
  uint32_t input_or_mask;      /* OR of every input word (range mode: shards combine these) */
  int      overflow;           /* 1 when out_capacity was too small: nothing was written */
  /* block x channels that took a fallback path: the reference's FFT autocorrelation (listings, duplicates
   * possible), lag sums in the reference's order, scalar long-term lag sums (residual >= 2^23) */
  uint32_t fallback_ltfft, fallback_exact_autocorr, fallback_scalar_ltcorr;
  /* optional debug export (host pointers, may be NULL) */
  struct SlabBlockRecord* records;
  uint32_t max_records;
  int32_t* const* residual_out;   /* host planar, num_samples each */
} SlabEncodeJob;

/* Mirrors oracle OraBlock / RefWBBlock field for field so tests can diff them directly. */
typedef struct SlabBlockRecord {
  uint32_t sample_offset, num_samples, block_type, block_size, byte_offset;
  uint32_t rshift[8];
  uint32_t pitch[8];
  int32_t  parcor_code[8][65];
  int32_t  lt_q31[8][8];
  uint64_t rice_init[8];
  double   parcor[8][65];
  double   lt[8][8];
} SlabBlockRecord;

int slab_encode(SlabCtx* ctx, SlabEncodeJob* job);

/* test hook: the encoder's long-term analysis on a caller-supplied residual (one block, one channel) */
int slab_debug_longterm(SlabCtx* ctx, const int32_t* data, uint32_t n, uint32_t taps, uint32_t fft_size,
                        uint32_t* pitch, double* coef);

#ifdef __cplusplus
}
#endif
#endif
