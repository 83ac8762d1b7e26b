/*
 * slab_host.c - host side of libsla_b200.so in plain C: the SLA public API (handles, capacity and
 * argument checks, the 43-byte container header, the block chain walk) on top of the CUDA layer
 * behind slab_device.h.  No sample ever passes through a CPU codec here: every encode/decode call
 * ends in slab_encode()/slab_decode(), and handle creation fails when there is no CUDA device.
 *
 * Mirrors, function by function, src/SLAEncoder.c:56-292,804-932 and src/SLADecoder.c:68-305,660-732
 * of the reference (same status codes in the same situations; see tests/test_api_errors.py).
 */
#include "sla_b200.h"
#include "slab_device.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define FLAG_WAVE_FORMAT   1u
#define FLAG_ENCODE_PARAM  2u
#define MIN_BLOCK_SAMPLES  2048u     /* SLAInternal.h:15 */
#define MIN_BLOCK_HEADER   11u       /* SLAInternal.h:35 */

struct SLAEncoder {
  struct SLAEncoderConfig   config;
  struct SLAWaveFormat      wave_format;
  struct SLAEncodeParameter encode_param;
  uint32_t                  status;
  SlabCtx*                  ctx;
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
  uint32_t*                 chain;        /* host block table: off | smp | n, grown on demand */
  uint32_t                  chain_cap;
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

const char* SLAB200_LastError(void) { return slab_last_error(); }

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
  if (encoder == NULL) return;
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
  if ((rc = encoder_precheck(encoder)) != SLA_APIRESULT_OK) return rc;

  fill_job(encoder, &job);
  job.input = input; job.input_on_device = on_device; job.num_samples = num_samples;
  job.out = data; job.out_on_device = on_device; job.out_capacity = data_size;
  job.out_offset = SLA_HEADER_SIZE;
  job.records = (struct SlabBlockRecord*)encoder->dbg_records;
  job.max_records = encoder->dbg_max_records;
  job.residual_out = encoder->dbg_residual;
  if (num_samples > 0) {
    if (slab_encode(encoder->ctx, &job) != 0) {
      fprintf(stderr, "SLAEncoder_EncodeWhole: %s\n", slab_last_error());
      return SLA_APIRESULT_NG;
    }
    if (job.overflow) return SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE;   /* SLAEncoder.c:848,915 */
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

/* One block with the handle's current offset_lshift and no partition search, SLAEncoder.c:458-801 */
SLAApiResult SLAEncoder_EncodeBlock(struct SLAEncoder* encoder, const int32_t* const* input,
    uint32_t num_samples, uint8_t* data, uint32_t data_size, uint32_t* output_size)
{
  SlabEncodeJob job;
  SLAApiResult rc;
  if (encoder == NULL || input == NULL || data == NULL || output_size == NULL)
    return SLA_APIRESULT_INVALID_ARGUMENT;
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
  if (decoder == NULL) return;
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

SLAApiResult SLADecoder_DecodeWhole(struct SLADecoder* decoder, const uint8_t* data, uint32_t data_size,
    int32_t** buffer, uint32_t buffer_num_samples, uint32_t* output_num_samples)
{
  struct SLAHeaderInfo header;
  SlabDecodeJob job;
  SLAApiResult rc, walk_rc = SLA_APIRESULT_OK;
  uint32_t off = SLA_HEADER_SIZE, smp = 0, nb = 0, cap;

  if (decoder == NULL || buffer == NULL || data == NULL || output_num_samples == NULL)
    return SLA_APIRESULT_INVALID_ARGUMENT;
  if ((rc = SLADecoder_DecodeHeader(data, data_size, &header)) != SLA_APIRESULT_OK) return rc;
  if ((rc = decoder_header_setup(decoder, &header)) != SLA_APIRESULT_OK) return rc;

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
    if (slab_decode(decoder->ctx, &job) != 0) {
      fprintf(stderr, "SLADecoder_DecodeWhole: %s\n", slab_last_error());
      return SLA_APIRESULT_NG;
    }
    if (job.first_bad_block != 0xFFFFFFFFu) return (SLAApiResult)job.first_bad_code;
  }
  if (walk_rc != SLA_APIRESULT_OK) return walk_rc;
  *output_num_samples = smp;
  if (decoder->config.verpose_flag != 0) { printf("progress:100%% \r"); fflush(stdout); }
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

void SLAB200_Decoder_EnableProfile(struct SLADecoder* decoder, int on) { if (decoder) slab_set_profile(decoder->ctx, on); }
uint32_t SLAB200_Decoder_GetProfile(const struct SLADecoder* decoder, const char** names, float* ms, uint32_t max_entries)
{
  return decoder ? slab_get_profile(decoder->ctx, names, ms, max_entries) : 0;
}

/* ================================================================ streaming decoder stubs ==== */
struct SLAStreamingDecoder* SLAStreamingDecoder_Create(const struct SLAStreamingDecoderConfig* config)
{
  (void)config;
  return NULL;
}
void SLAStreamingDecoder_Destroy(struct SLAStreamingDecoder* d) { (void)d; }
SLAApiResult SLAStreamingDecoder_SetWaveFormat(struct SLAStreamingDecoder* d, const struct SLAWaveFormat* w)
{ (void)d; (void)w; return SLA_APIRESULT_NG; }
SLAApiResult SLAStreamingDecoder_SetEncodeParameter(struct SLAStreamingDecoder* d, const struct SLAEncodeParameter* p)
{ (void)d; (void)p; return SLA_APIRESULT_NG; }
SLAApiResult SLAStreamingDecoder_EstimateMinimumNessesaryDataSize(struct SLAStreamingDecoder* d, uint32_t* v)
{ (void)d; (void)v; return SLA_APIRESULT_NG; }
SLAApiResult SLAStreamingDecoder_EstimateDecodableNumSamples(struct SLAStreamingDecoder* d, uint32_t* v)
{ (void)d; (void)v; return SLA_APIRESULT_NG; }
SLAApiResult SLAStreamingDecoder_GetOutputNumSamplesPerDecode(struct SLAStreamingDecoder* d, uint32_t* v)
{ (void)d; (void)v; return SLA_APIRESULT_NG; }
SLAApiResult SLAStreamingDecoder_AppendDataFragment(struct SLAStreamingDecoder* d, const uint8_t* p, uint32_t n)
{ (void)d; (void)p; (void)n; return SLA_APIRESULT_NG; }
SLAApiResult SLAStreamingDecoder_CollectDataFragment(struct SLAStreamingDecoder* d, const uint8_t** p, uint32_t* n)
{ (void)d; (void)p; (void)n; return SLA_APIRESULT_NG; }
SLAApiResult SLAStreamingDecoder_GetRemainDataSize(struct SLAStreamingDecoder* d, uint32_t* v)
{ (void)d; (void)v; return SLA_APIRESULT_NG; }
SLAApiResult SLAStreamingDecoder_Decode(struct SLAStreamingDecoder* d, int32_t** b, uint32_t n, uint32_t* o)
{ (void)d; (void)b; (void)n; (void)o; return SLA_APIRESULT_NG; }
