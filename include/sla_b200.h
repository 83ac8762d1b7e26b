/*
 * sla_b200.h - C ABI of libsla_b200.so: the SLA block encode/decode path on NVIDIA B200 (sm_100a).
 *
 * Section 1 is the drop-in boundary: the same symbols, structure layouts, enumerator values and
 * error behaviour as the reference library's public headers, so a program written against
 *   src/include/public/SLA.h          (result codes, wave format / encode parameter / header structs)
 *   src/include/public/SLAEncoder.h   (SLAEncoder_* : lines 28-53)
 *   src/include/public/SLADecoder.h   (SLADecoder_* : lines 39-59, SLAStreamingDecoder_* : 62-101)
 * links against this library unchanged (the shim headers SLA.h / SLAEncoder.h / SLADecoder.h in
 * this directory simply include this file).  All sample buffers at this boundary are HOST memory,
 * planar int32, left-justified to 32 bits, exactly as in the reference.
 *
 * Section 2 adds entry points that have no reference counterpart: device-resident input/output
 * (what bench.py times as the kernel-only figure), shard-range encoding for multi-GPU runs, a
 * debug export of per-block encoder intermediates for parity tests, and timing probes.
 *
 * There is no CPU implementation behind any of these calls: the Create functions return NULL when
 * no CUDA device is usable (SLAB200_LastError() says why).
 */
#ifndef SLA_B200_H_INCLUDED
#define SLA_B200_H_INCLUDED

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ===================================================================== 1. drop-in boundary ==== */

#define SLA_VERSION_STRING          "1.0.0"
#define SLA_FORMAT_VERSION          1
#define SLA_HEADER_SIZE             43
#define SLA_BLOCK_HEADER_SIZE       10
#define SLA_NUM_SAMPLES_INVALID     0xFFFFFFFF
#define SLA_NUM_BLOCKS_INVALID      0xFFFFFFFF
#define SLA_MAX_BLOCK_SIZE_INVAILD  0xFFFFFFFF
#define SLA_ENCODER_VERSION_STRING  "0.0.1(beta)"
#define SLA_DECODER_VERSION_STRING  "0.0.1(beta)"

/* replaces SLA.h:22-23 */
#define SLA_CalculateSufficientBlockSize(num_channels, num_samples, bit_per_sample) \
  (2 * (num_channels) * (num_samples) * ((bit_per_sample) / 8))

/* replaces SLA.h:26-43 (values are part of the ABI) */
typedef enum SLAApiResultTag {
  SLA_APIRESULT_OK = 0,
  SLA_APIRESULT_NG,
  SLA_APIRESULT_INVALID_ARGUMENT,
  SLA_APIRESULT_EXCEED_HANDLE_CAPACITY,
  SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE,
  SLA_APIRESULT_INVAILD_CHPROCESSMETHOD,
  SLA_APIRESULT_FAILED_TO_CALCULATE_COEF,
  SLA_APIRESULT_FAILED_TO_PREDICT,
  SLA_APIRESULT_FAILED_TO_SYNTHESIZE,
  SLA_APIRESULT_INSUFFICIENT_DATA_SIZE,
  SLA_APIRESULT_INVALID_HEADER_FORMAT,
  SLA_APIRESULT_DETECT_DATA_CORRUPTION,
  SLA_APIRESULT_FAILED_TO_FIND_SYNC_CODE,
  SLA_APIRESULT_INVALID_WINDOWFUNCTION_TYPE,
  SLA_APIRESULT_NO_DATA_FRAGMENTS,
  SLA_APIRESULT_PARAMETER_NOT_SET
} SLAApiResult;

/* replaces SLA.h:46-58 */
typedef enum SLAChannelProcessMethodTag {
  SLA_CHPROCESSMETHOD_NONE = 0,
  SLA_CHPROCESSMETHOD_STEREO_MS
} SLAChannelProcessMethod;

typedef enum SLAWindowFunctionTypeTag {
  SLA_WINDOWFUNCTIONTYPE_RECTANGULAR = 0,
  SLA_WINDOWFUNCTIONTYPE_SIN,
  SLA_WINDOWFUNCTIONTYPE_HANN,
  SLA_WINDOWFUNCTIONTYPE_BLACKMAN,
  SLA_WINDOWFUNCTIONTYPE_VORBIS
} SLAWindowFunctionType;

/* replaces SLA.h:61-86 (field order and types are part of the ABI) */
struct SLAWaveFormat {
  uint32_t num_channels;
  uint32_t bit_per_sample;
  uint32_t sampling_rate;
  uint8_t  offset_lshift;
};

struct SLAEncodeParameter {
  uint32_t                parcor_order;
  uint32_t                longterm_order;         /* taps; odd */
  uint32_t                lms_order_per_filter;   /* power of two >= 4 */
  SLAChannelProcessMethod ch_process_method;
  SLAWindowFunctionType   window_function_type;
  uint32_t                max_num_block_samples;
};

struct SLAHeaderInfo {
  struct SLAWaveFormat      wave_format;
  struct SLAEncodeParameter encode_param;
  uint32_t                  num_samples;
  uint32_t                  num_blocks;
  uint32_t                  max_block_size;
  uint32_t                  max_bit_per_second;
};

/* replaces SLAEncoder.h:11-21 */
struct SLAEncoder;
struct SLAEncoderConfig {
  uint32_t max_num_channels;
  uint32_t max_num_block_samples;
  uint32_t max_parcor_order;
  uint32_t max_longterm_order;
  uint32_t max_lms_order_per_filter;
  uint8_t  verpose_flag;
};

/* replaces SLADecoder.h:11-32 */
struct SLADecoder;
struct SLAStreamingDecoder;
struct SLADecoderConfig {
  uint32_t max_num_channels;
  uint32_t max_num_block_samples;
  uint32_t max_parcor_order;
  uint32_t max_longterm_order;
  uint32_t max_lms_order_per_filter;
  uint8_t  enable_crc_check;
  uint8_t  verpose_flag;
};
struct SLAStreamingDecoderConfig {
  struct SLADecoderConfig core_config;
  float                   decode_interval_hz;
  uint32_t                max_bit_per_sample;
};

/* ---- encoder: replaces SLAEncoder.h:28-53 / src/SLAEncoder.c:56-932 ----
 * Supported envelope (everything the reference's presets and CLI use): 1..8 channels, PARCOR order 1..64, 1/3/5/7
 * long-term taps, 4/8/16/32 LMS taps, max_num_block_samples 2048..16384.  Blocks above 16384 samples (the partition
 * search tables hold 17 nodes) are refused with SLA_APIRESULT_EXCEED_HANDLE_CAPACITY, as the reference refuses
 * parameters above its handle's configuration; the other out-of-range parameters return the reference's own codes
 * (even tap counts: FAILED_TO_CALCULATE_COEF, LMS sizes that are not a power of two >= 4: FAILED_TO_PREDICT). */
struct SLAEncoder* SLAEncoder_Create(const struct SLAEncoderConfig* config);
void SLAEncoder_Destroy(struct SLAEncoder* encoder);
SLAApiResult SLAEncoder_SetWaveFormat(struct SLAEncoder* encoder, const struct SLAWaveFormat* wave_format);
SLAApiResult SLAEncoder_SetEncodeParameter(struct SLAEncoder* encoder, const struct SLAEncodeParameter* encode_param);
SLAApiResult SLAEncoder_EncodeHeader(const struct SLAHeaderInfo* header, uint8_t* data, uint32_t data_size);
SLAApiResult SLAEncoder_EncodeBlock(struct SLAEncoder* encoder, const int32_t* const* input,
    uint32_t num_samples, uint8_t* data, uint32_t data_size, uint32_t* output_size);
SLAApiResult SLAEncoder_EncodeWhole(struct SLAEncoder* encoder, const int32_t* const* input,
    uint32_t num_samples, uint8_t* data, uint32_t data_size, uint32_t* output_size);

/* ---- decoder: replaces SLADecoder.h:39-59 / src/SLADecoder.c:68-732 ---- */
SLAApiResult SLADecoder_DecodeHeader(const uint8_t* data, uint32_t data_size, struct SLAHeaderInfo* header_info);
struct SLADecoder* SLADecoder_Create(const struct SLADecoderConfig* config);
void SLADecoder_Destroy(struct SLADecoder* decoder);
SLAApiResult SLADecoder_SetWaveFormat(struct SLADecoder* decoder, const struct SLAWaveFormat* wave_format);
SLAApiResult SLADecoder_SetEncodeParameter(struct SLADecoder* decoder, const struct SLAEncodeParameter* encode_param);
SLAApiResult SLADecoder_DecodeWhole(struct SLADecoder* decoder, const uint8_t* data, uint32_t data_size,
    int32_t** buffer, uint32_t buffer_num_samples, uint32_t* output_num_samples);

/* ---- streaming decoder (SLADecoder.h:62-101): a host layer over the block decoder - the reference's fragment queue,
 * estimates and result codes; complete blocks are decoded whole on the device and Decode serves its
 * quota from that cache (see INTEGRATION.md section 2 for the one behavioural difference). ---- */
struct SLAStreamingDecoder* SLAStreamingDecoder_Create(const struct SLAStreamingDecoderConfig* config);
void SLAStreamingDecoder_Destroy(struct SLAStreamingDecoder* decoder);
SLAApiResult SLAStreamingDecoder_SetWaveFormat(struct SLAStreamingDecoder* decoder, const struct SLAWaveFormat* wave_format);
SLAApiResult SLAStreamingDecoder_SetEncodeParameter(struct SLAStreamingDecoder* decoder, const struct SLAEncodeParameter* encode_param);
SLAApiResult SLAStreamingDecoder_EstimateMinimumNessesaryDataSize(struct SLAStreamingDecoder* decoder, uint32_t* estimate_data_size);
SLAApiResult SLAStreamingDecoder_EstimateDecodableNumSamples(struct SLAStreamingDecoder* decoder, uint32_t* estimate_num_samples);
SLAApiResult SLAStreamingDecoder_GetOutputNumSamplesPerDecode(struct SLAStreamingDecoder* decoder, uint32_t* output_num_samples);
SLAApiResult SLAStreamingDecoder_AppendDataFragment(struct SLAStreamingDecoder* decoder, const uint8_t* data, uint32_t data_size);
SLAApiResult SLAStreamingDecoder_CollectDataFragment(struct SLAStreamingDecoder* decoder, const uint8_t** data_ptr, uint32_t* data_size);
SLAApiResult SLAStreamingDecoder_GetRemainDataSize(struct SLAStreamingDecoder* decoder, uint32_t* remain_data_size);
SLAApiResult SLAStreamingDecoder_Decode(struct SLAStreamingDecoder* decoder, int32_t** buffer, uint32_t buffer_num_samples, uint32_t* num_output_samples);

/* ===================================================================== 2. B200 extensions ===== */

/* Text of the last failure inside the CUDA layer (empty string when none). */
const char* SLAB200_LastError(void);

/* Same contract as SLAEncoder_EncodeWhole / SLADecoder_DecodeWhole but every sample plane and the
 * byte stream are DEVICE pointers on the handle's CUDA device - the device that was current when the
 * handle was created; every entry point makes it current again in the calling thread - (input[] /
 * buffer[] themselves are host arrays of device pointers).  The 43-byte file header is written by the
 * host into data[0..43).
 * Stream ordering: a handle works on its own non-blocking streams, which are NOT ordered against the
 * caller's streams (nor against the legacy default stream).  The caller must have synchronised the
 * work that produces d_input / d_data before the call; the call itself is synchronous - outputs are
 * complete on return. */
SLAApiResult SLAB200_Encoder_EncodeWholeDevice(struct SLAEncoder* encoder, const int32_t* const* d_input,
    uint32_t num_samples, uint8_t* d_data, uint32_t data_size, uint32_t* output_size);
SLAApiResult SLAB200_Decoder_DecodeWholeDevice(struct SLADecoder* decoder, const uint8_t* d_data,
    uint32_t data_size, int32_t** d_buffer, uint32_t buffer_num_samples, uint32_t* output_num_samples);

/* Raw-PCM entry points: the samples cross the API as interleaved
 * little-endian PCM in HOST memory - the data chunk of a WAV file: frames of num_channels samples,
 * 8-bit unsigned or 16/24/32-bit signed - instead of planar left-justified int32.  They replace the
 * reference's WAV sample loops (src/wav.c:208-252 with :392-417 in, :630-668 out) together with the
 * whole-file call: only the PCM bytes cross PCIe and the (de)interleave runs on the device.  The
 * stream is byte-identical to what SLAEncoder_EncodeWhole produces from the converted planes.
 * bit_per_sample of the wave format must be 8, 16, 24 or 32. */
SLAApiResult SLAB200_Encoder_EncodePCM(struct SLAEncoder* encoder, const void* pcm, uint32_t num_samples,
    uint8_t* data, uint32_t data_size, uint32_t* output_size);
SLAApiResult SLAB200_Decoder_DecodePCM(struct SLADecoder* decoder, const uint8_t* data, uint32_t data_size,
    void* pcm, uint32_t buffer_num_samples, uint32_t* output_num_samples);

/* Batch decode (BASELINE config 5: a corpus of many short files).  One call decodes every item; files
 * with the same stream parameters are concatenated on the device - one block table, one launch of each
 * kernel for hundreds of files - so short files do not pay a whole-file latency each.  Every item gets its
 * own result code and sample count; the function itself fails only on invalid arguments or a device
 * error.  Output is interleaved little-endian PCM as in SLAB200_Decoder_DecodePCM. */
struct SLAB200BatchItem {
  const uint8_t* data;            /* in:  one .sla stream in host memory */
  uint32_t       data_size;
  void*          pcm;             /* in:  host buffer with room for capacity_samples frames */
  uint32_t       capacity_samples;
  uint32_t       output_num_samples;   /* out */
  SLAApiResult   result;               /* out: what SLADecoder_DecodeWhole would have returned for this file */
};
SLAApiResult SLAB200_Decoder_DecodeBatchPCM(struct SLADecoder* decoder, struct SLAB200BatchItem* items,
    uint32_t num_items);

/* Batch encode: many files of the handle's wave format and encode parameters in one call.  The files are
 * laid back to back in one set of device planes and go through ONE launch sequence per group of up to 48 M
 * frames (segment chain, offset_lshift and header statistics per file; everything else per block, as within a
 * single file), so a corpus of short files runs at the rate of one long file instead of paying every file's
 * launch and chain latency; groups are spread over W internal contexts (SLAB200_BATCH_ENC_WORKERS, default 6)
 * so that the copies of one overlap the kernels of another.  Every stream equals the one
 * SLAB200_Encoder_EncodePCM writes for that file; each item gets its own result code; the function itself fails
 * only on invalid arguments or a device error. */
struct SLAB200EncodeItem {
  const void*  pcm;               /* in:  interleaved little-endian PCM, num_samples frames (host) */
  uint32_t     num_samples;
  uint8_t*     data;              /* in:  host buffer for the .sla stream */
  uint32_t     data_size;
  uint32_t     output_size;       /* out */
  SLAApiResult result;            /* out */
};
SLAApiResult SLAB200_Encoder_EncodeBatchPCM(struct SLAEncoder* encoder, struct SLAB200EncodeItem* items,
    uint32_t num_items);

/* Shard-range encode for multi-GPU runs: encodes the blocks of one contiguous sample range of a
 * longer file (host pointers already offset to the range) with an offset_lshift agreed across
 * shards, writing bare blocks (no file header) to data.  The caller stitches the shards and writes
 * the header from the returned statistics. */
struct SLAB200RangeResult {
  uint32_t num_blocks, total_bytes, max_block_size, max_bit_per_second, input_or_mask;
};
SLAApiResult SLAB200_Encoder_InputOrMask(struct SLAEncoder* encoder, const int32_t* const* input,
    uint32_t num_samples, uint32_t* or_mask);
SLAApiResult SLAB200_Encoder_EncodeRange(struct SLAEncoder* encoder, const int32_t* const* input,
    uint32_t num_samples, uint32_t offset_lshift, uint8_t* data, uint32_t data_size,
    struct SLAB200RangeResult* result);

/* One shard of a file that is split over several GPUs (one process each), chain-exact: the stitched
 * stream is byte-identical to what one encoder writes for the whole file, whatever the leading-silence
 * rule (src/SLAEncoder.c:393-408) does to the segment grid.
 *   input[c]      first sample of this shard's range: its nominal range [begin, end) - boundaries are
 *                 multiples of max_num_block_samples - plus one more block (or up to the end of the file),
 *                 because the shard's last segment may run past `end`; host or device planes
 *   chain_start   where this shard's segment chain starts, relative to `begin`: 0 for the first shard,
 *                 otherwise what the previous shard reported through on_chain (it may lie past soft_end:
 *                 the shard is then empty and passes the value on)
 *   soft_end      end - begin for every shard but the last one (0 there): no segment starts at or after it
 *   offset_lshift agreed across the shards beforehand (min over shards of the trailing-zero count of
 *                 SLAB200_Encoder_InputOrMask, src/SLAEncoder.c:425-455)
 *   on_chain      called as soon as this shard's chain is known - long before its blocks are encoded -
 *                 with the position, relative to `begin`, where the next shard's chain starts; the caller
 *                 sends it to the next rank (4 bytes)
 * The blocks (no file header) are written to data (host or device).  The caller gathers the
 * {num_blocks, total_bytes, max_block_size, max_bit_per_second} of all shards, derives the byte offsets
 * and the header (SLAEncoder_EncodeHeader), and every rank writes its span at its offset. */
struct SLAB200ShardResult {
  uint32_t num_blocks, total_bytes, max_block_size, max_bit_per_second, next_start;
};
typedef void (*SLAB200ChainCallback)(void* user, uint32_t next_start);
SLAApiResult SLAB200_Encoder_EncodeShard(struct SLAEncoder* encoder, const int32_t* const* input, int input_on_device,
    uint32_t num_samples, uint32_t chain_start, uint32_t soft_end, uint32_t offset_lshift,
    uint8_t* data, int data_on_device, uint32_t data_size,
    SLAB200ChainCallback on_chain, void* user, struct SLAB200ShardResult* result);
/* OR mask of device-resident planes (the shard's nominal range) */
SLAApiResult SLAB200_Encoder_InputOrMaskDevice(struct SLAEncoder* encoder, const int32_t* const* d_input,
    uint32_t num_samples, uint32_t* or_mask);
/* device -> host copy on the handle's stream, synchronous (page-locked or pageable destination): how a
 * rank places its span of the stitched stream */
SLAApiResult SLAB200_Encoder_Download(struct SLAEncoder* encoder, void* dst_host, const void* src_device, uint32_t bytes);

/* Debug export: per-block intermediates of the last SLAEncoder_EncodeWhole call on this handle are
 * recorded into `records` (layout identical to oracle/sla_oracle.h OraBlock).  Pass NULL to stop. */
struct SLAB200BlockRecord {
  uint32_t sample_offset, num_samples, block_type, block_size, byte_offset;
  uint32_t rshift[8];
  uint32_t pitch[8];
  int32_t  parcor_code[8][65];
  int32_t  lt_q31[8][8];
  uint64_t rice_init[8];
  double   parcor[8][65];
  double   lt[8][8];
};
void SLAB200_Encoder_SetDebugExport(struct SLAEncoder* encoder, struct SLAB200BlockRecord* records,
    uint32_t max_records, int32_t* const* residual_out);

/* Test hook: the encoder's long-term (pitch) analysis on a caller-supplied residual of one block and one
 * channel (host pointer, <= 16384 samples) - SLALongTermCalculator_CalculateCoef as the encoder handle
 * configures it (src/SLAPredictor.c:791-980, src/SLAEncoder.c:110); lets the reference's pitch KAT
 * (test/test_SLAPredictor.c:717-768) run against the device kernels.  pitch_period = 0 where the reference
 * reports a failure or a period the encoder discards (SLAEncoder.c:629-632). */
SLAApiResult SLAB200_Debug_LongTerm(struct SLAEncoder* encoder, const int32_t* residual, uint32_t num_samples,
    uint32_t num_taps, uint32_t* pitch_period, double* coef);

/* Device time of the last whole-file call on a handle, in milliseconds: [0] host->device,
 * [1] kernels, [2] device->host; and the number of kernel launches it made. */
void SLAB200_Encoder_LastTiming(const struct SLAEncoder* encoder, float ms[3], uint32_t* launches);
void SLAB200_Decoder_LastTiming(const struct SLADecoder* decoder, float ms[3], uint32_t* launches);

/* Last SLAB200_Decoder_DecodeBatchPCM call: device time of the kernels summed over its groups (groups
 * run on several contexts at once, so this is an upper bound of the device-busy time) and launches. */
void SLAB200_Decoder_LastBatchTiming(const struct SLADecoder* decoder, float* kernel_ms, uint32_t* launches);

/* Last single-pass whole-file encode on this handle: block x channels that left the fast path for one of the
 * exactness fallbacks - [0] the reference's FFT autocorrelation for the pitch analysis (listings; a block x
 * channel can be listed twice), [1] PARCOR lag sums in the reference's summation order, [2] scalar instead of
 * tensor-core long-term lag sums (residual >= 2^23). */
void SLAB200_Encoder_LastFallbacks(const struct SLAEncoder* encoder, uint32_t counts[3]);

/* Per-kernel device timing: when enabled, every kernel launch of the following whole-file calls is
 * bracketed by CUDA events on the launching stream; GetProfile returns the launch list in order
 * (kernel label, milliseconds) of the last call and the number of entries. */
void SLAB200_Encoder_EnableProfile(struct SLAEncoder* encoder, int on);
uint32_t SLAB200_Encoder_GetProfile(const struct SLAEncoder* encoder, const char** names, float* ms, uint32_t max_entries);
void SLAB200_Decoder_EnableProfile(struct SLADecoder* decoder, int on);
uint32_t SLAB200_Decoder_GetProfile(const struct SLADecoder* decoder, const char** names, float* ms, uint32_t max_entries);

#ifdef __cplusplus
}
#endif

#endif /* SLA_B200_H_INCLUDED */
