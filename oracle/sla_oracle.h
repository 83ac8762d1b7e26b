/*
 * sla_oracle.h - TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the SLA block encode/decode path (reference: aikiriao/SLA, files cited at
 * each function in sla_oracle.c).  It is the parity checker for the CUDA library and nothing else:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link or call it; the
 * product (libsla_b200.so) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this restatement byte-for-byte and
 * double-for-double against the unmodified reference compiled into oracle/_ref (whole streams,
 * per-block intermediates, decode output) and against the committed vectors in tests/golden/.
 */
#ifndef SLA_ORACLE_H
#define SLA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAX_CH    8
#define ORA_MAX_ORD   64
#define ORA_MAX_TAPS  8

enum { ORA_BLOCK_COMPRESS = 0, ORA_BLOCK_SILENT = 1, ORA_BLOCK_RAW = 2 };

/* Everything that determines a stream besides the samples. */
typedef struct OraParams {
  uint32_t num_channels;
  uint32_t bits_per_sample;
  uint32_t sampling_rate;
  uint32_t parcor_order;
  uint32_t longterm_order;
  uint32_t lms_order;
  uint32_t ch_process;        /* 0 none, 1 stereo mid/side */
  uint32_t window_type;       /* 0 rect, 1 sin, 2 hann, 3 blackman, 4 vorbis */
  uint32_t max_block_samples;
  uint32_t fft_size;          /* reference handle property: roundup_pow2(2 * handle max block) */
} OraParams;

/* Same layout as struct RefWBBlock in ref_whitebox.c (tests share one ctypes definition). */
typedef struct OraBlock {
  uint32_t sample_offset;
  uint32_t num_samples;
  uint32_t block_type;
  uint32_t block_size;
  uint32_t byte_offset;
  uint32_t rshift[ORA_MAX_CH];
  uint32_t pitch[ORA_MAX_CH];
  int32_t  parcor_code[ORA_MAX_CH][ORA_MAX_ORD + 1];
  int32_t  lt_q31[ORA_MAX_CH][ORA_MAX_TAPS];
  uint64_t rice_init[ORA_MAX_CH];
  double   parcor[ORA_MAX_CH][ORA_MAX_ORD + 1];
  double   lt[ORA_MAX_CH][ORA_MAX_TAPS];
} OraBlock;

typedef struct OraHeader {
  uint32_t num_channels, num_samples, sampling_rate, bits_per_sample, offset_lshift;
  uint32_t parcor_order, longterm_order, lms_order, ch_process;
  uint32_t num_blocks, max_block_samples, max_block_size, max_bit_per_second;
} OraHeader;

/* primitives */
uint16_t ora_crc16(const uint8_t* data, uint64_t size);
uint32_t ora_lshift_offset(const int32_t* const* input, uint32_t nch, uint32_t n, uint32_t bps);
void     ora_make_window(uint32_t type, double* w, uint32_t n);
void     ora_autocorr(const double* x, uint32_t n, double* r, uint32_t nlags);
void     ora_parcor_double(const double* x, uint32_t n, double* parcor, uint32_t order);
double   ora_code_length(const double* x, uint32_t n, uint32_t bps, const double* parcor, uint32_t order);
void     ora_realft(double* data, uint32_t n, int sign);
/* returns 0 ok, 1 "failed to calculate" (encoder then disables the long-term stage) */
int      ora_longterm_analyse(const int32_t* res, uint32_t n, uint32_t fft_size, uint32_t taps,
                              uint32_t* pitch, double* coef);
int      ora_dijkstra(const double* adj, uint32_t stride, uint32_t nnodes, uint32_t* path, double* cost);

/* integer filters, in place unless noted */
void ora_ms_forward(int32_t* l, int32_t* r, uint32_t n);
void ora_ms_inverse(int32_t* m, int32_t* s, uint32_t n);
void ora_preemphasis(int32_t* x, uint32_t n);
void ora_deemphasis(int32_t* x, uint32_t n);
void ora_parcor_predict(const int32_t* x, uint32_t n, const int32_t* coef, uint32_t order, int32_t* res);
void ora_parcor_synth(const int32_t* res, uint32_t n, const int32_t* coef, uint32_t order, int32_t* out);
void ora_longterm_filter(const int32_t* in, uint32_t n, uint32_t pitch, const int32_t* coef_q31,
                         uint32_t taps, int32_t* out, int synth);
void ora_lms_filter(const int32_t* in, uint32_t n, uint32_t order, int32_t* out, int synth);

/* partition search for one segment (input = planar left-justified, already offset to the segment) */
int ora_search_partitions(const OraParams* p, const int32_t* const* input, uint32_t seg_samples,
                          uint32_t min_block, uint32_t* nparts, uint32_t* parts);

/* one block; returns bytes written or 0 on overflow */
uint32_t ora_encode_block(const OraParams* p, uint32_t lshift, const int32_t* const* input,
                          uint32_t n, uint8_t* out, uint32_t cap, OraBlock* info,
                          int32_t* const* residual_out);

/* whole file. returns 0 ok, <0 on error (-4 = insufficient buffer). */
int ora_encode_whole(const OraParams* p, const int32_t* const* input, uint32_t num_samples,
                     uint8_t* out, uint32_t cap, uint32_t* out_size,
                     OraBlock* blocks, uint32_t max_blocks, uint32_t* num_blocks,
                     int32_t* const* residual_out);

/* 0 ok, 1 bad signature/version, 2 header crc mismatch (fields still filled), 3 too short */
int ora_decode_header(const uint8_t* data, uint32_t size, OraHeader* h);

/* whole-file decode to planar left-justified int32. returns 0 ok; 10 sync lost; 11 crc; 12 short data;
 * 13 output too small. */
int ora_decode_whole(const uint8_t* data, uint32_t size, int check_crc,
                     int32_t* const* out, uint32_t out_capacity, uint32_t* out_samples,
                     OraBlock* blocks, uint32_t max_blocks, uint32_t* num_blocks);

#ifdef __cplusplus
}
#endif
#endif
