/*
 * ref_whitebox.c - TEST INFRASTRUCTURE ONLY (never linked into libsla_b200.so).
 *
 * White-box view of the UNMODIFIED reference encoder.  The reference's own unit tests reach
 * file-static state by #include-ing the .c file (test/test_SLAEncoder.c:6); this shim does the same
 * so that the parity tests can see what the public API hides:
 *   - the block partition chosen for every segment,
 *   - per block x channel: PARCOR doubles before quantisation, quantised codes, rshift,
 *     pitch period, long-term taps (double and Q31), initial Rice parameter,
 *   - the final residual handed to the entropy coder.
 *
 * Nothing of the reference is copied here: the sources are compiled where they lie
 * (see oracle/Makefile, -I$(REF)/src).  The driver loop below is ours; it calls the reference's
 * static functions SLAEncoder_SearchOptimalBlockPartitions / SLAEncoder_CalculateLeftShiftOffset
 * and the public SLAEncoder_EncodeBlock in the order SLAEncoder_EncodeWhole does
 * (src/SLAEncoder.c:804-932) and snapshots the handle after every block.
 */
#include "SLAUtility.c"
#include "SLABitStream.c"
#include "SLACoder.c"
#include "SLAPredictor.c"
#include "SLAEncoder.c"
#include "SLADecoder.c"

#define WB_MAX_CH   8
#define WB_MAX_ORD  64
#define WB_MAX_TAPS 8

struct RefWBBlock {
  uint32_t sample_offset;
  uint32_t num_samples;
  uint32_t block_type;
  uint32_t block_size;
  uint32_t byte_offset;
  uint32_t rshift[WB_MAX_CH];
  uint32_t pitch[WB_MAX_CH];
  int32_t  parcor_code[WB_MAX_CH][WB_MAX_ORD + 1];
  int32_t  lt_q31[WB_MAX_CH][WB_MAX_TAPS];
  uint64_t rice_init[WB_MAX_CH];
  double   parcor[WB_MAX_CH][WB_MAX_ORD + 1];
  double   lt[WB_MAX_CH][WB_MAX_TAPS];
};

uint32_t RefWB_SizeofBlock(void) { return (uint32_t)sizeof(struct RefWBBlock); }

/* Whole-file encode through the reference's own functions, recording every block.
 * residual_out (optional) receives the coder input per channel, file-length planar. */
int32_t RefWB_EncodeWhole(
    const struct SLAEncoderConfig* config,
    const struct SLAWaveFormat* wave_format,
    const struct SLAEncodeParameter* encode_param,
    const int32_t* const* input, uint32_t num_samples,
    uint8_t* data, uint32_t data_size, uint32_t* output_size,
    struct RefWBBlock* blocks, uint32_t max_blocks, uint32_t* num_blocks_out,
    int32_t* const* residual_out)
{
  struct SLAEncoder* enc;
  struct SLAHeaderInfo hdr;
  const int32_t* ptr[WB_MAX_CH];
  uint32_t pos = 0, nblk = 0, out = SLA_HEADER_SIZE, biggest = 0, peak_bps = 0;
  uint32_t nch, ch, k, p, nparts;
  int32_t rc;

  if ((enc = SLAEncoder_Create(config)) == NULL) { return -1; }
  if ((rc = (int32_t)SLAEncoder_SetWaveFormat(enc, wave_format)) != 0) { goto done; }
  if ((rc = (int32_t)SLAEncoder_SetEncodeParameter(enc, encode_param)) != 0) { goto done; }
  nch = wave_format->num_channels;

  enc->wave_format.offset_lshift
    = (uint8_t)SLAEncoder_CalculateLeftShiftOffset(enc, input, num_samples);

  while (pos < num_samples) {
    uint32_t left = num_samples - pos;
    uint32_t seg  = SLAUTILITY_MIN(enc->encode_param.max_num_block_samples, left);
    uint32_t minb = (uint32_t)SLAUTILITY_MIN(SLA_MIN_BLOCK_NUM_SAMPLES, left);
    if (out >= data_size) { rc = (int32_t)SLA_APIRESULT_INSUFFICIENT_BUFFER_SIZE; goto done; }
    for (ch = 0; ch < nch; ch++) { ptr[ch] = &input[ch][pos]; }
    rc = (int32_t)SLAEncoder_SearchOptimalBlockPartitions(enc, ptr, seg, minb,
        SLA_SEARCH_BLOCK_NUM_SAMPLES_DELTA, seg, &nparts, enc->num_block_partition_samples);
    if (rc != 0) { goto done; }
    for (p = 0; p < nparts; p++) {
      uint32_t n = enc->num_block_partition_samples[p], bsize = 0, bps;
      for (ch = 0; ch < nch; ch++) { ptr[ch] = &input[ch][pos]; }
      rc = (int32_t)SLAEncoder_EncodeBlock(enc, ptr, n, &data[out], data_size - out, &bsize);
      if (rc != 0) { goto done; }
      if (blocks != NULL && nblk < max_blocks) {
        struct RefWBBlock* b = &blocks[nblk];
        memset(b, 0, sizeof(*b));
        b->sample_offset = pos; b->num_samples = n;
        b->block_type = (uint32_t)enc->block_data_type;
        b->block_size = bsize;  b->byte_offset = out;
        for (ch = 0; ch < nch; ch++) {
          b->rshift[ch] = enc->parcor_rshift[ch];
          b->pitch[ch]  = enc->pitch_period[ch];
          b->rice_init[ch] = enc->coder->init_rice_parameter[ch][0];
          for (k = 0; k <= enc->encode_param.parcor_order && k <= WB_MAX_ORD; k++) {
            b->parcor[ch][k] = enc->parcor_coef[ch][k];
            b->parcor_code[ch][k] = enc->parcor_coef_code[ch][k];
          }
          for (k = 0; k < enc->encode_param.longterm_order && k < WB_MAX_TAPS; k++) {
            b->lt[ch][k] = enc->longterm_coef[ch][k];
            b->lt_q31[ch][k] = enc->longterm_coef_int32[ch][k];
          }
        }
      }
      if (residual_out != NULL && enc->block_data_type == SLA_BLOCK_DATA_TYPE_COMPRESSDATA) {
        for (ch = 0; ch < nch; ch++) {
          memcpy(&residual_out[ch][pos], enc->residual[ch], sizeof(int32_t) * n);
        }
      }
      out += bsize; pos += n; nblk++;
      if (bsize > biggest) { biggest = bsize; }
      bps = (8 * bsize * enc->wave_format.sampling_rate) / n;
      if (bps > peak_bps) { peak_bps = bps; }
    }
  }

  hdr.wave_format = enc->wave_format;
  hdr.encode_param = enc->encode_param;
  hdr.num_samples = num_samples;
  hdr.num_blocks = nblk;
  hdr.max_block_size = biggest;
  hdr.max_bit_per_second = peak_bps;
  rc = (int32_t)SLAEncoder_EncodeHeader(&hdr, data, data_size);
  *output_size = out;
  if (num_blocks_out != NULL) { *num_blocks_out = nblk; }

done:
  SLAEncoder_Destroy(enc);
  return rc;
}

/* Thin probes into reference statics, used to pin individual oracle functions. */
void RefWB_AutoCorrelation(const double* data, uint32_t n, double* out, uint32_t num_lags)
{
  (void)LPC_CalculateAutoCorrelation(data, n, out, num_lags);
}

int32_t RefWB_Parcor(const double* data, uint32_t n, double* parcor, uint32_t order)
{
  struct SLALPCCalculator* c = SLALPCCalculator_Create(order);
  int32_t rc = (int32_t)SLALPCCalculator_CalculatePARCORCoefDouble(c, data, n, parcor, order);
  SLALPCCalculator_Destroy(c);
  return rc;
}

int32_t RefWB_LongTerm(const int32_t* data, uint32_t n, uint32_t fft_size, uint32_t taps,
    uint32_t* pitch, double* coef)
{
  struct SLALongTermCalculator* c = SLALongTermCalculator_Create(fft_size,
      SLALONGTERM_MAX_PERIOD, SLALONGTERM_NUM_PITCH_CANDIDATES, taps);
  int32_t rc = (int32_t)SLALongTermCalculator_CalculateCoef(c, data, n, pitch, coef, taps);
  SLALongTermCalculator_Destroy(c);
  return rc;
}
