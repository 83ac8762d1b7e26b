/*
 * sla_b200_cli.c - command-line front end of libsla_b200.
 *
 * Same user interface as the reference tool (src/main.c:28-70 options, :420-537 main): -e / -d,
 * -m PRESET, -p / -q, -c yes|no, -h, -v, INPUT OUTPUT, same presets (src/main.c:57-64), same
 * handle capacities (:92-98, :186-193), same messages and exit codes.  What differs is the data path:
 * the reference parses the WAV file sample by sample through a bit reader into planar int32
 * (src/wav.c:208-252) and writes it back the same way (:630-668); here the data chunk of the WAV file
 * goes to the GPU as it lies in the file (SLAB200_Encoder_EncodePCM) and the decoder's interleaved PCM
 * lands directly behind a 44-byte RIFF header (SLAB200_Decoder_DecodePCM).  The files produced are
 * byte-identical to the reference tool's.
 *
 * One addition: -b DIR (batch).  Every positional argument is then an input file and the outputs are
 * DIR/<input stem>.wav or .sla; decoding goes through SLAB200_Decoder_DecodeBatchPCM, which decodes
 * files of equal stream parameters together.
 *
 * -s (the reference's streaming decode) runs SLAStreamingDecoder_* exactly as src/main.c:275-420 does.
 */
#define _POSIX_C_SOURCE 200809L
#include "sla_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

/* ------------------------------------------------------------------ options ---- */
struct Option {
  char        short_name;
  const char* long_name;
  int         takes_value;
  const char* help;
  int         seen;
  const char* value;
};

static struct Option g_options[] = {
  { 'e', "encode",    0, "Encode mode", 0, NULL },
  { 'd', "decode",    0, "Decode mode", 0, NULL },
  { 'm', "mode",      1, "Specify compress mode: 0(fast decode), ..., 4(high compression) default:2", 0, NULL },
  { 'p', "verpose",   0, "Verpose mode(try to display all information)", 0, NULL },
  { 'q', "quiet",     0, "Quiet mode(suppress outputs)", 0, NULL },
  { 'c', "crc-check", 1, "Whether to check CRC16 at decoding(yes or no) default:yes", 0, NULL },
  { 'h', "help",      0, "Show command help message", 0, NULL },
  { 'v', "version",   0, "Show version information", 0, NULL },
  { 's', "streaming", 0, "Use streaming decode(for debug; 120fps)", 0, NULL },
  { 'b', "batch",     1, "Batch mode: every other argument is an input file, outputs go to this directory", 0, NULL },
};
#define NUM_OPTIONS ((int)(sizeof(g_options) / sizeof(g_options[0])))
#define MAX_FILES 65536

static struct Option* option(const char* long_name)
{
  int i;
  for (i = 0; i < NUM_OPTIONS; i++)
    if (strcmp(g_options[i].long_name, long_name) == 0) return &g_options[i];
  return NULL;
}

/* value of an option that needs one: the next argument, which must not look like an option */
static int take_value(struct Option* o, int argc, char** argv, int* at, const char* shown)
{
  if (*at + 1 >= argc || argv[*at + 1][0] == '-') {
    fprintf(stderr, "%s: Option %s needs argument. \n", argv[0], shown);
    return -1;
  }
  o->value = argv[++*at];
  return 0;
}

/* grammar of src/command_line_parser.c:150-290: "--name", "--name=value", "--name value", bundles of
 * short options "-eq" where one that takes a value must come last, everything else is a file name */
static int parse_arguments(int argc, char** argv, const char** files, int max_files, int* num_files)
{
  int at, i;
  char shown[64];
  *num_files = 0;
  for (at = 1; at < argc; at++) {
    const char* arg = argv[at];
    if (strncmp(arg, "--", 2) == 0) {
      const char* name = arg + 2;
      const char* eq = strchr(name, '=');
      size_t len = eq ? (size_t)(eq - name) : strlen(name);
      struct Option* o = NULL;
      for (i = 0; i < NUM_OPTIONS; i++)
        if (strlen(g_options[i].long_name) == len && strncmp(g_options[i].long_name, name, len) == 0
            && (!eq || g_options[i].takes_value)) { o = &g_options[i]; break; }
      if (!o) { fprintf(stderr, "%s: Unknown long option - \"%s\" \n", argv[0], name); return -1; }
      if (o->seen) { fprintf(stderr, "%s: Option \"%s\" multiply specified. \n", argv[0], o->long_name); return -1; }
      if (eq) o->value = eq + 1;
      else if (o->takes_value) {
        snprintf(shown, sizeof(shown), "\"%s\"", o->long_name);
        if (take_value(o, argc, argv, &at, shown) != 0) return -1;
      }
      o->seen = 1;
    } else if (arg[0] == '-') {
      const char* p;
      for (p = arg + 1; *p; p++) {
        struct Option* o = NULL;
        for (i = 0; i < NUM_OPTIONS; i++) if (g_options[i].short_name == *p) { o = &g_options[i]; break; }
        if (!o) { fprintf(stderr, "%s: Unknown short option - \'%c\' \n", argv[0], *p); return -1; }
        if (o->seen) { fprintf(stderr, "%s: Option \'%c\' multiply specified. \n", argv[0], *p); return -1; }
        if (o->takes_value) {
          if (p[1] != '\0') {
            fprintf(stderr, "%s: Option \'%c\' needs argument. Please specify tail of short option sequence.\n", argv[0], *p);
            return -1;
          }
          snprintf(shown, sizeof(shown), "\'%c\'", *p);
          if (take_value(o, argc, argv, &at, shown) != 0) return -1;
        }
        o->seen = 1;
      }
    } else {
      if (*num_files >= max_files) { fprintf(stderr, "%s: Too many strings specified. \n", argv[0]); return -1; }
      files[(*num_files)++] = arg;
    }
  }
  return 0;
}

static void print_usage(char** argv) { printf("Usage: %s [options] INPUT_FILE_NAME OUTPUT_FILE_NAME \n", argv[0]); }

static void print_options(void)
{
  int i;
  char name[64];
  for (i = 0; i < NUM_OPTIONS; i++) {
    snprintf(name, sizeof(name), "  -%c, --%s", g_options[i].short_name, g_options[i].long_name);
    printf("%-20s %-18s  %s \n", name, g_options[i].takes_value ? "(needs argument)" : "", g_options[i].help);
  }
}

/* ------------------------------------------------------------------ presets (src/main.c:57-64) ---- */
static const struct SLAEncodeParameter g_presets[] = {
  {  8, 1, 4, SLA_CHPROCESSMETHOD_NONE,      SLA_WINDOWFUNCTIONTYPE_RECTANGULAR,  4096 },
  {  8, 1, 8, SLA_CHPROCESSMETHOD_STEREO_MS, SLA_WINDOWFUNCTIONTYPE_SIN,         12288 },
  { 16, 1, 8, SLA_CHPROCESSMETHOD_STEREO_MS, SLA_WINDOWFUNCTIONTYPE_SIN,         12288 },
  { 32, 3, 8, SLA_CHPROCESSMETHOD_STEREO_MS, SLA_WINDOWFUNCTIONTYPE_SIN,         12288 },
  { 32, 3, 8, SLA_CHPROCESSMETHOD_STEREO_MS, SLA_WINDOWFUNCTIONTYPE_SIN,         16384 },
};
#define NUM_PRESETS ((uint32_t)(sizeof(g_presets) / sizeof(g_presets[0])))
#define DEFAULT_PRESET 2u

/* ------------------------------------------------------------------ files ---- */
static uint8_t* read_file(const char* name, size_t* size)
{
  FILE* fp = fopen(name, "rb");
  struct stat st;
  uint8_t* buf;
  if (!fp) return NULL;
  if (fstat(fileno(fp), &st) != 0 || st.st_size < 0) { fclose(fp); return NULL; }
  buf = (uint8_t*)malloc((size_t)st.st_size + 1u);
  if (!buf) { fclose(fp); return NULL; }
  *size = fread(buf, 1, (size_t)st.st_size, fp);
  fclose(fp);
  if (*size != (size_t)st.st_size) { free(buf); return NULL; }
  return buf;
}

static int write_file(const char* name, const uint8_t* a, size_t na, const uint8_t* b, size_t nb)
{
  FILE* fp = fopen(name, "wb");
  int ok;
  if (!fp) return -1;
  ok = (na == 0 || fwrite(a, 1, na, fp) == na) && (nb == 0 || fwrite(b, 1, nb, fp) == nb);
  if (fclose(fp) != 0) ok = 0;
  return ok ? 0 : -1;
}

static uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void put16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

struct WavInfo {
  uint32_t num_channels, sampling_rate, bits_per_sample, num_samples;
  const uint8_t* pcm;            /* the data chunk's payload, inside the file image */
};

/* Accepts what src/wav.c:107-205 accepts: "RIFF" size "WAVE", the "fmt " chunk first (format id 1, an
 * extension beyond 16 bytes skipped with the same warning), then any chunks up to "data" (sizes taken
 * as they are, without RIFF's pad byte, as the reference does); 8/16/24/32 bits (:220-236). */
static int parse_wav(const uint8_t* f, size_t size, struct WavInfo* w)
{
  size_t at = 12;
  uint32_t fmt_size, data_size, frame;
  if (size < 36 || memcmp(f, "RIFF", 4) != 0 || memcmp(f + 8, "WAVE", 4) != 0 || memcmp(f + 12, "fmt ", 4) != 0) return -1;
  fmt_size = le32(f + 16);
  if (le16(f + 20) != 1u) return -1;
  w->num_channels = le16(f + 22);
  w->sampling_rate = le32(f + 24);
  w->bits_per_sample = le16(f + 34);
  at = 36;
  if ((int32_t)fmt_size > 16) {
    fprintf(stderr, "Warning: skip fmt chunk extention (unsupported). \n");
    at += fmt_size - 16u;
  }
  for (;;) {
    if (at + 8 > size) return -1;
    if (memcmp(f + at, "data", 4) == 0) break;
    at += 8u + (size_t)le32(f + at + 4);
  }
  data_size = le32(f + at + 4);
  at += 8;
  if (w->bits_per_sample != 8 && w->bits_per_sample != 16 && w->bits_per_sample != 24 && w->bits_per_sample != 32) return -1;
  frame = (w->bits_per_sample / 8u) * w->num_channels;
  if (frame == 0) return -1;
  w->num_samples = data_size / frame;
  if ((size_t)w->num_samples * frame > size - at) return -1;       /* the reference runs into end of file here */
  w->pcm = f + at;
  return 0;
}

/* the 44 bytes of src/wav.c:545-627 */
static void make_wav_header(uint8_t* h, uint32_t nch, uint32_t rate, uint32_t bits, uint32_t num_samples)
{
  const uint32_t frame = (bits / 8u) * nch, pcm_size = num_samples * frame;
  memcpy(h, "RIFF", 4); put32(h + 4, pcm_size + 44u - 8u);
  memcpy(h + 8, "WAVEfmt ", 8); put32(h + 16, 16); put16(h + 20, 1); put16(h + 22, nch);
  put32(h + 24, rate); put32(h + 28, rate * frame); put16(h + 32, frame); put16(h + 34, bits);
  memcpy(h + 36, "data", 4); put32(h + 40, pcm_size);
}

/* ------------------------------------------------------------------ encode (src/main.c:72-166) ---- */
static struct SLAEncoder* make_encoder(uint8_t verbose)
{
  struct SLAEncoderConfig config;
  struct SLAEncoder* e;
  config.max_num_channels = 8; config.max_num_block_samples = 16384; config.max_parcor_order = 48;
  config.max_longterm_order = 5; config.max_lms_order_per_filter = 40; config.verpose_flag = verbose;
  if ((e = SLAEncoder_Create(&config)) == NULL)
    fprintf(stderr, "Failed to create encoder handle. %s\n", SLAB200_LastError());
  return e;
}

static int encode_file(struct SLAEncoder* encoder, const char* in_name, const char* out_name, uint32_t preset_no, uint8_t verbose)
{
  size_t file_size = 0;
  uint8_t* file = read_file(in_name, &file_size);
  struct WavInfo wav;
  struct SLAWaveFormat wf;
  struct SLAEncodeParameter ep;
  uint8_t* out;
  uint32_t out_capacity, out_size = 0;
  SLAApiResult ret;
  if (!file || parse_wav(file, file_size, &wav) != 0) {
    fprintf(stderr, "Failed to open %s \n", in_name);
    free(file);
    return 1;
  }
  memset(&wf, 0, sizeof(wf));
  wf.num_channels = wav.num_channels; wf.bit_per_sample = wav.bits_per_sample; wf.sampling_rate = wav.sampling_rate;
  if ((ret = SLAEncoder_SetWaveFormat(encoder, &wf)) != SLA_APIRESULT_OK) {
    fprintf(stderr, "Failed to set wave parameter: %d \n", ret);
    free(file);
    return 1;
  }
  ep = g_presets[preset_no];
  if (!(wav.num_channels == 2 && ep.ch_process_method == SLA_CHPROCESSMETHOD_STEREO_MS))
    ep.ch_process_method = SLA_CHPROCESSMETHOD_NONE;            /* MS only for stereo sources (src/main.c:125-131) */
  if ((ret = SLAEncoder_SetEncodeParameter(encoder, &ep)) != SLA_APIRESULT_OK) {
    fprintf(stderr, "Failed to set encode parameter: %d \n", ret);
    free(file);
    return 1;
  }
  out_capacity = (uint32_t)(2u * file_size);                    /* src/main.c:139-142 */
  out = (uint8_t*)malloc(out_capacity ? out_capacity : 1u);
  if (!out) { fprintf(stderr, "Encoding error! out of memory \n"); free(file); return 1; }
  ret = SLAB200_Encoder_EncodePCM(encoder, wav.pcm, wav.num_samples, out, out_capacity, &out_size);
  if (ret != SLA_APIRESULT_OK) {
    fprintf(stderr, "Encoding error! %d \n", ret);
    free(out); free(file);
    return 1;
  }
  if (write_file(out_name, out, out_size, NULL, 0) != 0) {
    fprintf(stderr, "Failed to write %s \n", out_name);
    free(out); free(file);
    return 1;
  }
  if (verbose) printf("Encode succuess! size:%d -> %d \n", (uint32_t)file_size, out_size);
  free(out); free(file);
  return 0;
}

/* ------------------------------------------------------------------ decode (src/main.c:168-272) ---- */
static void decoder_config(struct SLADecoderConfig* config, uint8_t crc, uint8_t verbose)
{
  config->max_num_channels = 8; config->max_num_block_samples = 16384; config->max_parcor_order = 48;
  config->max_longterm_order = 5; config->max_lms_order_per_filter = 40;
  config->enable_crc_check = crc; config->verpose_flag = verbose;
}

static void print_header(const struct SLAHeaderInfo* h)
{
  printf("Num Channels:                %d \n", h->wave_format.num_channels);
  printf("Bit Per Sample:              %d \n", h->wave_format.bit_per_sample);
  printf("Sampling Rate:               %d \n", h->wave_format.sampling_rate);
  printf("Offset Left Shift:           %d \n", h->wave_format.offset_lshift);
  printf("PARCOR Order:                %d \n", h->encode_param.parcor_order);
  printf("Longterm Order:              %d \n", h->encode_param.longterm_order);
  printf("LMS Order Par Filter:        %d \n", h->encode_param.lms_order_per_filter);
  printf("Channel Process Method:      %d \n", h->encode_param.ch_process_method);
  printf("Max Number of Block Samples: %d \n", h->encode_param.max_num_block_samples);
  printf("Number of Samples:           %d \n", h->num_samples);
  printf("Number of Blocks:            %d \n", h->num_blocks);
  printf("Max Block Size:              %d \n", h->max_block_size);
  printf("Max Bit Per Second(bps):     %d \n", h->max_bit_per_second);
}

/* a WAV image of header.num_samples frames of digital silence behind its 44-byte header (what the
 * reference's calloc'ed planes become when written, src/wav.c:347-389) */
static uint8_t* new_wav_image(const struct SLAHeaderInfo* h, size_t* pcm_size)
{
  const uint32_t bits = h->wave_format.bit_per_sample;
  uint8_t* img;
  if (bits != 8 && bits != 16 && bits != 24 && bits != 32) return NULL;
  *pcm_size = (size_t)h->num_samples * (bits / 8u) * h->wave_format.num_channels;
  if ((img = (uint8_t*)malloc(44u + *pcm_size + 1u)) == NULL) return NULL;
  make_wav_header(img, h->wave_format.num_channels, h->wave_format.sampling_rate, bits, h->num_samples);
  memset(img + 44, bits == 8 ? 0x80 : 0, *pcm_size);
  return img;
}

static int decode_file(const char* in_name, const char* out_name, uint8_t crc, uint8_t verbose)
{
  struct SLADecoderConfig config;
  struct SLADecoder* decoder;
  struct SLAHeaderInfo header;
  size_t size = 0, pcm_size = 0;
  uint8_t *data, *img;
  uint32_t decoded = 0;
  SLAApiResult ret;
  decoder_config(&config, crc, verbose);
  if ((decoder = SLADecoder_Create(&config)) == NULL) {
    fprintf(stderr, "Failed to create decoder handle. %s\n", SLAB200_LastError());
    return 1;
  }
  if ((data = read_file(in_name, &size)) == NULL) {
    fprintf(stderr, "Failed to open %s \n", in_name);
    SLADecoder_Destroy(decoder);
    return 1;
  }
  if ((ret = SLADecoder_DecodeHeader(data, (uint32_t)size, &header)) != SLA_APIRESULT_OK) {
    fprintf(stderr, "Failed to get header information: %d \n", ret);
    goto fail;
  }
  if (verbose) print_header(&header);
  if ((img = new_wav_image(&header, &pcm_size)) == NULL) {
    fprintf(stderr, "Failed to create wav handle. \n");
    goto fail;
  }
  if ((ret = SLADecoder_SetWaveFormat(decoder, &header.wave_format)) != SLA_APIRESULT_OK) {
    fprintf(stderr, "Failed to set wave parameter: %d \n", ret);
    free(img); goto fail;
  }
  if ((ret = SLADecoder_SetEncodeParameter(decoder, &header.encode_param)) != SLA_APIRESULT_OK) {
    fprintf(stderr, "Failed to set encode parameter: %d \n", ret);
    free(img); goto fail;
  }
  if ((ret = SLAB200_Decoder_DecodePCM(decoder, data, (uint32_t)size, img + 44, header.num_samples, &decoded)) != SLA_APIRESULT_OK) {
    fprintf(stderr, "Decoding error! %d \n", ret);
    free(img); goto fail;
  }
  if (write_file(out_name, img, 44u + pcm_size, NULL, 0) != 0) {
    fprintf(stderr, "Failed to write wav file. \n");
    free(img); goto fail;
  }
  free(img); free(data);
  SLADecoder_Destroy(decoder);
  return 0;
fail:
  free(data);
  SLADecoder_Destroy(decoder);
  return 1;
}

/* ------------------------------------------------------------------ streaming decode (src/main.c:275-420) ---- */
static int streaming_decode_file(const char* in_name, const char* out_name, uint8_t crc, uint8_t verbose)
{
  struct SLAStreamingDecoderConfig sc;
  struct SLAStreamingDecoder* decoder;
  struct SLAHeaderInfo header;
  size_t size = 0, pcm_size = 0;
  uint8_t *data, *img = NULL;
  int32_t* planes[8];
  uint32_t nch, ch, bytes, sample_progress = 0, data_progress = SLA_HEADER_SIZE, per_decode = 0, i;
  SLAApiResult ret;
  int rc = 1;
  memset(planes, 0, sizeof(planes));
  decoder_config(&sc.core_config, crc, verbose);
  sc.decode_interval_hz = 120.0f;
  sc.max_bit_per_sample = 24;
  if ((decoder = SLAStreamingDecoder_Create(&sc)) == NULL) {
    fprintf(stderr, "Failed to create decoder handle. %s\n", SLAB200_LastError());
    return 1;
  }
  if ((data = read_file(in_name, &size)) == NULL) {
    fprintf(stderr, "Failed to open %s \n", in_name);
    SLAStreamingDecoder_Destroy(decoder);
    return 1;
  }
  if ((ret = SLADecoder_DecodeHeader(data, (uint32_t)size, &header)) != SLA_APIRESULT_OK) {
    fprintf(stderr, "Failed to get header information: %d \n", ret);
    goto done;
  }
  if (verbose) print_header(&header);
  if ((img = new_wav_image(&header, &pcm_size)) == NULL) { fprintf(stderr, "Failed to create wav handle. \n"); goto done; }
  if ((ret = SLAStreamingDecoder_SetWaveFormat(decoder, &header.wave_format)) != SLA_APIRESULT_OK) {
    fprintf(stderr, "Failed to set wave parameter: %d \n", ret);
    goto done;
  }
  if ((ret = SLAStreamingDecoder_SetEncodeParameter(decoder, &header.encode_param)) != SLA_APIRESULT_OK) {
    fprintf(stderr, "Failed to set encode parameter: %d \n", ret);
    goto done;
  }
  if (SLAStreamingDecoder_GetOutputNumSamplesPerDecode(decoder, &per_decode) != SLA_APIRESULT_OK || per_decode == 0) {
    fprintf(stderr, "Failed to get number of samples per decode. \n");
    goto done;
  }
  nch = header.wave_format.num_channels; bytes = header.wave_format.bit_per_sample / 8u;
  for (ch = 0; ch < nch; ch++)
    if ((planes[ch] = (int32_t*)malloc(sizeof(int32_t) * per_decode)) == NULL) goto done;
  /* the feeding pattern of src/main.c:368-405: the first fragment is one maximum block, afterwards what
   * the decoder estimates it needs for one Decode call; one Decode per fragment */
  while (sample_progress < header.num_samples) {
    uint32_t give = 0, got = 0;
    const uint8_t* used; uint32_t used_size;
    if (sample_progress == 0 && data_progress == SLA_HEADER_SIZE) give = header.max_block_size;
    else if (SLAStreamingDecoder_EstimateMinimumNessesaryDataSize(decoder, &give) != SLA_APIRESULT_OK) goto done;
    if (give > size - data_progress) give = (uint32_t)(size - data_progress);
    if ((ret = SLAStreamingDecoder_AppendDataFragment(decoder, data + data_progress, give)) != SLA_APIRESULT_OK) {
      fprintf(stderr, "Failed to append data fragment: %d \n", ret);
      goto done;
    }
    data_progress += give;
    if ((ret = SLAStreamingDecoder_Decode(decoder, planes, per_decode, &got)) != SLA_APIRESULT_OK) {
      fprintf(stderr, "Streaming Decode failed! ret:%d \n", ret);
      goto done;
    }
    if (got > header.num_samples - sample_progress) got = header.num_samples - sample_progress;
    for (i = 0; i < got; i++)
      for (ch = 0; ch < nch; ch++) {
        uint8_t* p = img + 44 + ((size_t)(sample_progress + i) * nch + ch) * bytes;
        const uint32_t u = (uint32_t)planes[ch][i];
        switch (bytes) {                                   /* src/wav.c:420-437 */
          case 1: p[0] = (uint8_t)((u >> 24) + 128u); break;
          case 2: p[0] = (uint8_t)(u >> 16); p[1] = (uint8_t)(u >> 24); break;
          case 3: p[0] = (uint8_t)(u >> 8); p[1] = (uint8_t)(u >> 16); p[2] = (uint8_t)(u >> 24); break;
          default: put32(p, u); break;
        }
      }
    sample_progress += got;
    while (SLAStreamingDecoder_CollectDataFragment(decoder, &used, &used_size) == SLA_APIRESULT_OK) { }
    if (verbose) { printf("progress: %4.1f %% \r", (double)sample_progress / header.num_samples * 100.0); fflush(stdout); }
  }
  if (write_file(out_name, img, 44u + pcm_size, NULL, 0) != 0) { fprintf(stderr, "Failed to write wav file. \n"); goto done; }
  rc = 0;
done:
  for (ch = 0; ch < 8; ch++) free(planes[ch]);
  free(img); free(data);
  SLAStreamingDecoder_Destroy(decoder);
  return rc;
}

/* ------------------------------------------------------------------ batch ---- */
static char* output_name(const char* dir, const char* in_name, const char* ext)
{
  const char* base = strrchr(in_name, '/');
  const char* dot;
  size_t stem;
  char* out;
  base = base ? base + 1 : in_name;
  dot = strrchr(base, '.');
  stem = dot ? (size_t)(dot - base) : strlen(base);
  out = (char*)malloc(strlen(dir) + 1u + stem + strlen(ext) + 1u);
  if (out) sprintf(out, "%s/%.*s%s", dir, (int)stem, base, ext);
  return out;
}

/* files of one wave format go through one SLAB200_Encoder_EncodeBatchPCM call (runs of equal format) */
static int batch_encode(const char* dir, const char** files, int n, uint32_t preset_no, uint8_t verbose)
{
  struct SLAEncoder* encoder = make_encoder(0);
  uint8_t** file = (uint8_t**)calloc((size_t)n, sizeof(*file));
  struct WavInfo* wav = (struct WavInfo*)calloc((size_t)n, sizeof(*wav));
  struct SLAB200EncodeItem* items = (struct SLAB200EncodeItem*)calloc((size_t)n, sizeof(*items));
  int i, j, failed = 0;
  if (!encoder || !file || !wav || !items) return 1;
  for (i = 0; i < n; i++) {
    size_t size = 0;
    file[i] = read_file(files[i], &size);
    if (!file[i] || parse_wav(file[i], size, &wav[i]) != 0) {
      fprintf(stderr, "Failed to open %s \n", files[i]);
      free(file[i]); file[i] = NULL;
      continue;
    }
    items[i].pcm = wav[i].pcm; items[i].num_samples = wav[i].num_samples;
    items[i].data_size = (uint32_t)(2u * size);                 /* src/main.c:139-142 */
    items[i].data = (uint8_t*)malloc(items[i].data_size ? items[i].data_size : 1u);
    if (!items[i].data) { free(file[i]); file[i] = NULL; }
  }
  for (i = 0; i < n; i = j) {
    struct SLAWaveFormat wf;
    struct SLAEncodeParameter ep = g_presets[preset_no];
    SLAApiResult ret;
    j = i + 1;
    if (!file[i]) { failed++; continue; }
    while (j < n && file[j] && wav[j].num_channels == wav[i].num_channels && wav[j].bits_per_sample == wav[i].bits_per_sample
           && wav[j].sampling_rate == wav[i].sampling_rate) j++;
    memset(&wf, 0, sizeof(wf));
    wf.num_channels = wav[i].num_channels; wf.bit_per_sample = wav[i].bits_per_sample; wf.sampling_rate = wav[i].sampling_rate;
    if (!(wf.num_channels == 2 && ep.ch_process_method == SLA_CHPROCESSMETHOD_STEREO_MS)) ep.ch_process_method = SLA_CHPROCESSMETHOD_NONE;
    if ((ret = SLAEncoder_SetWaveFormat(encoder, &wf)) != SLA_APIRESULT_OK) {
      fprintf(stderr, "%s: Failed to set wave parameter: %d \n", files[i], ret);
      failed += j - i;
      continue;
    }
    if ((ret = SLAEncoder_SetEncodeParameter(encoder, &ep)) != SLA_APIRESULT_OK) {
      fprintf(stderr, "%s: Failed to set encode parameter: %d \n", files[i], ret);
      failed += j - i;
      continue;
    }
    if ((ret = SLAB200_Encoder_EncodeBatchPCM(encoder, items + i, (uint32_t)(j - i))) != SLA_APIRESULT_OK) {
      fprintf(stderr, "Encoding error! %d %s\n", ret, SLAB200_LastError());
      failed += j - i;
      continue;
    }
    {
      int k;
      for (k = i; k < j; k++) {
        char* out = output_name(dir, files[k], ".sla");
        if (items[k].result != SLA_APIRESULT_OK) {
          fprintf(stderr, "%s: Encoding error! %d \n", files[k], items[k].result);
          failed++;
        } else if (!out || write_file(out, items[k].data, items[k].output_size, NULL, 0) != 0) {
          fprintf(stderr, "Failed to write %s \n", out ? out : files[k]);
          failed++;
        }
        free(out);
      }
    }
  }
  for (i = 0; i < n; i++) { free(items[i].data); free(file[i]); }
  free(items); free(wav); free(file);
  SLAEncoder_Destroy(encoder);
  if (verbose) printf("Batch encode: %d of %d files encoded \n", n - failed, n);
  return failed ? 1 : 0;
}

static int batch_decode(const char* dir, const char** files, int n, uint8_t crc, uint8_t verbose)
{
  struct SLADecoderConfig config;
  struct SLADecoder* decoder;
  struct SLAB200BatchItem* items = (struct SLAB200BatchItem*)calloc((size_t)n, sizeof(*items));
  uint8_t** data = (uint8_t**)calloc((size_t)n, sizeof(*data));
  uint8_t** img = (uint8_t**)calloc((size_t)n, sizeof(*img));
  size_t* pcm_size = (size_t*)calloc((size_t)n, sizeof(*pcm_size));
  int i, failed = 0;
  SLAApiResult ret;
  decoder_config(&config, crc, 0);
  decoder = SLADecoder_Create(&config);
  if (!decoder || !items || !data || !img || !pcm_size) {
    fprintf(stderr, "Failed to create decoder handle. %s\n", SLAB200_LastError());
    return 1;
  }
  for (i = 0; i < n; i++) {
    struct SLAHeaderInfo header;
    size_t size = 0;
    items[i].result = SLA_APIRESULT_NG;
    if ((data[i] = read_file(files[i], &size)) == NULL) { fprintf(stderr, "Failed to open %s \n", files[i]); continue; }
    if ((ret = SLADecoder_DecodeHeader(data[i], (uint32_t)size, &header)) != SLA_APIRESULT_OK) {
      fprintf(stderr, "%s: Failed to get header information: %d \n", files[i], ret);
      free(data[i]); data[i] = NULL;
      continue;
    }
    if ((img[i] = new_wav_image(&header, &pcm_size[i])) == NULL) {
      fprintf(stderr, "%s: Failed to create wav handle. \n", files[i]);
      free(data[i]); data[i] = NULL;
      continue;
    }
    items[i].data = data[i]; items[i].data_size = (uint32_t)size;
    items[i].pcm = img[i] + 44; items[i].capacity_samples = header.num_samples;
  }
  /* items that did not load keep data == NULL: compact the ones that did */
  {
    struct SLAB200BatchItem* live = (struct SLAB200BatchItem*)calloc((size_t)n, sizeof(*live));
    int* where = (int*)calloc((size_t)n, sizeof(int));
    int m = 0;
    if (!live || !where) return 1;
    for (i = 0; i < n; i++) if (data[i]) { live[m] = items[i]; where[m++] = i; }
    if (m > 0 && (ret = SLAB200_Decoder_DecodeBatchPCM(decoder, live, (uint32_t)m)) != SLA_APIRESULT_OK) {
      fprintf(stderr, "Decoding error! %d %s\n", ret, SLAB200_LastError());
      return 1;
    }
    for (i = 0; i < m; i++) items[where[i]] = live[i];
    free(live); free(where);
  }
  for (i = 0; i < n; i++) {
    char* out;
    if (!data[i]) { failed++; continue; }
    if (items[i].result != SLA_APIRESULT_OK) {
      fprintf(stderr, "%s: Decoding error! %d \n", files[i], items[i].result);
      failed++;
    } else if ((out = output_name(dir, files[i], ".wav")) == NULL || write_file(out, img[i], 44u + pcm_size[i], NULL, 0) != 0) {
      fprintf(stderr, "%s: Failed to write wav file. \n", files[i]);
      failed++;
      free(out);
    } else free(out);
    free(img[i]); free(data[i]);
  }
  free(items); free(data); free(img); free(pcm_size);
  SLADecoder_Destroy(decoder);
  if (verbose) printf("Batch decode: %d of %d files decoded \n", n - failed, n);
  return failed ? 1 : 0;
}

/* ------------------------------------------------------------------ main (src/main.c:434-537) ---- */
int main(int argc, char** argv)
{
  static const char* files[MAX_FILES];
  int num_files = 0;
  uint8_t verbose = 1;
  const int is_batch_possible = 1;
  if (argc == 1) { print_usage(argv); return 1; }
  if (parse_arguments(argc, argv, files, MAX_FILES, &num_files) != 0) return 1;
  if (option("help")->seen) {
    print_usage(argv);
    printf("options: \n");
    print_options();
    return 0;
  }
  if (option("version")->seen) {
    printf("SLA - Solitary Lossless Audio Compressor Version %s \n", SLA_VERSION_STRING);
    return 0;
  }
  if (num_files < 1) { fprintf(stderr, "%s: input file must be specified. \n", argv[0]); return 1; }
  if (!(is_batch_possible && option("batch")->seen)) {
    if (num_files < 2) { fprintf(stderr, "%s: output file must be specified. \n", argv[0]); return 1; }
    if (num_files > 2) { fprintf(stderr, "%s: Too many strings specified. \n", argv[0]); return 1; }
  }
  if (option("decode")->seen && option("encode")->seen) {
    fprintf(stderr, "%s: encode and decode mode cannot specify simultaneously. \n", argv[0]);
    return 1;
  }
  if (option("verpose")->seen) verbose = 1;
  else if (option("quiet")->seen) verbose = 0;

  if (option("decode")->seen) {
    uint8_t crc = 1;
    if (option("crc-check")->seen) crc = (strcmp(option("crc-check")->value, "yes") == 0) ? 1 : 0;
    if (option("batch")->seen) {
      if (batch_decode(option("batch")->value, files, num_files, crc, verbose) != 0) {
        fprintf(stderr, "%s: failed to decode some files. \n", argv[0]);
        return 1;
      }
    } else if (option("streaming")->seen) {
      if (streaming_decode_file(files[0], files[1], crc, verbose) != 0) {
        fprintf(stderr, "%s: failed to streaming decode %s. \n", argv[0], files[0]);
        return 1;
      }
    } else if (decode_file(files[0], files[1], crc, verbose) != 0) {
      fprintf(stderr, "%s: failed to decode %s. \n", argv[0], files[0]);
      return 1;
    }
  } else if (option("encode")->seen) {
    uint32_t preset_no = DEFAULT_PRESET;
    if (option("mode")->seen) {
      preset_no = (uint32_t)strtol(option("mode")->value, NULL, 10);
      if (preset_no >= NUM_PRESETS) { fprintf(stderr, "%s: encode preset number is out of range. \n", argv[0]); return 1; }
    }
    if (option("batch")->seen) {
      if (batch_encode(option("batch")->value, files, num_files, preset_no, verbose) != 0) return 1;
    } else {
      struct SLAEncoder* encoder = make_encoder(verbose);
      int rc;
      if (!encoder) return 1;
      rc = encode_file(encoder, files[0], files[1], preset_no, verbose);
      SLAEncoder_Destroy(encoder);
      if (rc != 0) return 1;
    }
  } else {
    fprintf(stderr, "%s: decode(-d) or encode(-e) option must be specified. \n", argv[0]);
    return 1;
  }
  return 0;
}
