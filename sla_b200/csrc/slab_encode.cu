/*
 * slab_encode.cu - launch sequence of SLAEncoder_EncodeWhole / EncodeBlock on the GPU
 * (reference: src/SLAEncoder.c:458-932).  Kernels live in slab_encode_kernels*.cuh.
 *
 * Two host synchronisations per call: (1) after the segment chain (segment count sizes the search
 * launches) and (2) after the partition search (block table; the host provides the analysis windows
 * for the distinct block lengths, computed with the host libm exactly as the reference does).  A
 * third, final one returns sizes and statistics.
 */
#include "slab_ctx.cuh"
#include "slab_encode_kernels2.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

enum {
  EA_INPUT = 0, EA_MISC, EA_FLAGS, EA_SEG_START, EA_SEG_LEN, EA_SEG_KIND, EA_PP, EA_TT, EA_ADJ,
  EA_SEG_NPARTS, EA_SEG_PARTS, EA_SEG_BLK0, EA_BLK_START, EA_BLK_LEN, EA_BLK_FLAG, EA_BLK_WIN,
  EA_CHAN, EA_PARCOR_D, EA_CODE, EA_KQ, EA_BLK_TYPE, EA_R1, EA_R3, EA_LT_D, EA_LTQ, EA_BLK_MODE,
  EA_BLK_HDR, EA_META, EA_BLK_SIZE, EA_BLK_OFF, EA_OUT, EA_ACORR, EA_MAXABS, EA_LTAC, EA_BLK_PST, EA_RISK, EA_FFT, EA_LPC_RISK, EA_DEFER, EA_LT_WIDE,
  EA_FILE_TAB, EA_CHUNK_FILE, EA_SEG_SLOTS, EA_BLK_LSHIFT, EA_COUNT_
};

static_assert(EA_COUNT_ <= SLAB_NUM_ARENAS - SLAB_USER_BUFFERS, "encoder arenas collide with the user buffers");

#define SLAB_PI 3.1415926535897932384626433832795029     /* SLAUtility.h:13 */

/* Analysis window for one block length, computed on the host with the host libm so that it is
 * bit-identical to the reference's (SLAUtility.c:88-189); cached per handle. */
static const double* get_window(SlabCtx* ctx, uint32_t type, uint32_t n)
{
  if (type == 0) return NULL;                   /* rectangular: multiply by 1.0 is the identity */
  for (uint32_t i = 0; i < ctx->num_windows; i++)
    if (ctx->windows[i].type == type && ctx->windows[i].length == n) return ctx->windows[i].dev;
  double* h = (double*)malloc(sizeof(double) * n);
  if (!h) return NULL;
  if (n == 1) h[0] = 1.0;
  else for (uint32_t i = 0; i < n; i++) {
    const double x = (double)i / (n - 1);
    switch (type) {
      case 1: h[i] = sin(SLAB_PI * x); break;
      case 2: h[i] = 0.5f - 0.5f * cos(2.0f * SLAB_PI * x); break;
      case 3: h[i] = 0.42f - 0.5f * cos(2.0f * SLAB_PI * x) + 0.08f * cos(4.0f * SLAB_PI * x); break;
      default: h[i] = sin((SLAB_PI / 2.0f) * sin(SLAB_PI * x) * sin(SLAB_PI * x)); break;
    }
  }
  double* d = NULL;
  if (cudaMalloc((void**)&d, sizeof(double) * n) != cudaSuccess) { free(h); return NULL; }
  cudaMemcpyAsync(d, h, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  free(h);
  if (ctx->num_windows == ctx->cap_windows) {
    const uint32_t ncap = ctx->cap_windows ? ctx->cap_windows * 2 : 32;
    SlabCtx::WindowEntry* grown = (SlabCtx::WindowEntry*)realloc(ctx->windows, sizeof(SlabCtx::WindowEntry) * ncap);
    if (!grown) { cudaFree(d); return NULL; }
    ctx->windows = grown; ctx->cap_windows = ncap;
  }
  ctx->windows[ctx->num_windows].type = type;
  ctx->windows[ctx->num_windows].length = n;
  ctx->windows[ctx->num_windows].dev = d;
  ctx->num_windows++;
  return d;
}

/* SLAB200_PACK_FAST=0 sends every block through the general packing kernel (A/B measurements) */
static bool env_pack_fast(void)
{
  const char* v = getenv("SLAB200_PACK_FAST");
  return !(v != NULL && v[0] == '0');
}
/* SLAB200_DEBUG_OFF (A/B measurements): bit 0 = no faithful-FFT fallback, bit 1 = no exact-order autocorrelation
 * fallback, bit 2 = scalar long-term lag sums only (no tensor-core kernel) */
static unsigned env_debug_off(void)
{
  const char* v = getenv("SLAB200_DEBUG_OFF");
  return v ? (unsigned)strtoul(v, NULL, 10) : 0u;
}

/* Trigonometric factors of the reference's FFT pair (Numerical Recipes four1 / realft as vendored in
 * SLAUtility.c:220-319) for one transform size: each stage's (wr, wi) sequence, produced with the
 * reference's own recurrences from sin() of the HOST libm - the only way the device butterflies can
 * reproduce the reference's doubles bit for bit.  cf / ci: complex stages for isign +1 / -1, the stage with
 * butterfly distance `half` at offset half - 1; rf / ri: the real-transform pass, i = 2 .. n/4. */
static int get_fft_tables(SlabCtx* ctx, uint32_t n, LtFftTables* out)
{
  if (ctx->fft_tab_size != n) {
    for (int i = 0; i < 4; i++) if (ctx->fft_tab[i]) { cudaFree(ctx->fft_tab[i]); ctx->fft_tab[i] = NULL; }
    ctx->fft_tab_size = 0;
    const uint32_t nn = n >> 1;
    const size_t nc = (size_t)(nn - 1u) * 2u, nr = (size_t)((n >> 2) - 1u) * 2u;
    double* h = (double*)malloc(sizeof(double) * (nc > nr ? nc : nr));
    if (!h) return -1;
    for (int t = 0; t < 4; t++) {
      const size_t cnt = t < 2 ? nc : nr;
      if (t < 2) {
        const double isign = t == 0 ? 1.0 : -1.0;
        for (uint32_t half = 1; half < nn; half <<= 1) {
          const uint32_t mmax = 2u * half;
          const double theta = isign * (6.28318530717959 / (double)mmax);
          double wt = sin(0.5 * theta);
          const double wpr = -2.0 * wt * wt, wpi = sin(theta);
          double wr = 1.0, wi = 0.0;
          for (uint32_t mc = 0; mc < half; mc++) {
            h[2u * (size_t)(half - 1u + mc)] = wr; h[2u * (size_t)(half - 1u + mc) + 1u] = wi;
            wt = wr;
            wr = wt * wpr - wi * wpi + wr;
            wi = wi * wpr + wt * wpi + wi;
          }
        }
      } else {
        double theta = 3.141592653589793 / (double)(n >> 1);
        if (t == 3) theta = -theta;
        double wt = sin(0.5 * theta);
        const double wpr = -2.0 * wt * wt, wpi = sin(theta);
        double wr = 1.0 + wpr, wi = wpi;
        for (uint32_t i = 2; i <= (n >> 2); i++) {
          h[2u * (size_t)(i - 2u)] = wr; h[2u * (size_t)(i - 2u) + 1u] = wi;
          wt = wr;
          wr = wt * wpr - wi * wpi + wr;
          wi = wi * wpr + wt * wpi + wi;
        }
      }
      if (cudaMalloc((void**)&ctx->fft_tab[t], sizeof(double) * (cnt ? cnt : 2)) != cudaSuccess) { free(h); return -1; }
      cudaMemcpyAsync(ctx->fft_tab[t], h, sizeof(double) * cnt, cudaMemcpyHostToDevice, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
    }
    free(h);
    ctx->fft_tab_size = n;
  }
  out->cf = ctx->fft_tab[0]; out->ci = ctx->fft_tab[1]; out->rf = ctx->fft_tab[2]; out->ri = ctx->fft_tab[3];
  return 0;
}

/* E6a + E6b + E6a': exact lag sums and the pitch / tap solve for every block x channel, then the
 * reference's own FFT autocorrelation for the few whose decisions its round-off could turn */
static int run_longterm(SlabCtx* ctx, const EncShape& sh, uint32_t fft_size, uint32_t nblocks, size_t nbc, uint32_t maxlen,
    const uint32_t* d_blk_pst, const uint32_t* d_blk_len, const uint32_t* d_type, const int32_t* d_r1, double* d_ltac,
    EncChan* d_chan, double* d_ltd, int32_t* d_ltq, uint32_t* d_risk_count, cudaStream_t serial_stream)
{
  /* block rounded up to 16 equal ranges of whole 17-step turns, plus the look-ahead of the widest lag */
  const size_t lt_steps = ((((size_t)maxlen + LT_PARTS - 1u) / LT_PARTS + LT_TILE - 1u) / LT_TILE) * LT_TILE;
  const size_t smem = sizeof(int32_t) * (LT_PARTS * lt_steps + LT_LAGS_PAD + LT_TILE + 16u);
  const bool faithful = (env_debug_off() & 1u) == 0 && fft_size >= 8u && fft_size <= (1u << 18) && (fft_size & (fft_size - 1u)) == 0 && fft_size >= maxlen;
  uint32_t* d_risk = NULL;
  if (faithful) {
    d_risk = slab_arena_as<uint32_t>(ctx, EA_RISK, 2u * nbc + 2u);      /* both kernels may list a block x channel */
    if (!d_risk) return -1;
  }
  if (slab_opt_in_smem(k_enc_ltcorr, smem)) return -1;
  const uint32_t* d_wide = NULL;
  if ((env_debug_off() & 4u) == 0 && maxlen <= LTM_M * LTM_MAXF) {
    /* tensor-core form for residuals below 2^23; it flags the block x channels it leaves to the scalar kernel */
    uint32_t* w = slab_arena_as<uint32_t>(ctx, EA_LT_WIDE, nbc + 1u);
    if (!w || slab_opt_in_smem(k_enc_ltcorr_mma, LTM_SMEM(3))) return -1;
    SLAB_RUN(ctx, "E6a k_enc_ltcorr_mma", k_enc_ltcorr_mma, (unsigned)nbc, 256, LTM_SMEM(3), sh, d_blk_pst, d_blk_len, d_type, d_r1, d_ltac,
             d_risk, d_risk_count, w, d_risk_count + (M_LT_WIDE - M_RISK));
    d_wide = w;
  }
  SLAB_RUN(ctx, "E6a k_enc_ltcorr", k_enc_ltcorr, (unsigned)nbc, LT_THREADS, smem, sh, d_blk_pst, d_blk_len, d_type, d_r1, d_ltac,
           d_risk, d_risk_count, d_wide);
  /* from here to the packing kernel everything is one thread per block x channel (or less): in chunk mode
   * these run on the high-priority stream, next to the bulk kernels of the other chunks in flight */
  if (serial_stream != NULL && slab_hop(ctx, serial_stream) != 0) return -1;
  SLAB_RUN(ctx, "E6b k_enc_ltsolve", k_enc_ltsolve, slab_div_up(nbc, 4), 128, 0, sh, nblocks, d_type, d_ltac, d_chan, d_ltd, d_ltq,
           d_risk, d_risk_count);
  if (faithful) {
    LtFftTables tb;
    if (get_fft_tables(ctx, fft_size, &tb) != 0) { slab_set_error("sla_b200: FFT table allocation failed"); return -1; }
    /* one CTA of 1024 threads per transform pair is the fastest shape measured (C2: 0.92 ms; 512 / 256 / 128 threads
     * with three CTAs per SM: 1.07 / 1.12 / 1.19 ms) - SLAB200_LTFFT_THREADS repeats the measurement */
    static const unsigned fft_threads = []{ const char* v = getenv("SLAB200_LTFFT_THREADS"); unsigned t = v ? (unsigned)strtoul(v, NULL, 10) : 1024u;
                                            return (t == 128u || t == 256u || t == 512u || t == 1024u) ? t : 1024u; }();
    const unsigned per_sm = fft_threads >= 1024u ? 2u : 3u;
    unsigned grid = nbc < 148u * per_sm ? (unsigned)nbc : 148u * per_sm;
    double* d_fft = slab_arena_as<double>(ctx, EA_FFT, (size_t)grid * 2u * fft_size);
    const size_t fsm = sizeof(double) * 2u * ((fft_size >> 1) < LTFFT_GROUP ? (fft_size >> 1) : LTFFT_GROUP);
    if (!d_fft || slab_opt_in_smem(k_enc_ltfft, fsm)) return -1;
    SLAB_RUN(ctx, "E6c k_enc_ltfft", k_enc_ltfft, grid, fft_threads, fsm, sh, fft_size, d_blk_pst, d_blk_len, d_r1, d_risk, d_risk_count,
             d_fft, tb, d_ltac, d_chan, d_ltd, d_ltq);
  }
  return 0;
}

template <typename K> static int opt_in_smem(K kernel, size_t bytes) { return slab_opt_in_smem(kernel, bytes); }

#define ARENA(T, slot, count) slab_arena_as<T>(ctx, slot, (size_t)(count));

static int slab_encode_impl(SlabCtx* ctx, SlabEncodeJob* job);

/* the launch sequence below may leave the context on its high-priority stream: callers always find the main one */
extern "C" int slab_encode(SlabCtx* ctx, SlabEncodeJob* job)
{
  const int rc = slab_encode_impl(ctx, job);
  ctx->stream = ctx->stream_main;
  return rc;
}

static int slab_encode_impl(SlabCtx* ctx, SlabEncodeJob* job)
{
  EncShape sh;
  memset(&sh, 0, sizeof(sh));
  sh.nch = job->num_channels; sh.bits = job->bits_per_sample; sh.rate = job->sampling_rate;
  sh.P = job->parcor_order; sh.T = job->longterm_order; sh.lms = job->lms_order;
  sh.ms = (job->ch_process == 1); sh.window_type = job->window_type; sh.maxblk = job->max_block_samples;
  sh.N = job->num_samples;
  sh.nnmax = (sh.maxblk + SLAB_GRID - 1) / SLAB_GRID + 1u;
  sh.wide = sh.bits > 24u;
  sh.ac_scale = ldexp(1.0, -62) * (double)(job->fft_size / 2u);
  const int pmax = sh.P <= 8 ? 8 : sh.P <= 16 ? 16 : sh.P <= 32 ? 32 : 64;
  sh.pstride = (uint32_t)pmax + 1u;
  job->overflow = 0; job->num_blocks = 0; job->total_bytes = 0; job->max_block_size = 0;
  job->max_bit_per_second = 0; job->input_or_mask = 0; job->offset_lshift = 0;
  ctx->launches = 0;
  slab_prof_reset(ctx);
  if (sh.nch < 1 || sh.nch > SLAB_MAX_CH || sh.P < 1 || sh.P > SLAB_MAX_PARCOR || sh.T > SLAB_MAX_TAPS ||
      (sh.T & 1u) == 0 || sh.lms < 4 || sh.lms > SLAB_MAX_LMS || (sh.lms & (sh.lms - 1)) != 0 ||
      sh.bits < 1 || sh.bits > 32 || sh.N == 0 || sh.maxblk < SLAB_MIN_BLOCK || sh.maxblk > 16384u ||
      (job->single_block && sh.N > 16384u)) {
    slab_set_error("sla_b200: encode parameters outside the supported envelope "
                   "(1..8 ch, PARCOR 1..64, odd taps <= 7, LMS 4/8/16/32, block 2048..16384)");
    return -1;
  }
  const uint32_t N = sh.N, nch = sh.nch;
  const uint32_t nfiles = job->num_files;
  if (nfiles > 0) {
    bool ok = job->file_start != NULL && job->file_len != NULL && job->files != NULL && !job->single_block && !job->mask_only &&
              job->on_consumed == NULL && job->forced_lshift < 0 && job->first_sample == 0 && job->soft_end == 0 &&
              job->records == NULL && job->residual_out == NULL;
    for (uint32_t f = 0; ok && f < nfiles; f++) {
      const uint64_t end = (uint64_t)job->file_start[f] + job->file_len[f];
      ok = (job->file_start[f] & (SLAB_GRID - 1u)) == 0 && job->file_len[f] > 0 &&
           end <= ((f + 1u < nfiles) ? job->file_start[f + 1u] : N);
    }
    if (!ok) { slab_set_error("sla_b200: merged encode: file table or mode not supported"); return -1; }
    memset(job->files, 0, sizeof(SlabFileResult) * nfiles);
  }
  ctx->stream = ctx->stream_main;        /* an earlier call that failed half-way may have left the other stream selected */
  if (job->high_priority && ctx->stream_hi != NULL && job->input_on_device) {
    if (slab_hop(ctx, ctx->stream_hi) != 0) return -1;      /* ordered after what the caller queued on the main stream */
  }
  cudaStream_t st = ctx->stream;
  SLAB_CUDA_TRY(cudaEventRecord(ctx->ev[0], st));

  /* ---- input planes on the device ---- */
  InPtrs in;
  memset(&in, 0, sizeof(in));
  bool vec = true;
  if (job->input_on_device) {
    for (uint32_t c = 0; c < nch; c++) { in.p[c] = job->input[c]; vec = vec && (((uintptr_t)in.p[c] & 15u) == 0); }
  } else {
    const size_t plane = ((size_t)N + 3u) & ~(size_t)3u;
    int32_t* d_in = ARENA(int32_t, EA_INPUT, plane * nch);
    if (!d_in) return -1;
    for (uint32_t c = 0; c < nch; c++) {
      SLAB_CUDA_TRY(cudaMemcpyAsync(d_in + plane * c, job->input[c], (size_t)N * 4u, cudaMemcpyHostToDevice, st));
      in.p[c] = d_in + plane * c;
    }
  }
  SLAB_CUDA_TRY(cudaEventRecord(ctx->ev[1], st));

  uint32_t* d_misc = ARENA(uint32_t, EA_MISC, M_COUNT);
  uint32_t* h_misc = (uint32_t*)slab_pinned(ctx, 4096 + sizeof(uint32_t) * (size_t)nfiles);
  const uint32_t nchunks = (N + SLAB_GRID - 1) / SLAB_GRID;
  uint32_t* d_flags = ARENA(uint32_t, EA_FLAGS, nchunks + 1u);
  if (!d_misc || !h_misc || !d_flags) return -1;
  /* merged job: file table (start | len | first segment slot), per-file OR mask / segment count / first
   * segment, and the file of every 1024-sample chunk */
  uint32_t* d_file_tab = NULL; uint32_t* d_file_or = NULL; uint32_t* d_file_nseg = NULL; uint32_t* d_file_seg0 = NULL;
  uint32_t* d_chunk_file = NULL;
  uint32_t seg_slots = 0;
  if (nfiles > 0) {
    uint32_t* h = (uint32_t*)malloc(sizeof(uint32_t) * 3u * (size_t)nfiles);
    d_file_tab = ARENA(uint32_t, EA_FILE_TAB, 6u * (size_t)nfiles);
    d_chunk_file = ARENA(uint32_t, EA_CHUNK_FILE, nchunks + 1u);
    if (!h || !d_file_tab || !d_chunk_file) { free(h); return -1; }
    d_file_or = d_file_tab + 3u * (size_t)nfiles; d_file_nseg = d_file_or + nfiles; d_file_seg0 = d_file_nseg + nfiles;
    for (uint32_t f = 0; f < nfiles; f++) {
      h[3u * f] = job->file_start[f]; h[3u * f + 1u] = job->file_len[f]; h[3u * f + 2u] = seg_slots;
      seg_slots += job->file_len[f] / SLAB_MIN_BLOCK + 2u;
    }
    /* only the small table crosses PCIe (pageable source: staged before the call returns); the chunk map, 4 bytes
     * per 1024 samples, is built on the device */
    cudaError_t fe = cudaMemcpyAsync(d_file_tab, h, sizeof(uint32_t) * 3u * nfiles, cudaMemcpyHostToDevice, st);
    if (fe == cudaSuccess) fe = cudaMemsetAsync(d_file_or, 0, sizeof(uint32_t) * 3u * nfiles, st);
    free(h);
    SLAB_CUDA_TRY(fe);
    SLAB_RUN(ctx, "E0 k_enc_chunk_map", k_enc_chunk_map, slab_div_up(nfiles, 128), 128, 0, (const uint32_t*)d_file_tab, 3u, nfiles, nchunks, d_chunk_file);
  }
  /* The OR mask and the segment chain are what the next chunk of a pipelined call waits for
   * (on_consumed): in chunk mode they run on the context's high-priority stream, so they do not queue
   * behind the bulk kernels of the chunks already in flight.  The host synchronises that stream before
   * anything else of this call is launched, which orders the two streams. */
  const bool chain_hi = job->on_consumed != NULL && ctx->stream_hi != NULL && job->input_on_device && st != ctx->stream_hi;
  if (chain_hi) {
    /* inputs produced on the main stream (the PCM de-interleave) come first */
    SLAB_CUDA_TRY(cudaEventRecord(ctx->ev_join, st));
    SLAB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream_hi, ctx->ev_join, 0));
    ctx->stream = ctx->stream_hi;
  }
  cudaStream_t st_chain = ctx->stream;
  int chain_rc = 0;
  do {
    if (cudaMemsetAsync(d_misc, 0, sizeof(uint32_t) * M_COUNT, st_chain) != cudaSuccess) { chain_rc = -1; break; }
    /* ---- E0 ---- */
    {
      unsigned grid_scan = slab_div_up(nchunks, 8);
      if (grid_scan > 148u * 8u) grid_scan = 148u * 8u;       /* persistent: 8 CTAs of 8 warps per SM */
      if (vec) SLAB_RUN_RC(chain_rc, ctx, "E0 k_enc_scan", (k_enc_scan<true>), grid_scan, 256, 0, in, nch, N, d_flags, d_misc, (const uint32_t*)d_chunk_file, d_file_or);
      else     SLAB_RUN_RC(chain_rc, ctx, "E0 k_enc_scan", (k_enc_scan<false>), grid_scan, 256, 0, in, nch, N, d_flags, d_misc, (const uint32_t*)d_chunk_file, d_file_or);
      if (chain_rc) break;
    }
  } while (0);
  if (chain_rc) { ctx->stream = st; slab_set_error("sla_b200: launch of the scan failed"); return -1; }

  /* ---- E2: segment chain ---- */
  const uint32_t seg_cap = (nfiles > 0 ? seg_slots : N / SLAB_MIN_BLOCK) + 2u;
  uint32_t* d_seg_start = ARENA(uint32_t, EA_SEG_START, seg_cap);
  uint32_t* d_seg_len = ARENA(uint32_t, EA_SEG_LEN, seg_cap);
  uint32_t* d_seg_kind = ARENA(uint32_t, EA_SEG_KIND, seg_cap);
  if (!d_seg_start || !d_seg_len || !d_seg_kind) { ctx->stream = st; return -1; }
  if (nfiles > 0) {
    uint32_t* d_slots = ARENA(uint32_t, EA_SEG_SLOTS, 3u * (size_t)seg_cap);
    if (!d_slots) return -1;
    do {
      SLAB_RUN_RC(chain_rc, ctx, "E2 k_enc_segments", k_enc_segments, nfiles, 32, 0, in, nch, N, sh.maxblk, 0u, N, d_flags,
                  d_slots, d_slots + seg_cap, d_slots + 2u * (size_t)seg_cap, d_misc, (const uint32_t*)d_file_tab, d_file_nseg);
      if (chain_rc) break;
      SLAB_RUN_RC(chain_rc, ctx, "E2 k_scan_u32", k_scan_u32, 1, 1024, 0, (const uint32_t*)d_file_nseg, d_file_seg0, nfiles, d_misc + M_NSEG);
      if (chain_rc) break;
      SLAB_RUN_RC(chain_rc, ctx, "E2 k_enc_compact_segments", k_enc_compact_segments, nfiles, 128, 0, (const uint32_t*)d_file_tab,
                  (const uint32_t*)d_file_nseg, (const uint32_t*)d_file_seg0, (const uint32_t*)d_slots, (const uint32_t*)(d_slots + seg_cap),
                  (const uint32_t*)(d_slots + 2u * (size_t)seg_cap), d_seg_start, d_seg_len, d_seg_kind);
    } while (0);
  } else if (!job->single_block && !job->mask_only) {
    const uint32_t stop = (job->soft_end != 0 && job->soft_end < N) ? job->soft_end : N;
    SLAB_RUN_RC(chain_rc, ctx, "E2 k_enc_segments", k_enc_segments, 1, 32, 0, in, nch, N, sh.maxblk, job->first_sample, stop, d_flags, d_seg_start, d_seg_len, d_seg_kind, d_misc,
                (const uint32_t*)nullptr, (uint32_t*)nullptr);
  }
  ctx->stream = st;
  if (chain_rc) { slab_set_error("sla_b200: launch of the segment chain failed"); return -1; }
  SLAB_CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(uint32_t) * M_COUNT, cudaMemcpyDeviceToHost, st_chain));
  if (nfiles > 0) SLAB_CUDA_TRY(cudaMemcpyAsync(h_misc + 1024, d_file_or, sizeof(uint32_t) * nfiles, cudaMemcpyDeviceToHost, st_chain));
  SLAB_CUDA_TRY(cudaStreamSynchronize(st_chain));                              /* sync (1) */
  const uint32_t or_mask = h_misc[M_ORMASK];
  job->input_or_mask = or_mask;
  if (job->mask_only) return 0;
  job->consumed_samples = job->single_block ? N : h_misc[M_CONSUMED];
  if (job->on_consumed) job->on_consumed(job->user, job->consumed_samples);
  /* offset_lshift, SLAEncoder.c:425-455 */
  uint32_t lshift = 0;
  if (job->forced_lshift >= 0) lshift = (uint32_t)job->forced_lshift;
  else if (or_mask != 0) {
    uint32_t ntz = 0;
    while (((or_mask >> ntz) & 1u) == 0) ntz++;
    lshift = sh.bits - (32u - ntz);
  }
  if (lshift >= sh.bits) { slab_set_error("sla_b200: input has bits below its declared width"); return -1; }
  sh.lshift = lshift;
  job->offset_lshift = lshift;
  for (uint32_t f = 0; f < nfiles; f++) {              /* the same rule, file by file */
    const uint32_t m = h_misc[1024 + f];
    uint32_t fl = 0;
    if (m != 0) { uint32_t ntz = 0; while (((m >> ntz) & 1u) == 0) ntz++; fl = sh.bits - (32u - ntz); }
    if (fl >= sh.bits) { slab_set_error("sla_b200: input %u has bits below its declared width", f); return -1; }
    job->files[f].offset_lshift = fl;
    if (f + 1u == nfiles) job->offset_lshift = fl;
  }

  uint32_t nblocks = 0;
  const uint32_t blk_cap = N / SLAB_MIN_BLOCK + seg_cap + 2u;
  uint32_t* d_blk_start = ARENA(uint32_t, EA_BLK_START, blk_cap);
  uint32_t* d_blk_len = ARENA(uint32_t, EA_BLK_LEN, blk_cap);
  uint32_t* d_blk_flag = ARENA(uint32_t, EA_BLK_FLAG, blk_cap);
  if (!d_blk_start || !d_blk_len || !d_blk_flag) return -1;
  uint32_t* h_blk = NULL;       /* [len | flag | start] x nblocks in pinned memory */

  if (job->single_block) {
    nblocks = 1;
    h_blk = (uint32_t*)slab_pinned(ctx, 4096 + 64);
    if (!h_blk) return -1;
    h_blk += 1024;              /* keep clear of h_misc */
    h_blk[0] = N; h_blk[1] = 0; h_blk[2] = 0;
    SLAB_CUDA_TRY(cudaMemcpyAsync(d_blk_len, &h_blk[0], 4, cudaMemcpyHostToDevice, st));
    SLAB_CUDA_TRY(cudaMemcpyAsync(d_blk_flag, &h_blk[1], 4, cudaMemcpyHostToDevice, st));
    SLAB_CUDA_TRY(cudaMemcpyAsync(d_blk_start, &h_blk[2], 4, cudaMemcpyHostToDevice, st));
  } else {
    /* ---- E3: partition search ---- */
    const uint32_t nseg = h_misc[M_NSEG];
    const uint32_t lags = sh.P + 1u;
    const size_t sums = (size_t)nseg * nch * sh.nnmax * lags;
    unsigned long long* d_pp = ARENA(unsigned long long, EA_PP, sums);
    unsigned long long* d_tt = ARENA(unsigned long long, EA_TT, sums);
    double* d_adj = ARENA(double, EA_ADJ, (size_t)nseg * sh.nnmax * sh.nnmax);
    uint32_t* d_nparts = ARENA(uint32_t, EA_SEG_NPARTS, nseg + 1u);
    uint32_t* d_parts = ARENA(uint32_t, EA_SEG_PARTS, (size_t)nseg * sh.nnmax);
    uint32_t* d_blk0 = ARENA(uint32_t, EA_SEG_BLK0, nseg + 1u);
    if (!d_pp || !d_tt || !d_adj || !d_nparts || !d_parts || !d_blk0) return -1;
    const size_t ysize = sh.wide ? sizeof(double) : sizeof(int32_t);
    const size_t smem = sizeof(long long) * (size_t)(sh.nnmax - 1u) * lags + ysize * ((((size_t)sh.maxblk + SLAB_GRID - 1u) & ~(size_t)(SLAB_GRID - 1u)) + 128u);
    dim3 grid_ls(nseg, nch);
    const unsigned grid_e = nseg;
#define RUN_LAGSUMS(W, G)                                                                                   \
    do {                                                                                                  \
      if (opt_in_smem(k_enc_lagsums<W, G>, smem)) return -1;                                              \
      SLAB_RUN(ctx, "E3a k_enc_lagsums", (k_enc_lagsums<W, G>), grid_ls, 256, smem, in, sh, d_seg_start, d_seg_len, d_seg_kind, d_pp, d_tt); \
    } while (0)
    if (!sh.wide) {
      if (lags == 9u) RUN_LAGSUMS(false, 9);
      else if (lags == 17u) RUN_LAGSUMS(false, 17);
      else if (lags == 33u) RUN_LAGSUMS(false, 33);
      else RUN_LAGSUMS(false, 0);
      SLAB_RUN(ctx, "E3b k_enc_edges", (k_enc_edges<false>), grid_e, 128, 0, sh, d_seg_start, d_seg_len, d_seg_kind, d_pp, d_tt, d_adj);
    } else {
      RUN_LAGSUMS(true, 0);
      SLAB_RUN(ctx, "E3b k_enc_edges", (k_enc_edges<true>), grid_e, 128, 0, sh, d_seg_start, d_seg_len, d_seg_kind, d_pp, d_tt, d_adj);
    }
#undef RUN_LAGSUMS
    SLAB_RUN(ctx, "E3c k_enc_dijkstra", k_enc_dijkstra, slab_div_up(nseg, 64), 64, 0, sh, nseg, d_seg_len, d_seg_kind, d_adj, d_nparts, d_parts);
    SLAB_RUN(ctx, "E3d k_scan_u32", k_scan_u32, 1, 1024, 0, d_nparts, d_blk0, nseg, d_misc + M_NBLOCKS);
    SLAB_RUN(ctx, "E3d k_enc_fill_blocks", k_enc_fill_blocks, slab_div_up(nseg, 128), 128, 0, sh, nseg, d_seg_start, d_seg_kind, d_nparts, d_parts, d_blk0, d_blk_start, d_blk_len, d_blk_flag);
    SLAB_CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(uint32_t) * M_COUNT, cudaMemcpyDeviceToHost, st));
    SLAB_CUDA_TRY(cudaStreamSynchronize(st));
    nblocks = h_misc[M_NBLOCKS];
    if (nblocks == 0 || nblocks > blk_cap) { slab_set_error("sla_b200: partition search produced %u blocks", nblocks); return -1; }
    h_blk = (uint32_t*)slab_pinned(ctx, 4096 + (size_t)nblocks * 16u + 64);      /* fourth column: block sizes of a merged job */
    if (!h_blk) return -1;
    h_misc = h_blk;             /* the pinned buffer may have moved */
    h_blk += 1024;
    SLAB_CUDA_TRY(cudaMemcpyAsync(h_blk, d_blk_len, (size_t)nblocks * 4u, cudaMemcpyDeviceToHost, st));
    SLAB_CUDA_TRY(cudaMemcpyAsync(h_blk + nblocks, d_blk_flag, (size_t)nblocks * 4u, cudaMemcpyDeviceToHost, st));
    SLAB_CUDA_TRY(cudaMemcpyAsync(h_blk + 2u * (size_t)nblocks, d_blk_start, (size_t)nblocks * 4u, cudaMemcpyDeviceToHost, st));
    SLAB_CUDA_TRY(cudaStreamSynchronize(st));                                  /* sync (2) */
  }
  job->num_blocks = nblocks;

  /* ---- analysis windows for the distinct block lengths ---- */
  const double** h_win = (const double**)malloc(sizeof(double*) * nblocks);
  uint32_t* h_pst = (uint32_t*)slab_host_scratch(ctx, sizeof(uint32_t) * (3u * (size_t)nblocks + 3u));
  if (!h_win || !h_pst) { free(h_win); return -1; }
  uint32_t* h_blk_lshift = h_pst + nblocks + 1u;       /* merged job: offset_lshift and file of every block */
  uint32_t* h_blk_file = h_blk_lshift + nblocks + 1u;
  if (nfiles > 0) {
    uint32_t f = 0;
    for (uint32_t b = 0; b < nblocks; b++) {
      const uint32_t start = h_blk[2u * (size_t)nblocks + b];
      while (f + 1u < nfiles && start >= job->file_start[f + 1u]) f++;
      h_blk_file[b] = f; h_blk_lshift[b] = job->files[f].offset_lshift;
      job->files[f].num_blocks++;
    }
  }
  uint32_t maxlen = 0, padded = 0;
  for (uint32_t b = 0; b < nblocks; b++) {
    const uint32_t len = h_blk[b], flag = job->single_block ? 0u : h_blk[nblocks + b];
    h_pst[b] = padded;
    padded += (len + 7u) & ~7u;
    if (len > maxlen) maxlen = len;
    h_win[b] = NULL;
    if (flag == 0 && sh.window_type != 0) {
      /* consecutive blocks usually share a length: check the previous one before the cache */
      if (b > 0 && h_blk[b - 1] == len && h_win[b - 1] != NULL) h_win[b] = h_win[b - 1];
      else {
        h_win[b] = get_window(ctx, sh.window_type, len);
        if (!h_win[b]) { free(h_win); slab_set_error("sla_b200: window table allocation failed"); return -1; }
      }
    }
  }
  const double** d_win = (const double**)slab_arena(ctx, EA_BLK_WIN, sizeof(double*) * nblocks);
  uint32_t* d_blk_pst = ARENA(uint32_t, EA_BLK_PST, nblocks + 1u);
  if (!d_win || !d_blk_pst) { free(h_win); return -1; }
  cudaError_t we = cudaMemcpyAsync(d_win, h_win, sizeof(double*) * nblocks, cudaMemcpyHostToDevice, st);
  if (we == cudaSuccess) we = cudaMemcpyAsync(d_blk_pst, h_pst, sizeof(uint32_t) * nblocks, cudaMemcpyHostToDevice, st);
  if (we == cudaSuccess && nfiles > 0) {
    uint32_t* d_bl = ARENA(uint32_t, EA_BLK_LSHIFT, nblocks + 1u);
    if (!d_bl) { free(h_win); return -1; }
    we = cudaMemcpyAsync(d_bl, h_blk_lshift, sizeof(uint32_t) * nblocks, cudaMemcpyHostToDevice, st);
    sh.blk_lshift = d_bl;
  }
  if (we == cudaSuccess) we = cudaStreamSynchronize(st);
  free(h_win);
  SLAB_CUDA_TRY(we);
  const size_t NP = ((size_t)padded + 15u) & ~(size_t)7u;    /* plane stride of r1 / r3 / meta */
  sh.NP = (uint32_t)NP;

  /* ---- per block x channel state ---- */
  const size_t nbc = (size_t)nblocks * nch;
  EncChan* d_chan = ARENA(EncChan, EA_CHAN, nbc);
  double* d_parcor = ARENA(double, EA_PARCOR_D, nbc * (SLAB_MAX_PARCOR + 1));
  int32_t* d_code = ARENA(int32_t, EA_CODE, nbc * (SLAB_MAX_PARCOR + 1));
  int32_t* d_kq = ARENA(int32_t, EA_KQ, nbc * sh.pstride);
  uint32_t* d_type = ARENA(uint32_t, EA_BLK_TYPE, nblocks);
  int32_t* d_r1 = ARENA(int32_t, EA_R1, NP * nch);
  int32_t* d_r3 = ARENA(int32_t, EA_R3, NP * nch);
  double* d_ltd = ARENA(double, EA_LT_D, nbc * 8);
  int32_t* d_ltq = ARENA(int32_t, EA_LTQ, nbc * 8);
  uint32_t* d_mode = ARENA(uint32_t, EA_BLK_MODE, nblocks);
  uint32_t* d_hdr = ARENA(uint32_t, EA_BLK_HDR, nblocks);
  uint16_t* d_meta = ARENA(uint16_t, EA_META, NP * nch);
  uint32_t* d_size = ARENA(uint32_t, EA_BLK_SIZE, nblocks + 1u);
  uint32_t* d_off = ARENA(uint32_t, EA_BLK_OFF, nblocks + 1u);
  double* d_acorr = ARENA(double, EA_ACORR, nbc * (SLAB_MAX_PARCOR + 1));
  uint32_t* d_maxabs = ARENA(uint32_t, EA_MAXABS, nbc);
  double* d_ltac = ARENA(double, EA_LTAC, nbc * 264u);
  if (!d_chan || !d_parcor || !d_code || !d_kq || !d_type || !d_r1 || !d_r3 || !d_ltd || !d_ltq || !d_mode ||
      !d_hdr || !d_meta || !d_size || !d_off || !d_acorr || !d_maxabs || !d_ltac) return -1;
  SLAB_CUDA_TRY(cudaMemsetAsync(d_chan, 0, sizeof(EncChan) * nbc, st));
  SLAB_CUDA_TRY(cudaMemsetAsync(d_ltd, 0, sizeof(double) * nbc * 8, st));
  SLAB_CUDA_TRY(cudaMemsetAsync(d_ltq, 0, sizeof(int32_t) * nbc * 8, st));
  if (job->records) {
    SLAB_CUDA_TRY(cudaMemsetAsync(d_parcor, 0, sizeof(double) * nbc * (SLAB_MAX_PARCOR + 1), st));
    SLAB_CUDA_TRY(cudaMemsetAsync(d_code, 0, sizeof(int32_t) * nbc * (SLAB_MAX_PARCOR + 1), st));
  }
  if (job->residual_out) {
    SLAB_CUDA_TRY(cudaMemsetAsync(d_r3, 0, sizeof(int32_t) * NP * nch, st));
  }

  /* output staging */
  uint32_t cap = job->out_capacity > job->out_offset ? job->out_capacity - job->out_offset : 0u;
  uint8_t* d_out;
  if (job->out_on_device) d_out = job->out + job->out_offset;
  else {
    const uint64_t bound = 2ull * nch * N * ((sh.bits + 7u) / 8u) + (uint64_t)nblocks * 1024u + 65536u;
    if ((uint64_t)cap > bound) cap = (uint32_t)bound;
    d_out = (uint8_t*)slab_arena(ctx, EA_OUT, (size_t)cap + 64u);
    if (!d_out) return -1;
  }
  sh.out_cap = cap;

  /* ---- E4 ---- */
  {
    const size_t smem = sizeof(double) * ((size_t)maxlen + 2u * 33u + 16u);
    uint32_t* d_lpc_risk = slab_arena_as<uint32_t>(ctx, EA_LPC_RISK, nbc + 1u);
    if (!d_lpc_risk) return -1;
#define RUN_ANALYSIS(L)                                                                                   \
    do {                                                                                                  \
      if (opt_in_smem(k_enc_autocorr<L, false>, smem) || opt_in_smem(k_enc_autocorr<L, true>, smem)) return -1; \
      SLAB_RUN(ctx, "E4a k_enc_autocorr", (k_enc_autocorr<L, false>), (unsigned)nbc, 256, smem, in, sh, d_blk_start, \
               d_blk_len, d_blk_flag, d_win, d_acorr, d_maxabs, (const uint32_t*)nullptr);                \
      SLAB_RUN(ctx, "E4b k_enc_lpc", k_enc_lpc<false>, slab_div_up(nbc, 64), 64, 0, sh, nblocks, d_blk_len, d_blk_flag, \
               d_acorr, d_maxabs, d_chan, d_parcor, d_code, d_kq, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr); \
      if (env_debug_off() & 2u) break;                                                                    \
      SLAB_RUN(ctx, "E4b k_enc_lpc_risk", k_enc_lpc<true>, slab_div_up(nbc, 64), 64, 0, sh, nblocks, d_blk_len, d_blk_flag, \
               d_acorr, d_maxabs, d_chan, d_parcor, d_code, d_kq, d_lpc_risk, (const uint32_t*)nullptr, d_misc + M_LPC_RISK); \
      /* the few block x channels whose recursion is badly conditioned: lag sums in the reference's order */ \
      SLAB_RUN(ctx, "E4c k_enc_autocorr_exact", (k_enc_autocorr<L, true>), (unsigned)nbc, 256, smem, in, sh, d_blk_start, \
               d_blk_len, d_blk_flag, d_win, d_acorr, d_maxabs, (const uint32_t*)d_lpc_risk);             \
      SLAB_RUN(ctx, "E4d k_enc_lpc_exact", k_enc_lpc<false>, slab_div_up(nbc, 64), 64, 0, sh, nblocks, d_blk_len, d_blk_flag, \
               d_acorr, d_maxabs, d_chan, d_parcor, d_code, d_kq, (uint32_t*)nullptr, (const uint32_t*)d_lpc_risk, (uint32_t*)nullptr); \
    } while (0)
    if (sh.P <= 8) RUN_ANALYSIS(9);
    else if (sh.P <= 16) RUN_ANALYSIS(17);
    else if (sh.P <= 32) RUN_ANALYSIS(33);
    else RUN_ANALYSIS(0);
#undef RUN_ANALYSIS
  }
  SLAB_RUN(ctx, "E4 k_enc_blocktype", k_enc_blocktype, slab_div_up(nblocks, 128), 128, 0, sh, nblocks, d_blk_flag, d_chan, d_type);
  /* ---- E5 ---- */
  {
    const uint32_t spb = (maxlen + SLAB_SLICE - 1) / SLAB_SLICE;
    const unsigned grid = slab_div_up((uint64_t)nbc * spb, 128);
    switch (pmax) {
      case 8: SLAB_RUN(ctx, "E5 k_enc_parcor", (k_enc_parcor<8>), grid, 128, 0, in, sh, nblocks, spb, d_blk_start, d_blk_pst, d_blk_len, d_type, d_kq, d_r1); break;
      case 16: SLAB_RUN(ctx, "E5 k_enc_parcor", (k_enc_parcor<16>), grid, 128, 0, in, sh, nblocks, spb, d_blk_start, d_blk_pst, d_blk_len, d_type, d_kq, d_r1); break;
      case 32: SLAB_RUN(ctx, "E5 k_enc_parcor", (k_enc_parcor<32>), grid, 128, 0, in, sh, nblocks, spb, d_blk_start, d_blk_pst, d_blk_len, d_type, d_kq, d_r1); break;
      default: SLAB_RUN(ctx, "E5 k_enc_parcor", (k_enc_parcor<64>), grid, 128, 0, in, sh, nblocks, spb, d_blk_start, d_blk_pst, d_blk_len, d_type, d_kq, d_r1); break;
    }
  }
  /* ---- E6 ---- */
  const cudaStream_t serial_stream = (job->on_consumed != NULL && ctx->stream_hi != NULL && job->input_on_device && st != ctx->stream_hi) ? ctx->stream_hi : NULL;
  if (run_longterm(ctx, sh, job->fft_size, nblocks, nbc, maxlen, d_blk_pst, d_blk_len, d_type, d_r1, d_ltac, d_chan, d_ltd, d_ltq,
                   d_misc + M_RISK, serial_stream) != 0) { ctx->stream = st; return -1; }
  /* ---- E7/E8 ---- */
  {
    const unsigned grid = slab_div_up(nbc, 64);
#define RUN_LTLMS(N, TP) SLAB_RUN(ctx, "E7 k_enc_ltlms", (k_enc_ltlms<N, TP>), grid, 64, 0, sh, nblocks, d_blk_pst, d_blk_len, d_type, d_ltq, d_r1, d_r3, d_chan)
    if (sh.lms == 4 || sh.lms == 8) {
      if (sh.lms == 4) { if (sh.T <= 1) RUN_LTLMS(4, 1); else if (sh.T <= 3) RUN_LTLMS(4, 3); else RUN_LTLMS(4, 7); }
      else             { if (sh.T <= 1) RUN_LTLMS(8, 1); else if (sh.T <= 3) RUN_LTLMS(8, 3); else RUN_LTLMS(8, 7); }
    } else {
      SLAB_RUN(ctx, "E7 k_enc_ltlms_generic", k_enc_ltlms_generic, grid, 64, 0, sh, nblocks, d_blk_pst, d_blk_len, d_type, d_ltq, d_r1, d_r3, d_chan);
    }
#undef RUN_LTLMS
  }
  /* ---- E9 ---- */
  SLAB_RUN(ctx, "E9 k_enc_riceprep", k_enc_riceprep, slab_div_up(nblocks, 128), 128, 0, sh, nblocks, d_blk_len, d_type, d_chan, d_mode, d_hdr);
  if (opt_in_smem(k_enc_ricetrace, sizeof(RiceTraceSmem))) return -1;
  SLAB_RUN(ctx, "E9 k_enc_ricetrace", k_enc_ricetrace, slab_div_up(nbc, 32), 32, sizeof(RiceTraceSmem), sh, nblocks, d_blk_pst, d_blk_len, d_type, d_mode, d_r3, d_chan, d_meta);
  SLAB_RUN(ctx, "E9 k_enc_blocksizes", k_enc_blocksizes, slab_div_up(nblocks, 128), 128, 0, sh, nblocks, d_blk_len, d_type, d_hdr, d_chan, d_size, d_misc);
  SLAB_RUN(ctx, "E9 k_scan_u32", k_scan_u32, 1, 1024, 0, d_size, d_off, nblocks, d_misc + M_TOTAL_BYTES);
  SLAB_RUN(ctx, "E9 k_enc_check_capacity", k_enc_check_capacity, 1, 32, 0, sh, d_misc);
  if (slab_hop(ctx, st) != 0) { ctx->stream = st; return -1; }          /* back to the main stream for the bulk kernels */
  {
    /* mono / stereo: recursive-Rice blocks by k_enc_pack_rice, everything else (and the blocks it defers) by
     * the general kernel */
    uint32_t* d_defer = NULL;
    if (nch <= 2u && env_pack_fast()) {
      d_defer = slab_arena_as<uint32_t>(ctx, EA_DEFER, nblocks + 1u);
      if (!d_defer) { ctx->stream = st; return -1; }
      if (nch == 1) SLAB_RUN(ctx, "E9 k_enc_pack_rice", k_enc_pack_rice<1>, nblocks, 256, 0, sh, d_blk_pst, d_blk_len, d_type, d_mode, d_hdr, d_off, d_chan, d_code, d_ltq, d_r3, d_meta, d_misc, d_out, d_defer);
      else SLAB_RUN(ctx, "E9 k_enc_pack_rice", k_enc_pack_rice<2>, nblocks, 256, 0, sh, d_blk_pst, d_blk_len, d_type, d_mode, d_hdr, d_off, d_chan, d_code, d_ltq, d_r3, d_meta, d_misc, d_out, d_defer);
    }
    if (nch == 1) SLAB_RUN(ctx, "E9 k_enc_pack", k_enc_pack<1>, nblocks, 256, 0, in, sh, d_blk_start, d_blk_pst, d_blk_len, d_type, d_mode, d_hdr, d_size, d_off, d_chan, d_code, d_ltq, d_r3, d_meta, d_misc, d_out, (const uint32_t*)d_defer);
    else if (nch == 2) SLAB_RUN(ctx, "E9 k_enc_pack", k_enc_pack<2>, nblocks, 256, 0, in, sh, d_blk_start, d_blk_pst, d_blk_len, d_type, d_mode, d_hdr, d_size, d_off, d_chan, d_code, d_ltq, d_r3, d_meta, d_misc, d_out, (const uint32_t*)d_defer);
    else SLAB_RUN(ctx, "E9 k_enc_pack", k_enc_pack<SLAB_MAX_CH>, nblocks, 256, 0, in, sh, d_blk_start, d_blk_pst, d_blk_len, d_type, d_mode, d_hdr, d_size, d_off, d_chan, d_code, d_ltq, d_r3, d_meta, d_misc, d_out, (const uint32_t*)nullptr);
  }
  /* ---- E10 ---- */
  SLAB_RUN(ctx, "E10 k_enc_crc", k_enc_crc, slab_div_up((uint64_t)nblocks * 32u, 128), 128, 0, nblocks, d_size, d_off, d_misc, d_out);
  SLAB_CUDA_TRY(cudaEventRecord(ctx->ev[2], st));

  SLAB_CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(uint32_t) * M_COUNT, cudaMemcpyDeviceToHost, st));
  SLAB_CUDA_TRY(cudaStreamSynchronize(st));                                    /* sync (3) */
  job->total_bytes = h_misc[M_TOTAL_BYTES];
  job->fallback_ltfft = h_misc[M_RISK]; job->fallback_exact_autocorr = h_misc[M_LPC_RISK]; job->fallback_scalar_ltcorr = h_misc[M_LT_WIDE];
  job->max_block_size = h_misc[M_MAX_BLOCK];
  job->max_bit_per_second = h_misc[M_MAX_BPS];
  if (h_misc[M_OVERFLOW]) { job->overflow = 1; job->total_bytes = 0; return 0; }
  if (nfiles > 0) {
    /* per-file statistics from the block sizes, SLAEncoder.c:887-898 (uint32 wrap included) */
    uint32_t* h_size = h_blk + 3u * (size_t)nblocks;
    uint32_t off = 0;
    SLAB_CUDA_TRY(cudaMemcpyAsync(h_size, d_size, sizeof(uint32_t) * nblocks, cudaMemcpyDeviceToHost, st));
    SLAB_CUDA_TRY(cudaStreamSynchronize(st));
    for (uint32_t b = 0; b < nblocks; b++) {
      SlabFileResult* fr = &job->files[h_blk_file[b]];
      const uint32_t size = h_size[b], bps = (8u * size * sh.rate) / h_blk[b];
      if (fr->num_bytes == 0) fr->byte_offset = off;
      fr->num_bytes += size; off += size;
      if (size > fr->max_block_size) fr->max_block_size = size;
      if (bps > fr->max_bit_per_second) fr->max_bit_per_second = bps;
    }
  }
  if (!job->out_on_device)
    SLAB_CUDA_TRY(cudaMemcpyAsync(job->out + job->out_offset, d_out, job->total_bytes, cudaMemcpyDeviceToHost, st));
  SLAB_CUDA_TRY(cudaEventRecord(ctx->ev[3], st));
  SLAB_CUDA_TRY(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ctx->last_ms[SLAB_T_H2D], ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->last_ms[SLAB_T_KERNELS], ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&ctx->last_ms[SLAB_T_D2H], ctx->ev[2], ctx->ev[3]);
  slab_prof_collect(ctx);

  /* ---- optional debug export ---- */
  if (job->records && job->max_records) {
    const uint32_t nrec = nblocks < job->max_records ? nblocks : job->max_records;
    EncChan* h_chan = (EncChan*)malloc(sizeof(EncChan) * nbc);
    double* h_parcor = (double*)malloc(sizeof(double) * nbc * (SLAB_MAX_PARCOR + 1));
    int32_t* h_code = (int32_t*)malloc(sizeof(int32_t) * nbc * (SLAB_MAX_PARCOR + 1));
    double* h_ltd = (double*)malloc(sizeof(double) * nbc * 8);
    int32_t* h_ltq = (int32_t*)malloc(sizeof(int32_t) * nbc * 8);
    uint32_t* h_tab = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nblocks * 5u);
    if (h_chan && h_parcor && h_code && h_ltd && h_ltq && h_tab) {
      cudaMemcpy(h_chan, d_chan, sizeof(EncChan) * nbc, cudaMemcpyDeviceToHost);
      cudaMemcpy(h_parcor, d_parcor, sizeof(double) * nbc * (SLAB_MAX_PARCOR + 1), cudaMemcpyDeviceToHost);
      cudaMemcpy(h_code, d_code, sizeof(int32_t) * nbc * (SLAB_MAX_PARCOR + 1), cudaMemcpyDeviceToHost);
      cudaMemcpy(h_ltd, d_ltd, sizeof(double) * nbc * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(h_ltq, d_ltq, sizeof(int32_t) * nbc * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(h_tab, d_blk_start, sizeof(uint32_t) * nblocks, cudaMemcpyDeviceToHost);
      cudaMemcpy(h_tab + nblocks, d_blk_len, sizeof(uint32_t) * nblocks, cudaMemcpyDeviceToHost);
      cudaMemcpy(h_tab + 2u * (size_t)nblocks, d_type, sizeof(uint32_t) * nblocks, cudaMemcpyDeviceToHost);
      cudaMemcpy(h_tab + 3u * (size_t)nblocks, d_size, sizeof(uint32_t) * nblocks, cudaMemcpyDeviceToHost);
      cudaMemcpy(h_tab + 4u * (size_t)nblocks, d_off, sizeof(uint32_t) * nblocks, cudaMemcpyDeviceToHost);
      for (uint32_t b = 0; b < nrec; b++) {
        SlabBlockRecord* r = &job->records[b];
        memset(r, 0, sizeof(*r));
        r->sample_offset = h_tab[b]; r->num_samples = h_tab[nblocks + b];
        r->block_type = h_tab[2u * (size_t)nblocks + b]; r->block_size = h_tab[3u * (size_t)nblocks + b];
        r->byte_offset = job->out_offset + h_tab[4u * (size_t)nblocks + b];
        for (uint32_t c = 0; c < nch; c++) {
          const size_t bc = (size_t)b * nch + c;
          r->rshift[c] = h_chan[bc].rshift; r->pitch[c] = h_chan[bc].pitch; r->rice_init[c] = h_chan[bc].rice_init;
          for (uint32_t k = 0; k <= sh.P; k++) {
            r->parcor[c][k] = h_parcor[bc * (SLAB_MAX_PARCOR + 1) + k];
            r->parcor_code[c][k] = h_code[bc * (SLAB_MAX_PARCOR + 1) + k];
          }
          for (uint32_t k = 0; k < sh.T; k++) { r->lt[c][k] = h_ltd[bc * 8 + k]; r->lt_q31[c][k] = h_ltq[bc * 8 + k]; }
        }
      }
    }
    free(h_chan); free(h_parcor); free(h_code); free(h_ltd); free(h_ltq); free(h_tab);
  }
  if (job->residual_out) {
    /* the residual planes are block-padded on the device: copy block by block */
    for (uint32_t c = 0; c < nch; c++)
      for (uint32_t b = 0; b < nblocks; b++) {
        const uint32_t start = job->single_block ? 0u : h_blk[2u * (size_t)nblocks + b];
        cudaMemcpy(job->residual_out[c] + start, d_r3 + (size_t)c * NP + h_pst[b], (size_t)h_blk[b] * 4u, cudaMemcpyDeviceToHost);
      }
  }
  SLAB_CUDA_TRY(cudaGetLastError());
  return 0;
}

/* Test hook: the long-term analysis (E6a + E6b) of the encoder on a caller-supplied residual, as one
 * block of one channel - what SLALongTermCalculator_CalculateCoef does in the reference
 * (src/SLAPredictor.c:791-980; its own KAT: test/test_SLAPredictor.c:717-768).  pitch = 0 when the
 * reference would report a failure or a period the encoder does not use (SLAEncoder.c:629-632). */
extern "C" int slab_debug_longterm(SlabCtx* ctx, const int32_t* data, uint32_t n, uint32_t taps, uint32_t fft_size,
                                   uint32_t* pitch, double* coef)
{
  if (n == 0 || n > 16384u || taps < 1 || taps > SLAB_MAX_TAPS || (taps & 1u) == 0) {
    slab_set_error("sla_b200: long-term test hook: 1..16384 samples, odd taps <= 7");
    return -1;
  }
  EncShape sh;
  memset(&sh, 0, sizeof(sh));
  sh.nch = 1; sh.bits = 16; sh.rate = 44100; sh.P = 1; sh.T = taps; sh.lms = 4; sh.maxblk = 16384; sh.N = n;
  sh.ac_scale = ldexp(1.0, -62) * (double)(fft_size / 2u);
  const size_t NP = ((size_t)n + 15u) & ~(size_t)7u;
  sh.NP = (uint32_t)NP;
  cudaStream_t st = ctx->stream;
  int32_t* d_r1 = ARENA(int32_t, EA_R1, NP);
  uint32_t* d_tab = ARENA(uint32_t, EA_BLK_PST, 4);           /* pst | len | type */
  EncChan* d_chan = ARENA(EncChan, EA_CHAN, 1);
  double* d_ltd = ARENA(double, EA_LT_D, 8);
  int32_t* d_ltq = ARENA(int32_t, EA_LTQ, 8);
  double* d_ltac = ARENA(double, EA_LTAC, 264u);
  uint32_t* h = (uint32_t*)slab_pinned(ctx, 4096);
  if (!d_r1 || !d_tab || !d_chan || !d_ltd || !d_ltq || !d_ltac || !h) return -1;
  h[0] = 0; h[1] = n; h[2] = SLAB_BLOCK_COMPRESS;
  SLAB_CUDA_TRY(cudaMemsetAsync(d_r1, 0, NP * sizeof(int32_t), st));
  SLAB_CUDA_TRY(cudaMemcpyAsync(d_r1, data, (size_t)n * 4u, cudaMemcpyHostToDevice, st));
  SLAB_CUDA_TRY(cudaMemcpyAsync(d_tab, h, 12, cudaMemcpyHostToDevice, st));
  SLAB_CUDA_TRY(cudaMemsetAsync(d_chan, 0, sizeof(EncChan), st));
  SLAB_CUDA_TRY(cudaMemsetAsync(d_ltd, 0, sizeof(double) * 8, st));
  uint32_t* d_misc = ARENA(uint32_t, EA_MISC, M_COUNT);
  if (!d_misc) return -1;
  SLAB_CUDA_TRY(cudaMemsetAsync(d_misc, 0, sizeof(uint32_t) * M_COUNT, st));
  if (run_longterm(ctx, sh, fft_size, 1u, 1u, n, d_tab, d_tab + 1, d_tab + 2, d_r1, d_ltac, d_chan, d_ltd, d_ltq, d_misc + M_RISK, NULL) != 0) return -1;
  EncChan* h_chan = (EncChan*)(h + 16);
  double* h_ltd = (double*)(h + 64);
  SLAB_CUDA_TRY(cudaMemcpyAsync(h_chan, d_chan, sizeof(EncChan), cudaMemcpyDeviceToHost, st));
  SLAB_CUDA_TRY(cudaMemcpyAsync(h_ltd, d_ltd, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
  SLAB_CUDA_TRY(cudaStreamSynchronize(st));
  *pitch = h_chan->pitch;
  for (uint32_t k = 0; k < taps; k++) coef[k] = h_ltd[k];
  return 0;
}
