/*
 * slab_decode_kernels.cuh - device code shared by the decoder's translation units: the per-lane
 * entropy decoder (DeLane) and the per-lane synthesis cascade (SynthLane).  slab_decode.cu builds the
 * stand-alone kernels from them, slab_decode_fused.cu the fused entropy + synthesis kernel.
 */
#ifndef SLAB_DECODE_KERNELS_CUH
#define SLAB_DECODE_KERNELS_CUH

#include "slab_common.cuh"
#include "slab_ctx.cuh"

/* SLAApiResult values used on the device (SLA.h:26-43) */
#define SLAB_RES_INSUFFICIENT_DATA   9u
#define SLAB_RES_DATA_CORRUPTION     11u
#define SLAB_RES_SYNC_CODE           12u
#define SLAB_RES_INSUFFICIENT_BUFFER 4u

struct DecShape {
  uint32_t nch, bits, lshift, P, T, lms, ms, check_crc;
  uint32_t nblocks, total_samples, stream_size, nwords, pstride;
  uint32_t NP;      /* stride of the work planes: blocks start on multiples of 8 samples there */
};

struct OutPtrs { int32_t* p[SLAB_MAX_CH]; };

/* The reference steps to the next block by the bytes its bit reader consumed (SLADecoder.c:651,717),
 * not by the size field.  They agree on every well-formed stream; when they do not (corrupt data
 * with the CRC check off) the reference loses the sync code at the next block. */
__device__ __forceinline__ void k_dec_check_consumed(const SlabBitReader& br, uint32_t blk_off,
    uint32_t size_field, uint32_t* err)
{
  const uint64_t consumed = br.byte_pos() - blk_off;
  if (consumed != (uint64_t)size_field + 6u && *err == 0) *err = SLAB_RES_SYNC_CODE;
}

/* ------------------------------------------------------------------ D1b: header + entropy decode */
/* One lane per block (the channels of a block share one bit stream, sample-interleaved), one warp per
 * CTA.  The stage is a pure recurrence -
 * where code i + 1 starts is known only after code i has been decoded - so what is tuned here is the
 * length of that dependent chain and the instruction count per code:
 *   window (1 funnel shift) -> leading zeros -> parameter select -> bits used -> advance (add,
 *   compare, predicated rotate) -> next window.
 * The remainder extraction, the value, both running-mean updates and the next exponents hang off the
 * chain.  Escapes (run of 16) and codes longer than 32 bits leave through one rarely taken branch. */
struct DeRiceState { uint32_t p0, p1, k0, k1; };

/* Off the fast path: the escape code (a run of 16, then a gamma code, SLACoder.c:141-162) in two
 * more window steps when the gamma part fits one window, which is every quotient below 2^16; or a
 * code longer than the 32-bit window; or, in damaged streams only, runs that need the loop. */
__device__ __forceinline__ void de_rice_slow(SlabBitReader& br, uint32_t lz, uint32_t k0, uint32_t k1, uint32_t& q, uint32_t& r)
{
  if (lz == 16u) {
    br.advance(17u);
    const uint32_t g = slab_lz_nonzero(br.window());           /* digits after the leading one */
    if (g < 16u) {
      br.advance(g + 1u);
      q = 15u + (1u << g) + br.get(g);
    } else {
      const uint32_t nd = br.zero_run() + 1u;
      q = 16u + (uint32_t)((1ull << ((nd - 1u) & 63u)) + br.get(nd - 1u > 32u ? 32u : nd - 1u) - 1ull);
    }
  } else {
    q = br.zero_run();
    if (q == 16u) {
      const uint32_t nd = br.zero_run() + 1u;
      if (nd > 1u) q += (uint32_t)((1ull << ((nd - 1u) & 63u)) + br.get(nd - 1u > 32u ? 32u : nd - 1u) - 1ull);
    }
  }
  r = br.get(q ? k1 : k0);
}

/* one recursive-Rice code, SLACoder.c:273-318 */
__device__ __forceinline__ uint32_t de_rice_code(SlabBitReader& br, DeRiceState& st)
{
  const uint32_t W = br.window();
  const uint32_t lz = slab_lz_nonzero(W);                 /* 0xffffffff for an all-zero window */
  const uint32_t k = lz ? st.k1 : st.k0;
  const uint32_t used = lz + 1u + k;
  uint32_t q, r;
  if (__builtin_expect(lz < 16u && used <= 32u, 1)) {
    q = lz;
    r = slab_shr_c(W << (lz + 1u), 32u - k);              /* k == 0 -> 0 */
    br.advance(used);
  } else {
    de_rice_slow(br, lz, st.k0, st.k1, q, r);
  }
  const uint32_t tail = ((q - 1u) << st.k1) + r;
  const uint32_t v = q ? (1u << st.k0) + tail : r;
  const uint32_t p1n = slab_rice_update32(st.p1, tail);
  st.p0 = slab_rice_update32(st.p0, v);
  st.p1 = q ? p1n : st.p1;
  st.k0 = slab_rice_k32(st.p0);
  st.k1 = slab_rice_k32(st.p1);
  return v;
}

/* one fixed-parameter Golomb code, SLACoder.c:85-117 */
__device__ __forceinline__ uint32_t de_golomb_code(SlabBitReader& br, uint32_t mm)
{
  const uint32_t q = br.zero_run();
  if ((mm & (mm - 1u)) == 0) return q * mm + br.get(slab_log2ceil(mm));
  const uint32_t bb = slab_log2ceil(mm); const uint32_t cut = (1u << bb) - mm;
  uint32_t rest = br.get(bb - 1u);
  if (rest >= cut) rest = ((rest << 1) + br.get(1)) - cut;
  return q * mm + rest;
}

/* Per-lane entropy state: one block's bit reader and the coder state of its channels.  Shared by
 * the stand-alone entropy kernel (k_dec_entropy) and the fused entropy + synthesis kernel
 * (k_dec_block), which differ only in where the decoded residuals go. */
enum { DE_IDLE = 0, DE_RICE, DE_GOLOMB, DE_RAW };

struct DeOutArrays {
  uint32_t* type; int32_t* kq; int32_t* ltq; uint32_t* pitch; uint32_t* err;
};

template <int NCH>
struct DeLane {
  SlabBitReader br;
  DeRiceState st[NCH];
  uint32_t aux[NCH];          /* fixed Golomb parameter, or raw bit width */
  uint32_t n, mode, size_field, blk_off, b;

  /* block header, SLADecoder.c:309-420; leaves mode = DE_IDLE for silent and unusable blocks */
  __device__ __forceinline__ void begin(unsigned char* ring, const uint32_t* words, const DecShape& sh, uint32_t block,
      const uint32_t* __restrict__ blk_off_in, const uint32_t* __restrict__ blk_n, const DeOutArrays& o)
  {
    b = block; mode = DE_IDLE; n = 0;
    blk_off = blk_off_in[b];
    br.init(ring, words, sh.nwords, blk_off);
    const uint32_t sync = br.get(16);
    size_field = br.get(32);
    (void)br.get(16);
    const uint32_t nn = br.get(16);
    const uint32_t type = br.get(2);
    o.type[b] = type;
    if (sync != 0xFFFFu) { o.err[b] = SLAB_RES_SYNC_CODE; return; }
    if (nn != blk_n[b] || type > SLAB_BLOCK_RAW) { if (o.err[b] == 0) o.err[b] = SLAB_RES_DATA_CORRUPTION; return; }
    n = nn;
    if (type == SLAB_BLOCK_COMPRESS) {
#pragma unroll 1
      for (int c = 0; c < NCH; c++) {
        const uint32_t bc = b * NCH + c;
        br.topup();                                              /* one channel header: < 180 bytes */
        const uint32_t rsh = br.get(4);
        int32_t* kq = o.kq + (size_t)bc * sh.pstride;
        kq[0] = 0;
#pragma unroll 1
        for (uint32_t k = 1; k <= sh.P; k++) {
          const uint32_t qb = (k < 4u) ? 16u : 8u;              /* SLAInternal.h:38 */
          const int32_t q = slab_unzigzag(br.get(qb));
          kq[k] = (int32_t)((uint32_t)q << (16u - qb)) >> rsh;  /* SLADecoder.c:384-389 */
        }
#pragma unroll 1
        for (uint32_t k = sh.P + 1; k < sh.pstride; k++) kq[k] = 0;
        uint32_t pitch = 0;
        if (br.get(1)) {
          pitch = br.get(10);
#pragma unroll 1
          for (uint32_t k = 0; k < sh.T; k++)
            o.ltq[(size_t)bc * 8 + k] = (int32_t)((uint32_t)slab_unzigzag(br.get(16)) << 16);
        }
        o.pitch[bc] = pitch;
        const uint32_t init = br.get(sh.bits) << 8;              /* SLACoder.c:18-20: 32-bit shift */
#pragma unroll
        for (int cc = 0; cc < NCH; cc++)                         /* static register index */
          if (cc == c) { st[cc].p0 = st[cc].p1 = init; }
      }
    }
    br.align_byte();
    if (type == SLAB_BLOCK_SILENT) { n = 0; finish(o); return; }
    if (type == SLAB_BLOCK_RAW) {
#pragma unroll
      for (int c = 0; c < NCH; c++) aux[c] = sh.bits - sh.lshift + ((c == 1 && sh.ms) ? 1u : 0u);
      mode = DE_RAW;
      return;
    }
    uint64_t avg = 0;
#pragma unroll
    for (int c = 0; c < NCH; c++) avg += slab_rice_param(st[c].p0);
    avg /= NCH;
    if (avg > 8) {                                               /* SLACoder.c:491 */
#pragma unroll
      for (int c = 0; c < NCH; c++) { st[c].k0 = slab_rice_k32(st[c].p0); st[c].k1 = st[c].k0; }
      mode = DE_RICE;
    } else {
#pragma unroll
      for (int c = 0; c < NCH; c++) aux[c] = slab_rice_param(st[c].p0);
      mode = DE_GOLOMB;
    }
  }

  /* decode samples [s0, s1) of every channel; sink.put(c, s, value).  PER = samples between two
   * top-ups of the ring (at most 32 codes). */
  template <class Sink>
  __device__ __forceinline__ void span(uint32_t s0, uint32_t s1, Sink& sink)
  {
    constexpr uint32_t PER = (NCH == 1) ? 32u : (NCH == 2) ? 16u : (NCH <= 4) ? 8u : 4u;
#define SLAB_DECODE_SPAN(DECODE_ONE)                                                               \
    _Pragma("unroll 1") for (uint32_t i = s0; i < s1; i += PER) {                                  \
      br.topup();                                                                                  \
      const uint32_t end = (s1 - i < PER) ? s1 : i + PER;                                          \
      _Pragma("unroll 1") for (uint32_t s = i; s < end; s++) {                                     \
        _Pragma("unroll") for (int c = 0; c < NCH; c++) {                                          \
          DECODE_ONE;                                                                              \
          sink.put(c, s, slab_unzigzag(v));                                                        \
        }                                                                                          \
      }                                                                                            \
    }
    if (mode == DE_RICE) { SLAB_DECODE_SPAN(const uint32_t v = de_rice_code(br, st[c])) }
    else if (mode == DE_GOLOMB) { SLAB_DECODE_SPAN(const uint32_t v = de_golomb_code(br, aux[c])) }
    else if (mode == DE_RAW) { SLAB_DECODE_SPAN(const uint32_t v = br.get(aux[c])) }
#undef SLAB_DECODE_SPAN
  }

  __device__ __forceinline__ void finish(const DeOutArrays& o)
  {
    k_dec_check_consumed(br, blk_off, size_field, &o.err[b]);
  }
};

struct DeGlobalSink {
  int32_t* base; size_t stride;
  __device__ __forceinline__ void put(int c, uint32_t s, int32_t v) const { base[(size_t)c * stride + s] = v; }
};

/* ------------------------------------------------------------------ D2: synthesis cascade */
/* One thread per block x channel; LMS -> long-term -> PARCOR -> de-emphasis fused per sample with all
 * filter state in registers.  Samples go in chunks of LMS_N: the chunk's residuals and the long-term
 * history are loaded up front so that their latency overlaps; the LMS delay lines are ring buffers
 * indexed at compile time after unrolling.  The main loop covers whole chunks without per-sample
 * bounds checks; the priming chunk, the tail and pitch lags shorter than a chunk go through a checked
 * variant of the same code. */
template <int LMS_N>
struct LmsRing {
  int32_t cx[LMS_N], cp[LMS_N], hx[LMS_N], hp[LMS_N], sx[LMS_N], sp[LMS_N];
};

/* one sign-LMS synthesis step at ring slot U (SLAPredictor.c:1390-1453): returns the output sample */
template <int LMS_N>
__device__ __forceinline__ int32_t lms_synth_step(LmsRing<LMS_N>& st, const int U, int32_t resid)
{
  uint32_t a0 = 1u << 9, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
  for (int i = 0; i < LMS_N; i += 2) {
    a0 += (uint32_t)st.cx[i] * (uint32_t)st.hx[(U - 1 - i + 2 * LMS_N) % LMS_N];
    a1 += (uint32_t)st.cp[i] * (uint32_t)st.hp[(U - 1 - i + 2 * LMS_N) % LMS_N];
    a2 += (uint32_t)st.cx[i + 1] * (uint32_t)st.hx[(U - 2 - i + 2 * LMS_N) % LMS_N];
    a3 += (uint32_t)st.cp[i + 1] * (uint32_t)st.hp[(U - 2 - i + 2 * LMS_N) % LMS_N];
  }
  const int32_t pred = (int32_t)((a0 + a1) + (a2 + a3)) >> 10;
  const int32_t v = (int32_t)((uint32_t)resid + (uint32_t)pred);
  const uint32_t mag = (resid < 0) ? (0u - (uint32_t)resid) : (uint32_t)resid;
  const int32_t step = slab_sgn(resid) * (int32_t)(slab_bitlen(mag) >> 1);
#pragma unroll
  for (int i = 0; i < LMS_N; i++) {
    st.cx[i] += step * st.sx[(U - 1 - i + 2 * LMS_N) % LMS_N];
    st.cp[i] += step * st.sp[(U - 1 - i + 2 * LMS_N) % LMS_N];
  }
  st.hx[U] = v; st.hp[U] = pred; st.sx[U] = slab_sgn(v); st.sp[U] = slab_sgn(pred);
  return v;
}

/* PARCOR lattice synthesis of one sample (SLAPredictor.c:722-736), zero-padded to PMAX stages.  The
 * products that feed the forward chain only need the previous sample's backward errors, so they are
 * all issued first; the chain itself is PMAX dependent adds. */
template <int PMAX>
__device__ __forceinline__ int32_t parcor_synth_step(const int32_t* kk, int32_t* bw, int32_t in)
{
  int32_t t[PMAX + 1], fs[PMAX + 1];
#pragma unroll
  for (int m = 1; m <= PMAX; m++) t[m] = slab_latmul(kk[m], bw[m - 1]);
  fs[PMAX] = in + t[PMAX];
#pragma unroll
  for (int m = PMAX - 1; m >= 1; m--) fs[m] = fs[m + 1] + t[m];
  /* fs[m] = forward error after stage m; b[m] = b[m-1](old) - k[m] * fs[m] */
#pragma unroll
  for (int m = PMAX; m >= 1; m--) bw[m] = bw[m - 1] - slab_latmul(kk[m], fs[m]);
  bw[0] = fs[1];
  return fs[1];
}

template <int LMS_N, int PMAX, int TAPS, bool CHECKED>
__device__ __forceinline__ void synth_chunk(LmsRing<LMS_N>& st, const int32_t* kk, int32_t* bw,
    const int32_t* ltc, int32_t& emph_prev, int32_t* x, int32_t* lt_hist, uint32_t s0, uint32_t n,
    uint32_t delay, bool use_lt, bool lt_far, bool prime, bool filter, const int32_t* tile_src)
{
  int32_t rin[LMS_N], hist[LMS_N + TAPS - 1], lto[LMS_N], res[LMS_N];
  if (tile_src != nullptr) {
    /* fused kernel: the chunk's residuals sit in this lane's row of the shared-memory tile (explicit
     * shared-space loads: through the generic pointer they would be generic LDs with global-load
     * latency on the long scoreboard) */
#ifdef SLAB_EMUL
#pragma unroll
    for (int u = 0; u < LMS_N; u++) rin[u] = tile_src[u];
#else
    const uint32_t taddr = (uint32_t)__cvta_generic_to_shared(tile_src);
#pragma unroll
    for (int u = 0; u < LMS_N; u++) asm volatile("ld.shared.s32 %0, [%1];" : "=r"(rin[u]) : "r"(taddr + 4u * (uint32_t)u));
#endif
  } else {
    /* 128-bit accesses: the block's slot in the work planes is 32-byte aligned and padded to a
     * multiple of 8 samples, so a chunk never leaves it */
    const int4* xv = reinterpret_cast<const int4*>(x);
#pragma unroll
    for (int q = 0; q < LMS_N / 4; q++) {
      const int4 t = xv[(s0 >> 2) + q];
      rin[4 * q] = t.x; rin[4 * q + 1] = t.y; rin[4 * q + 2] = t.z; rin[4 * q + 3] = t.w;
    }
  }
#pragma unroll
  for (int u = 0; u < LMS_N + TAPS - 1; u++) hist[u] = 0;
  if (use_lt && lt_far) {
#pragma unroll
    for (int u = 0; u < LMS_N + TAPS - 1; u++) {
      const uint32_t idx = s0 + (uint32_t)u;
      hist[u] = (idx >= delay) ? lt_hist[idx - delay] : 0;
    }
  }
#pragma unroll
  for (int u = 0; u < LMS_N; u++) {
    const uint32_t s = s0 + (uint32_t)u;
    const int32_t resid = rin[u];
    int32_t v = resid;
    if (filter) {
      if (CHECKED && prime) { st.hx[u] = st.hp[u] = resid; st.sx[u] = st.sp[u] = slab_sgn(resid); }
      else v = lms_synth_step<LMS_N>(st, u, resid);
    }
    if (use_lt) {                                   /* SLAPredictor.c:1031-1108, recursive on its own output */
      if (s >= delay && (!CHECKED || s < n)) {
        long long acc = 1ll << 30;
        if (lt_far) {
#pragma unroll
          for (int j = 0; j < TAPS; j++) acc = slab_mad_wide(ltc[j], hist[u + j], acc);
        } else {
#pragma unroll
          for (int j = 0; j < TAPS; j++) acc = slab_mad_wide(ltc[j], lt_hist[s - delay + j], acc);
        }
        v = (int32_t)((uint32_t)v + (uint32_t)(int32_t)(acc >> 31));
      }
      if (lt_far) lto[u] = v;
      else if (!CHECKED || s < n) lt_hist[s] = v;
    }
    int32_t f = parcor_synth_step<PMAX>(kk, bw, v);
    f = (int32_t)((uint32_t)f + (uint32_t)slab_emph(emph_prev));       /* SLAPredictor.c:1781-1786 */
    emph_prev = f;
    res[u] = f;
  }
  int4* ov = reinterpret_cast<int4*>(x);
  int4* hv = reinterpret_cast<int4*>(lt_hist);
#pragma unroll
  for (int q = 0; q < LMS_N / 4; q++) {
    ov[(s0 >> 2) + q] = make_int4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
    if (use_lt && lt_far) hv[(s0 >> 2) + q] = make_int4(lto[4 * q], lto[4 * q + 1], lto[4 * q + 2], lto[4 * q + 3]);
  }
}

/* Per-lane synthesis state of one block x channel (everything in registers). */
template <int LMS_N, int PMAX, int TAPS>
struct SynthLane {
  int32_t kk[PMAX + 1], bw[PMAX + 1], ltc[TAPS];
  LmsRing<LMS_N> st;
  int32_t emph_prev;
  int32_t* x; int32_t* lt_hist;
  uint32_t n, delay;
  bool use_lt, lt_far, filter;

  __device__ __forceinline__ void begin(const DecShape& sh, uint32_t bc, uint32_t nsamp, int32_t* work_row, int32_t* hist_row,
      const int32_t* kq_in, const int32_t* ltq_in, const uint32_t* pitch_in)      /* coherent loads: see k_dec_block */
  {
    n = nsamp; x = work_row; lt_hist = hist_row;
#pragma unroll
    for (int m = 0; m <= PMAX; m++) { kk[m] = kq_in[(size_t)bc * sh.pstride + m]; bw[m] = 0; }
    const uint32_t pitch = pitch_in[bc];
    delay = pitch + (sh.T >> 1);
    use_lt = pitch != 0;
    lt_far = delay >= (uint32_t)LMS_N + sh.T - 1u;         /* taps never reach into the chunk */
#pragma unroll
    for (int j = 0; j < TAPS; j++) ltc[j] = (use_lt && (uint32_t)j < sh.T) ? ltq_in[(size_t)bc * 8 + j] : 0;
#pragma unroll
    for (int i = 0; i < LMS_N; i++) { st.cx[i] = st.cp[i] = 0; st.hx[i] = st.hp[i] = st.sx[i] = st.sp[i] = 0; }
    emph_prev = 0;
    filter = n > (uint32_t)LMS_N;          /* SLAPredictor.c:1366-1387: short blocks pass through */
  }
  /* samples [s0, s0 + LMS_N): the priming chunk, whole chunks without bounds checks, checked tail */
  __device__ __forceinline__ void chunk(uint32_t s0, const int32_t* tile_src)
  {
    if (s0 >= n) return;
    if (s0 == 0)
      synth_chunk<LMS_N, PMAX, TAPS, true>(st, kk, bw, ltc, emph_prev, x, lt_hist, 0, n, delay, use_lt, lt_far, true, filter, tile_src);
    else if (s0 + LMS_N <= n && (!use_lt || lt_far))
      synth_chunk<LMS_N, PMAX, TAPS, false>(st, kk, bw, ltc, emph_prev, x, lt_hist, s0, n, delay, use_lt, true, false, filter, tile_src);
    else
      synth_chunk<LMS_N, PMAX, TAPS, true>(st, kk, bw, ltc, emph_prev, x, lt_hist, s0, n, delay, use_lt, lt_far, false, filter, tile_src);
  }
};

/* slab_decode_fused.cu: 1 = parameter set not covered, 0 = launched, -1 = CUDA error */
int slab_decode_fused(SlabCtx* ctx, const DecShape& sh, int pmax, const uint32_t* words, const uint32_t* blk_off,
    const uint32_t* blk_pst, const uint32_t* blk_n, int32_t* work, int32_t* scratch, uint32_t* type, int32_t* kq,
    int32_t* ltq, uint32_t* pitch, uint32_t* err);

#endif
