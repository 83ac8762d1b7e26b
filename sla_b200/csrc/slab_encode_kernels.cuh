/*
 * slab_encode_kernels.cuh - the encoder's device code (reference: src/SLAEncoder.c:356-932,
 * src/SLAPredictor.c:189-468,557-607,791-980,1031-1108,1202-1331,1521-1705, src/SLACoder.c:45-82,
 * 120-138,224-270,361-467, src/SLAUtility.c:322-339,370-412,487-696).
 *
 *   E0  k_enc_scan          OR mask of every sample (-> offset_lshift) + per-1024-chunk non-zero flags
 *   E2  k_enc_segments      segment start chain incl. the leading-silence rule (one warp, speculative)
 *   E3a k_enc_lagsums       per segment x channel: exact integer lag sums per 1024-chunk + boundary terms
 *   E3b k_enc_edges         per (segment, edge): autocorrelation from chunk sums, Levinson, code length
 *   E3c k_enc_dijkstra      per segment: shortest path -> block partition
 *   E3d k_enc_fill_blocks   block table (after a scan of partition counts)
 *   E4  k_enc_analysis      per block x channel: window, pre-emphasis, autocorrelation, Levinson,
 *                           RAW decision, bit width, coefficient quantisation
 *   E5  k_enc_parcor        time-parallel int32 pre-emphasis + PARCOR lattice analysis
 *   E6  k_enc_longterm      per block x channel: exact 260-lag autocorrelation, pitch pick, tap solve
 *   E7/8 k_enc_ltlms        per block x channel: long-term FIR + sign-LMS (state in registers)
 *   E9  k_enc_riceprep / k_enc_ricetrace / k_enc_blocksizes / k_enc_pack   entropy coding split into
 *                           parameter trace + code lengths, prefix-scanned offsets, bit packing
 *   E10 k_enc_crc           per-block CRC-16 and size/CRC patch
 */
#ifndef SLAB_ENCODE_KERNELS_CUH
#define SLAB_ENCODE_KERNELS_CUH

#include "slab_common.cuh"

#include <float.h>
#include <math.h>
#include <type_traits>

/* misc[] scalar slots in device memory */
enum {
  M_ORMASK = 0, M_NSEG, M_NBLOCKS, M_TOTAL_BYTES, M_MAX_BLOCK, M_MAX_BPS, M_OVERFLOW, M_CONSUMED, M_RISK, M_LPC_RISK, M_LT_WIDE, M_COUNT = 16
};

#define SLAB_BIGWEIGHT 16777216.0            /* SLAPredictor.c:16 */
#define SLAB_SLICE     256u                  /* samples per thread in the time-parallel lattice */

struct InPtrs { const int32_t* p[SLAB_MAX_CH]; };

struct EncShape {
  uint32_t nch, bits, rate, P, T, lms, ms, window_type, maxblk;
  uint32_t N;                 /* samples per channel in this job */
  uint32_t NP;                /* stride of the intermediate planes: every block starts on a multiple of 8
                                 samples there, so per-thread accesses can be 128-bit */
  uint32_t nnmax;             /* ceil(maxblk / 1024) + 1 */
  uint32_t pstride;           /* P + 1 rounded to the lattice template size + 1 */
  uint32_t lshift;            /* valid after the host read the OR mask */
  uint32_t wide;              /* bits > 24: lag sums in double instead of exact int64 */
  double   ac_scale;          /* 2^-62 * fft_size / 2: scale of the reference's FFT autocorrelation */
  uint32_t out_cap;           /* bytes available for blocks */
  const uint32_t* blk_lshift; /* merged multi-file job: offset_lshift of each block's file (else NULL: lshift) */
};

__device__ __forceinline__ uint32_t enc_lshift(const EncShape& sh, uint32_t block)
{
  return sh.blk_lshift != nullptr ? sh.blk_lshift[block] : sh.lshift;
}

/* (shifted, mid/side transformed) integer sample of channel c, SLAEncoder.c:505-517 */
__device__ __forceinline__ int32_t enc_sample(const InPtrs& in, uint32_t c, uint32_t ms, uint32_t shift, size_t n)
{
  if (!ms) return in.p[c][n] >> shift;
  const int32_t l = in.p[0][n] >> shift, r = in.p[1][n] >> shift;
  return (c == 0) ? ((l + r) >> 1) : (l - r);          /* SLAUtility.c:403-404 */
}

/* ------------------------------------------------------------------------------------ E0 */
/* flags[chunk] = bit g set when samples [32g, 32g+32) of the 1024-sample chunk hold a non-zero value
 * in any channel (so flags[chunk] != 0 <=> the chunk is not silent).
 * One warp per chunk: every lane issues all of its 128-bit loads (8 per channel) before it looks at
 * any of them, so a CTA of 8 warps keeps 64 KB per channel pair in flight; the reduction is warp
 * shuffles only - no shared memory, no CTA barrier. */
template <bool VEC>
__global__ void __launch_bounds__(256) k_enc_scan(InPtrs in, uint32_t nch, uint32_t N,
    uint32_t* __restrict__ flags, uint32_t* __restrict__ misc,
    const uint32_t* __restrict__ chunk_file, uint32_t* __restrict__ file_or)   /* merged job: an OR mask per file */
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t nchunks = (N + SLAB_GRID - 1u) / SLAB_GRID;
  uint32_t all = 0;
  /* persistent: the grid is a few CTAs per SM and every warp strides over the chunks */
  for (uint32_t chunk = blockIdx.x * 8u + (threadIdx.x >> 5); chunk < nchunks; chunk += gridDim.x * 8u) {
  const size_t base = (size_t)chunk * SLAB_GRID;
  uint32_t acc = 0, fine = 0;
  if (VEC && base + SLAB_GRID <= N) {
    for (uint32_t c = 0; c < nch; c++) {
      const int4* src = reinterpret_cast<const int4*>(in.p[c] + base) + lane;
      int4 v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = __ldcs(src + 32 * j);       /* read once: streaming */
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t o = (uint32_t)(v[j].x | v[j].y | v[j].z | v[j].w);
        acc |= o;
        if (o) fine |= 1u << (4 * j + (lane >> 3));     /* int4 (32 j + lane) = samples 128 j + 4 lane .. */
      }
    }
  } else {
    for (uint32_t c = 0; c < nch; c++)
      for (uint32_t i = lane; i < SLAB_GRID && base + i < N; i += 32) {
        const uint32_t v = (uint32_t)in.p[c][base + i];
        acc |= v;
        if (v) fine |= 1u << (i >> 5);
      }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    acc |= __shfl_xor_sync(SLAB_FULL_MASK, acc, d);
    fine |= __shfl_xor_sync(SLAB_FULL_MASK, fine, d);
  }
  if (lane == 0) {
    flags[chunk] = fine;
    if (chunk_file != nullptr && acc) atomicOr(&file_or[chunk_file[chunk]], acc);
  }
  all |= acc;
  }
  if (lane == 0 && all) atomicOr(&misc[M_ORMASK], all);
}

/* ------------------------------------------------------------------------------------ E2 */
/* Segment chain of SLAEncoder_EncodeWhole (SLAEncoder.c:846-869) with the leading-silence rule of
 * SLAEncoder_SearchOptimalBlockPartitions (:393-408).  One warp.  Fast path: lanes test 128
 * consecutive grid positions at once from the chunk flags.  A candidate silent start is resolved
 * from the 32-sample group bits of up to 32 chunks in one round trip; samples are only read for the
 * one group that decides. */
__global__ void __launch_bounds__(32) k_enc_segments(InPtrs in, uint32_t nch, uint32_t N, uint32_t maxblk,
    uint32_t first, uint32_t stop,      /* chunk mode: chain starts at `first`, no segment starts at or after `stop` */
    const uint32_t* __restrict__ flags, uint32_t* __restrict__ seg_start, uint32_t* __restrict__ seg_len,
    uint32_t* __restrict__ seg_kind, uint32_t* __restrict__ misc,
    const uint32_t* __restrict__ file_tab, uint32_t* __restrict__ file_nseg)
{
  const uint32_t lane = threadIdx.x;
  if (file_tab != nullptr) {
    /* merged job: one warp per file, chain over [start, start + len), segments into the file's own slots
     * (start | len | first slot); k_enc_compact_segments closes the gaps */
    first = file_tab[3u * blockIdx.x];
    N = first + file_tab[3u * blockIdx.x + 1u];
    stop = N;
    const uint32_t slot0 = file_tab[3u * blockIdx.x + 2u];
    seg_start += slot0; seg_len += slot0; seg_kind += slot0;
  }
  const uint32_t nchunks = (N + SLAB_GRID - 1) / SLAB_GRID;
  uint64_t s = first;
  uint32_t count = 0;
  while (s < stop) {
    /* 128 grid positions per round trip: four independent flag loads per lane */
    uint32_t m4[4];
    uint64_t sk4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      sk4[j] = s + (uint64_t)(lane + 32u * (uint32_t)j) * maxblk;
      m4[j] = 0;
      if (sk4[j] < stop && (uint64_t)N - sk4[j] >= SLAB_MIN_BLOCK) {
        const uint64_t c1 = (sk4[j] + SLAB_GRID - 1) / SLAB_GRID;   /* aligned chunk inside [sk, sk + 2048) */
        m4[j] = flags[c1];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) m4[j] = __ballot_sync(SLAB_FULL_MASK, m4[j] != 0);
    uint32_t f = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (f == 32u * (uint32_t)j) f += (m4[j] == 0xffffffffu) ? 32u : (uint32_t)(__ffs((int)~m4[j]) - 1);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t idx = lane + 32u * (uint32_t)j;
      if (idx < f) {
        const uint32_t left = (uint32_t)(N - sk4[j]);
        seg_start[count + idx] = (uint32_t)sk4[j];
        seg_len[count + idx] = left < maxblk ? left : maxblk;
        seg_kind[count + idx] = 0;
      }
    }
    count += f;
    s += (uint64_t)f * maxblk;
    if (f == 128u || s >= stop) continue;
    /* exact: z = offset of the first non-zero sample in [s, s + seglen), or seglen */
    const uint32_t left = (uint32_t)(N - s);
    const uint32_t seglen = left < maxblk ? left : maxblk;
    const uint32_t minb = left < SLAB_MIN_BLOCK ? left : SLAB_MIN_BLOCK;
    const uint64_t end = s + seglen;
    const uint32_t c_first = (uint32_t)(s / SLAB_GRID);
    /* lane L owns chunk c_first + L (a segment spans at most 17 chunks) */
    const uint32_t my_chunk = c_first + lane;
    uint32_t word = 0;
    if (my_chunk < nchunks && (uint64_t)my_chunk * SLAB_GRID < end) word = flags[my_chunk];
    if (lane == 0) {                                  /* groups that end at or before s */
      const uint32_t g = (uint32_t)(s % SLAB_GRID) >> 5;
      word &= ~((g == 0) ? 0u : ((1u << g) - 1u));
    }
    uint32_t z = seglen;
    for (int guard = 0; guard < 1024; guard++) {
      uint64_t cand = ~0ull;                          /* start of my first flagged group */
      if (word) cand = (uint64_t)my_chunk * SLAB_GRID + 32u * (uint32_t)(__ffs((int)word) - 1);
      uint64_t best = cand;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) { const uint64_t o = __shfl_xor_sync(SLAB_FULL_MASK, best, d); best = o < best ? o : best; }
      if (best == ~0ull || best >= end) break;
      const uint64_t pos = best + lane;
      int nz = 0;
      if (pos >= s && pos < end)
        for (uint32_t c = 0; c < nch; c++) nz |= (in.p[c][pos] != 0);
      const uint32_t b = __ballot_sync(SLAB_FULL_MASK, nz);
      if (b) { z = (uint32_t)(best + (uint32_t)(__ffs((int)b) - 1) - s); break; }
      if (cand == best) word &= word - 1u;            /* that group had nothing inside the range */
    }
    if (lane == 0) {
      seg_start[count] = (uint32_t)s;
      if (z >= minb) { seg_len[count] = z; seg_kind[count] = 1; }
      else { seg_len[count] = seglen; seg_kind[count] = 0; }
    }
    s += (z >= minb) ? z : seglen;
    count += 1;
  }
  if (lane == 0) {
    if (file_nseg != nullptr) file_nseg[blockIdx.x] = count;
    else { misc[M_NSEG] = count; misc[M_CONSUMED] = (uint32_t)(s < N ? s : N); }
  }
}

/* merged job: the file of every 1024-sample chunk, from the file table (start at tab[stride * f]; starts are
 * ascending multiples of 1024; the chunks before the first start belong to file 0) */
__global__ void __launch_bounds__(128) k_enc_chunk_map(const uint32_t* __restrict__ tab, uint32_t stride, uint32_t nfiles,
    uint32_t nchunks, uint32_t* __restrict__ chunk_file)
{
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nfiles) return;
  const uint32_t c0 = (f == 0u) ? 0u : (tab[stride * f] >> 10);
  const uint32_t c1 = (f + 1u < nfiles) ? (tab[stride * (f + 1u)] >> 10) : nchunks;
  for (uint32_t ch = c0; ch < c1; ch++) chunk_file[ch] = f;
}

/* merged job: the segments of file f move from its slots to [file_seg0[f], file_seg0[f] + file_nseg[f]) */
__global__ void __launch_bounds__(128) k_enc_compact_segments(const uint32_t* __restrict__ file_tab,
    const uint32_t* __restrict__ file_nseg, const uint32_t* __restrict__ file_seg0,
    const uint32_t* __restrict__ slot_start, const uint32_t* __restrict__ slot_len, const uint32_t* __restrict__ slot_kind,
    uint32_t* __restrict__ seg_start, uint32_t* __restrict__ seg_len, uint32_t* __restrict__ seg_kind)
{
  const uint32_t f = blockIdx.x, from = file_tab[3u * f + 2u], to = file_seg0[f], n = file_nseg[f];
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    seg_start[to + i] = slot_start[from + i]; seg_len[to + i] = slot_len[from + i]; seg_kind[to + i] = slot_kind[from + i];
  }
}

/* ------------------------------------------------------------------------------------ E3a */
/* Search-path autocorrelation.  The search data are PCM * 2^-31 (mid = (L+R)/2, side = L-R): dyadic
 * rationals, so lag sums are exact integers and order-independent; every edge's
 * autocorrelation is sum_c P_c(k) - T_j(k) with P_c the lag sums of 1024-sample chunk c and T_j the
 * terms that straddle boundary j.  Output: inclusive chunk prefixes PP[node][k] and TT[node][k]. */
template <bool WIDE, int LG>      /* LG = P + 1 when it is 9, 17 or 33 (register-tiled path), else 0 */
__global__ void __launch_bounds__(256, (LG > 17 ? 1 : 3)) k_enc_lagsums(InPtrs in, EncShape sh,
    const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_len,
    const uint32_t* __restrict__ seg_kind, unsigned long long* __restrict__ PP,
    unsigned long long* __restrict__ TT)
{
  typedef typename std::conditional<WIDE, double, int32_t>::type Y;
  typedef typename std::conditional<WIDE, double, long long>::type A;
  SLAB_DYN_SMEM(unsigned char, smem);
  const uint32_t seg = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
  if (seg_kind[seg] != 0) return;
  const uint32_t L = seg_len[seg], lags = sh.P + 1u;
  const uint32_t nchunks = (L + SLAB_GRID - 1) / SLAB_GRID, nn = nchunks + 1u;
  A* S = reinterpret_cast<A*>(smem);                                /* [nchunks][lags] */
  Y* y = reinterpret_cast<Y*>(smem + sizeof(A) * (size_t)(sh.nnmax - 1u) * lags);
  const size_t s0 = seg_start[seg];
  const uint32_t shift = 32u - sh.bits;
  constexpr bool TILED = !WIDE && LG > 0;
  if (!TILED) {
    /* staging: eight samples per thread and channel are requested before the first one is used */
    for (uint32_t n0 = tid; n0 < L; n0 += 8u * blockDim.x) {
      int32_t a[8], d[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t n = n0 + (uint32_t)j * blockDim.x;
        a[j] = 0; d[j] = 0;
        if (n < L) {
          if (!sh.ms) a[j] = in.p[c][s0 + n];
          else { a[j] = in.p[0][s0 + n]; d[j] = in.p[1][s0 + n]; }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t n = n0 + (uint32_t)j * blockDim.x;
        if (n < L) {
          if (!sh.ms) y[n] = (Y)(a[j] >> shift);
          else {
            const long long l = a[j] >> shift, r = d[j] >> shift;
            y[n] = (Y)((c == 0) ? (l + r) : (l - r));                /* mid kept un-halved: scale 2^-bits */
          }
        }
      }
    }
    for (uint32_t i = tid; i < nchunks * lags; i += blockDim.x) S[i] = (A)0;
    /* zero from the end of the segment to the end of the staging area (maxblk rounded up to 1024, + 128) */
    for (uint32_t i = L + tid; i < ((sh.maxblk + SLAB_GRID - 1u) & ~(SLAB_GRID - 1u)) + 128u; i += blockDim.x) y[i] = (Y)0;
    __syncthreads();
  }
  if (!WIDE && LG > 0) {
    /* register-tiled: a thread owns 64 consecutive samples and all LG lags.  It walks them sixteen at a time
     * with a window of 16 + LG - 1 samples in registers: 16 x LG multiply-adds on compile-time register
     * indices per sixteen shared-memory loads, no predicate anywhere - samples past the segment end are zero
     * in the staging area, so every run and every lag covers the same range.  The sixteen 64-sample runs of
     * a chunk sit in sixteen adjacent lanes: their partial sums meet in a shuffle reduction and the chunk's
     * sums are written once, without atomics. */
    constexpr int T = LG > 0 ? LG : 1;
    constexpr int WIN = 16 + T - 1;
    for (uint32_t t = tid; t < ((nchunks * 16u + 31u) & ~31u); t += blockDim.x) {     /* whole warps take part */
      const uint32_t ch = t >> 4, lo = ch * SLAB_GRID + (t & 15u) * 64u;
      {
        /* the warp stages what it is about to use - its two chunks and 64 samples of look-ahead (the same
         * values the next warp writes there) - so there is no CTA barrier between loading and multiplying:
         * while one warp waits for its loads the others compute.  Eight samples per lane and channel are
         * requested before the first one is used; positions past the segment end become zero. */
        const uint32_t s_lo = ((t & ~31u) >> 4) * SLAB_GRID;
        uint32_t s_hi = s_lo + 2u * SLAB_GRID + 64u;
        if (s_hi > nchunks * SLAB_GRID + 64u) s_hi = nchunks * SLAB_GRID + 64u;
        for (uint32_t n0 = s_lo + (tid & 31u); n0 < s_hi; n0 += 8u * 32u) {
          int32_t a[8], d[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const uint32_t n = n0 + 32u * (uint32_t)j;
            a[j] = 0; d[j] = 0;
            if (n < L) {
              if (!sh.ms) a[j] = in.p[c][s0 + n];
              else { a[j] = in.p[0][s0 + n]; d[j] = in.p[1][s0 + n]; }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const uint32_t n = n0 + 32u * (uint32_t)j;
            if (n < s_hi) {
              if (!sh.ms) y[n] = (Y)(a[j] >> shift);
              else {
                const long long l = a[j] >> shift, r = d[j] >> shift;
                y[n] = (Y)((c == 0) ? (l + r) : (l - r));            /* mid kept un-halved: scale 2^-bits */
              }
            }
          }
        }
        __syncwarp();
      }
      long long acc[T];
#pragma unroll
      for (int k = 0; k < T; k++) acc[k] = 0;
      if (ch < nchunks) {
        int32_t w[WIN];
#pragma unroll
        for (int j = 0; j < T - 1; j++) w[j] = (int32_t)y[lo + j];
#pragma unroll 1
        for (uint32_t i = lo; i < lo + 64u; i += 16u) {
#pragma unroll
          for (int j = T - 1; j < WIN; j++) w[j] = (int32_t)y[i + j];
#pragma unroll
          for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int k = 0; k < T; k++) acc[k] = slab_mad_wide(w[r], w[r + k], acc[k]);
          }
#pragma unroll
          for (int j = 0; j < T - 1; j++) w[j] = w[j + 16];
        }
      }
#pragma unroll
      for (int k = 0; k < T; k++) {
        long long v = acc[k];
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) v += __shfl_xor_sync(SLAB_FULL_MASK, v, d);
        if ((t & 15u) == 0u && ch < nchunks) S[ch * lags + k] = (A)v;
      }
    }
  } else if (!WIDE) {
    const uint32_t total = nchunks * 4u * lags;
    for (uint32_t t = tid; t < total; t += blockDim.x) {
      const uint32_t k = t % lags, cs = t / lags, ch = cs >> 2, sub = cs & 3u;
      const uint32_t lo = ch * SLAB_GRID + sub * 256u;
      uint32_t hi = lo + 256u;
      if (hi + k > L) hi = (L > k) ? L - k : 0u;
      A acc = 0;
      for (uint32_t n = lo; n < hi; n++) acc += (A)y[n] * (A)y[n + k];
      if (acc != 0) atomicAdd(reinterpret_cast<unsigned long long*>(&S[ch * lags + k]), (unsigned long long)acc);
    }
  } else {
    const uint32_t total = nchunks * lags;
    for (uint32_t t = tid; t < total; t += blockDim.x) {
      const uint32_t k = t % lags, ch = t / lags;
      const uint32_t lo = ch * SLAB_GRID;
      uint32_t hi = lo + SLAB_GRID;
      if (hi + k > L) hi = (L > k) ? L - k : 0u;
      A acc = 0;
      for (uint32_t n = lo; n < hi; n++) acc += (A)y[n] * (A)y[n + k];
      S[ch * lags + k] = acc;
    }
  }
  __syncthreads();
  const size_t obase = ((size_t)seg * sh.nch + c) * sh.nnmax * lags;
  for (uint32_t k = tid; k < lags; k += blockDim.x) {
    A run = 0;
    A* pp = reinterpret_cast<A*>(PP + obase);
    pp[k] = run;
    for (uint32_t ch = 0; ch < nchunks; ch++) { run += S[ch * lags + k]; pp[(size_t)(ch + 1u) * lags + k] = run; }
  }
  for (uint32_t t = tid; t < nn * lags; t += blockDim.x) {
    const uint32_t k = t % lags, j = t / lags;
    const uint32_t e = j * SLAB_GRID;
    A acc = 0;
    if (j > 0 && e < L) {
      const uint32_t lo = e - (k < e ? k : e);
      for (uint32_t n = lo; n < e; n++) if (n + k < L) acc += (A)y[n] * (A)y[n + k];
    }
    reinterpret_cast<A*>(TT + obase)[(size_t)j * lags + k] = acc;
  }
}

/* PARCOR from autocorrelation: LPC_CalculateCoef + LPC_LevinsonDurbinRecursion,
 * SLAPredictor.c:217-328, same operation order (u/v vectors folded into one update). */
__device__ inline void enc_levinson(const double* r, uint32_t nsamples, uint32_t order, double* parcor,
                                    double* a, double* t)
{
  for (uint32_t i = 0; i <= order; i++) parcor[i] = 0.0;
  if (nsamples < order) return;
  if (fabs(r[0]) < (double)FLT_EPSILON) return;
  for (uint32_t i = 0; i < order + 2u; i++) a[i] = 0.0;
  a[0] = 1.0;
  a[1] = -r[1] / r[0];
  parcor[1] = r[1] / r[0];
  double e = r[0] + r[1] * a[1];
  for (uint32_t d = 1; d < order; d++) {
    double g = 0.0;
    for (uint32_t i = 0; i < d + 1u; i++) g += a[i] * r[d + 1u - i];
    g /= (-e);
    e = (1.0 - g * g) * e;
    for (uint32_t i = 0; i < d + 2u; i++) t[i] = a[i];
    a[0] = 1.0 + g * 0.0;
    for (uint32_t i = 1; i <= d; i++) a[i] = t[i] + g * t[d + 1u - i];
    a[d + 1u] = 0.0 + g * 1.0;
    parcor[d + 1u] = -g;
  }
}

/* SLALPCCalculator_EstimateCodeLength, SLAPredictor.c:416-468; power = sum of squares (= r[0]) */
__device__ inline double enc_code_length(double power, uint32_t nsamples, uint32_t bits,
                                         const double* parcor, uint32_t order)
{
  double pw = power * exp2((double)(2u * (bits - 1u)));
  if (fabs(pw) <= (double)FLT_MIN) return 0.0;
  pw = log(pw) * 1.4426950408889634 - log((double)nsamples) * 1.4426950408889634;
  double vr = 0.0;
  for (uint32_t k = 1; k <= order; k++) vr += log(1.0 - parcor[k] * parcor[k]) * 1.4426950408889634;
  double len = 1.9426950408889634 + 0.5 * (pw + vr);
  len /= 8;
  if (len <= 0) len = 1.0 / 8;
  return len;
}

/* ------------------------------------------------------------------------------------ E3b */
/* Same recursion with the coefficient update done in place, pair by pair (a[i], a[d+1-i]): every new
 * coefficient is still t[i] + g * t[d+1-i] evaluated once, so the doubles are identical, but the copy
 * of the coefficient vector (half of the recursion's memory traffic) is gone. */
__device__ inline void enc_levinson_inplace(const double* r, uint32_t nsamples, uint32_t order, double* parcor, double* a)
{
  for (uint32_t i = 0; i <= order; i++) parcor[i] = 0.0;
  if (nsamples < order) return;
  if (fabs(r[0]) < (double)FLT_EPSILON) return;
  a[0] = 1.0;
  a[1] = -r[1] / r[0];
  parcor[1] = r[1] / r[0];
  double e = r[0] + r[1] * a[1];
  for (uint32_t d = 1; d < order; d++) {
    double g = 0.0;
    for (uint32_t i = 0; i < d + 1u; i++) g += a[i] * r[d + 1u - i];
    g /= (-e);
    e = (1.0 - g * g) * e;
    uint32_t lo = 1, hi = d;
    for (; lo < hi; lo++, hi--) {
      const double x = a[lo], y = a[hi];
      a[lo] = x + g * y;
      a[hi] = y + g * x;
    }
    if (lo == hi) a[lo] = a[lo] + g * a[lo];
    a[d + 1u] = 0.0 + g * 1.0;
    parcor[d + 1u] = -g;
  }
}

/* E3b, one CTA per segment: the edges the search may use (SLAPredictor.c:1615-1663) are listed first,
 * then every (edge, channel) pair is one work item - no lane idles on the pairs of the node matrix
 * that are not edges - and the per-channel terms are summed in channel order. */
#define EDGE_MAX_NODES 17u                      /* 16384 / 1024 + 1 */
#define EDGE_MAX_PAIRS (EDGE_MAX_NODES * (EDGE_MAX_NODES - 1u) / 2u)
template <bool WIDE>
__global__ void __launch_bounds__(256) k_enc_edges(EncShape sh,
    const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_len,
    const uint32_t* __restrict__ seg_kind, const unsigned long long* __restrict__ PP,
    const unsigned long long* __restrict__ TT, double* __restrict__ adj)
{
  typedef typename std::conditional<WIDE, double, long long>::type A;
  __shared__ uint16_t pair_ij[EDGE_MAX_PAIRS];
  __shared__ double term[EDGE_MAX_PAIRS * SLAB_MAX_CH];
  __shared__ uint32_t npairs;
  const uint32_t seg = blockIdx.x, tid = threadIdx.x;
  if (seg_kind[seg] != 0) return;
  const uint32_t L = seg_len[seg], nn = (L + SLAB_GRID - 1) / SLAB_GRID + 1u, lags = sh.P + 1u;
  /* min(samples left in the file, minimum block): a segment is shorter than the minimum block only when it is
   * all that is left of its file, so its own length decides (a merged job holds many files) */
  const uint32_t minb = L < SLAB_MIN_BLOCK ? L : SLAB_MIN_BLOCK;
  double* out = adj + (size_t)seg * sh.nnmax * sh.nnmax;
  if (tid == 0) npairs = 0;
  __syncthreads();
  for (uint32_t t = tid; t < sh.nnmax * sh.nnmax; t += blockDim.x) {
    const uint32_t i = t / sh.nnmax, j = t % sh.nnmax;
    if (i >= nn || j >= nn) continue;
    out[t] = SLAB_BIGWEIGHT;
    if (j <= i) continue;
    uint32_t len = (j - i) * SLAB_GRID;
    if (len > L - i * SLAB_GRID) len = L - i * SLAB_GRID;
    if (len < minb || len > L) continue;                           /* SLAPredictor.c:1626-1630 */
    pair_ij[atomicAdd(&npairs, 1u)] = (uint16_t)((i << 8) | j);
  }
  __syncthreads();
  const uint32_t np = npairs;
  for (uint32_t w = tid; w < np * sh.nch; w += blockDim.x) {
    const uint32_t e = w / sh.nch, c = w - e * sh.nch;
    const uint32_t i = pair_ij[e] >> 8, j = pair_ij[e] & 0xFFu;
    uint32_t len = (j - i) * SLAB_GRID;
    if (len > L - i * SLAB_GRID) len = L - i * SLAB_GRID;
    double r[SLAB_MAX_PARCOR + 2], a[SLAB_MAX_PARCOR + 2], parcor[SLAB_MAX_PARCOR + 1];
    const size_t base = ((size_t)seg * sh.nch + c) * sh.nnmax * lags;
    const A* pp = reinterpret_cast<const A*>(PP + base);
    const A* tt = reinterpret_cast<const A*>(TT + base);
    /* value = integer * 2^-(bits-1) (plain, side) or integer * 2^-bits (mid) */
    const double scale = exp2(-2.0 * (double)((sh.ms && c == 0) ? sh.bits : sh.bits - 1u));
    for (uint32_t k = 0; k < lags; k++) {
      const A v = pp[(size_t)j * lags + k] - pp[(size_t)i * lags + k] - tt[(size_t)j * lags + k];
      r[k] = (double)v * scale;
    }
    enc_levinson_inplace(r, len, sh.P, parcor, a);
    term[e * sh.nch + c] = len * enc_code_length(r[0], len, sh.bits, parcor, sh.P);
  }
  __syncthreads();
  for (uint32_t e = tid; e < np; e += blockDim.x) {
    double total = 0.0;
    for (uint32_t c = 0; c < sh.nch; c++) total += term[e * sh.nch + c];
    total += 50;       /* SLAPredictor.c:20 */
    total += 300;      /* SLAInternal.h:29  */
    out[(pair_ij[e] >> 8) * sh.nnmax + (pair_ij[e] & 0xFFu)] = total;
  }
}

/* ------------------------------------------------------------------------------------ E3c */
/* SLAOptimalEncodeEstimator_ApplyDijkstraMethod + path unroll, SLAPredictor.c:1521-1581,1671-1695 */
__global__ void __launch_bounds__(64) k_enc_dijkstra(EncShape sh, uint32_t nseg,
    const uint32_t* __restrict__ seg_len, const uint32_t* __restrict__ seg_kind,
    const double* __restrict__ adj, uint32_t* __restrict__ seg_nparts, uint32_t* __restrict__ seg_parts)
{
  const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= nseg) return;
  uint32_t* parts = seg_parts + (size_t)seg * sh.nnmax;
  const uint32_t L = seg_len[seg];
  if (seg_kind[seg] != 0) { seg_nparts[seg] = 1; parts[0] = L; return; }
  const uint32_t nn = (L + SLAB_GRID - 1) / SLAB_GRID + 1u;
  const double* A = adj + (size_t)seg * sh.nnmax * sh.nnmax;
  double cost[SLAB_MAX_NODES];
  uint32_t path[SLAB_MAX_NODES];
  uint8_t done[SLAB_MAX_NODES];
  for (uint32_t i = 0; i < nn; i++) { done[i] = 0; path[i] = 0xFFFFFFFFu; cost[i] = SLAB_BIGWEIGHT; }
  cost[0] = 0.0;
  uint32_t cur = 0;
  for (uint32_t guard = 0; guard <= nn; guard++) {
    double best = SLAB_BIGWEIGHT;
    for (uint32_t i = 0; i < nn; i++) if (!done[i] && cost[i] < best) { best = cost[i]; cur = i; }
    if (cur == nn - 1u) break;
    for (uint32_t i = 0; i < nn; i++) {
      const double via = A[cur * sh.nnmax + i] + cost[cur];
      if (cost[i] > via) { cost[i] = via; path[i] = cur; }
    }
    done[cur] = 1;
  }
  uint32_t hops = 0;
  for (uint32_t node = nn - 1u; node != 0 && node != 0xFFFFFFFFu && hops < nn; node = path[node]) hops++;
  uint32_t node = nn - 1u;
  for (uint32_t i = 0; i < hops; i++) {
    const uint32_t from = path[node];
    uint32_t len = (node - from) * SLAB_GRID;
    if (len > L - from * SLAB_GRID) len = L - from * SLAB_GRID;
    parts[hops - 1u - i] = len;
    node = from;
  }
  seg_nparts[seg] = hops;
}

/* single-CTA exclusive scan of uint32 (counts are small; n is at most a few million) */
__global__ void __launch_bounds__(1024) k_scan_u32(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
    uint32_t n, uint32_t* __restrict__ total_out)
{
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024u) {
    const uint32_t i = base + tid;
    const uint32_t v = (i < n) ? in[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(SLAB_FULL_MASK, x, d);
      if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31u) warp_sum[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(SLAB_FULL_MASK, w, d);
        if (lane >= (uint32_t)d) w += y;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const uint32_t prefix = carry + (wid ? warp_sum[wid - 1u] : 0u) + x - v;
    if (i < n) out[i] = prefix;
    __syncthreads();
    if (tid == 1023u) carry = prefix + v;
    __syncthreads();
  }
  if (tid == 0 && total_out) *total_out = carry;
}

/* ------------------------------------------------------------------------------------ E3d */
__global__ void __launch_bounds__(128) k_enc_fill_blocks(EncShape sh, uint32_t nseg,
    const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_kind,
    const uint32_t* __restrict__ seg_nparts, const uint32_t* __restrict__ seg_parts,
    const uint32_t* __restrict__ seg_blk0, uint32_t* __restrict__ blk_start, uint32_t* __restrict__ blk_len,
    uint32_t* __restrict__ blk_flag)
{
  const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= nseg) return;
  uint32_t pos = seg_start[seg], b = seg_blk0[seg];
  for (uint32_t k = 0; k < seg_nparts[seg]; k++, b++) {
    const uint32_t len = seg_parts[(size_t)seg * sh.nnmax + k];
    blk_start[b] = pos; blk_len[b] = len; blk_flag[b] = seg_kind[seg];
    pos += len;
  }
}

#endif
