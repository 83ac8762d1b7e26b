/*
 * slab_decode.cu - SLADecoder_DecodeWhole on the GPU (reference: src/SLADecoder.c:309-732,
 * src/SLACoder.c:85-162,273-318,470-506, src/SLAPredictor.c:722-736,1031-1108,1334-1463,1768-1791).
 *
 * Blocks are self-contained (all predictor/coder state resets at a block start,
 * SLADecoder.c:569-581), so the unit of parallelism is the block for the entropy stage - channels
 * share one bitstream, sample-interleaved - and block x channel for the synthesis cascade.
 *
 *   D0  k_dec_walk      block chain -> offset table (only when the caller has no host copy)
 *   D1a k_dec_crc       one warp per block: sliced CRC-16 combined with x^(8n) mod P
 *   D1b k_dec_entropy   one thread per block: block header + Rice/Golomb/raw decode
 *   D2  k_dec_synth     one thread per block x channel: LMS -> long-term -> PARCOR -> de-emphasis,
 *                       all filter state in registers
 *   D3  k_dec_output    streaming: MS->LR, left shift, store planar int32
 */
#include "slab_decode_kernels.cuh"

#include <string.h>

enum {
  DA_STREAM = 0, DA_BLK_OFF, DA_BLK_SMP, DA_BLK_N, DA_WORK, DA_OUT, DA_TYPE, DA_KQ, DA_LTQ, DA_PITCH,
  DA_ERR, DA_COUNTERS, DA_BLK_PST,
  DA_W_POS, DA_W_NEXT, DA_W_N, DA_W_J0, DA_W_J1, DA_W_ORD, DA_W_HKEY, DA_W_HVAL, DA_W_BNEXT
};

/* ------------------------------------------------------------------ D0: device-side chain walk */
/* counters[0] = blocks, counters[1] = samples, counters[2] = error code, counters[3] = padded samples,
 * counters[6] = largest block (samples per channel) on the chain */
__global__ void k_dec_walk(const uint8_t* stream, uint32_t stream_size, uint32_t max_samples,
                           uint32_t max_blocks, uint32_t* blk_off, uint32_t* blk_smp, uint32_t* blk_n,
                           uint32_t* blk_pst, uint32_t* counters)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint32_t off = 43, smp = 0, nb = 0, err = 0, padded = 0, maxn = 0;
  while (smp < max_samples && nb < max_blocks) {
    if (off > stream_size || stream_size - off < 11u) { err = SLAB_RES_INSUFFICIENT_DATA; break; }
    const uint8_t* b = stream + off;
    if (b[0] != 0xFF || b[1] != 0xFF) { err = SLAB_RES_SYNC_CODE; break; }
    uint32_t size = (((uint32_t)b[2] << 24) | ((uint32_t)b[3] << 16) | ((uint32_t)b[4] << 8) | b[5]) + 6u;
    uint32_t n = ((uint32_t)b[8] << 8) | b[9];
    /* a size field of 0xFFFFFFFA.. wraps: same rule as the host walk (block >= 10 bytes, inside the stream) */
    if (size > stream_size - off || size < 10u) { err = SLAB_RES_INSUFFICIENT_DATA; break; }
    if (n > max_samples - smp) { err = SLAB_RES_INSUFFICIENT_BUFFER; break; }
    blk_off[nb] = off; blk_smp[nb] = smp; blk_n[nb] = n; blk_pst[nb] = padded;
    nb++; smp += n; off += size; padded += (n + 7u) & ~7u;
    maxn = n > maxn ? n : maxn;
  }
  counters[0] = nb; counters[1] = smp; counters[2] = err; counters[3] = padded; counters[6] = maxn;
}

/* ------------------------------------------------------------------ D0 in parallel */
/* The serial walk above pays one DRAM round trip per block (12 920 dependent misses on a 1-hour
 * file).  The parallel form: (1) k_dec_findsync scans the whole stream once for positions that look
 * like a block start exactly as the walk would accept them (sync code, room for the header, size field
 * inside the stream) and files them in a hash table keyed by position; (2) k_dec_chain, one CTA,
 * resolves every candidate's successor through the table, then marks the candidates reachable from
 * byte 43 by pointer doubling - round r pushes the mark 2^r blocks ahead and squares the jump table,
 * and the mark itself is the block's ordinal - and finally applies the walk's stopping rules
 * (sample budget, block budget, error at the first position that is not a candidate).
 * Random data contains the sync pattern once per 64 KiB, so the candidate set is the true blocks plus
 * a few thousand impostors that the chain never reaches. */
#define DW_NONE 0xFFFFFFFFu

struct DecWalk {
  uint32_t cap;            /* candidate capacity */
  uint32_t hmask;          /* hash table size - 1 (power of two, at least 2 * cap) */
  uint32_t* pos; uint32_t* next; uint32_t* nsmp;      /* per candidate */
  uint32_t* jmp0; uint32_t* jmp1; uint32_t* ord;      /* per candidate */
  uint32_t* hkey; uint32_t* hval;                     /* hash table, zero = empty */
  uint32_t* bnext;         /* per block: where the following block starts */
};

__device__ __forceinline__ uint32_t dw_hash(uint32_t pos, uint32_t mask) { return (pos * 2654435761u >> 7) & mask; }

/* counters[4] = candidates found, counters[5] = overflow flag */
__global__ void __launch_bounds__(256) k_dec_findsync(const uint8_t* __restrict__ stream, uint32_t stream_size,
    DecWalk w, uint32_t* __restrict__ counters)
{
  const uint32_t groups = (stream_size + 15u) >> 4;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
    const uint4 v = *reinterpret_cast<const uint4*>(stream + (size_t)g * 16u);   /* image is padded */
    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
    /* bit i of m: byte i of the group is 0xFF */
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t x = wv[k];
      const uint32_t t = x & (x >> 4) & 0x0F0F0F0Fu;           /* low nibble of each byte: both nibbles 0xF */
      const uint32_t f = t & (t >> 2) & 0x03030303u;
      const uint32_t e = f & (f >> 1) & 0x01010101u;
      m |= ((e | (e >> 7) | (e >> 14) | (e >> 21)) & 0xFu) << (4 * k);   /* gather the four byte flags */
    }
    if (m == 0u) continue;
    const uint32_t nxt = stream[(size_t)g * 16u + 16u] == 0xFFu ? 1u : 0u;
    uint32_t pairs = m & ((m >> 1) | (nxt << 15));
    while (pairs) {
      const uint32_t i = (uint32_t)__ffs((int)pairs) - 1u;
      pairs &= pairs - 1u;
      const uint32_t p = g * 16u + i;
      if (p < 43u || p > stream_size || stream_size - p < 11u) continue;
      const uint8_t* b = stream + p;
      const uint32_t size = (((uint32_t)b[2] << 24) | ((uint32_t)b[3] << 16) | ((uint32_t)b[4] << 8) | b[5]) + 6u;
      if (size > stream_size - p || size < 10u) continue;        /* the walk's rule: >= 10 bytes, no wrap */
      const uint32_t idx = atomicAdd(&counters[4], 1u);
      if (idx >= w.cap) { counters[5] = 1u; continue; }
      w.pos[idx] = p; w.next[idx] = p + size; w.nsmp[idx] = ((uint32_t)b[8] << 8) | b[9];
      uint32_t slot = dw_hash(p, w.hmask);
      while (atomicCAS(&w.hkey[slot], 0u, p) != 0u) slot = (slot + 1u) & w.hmask;
      w.hval[slot] = idx;
    }
  }
}

__device__ __forceinline__ uint32_t dw_lookup(const DecWalk& w, uint32_t pos)
{
  uint32_t slot = dw_hash(pos, w.hmask);
  for (;;) {
    const uint32_t k = w.hkey[slot];
    if (k == pos) return w.hval[slot];
    if (k == 0u) return DW_NONE;
    slot = (slot + 1u) & w.hmask;
  }
}

/* exclusive scan of one value per thread over the CTA (1024 threads), returns the CTA total */
__device__ __forceinline__ uint32_t dw_block_scan(uint32_t v, uint32_t* warp_sum, uint32_t& excl)
{
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(SLAB_FULL_MASK, x, d);
    if (lane >= (uint32_t)d) x += y;
  }
  if (lane == 31u) warp_sum[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t t = warp_sum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(SLAB_FULL_MASK, t, d);
      if (lane >= (uint32_t)d) t += y;
    }
    warp_sum[lane] = t;
  }
  __syncthreads();
  excl = (wid ? warp_sum[wid - 1u] : 0u) + x - v;
  const uint32_t total = warp_sum[31];
  __syncthreads();
  return total;
}

/* one CTA of 1024 threads; writes the same four counters as k_dec_walk */
__global__ void __launch_bounds__(1024) k_dec_chain(const uint8_t* __restrict__ stream, uint32_t stream_size,
    uint32_t max_samples, uint32_t max_blocks, DecWalk w, uint32_t* blk_off, uint32_t* blk_smp,
    uint32_t* blk_n, uint32_t* blk_pst, uint32_t* counters)
{
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t s_len, s_end, s_carry_s, s_carry_p, s_maxn;
  const uint32_t tid = threadIdx.x;
  const uint32_t ncand = counters[4] < w.cap ? counters[4] : w.cap;
  if (tid == 0) { s_len = 0; s_end = DW_NONE; s_carry_s = 0; s_carry_p = 0; s_maxn = 0; }
  for (uint32_t c = tid; c < ncand; c += 1024u) {
    w.jmp0[c] = dw_lookup(w, w.next[c]);
    w.ord[c] = (w.pos[c] == 43u) ? 0u : DW_NONE;
  }
  __syncthreads();
  uint32_t* ja = w.jmp0; uint32_t* jb = w.jmp1;
  for (uint32_t r = 0; r < 32u && (r == 0u || (1u << (r - 1u)) < ncand); r++) {
    for (uint32_t c = tid; c < ncand; c += 1024u) {
      const uint32_t t = ja[c];
      const uint32_t oc = w.ord[c];
      if (t != DW_NONE && oc != DW_NONE) w.ord[t] = oc + (1u << r);
      jb[c] = (t != DW_NONE) ? ja[t] : DW_NONE;
    }
    __syncthreads();
    uint32_t* tmp = ja; ja = jb; jb = tmp;
  }
  /* blocks in stream order */
  for (uint32_t c = tid; c < ncand; c += 1024u) {
    const uint32_t oc = w.ord[c];
    if (oc != DW_NONE && oc < max_blocks) {
      blk_off[oc] = w.pos[c]; blk_n[oc] = w.nsmp[c]; w.bnext[oc] = w.next[c];
      atomicMax(&s_len, oc + 1u);
    }
  }
  __syncthreads();
  const uint32_t len = s_len;
  /* sample and padded-sample offsets; the first block the walk would not accept */
  for (uint32_t base = 0; base < len; base += 1024u) {
    const uint32_t i = base + tid;
    const uint32_t n = (i < len) ? blk_n[i] : 0u;
    uint32_t es, ep;
    const uint32_t ts = dw_block_scan(n, warp_sum, es);
    const uint32_t tp = dw_block_scan((n + 7u) & ~7u, warp_sum, ep);
    const uint32_t smp = s_carry_s + es, pst = s_carry_p + ep;
    if (i < len) {
      blk_smp[i] = smp; blk_pst[i] = pst;
      atomicMax(&s_maxn, n);
      if (smp >= max_samples || n > max_samples - smp) atomicMin(&s_end, i);
    }
    __syncthreads();
    if (tid == 0) { s_carry_s += ts; s_carry_p += tp; }
    __syncthreads();
  }
  if (tid != 0) return;
  uint32_t nb, total, padded, err = 0;
  if (s_end != DW_NONE) {
    nb = s_end; total = blk_smp[nb]; padded = blk_pst[nb];
    if (total < max_samples) err = SLAB_RES_INSUFFICIENT_BUFFER;      /* block nb does not fit the sample budget */
  } else {
    nb = len; total = s_carry_s; padded = s_carry_p;
    if (total < max_samples && nb < max_blocks) {
      /* the walk would look for a block here and not find an acceptable one */
      const uint32_t off = nb ? w.bnext[nb - 1u] : 43u;
      if (off > stream_size || stream_size - off < 11u) err = SLAB_RES_INSUFFICIENT_DATA;
      else if (stream[off] != 0xFF || stream[off + 1u] != 0xFF) err = SLAB_RES_SYNC_CODE;
      else err = SLAB_RES_INSUFFICIENT_DATA;                            /* size field runs past the stream */
    }
  }
  counters[0] = nb; counters[1] = total; counters[2] = err; counters[3] = padded; counters[6] = s_maxn;
}

/* ------------------------------------------------------------------ D1a: per-block CRC check */
__global__ void __launch_bounds__(128) k_dec_crc(const uint8_t* __restrict__ stream, DecShape sh,
    const uint32_t* __restrict__ blk_off, uint32_t* __restrict__ err)
{
  __shared__ SlabCrcTables tb;
  slab_crc16_build_tables(&tb);
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (warp >= sh.nblocks) return;                      /* whole warps leave together */
  const uint8_t* b = stream + blk_off[warp];
  const uint32_t size = (((uint32_t)b[2] << 24) | ((uint32_t)b[3] << 16) | ((uint32_t)b[4] << 8) | b[5]) + 6u;
  const uint32_t stored = ((uint32_t)b[6] << 8) | b[7];
  const uint32_t total = size >= 8u ? size - 8u : 0u;
  const uint32_t slice = (total + 31u) / 32u;
  uint32_t lo = lane * slice, hi = lo + slice;
  if (lo > total) lo = total;
  if (hi > total) hi = total;
  uint32_t crc = slab_crc16_run(&tb, b + 8u + lo, hi - lo);
  crc = slab_crc16_mul(crc, slab_crc16_xpow8(total - hi));
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) crc ^= __shfl_xor_sync(SLAB_FULL_MASK, crc, d);
  if (lane == 0 && crc != stored) err[warp] = SLAB_RES_DATA_CORRUPTION;
}

template <int NCH>
__global__ void __launch_bounds__(32) k_dec_entropy(const uint32_t* __restrict__ words, DecShape sh,
    const uint32_t* __restrict__ blk_off, const uint32_t* __restrict__ blk_pst,
    const uint32_t* __restrict__ blk_n,
    int32_t* __restrict__ work, uint32_t* __restrict__ type_out, int32_t* __restrict__ kq_out,
    int32_t* __restrict__ ltq_out, uint32_t* __restrict__ pitch_out, uint32_t* __restrict__ err)
{
  __shared__ __align__(16) unsigned char rings[32u * SLAB_BR_RING];
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= sh.nblocks) return;
  DeOutArrays o; o.type = type_out; o.kq = kq_out; o.ltq = ltq_out; o.pitch = pitch_out; o.err = err;
  DeLane<NCH> L;
  L.begin(rings + threadIdx.x * SLAB_BR_RING, words, sh, b, blk_off, blk_n, o);
  if (L.mode == DE_IDLE) return;
  DeGlobalSink sink; sink.base = work + blk_pst[b]; sink.stride = sh.NP;
  L.span(0, L.n, sink);
  L.finish(o);
}

template <int LMS_N, int PMAX, int TAPS>
__global__ void __launch_bounds__(64) k_dec_synth(DecShape sh,
    const uint32_t* __restrict__ blk_smp, const uint32_t* __restrict__ blk_n,
    const uint32_t* __restrict__ type_in, const int32_t* __restrict__ kq_in,
    const int32_t* __restrict__ ltq_in, const uint32_t* __restrict__ pitch_in,
    const uint32_t* __restrict__ err,
    int32_t* __restrict__ work, int32_t* __restrict__ scratch)
{
  const uint32_t bc = blockIdx.x * blockDim.x + threadIdx.x;
  if (bc >= sh.nblocks * sh.nch) return;
  const uint32_t b = bc / sh.nch, c = bc - b * sh.nch;
  if (type_in[b] != SLAB_BLOCK_COMPRESS || err[b] != 0) return;
  SynthLane<LMS_N, PMAX, TAPS> S;
  S.begin(sh, bc, blk_n[b], work + (size_t)c * sh.NP + blk_smp[b], scratch + (size_t)c * sh.NP + blk_smp[b],
          kq_in, ltq_in, pitch_in);                                /* blk_smp = padded starts here */
  S.chunk(0, nullptr);
  uint32_t s0 = LMS_N;
  if (!S.use_lt || S.lt_far) {
    for (; s0 + LMS_N <= S.n; s0 += LMS_N)
      synth_chunk<LMS_N, PMAX, TAPS, false>(S.st, S.kk, S.bw, S.ltc, S.emph_prev, S.x, S.lt_hist, s0, S.n, S.delay, S.use_lt, true, false, S.filter, nullptr);
  }
  for (; s0 < S.n; s0 += LMS_N) S.chunk(s0, nullptr);
}

/* Generic fallback for parameter sets outside the specialised instantiations (LMS 16/32, PARCOR
 * order above 32, more than 3 long-term taps): runtime loops, state in local memory. */
__global__ void __launch_bounds__(64) k_dec_synth_generic(DecShape sh,
    const uint32_t* __restrict__ blk_smp, const uint32_t* __restrict__ blk_n,
    const uint32_t* __restrict__ type_in, const int32_t* __restrict__ kq_in,
    const int32_t* __restrict__ ltq_in, const uint32_t* __restrict__ pitch_in,
    const uint32_t* __restrict__ err,
    int32_t* __restrict__ work, int32_t* __restrict__ scratch)
{
  const uint32_t bc = blockIdx.x * blockDim.x + threadIdx.x;
  if (bc >= sh.nblocks * sh.nch) return;
  const uint32_t b = bc / sh.nch, c = bc - b * sh.nch;
  if (type_in[b] != SLAB_BLOCK_COMPRESS || err[b] != 0) return;
  const uint32_t n = blk_n[b], N = sh.lms, P = sh.P, T = sh.T;
  int32_t* x = work + (size_t)c * sh.NP + blk_smp[b];                /* blk_smp = padded starts here */
  int32_t* lt_hist = scratch + (size_t)c * sh.NP + blk_smp[b];
  int32_t kk[SLAB_MAX_PARCOR + 1], bw[SLAB_MAX_PARCOR + 1];
  for (uint32_t m = 0; m <= P; m++) { kk[m] = kq_in[(size_t)bc * sh.pstride + m]; bw[m] = 0; }
  const uint32_t pitch = pitch_in[bc], delay = pitch + (T >> 1);
  int32_t ltc[SLAB_MAX_TAPS];
  for (uint32_t j = 0; j < SLAB_MAX_TAPS; j++) ltc[j] = (pitch != 0 && j < T) ? ltq_in[(size_t)bc * 8 + j] : 0;
  int32_t cx[SLAB_MAX_LMS], cp[SLAB_MAX_LMS], hx[SLAB_MAX_LMS], hp[SLAB_MAX_LMS];   /* ring: slot = time mod N */
  for (uint32_t i = 0; i < N; i++) { cx[i] = cp[i] = 0; hx[i] = hp[i] = 0; }
  int32_t emph_prev = 0;
  for (uint32_t s = 0; s < n; s++) {
    const int32_t resid = x[s];
    int32_t v = resid;
    if (n > N) {
      const uint32_t slot = s & (N - 1u);
      if (s < N) { hx[slot] = hp[slot] = resid; }
      else {
        uint32_t acc = 1u << 9;
        for (uint32_t i = 0; i < N; i++) {
          const uint32_t q = (s - 1u - i) & (N - 1u);
          acc += (uint32_t)cx[i] * (uint32_t)hx[q] + (uint32_t)cp[i] * (uint32_t)hp[q];
        }
        const int32_t pred = (int32_t)acc >> 10;
        v = (int32_t)((uint32_t)resid + (uint32_t)pred);
        const uint32_t mag = (resid < 0) ? (0u - (uint32_t)resid) : (uint32_t)resid;
        const int32_t step = slab_sgn(resid) * (int32_t)(slab_bitlen(mag) >> 1);
        for (uint32_t i = 0; i < N; i++) {
          const uint32_t q = (s - 1u - i) & (N - 1u);
          cx[i] += step * slab_sgn(hx[q]); cp[i] += step * slab_sgn(hp[q]);
        }
        hx[slot] = v; hp[slot] = pred;
      }
    }
    if (pitch != 0) {
      if (s >= delay) {
        long long acc = 1ll << 30;
        for (uint32_t j = 0; j < T; j++) acc = slab_mad_wide(ltc[j], lt_hist[s - delay + j], acc);
        v = (int32_t)((uint32_t)v + (uint32_t)(int32_t)(acc >> 31));
      }
      lt_hist[s] = v;
    }
    int32_t f = v;
    for (uint32_t m = P; m >= 1; m--) {
      f += slab_latmul(kk[m], bw[m - 1]);
      bw[m] = bw[m - 1] - slab_latmul(kk[m], f);
    }
    bw[0] = f;
    f = (int32_t)((uint32_t)f + (uint32_t)slab_emph(emph_prev));
    emph_prev = f;
    x[s] = f;
  }
}

/* ------------------------------------------------------------------ D3: MS->LR, shift, store */
/* CTA = 1024 consecutive samples of one block; thread t handles samples t, t + 256, t + 512, t + 768,
 * so every load and store instruction of a warp covers 128 contiguous bytes whatever the block's
 * alignment in the caller's planes (blocks may start at any sample there).  All loads of a thread are
 * issued before the first store. */
template <int NCH, bool MS>
__global__ void __launch_bounds__(256) k_dec_output(DecShape sh,
    const uint32_t* __restrict__ blk_smp, const uint32_t* __restrict__ blk_pst,
    const uint32_t* __restrict__ blk_n,
    const uint32_t* __restrict__ type_in, const int32_t* __restrict__ work, OutPtrs out)
{
  const uint32_t b = blockIdx.x;
  const uint32_t n = blk_n[b];
  const uint32_t i0 = blockIdx.y * 1024u + threadIdx.x;
  if (blockIdx.y * 1024u >= n) return;
  const uint32_t type = type_in[b];
  const uint32_t up = 32u - sh.bits + sh.lshift;
  const size_t pos = (size_t)blk_smp[b] + i0;          /* in the caller's planes */
  const size_t wpos = (size_t)blk_pst[b] + i0;         /* in the padded work planes */
  const uint32_t nch = NCH ? (uint32_t)NCH : sh.nch;
  if (type == SLAB_BLOCK_SILENT) {
    for (uint32_t c = 0; c < nch; c++)
#pragma unroll
      for (uint32_t k = 0; k < 4u; k++)
        if (i0 + 256u * k < n) __stcs(out.p[c] + pos + 256u * k, 0);
    return;
  }
  if (NCH == 2) {
    int32_t a[4], d[4];
#pragma unroll
    for (uint32_t k = 0; k < 4u; k++) {
      const bool live = i0 + 256u * k < n;
      a[k] = live ? __ldcs(work + wpos + 256u * k) : 0;
      d[k] = live ? __ldcs(work + (size_t)sh.NP + wpos + 256u * k) : 0;
    }
#pragma unroll
    for (uint32_t k = 0; k < 4u; k++) {
      if (i0 + 256u * k < n) {
        int32_t l = a[k], r = d[k];
        if (MS) {
          const int32_t side = d[k];
          const int32_t mid = (int32_t)(((uint32_t)a[k] << 1) | ((uint32_t)side & 1u));   /* SLAUtility.c:427-432 */
          l = (mid + side) >> 1; r = (mid - side) >> 1;
        }
        __stcs(out.p[0] + pos + 256u * k, (int32_t)((uint32_t)l << up));
        __stcs(out.p[1] + pos + 256u * k, (int32_t)((uint32_t)r << up));
      }
    }
  } else {
    for (uint32_t c = 0; c < nch; c++) {
      int32_t a[4];
#pragma unroll
      for (uint32_t k = 0; k < 4u; k++)
        a[k] = (i0 + 256u * k < n) ? __ldcs(work + (size_t)c * sh.NP + wpos + 256u * k) : 0;
#pragma unroll
      for (uint32_t k = 0; k < 4u; k++)
        if (i0 + 256u * k < n) __stcs(out.p[c] + pos + 256u * k, (int32_t)((uint32_t)a[k] << up));
    }
  }
}

/* ------------------------------------------------------------------ host-side launch sequence */
/* specialised instantiations: LMS order 4/8, PARCOR order <= 32, <= 3 taps (all reference presets) */
template <int LMS_N, int TAPS>
static int launch_synth_t(SlabCtx* ctx, const DecShape& sh, int pmax, const uint32_t* blk_smp,
    const uint32_t* blk_n, const uint32_t* type, const int32_t* kq, const int32_t* ltq,
    const uint32_t* pitch, const uint32_t* err, int32_t* work, int32_t* scratch)
{
  const unsigned threads = 64, grid = slab_div_up((uint64_t)sh.nblocks * sh.nch, threads);
  switch (pmax) {
    case 8:  SLAB_RUN(ctx, "D2 k_dec_synth", (k_dec_synth<LMS_N, 8, TAPS>), grid, threads, 0, sh, blk_smp, blk_n, type, kq, ltq, pitch, err, work, scratch); break;
    case 16: SLAB_RUN(ctx, "D2 k_dec_synth", (k_dec_synth<LMS_N, 16, TAPS>), grid, threads, 0, sh, blk_smp, blk_n, type, kq, ltq, pitch, err, work, scratch); break;
    default: SLAB_RUN(ctx, "D2 k_dec_synth", (k_dec_synth<LMS_N, 32, TAPS>), grid, threads, 0, sh, blk_smp, blk_n, type, kq, ltq, pitch, err, work, scratch); break;
  }
  return 0;
}

static int launch_synth(SlabCtx* ctx, const DecShape& sh, int pmax, const uint32_t* blk_smp,
    const uint32_t* blk_n, const uint32_t* type, const int32_t* kq, const int32_t* ltq,
    const uint32_t* pitch, const uint32_t* err, int32_t* work, int32_t* scratch)
{
  if ((sh.lms == 4 || sh.lms == 8) && pmax <= 32 && sh.T <= 3) {
    if (sh.lms == 4) {
      if (sh.T <= 1) return launch_synth_t<4, 1>(ctx, sh, pmax, blk_smp, blk_n, type, kq, ltq, pitch, err, work, scratch);
      else return launch_synth_t<4, 3>(ctx, sh, pmax, blk_smp, blk_n, type, kq, ltq, pitch, err, work, scratch);
    } else {
      if (sh.T <= 1) return launch_synth_t<8, 1>(ctx, sh, pmax, blk_smp, blk_n, type, kq, ltq, pitch, err, work, scratch);
      else return launch_synth_t<8, 3>(ctx, sh, pmax, blk_smp, blk_n, type, kq, ltq, pitch, err, work, scratch);
    }
  } else {
    const unsigned threads = 64, grid = slab_div_up((uint64_t)sh.nblocks * sh.nch, threads);
    SLAB_RUN(ctx, "D2 k_dec_synth_generic", k_dec_synth_generic, grid, threads, 0, sh, blk_smp, blk_n, type, kq, ltq, pitch, err, work, scratch);
  }
  return 0;
}

template <int NCH>
static int launch_entropy(SlabCtx* ctx, const DecShape& sh, const uint32_t* words,
    const uint32_t* blk_off, const uint32_t* blk_smp, const uint32_t* blk_n, int32_t* work,
    uint32_t* type, int32_t* kq, int32_t* ltq, uint32_t* pitch, uint32_t* err)
{
  SLAB_RUN(ctx, "D1b k_dec_entropy", (k_dec_entropy<NCH>), slab_div_up(sh.nblocks, 32), 32, 0, words, sh, blk_off, blk_smp, blk_n,
           work, type, kq, ltq, pitch, err);
  return 0;
}

extern "C" int slab_decode(SlabCtx* ctx, SlabDecodeJob* job)
{
  DecShape sh;
  memset(&sh, 0, sizeof(sh));
  sh.nch = job->num_channels; sh.bits = job->bits_per_sample; sh.lshift = job->offset_lshift;
  sh.P = job->parcor_order; sh.T = job->longterm_order; sh.lms = job->lms_order;
  sh.ms = (job->ch_process == 1); sh.check_crc = job->check_crc;
  sh.stream_size = job->stream_size;
  sh.nwords = (((job->stream_size + 3u) / 4u + 15u) & ~15u) + 32u;   /* whole 64-byte chunks + two spare */
  job->first_bad_block = 0xFFFFFFFFu; job->first_bad_code = 0;
  job->decoded_blocks = 0; job->decoded_samples = 0;
  ctx->launches = 0;
  slab_prof_reset(ctx);
  if (sh.nch < 1 || sh.nch > SLAB_MAX_CH || sh.P > SLAB_MAX_PARCOR || sh.T > SLAB_MAX_TAPS ||
      sh.lms > SLAB_MAX_LMS || (sh.lms & (sh.lms - 1)) != 0 || sh.lms < 4) {
    slab_set_error("sla_b200: decode parameters outside the supported envelope");
    return -1;
  }
  const int pmax = sh.P <= 8 ? 8 : sh.P <= 16 ? 16 : sh.P <= 32 ? 32 : 64;
  sh.pstride = (uint32_t)pmax + 1u;

  SLAB_CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
  /* stream image: word-aligned, zero padded so the bit reader may over-read safely */
  uint8_t* d_stream = (uint8_t*)slab_arena(ctx, DA_STREAM, (size_t)sh.nwords * 4u);
  if (!d_stream) return -1;
  {
    const size_t tail = (size_t)(job->stream_size & ~63u);
    SLAB_CUDA_TRY(cudaMemsetAsync(d_stream + tail, 0, (size_t)sh.nwords * 4u - tail, ctx->stream));
  }
  SLAB_CUDA_TRY(cudaMemcpyAsync(d_stream, job->stream, job->stream_size,
                                job->stream_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                ctx->stream));

  uint32_t nblocks = job->num_blocks, total = job->total_samples;
  const int device_walk = (job->blk_byte_off == NULL);
  uint32_t max_blocks = nblocks;
  if (device_walk) {
    max_blocks = job->max_samples + 1u;
    if (max_blocks > (1u << 22) || max_blocks == 0) max_blocks = 1u << 22;
  }
  uint32_t* d_off = slab_arena_as<uint32_t>(ctx, DA_BLK_OFF, max_blocks ? max_blocks : 1);
  uint32_t* d_smp = slab_arena_as<uint32_t>(ctx, DA_BLK_SMP, max_blocks ? max_blocks : 1);
  uint32_t* d_n   = slab_arena_as<uint32_t>(ctx, DA_BLK_N, max_blocks ? max_blocks : 1);
  uint32_t* d_pst = slab_arena_as<uint32_t>(ctx, DA_BLK_PST, max_blocks ? max_blocks : 1);
  uint32_t* d_cnt = slab_arena_as<uint32_t>(ctx, DA_COUNTERS, 8);
  uint32_t* h_pin = (uint32_t*)slab_pinned(ctx, 64);
  if (!d_off || !d_smp || !d_n || !d_pst || !d_cnt || !h_pin) return -1;
  uint32_t walk_err = 0, padded = 0, max_n = 0;
  if (device_walk) {
    /* candidate capacity: one per KiB of stream plus slack; denser streams (long runs of tiny
     * blocks) fall back to the serial walk */
    DecWalk w;
    w.cap = job->stream_size / 1024u + 4096u;
    uint32_t hsize = 1u;
    while (hsize < 2u * w.cap) hsize <<= 1;
    w.hmask = hsize - 1u;
    w.pos = slab_arena_as<uint32_t>(ctx, DA_W_POS, w.cap);
    w.next = slab_arena_as<uint32_t>(ctx, DA_W_NEXT, w.cap);
    w.nsmp = slab_arena_as<uint32_t>(ctx, DA_W_N, w.cap);
    w.jmp0 = slab_arena_as<uint32_t>(ctx, DA_W_J0, w.cap);
    w.jmp1 = slab_arena_as<uint32_t>(ctx, DA_W_J1, w.cap);
    w.ord = slab_arena_as<uint32_t>(ctx, DA_W_ORD, w.cap);
    w.hkey = slab_arena_as<uint32_t>(ctx, DA_W_HKEY, hsize);
    w.hval = slab_arena_as<uint32_t>(ctx, DA_W_HVAL, hsize);
    w.bnext = slab_arena_as<uint32_t>(ctx, DA_W_BNEXT, max_blocks);
    if (!w.pos || !w.next || !w.nsmp || !w.jmp0 || !w.jmp1 || !w.ord || !w.hkey || !w.hval || !w.bnext) return -1;
    SLAB_CUDA_TRY(cudaMemsetAsync(w.hkey, 0, (size_t)hsize * 4u, ctx->stream));
    SLAB_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 32, ctx->stream));
    {
      const unsigned groups = slab_div_up(job->stream_size, 16);
      unsigned grid = slab_div_up(groups, 256);
      if (grid > 148u * 16u) grid = 148u * 16u;
      if (grid == 0) grid = 1;
      SLAB_RUN(ctx, "D0 k_dec_findsync", k_dec_findsync, grid, 256, 0, d_stream, job->stream_size, w, d_cnt);
    }
    SLAB_RUN(ctx, "D0 k_dec_chain", k_dec_chain, 1, 1024, 0, d_stream, job->stream_size, job->max_samples, max_blocks,
             w, d_off, d_smp, d_n, d_pst, d_cnt);
    SLAB_CUDA_TRY(cudaMemcpyAsync(h_pin, d_cnt, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SLAB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_pin[5] != 0) {
      SLAB_RUN(ctx, "D0 k_dec_walk", k_dec_walk, 1, 32, 0, d_stream, job->stream_size, job->max_samples, max_blocks,
               d_off, d_smp, d_n, d_pst, d_cnt);
    }
    SLAB_CUDA_TRY(cudaMemcpyAsync(h_pin, d_cnt, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SLAB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    nblocks = h_pin[0]; total = h_pin[1]; walk_err = h_pin[2]; padded = h_pin[3]; max_n = h_pin[6];
  } else if (nblocks) {
    /* padded block starts for the work planes */
    uint32_t* h_pst = (uint32_t*)slab_host_scratch(ctx, sizeof(uint32_t) * nblocks);
    if (!h_pst) return -1;
    for (uint32_t b = 0; b < nblocks; b++) {
      h_pst[b] = padded; padded += (job->blk_nsmp[b] + 7u) & ~7u;
      if (job->blk_nsmp[b] > max_n) max_n = job->blk_nsmp[b];
    }
    SLAB_CUDA_TRY(cudaMemcpyAsync(d_off, job->blk_byte_off, nblocks * 4u, cudaMemcpyHostToDevice, ctx->stream));
    SLAB_CUDA_TRY(cudaMemcpyAsync(d_smp, job->blk_smp_off, nblocks * 4u, cudaMemcpyHostToDevice, ctx->stream));
    SLAB_CUDA_TRY(cudaMemcpyAsync(d_n, job->blk_nsmp, nblocks * 4u, cudaMemcpyHostToDevice, ctx->stream));
    SLAB_CUDA_TRY(cudaMemcpyAsync(d_pst, h_pst, nblocks * 4u, cudaMemcpyHostToDevice, ctx->stream));
  }
  SLAB_CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
  sh.nblocks = nblocks; sh.total_samples = total;
  sh.NP = (padded + 15u) & ~7u;
  job->decoded_blocks = nblocks; job->decoded_samples = total;

  if (nblocks > 0 && total > 0) {
    const size_t plane = (size_t)total, wplane = (size_t)sh.NP;
    int32_t* d_work = slab_arena_as<int32_t>(ctx, DA_WORK, wplane * sh.nch);
    uint32_t* d_type = slab_arena_as<uint32_t>(ctx, DA_TYPE, nblocks);
    int32_t* d_kq = slab_arena_as<int32_t>(ctx, DA_KQ, (size_t)nblocks * sh.nch * sh.pstride);
    int32_t* d_ltq = slab_arena_as<int32_t>(ctx, DA_LTQ, (size_t)nblocks * sh.nch * 8);
    uint32_t* d_pitch = slab_arena_as<uint32_t>(ctx, DA_PITCH, (size_t)nblocks * sh.nch);
    uint32_t* d_err = slab_arena_as<uint32_t>(ctx, DA_ERR, nblocks);
    if (!d_work || !d_type || !d_kq || !d_ltq || !d_pitch || !d_err) return -1;
    OutPtrs out;
    memset(&out, 0, sizeof(out));
    int32_t* d_out = NULL;
    if (job->out_on_device) {
      for (uint32_t c = 0; c < sh.nch; c++) out.p[c] = job->out[c];
    } else {
      d_out = slab_arena_as<int32_t>(ctx, DA_OUT, wplane * sh.nch);     /* also the long-term scratch: wplane >= plane */
      if (!d_out) return -1;
      for (uint32_t c = 0; c < sh.nch; c++) out.p[c] = d_out + plane * c;
    }
    /* the long-term stage keeps its output history in a per-channel scratch plane */
    int32_t* d_scratch = job->out_on_device ? slab_arena_as<int32_t>(ctx, DA_OUT, wplane * sh.nch) : d_out;
    if (!d_scratch) return -1;

    SLAB_CUDA_TRY(cudaMemsetAsync(d_err, 0, nblocks * 4u, ctx->stream));
    if (sh.check_crc) {
      SLAB_RUN(ctx, "D1a k_dec_crc", k_dec_crc, slab_div_up((uint64_t)nblocks * 32u, 128), 128, 0, d_stream, sh, d_off, d_err);
    }
    const uint32_t* words = (const uint32_t*)d_stream;
    /* mono / stereo with preset-class parameters: entropy decode and synthesis overlapped in one kernel */
    const int fused = slab_decode_fused(ctx, sh, pmax, words, d_off, d_pst, d_n, d_work, d_scratch, d_type, d_kq,
                                        d_ltq, d_pitch, d_err);
    if (fused < 0) return -1;
    if (fused == 1) {
      switch (sh.nch) {
        case 1: if (launch_entropy<1>(ctx, sh, words, d_off, d_pst, d_n, d_work, d_type, d_kq, d_ltq, d_pitch, d_err) != 0) return -1; break;
        case 2: if (launch_entropy<2>(ctx, sh, words, d_off, d_pst, d_n, d_work, d_type, d_kq, d_ltq, d_pitch, d_err) != 0) return -1; break;
        case 3: if (launch_entropy<3>(ctx, sh, words, d_off, d_pst, d_n, d_work, d_type, d_kq, d_ltq, d_pitch, d_err) != 0) return -1; break;
        case 4: if (launch_entropy<4>(ctx, sh, words, d_off, d_pst, d_n, d_work, d_type, d_kq, d_ltq, d_pitch, d_err) != 0) return -1; break;
        case 5: if (launch_entropy<5>(ctx, sh, words, d_off, d_pst, d_n, d_work, d_type, d_kq, d_ltq, d_pitch, d_err) != 0) return -1; break;
        case 6: if (launch_entropy<6>(ctx, sh, words, d_off, d_pst, d_n, d_work, d_type, d_kq, d_ltq, d_pitch, d_err) != 0) return -1; break;
        case 7: if (launch_entropy<7>(ctx, sh, words, d_off, d_pst, d_n, d_work, d_type, d_kq, d_ltq, d_pitch, d_err) != 0) return -1; break;
        default: if (launch_entropy<8>(ctx, sh, words, d_off, d_pst, d_n, d_work, d_type, d_kq, d_ltq, d_pitch, d_err) != 0) return -1; break;
      }
      if (launch_synth(ctx, sh, pmax, d_pst, d_n, d_type, d_kq, d_ltq, d_pitch, d_err, d_work, d_scratch) != 0) return -1;
    }
    {
      /* rows for the largest block actually on the chain: a block header may carry up to 65535 samples
       * whatever the container header's max_num_block_samples says, and the reference decodes such a
       * block as long as it fits the handle (SLADecoder.c:633) */
      dim3 grid(nblocks, slab_div_up(max_n ? max_n : 1u, 1024));
      if (sh.nch == 2 && sh.ms) SLAB_RUN(ctx, "D3 k_dec_output", (k_dec_output<2, true>), grid, 256, 0, sh, d_smp, d_pst, d_n, d_type, d_work, out);
      else if (sh.nch == 2) SLAB_RUN(ctx, "D3 k_dec_output", (k_dec_output<2, false>), grid, 256, 0, sh, d_smp, d_pst, d_n, d_type, d_work, out);
      else SLAB_RUN(ctx, "D3 k_dec_output", (k_dec_output<0, false>), grid, 256, 0, sh, d_smp, d_pst, d_n, d_type, d_work, out);
    }
    SLAB_CUDA_TRY(cudaEventRecord(ctx->ev[2], ctx->stream));

    uint32_t* h_err = (uint32_t*)slab_pinned(ctx, (size_t)nblocks * 4u + 64);
    if (!h_err) return -1;
    SLAB_CUDA_TRY(cudaMemcpyAsync(h_err, d_err, nblocks * 4u, cudaMemcpyDeviceToHost, ctx->stream));
    if (!job->out_on_device)
      for (uint32_t c = 0; c < sh.nch; c++)
        SLAB_CUDA_TRY(cudaMemcpyAsync(job->out[c], d_out + plane * c, plane * 4u, cudaMemcpyDeviceToHost, ctx->stream));
    SLAB_CUDA_TRY(cudaEventRecord(ctx->ev[3], ctx->stream));
    SLAB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (uint32_t b = 0; b < nblocks; b++)
      if (h_err[b] != 0) { job->first_bad_block = b; job->first_bad_code = h_err[b]; break; }
    if (job->blk_err_out) memcpy(job->blk_err_out, h_err, (size_t)nblocks * 4u);
    cudaEventElapsedTime(&ctx->last_ms[SLAB_T_H2D], ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->last_ms[SLAB_T_KERNELS], ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->last_ms[SLAB_T_D2H], ctx->ev[2], ctx->ev[3]);
    slab_prof_collect(ctx);
  } else {
    SLAB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  if (device_walk && walk_err != 0 && job->first_bad_block == 0xFFFFFFFFu) {
    job->first_bad_block = nblocks; job->first_bad_code = walk_err;
  }
  SLAB_CUDA_TRY(cudaGetLastError());
  return 0;
}
