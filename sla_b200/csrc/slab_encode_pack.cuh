/* slab_encode_pack.cuh - E9 bit packing of recursive-Rice blocks, mono and stereo (the common case of
 * SLACoder_PutDataArray, src/SLACoder.c:224-270,429-467, with the MSB-first writer of
 * src/include/private/SLABitStream.h:166-216).
 *
 * One CTA per block, tiles of 2048 samples.  A thread owns EIGHT consecutive samples of the tile (all
 * channels: 8 or 16 codes): two 128-bit loads per channel bring the residuals, one brings the eight pairs of
 * Rice exponents the trace kernel left.  Pass 1 adds up the thread's code lengths; a block scan turns the
 * 256 totals into bit offsets; pass 2 assembles the codes in a 64-bit register accumulator and ORs every
 * completed 32-bit word into the shared-memory stage: one atomic per word (about ten per thread and tile)
 * where the general kernel k_enc_pack does one or two per code, and one scan + four barriers per 2048
 * samples instead of per 256.  The tile is flushed as aligned 32-bit words.
 * Escapes (quotient >= 16: 16 zeros, a one, a gamma code) go through the same accumulator piecewise.
 * A tile whose bits exceed the 32 KB stage is written by k_enc_pack's byte-serial path instead: the block is
 * then left to that kernel (flagged in `defer`). */
#ifndef SLAB_ENCODE_PACK_CUH
#define SLAB_ENCODE_PACK_CUH

#define PACK2_ROWS   8u
#define PACK2_TILE   (256u * PACK2_ROWS)

struct PackAcc {
  uint32_t* stage;          /* shared-memory words, MSB-first, zeroed before the tile */
  uint64_t  acc;            /* pending bits, left-aligned */
  uint32_t  nbits;          /* bits pending in acc, including the bit offset inside the first word */
  uint32_t  word;           /* index of the word the pending bits start in */
  __device__ __forceinline__ void begin(uint32_t* s, uint32_t bitpos)
  {
    stage = s; word = bitpos >> 5; nbits = bitpos & 31u; acc = 0;
  }
  /* a completed word: OR-ed in (the first and the last word of a span are shared with the neighbours; an
   * OR everywhere keeps the path free of branches - uncontended shared-memory atomics are cheap) */
  __device__ __forceinline__ void flush_word()
  {
    const uint32_t w = (uint32_t)(acc >> 32);
    if (w) atomicOr(&stage[word], w);
    word++; acc <<= 32; nbits -= 32u;
  }
  __device__ __forceinline__ void zeros(uint32_t n)        /* any n */
  {
    nbits += n;
    while (nbits >= 32u) flush_word();
  }
  /* n <= 32 bits of v (v < 2^n); at most 31 bits are pending */
  __device__ __forceinline__ void put(uint32_t v, uint32_t n)
  {
    acc |= (uint64_t)v << (64u - nbits - n);
    nbits += n;
    if (nbits >= 32u) flush_word();
  }
  __device__ __forceinline__ void end()
  {
    if (nbits) { const uint32_t w = (uint32_t)(acc >> 32); if (w) atomicOr(&stage[word], w); }
  }
};

/* the rare codes: a quotient of 16 or more (16 zeros, a one, a gamma code, SLACoder.c:120-138) or more than 32
 * bits in all; out of line so that the sixteen unrolled emitters of a thread stay small */
__device__ __noinline__ void pack2_slow(PackAcc* A, uint32_t q, uint32_t k, uint32_t low)
{
  if (q < 16u) {
    A->zeros(q);
    A->put(1u, 1u);
  } else {
    A->zeros(16u);
    const uint32_t g = q - 16u;
    if (g == 0) A->put(3u, 2u);                            /* the terminating one, then gamma(0) = 1 */
    else {
      const uint32_t nd = slab_log2ceil(g + 2u);
      A->put(1u, 1u);
      A->zeros(nd - 1u);
      if (nd > 16u) { A->put((g + 1u) >> 16, nd - 16u); A->put((g + 1u) & 0xFFFFu, 16u); }
      else A->put(g + 1u, nd);
    }
  }
  if (k > 16u) { A->put(low >> 16, k - 16u); A->put(low & 0xFFFFu, 16u); }
  else if (k) A->put(low, k);
}

template <int CH>
__global__ void __launch_bounds__(256, 3) k_enc_pack_rice(EncShape sh,
    const uint32_t* __restrict__ blk_pst, const uint32_t* __restrict__ blk_len,
    const uint32_t* __restrict__ blk_type, const uint32_t* __restrict__ blk_mode,
    const uint32_t* __restrict__ blk_hdr_bytes, const uint32_t* __restrict__ blk_off,
    const EncChan* __restrict__ chan, const int32_t* __restrict__ code_in, const int32_t* __restrict__ ltq_in,
    const int32_t* __restrict__ r3, const uint16_t* __restrict__ meta,
    const uint32_t* __restrict__ misc, uint8_t* __restrict__ out, uint32_t* __restrict__ defer)
{
  __shared__ uint32_t stage[PACK_STAGE_WORDS + 4];
  __shared__ uint32_t warp_tot[8];
  __shared__ int s_defer;
  if (misc[M_OVERFLOW]) return;
  const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  const uint32_t type = blk_type[b], mode = blk_mode[b];
  if (tid == 0) defer[b] = 1u;                         /* until this kernel has finished the block */
  if (type != SLAB_BLOCK_COMPRESS || !mode) return;    /* k_enc_pack writes every other kind of block */
  const uint32_t n = blk_len[b], hdrb = blk_hdr_bytes[b];
  uint8_t* dst = out + blk_off[b];
  const size_t p0 = blk_pst[b];

  /* ---- pass 0: would every tile fit the stage?  (lengths only; almost always yes) ---- */
  if (tid == 0) s_defer = 0;
  for (uint32_t i = tid; i < PACK_STAGE_WORDS + 4u; i += 256u) stage[i] = 0;
  __syncthreads();

  /* ---- header: thread 0 builds it in the stage, everyone copies it out (SLAEncoder.c:685-737) ---- */
  if (tid == 0) {
    uint64_t p = 0;
    pack_put(stage, p, 0xFFFFu, 16); p += 16;
    p += 32 + 16;                                  /* size and CRC are patched by k_enc_crc */
    pack_put(stage, p, n, 16); p += 16;
    pack_put(stage, p, type, 2); p += 2;
    for (uint32_t c = 0; c < sh.nch; c++) {
      const EncChan& ch = chan[b * sh.nch + c];
      const int32_t* pc = code_in + (size_t)(b * sh.nch + c) * (SLAB_MAX_PARCOR + 1);
      pack_put(stage, p, ch.rshift, 4); p += 4;
      for (uint32_t k = 1; k <= sh.P; k++) {
        const uint32_t qb = (k < 4u) ? 16u : 8u;
        pack_put(stage, p, slab_zigzag(pc[k]) & ((1u << qb) - 1u), qb); p += qb;
      }
      if (ch.pitch >= 3u) {
        pack_put(stage, p, 1u, 1); p += 1;
        pack_put(stage, p, ch.pitch, 10); p += 10;
        for (uint32_t k = 0; k < sh.T; k++) {
          pack_put(stage, p, slab_zigzag(ltq_in[(size_t)(b * sh.nch + c) * 8 + k] >> 16) & 0xFFFFu, 16); p += 16;
        }
      } else { p += 1; }
      const uint32_t par = slab_rice_param(ch.rice_init);
      pack_put(stage, p, (sh.bits >= 32u) ? par : (par & ((1u << sh.bits) - 1u)), sh.bits); p += sh.bits;
    }
  }
  __syncthreads();
  const uint32_t hdr_words = (hdrb + 3u) >> 2;
  uint32_t hdr_keep = 0;                               /* my word of the header, written once the block is known to fit */
  if (tid < hdr_words) hdr_keep = stage[tid];
  __syncthreads();
  for (uint32_t i = tid; i < hdr_words + 1u; i += 256u) stage[i] = 0;
  __syncthreads();

  uint64_t byte_cursor = hdrb;
  uint32_t carry_bits = 0;
  bool header_out = false;
  const int4* rv[CH];
  const uint4* mv[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) {
    rv[c] = reinterpret_cast<const int4*>(r3 + (size_t)c * sh.NP + p0);
    mv[c] = reinterpret_cast<const uint4*>(meta + (size_t)c * sh.NP + p0);
  }
  for (uint32_t t0 = 0; t0 < n; t0 += PACK2_TILE) {
    const uint32_t s0 = t0 + tid * PACK2_ROWS;
    const uint32_t cnt = (s0 < n) ? ((n - s0 < PACK2_ROWS) ? n - s0 : PACK2_ROWS) : 0u;
    /* residuals (zigzag) and exponent pairs stay in registers: every loop below is fully unrolled */
    uint32_t val[CH][PACK2_ROWS], met[CH][PACK2_ROWS / 2u];
#pragma unroll
    for (int c = 0; c < CH; c++) {
      int4 a = make_int4(0, 0, 0, 0), bq = make_int4(0, 0, 0, 0);
      uint4 m = make_uint4(0, 0, 0, 0);
      if (cnt) { a = rv[c][(s0 >> 2)]; bq = rv[c][(s0 >> 2) + 1u]; m = mv[c][s0 >> 3]; }
      val[c][0] = slab_zigzag(a.x); val[c][1] = slab_zigzag(a.y); val[c][2] = slab_zigzag(a.z); val[c][3] = slab_zigzag(a.w);
      val[c][4] = slab_zigzag(bq.x); val[c][5] = slab_zigzag(bq.y); val[c][6] = slab_zigzag(bq.z); val[c][7] = slab_zigzag(bq.w);
      met[c][0] = m.x; met[c][1] = m.y; met[c][2] = m.z; met[c][3] = m.w;
    }
#define PACK2_MET(c, r) (((r) & 1) ? (met[c][(r) >> 1] >> 16) : (met[c][(r) >> 1] & 0xFFFFu))
    /* pass 1: bits of my codes */
    uint32_t mine = 0;
#pragma unroll
    for (int r = 0; r < (int)PACK2_ROWS; r++) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        const uint32_t mt = PACK2_MET(c, r);
        const uint32_t k0 = mt & 31u, k1 = mt >> 5, v = val[c][r];
        const bool second = v >= (1u << k0);
        const uint32_t rest = v - (1u << k0);
        const uint32_t q = 1u + (rest >> k1);
        uint32_t len = second ? q + 1u + k1 : 1u + k0;
        if (second && q >= 16u) len = 17u + enc_gamma_len(q - 16u) + k1;
        mine += ((uint32_t)r < cnt) ? len : 0u;
      }
    }
    uint32_t x = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t yv = __shfl_up_sync(SLAB_FULL_MASK, x, d); if (lane >= (uint32_t)d) x += yv; }
    if (lane == 31u) warp_tot[wid] = x;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8u; w++) { if (w < wid) before += warp_tot[w]; total += warp_tot[w]; }
    if (carry_bits + total > PACK_STAGE_WORDS * 32u) {
      /* giant escapes: the general kernel's byte-serial path writes this block (nothing was written yet
       * when this happens in the first tile; later tiles leave the block to be rewritten as a whole) */
      if (tid == 0) s_defer = 1;
    }
    __syncthreads();
    if (s_defer) return;                             /* defer[b] stays 1 */
    if (!header_out) {
      /* the header bytes, once */
      header_out = true;
      if (tid < hdr_words) {
        const uint32_t w = hdr_keep;
#pragma unroll
        for (uint32_t k = 0; k < 4u; k++) if (4u * tid + k < hdrb) dst[4u * tid + k] = (uint8_t)(w >> (24u - 8u * k));
      }
    }
    /* pass 2: my codes through the accumulator, unrolled.  Both kinds of code have one form - q zeros, then
     * 1 + k bits - and go in with ONE shift-and-or when they have at most 32 bits (at most 31 are pending, the
     * accumulator holds 64), followed by one test for a completed word; the rest (escapes, q >= 16, and codes
     * beyond 32 bits) are an out-of-line call, so the sixteen copies of this body stay small. */
    if (cnt) {
      PackAcc A;
      A.begin(stage, carry_bits + before + (x - mine));
#pragma unroll
      for (int r = 0; r < (int)PACK2_ROWS; r++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
          if ((uint32_t)r < cnt) {
            const uint32_t mt = PACK2_MET(c, r);
            const uint32_t k0 = mt & 31u, k1 = mt >> 5, v = val[c][r];
            const bool second = v >= (1u << k0);
            const uint32_t rest = v - (1u << k0);
            const uint32_t q = second ? 1u + (rest >> k1) : 0u;
            const uint32_t k = second ? k1 : k0;
            const uint32_t low = (second ? rest : v) & ((1u << k) - 1u);
            const uint32_t len = q + 1u + k;
            if (q < 16u && len <= 32u) {
              A.acc |= (uint64_t)((1u << k) | low) << (64u - A.nbits - len);
              A.nbits += len;
              if (A.nbits >= 32u) A.flush_word();
            } else {
              pack2_slow(&A, q, k, low);
            }
          }
        }
      }
      A.end();
    }
#undef PACK2_MET
    __syncthreads();
    const uint32_t nbits = carry_bits + total;
    const uint32_t full = nbits >> 3;
    {
      /* whole bytes of the stage go out as aligned 32-bit words (big-endian words in the stage: a funnel
       * shift picks four stream bytes at any byte offset, a byte permute puts them in memory order) */
      uint8_t* out0 = dst + byte_cursor;
      uint32_t head = (4u - (uint32_t)((size_t)out0 & 3u)) & 3u;
      if (head > full) head = full;
      const uint32_t nw = (full - head) >> 2, done = head + 4u * nw;
      if (tid < head) out0[tid] = (uint8_t)(stage[tid >> 2] >> (24u - 8u * (tid & 3u)));
      for (uint32_t w = tid; w < nw; w += 256u) {
        const uint32_t i = head + 4u * w;
        const uint32_t be = __funnelshift_l(stage[(i >> 2) + 1u], stage[i >> 2], 8u * (i & 3u));
        *reinterpret_cast<uint32_t*>(out0 + i) = __byte_perm(be, 0, 0x0123);
      }
      if (tid < full - done) {
        const uint32_t i = done + tid;
        out0[i] = (uint8_t)(stage[i >> 2] >> (24u - 8u * (i & 3u)));
      }
    }
    const uint32_t tail = (nbits & 7u) ? ((stage[full >> 2] >> (24u - 8u * (full & 3u))) & 0xFFu) : 0u;
    __syncthreads();
    const uint32_t used_words = (nbits + 31u) / 32u + 1u;
    for (uint32_t i = tid; i < used_words; i += 256u) stage[i] = 0;
    __syncthreads();
    if (tid == 0) stage[0] = tail << 24;
    byte_cursor += full;
    carry_bits = nbits & 7u;
    __syncthreads();
  }
  if (tid == 0) {
    if (carry_bits) dst[byte_cursor] = (uint8_t)(stage[0] >> 24);
    defer[b] = 0u;
  }
}

#endif
