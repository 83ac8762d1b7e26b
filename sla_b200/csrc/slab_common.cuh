/*
 * slab_common.cuh - device-side arithmetic shared by the encoder and decoder kernels.
 * Each helper names the reference expression it must reproduce bit-for-bit.
 */
#ifndef SLAB_COMMON_CUH
#define SLAB_COMMON_CUH

#include "slab_cuda.h"

#define SLAB_MAX_CH     8
#define SLAB_MAX_PARCOR 64      /* handle capacity limit we accept (reference CLI uses 48) */
#define SLAB_MAX_TAPS   7
#define SLAB_MAX_LMS    32
#define SLAB_GRID       1024u   /* SLA_SEARCH_BLOCK_NUM_SAMPLES_DELTA, SLAInternal.h:16 */
#define SLAB_MIN_BLOCK  2048u   /* SLA_MIN_BLOCK_NUM_SAMPLES, SLAInternal.h:15 */
#define SLAB_MAX_NODES  66      /* 65536/1024 + 2 */
#define SLAB_MAX_PITCH  256u    /* SLALONGTERM_MAX_PERIOD, SLAInternal.h:9 */
#define SLAB_NUM_LTLAGS 260u    /* lags 0..259 feed the pitch picker and the tap solve */

enum { SLAB_BLOCK_COMPRESS = 0, SLAB_BLOCK_SILENT = 1, SLAB_BLOCK_RAW = 2 };

/* SLAUTILITY_SINT32_TO_UINT32 / UINT32_TO_SINT32, SLAUtility.h:37-39 */
__host__ __device__ __forceinline__ uint32_t slab_zigzag(int32_t s)
{
  return (s < 0) ? (uint32_t)(-(s << 1)) - 1u : (uint32_t)(s << 1);
}
__host__ __device__ __forceinline__ int32_t slab_unzigzag(uint32_t u)
{
  return (int32_t)(u >> 1) ^ -(int32_t)(u & 1u);
}
__host__ __device__ __forceinline__ int32_t slab_sgn(int32_t v) { return (v > 0) - (v < 0); }

__device__ __forceinline__ uint32_t slab_bitlen(uint32_t x) { return 32u - (uint32_t)__clz((int)x); }
/* SLAUTILITY_LOG2CEIL(x) = 32 - nlz(x - 1), SLAUtility.h:53 */
__device__ __forceinline__ uint32_t slab_log2ceil(uint32_t x) { return slab_bitlen(x - 1u); }

/* PARCOR lattice product: (k * v + 2^14) >> 15 in wrapping int32, SLAPredictor.c:590,728 */
__device__ __forceinline__ int32_t slab_latmul(int32_t k, int32_t v)
{
  return (int32_t)((uint32_t)k * (uint32_t)v + (1u << 14)) >> 15;
}

/* Non-binding L1 prefetch of the line holding p.  Register-destination loads cannot stay in flight
 * across a loop back-edge (ptxas drains their scoreboards at the branch), so data that is needed a
 * few iterations later is pulled into L1 with this instead; the later load is then an L1 hit. */
__device__ __forceinline__ void slab_prefetch_l1(const void* p)
{
#ifndef SLAB_EMUL
  asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#else
  (void)p;
#endif
}

/* acc + a * b with a, b 32-bit signed and a 64-bit accumulator: one IMAD.WIDE */
__device__ __forceinline__ long long slab_mad_wide(int32_t a, int32_t b, long long acc)
{
#ifdef SLAB_EMUL
  return acc + (long long)a * (long long)b;
#else
  long long r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(acc));
  return r;
#endif
}

/* (prev * 31) >> 5, SLAPredictor.c:1758,1785 */
__device__ __forceinline__ int32_t slab_emph(int32_t prev) { return (int32_t)((uint32_t)prev * 31u) >> 5; }

/* log2 of the Rice modulus for a Q8 running mean, SLACoder.c:30-31:
 * m = roundup_pow2(max(((p >> 1) + 128) >> 8, 1)) */
__device__ __forceinline__ uint32_t slab_rice_k(uint64_t p)
{
  uint32_t m = (uint32_t)(((p >> 1) + 128u) >> 8);
  m = m < 1u ? 1u : m;
  return slab_log2ceil(m);
}
/* p <- (119 p + 9 * (v << 8) + 64) >> 7 with the 32-bit wraps of SLACoder.c:14,27 */
__device__ __forceinline__ uint64_t slab_rice_update(uint64_t p, uint32_t v)
{
  uint32_t w = 9u * (uint32_t)(v << 8);
  return (119ull * p + (uint64_t)w + 64ull) >> 7;
}
/* SLACODER_PARAMETER_GET, SLACoder.c:22-23 */
__device__ __forceinline__ uint32_t slab_rice_param(uint64_t p)
{
  uint32_t m = (uint32_t)((p + 128u) >> 8);
  return m < 1u ? 1u : m;
}

/* ---------------- CRC-16/IBM (reflected 0xA001, init 0, no xor-out), SLAUtility.c:322-339 ------- */
__host__ __device__ __forceinline__ uint32_t slab_crc16_byte(uint32_t crc, uint32_t byte)
{
  crc ^= byte;
#pragma unroll
  for (int i = 0; i < 8; i++) crc = (crc >> 1) ^ ((crc & 1u) ? 0xA001u : 0u);
  return crc;
}
/* product of two residues mod the CRC polynomial, reflected bit order (x^0 is bit 15) */
__host__ __device__ __forceinline__ uint32_t slab_crc16_mul(uint32_t a, uint32_t b)
{
  uint32_t p = 0;
  for (uint32_t m = 0x8000u; m != 0; m >>= 1) {
    if (a & m) p ^= b;
    b = (b >> 1) ^ ((b & 1u) ? 0xA001u : 0u);
  }
  return p;
}
/* x^(8 * nbytes) mod P: appending nbytes zero bytes multiplies the register by this */
__host__ __device__ __forceinline__ uint32_t slab_crc16_xpow8(uint32_t nbytes)
{
  uint32_t result = 0x8000u, base = 0x4000u;     /* 1 and x */
  uint64_t e = (uint64_t)nbytes * 8u;
  while (e) {
    if (e & 1u) result = slab_crc16_mul(result, base);
    base = slab_crc16_mul(base, base);
    e >>= 1;
  }
  return result;
}

/* ---------------- MSB-first bit reader over a 16-byte aligned, zero-padded device stream --------- */
/* The stream is fetched 16 bytes at a time (one L1 wavefront per 128 bits instead of one per 32) into
 * `cur`; the following 16 bytes are always already in flight in `nxt`, so a top-up never waits on the
 * load it issues.  The top-up itself is branch-free except when a 16-byte group is exhausted. */
struct SlabBitReader {
  const uint4* wv;
  uint64_t buf;        /* next bit = bit 63 */
  uint32_t navail;
  uint32_t next;       /* index of the next 32-bit word to enter the buffer */
  uint32_t nquads;     /* 16-byte groups available (stream is zero padded up to this) */
  uint4 cur, nxt;

  __device__ __forceinline__ uint4 load_quad(uint32_t q) const
  {
    return (q < nquads) ? wv[q] : make_uint4(0u, 0u, 0u, 0u);
  }
  __device__ __forceinline__ uint32_t pick() const
  {
    const uint32_t k = next & 3u;
    const uint32_t lo = (k & 1u) ? cur.y : cur.x, hi = (k & 1u) ? cur.w : cur.z;
    return __byte_perm((k & 2u) ? hi : lo, 0, 0x0123);
  }
  __device__ __forceinline__ void advance()
  {
    next++;
    if ((next & 3u) == 0u) { cur = nxt; nxt = load_quad((next >> 2) + 1u); }
  }
  __device__ __forceinline__ void init(const uint32_t* words, uint32_t total_words, uint64_t byte_off)
  {
    wv = reinterpret_cast<const uint4*>(words); nquads = total_words >> 2;
    next = (uint32_t)(byte_off >> 2);
    cur = load_quad(next >> 2); nxt = load_quad((next >> 2) + 1u);
    const uint32_t skip = (uint32_t)(byte_off & 3u) * 8u;
    buf = ((uint64_t)pick() << 32) << skip;
    navail = 32u - skip;
    advance();
  }
  __device__ __forceinline__ void refill()
  {
    if (navail <= 32u) {
      buf |= (uint64_t)pick() << (32u - navail);
      navail += 32u;
      advance();
    }
  }
  /* n in [0, 32] */
  __device__ __forceinline__ uint32_t get(uint32_t n)
  {
    refill();
    const uint32_t v = (uint32_t)((buf >> 1) >> (63u - n));     /* n == 0 -> 0 */
    buf <<= n; navail -= n;
    return v;
  }
  /* zeros before the next 1 bit; the 1 is consumed (SLABitReader_GetZeroRunLength) */
  __device__ __forceinline__ uint32_t zero_run()
  {
    refill();
    uint32_t lz = (uint32_t)__clzll((long long)buf);
    if (lz < navail) {                              /* common: terminator inside the buffer */
      buf = (buf << lz) << 1; navail -= lz + 1u;
      return lz;
    }
    uint32_t run = 0;
    for (;;) {
      run += navail; buf = 0; navail = 0;
      if ((next >> 2) >= nquads) return run;
      refill();
      lz = (uint32_t)__clzll((long long)buf);
      if (lz < navail) {
        buf = (buf << lz) << 1; navail -= lz + 1u;
        return run + lz;
      }
    }
  }
  __device__ __forceinline__ void align_byte()
  {
    const uint32_t drop = navail & 7u;
    buf <<= drop; navail -= drop;
  }
  /* bytes consumed since the stream start, rounding a partial byte up */
  __device__ __forceinline__ uint64_t byte_pos() const
  {
    return ((uint64_t)next * 32u - navail + 7u) >> 3;
  }
};

#endif
