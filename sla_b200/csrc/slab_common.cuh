/*
 * slab_common.cuh - device-side arithmetic shared by the encoder and decoder kernels.
 * Each helper names the reference expression it must reproduce bit-for-bit.
 */
#ifndef SLAB_COMMON_CUH
#define SLAB_COMMON_CUH

#include "slab_cuda.h"
#include "slab_lanestream.cuh"

#include <string.h>

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
/* The same two on a 32-bit register: starting below 2^32 the running mean stays below 2^32
 * (p' < (119 * 2^32 + 2^32 + 64) / 128), so only the 39-bit intermediate needs a wide multiply. */
__device__ __forceinline__ uint32_t slab_rice_update32(uint32_t p, uint32_t v)
{
  return (uint32_t)(((uint64_t)p * 119u + (uint64_t)(v * 2304u) + 64u) >> 7);
}
/* log2ceil(max((p + 256) >> 9, 1)) = max(bitlen(max(p, 256) - 256) - 9, 0) */
__device__ __forceinline__ uint32_t slab_rice_k32(uint32_t p);
/* SLACODER_PARAMETER_GET, SLACoder.c:22-23 */
__device__ __forceinline__ uint32_t slab_rice_param(uint64_t p)
{
  uint32_t m = (uint32_t)((p + 128u) >> 8);
  return m < 1u ? 1u : m;
}

/* ---------------- software model of an x87 extended-precision accumulator -----------------------
 * The reference forms the refinement residual of its small linear solver in `long double`
 * (SLAUtility.c:627): on x86-64 that is the 80-bit x87 format, 64-bit significand, round to nearest
 * even after every addition.  For nearly singular tap systems the low bits of that residual reach the
 * Q15 tap codes, so the accumulation is reproduced exactly with 64-bit integer arithmetic (verified
 * against native long double on 4 million random sums, tools/x87_check.cpp). */
__host__ __device__ __forceinline__ int slab_clz64(unsigned long long v)
{
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return __builtin_clzll(v);
#endif
}
__host__ __device__ __forceinline__ unsigned long long slab_d2u(double d)
{
#if defined(__CUDA_ARCH__)
  return (unsigned long long)__double_as_longlong(d);
#else
  unsigned long long u; memcpy(&u, &d, 8); return u;
#endif
}
__host__ __device__ __forceinline__ double slab_u2d(unsigned long long u)
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, 8); return d;
#endif
}
struct SlabX87 { unsigned long long m; int e; int s; };      /* value = (-1)^s * m * 2^(e - 63), m normalised (bit 63 set) or 0 */

__host__ __device__ __forceinline__ SlabX87 slab_x87_from_double(double d)
{
  SlabX87 r; r.m = 0; r.e = 0; r.s = 0;
  unsigned long long bits = slab_d2u(d);
  r.s = (int)(bits >> 63);
  const int be = (int)((bits >> 52) & 0x7FFu);
  unsigned long long frac = bits & 0xFFFFFFFFFFFFFull;
  if (be == 0) {
    if (frac == 0) return r;                                  /* zero */
    const int lz = slab_clz64(frac);                             /* subnormal */
    r.m = frac << lz; r.e = -1022 - 52 + (63 - lz);
    return r;
  }
  r.m = ((1ull << 52) | frac) << 11;
  r.e = be - 1023;
  return r;
}

__host__ __device__ __forceinline__ double slab_x87_to_double(SlabX87 a)
{
  if (a.m == 0) return a.s ? -0.0 : 0.0;
  unsigned long long keep = a.m >> 11, rest = a.m & 0x7FFull;
  int e = a.e;
  if (rest > 0x400ull || (rest == 0x400ull && (keep & 1ull))) {
    keep++;
    if (keep >> 53) { keep >>= 1; e++; }
  }
  const unsigned long long bits = ((unsigned long long)a.s << 63) | ((unsigned long long)(e + 1023) << 52) | (keep & 0xFFFFFFFFFFFFFull);
  return slab_u2d(bits);
}

/* 128-bit helpers on (hi, lo) pairs of 64-bit words */
__host__ __device__ __forceinline__ void slab_u128_shr(unsigned long long* hi, unsigned long long* lo, int s, int* sticky)   /* 0 < s < 128 */
{
  unsigned long long h = *hi, l = *lo, lost;
  if (s >= 64) {
    lost = l | ((s > 64) ? (h << (128 - s)) : 0ull);
    l = (s == 64) ? h : (h >> (s - 64));
    h = 0;
  } else {
    lost = l << (64 - s);
    l = (l >> s) | (h << (64 - s));
    h >>= s;
  }
  if (lost) *sticky = 1;
  *hi = h; *lo = l;
}

__host__ __device__ __forceinline__ SlabX87 slab_x87_add(SlabX87 a, SlabX87 b)
{
  if (a.m == 0) return b;
  if (b.m == 0) return a;
  if (b.e > a.e || (b.e == a.e && b.m > a.m)) { SlabX87 t = a; a = b; b = t; }
  const int shift = a.e - b.e;
  /* 128-bit fixed point: the significand in the high word */
  unsigned long long ah = a.m, al = 0, bh = b.m, bl = 0;
  int sticky = 0;
  if (shift >= 128) { bh = 0; bl = 0; sticky = 1; }
  else if (shift > 0) slab_u128_shr(&bh, &bl, shift, &sticky);
  SlabX87 r; r.s = a.s; r.e = a.e;
  unsigned long long sh, sl;
  if (a.s == b.s) {
    sl = al + bl;
    const unsigned long long c0 = (sl < al) ? 1ull : 0ull;
    sh = ah + bh;
    unsigned long long c1 = (sh < ah) ? 1ull : 0ull;
    const unsigned long long sh2 = sh + c0;
    if (sh2 < sh) c1 = 1ull;
    sh = sh2;
    if (c1) {                                                 /* carry out of 128 bits */
      if (sl & 1ull) sticky = 1;
      sl = (sl >> 1) | (sh << 63);
      sh = (sh >> 1) | (1ull << 63);
      r.e++;
    }
  } else {
    /* a >= b in magnitude */
    sl = al - bl;
    const unsigned long long br = (al < bl) ? 1ull : 0ull;
    sh = ah - bh - br;
    if (sticky) {                                             /* the lost low bits of b borrow; they stay sticky */
      if (sl == 0) sh--;
      sl--;
    }
    if (sh == 0 && sl == 0 && !sticky) { r.m = 0; r.e = 0; r.s = 0; return r; }
  }
  /* normalise */
  if (sh == 0) { sh = sl; sl = 0; r.e -= 64; }
  if (sh == 0) { r.m = 0; r.e = 0; r.s = 0; return r; }
  const int lz = slab_clz64(sh);
  if (lz) { sh = (sh << lz) | (sl >> (64 - lz)); sl <<= lz; r.e -= lz; }
  /* round to nearest even on the 64-bit significand */
  const unsigned long long half = 1ull << 63;
  const int above = sl > half || (sl == half && sticky);
  const int tie = sl == half && !sticky;
  if (above || (tie && (sh & 1ull))) {
    sh++;
    if (sh == 0) { sh = half; r.e++; }
  }
  r.m = sh;
  return r;
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

/* ---- slicing-by-4 for the same CRC: four 256-entry tables in shared memory, T[k][x] = register after
 * byte x followed by k zero bytes.  One 32-bit word of the stream costs four table loads and a few
 * XORs instead of 32 shift/XOR steps. ---- */
struct SlabCrcTables { uint16_t t[4][256]; };

/* every thread of the CTA must call this; ends with a barrier */
__device__ __forceinline__ void slab_crc16_build_tables(SlabCrcTables* tb)
{
  for (uint32_t x = threadIdx.x; x < 256u; x += blockDim.x) {
    uint32_t c = slab_crc16_byte(0u, x);
    tb->t[0][x] = (uint16_t)c;
#pragma unroll
    for (int k = 1; k < 4; k++) { c = slab_crc16_byte(c, 0u); tb->t[k][x] = (uint16_t)c; }
  }
  __syncthreads();
}
__device__ __forceinline__ uint32_t slab_crc16_step1(const SlabCrcTables* tb, uint32_t crc, uint32_t byte)
{
  return (uint32_t)tb->t[0][(crc ^ byte) & 0xFFu] ^ (crc >> 8);
}
/* w = four stream bytes, first byte in the low bits */
__device__ __forceinline__ uint32_t slab_crc16_step4(const SlabCrcTables* tb, uint32_t crc, uint32_t w)
{
  const uint32_t x = crc ^ w;
  return (uint32_t)tb->t[3][x & 0xFFu] ^ (uint32_t)tb->t[2][(x >> 8) & 0xFFu] ^
         (uint32_t)tb->t[1][(x >> 16) & 0xFFu] ^ (uint32_t)tb->t[0][x >> 24];
}
/* CRC register after bytes p[0 .. n), starting from 0; 128-bit loads once p is 16-byte aligned */
__device__ __forceinline__ uint32_t slab_crc16_run(const SlabCrcTables* tb, const uint8_t* __restrict__ p, uint32_t n)
{
  uint32_t crc = 0, i = 0;
  const uint32_t head = (16u - (uint32_t)((size_t)p & 15u)) & 15u;
  for (; i < n && i < head; i++) crc = slab_crc16_step1(tb, crc, p[i]);
  for (; i + 16u <= n; i += 16u) {
    const uint4 v = *reinterpret_cast<const uint4*>(p + i);
    crc = slab_crc16_step4(tb, crc, v.x);
    crc = slab_crc16_step4(tb, crc, v.y);
    crc = slab_crc16_step4(tb, crc, v.z);
    crc = slab_crc16_step4(tb, crc, v.w);
  }
  for (; i < n; i++) crc = slab_crc16_step1(tb, crc, p[i]);
  return crc;
}

/* ---------------- MSB-first bit reader over a 64-byte aligned, zero-padded device stream --------- */
/* One reader per lane; every lane of a warp walks its own block.  The lane's part of the stream is
 * pulled into a private 1 KiB ring in shared memory by 64-byte groups of cp.async copies issued a
 * whole top-up period ahead of their use, so global-memory latency never reaches the decode
 * recurrence.  The bit window is three consecutive stream words in registers - w0 and w1 in value
 * (big-endian) order, w2r still as loaded - plus the number of bits of w0 consumed so far: the next
 * 32 bits are one funnel shift away, and consuming up to 32 bits is an add, a compare and a
 * branch-free rotate.  The rotate's refill (word widx + 2, a shared-memory load) is not looked at
 * until the following rotate, so its latency is off the recurrence as well.
 *
 * Protocol (per lane): init(); then topup() at least once per SLAB_BR_PERIOD_BYTES consumed. */
#define SLAB_BR_RING         1024u              /* bytes per lane; rings are 16-byte aligned */
#define SLAB_BR_CHUNK        64u                /* copy granularity */
#define SLAB_BR_CHUNKS       (SLAB_BR_RING / SLAB_BR_CHUNK)
#define SLAB_BR_PERIOD_BYTES 448u               /* 32 well-formed codes of at most 14 bytes */
#define SLAB_BR_NEED_CHUNKS  (SLAB_BR_PERIOD_BYTES / SLAB_BR_CHUNK + 1u)

__device__ __forceinline__ uint32_t slab_shr_c(uint32_t v, uint32_t n)      /* v >> n, n in [0, 32] */
{
#ifdef SLAB_EMUL
  return n >= 32u ? 0u : v >> n;
#else
  return __funnelshift_rc(v, 0u, n);
#endif
}
/* position of the most significant set bit, 0xffffffff for 0 (one FLO) */
__device__ __forceinline__ uint32_t slab_msb(uint32_t v)
{
#ifdef SLAB_EMUL
  return v ? 31u - (uint32_t)__builtin_clz(v) : 0xffffffffu;
#else
  uint32_t r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
#endif
}
/* number of leading zeros for v != 0, 0xffffffff for 0 (one FLO.SH) */
__device__ __forceinline__ uint32_t slab_lz_nonzero(uint32_t v)
{
#ifdef SLAB_EMUL
  return v ? (uint32_t)__builtin_clz(v) : 0xffffffffu;
#else
  uint32_t r;
  asm("bfind.shiftamt.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
#endif
}

#ifdef SLAB_EMUL
typedef unsigned char* slab_ring_t;
#else
typedef uint32_t slab_ring_t;                   /* shared-window address */
#endif

__device__ __forceinline__ void slab_br_copy_chunk(slab_ring_t ring, const unsigned char* src, uint32_t chunk)
{
  const uint32_t off = (chunk * SLAB_BR_CHUNK) & (SLAB_BR_RING - 1u);
  const unsigned char* g = src + (size_t)chunk * SLAB_BR_CHUNK;
#ifdef SLAB_EMUL
  memcpy(ring + off, g, SLAB_BR_CHUNK);
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\t"
               "cp.async.cg.shared.global [%0+16], [%1+16], 16;\n\t"
               "cp.async.cg.shared.global [%0+32], [%1+32], 16;\n\t"
               "cp.async.cg.shared.global [%0+48], [%1+48], 16;"
               :: "r"(ring + off), "l"(g) : "memory");
#endif
}
/* fill every free slot of a lane's ring and wait for the copies (start of a block; rare afterwards).
 * Kept out of line: the readers call it from rarely taken branches only. */
static __device__ __noinline__ uint32_t slab_br_fill(uint32_t fetch, uint32_t widx, uint32_t nchunks, slab_ring_t ring, const unsigned char* src)
{
#pragma unroll 1
  while (fetch - (widx >> 4) < SLAB_BR_CHUNKS && fetch < nchunks) { slab_br_copy_chunk(ring, src, fetch); fetch++; }
  slab_cp_async_commit();
  slab_cp_async_wait<0>();
  return fetch;
}

struct SlabBitReader {
  uint32_t w0, w1;            /* stream words widx, widx + 1 in value order */
  uint32_t w2r;               /* stream word widx + 2 as loaded (byte-swapped when it becomes w1) */
  uint32_t o;                 /* bits of w0 already consumed: 0..31 */
  uint32_t widx;
  uint32_t fetch;             /* next 64-byte chunk of the stream to copy into the ring */
  uint32_t nchunks;           /* chunks in the stream image; nothing is fetched beyond them */
  slab_ring_t ring;           /* this lane's ring */
  const unsigned char* src;   /* stream image (64-byte aligned, zero padded) */

  __device__ __forceinline__ uint32_t ring_word(uint32_t w) const
  {
    uint32_t v;
#ifdef SLAB_EMUL
    memcpy(&v, ring + ((w << 2) & (SLAB_BR_RING - 1u)), 4);
#else
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ring + ((w << 2) & (SLAB_BR_RING - 1u))));
#endif
    return v;
  }
  __device__ __forceinline__ void fill() { fetch = slab_br_fill(fetch, widx, nchunks, ring, src); }
  /* Everything issued by earlier calls has had a whole period to land, so the wait is free; then
   * refill up to two of the chunks freed since.  If less than one period's worth had been fetched
   * (a lane that outran 128 bytes per period) fill the ring and wait for it. */
  __device__ __forceinline__ void topup()
  {
    slab_cp_async_wait<0>();
    const uint32_t have = fetch - (widx >> 4);
#pragma unroll
    for (int j = 0; j < 2; j++) {
      if (fetch - (widx >> 4) < SLAB_BR_CHUNKS && fetch < nchunks) { slab_br_copy_chunk(ring, src, fetch); fetch++; }
    }
    slab_cp_async_commit();
    if (have < SLAB_BR_NEED_CHUNKS && fetch < nchunks) fill();
  }
  __device__ __forceinline__ void init(unsigned char* lane_ring, const void* stream, uint32_t total_words, uint64_t byte_off)
  {
#ifdef SLAB_EMUL
    ring = lane_ring;
#else
    ring = (uint32_t)__cvta_generic_to_shared(lane_ring);
#endif
    src = reinterpret_cast<const unsigned char*>(stream); nchunks = total_words >> 4;
    widx = (uint32_t)(byte_off >> 2);
    o = (uint32_t)(byte_off & 3u) * 8u;
    fetch = widx >> 4;
    fill();
    w0 = __byte_perm(ring_word(widx), 0, 0x0123); w1 = __byte_perm(ring_word(widx + 1u), 0, 0x0123);
    w2r = ring_word(widx + 2u);
  }
  /* the next 32 bits of the stream */
  __device__ __forceinline__ uint32_t window() const { return __funnelshift_l(w1, w0, o); }
  /* drop n <= 32 bits.  Written without a branch: the rotate is two selects, the refill load is
   * predicated by hand (the compiler would otherwise wrap the whole rotate in a divergent branch). */
  __device__ __forceinline__ void advance(uint32_t n)
  {
    o += n;
    const uint32_t adv = o >> 5;                               /* 0 or 1 */
    widx += adv;
    o &= 31u;
    w0 = adv ? w1 : w0;
    w1 = adv ? __byte_perm(w2r, 0, 0x0123) : w1;
#ifdef SLAB_EMUL
    if (adv) w2r = ring_word(widx + 2u);
#else
    /* volatile: keeps its order against the (volatile) cp.async wait / commit of topup() and fill() */
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.u32 %0, [%1];\n\t}"
        : "+r"(w2r) : "r"(ring + (((widx + 2u) << 2) & (SLAB_BR_RING - 1u))), "r"(adv));
#endif
  }
  /* n in [0, 32] */
  __device__ __forceinline__ uint32_t get(uint32_t n)
  {
    const uint32_t v = slab_shr_c(window(), 32u - n);           /* n == 0 -> 0 */
    advance(n);
    return v;
  }
  /* zeros before the next 1 bit; the 1 is consumed (SLABitReader_GetZeroRunLength).  Runs longer
   * than a window only occur in damaged streams: they are followed to the end of the stream image. */
  __device__ __forceinline__ uint32_t zero_run()
  {
    uint32_t run = 0;
#pragma unroll 1
    for (;;) {
      const uint32_t w = window();
      if (w != 0u) {
        const uint32_t lz = (uint32_t)__clz((int)w);
        advance(lz + 1u);
        return run + lz;
      }
      run += 32u; advance(32u);
      if ((widx >> 4) >= nchunks) return run;
      if ((widx & 15u) == 0u) fill();
    }
  }
  __device__ __forceinline__ void align_byte() { advance((8u - (o & 7u)) & 7u); }
  /* bytes consumed since the stream start, rounding a partial byte up */
  __device__ __forceinline__ uint64_t byte_pos() const
  {
    return ((uint64_t)widx * 32u + o + 7u) >> 3;
  }
};

__device__ __forceinline__ uint32_t slab_rice_k32(uint32_t p)
{
  const uint32_t t = (p > 256u ? p : 256u) - 256u;
  const int32_t k = (int32_t)slab_msb(t) - 8;
  return (uint32_t)(k < 0 ? 0 : k);
}

#endif
