/*
 * slab_decode_fused.cu - D1b + D2 in one kernel (reference: src/SLADecoder.c:309-537).
 *
 * Entropy decode (one lane per block) and the synthesis cascade (one lane per block x channel) are
 * both pure recurrences whose run time is a dependent-instruction chain, not bandwidth; run one after
 * the other they cost the sum of the two chains and move the residual through HBM in between.  Here a
 * CTA takes 32 blocks: warp 0 decodes their bit streams tile by tile into shared memory, warps
 * 1..NCH (one per channel) run LMS -> long-term -> PARCOR -> de-emphasis on the tile decoded one step
 * earlier, so the two chains overlap and the residual never leaves the SM.
 *
 *   iteration t:   warp 0      decodes tile t      -> tile buffer t & 1
 *                  warps 1..   synthesise tile t-1 <- tile buffer (t-1) & 1 -> work planes (HBM)
 *                  __syncthreads()
 *
 * Instantiated for 1, 2 and 8 channels and the coefficient-count classes of every reference preset; other
 * parameter sets take the separate kernels in slab_decode.cu.
 */
#include "slab_decode_kernels.cuh"

#include <string.h>

template <int NCH> struct FusedGeom {
  static constexpr uint32_t TS = (NCH <= 2) ? 64u : 32u;      /* samples per tile */
  static constexpr uint32_t ROW = TS + 1u;                    /* odd stride: lanes hit distinct banks */
  static constexpr size_t   TILE_WORDS = 2u * NCH * 32u * ROW;
  static constexpr size_t   SMEM = 32u * SLAB_BR_RING + TILE_WORDS * sizeof(int32_t);
};

struct DeTileSink {
  int32_t* row0;            /* this lane's row of channel 0 in the current buffer, minus the tile start */
  uint32_t chan_stride;     /* words between the rows of consecutive channels */
  __device__ __forceinline__ void put(int c, uint32_t s, int32_t v) const { row0[(uint32_t)c * chan_stride + s] = v; }
};

template <int NCH, int LMS_N, int PMAX, int TAPS>
__global__ void __launch_bounds__(32 * (NCH + 1)) k_dec_block(const uint32_t* __restrict__ words, DecShape sh,
    const uint32_t* __restrict__ blk_off, const uint32_t* __restrict__ blk_pst,
    const uint32_t* __restrict__ blk_n,
    int32_t* work, int32_t* scratch, uint32_t* type_out, int32_t* kq_out, int32_t* ltq_out, uint32_t* pitch_out,
    uint32_t* err)      /* no __restrict__: the entropy warp writes what the synthesis warps read */
{
  typedef FusedGeom<NCH> G;
  SLAB_DYN_SMEM(unsigned char, smem);            /* rings (1 KiB aligned) | tiles */
  __shared__ uint32_t s_n[32], s_mode[32], s_ntiles;
  unsigned char* rings = smem;
  int32_t* tiles = reinterpret_cast<int32_t*>(smem + 32u * SLAB_BR_RING);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t b = blockIdx.x * 32u + lane;
  const bool valid = b < sh.nblocks;
  DeOutArrays o; o.type = type_out; o.kq = kq_out; o.ltq = ltq_out; o.pitch = pitch_out; o.err = err;

  if (warp == 0) {
    /* ---------------- entropy warp ---------------- */
    DeLane<NCH> L;
    L.mode = DE_IDLE; L.n = 0;
    if (valid) L.begin(rings + lane * SLAB_BR_RING, words, sh, b, blk_off, blk_n, o);
    const uint32_t n = (L.mode != DE_IDLE) ? L.n : 0u;
    s_n[lane] = n; s_mode[lane] = L.mode;
    const uint32_t nt = slab_warp_max((n + G::TS - 1u) / G::TS);
    if (lane == 0) s_ntiles = nt;
    __syncthreads();                              /* headers parsed: coefficients are in global memory */
    for (uint32_t t = 0; t <= nt; t++) {
      const uint32_t s0 = t * G::TS;
      if (t < nt && s0 < n) {
        DeTileSink sink;
        sink.row0 = tiles + ((size_t)(t & 1u) * NCH * 32u + lane) * G::ROW - s0;
        sink.chan_stride = 32u * G::ROW;
        L.span(s0, (n - s0 < G::TS) ? n : s0 + G::TS, sink);
      }
      __syncthreads();
    }
    if (L.mode != DE_IDLE) L.finish(o);
  } else {
    /* ---------------- synthesis warps: warp w handles channel w - 1 of the CTA's 32 blocks ---------------- */
    const uint32_t c = warp - 1u;
    __syncthreads();
    const uint32_t n = s_n[lane], mode = s_mode[lane], nt = s_ntiles;
    const uint32_t bc = b * NCH + c;
    int32_t* row = work + (size_t)c * sh.NP + (valid ? blk_pst[b] : 0u);
    const bool coded = (mode == DE_RICE || mode == DE_GOLOMB);
    const bool filter_on = valid && coded && err[b] == 0;       /* a block that failed its CRC stays a residual */
    SynthLane<LMS_N, PMAX, TAPS> S;
    if (filter_on)
      S.begin(sh, bc, n, row, scratch + (size_t)c * sh.NP + blk_pst[b], kq_out, ltq_out, pitch_out);
    for (uint32_t t = 0; t <= nt; t++) {
      if (t >= 1u) {
        const uint32_t s0 = (t - 1u) * G::TS;
        const int32_t* src = tiles + ((size_t)((t - 1u) & 1u) * NCH * 32u + (size_t)c * 32u + lane) * G::ROW;
        if (filter_on) {
#pragma unroll 1
          for (uint32_t u = 0; u < G::TS; u += LMS_N) S.chunk(s0 + u, src + u);
        } else if (mode != DE_IDLE && s0 < n) {
          const uint32_t cnt = (n - s0 < G::TS) ? n - s0 : G::TS;
          for (uint32_t u = 0; u < cnt; u++) row[s0 + u] = src[u];    /* raw samples (or an unfiltered residual) */
        }
      }
      __syncthreads();
    }
  }
}

/* ------------------------------------------------------------------ launch ---- */
struct FusedArgs {
  const uint32_t* words; const uint32_t* blk_off; const uint32_t* blk_pst; const uint32_t* blk_n;
  int32_t* work; int32_t* scratch; uint32_t* type; int32_t* kq; int32_t* ltq; uint32_t* pitch; uint32_t* err;
};

template <int NCH, int LMS_N, int PMAX, int TAPS>
static int launch_one(SlabCtx* ctx, const DecShape& sh, const FusedArgs& a)
{
  typedef FusedGeom<NCH> G;
  auto kp = k_dec_block<NCH, LMS_N, PMAX, TAPS>;
  if (slab_opt_in_smem(kp, G::SMEM) != 0) return -1;
  SLAB_RUN(ctx, "D1b+D2 k_dec_block", kp, slab_div_up(sh.nblocks, 32), 32 * (NCH + 1), G::SMEM, a.words, sh, a.blk_off,
           a.blk_pst, a.blk_n, a.work, a.scratch, a.type, a.kq, a.ltq, a.pitch, a.err);
  return 0;
}

template <int NCH, int LMS_N, int TAPS>
static int launch_p(SlabCtx* ctx, const DecShape& sh, int pmax, const FusedArgs& a)
{
  switch (pmax) {
    case 8:  return launch_one<NCH, LMS_N, 8, TAPS>(ctx, sh, a);
    case 16: return launch_one<NCH, LMS_N, 16, TAPS>(ctx, sh, a);
    default: return launch_one<NCH, LMS_N, 32, TAPS>(ctx, sh, a);
  }
}

template <int NCH>
static int launch_n(SlabCtx* ctx, const DecShape& sh, int pmax, const FusedArgs& a)
{
  if (sh.lms == 4) return sh.T <= 1 ? launch_p<NCH, 4, 1>(ctx, sh, pmax, a) : launch_p<NCH, 4, 3>(ctx, sh, pmax, a);
  return sh.T <= 1 ? launch_p<NCH, 8, 1>(ctx, sh, pmax, a) : launch_p<NCH, 8, 3>(ctx, sh, pmax, a);
}

/* 1 = not applicable to this parameter set (the caller runs the separate kernels), 0 = launched, -1 = error */
int slab_decode_fused(SlabCtx* ctx, const DecShape& sh, int pmax, const uint32_t* words, const uint32_t* blk_off,
    const uint32_t* blk_pst, const uint32_t* blk_n, int32_t* work, int32_t* scratch, uint32_t* type, int32_t* kq,
    int32_t* ltq, uint32_t* pitch, uint32_t* err)
{
  if (!((sh.lms == 4 || sh.lms == 8) && pmax <= 32 && sh.T <= 3 && (sh.nch == 1 || sh.nch == 2 || sh.nch == 8))) return 1;
  FusedArgs a = { words, blk_off, blk_pst, blk_n, work, scratch, type, kq, ltq, pitch, err };
  if (sh.nch == 8) return launch_n<8>(ctx, sh, pmax, a);
  return sh.nch == 1 ? launch_n<1>(ctx, sh, pmax, a) : launch_n<2>(ctx, sh, pmax, a);
}
