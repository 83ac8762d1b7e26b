/* slab_encode_kernels2.cuh - encoder kernels E4..E10 (see slab_encode_kernels.cuh for the map). */
#ifndef SLAB_ENCODE_KERNELS2_CUH
#define SLAB_ENCODE_KERNELS2_CUH

#include "slab_encode_kernels.cuh"
#include "slab_lanestream.cuh"

/* per block x channel analysis results */
struct EncChan {
  uint32_t flags;      /* bit0: some sample non-zero, bit1: estimated ratio >= 0.95 (RAW) */
  uint32_t rshift;
  uint32_t pitch;      /* 0 = long-term stage unused */
  uint32_t rice_init;  /* Q8 parameter as the reference stores it: (uint32)(mean << 8) */
  unsigned long long zsum;   /* sum of zig-zagged residuals */
  unsigned long long bits;   /* entropy-coded bits of this channel */
};

/* double -> int32 the way x86-64 cvttsd2si does it (NaN / out of range -> INT_MIN), because the
 * reference casts Levinson output unchecked (SLAEncoder.c:578-582) */
__device__ __forceinline__ int32_t enc_d2i_x86(double v)
{
  if (!(v > -2147483649.0 && v < 2147483648.0)) return (int32_t)0x80000000u;
  return (int32_t)v;
}
/* SLAUtility_Round, SLAUtility.c:436-439 */
__device__ __forceinline__ double enc_round(double d) { return (d >= 0.0) ? floor(d + 0.5) : -floor(-d + 0.5); }

/* a fixed value in [-1, 1) per index that follows no pattern a signal could share */
__device__ __forceinline__ double enc_unstructured(uint32_t k)
{
  return (double)(((k + 1u) * 2654435761u >> 16) & 0xFFFFu) * (1.0 / 32768.0) - 1.0;
}

template <typename T> __device__ __forceinline__ T enc_warp_sum(T v)
{
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(SLAB_FULL_MASK, v, d);
  return v;
}

/* ------------------------------------------------------------------------------------ E4 */
/* E4a, one CTA per block x channel: d[n] = windowed double signal in shared memory; the pre-emphasised
 * value e[n] = d[n] - d[n-1] * 31/32 is formed when it enters a thread's register window, so both are
 * bit-identical to the reference's input_double after SLAEncoder.c:540-543.  Autocorrelation by a
 * fixed-order parallel reduction (the reference's folded serial order differs by ~1e-13 relative):
 * LAGS > 0: every thread owns a contiguous run of samples and slides a LAGS-wide register window over
 * it (one shared-memory load per LAGS FMAs, window rotation resolved at compile time);
 * LAGS == 0: generic strided fallback for orders above 32. */
template <int LAGS, bool EXACT = false>
__global__ void __launch_bounds__(256) k_enc_autocorr(InPtrs in, EncShape sh,
    const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ blk_len,
    const uint32_t* __restrict__ blk_flag, const double* const* __restrict__ blk_win,
    double* __restrict__ acorr_out, uint32_t* __restrict__ maxabs_out, const uint32_t* __restrict__ only)
{
  /* EXACT: the listed block x channels only (only[bc] != 0), lag sums in the reference's own order */
  if (EXACT && only[blockIdx.x] == 0u) return;
  SLAB_DYN_SMEM(double, dsm);                   /* dsm[0] = d[-1] = 0, dsm[i + 1] = d[i] */
  __shared__ double red[8];
  __shared__ double part[8 * (LAGS > 0 ? LAGS : 1)];
  __shared__ uint32_t red_u[8];
  const uint32_t bc = blockIdx.x, b = bc / sh.nch, c = bc - b * sh.nch, tid = threadIdx.x;
  const uint32_t lane = tid & 31u, wid = tid >> 5;
  if (blk_flag[b] != 0) {                       /* leading-silence block: all zero by construction */
    if (tid == 0) maxabs_out[bc] = 0;
    return;
  }
  const uint32_t n = blk_len[b];
  const size_t s0 = blk_start[b];
  const uint32_t shift = 32u - sh.bits + enc_lshift(sh, b);
  const double* win = blk_win[b];
  const double emph = 0.96875;                  /* (2^5 - 1) * 2^-5, SLAPredictor.c:1803 */
  const double two_m31 = 4.656612873077392578125e-10;
  constexpr uint32_t PAD = 2u * (LAGS > 0 ? LAGS : 1) + 4u;
  uint32_t maxabs = 0;
  if (tid == 0) dsm[0] = 0.0;
  /* staging: eight samples per thread are requested from every source (both input planes, the window
   * table) before the first one is converted, so the loads of a thread overlap */
  const int32_t* pa = in.p[sh.ms ? 0 : c] + s0;
  const int32_t* pb = in.p[sh.ms ? 1 : c] + s0;
  for (uint32_t i0 = tid; i0 < n; i0 += 8u * 256u) {
    int32_t ra[8], rb[8];
    double wv[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t i = i0 + 256u * (uint32_t)j;
      ra[j] = 0; rb[j] = 0; wv[j] = 1.0;
      if (i < n) {
        ra[j] = pa[i];
        if (sh.ms) rb[j] = pb[i];
        if (win) wv[j] = win[i];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t i = i0 + 256u * (uint32_t)j;
      if (i < n) {
        double cur;
        int32_t xi;
        if (!sh.ms) {
          xi = ra[j] >> shift;
          cur = (double)ra[j] * two_m31;
        } else {
          const int32_t l = ra[j] >> shift, r = rb[j] >> shift;
          xi = (c == 0) ? ((l + r) >> 1) : (l - r);                      /* SLAUtility.c:403-404 */
          const double dl = (double)ra[j] * two_m31, dr = (double)rb[j] * two_m31;
          cur = (c == 0) ? (dl + dr) / 2 : (dl - dr);                    /* SLAUtility.c:381-385 */
        }
        const uint32_t a = (xi < 0) ? (0u - (uint32_t)xi) : (uint32_t)xi;
        maxabs = a > maxabs ? a : maxabs;
        if (win) cur *= wv[j];
        dsm[i + 1u] = cur;
      }
    }
  }
  for (uint32_t i = n + tid; i < n + PAD; i += 256) dsm[i + 1u] = 0.0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { const uint32_t o = __shfl_xor_sync(SLAB_FULL_MASK, maxabs, d); maxabs = o > maxabs ? o : maxabs; }
  if (lane == 0) red_u[wid] = maxabs;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; w++) maxabs = red_u[w] > maxabs ? red_u[w] : maxabs;
    maxabs_out[bc] = maxabs;
  }
  const uint32_t lags = sh.P + 1u;
  double* out = acorr_out + (size_t)bc * (SLAB_MAX_PARCOR + 1);
  /* e(j) for j in [0, n), 0 beyond */
#define EMPH_AT(j) (((j) < n) ? (dsm[(j) + 1u] - dsm[(j)] * emph) : 0.0)
  if (EXACT) {
    /* LPC_CalculateAutoCorrelation, SLAPredictor.c:331-388, term by term: one thread per lag, one serial
     * accumulator - the folded pairs first (i outer, l inner), then the plain tail.  The doubles are the
     * reference's bit for bit; used for the block x channels whose Levinson recursion is so badly
     * conditioned that the rounding of a re-ordered sum would reach the quantised coefficients.
     * The pre-emphasised signal replaces the windowed one in place first (every thread a contiguous run,
     * right to left, its left neighbour saved beforehand). */
    {
      const uint32_t run = (n + 255u) / 256u, a = tid * run;
      const uint32_t b_end = (a + run < n) ? a + run : n;
      const double left = (a < n) ? dsm[a] : 0.0;              /* d[a - 1] */
      __syncthreads();
      for (uint32_t j = b_end; j > a; j--) {
        const uint32_t i = j - 1u;
        dsm[i + 1u] = dsm[i + 1u] - ((i == a) ? left : dsm[i]) * emph;
      }
      __syncthreads();
    }
#define E_AT(j) (((j) < n) ? dsm[(j) + 1u] : 0.0)
    if (tid < lags) {
      const uint32_t lag = tid;
      double acc = 0.0;
      if (lag == 0u) {
        for (uint32_t i = 0; i < n; i++) { const double e = E_AT(i); acc += e * e; }
      } else if (lag < n) {
        const uint32_t two = lag << 1;
        const uint32_t groups = (3u * lag < n) ? 1u + (n - 3u * lag) / two : 0u;
        const uint32_t span = groups * two;
        for (uint32_t i = 0; i < lag; i++)
          for (uint32_t l = 0; l < span; l += two)
            acc += E_AT(l + lag + i) * (E_AT(l + i) + E_AT(l + two + i));
        for (uint32_t i = 0; i < n - span - lag; i++) acc += E_AT(span + lag + i) * E_AT(span + i);
      }
      out[tid] = acc;
    }
#undef E_AT
  } else if (LAGS > 0) {
    constexpr int LG = LAGS > 0 ? LAGS : 1;
    uint32_t run = (n + 255u) / 256u;
    run |= 1u;                                     /* odd stride: no systematic bank conflicts */
    const uint32_t lo = tid * run;
    const uint32_t hi = (lo + run < n) ? lo + run : n;
    double acc[LG], w[LG];
#pragma unroll
    for (int k = 0; k < LG; k++) { acc[k] = 0.0; w[k] = (lo < n) ? EMPH_AT(lo + k) : 0.0; }
    for (uint32_t i = lo; i < hi; i += LG) {
#pragma unroll
      for (int r = 0; r < LG; r++) {
        if (i + r < hi) {
          const double x = w[r];
#pragma unroll
          for (int k = 0; k < LG; k++) acc[k] = fma(x, w[(r + k) % LG], acc[k]);
          w[r] = EMPH_AT(i + r + LG);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < LG; k++) {
      const double v = enc_warp_sum(acc[k]);
      if (lane == 0) part[wid * LG + k] = v;
    }
    __syncthreads();
    if (tid < lags) {
      double s = 0.0;
      for (int w8 = 0; w8 < 8; w8++) s += part[w8 * LG + tid];
      out[tid] = s;
    }
  } else {
    for (uint32_t k = 0; k < lags; k++) {
      double acc = 0.0;
      if (n > k) for (uint32_t i = tid; i < n - k; i += 256) acc = fma(EMPH_AT(i), EMPH_AT(i + k), acc);
      acc = enc_warp_sum(acc);
      if (lane == 0) red[wid] = acc;
      __syncthreads();
      if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += red[w];
        out[k] = s;
      }
      __syncthreads();
    }
  }
#undef EMPH_AT
}

/* E4b, one thread per block x channel: Levinson-Durbin, code-length estimate (RAW decision), bit
 * width -> rshift, coefficient quantisation (SLAEncoder.c:546-589).
 * The autocorrelation it starts from is a re-ordered (parallel) sum: about 1e-14 relative away from the
 * reference's serial one.  When `risk` is given, the recursion is repeated on an autocorrelation moved by
 * 1e-12 R(0) - a hundred times that - (k_enc_lpc<true>) and a block x channel whose quantised codes or RAW
 * decision move with it is flagged: k_enc_autocorr<.., true> then redoes its lag sums in the reference's
 * order and k_enc_lpc<false> runs again on the flagged ones (`only`). */
template <bool RISK>      /* RISK: nothing is written but risk[bc] - do the codes / the RAW decision move with the perturbation? */
__global__ void __launch_bounds__(64) k_enc_lpc(EncShape sh, uint32_t nblocks,
    const uint32_t* __restrict__ blk_len, const uint32_t* __restrict__ blk_flag,
    const double* __restrict__ acorr_in, const uint32_t* __restrict__ maxabs_in,
    EncChan* chan, double* parcor_out, int32_t* code_out, int32_t* kq_out, uint32_t* risk, const uint32_t* only,
    uint32_t* risk_count)
{
  const uint32_t bc = blockIdx.x * blockDim.x + threadIdx.x;
  if (bc >= nblocks * sh.nch) return;
  if (only != nullptr && only[bc] == 0u) return;
  const uint32_t b = bc / sh.nch;
  if (RISK) risk[bc] = 0u;
  if (blk_flag[b] != 0) {
    if (!RISK) { chan[bc].flags = 0; chan[bc].rshift = 0; chan[bc].pitch = 0; }
    return;
  }
  const uint32_t n = blk_len[b], maxabs = maxabs_in[bc];
  if (RISK && maxabs == 0) return;
  double R[SLAB_MAX_PARCOR + 2], a[SLAB_MAX_PARCOR + 2], t[SLAB_MAX_PARCOR + 2], parcor[SLAB_MAX_PARCOR + 1];
  for (uint32_t k = 0; k <= sh.P; k++) R[k] = acorr_in[(size_t)bc * (SLAB_MAX_PARCOR + 1) + k];
  if (RISK) {
    /* the offsets follow no pattern a signal could share (a uniform or alternating one would be a mere
     * rescaling for a DC or Nyquist tone and leave the recursion unchanged) */
    const double r0 = R[0];
    for (uint32_t k = 0; k <= sh.P; k++) R[k] = R[k] + 1e-12 * r0 * enc_unstructured(k);
  }
  enc_levinson(R, n, sh.P, parcor, a, t);
  double est = enc_code_length(R[0], n, sh.bits, parcor, sh.P);
  est = (8 * est) / sh.bits;
  uint32_t flags = (maxabs != 0) ? 1u : 0u;
  if (est >= (double)0.95f) flags |= 2u;                             /* SLAInternal.h:30 */
  const uint32_t bw = (maxabs > 0) ? slab_log2ceil(maxabs) + 1u : 1u;      /* SLAUtility.c:677-696 */
  const uint32_t rshift = (bw > 16u) ? bw - 16u : 0u;
  double* pd = parcor_out + (size_t)bc * (SLAB_MAX_PARCOR + 1);
  int32_t* pc = code_out + (size_t)bc * (SLAB_MAX_PARCOR + 1);
  int32_t* pk = kq_out + (size_t)bc * sh.pstride;
  uint32_t moved = 0;
  if (RISK) moved = (chan[bc].flags != flags) ? 1u : 0u;
  else { pd[0] = 0.0; pc[0] = 0; pk[0] = 0; }
  for (uint32_t k = 1; k <= sh.P; k++) {                             /* SLAEncoder.c:573-589 */
    const uint32_t qb = (k < 4u) ? 16u : 8u;
    const int32_t lim = 1 << (qb - 1u);
    int32_t q = enc_d2i_x86(enc_round(parcor[k] * exp2((double)(qb - 1u))));
    q = q < -lim ? -lim : q;
    q = q > lim - 1 ? lim - 1 : q;
    if (RISK) moved |= (pc[k] != q) ? 1u : 0u;
    else {
      pd[k] = parcor[k]; pc[k] = q;
      pk[k] = (int32_t)((uint32_t)q << (16u - qb)) >> rshift;
    }
  }
  if (RISK) { risk[bc] = moved; if (moved && risk_count) atomicAdd(risk_count, 1u); return; }
  for (uint32_t k = sh.P + 1u; k < sh.pstride; k++) pk[k] = 0;
  chan[bc].flags = flags; chan[bc].rshift = rshift; chan[bc].pitch = 0;
}

/* block type, SLAEncoder.c:520-528,562-565 */
__global__ void __launch_bounds__(128) k_enc_blocktype(EncShape sh, uint32_t nblocks,
    const uint32_t* __restrict__ blk_flag, const EncChan* __restrict__ chan, uint32_t* __restrict__ blk_type)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  uint32_t any_nz = 0, any_raw = 0;
  if (blk_flag[b] == 0)
    for (uint32_t c = 0; c < sh.nch; c++) { any_nz |= chan[b * sh.nch + c].flags & 1u; any_raw |= chan[b * sh.nch + c].flags & 2u; }
  blk_type[b] = !any_nz ? SLAB_BLOCK_SILENT : (any_raw ? SLAB_BLOCK_RAW : SLAB_BLOCK_COMPRESS);
}

/* ------------------------------------------------------------------------------------ E5 */
/* int32 pre-emphasis + PARCOR lattice analysis (SLAPredictor.c:1741-1765, 557-607).  The lattice is
 * feed-forward: f_P[n] depends on x[n-P-1 .. n] only, so every 256-sample slice restarts from a zero
 * state at least P samples early and is exact.  Four samples per step: 128-bit loads
 * of the input planes when the block is 16-byte aligned there, 128-bit stores into the padded
 * residual plane always. */
template <int PMAX>
__global__ void __launch_bounds__(128) k_enc_parcor(InPtrs in, EncShape sh, uint32_t nblocks, uint32_t spb,
    const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ blk_pst,
    const uint32_t* __restrict__ blk_len,
    const uint32_t* __restrict__ blk_type, const int32_t* __restrict__ kq_in, int32_t* __restrict__ r1)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t bc = (uint32_t)(t / spb), sl = (uint32_t)(t % spb);
  if (bc >= nblocks * sh.nch) return;
  const uint32_t b = bc / sh.nch, c = bc - b * sh.nch;
  if (blk_type[b] != SLAB_BLOCK_COMPRESS) return;
  const uint32_t n = blk_len[b], n0 = sl * SLAB_SLICE;
  if (n0 >= n) return;
  const uint32_t n1 = (n0 + SLAB_SLICE < n) ? n0 + SLAB_SLICE : n;
  const size_t s0 = blk_start[b];
  const uint32_t shift = 32u - sh.bits + enc_lshift(sh, b);
  int32_t kk[PMAX + 1], bw[PMAX + 1];
#pragma unroll
  for (int m = 0; m <= PMAX; m++) { kk[m] = kq_in[(size_t)bc * sh.pstride + m]; bw[m] = 0; }
  const uint32_t need = (sh.P + 3u) & ~3u;                  /* warm-up, kept a multiple of 4 */
  const uint32_t warm = (n0 < need) ? n0 : need;
  const uint32_t start = n0 - warm;
  int32_t prev = (start > 0) ? enc_sample(in, c, sh.ms, shift, s0 + start - 1) : 0;
  int4* dst = reinterpret_cast<int4*>(r1 + (size_t)c * sh.NP + blk_pst[b]);
  const int32_t* pl = in.p[sh.ms ? 0 : c];
  const int32_t* pr = in.p[sh.ms ? 1 : c];
  const bool vec = (((uintptr_t)(pl + s0) | (uintptr_t)(pr + s0)) & 15u) == 0;
  /* one lattice step over four consecutive samples; stores them once the warm-up is over */
#define PARCOR_STEP4(I, XS)                                                                      \
  do {                                                                                           \
    int32_t fo[4];                                                                               \
    _Pragma("unroll") for (int q = 0; q < 4; q++) {                                              \
      const int32_t x = (XS)[q];                                                                 \
      const int32_t y = (int32_t)((uint32_t)x - (uint32_t)slab_emph(prev));                      \
      prev = x;                                                                                  \
      int32_t f = y, b_old = bw[0];                                                              \
      _Pragma("unroll") for (int m = 1; m <= PMAX; m++) {                                        \
        const int32_t keep = bw[m];                                                              \
        const int32_t fm = f - slab_latmul(kk[m], b_old);                                        \
        bw[m] = b_old - slab_latmul(kk[m], f);                                                   \
        f = fm; b_old = keep;                                                                    \
      }                                                                                          \
      bw[0] = y;                                                                                 \
      fo[q] = f;                                                                                 \
    }                                                                                            \
    if ((I) >= n0) { int4 o; o.x = fo[0]; o.y = fo[1]; o.z = fo[2]; o.w = fo[3]; dst[(I) >> 2] = o; } \
  } while (0)
  uint32_t i = start;
  if (vec) {
    /* sixteen samples per iteration: all eight 128-bit loads are issued before the first lattice step,
     * so one memory round trip is paid per sixteen samples instead of per four */
    for (; i + 16u <= n1 && s0 + i + 16u <= sh.N; i += 16u) {
      int4 l[4], r[4];
#pragma unroll
      for (int g = 0; g < 4; g++) {
        l[g] = *reinterpret_cast<const int4*>(pl + s0 + i + 4 * g);
        if (sh.ms) r[g] = *reinterpret_cast<const int4*>(pr + s0 + i + 4 * g);
      }
#pragma unroll
      for (int g = 0; g < 4; g++) {
        int32_t xs[4];
        const int32_t la[4] = {l[g].x >> shift, l[g].y >> shift, l[g].z >> shift, l[g].w >> shift};
        if (!sh.ms) { xs[0] = la[0]; xs[1] = la[1]; xs[2] = la[2]; xs[3] = la[3]; }
        else {
          const int32_t ra[4] = {r[g].x >> shift, r[g].y >> shift, r[g].z >> shift, r[g].w >> shift};
#pragma unroll
          for (int q = 0; q < 4; q++) xs[q] = (c == 0) ? ((la[q] + ra[q]) >> 1) : (la[q] - ra[q]);
        }
        PARCOR_STEP4(i + 4u * (uint32_t)g, xs);
      }
    }
  }
  for (; i < n1; i += 4u) {
    int32_t xs[4];
#pragma unroll
    for (int q = 0; q < 4; q++) xs[q] = (s0 + i + q < sh.N) ? enc_sample(in, c, sh.ms, shift, s0 + i + q) : 0;
    PARCOR_STEP4(i, xs);
  }
#undef PARCOR_STEP4
}

/* ------------------------------------------------------------------------------------ E6 */
/* Small dense solve with the reference's pivoting quirks, SLAUtility.c:487-674, including the x87
 * long-double accumulation of the refinement residual (slab_common.cuh). */
__device__ inline int enc_lu_solve(double (*A)[SLAB_MAX_TAPS], double* bvec, uint32_t dim)
{
  double LU[SLAB_MAX_TAPS][SLAB_MAX_TAPS], scale[SLAB_MAX_TAPS], x[SLAB_MAX_TAPS], err[SLAB_MAX_TAPS];
  uint32_t piv[SLAB_MAX_TAPS];
  for (uint32_t r = 0; r < dim; r++) { for (uint32_t c = 0; c < dim; c++) LU[r][c] = A[r][c]; x[r] = bvec[r]; }
  for (uint32_t r = 0; r < dim; r++) {
    double big = 0.0;
    for (uint32_t c = 0; c < dim; c++) if (fabs(LU[r][c]) > big) big = fabs(LU[r][c]);
    if (fabs(big) <= (double)FLT_EPSILON) return -1;
    scale[r] = 1.0 / big;
  }
  for (uint32_t c = 0; c < dim; c++) {
    uint32_t r, best;
    double big = 0.0;
    for (r = 0; r < c; r++) {
      double s = LU[r][c];
      for (uint32_t k = 0; k < r; k++) s -= LU[r][k] * LU[k][c];
      LU[r][c] = s;
    }
    best = r;
    for (r = c; r < dim; r++) {
      double s = LU[r][c];
      for (uint32_t k = 0; k < c; k++) s -= LU[r][k] * LU[k][c];
      LU[r][c] = s;
      if (scale[r] * fabs(s) >= big) { big = scale[r] * fabs(s); best = r; }
    }
    if (c != best) {
      for (uint32_t k = 0; k < dim; k++) { const double tmp = LU[best][k]; LU[best][k] = LU[c][k]; LU[c][k] = tmp; }
      scale[best] = scale[c];
    }
    piv[c] = best;
    if (fabs(LU[c][c]) <= (double)FLT_EPSILON) return -1;
    if (c != dim - 1u) {
      const double inv = 1.0 / LU[c][c];
      for (r = c + 1u; r < dim; r++) LU[r][c] *= inv;
    }
  }
  for (uint32_t pass = 0; pass <= 2u; pass++) {
    double* v = (pass == 0) ? x : err;
    if (pass > 0) {
      for (uint32_t r = 0; r < dim; r++) {
        /* long double e = -b[r]; e += A[r][c] * x[c] (double products); err[r] = (double)e */
        SlabX87 ee = slab_x87_from_double(-bvec[r]);
        for (uint32_t c = 0; c < dim; c++) ee = slab_x87_add(ee, slab_x87_from_double(A[r][c] * x[c]));
        err[r] = slab_x87_to_double(ee);
      }
    }
    uint32_t first = 0;
    for (uint32_t r = 0; r < dim; r++) {
      const uint32_t p = piv[r];
      double s = v[p];
      v[p] = v[r];
      if (first != 0) { for (uint32_t c = first; c < r; c++) s -= LU[r][c] * v[c]; }
      else if (s != 0.0) { first = r; }
      v[r] = s;
    }
    for (uint32_t r = dim; r-- > 0; ) {
      double s = v[r];
      for (uint32_t c = r + 1u; c < dim; c++) s -= LU[r][c] * v[c];
      v[r] = s / LU[r][r];
    }
    if (pass > 0) for (uint32_t r = 0; r < dim; r++) x[r] -= err[r];
  }
  for (uint32_t r = 0; r < dim; r++) bvec[r] = x[r];
  return 0;
}

/* The pitch pick (SLAPredictor.c:867-924) with the scan over the lags done by a whole warp: every lane classifies
 * nine lags (negative -> positive crossing, positive -> negative crossing, local peak) into bit masks, then
 * lane 0 walks the lobes with bit scans instead of 256 dependent loads and compares.  Same decisions, same
 * order as the reference's loops, including its reads just past lag 255.
 * acs: the 260 lags in shared memory; masks: 3 x 9 words; cand: 264 entries.  Returns (in lane 0) the first
 * candidate reaching the maximum peak, 0 for a silent frame, 0xFFFFFFFF for "failed to calculate". */
__device__ __forceinline__ uint32_t enc_pitch_pick_warp(const double* acs, uint32_t* masks, uint16_t* cand, uint32_t lane)
{
  uint32_t* up = masks; uint32_t* down = masks + 9; uint32_t* peak = masks + 18;
#pragma unroll
  for (uint32_t w = 0; w < 9u; w++) {
    const uint32_t j = 32u * w + lane;
    bool u = false, d = false, pk = false;
    if (j >= 1u && j <= 258u) {
      const double c = acs[j], l = acs[j - 1u], r = acs[j + 1u];
      u = l < 0.0 && c > 0.0;
      d = c > 0.0 && r < 0.0;
      pk = c > l && c > r;
    }
    const uint32_t bu = __ballot_sync(SLAB_FULL_MASK, u), bd = __ballot_sync(SLAB_FULL_MASK, d), bp = __ballot_sync(SLAB_FULL_MASK, pk);
    if (lane == 0) { up[w] = bu; down[w] = bd; peak[w] = bp; }
  }
  __syncwarp();
  uint32_t result = 0;
  if (lane == 0) {
    /* first set bit of m at an index in [from, limit), or `none` */
    auto next_bit = [](const uint32_t* m, uint32_t from, uint32_t limit, uint32_t none) -> uint32_t {
      for (uint32_t w = from >> 5; w < 9u && 32u * w < limit; w++) {
        uint32_t bits = m[w];
        if (w == (from >> 5)) bits &= 0xFFFFFFFFu << (from & 31u);
        if (bits) { const uint32_t at = 32u * w + (uint32_t)(__ffs((int)bits) - 1); return at < limit ? at : none; }
      }
      return none;
    };
    if (fabs(acs[0]) <= (double)FLT_MIN) result = 0;
    else {
      uint32_t ncand = 0, i = 1;
      double peak_max = 0.0;
      while (i < SLAB_MAX_PITCH && ncand < SLAB_MAX_PITCH) {
        const uint32_t start = next_bit(up, i, SLAB_MAX_PITCH, SLAB_MAX_PITCH);
        uint32_t end = start + 1u;
        if (end < SLAB_MAX_PITCH) end = next_bit(down, end, SLAB_MAX_PITCH, SLAB_MAX_PITCH);
        uint32_t at = 0;
        double best = 0.0;
        for (uint32_t j = next_bit(peak, start, end + 1u, 0xFFFFu); j != 0xFFFFu; j = next_bit(peak, j + 1u, end + 1u, 0xFFFFu))
          if (acs[j] > best) { at = j; best = acs[j]; }
        if (at != 0) { cand[ncand++] = (uint16_t)at; if (best > peak_max) peak_max = best; }
        i = end + 1u;
      }
      if (ncand == 0) result = 0xFFFFFFFFu;
      else {
        uint32_t k = 0;
        for (; k < ncand; k++) if (acs[cand[k]] >= (double)1.0f * peak_max) break;
        result = cand[k];
      }
    }
  }
  return result;
}

/* Tap solve for the candidate the pitch pick chose (enc_pitch_pick_warp), SLAPredictor.c:855-865,926-977.
 * first_cand: the candidate, 0 for a silent frame, 0xFFFFFFFF when the pick found none.
 * returns 0 ok (pitch may be 0 for a silent frame), 1 "failed to calculate". */
__device__ __forceinline__ int enc_taps_solve(const double* ac, uint32_t taps, uint32_t first_cand, uint32_t* pitch, double* coef)
{
  *pitch = 0;
  for (uint32_t j = 0; j < SLAB_MAX_TAPS; j++) coef[j] = 0.0;
  if (first_cand == 0xFFFFFFFFu) return 1;
  if (fabs(ac[0]) <= (double)FLT_MIN) return 0;
  if (first_cand < taps / 2u + 1u) return 1;
  double Rm[SLAB_MAX_TAPS][SLAB_MAX_TAPS], v[SLAB_MAX_TAPS], mag = 0.0;
  for (uint32_t j = 0; j < taps; j++)
    for (uint32_t k = 0; k < taps; k++) Rm[j][k] = ac[(j >= k) ? (j - k) : (k - j)];
  for (uint32_t j = 0; j < taps; j++) v[j] = ac[j + first_cand - taps / 2u];
  if (enc_lu_solve(Rm, v, taps) != 0) return 1;
  for (uint32_t j = 0; j < taps; j++) mag += fabs(v[j]);
  if (mag >= 1.0) {
    for (uint32_t j = 0; j < taps; j++) v[j] = 0.0;
    v[taps / 2u] = ac[first_cand] / ac[0];
  }
  *pitch = first_cand;
  for (uint32_t j = 0; j < taps; j++) coef[j] = v[j];
  return 0;
}

#define LT_TILE     17u                                      /* lags per lane */
#define LT_GROUPS   16u                                      /* lanes of a half-warp: 16 x 17 = 272 lags >= 260 */
#define LT_PARTS    16u                                      /* sample ranges = half-warps of the CTA */
#define LT_LAGS_PAD (LT_GROUPS * LT_TILE)
#define LT_THREADS  (LT_GROUPS * LT_PARTS)

/* Lane g of a half-warp owns lags 17g .. 17g+16 over the half-warp's sample range.  Its window
 * w[0..16] holds y[i + 17g .. i + 17g + 16] at compile-time rotating slots.  At step i + r:
 *   x            = y[i + r]           = lane 0's w[r]                         -> one shuffle
 *   new w[r]     = y[i + r + 17(g+1)] = lane g+1's w[r] (the element it drops) -> one shuffle
 * so operands move between registers of neighbouring lanes (a systolic array); only lane 15 reads
 * shared memory, and - for the FP64 path - only it pays an integer-to-double conversion.  Every lane
 * then does 17 multiply-adds with nothing else on the arithmetic pipe.
 * All half-warps run the same number of steps (`steps`, a multiple of 17); ranges past the end of the
 * block read the zero padding. */
template <typename A, typename W>
__device__ __forceinline__ void lt_accumulate(const int32_t* y, uint32_t lo, uint32_t steps, uint32_t g, A* acc)
{
  W w[LT_TILE];
#pragma unroll
  for (int k = 0; k < (int)LT_TILE; k++) w[k] = (W)y[lo + g * LT_TILE + k];
  const int32_t* feed = y + lo + LT_GROUPS * LT_TILE;          /* what lane 15 pulls in: y[i + r + 272] */
  for (uint32_t i = 0; i < steps; i += LT_TILE) {
#pragma unroll
    for (int r = 0; r < (int)LT_TILE; r++) {
      const W x = __shfl_sync(SLAB_FULL_MASK, w[r], 0, 16);
#pragma unroll
      for (int k = 0; k < (int)LT_TILE; k++) {
        if (sizeof(A) == sizeof(double) && !std::is_integral<A>::value) acc[k] = (A)fma((double)x, (double)w[(r + k) % LT_TILE], (double)acc[k]);
        else acc[k] += (A)x * (A)w[(r + k) % LT_TILE];
      }
      W in = __shfl_down_sync(SLAB_FULL_MASK, w[r], 1, 16);
      if (g == LT_GROUPS - 1u) in = (W)feed[i + r];
      w[r] = in;
    }
  }
}

/* The same array on the 32-bit integer pipe, for blocks whose residual is small enough that the 17
 * products of one window turn cannot overflow an int32 (|r|^2 * 17 < 2^31, i.e. |r| < 11 239 - almost
 * every block of 16-bit material): IMAD issues at twice the rate of DFMA, the operands need one shuffle
 * instead of two, and nothing is converted.  After every turn the 17 partial sums are folded into the
 * 64-bit totals. */
__device__ __forceinline__ void lt_accumulate_i32(const int32_t* y, uint32_t lo, uint32_t steps, uint32_t g, long long* acc)
{
  int32_t w[LT_TILE];
#pragma unroll
  for (int k = 0; k < (int)LT_TILE; k++) w[k] = y[lo + g * LT_TILE + k];
  const int32_t* feed = y + lo + LT_GROUPS * LT_TILE;
  for (uint32_t i = 0; i < steps; i += LT_TILE) {
    int32_t a[LT_TILE];
#pragma unroll
    for (int k = 0; k < (int)LT_TILE; k++) a[k] = 0;
#pragma unroll
    for (int r = 0; r < (int)LT_TILE; r++) {
      const int32_t x = __shfl_sync(SLAB_FULL_MASK, w[r], 0, 16);
#pragma unroll
      for (int k = 0; k < (int)LT_TILE; k++) a[k] += x * w[(r + k) % LT_TILE];
      int32_t in = __shfl_down_sync(SLAB_FULL_MASK, w[r], 1, 16);
      if (g == LT_GROUPS - 1u) in = feed[i + r];
      w[r] = in;
    }
#pragma unroll
    for (int k = 0; k < (int)LT_TILE; k++) acc[k] += a[k];
  }
}

/* E6a, one CTA (256 threads) per block x channel: lags 0..259 of the PARCOR residual as exact integer
 * sums (the reference gets them, up to FFT round-off, from two 32768-point real FFTs), scaled like the
 * reference's un-normalised inverse transform so that its absolute thresholds apply unchanged.
 * half-warp = one sixteenth of the block, lane = 17 lags (see lt_accumulate).
 * Four arithmetic paths, chosen per block from max|r|: 32-bit IMAD with a 64-bit fold per window turn
 * while 17 products fit an int32 (exact; the fastest), FP64 FMA while every partial sum stays below
 * 2^53 (exact, and the FP64 pipe has twice the rate of IMAD.WIDE), int64 up to |r| < 2^24 (exact),
 * rounded double beyond. */
__global__ void __launch_bounds__(LT_THREADS) k_enc_ltcorr(EncShape sh,
    const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ blk_len,
    const uint32_t* __restrict__ blk_type, const int32_t* __restrict__ r1, double* __restrict__ ac_out,
    uint32_t* __restrict__ risk_list, uint32_t* __restrict__ risk_count, const uint32_t* __restrict__ only)
{
  if (only != nullptr && only[blockIdx.x] == 0u) return;       /* k_enc_ltcorr_mma has done this one */
  SLAB_DYN_SMEM(int32_t, y);
  __shared__ long long part_i[LT_PARTS][LT_LAGS_PAD];          /* doubles alias the same storage */
  __shared__ uint32_t red_u[16];
  __shared__ int s_risk;
  __shared__ uint32_t s_npeaks;
  __shared__ uint16_t s_peaks[264];
  const uint32_t bc = blockIdx.x, b = bc / sh.nch, c = bc - b * sh.nch, tid = threadIdx.x;
  if (blk_type[b] != SLAB_BLOCK_COMPRESS) return;
  const uint32_t n = blk_len[b];
  /* steps per half-warp: ceil(n / 16) rounded up to whole window turns */
  const uint32_t steps = (((n + LT_PARTS - 1u) / LT_PARTS + LT_TILE - 1u) / LT_TILE) * LT_TILE;
  const uint32_t filled = LT_PARTS * steps + LT_LAGS_PAD + LT_TILE;     /* everything any lane may read */
  const int32_t* src = r1 + (size_t)c * sh.NP + blk_start[b];        /* blk_start = padded starts here */
  uint32_t maxabs = 0;
#pragma unroll 4
  for (uint32_t i = tid; i < filled; i += blockDim.x) {
    const int32_t v = (i < n) ? src[i] : 0;
    y[i] = v;
    const uint32_t a = (v < 0) ? (0u - (uint32_t)v) : (uint32_t)v;
    maxabs = a > maxabs ? a : maxabs;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { const uint32_t o = __shfl_xor_sync(SLAB_FULL_MASK, maxabs, d); maxabs = o > maxabs ? o : maxabs; }
  if ((tid & 31u) == 0) red_u[tid >> 5] = maxabs;
  __syncthreads();
  maxabs = 0;
  for (uint32_t w = 0; w < (blockDim.x >> 5); w++) maxabs = red_u[w] > maxabs ? red_u[w] : maxabs;
  const bool fp_exact = (double)maxabs * (double)maxabs * (double)(steps + 1u) < 9007199254740992.0;   /* 2^53 */
  const bool int_exact = maxabs < (1u << 24);      /* products < 2^48, sums of <= 2^14 terms < 2^62 */
  {
    const uint32_t g = tid & 15u, p = tid >> 4, k0 = g * LT_TILE;
    const uint32_t lo = p * steps;
    if ((unsigned long long)maxabs * maxabs * LT_TILE < (1ull << 31)) {
      long long acc[LT_TILE];
#pragma unroll
      for (int k = 0; k < (int)LT_TILE; k++) acc[k] = 0;
      lt_accumulate_i32(y, lo, steps, g, acc);
#pragma unroll
      for (int k = 0; k < (int)LT_TILE; k++) part_i[p][k0 + k] = acc[k];
    } else if (fp_exact || !int_exact) {
      double acc[LT_TILE];
#pragma unroll
      for (int k = 0; k < (int)LT_TILE; k++) acc[k] = 0.0;
      lt_accumulate<double, double>(y, lo, steps, g, acc);
#pragma unroll
      for (int k = 0; k < (int)LT_TILE; k++)
        part_i[p][k0 + k] = fp_exact ? __double2ll_rz(acc[k]) : __double_as_longlong(acc[k]);
    } else {
      long long acc[LT_TILE];
#pragma unroll
      for (int k = 0; k < (int)LT_TILE; k++) acc[k] = 0;
      lt_accumulate<long long, int32_t>(y, lo, steps, g, acc);
#pragma unroll
      for (int k = 0; k < (int)LT_TILE; k++) part_i[p][k0 + k] = acc[k];
    }
  }
  __syncthreads();
  if (tid == 0) s_risk = 0;
  double* lagv = reinterpret_cast<double*>(y);                  /* the staged samples are no longer needed */
  for (uint32_t t = tid; t < SLAB_NUM_LTLAGS; t += blockDim.x) {
    double v;
    if (fp_exact || int_exact) {
      long long sum = 0;
      for (uint32_t p = 0; p < LT_PARTS; p++) sum += part_i[p][t];
      v = (double)sum;
    } else {
      v = 0.0;
      for (uint32_t p = 0; p < LT_PARTS; p++) v += __longlong_as_double(part_i[p][t]);
    }
    ac_out[(size_t)bc * 264u + t] = v * sh.ac_scale;
    lagv[t] = v;
  }
  __syncthreads();
  /* Which decisions of the pitch picker (SLAPredictor.c:867-924: sign tests, local-peak tests, peak
   * against peak) would the round-off of the reference's FFT autocorrelation be able to turn?  Its values
   * differ from these exact sums by about 1e-15 R(0); a comparison is at risk when the two exact values
   * are closer than 1e-11 R(0) - on real audio only exact ties and exact zeros are (quiet or periodic
   * synthetic blocks).  Such a block x channel is listed and re-analysed by k_enc_ltfft, which runs the
   * reference's transform itself. */
  if (risk_list != nullptr) {
    const double tol = fabs(lagv[0]) * 1e-11;
    if (tid == 0) s_npeaks = 0;
    __syncthreads();
    if (fabs(lagv[0]) > 0.0) {
      for (uint32_t t = tid; t < 258u; t += blockDim.x) {
        const double v = lagv[t];
        if (fabs(v) <= tol || fabs(v - lagv[t + 1u]) <= tol) s_risk = 1;
        /* a positive local peak: its rank among the other peaks decides the candidate and the maximum */
        if (t >= 1u && t < 257u && v > 0.0 && v > lagv[t - 1u] && v > lagv[t + 1u]) s_peaks[atomicAdd(&s_npeaks, 1u)] = (uint16_t)t;
      }
    }
    __syncthreads();
    const uint32_t np = s_npeaks;
    for (uint32_t i = tid; i + 1u < np; i += blockDim.x) {
      const double v = lagv[s_peaks[i]];
      for (uint32_t j = i + 1u; j < np; j++) if (fabs(v - lagv[s_peaks[j]]) <= tol) s_risk = 1;
    }
    __syncthreads();
    if (tid == 0 && s_risk) risk_list[atomicAdd(risk_count, 1u)] = bc;
  }
}

/* E6a': the reference's own autocorrelation for the listed block x channels - two real FFTs of the
 * handle's transform size around a power spectrum (SLAPredictor.c:791-853, SLAUtility.c:220-319: the
 * Numerical Recipes four1 / realft pair).  Every butterfly evaluates the reference's expression with the
 * reference's trigonometric factors - tables the host fills with the reference's recurrences and the host
 * libm - so the doubles are the reference's bit for bit, and with them every tie-break of the pitch picker.
 * A CTA takes one listed block x channel at a time; the working set (2 x 256 KB) lives in global memory (L2),
 * the first twelve stages of each transform run on 64 KB groups in shared memory.
 * Thread 0 then redoes pitch pick, tap solve and tap quantisation on the exact values. */
struct LtFftTables {
  const double* cf;      /* complex FFT, isign = +1: (wr, wi) per stage, stage with mmax at offset mmax/2 - 1 */
  const double* ci;      /* isign = -1 */
  const double* rf;      /* realft post-processing, forward: (wr, wi) for i = 2 .. n/4 */
  const double* ri;      /* inverse */
};

#define LTFFT_GROUP 4096u        /* complex points whose stages run in shared memory (64 KB) */

/* four1 on nn complex points: src -> dst (both global, distinct).  The bit-reversal permutation and the
 * stages with butterfly distance below LTFFT_GROUP work on contiguous groups of the permuted array, so a
 * group is gathered into shared memory, takes its 12 stages there and is written out once; only the last
 * stages (distance >= LTFFT_GROUP) run over global memory.  Every butterfly is the same expression with the
 * same factors as before - the doubles do not depend on where the operands live. */
__device__ __forceinline__ void ltfft_cfft(const double* src, double* dst, double* sm, uint32_t nn, const double* tw,
                                           uint32_t tid, uint32_t nthreads)
{
  const uint32_t lg = 31u - (uint32_t)__clz((int)nn);
  const uint32_t G = nn < LTFFT_GROUP ? nn : LTFFT_GROUP;
  for (uint32_t base = 0; base < nn; base += G) {
    for (uint32_t r = tid; r < G; r += nthreads) {
      const uint32_t j = __brev(base + r) >> (32u - lg);          /* the swap loop of four1, as a gather */
      sm[2u * r] = src[2u * j]; sm[2u * r + 1u] = src[2u * j + 1u];
    }
    __syncthreads();
    for (uint32_t half = 1u; half < G; half <<= 1) {              /* half = mmax / 2 in four1's terms */
      const double* w = tw + 2u * (size_t)(half - 1u);
      for (uint32_t b = tid; b < (G >> 1); b += nthreads) {
        const uint32_t mc = b & (half - 1u), g = b / half;
        const uint32_t ia = 2u * (g * 2u * half + mc), ik = ia + 2u * half;
        const double wr = w[2u * mc], wi = w[2u * mc + 1u];
        const double tr = wr * sm[ik] - wi * sm[ik + 1u];
        const double ti = wr * sm[ik + 1u] + wi * sm[ik];
        sm[ik] = sm[ia] - tr; sm[ik + 1u] = sm[ia + 1u] - ti;
        sm[ia] += tr; sm[ia + 1u] += ti;
      }
      __syncthreads();
    }
    for (uint32_t r = tid; r < 2u * G; r += nthreads) dst[2u * base + r] = sm[r];
    __syncthreads();
  }
  for (uint32_t half = G; half < nn; half <<= 1) {
    const double* w = tw + 2u * (size_t)(half - 1u);
    for (uint32_t b = tid; b < (nn >> 1); b += nthreads) {
      const uint32_t mc = b & (half - 1u), g = b / half;
      const uint32_t ia = 2u * (g * 2u * half + mc), ik = ia + 2u * half;
      const double wr = w[2u * mc], wi = w[2u * mc + 1u];
      const double tr = wr * dst[ik] - wi * dst[ik + 1u];
      const double ti = wr * dst[ik + 1u] + wi * dst[ik];
      dst[ik] = dst[ia] - tr; dst[ik + 1u] = dst[ia + 1u] - ti;
      dst[ia] += tr; dst[ia + 1u] += ti;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void ltfft_realft_post(double* d, uint32_t n, const double* tw, double c2, uint32_t tid, uint32_t nthreads)
{
  const double c1 = 0.5;
  for (uint32_t i = 2u + tid; i <= (n >> 2); i += nthreads) {
    /* 1-based indices of realft: i1 = 2i - 1, i2 = 2i, i3 = n + 3 - i2, i4 = i3 + 1 */
    const uint32_t i1 = 2u * i - 2u, i2 = i1 + 1u, i3 = n + 2u - 2u * i, i4 = i3 + 1u;
    const double wr = tw[2u * (i - 2u)], wi = tw[2u * (i - 2u) + 1u];
    const double h1r = c1 * (d[i1] + d[i3]), h1i = c1 * (d[i2] - d[i4]);
    const double h2r = -c2 * (d[i2] + d[i4]), h2i = c2 * (d[i1] - d[i3]);
    d[i1] = h1r + wr * h2r - wi * h2i;
    d[i2] = h1i + wr * h2i + wi * h2r;
    d[i3] = h1r - wr * h2r + wi * h2i;
    d[i4] = -h1i + wr * h2i + wi * h2r;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) k_enc_ltfft(EncShape sh, uint32_t fft_size,
    const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ blk_len,
    const int32_t* __restrict__ r1, const uint32_t* __restrict__ risk_list,
    const uint32_t* __restrict__ risk_count, double* __restrict__ scratch, LtFftTables tb,
    double* __restrict__ ac_out, EncChan* __restrict__ chan, double* __restrict__ lt_out, int32_t* __restrict__ ltq_out)
{
  SLAB_DYN_SMEM(double, fsm);
  __shared__ double s_ac[264];
  __shared__ uint32_t s_masks[27];
  __shared__ uint16_t s_cand[264];
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  const uint32_t count = *risk_count;
  double* d = scratch + (size_t)blockIdx.x * 2u * fft_size;       /* two buffers: the transforms go back and forth */
  double* e = d + fft_size;
  for (uint32_t item = blockIdx.x; item < count; item += gridDim.x) {
    const uint32_t bc = risk_list[item], b = bc / sh.nch, c = bc - b * sh.nch;
    const uint32_t n = blk_len[b];
    const int32_t* src = r1 + (size_t)c * sh.NP + blk_start[b];
    for (uint32_t i = tid; i < fft_size; i += nt) d[i] = (i < n) ? (double)src[i] * 4.656612873077392578125e-10 : 0.0;
    __syncthreads();
    /* forward: four1 on n/2 complex points (d -> e), then the real-transform pass */
    ltfft_cfft(d, e, fsm, fft_size >> 1, tb.cf, tid, nt);
    ltfft_realft_post(e, fft_size, tb.rf, -0.5, tid, nt);
    if (tid == 0) { const double h = e[0]; e[0] = h + e[1]; e[1] = h - e[1]; }
    __syncthreads();
    /* power spectrum, SLAPredictor.c:838-846 */
    for (uint32_t i = tid; i < (fft_size >> 1); i += nt) {
      if (i == 0) { e[0] *= e[0]; e[1] *= e[1]; }
      else { const double re = e[2u * i], im = e[2u * i + 1u]; e[2u * i] = re * re + im * im; e[2u * i + 1u] = 0.0; }
    }
    __syncthreads();
    /* inverse (un-normalised, as the reference leaves it): e -> d */
    ltfft_realft_post(e, fft_size, tb.ri, 0.5, tid, nt);
    if (tid == 0) { const double h = e[0]; e[0] = 0.5 * (h + e[1]); e[1] = 0.5 * (h - e[1]); }
    __syncthreads();
    ltfft_cfft(e, d, fsm, fft_size >> 1, tb.ci, tid, nt);
    for (uint32_t t = tid; t < 264u; t += nt) {
      const double v = (t < SLAB_NUM_LTLAGS) ? d[t] : 0.0;
      if (t < SLAB_NUM_LTLAGS) ac_out[(size_t)bc * 264u + t] = v;
      s_ac[t] = v;
    }
    __syncthreads();
    if (tid < 32u) {                                 /* warp 0: pitch pick, then lane 0: tap solve and quantisation */
      const uint32_t first_cand = enc_pitch_pick_warp(s_ac, s_masks, s_cand, tid);
      if (tid == 0) {
        uint32_t pitch = 0;
        double coef[SLAB_MAX_TAPS];
        const int rc = enc_taps_solve(s_ac, sh.T, first_cand, &pitch, coef);
        if (rc != 0 || pitch >= SLAB_MAX_PITCH) pitch = 0;
        for (uint32_t j = 0; j < sh.T; j++) {
          lt_out[(size_t)bc * 8 + j] = coef[j];
          ltq_out[(size_t)bc * 8 + j] = (int32_t)((uint32_t)enc_d2i_x86(enc_round(coef[j] * 32768.0)) << 16);
        }
        chan[bc].pitch = pitch;
      }
    }
    __syncthreads();
  }
}

/* How far do the taps move when the autocorrelation moves by 1e-12 relative (a thousand times the
 * round-off that separates the exact lag sums from the reference's FFT values)?  For tonal residuals the
 * normal equations of a 3- or 5-tap predictor are nearly singular and amplify that difference into the
 * Q15 codes; such a block x channel goes to k_enc_ltfft as well. */
__device__ inline bool enc_taps_sensitive(const double* ac, uint32_t taps, uint32_t pitch)
{
  if (taps < 2u || pitch < taps / 2u + 1u) return false;
  double sol[2][SLAB_MAX_TAPS], mag[2] = {0.0, 0.0};
  for (int pass = 0; pass < 2; pass++) {
    double Rm[SLAB_MAX_TAPS][SLAB_MAX_TAPS];
    const double eps = pass ? 1e-12 : 0.0;
    for (uint32_t j = 0; j < taps; j++)
      for (uint32_t k = 0; k < taps; k++) {
        const uint32_t lag = (j >= k) ? (j - k) : (k - j);
        Rm[j][k] = ac[lag] + eps * ac[0] * enc_unstructured(lag);
      }
    for (uint32_t j = 0; j < taps; j++) {
      const uint32_t lag = j + pitch - taps / 2u;
      sol[pass][j] = ac[lag] + eps * ac[0] * enc_unstructured(lag + 1000u);
    }
    if (enc_lu_solve(Rm, sol[pass], taps) != 0) return true;
    for (uint32_t j = 0; j < taps; j++) mag[pass] += fabs(sol[pass][j]);
  }
  /* the |c| >= 1 fallback to a single tap (SLAPredictor.c:958-970) is a threshold decision too */
  if ((mag[0] >= 1.0) != (mag[1] >= 1.0) || fabs(mag[0] - 1.0) < 3e-6) return true;
  if (mag[0] >= 1.0) return false;               /* the single centre tap is a plain ratio */
  for (uint32_t j = 0; j < taps; j++) if (fabs(sol[1][j] - sol[0][j]) > 3e-6) return true;
  return false;
}

/* E6b, one warp per block x channel: pitch pick (the warp), tap solve and tap quantisation (lane 0) */
__global__ void __launch_bounds__(128) k_enc_ltsolve(EncShape sh, uint32_t nblocks,
    const uint32_t* __restrict__ blk_type, const double* __restrict__ ac_in,
    EncChan* __restrict__ chan, double* __restrict__ lt_out, int32_t* __restrict__ ltq_out,
    uint32_t* __restrict__ risk_list, uint32_t* __restrict__ risk_count)
{
  __shared__ double s_ac[4][264];
  __shared__ uint32_t s_masks[4][27];
  __shared__ uint16_t s_cand[4][264];
  const uint32_t lane = threadIdx.x & 31u, wp = threadIdx.x >> 5;
  const uint32_t bc = blockIdx.x * 4u + wp;
  if (bc >= nblocks * sh.nch) return;                               /* whole warps leave together */
  if (blk_type[bc / sh.nch] != SLAB_BLOCK_COMPRESS) return;
  const double* ac = ac_in + (size_t)bc * 264u;
  for (uint32_t t = lane; t < 264u; t += 32u) s_ac[wp][t] = (t < SLAB_NUM_LTLAGS) ? ac[t] : 0.0;
  __syncwarp();
  const uint32_t first_cand = enc_pitch_pick_warp(s_ac[wp], s_masks[wp], s_cand[wp], lane);
  if (lane != 0) return;
  const double* acs = s_ac[wp];
  const uint32_t taps = sh.T;
  uint32_t pitch = 0;
  double coef[SLAB_MAX_TAPS];
  const int rc = enc_taps_solve(acs, taps, first_cand, &pitch, coef);
  if (risk_list != nullptr && taps > 1u && rc == 0 && pitch != 0 && pitch < SLAB_MAX_PITCH) {
    if (enc_taps_sensitive(acs, taps, pitch)) risk_list[atomicAdd(risk_count, 1u)] = bc;
  }
  if (rc != 0 || pitch >= SLAB_MAX_PITCH) pitch = 0;                 /* SLAEncoder.c:629-632 */
  for (uint32_t j = 0; j < taps; j++) {                              /* SLAEncoder.c:635-640 */
    lt_out[(size_t)bc * 8 + j] = coef[j];
    ltq_out[(size_t)bc * 8 + j] = (int32_t)((uint32_t)enc_d2i_x86(enc_round(coef[j] * 32768.0)) << 16);
  }
  chan[bc].pitch = pitch;
}

/* ------------------------------------------------------------------------------------ E7 + E8 */
/* Long-term FIR (SLAPredictor.c:1031-1108, is_predict) fused with the sign-LMS predictor
 * (SLAPredictor.c:1202-1331); one thread per block x channel, filter state in registers.
 * Samples are processed in chunks of LMS_N: all inputs of a chunk (and the long-term taps' history,
 * which is plain input here) are loaded up front so that their latency overlaps, and the delay lines
 * are ring buffers whose slot index is a compile-time constant after unrolling (no shifting). */
template <int LMS_N, int TAPS>      /* TAPS = long-term taps rounded up to 1, 3 or 7 */
__global__ void __launch_bounds__(64) k_enc_ltlms(EncShape sh, uint32_t nblocks,
    const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ blk_len,
    const uint32_t* __restrict__ blk_type, const int32_t* __restrict__ ltq_in,
    const int32_t* __restrict__ r1, int32_t* __restrict__ r3, EncChan* __restrict__ chan)
{
  const uint32_t bc = blockIdx.x * blockDim.x + threadIdx.x;
  if (bc >= nblocks * sh.nch) return;
  const uint32_t b = bc / sh.nch, c = bc - b * sh.nch;
  if (blk_type[b] != SLAB_BLOCK_COMPRESS) return;
  const uint32_t n = blk_len[b];
  const int32_t* x = r1 + (size_t)c * sh.NP + blk_start[b];          /* blk_start = padded starts here */
  int32_t* out = r3 + (size_t)c * sh.NP + blk_start[b];
  const uint32_t pitch = chan[bc].pitch, T = sh.T;
  const bool use_lt = pitch >= 3u;                                   /* SLAInternal.h:14 */
  const uint32_t delay = pitch + (T >> 1);
  int32_t ltc[TAPS];
#pragma unroll
  for (int j = 0; j < TAPS; j++) ltc[j] = (use_lt && (uint32_t)j < T) ? ltq_in[(size_t)bc * 8 + j] : 0;
  /* ring buffers: slot (t mod LMS_N) holds the value of time t */
  int32_t cx[LMS_N], cp[LMS_N], hx[LMS_N], hp[LMS_N], sx[LMS_N], sp[LMS_N];
#pragma unroll
  for (int i = 0; i < LMS_N; i++) { cx[i] = cp[i] = 0; hx[i] = hp[i] = sx[i] = sp[i] = 0; }
  unsigned long long zsum = 0;
  const bool filter = n > (uint32_t)LMS_N;
  /* 128-bit loads/stores: blocks start on multiples of 8 samples in the intermediate planes and are
   * padded to a multiple of 8, so a chunk never leaves the block's own storage */
  const int4* xv = reinterpret_cast<const int4*>(x);
  int4* ov = reinterpret_cast<int4*>(out);
  /* the chunk after the one being filtered is already on its way: both the input samples and the
   * pitch-delayed history come from the input plane, so their loads never depend on the filter state
   * and a whole chunk of arithmetic hides their latency */
  int4 nx[LMS_N / 4];
  int32_t nh[LMS_N + TAPS - 1];
  auto fetch = [&](uint32_t at) {
#pragma unroll
    for (int q = 0; q < LMS_N / 4; q++) nx[q] = xv[(at >> 2) + q];
    if (use_lt) {
#pragma unroll
      for (int u = 0; u < LMS_N + TAPS - 1; u++) {
        const uint32_t idx = at + (uint32_t)u;                      /* position at + u - delay */
        nh[u] = ((uint32_t)u < (uint32_t)LMS_N + T - 1u && idx >= delay && idx - delay < n) ? x[idx - delay] : 0;
      }
    }
  };
#pragma unroll
  for (int u = 0; u < LMS_N + TAPS - 1; u++) nh[u] = 0;
  if (n > 0) fetch(0);
  for (uint32_t s0 = 0; s0 < n; s0 += LMS_N) {
    int32_t xin[LMS_N], hist[LMS_N + TAPS - 1], res[LMS_N];
#pragma unroll
    for (int q = 0; q < LMS_N / 4; q++) {
      xin[4 * q] = nx[q].x; xin[4 * q + 1] = nx[q].y; xin[4 * q + 2] = nx[q].z; xin[4 * q + 3] = nx[q].w;
    }
#pragma unroll
    for (int u = 0; u < LMS_N + TAPS - 1; u++) hist[u] = nh[u];
    if (s0 + LMS_N < n) fetch(s0 + LMS_N);
#pragma unroll
    for (int u = 0; u < LMS_N; u++) {
      const uint32_t s = s0 + (uint32_t)u;
      int32_t v = xin[u];
      if (use_lt && s >= delay) {
        long long acc = 1ll << 30;
#pragma unroll
        for (int j = 0; j < TAPS; j++) acc += (long long)ltc[j] * (long long)hist[u + j];
        v = (int32_t)((uint32_t)v - (uint32_t)(int32_t)(acc >> 31));
      }
      int32_t resid = v;
      if (filter) {
        if (s0 == 0) {                            /* first LMS_N samples prime both delay lines */
          hx[u] = hp[u] = v; sx[u] = sp[u] = slab_sgn(v);
        } else {
          uint32_t acc0 = 1u << 9, acc1 = 0, acc2 = 0, acc3 = 0;
#pragma unroll
          for (int i = 0; i < LMS_N; i += 2) {
            acc0 += (uint32_t)cx[i] * (uint32_t)hx[(u - 1 - i + 2 * LMS_N) % LMS_N];
            acc1 += (uint32_t)cp[i] * (uint32_t)hp[(u - 1 - i + 2 * LMS_N) % LMS_N];
            acc2 += (uint32_t)cx[i + 1] * (uint32_t)hx[(u - 2 - i + 2 * LMS_N) % LMS_N];
            acc3 += (uint32_t)cp[i + 1] * (uint32_t)hp[(u - 2 - i + 2 * LMS_N) % LMS_N];
          }
          const int32_t pred = (int32_t)((acc0 + acc1) + (acc2 + acc3)) >> 10;
          resid = (int32_t)((uint32_t)v - (uint32_t)pred);
          const uint32_t mag = (resid < 0) ? (0u - (uint32_t)resid) : (uint32_t)resid;
          const int32_t step = slab_sgn(resid) * (int32_t)(slab_bitlen(mag) >> 1);
#pragma unroll
          for (int i = 0; i < LMS_N; i++) {
            cx[i] += step * sx[(u - 1 - i + 2 * LMS_N) % LMS_N];
            cp[i] += step * sp[(u - 1 - i + 2 * LMS_N) % LMS_N];
          }
          hx[u] = v; hp[u] = pred; sx[u] = slab_sgn(v); sp[u] = slab_sgn(pred);
        }
      }
      res[u] = resid;
      if (s < n) zsum += slab_zigzag(resid);
    }
#pragma unroll
    for (int q = 0; q < LMS_N / 4; q++) {
      int4 t;
      t.x = res[4 * q]; t.y = res[4 * q + 1]; t.z = res[4 * q + 2]; t.w = res[4 * q + 3];
      ov[(s0 >> 2) + q] = t;
    }
  }
  chan[bc].zsum = zsum;
}

/* Generic fallback for LMS orders 16 and 32: runtime loops, state in local memory. */
__global__ void __launch_bounds__(64) k_enc_ltlms_generic(EncShape sh, uint32_t nblocks,
    const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ blk_len,
    const uint32_t* __restrict__ blk_type, const int32_t* __restrict__ ltq_in,
    const int32_t* __restrict__ r1, int32_t* __restrict__ r3, EncChan* __restrict__ chan)
{
  const uint32_t bc = blockIdx.x * blockDim.x + threadIdx.x;
  if (bc >= nblocks * sh.nch) return;
  const uint32_t b = bc / sh.nch, c = bc - b * sh.nch;
  if (blk_type[b] != SLAB_BLOCK_COMPRESS) return;
  const uint32_t n = blk_len[b], N = sh.lms, T = sh.T;
  const int32_t* x = r1 + (size_t)c * sh.NP + blk_start[b];          /* blk_start = padded starts here */
  int32_t* out = r3 + (size_t)c * sh.NP + blk_start[b];
  const uint32_t pitch = chan[bc].pitch;
  const bool use_lt = pitch >= 3u;
  const uint32_t delay = pitch + (T >> 1);
  int32_t ltc[SLAB_MAX_TAPS];
  for (uint32_t j = 0; j < SLAB_MAX_TAPS; j++) ltc[j] = (use_lt && j < T) ? ltq_in[(size_t)bc * 8 + j] : 0;
  int32_t cx[SLAB_MAX_LMS], cp[SLAB_MAX_LMS], hx[SLAB_MAX_LMS], hp[SLAB_MAX_LMS];
  for (uint32_t i = 0; i < N; i++) { cx[i] = cp[i] = 0; hx[i] = hp[i] = 0; }
  unsigned long long zsum = 0;
  for (uint32_t s = 0; s < n; s++) {
    int32_t v = x[s];
    if (use_lt && s >= delay) {
      long long acc = 1ll << 30;
      for (uint32_t j = 0; j < T; j++) acc = slab_mad_wide(ltc[j], x[s - delay + j], acc);
      v = (int32_t)((uint32_t)v - (uint32_t)(int32_t)(acc >> 31));
    }
    int32_t resid = v;
    if (n > N) {
      const uint32_t slot = s & (N - 1u);
      if (s < N) { hx[slot] = hp[slot] = v; }
      else {
        uint32_t acc = 1u << 9;
        for (uint32_t i = 0; i < N; i++) {
          const uint32_t q = (s - 1u - i) & (N - 1u);
          acc += (uint32_t)cx[i] * (uint32_t)hx[q] + (uint32_t)cp[i] * (uint32_t)hp[q];
        }
        const int32_t pred = (int32_t)acc >> 10;
        resid = (int32_t)((uint32_t)v - (uint32_t)pred);
        const uint32_t mag = (resid < 0) ? (0u - (uint32_t)resid) : (uint32_t)resid;
        const int32_t step = slab_sgn(resid) * (int32_t)(slab_bitlen(mag) >> 1);
        for (uint32_t i = 0; i < N; i++) {
          const uint32_t q = (s - 1u - i) & (N - 1u);
          cx[i] += step * slab_sgn(hx[q]); cp[i] += step * slab_sgn(hp[q]);
        }
        hx[slot] = v; hp[slot] = pred;
      }
    }
    out[s] = resid;
    zsum += slab_zigzag(resid);
  }
  chan[bc].zsum = zsum;
}

/* ------------------------------------------------------------------------------------ E9 */
/* code length helpers ------------------------------------------------------------ */
__device__ __forceinline__ uint32_t enc_gamma_len(uint32_t g)          /* SLACoder.c:120-138 */
{
  return (g == 0) ? 1u : 2u * slab_log2ceil(g + 2u) - 1u;
}
/* adaptive code for v given (k0, k1): SLACoder.c:224-270 */
__device__ __forceinline__ uint64_t enc_rice_len(uint32_t v, uint32_t k0, uint32_t k1)
{
  if (v < (1u << k0)) return 1u + k0;
  const uint32_t q = 1u + ((v - (1u << k0)) >> k1);
  return (uint64_t)((q < 16u) ? q + 1u : 17u + enc_gamma_len(q - 16u)) + k1;
}
/* fixed Golomb code, SLACoder.c:45-82 */
__device__ __forceinline__ uint64_t enc_golomb_len(uint32_t v, uint32_t m)
{
  const uint32_t q = v / m, rest = v - q * m;
  if ((m & (m - 1u)) == 0) return (uint64_t)q + 1u + slab_log2ceil(m);
  const uint32_t bb = slab_log2ceil(m), cut = (1u << bb) - m;
  return (uint64_t)q + 1u + ((rest < cut) ? bb - 1u : bb);
}

/* per block: initial Rice parameter (SLACoder.c:361-385), coder mode (:442-447), header size */
__global__ void __launch_bounds__(128) k_enc_riceprep(EncShape sh, uint32_t nblocks,
    const uint32_t* __restrict__ blk_len, const uint32_t* __restrict__ blk_type,
    EncChan* __restrict__ chan, uint32_t* __restrict__ blk_mode, uint32_t* __restrict__ blk_hdr_bytes)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  uint32_t bits = 16u + 32u + 16u + 16u + 2u;
  uint32_t mode = 0;
  if (blk_type[b] == SLAB_BLOCK_COMPRESS) {
    const uint32_t n = blk_len[b];
    unsigned long long avg = 0;
    for (uint32_t c = 0; c < sh.nch; c++) {
      EncChan& ch = chan[b * sh.nch + c];
      unsigned long long mean = ch.zsum / n;
      const uint32_t init = (uint32_t)(mean > 1ull ? mean : 1ull);
      ch.rice_init = (uint32_t)(init << 8);
      avg += slab_rice_param(ch.rice_init);
      const uint32_t hi = sh.P < 3u ? sh.P : 3u;
      bits += 4u + hi * 16u + (sh.P - hi) * 8u + 1u + ((ch.pitch >= 3u) ? 10u + 16u * sh.T : 0u) + sh.bits;
    }
    avg /= sh.nch;
    mode = (avg > 8ull) ? 1u : 0u;
  }
  blk_mode[b] = mode;
  blk_hdr_bytes[b] = (bits + 7u) >> 3;
}

/* per block x channel: the sequential parameter trace.  Stores, per code, the two Rice exponents
 * (k0 | k1 << 5); code lengths and bits are re-derived from them in the packing kernel.
 * One warp per 32 block x channel rows; the residual rows arrive through the cp.async tile pipeline
 * of slab_lanestream.cuh and the exponent rows leave through a shared-memory tile, so no lane ever
 * waits on its own scattered global access inside the recurrence. */
#define RT_TILE   32u                      /* samples per row tile */
#define RT_STAGES 4
struct RiceTraceSmem {
  SlabLsRows rin, rout;
  unsigned char in[RT_STAGES][SlabLsGeom<RT_TILE * 4>::STAGE];
  unsigned char out[SlabLsGeom<RT_TILE * 2>::STAGE];
};

__global__ void __launch_bounds__(32) k_enc_ricetrace(EncShape sh, uint32_t nblocks,
    const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ blk_len,
    const uint32_t* __restrict__ blk_type, const uint32_t* __restrict__ blk_mode,
    const int32_t* __restrict__ r3, EncChan* __restrict__ chan, uint16_t* __restrict__ meta)
{
  typedef SlabLsGeom<RT_TILE * 4> GI;
  typedef SlabLsGeom<RT_TILE * 2> GO;
  SLAB_DYN_SMEM(RiceTraceSmem, sm);
  const uint32_t lane = threadIdx.x;
  const uint32_t bc = blockIdx.x * 32u + lane;
  const bool valid = bc < nblocks * sh.nch;
  const uint32_t b = valid ? bc / sh.nch : 0u, c = bc - b * sh.nch;
  const bool active = valid && blk_type[b] == SLAB_BLOCK_COMPRESS;
  const uint32_t n = active ? blk_len[b] : 0u;
  const uint32_t mode = active ? blk_mode[b] : 0u;
  const uint32_t npad = (n + 7u) & ~7u;
  const size_t slot = (size_t)c * sh.NP + (valid ? blk_start[b] : 0u);       /* blk_start = padded starts here */
  sm->rin.ptr[lane] = (unsigned long long)(r3 + slot);
  sm->rin.bytes[lane] = npad * 4u;
  sm->rout.ptr[lane] = (unsigned long long)(meta + slot);
  sm->rout.bytes[lane] = mode ? npad * 2u : 0u;
  /* the running means fit 32 bits (slab_rice_update32) */
  uint32_t p0 = active ? (uint32_t)chan[bc].rice_init : 0u, p1 = p0;
  const uint32_t gm = slab_rice_param(p0);                                  /* fixed-Golomb parameter */
  __syncwarp();
  const uint32_t ntiles = slab_warp_max((n + RT_TILE - 1u) / RT_TILE);
#pragma unroll
  for (int t = 0; t < RT_STAGES - 1; t++) {
    if ((uint32_t)t < ntiles) slab_ls_load<RT_TILE * 4>(&sm->rin, sm->in[t], (uint32_t)t, lane);
    slab_cp_async_commit();
  }
  unsigned long long bits = 0;
  for (uint32_t t = 0; t < ntiles; t++) {
    const uint32_t tn = t + RT_STAGES - 1u;
    if (tn < ntiles) slab_ls_load<RT_TILE * 4>(&sm->rin, sm->in[tn % RT_STAGES], tn, lane);
    slab_cp_async_commit();
    slab_cp_async_wait<RT_STAGES - 1>();
    __syncwarp();
    const int4* xv = reinterpret_cast<const int4*>(sm->in[t % RT_STAGES] + lane * GI::ROW);
    uint4* mv = reinterpret_cast<uint4*>(sm->out + lane * GO::ROW);
    const uint32_t base = t * RT_TILE;
    if (base < n) {
      if (mode) {
#pragma unroll
        for (uint32_t s0 = 0; s0 < RT_TILE; s0 += 8u) {
          if (base + s0 < n) {
            const int4 a = xv[s0 >> 2], bq = xv[(s0 >> 2) + 1u];
            const int32_t xin[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
            uint32_t mt[8];
            uint32_t group_bits = 0;                      /* 8 codes of at most ~100 bits each */
            const uint32_t live = n - (base + s0);        /* samples of this group inside the block (>= 1) */
#pragma unroll
            for (int u = 0; u < 8; u++) {
              const uint32_t v = slab_zigzag(xin[u]);
              const uint32_t k0 = slab_rice_k32(p0);
              const uint32_t k1r = slab_rice_k32(p1);
              const bool second = v >= (1u << k0);
              const uint32_t rest = v - (1u << k0);
              const uint32_t p1n = slab_rice_update32(p1, rest);
              p0 = slab_rice_update32(p0, v);
              p1 = second ? p1n : p1;
              const uint32_t k1 = second ? k1r : 0u;
              mt[u] = k0 | (k1 << 5);
              /* code length (enc_rice_len) without a branch on `second`; only the gamma escape branches */
              const uint32_t q = 1u + (rest >> k1);
              uint32_t len = second ? q + 1u + k1 : 1u + k0;
              if (second && q >= 16u) len = 17u + enc_gamma_len(q - 16u) + k1;
              group_bits += ((uint32_t)u < live) ? len : 0u;
            }
            bits += group_bits;
            uint4 packed;
            packed.x = mt[0] | (mt[1] << 16); packed.y = mt[2] | (mt[3] << 16);
            packed.z = mt[4] | (mt[5] << 16); packed.w = mt[6] | (mt[7] << 16);
            mv[s0 >> 3] = packed;
          }
        }
      } else {
        const int32_t* xs = reinterpret_cast<const int32_t*>(xv);
        const uint32_t cnt = (n - base < RT_TILE) ? n - base : RT_TILE;
        for (uint32_t u = 0; u < cnt; u++) bits += enc_golomb_len(slab_zigzag(xs[u]), gm);
      }
    }
    __syncwarp();
    slab_ls_store<RT_TILE * 2>(&sm->rout, sm->out, t, lane);
    __syncwarp();
  }
  if (valid) chan[bc].bits = active ? bits : 0ull;
}

/* per block: byte size; file statistics (SLAEncoder.c:887-898, uint32 wrap included) */
__global__ void __launch_bounds__(128) k_enc_blocksizes(EncShape sh, uint32_t nblocks,
    const uint32_t* __restrict__ blk_len, const uint32_t* __restrict__ blk_type,
    const uint32_t* __restrict__ blk_hdr_bytes, const EncChan* __restrict__ chan,
    uint32_t* __restrict__ blk_size, uint32_t* __restrict__ misc)
{
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const uint32_t n = blk_len[b], type = blk_type[b];
  unsigned long long bits = 0;
  if (type == SLAB_BLOCK_COMPRESS) {
    for (uint32_t c = 0; c < sh.nch; c++) bits += chan[b * sh.nch + c].bits;
  } else if (type == SLAB_BLOCK_RAW) {
    uint32_t row = 0;
    for (uint32_t c = 0; c < sh.nch; c++) row += sh.bits - enc_lshift(sh, b) + ((c == 1u && sh.ms) ? 1u : 0u);
    bits = (unsigned long long)row * n;
  }
  unsigned long long size = blk_hdr_bytes[b] + ((bits + 7ull) >> 3);
  if (size > 0xF0000000ull) size = 0xF0000000ull;
  blk_size[b] = (uint32_t)size;
  atomicMax(&misc[M_MAX_BLOCK], (uint32_t)size);
  atomicMax(&misc[M_MAX_BPS], (uint32_t)((8u * (uint32_t)size * sh.rate) / n));
}

__global__ void k_enc_check_capacity(EncShape sh, uint32_t* __restrict__ misc)
{
  if (blockIdx.x == 0 && threadIdx.x == 0) misc[M_OVERFLOW] = (misc[M_TOTAL_BYTES] > sh.out_cap) ? 1u : 0u;
}

/* ---- bit packing --------------------------------------------------------------- */
#define PACK_TILE       256u
#define PACK_STAGE_WORDS 8192u            /* 32 KB of staged bits per tile */

/* OR `nbits` (<= 32) low bits of v into an MSB-first word array at bit position pos */
__device__ __forceinline__ void pack_put(uint32_t* stage, uint64_t pos, uint32_t v, uint32_t nbits)
{
  if (nbits == 0) return;
  const uint32_t w = (uint32_t)(pos >> 5), sh = (uint32_t)(pos & 31u);
  const uint64_t wide = ((uint64_t)v << (64u - nbits)) >> sh;        /* nbits <= 32, sh <= 31 */
  const uint32_t hi = (uint32_t)(wide >> 32), lo = (uint32_t)wide;
  if (hi) atomicOr(&stage[w], hi);
  if (lo) atomicOr(&stage[w + 1u], lo);
}

/* A code is a sequence of zero runs and explicit bit fields; the same emitters drive the staged
 * (shared-memory) sink and the byte-serial sink of the oversize fallback. */
struct StageSink {
  uint32_t* stage; uint64_t pos;
  __device__ __forceinline__ void zeros(uint64_t n) { pos += n; }
  __device__ __forceinline__ void bits(uint32_t v, uint32_t n) { pack_put(stage, pos, v, n); pos += n; }
};
struct ByteSink {
  uint8_t* dst; uint64_t pos; uint32_t cur, nb;
  __device__ __forceinline__ void push(uint32_t bit)
  {
    cur |= bit << (7u - nb);
    if (++nb == 8u) { dst[pos++] = (uint8_t)cur; cur = 0; nb = 0; }
  }
  __device__ __forceinline__ void zeros(uint64_t n) { for (uint64_t i = 0; i < n; i++) push(0); }
  __device__ __forceinline__ void bits(uint32_t v, uint32_t n) { for (uint32_t k = n; k-- > 0; ) push((v >> k) & 1u); }
};

/* SLACoder.c:224-270 */
template <class Sink> __device__ __forceinline__ void emit_rice(Sink& s, uint32_t v, uint32_t k0, uint32_t k1)
{
  if (v < (1u << k0)) { s.bits((1u << k0) | v, 1u + k0); return; }
  const uint32_t rest = v - (1u << k0);
  const uint32_t q = 1u + (rest >> k1);
  if (q < 16u) { s.zeros(q); s.bits(1u, 1u); }
  else {
    s.zeros(16u); s.bits(1u, 1u);
    const uint32_t g = q - 16u;                                      /* gamma, SLACoder.c:120-138 */
    if (g == 0) s.bits(1u, 1u);
    else { const uint32_t nd = slab_log2ceil(g + 2u); s.zeros(nd - 1u); s.bits(g + 1u, nd); }
  }
  s.bits(rest & ((1u << k1) - 1u), k1);
}
/* SLACoder.c:45-82 */
template <class Sink> __device__ __forceinline__ void emit_golomb(Sink& s, uint32_t v, uint32_t m)
{
  const uint32_t q = v / m, rest = v - q * m;
  s.zeros(q); s.bits(1u, 1u);
  if ((m & (m - 1u)) == 0) { s.bits(rest, slab_log2ceil(m)); return; }
  const uint32_t bb = slab_log2ceil(m), cut = (1u << bb) - m;
  if (rest < cut) s.bits(rest, bb - 1u); else s.bits(rest + cut, bb);
}
template <int CH, class Sink> __device__ __forceinline__ void emit_row(Sink& s, uint32_t type, uint32_t mode, uint32_t nch,
                                                               const uint32_t* vals, const uint32_t* mets)
{
#pragma unroll
  for (uint32_t c = 0; c < (uint32_t)CH; c++) {     /* unrolled with a guard: the arrays stay in registers */
    if (c < nch) {
      if (type == SLAB_BLOCK_RAW) s.bits((mets[c] >= 32u) ? vals[c] : (vals[c] & ((1u << mets[c]) - 1u)), mets[c]);
      else if (mode) emit_rice(s, vals[c], mets[c] & 31u, mets[c] >> 5);
      else emit_golomb(s, vals[c], mets[c]);
    }
  }
}

/* One CTA per block: header bits, then the sample-interleaved codes in tiles of 256 samples; each
 * tile is assembled in shared memory (word atomics) at prefix-scanned bit offsets and flushed as whole
 * bytes; every output byte is written exactly once, by the CTA that owns the block. */
template <int CH>      /* channel count the row loops are unrolled to: 1, 2, or SLAB_MAX_CH for anything else */
__global__ void __launch_bounds__(256, (CH <= 2 ? 4 : 2)) k_enc_pack(InPtrs in, EncShape sh,
    const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ blk_pst,
    const uint32_t* __restrict__ blk_len,
    const uint32_t* __restrict__ blk_type, const uint32_t* __restrict__ blk_mode,
    const uint32_t* __restrict__ blk_hdr_bytes, const uint32_t* __restrict__ blk_size,
    const uint32_t* __restrict__ blk_off, const EncChan* __restrict__ chan,
    const int32_t* __restrict__ code_in, const int32_t* __restrict__ ltq_in,
    const int32_t* __restrict__ r3, const uint16_t* __restrict__ meta,
    const uint32_t* __restrict__ misc, uint8_t* __restrict__ out, const uint32_t* __restrict__ defer)
{
  __shared__ uint32_t stage[PACK_STAGE_WORDS + 4];
  __shared__ uint32_t warp_tot[8];
  __shared__ int too_big;
  if (misc[M_OVERFLOW]) return;
  const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  if (defer != nullptr && defer[b] == 0u) return;      /* k_enc_pack_rice has written this block */
  const uint32_t n = blk_len[b], type = blk_type[b], mode = blk_mode[b], hdrb = blk_hdr_bytes[b];
  uint8_t* dst = out + blk_off[b];
  const size_t s0 = blk_start[b], p0 = blk_pst[b];

  /* ---- header (thread 0 builds it in the stage, everyone copies it out) ---- */
  for (uint32_t i = tid; i < PACK_STAGE_WORDS + 4u; i += 256u) stage[i] = 0;
  __syncthreads();
  if (tid == 0) {
    uint64_t p = 0;
    pack_put(stage, p, 0xFFFFu, 16); p += 16;
    p += 32 + 16;                                  /* size and CRC are patched by k_enc_crc */
    pack_put(stage, p, n, 16); p += 16;
    pack_put(stage, p, type, 2); p += 2;
    if (type == SLAB_BLOCK_COMPRESS) {
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
  }
  __syncthreads();
  for (uint32_t i = tid; i < hdrb; i += 256u) dst[i] = (uint8_t)(stage[i >> 2] >> (24u - 8u * (i & 3u)));
  __syncthreads();
  if (type == SLAB_BLOCK_SILENT) return;

  /* ---- data ---- */
  uint64_t byte_cursor = hdrb;        /* bytes already written */
  uint32_t carry_bits = 0;            /* bits pending in stage[0]'s top byte */
  const uint32_t shift = 32u - sh.bits + enc_lshift(sh, b);
  const uint32_t nch = sh.nch;
  for (uint32_t i = tid; i < PACK_STAGE_WORDS + 4u; i += 256u) stage[i] = 0;
  __syncthreads();
  /* the raw inputs of a row (residual + Rice exponents, or PCM for a raw block) are requested one tile
   * ahead of their use, so their latency overlaps the barriers and the packing of the current tile */
  int32_t nxt_v[CH]; uint32_t nxt_m[CH];
#define PACK_LOAD_ROW(S)                                                                             \
  do {                                                                                               \
    _Pragma("unroll") for (uint32_t c = 0; c < (uint32_t)CH; c++) {                                   \
      nxt_v[c] = 0; nxt_m[c] = 0;                                                                    \
      if (c < nch && (S) < n) {                                                                      \
        if (type == SLAB_BLOCK_RAW) nxt_v[c] = enc_sample(in, c, sh.ms, shift, s0 + (S));            \
        else {                                                                                       \
          nxt_v[c] = r3[(size_t)c * sh.NP + p0 + (S)];                                               \
          if (mode) nxt_m[c] = meta[(size_t)c * sh.NP + p0 + (S)];                                   \
        }                                                                                            \
      }                                                                                              \
    }                                                                                                \
  } while (0)
  uint32_t golomb_m[CH];
#pragma unroll
  for (uint32_t c = 0; c < (uint32_t)CH; c++)
    golomb_m[c] = (c < nch && type != SLAB_BLOCK_RAW && !mode) ? slab_rice_param(chan[b * nch + c].rice_init) : 1u;
  PACK_LOAD_ROW(tid);
  for (uint32_t t0 = 0; t0 < n; t0 += PACK_TILE) {
    const uint32_t s = t0 + tid;
    const bool live = s < n;
    /* pass 1: my row length */
    uint64_t row = 0;
    uint32_t vals[CH], mets[CH], fq[CH], fk[CH], fv[CH];
    bool row_fast = (type == SLAB_BLOCK_COMPRESS) && mode;
#pragma unroll
    for (uint32_t c = 0; c < (uint32_t)CH; c++) {
      vals[c] = slab_zigzag(nxt_v[c]); mets[c] = nxt_m[c]; fq[c] = fk[c] = fv[c] = 0;
      if (live && c < nch) {
        if (type == SLAB_BLOCK_RAW) {
          mets[c] = 32u - shift + ((c == 1u && sh.ms) ? 1u : 0u);     /* bits - offset_lshift */
          row += mets[c];
        } else if (mode) {
          /* branch-free form of the common code (quotient below 16): `q` zeros, then 1 + k bits `V` */
          const uint32_t k0 = mets[c] & 31u, k1 = mets[c] >> 5;
          const bool second = vals[c] >= (1u << k0);
          const uint32_t rest = vals[c] - (1u << k0);
          const uint32_t q = second ? 1u + (rest >> k1) : 0u;
          const uint32_t k = second ? k1 : k0;
          fq[c] = q; fk[c] = k;
          fv[c] = (1u << k) | ((second ? rest : vals[c]) & ((1u << k) - 1u));
          if (q < 16u) row += q + 1u + k;
          else { row_fast = false; row += enc_rice_len(vals[c], k0, k1); }
        } else {
          mets[c] = golomb_m[c];
          row += enc_golomb_len(vals[c], mets[c]);
        }
      }
    }
    PACK_LOAD_ROW(s + PACK_TILE);
    /* a tile that cannot fit the stage (only with huge unary escapes) takes the serial path */
    const uint32_t row32 = (row > 0x007FFFFFull) ? 0x007FFFFFu : (uint32_t)row;
    uint32_t x = row32;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t yv = __shfl_up_sync(SLAB_FULL_MASK, x, d); if (lane >= (uint32_t)d) x += yv; }
    if (lane == 31u) warp_tot[wid] = x;
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (uint32_t w = 0; w < 8u; w++) { if (w < wid) before += warp_tot[w]; total += warp_tot[w]; }
    const uint64_t my_off = (uint64_t)carry_bits + before + (x - row32);
    if (tid == 0) too_big = ((uint64_t)carry_bits + total > (uint64_t)PACK_STAGE_WORDS * 32u) || (total >= 0x007FFFFFu);
    __syncthreads();
    if (!too_big) {
      if (live) {
        if (row_fast) {
          uint32_t pos = (uint32_t)my_off;
#pragma unroll
          for (uint32_t c = 0; c < (uint32_t)CH; c++) {
            if (c < nch) { pos += fq[c]; pack_put(stage, pos, fv[c], 1u + fk[c]); pos += 1u + fk[c]; }
          }
        } else {
          StageSink sink; sink.stage = stage; sink.pos = my_off;
          emit_row<CH>(sink, type, mode, nch, vals, mets);
        }
      }
      __syncthreads();
      const uint32_t nbits = carry_bits + total;
      const uint32_t full = nbits >> 3;
      {
        /* whole bytes of the stage go out as aligned 32-bit words (the stage holds big-endian words: a
         * funnel shift picks four stream bytes at any byte offset, a byte permute puts them in memory
         * order); at most three single bytes on either side */
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
    } else {
      /* oversize tile (giant unary escapes in fixed-Golomb mode): the rows go out one thread after
       * the other through a byte-serial writer; stage[0..2] carry the writer state between threads */
      for (uint32_t who = 0; who < PACK_TILE && t0 + who < n; who++) {
        if (tid == who) {
          ByteSink sink; sink.dst = dst; sink.pos = byte_cursor; sink.cur = stage[0] >> 24; sink.nb = carry_bits;
          emit_row<CH>(sink, type, mode, nch, vals, mets);
          stage[0] = sink.cur << 24;
          stage[1] = (uint32_t)(sink.pos - byte_cursor);
          stage[2] = sink.nb;
        }
        __syncthreads();
        byte_cursor += stage[1];
        carry_bits = stage[2];
        __syncthreads();
      }
      if (tid == 0) { stage[1] = 0; stage[2] = 0; }
      __syncthreads();
    }
  }
  if (carry_bits && tid == 0) dst[byte_cursor] = (uint8_t)(stage[0] >> 24);
#undef PACK_LOAD_ROW
}


/* ------------------------------------------------------------------------------------ E10 */
/* One warp per block: CRC-16/IBM over bytes [8, size) in 32 slices combined through x^(8n) mod P,
 * then the size-6 and CRC fields are patched in (SLAEncoder.c:787-795). */
__global__ void __launch_bounds__(128) k_enc_crc(uint32_t nblocks, const uint32_t* __restrict__ blk_size,
    const uint32_t* __restrict__ blk_off, const uint32_t* __restrict__ misc, uint8_t* __restrict__ out)
{
  __shared__ SlabCrcTables tb;
  slab_crc16_build_tables(&tb);
  if (misc[M_OVERFLOW]) return;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (warp >= nblocks) return;
  uint8_t* b = out + blk_off[warp];
  const uint32_t size = blk_size[warp];
  const uint32_t total = size - 8u;
  const uint32_t slice = (total + 31u) / 32u;
  uint32_t lo = lane * slice, hi = lo + slice;
  if (lo > total) lo = total;
  if (hi > total) hi = total;
  uint32_t crc = slab_crc16_run(&tb, b + 8u + lo, hi - lo);
  crc = slab_crc16_mul(crc, slab_crc16_xpow8(total - hi));
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) crc ^= __shfl_xor_sync(SLAB_FULL_MASK, crc, d);
  if (lane == 0) {
    const uint32_t field = size - 6u;
    b[2] = (uint8_t)(field >> 24); b[3] = (uint8_t)(field >> 16); b[4] = (uint8_t)(field >> 8); b[5] = (uint8_t)field;
    b[6] = (uint8_t)(crc >> 8); b[7] = (uint8_t)crc;
  }
}

#include "slab_encode_pack.cuh"
#include "slab_encode_ltmma.cuh"

#endif
