/* slab_encode_ltmma.cuh - E6a on the tensor cores: the 260 exact lag sums of the long-term analysis
 * (the autocorrelation of SLALongTermCalculator_CalculateCoef, src/SLAPredictor.c:791-853) as a banded Gram
 * product.
 *
 * The block's residual y[0..n) is cut into frames of 64 samples.  With A[f][j] = y[64 f + j] (j < 64) and
 * B[f][j'] = y[64 f + j'] (j' < 324, i.e. running into the following frames),
 *       C = A^T B,   C[j][j'] = sum_f y[64 f + j] y[64 f + j'],   R[k] = sum_j C[j][j + k]
 * - the lags are the diagonals of a 64 x 324 product whose reduction dimension is the frame index.  Only
 * the band 0 <= j' - j < 260 is needed: 35 diagonal strips of 16 x 8 tiles.
 * Exactness: integers only.  y is split into 8-bit limbs (two for |y| < 2^15, three for |y| < 2^23; top limb
 * signed, lower limbs unsigned), every limb pair is one s32-accumulating integer MMA
 * (mma.sync.m16n8k32 .s8/.u8), partial sums stay far below 2^31 (<= 256 frames of <= 2^16 per term, then 64
 * rows), and the limb products are recombined in int64.  The result equals the scalar kernel's bit for bit;
 * blocks with larger residuals keep the scalar kernel (k_enc_ltcorr).
 *
 * Measured on B200: mma.sync s8 sustains 574 T multiply-adds/s against 13 T/s of IMAD (tools/micro/imma_bench.cu),
 * so even with 4 (9) limb products and a 25 % wider product than the band the tensor cores are an order of
 * magnitude ahead of the direct O(260 n) form.
 *
 * Shared memory: per limb a transposed plane T[j][f] (frame index contiguous, row stride LTM_FP bytes with
 * LTM_FP / 4 = 4 mod 32, so that the 32 fragment loads of a warp - 8 rows x 4 consecutive words - hit 32
 * different banks).  An A fragment register is 4 consecutive frames of one row: one aligned word.  A B
 * fragment register is 4 consecutive frames of column j' = row (j' mod 64), shifted by j' / 64 frames: two
 * aligned words and a byte permute.  A warp owns a diagonal strip (tiles with n-tile * 8 - m-tile * 16 = d):
 * the tiles of a strip share their diagonals, so their products are summed in the accumulators and reduced
 * once: 128 shared-memory atomics per strip instead of per tile. */
#ifndef SLAB_ENCODE_LTMMA_CUH
#define SLAB_ENCODE_LTMMA_CUH

#define LTM_M        64u                       /* samples per frame */
#define LTM_MAXF     256u                      /* frames: 16384 / 64 */
#define LTM_FP       272u                      /* row stride in bytes: >= LTM_MAXF + 6 + 3, (LTM_FP / 4) % 32 == 4 */
#define LTM_NT       41u                       /* n-tiles of 8 columns: j' < 328 */
#define LTM_STRIPS   35u                       /* d = 0, 8, ..., 272 */

/* D(16x8, s32) += A(16x32, 8-bit, row) * B(32x8, 8-bit, col); AS / BS: operand is signed */
template <bool AS, bool BS>
__device__ __forceinline__ void ltm_mma(int32_t (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
#ifndef SLAB_EMUL
  if (AS && BS)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else if (AS && !BS)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else if (!AS && BS)
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  else
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
#else
  /* simulator: the fragment layout of PTX ISA "Matrix Fragments for mma.m16n8k32" spelled out.  Lane = 4 g + t
   * holds a0 = A[g][4t..4t+3], a1 = A[g+8][4t..], a2 = A[g][16+4t..], a3 = A[g+8][16+4t..], b0 = B[4t..4t+3][g],
   * b1 = B[16+4t..][g], c0 = C[g][2t], c1 = C[g][2t+1], c2 = C[g+8][2t], c3 = C[g+8][2t+1].  The lanes of a warp
   * deposit their fragments, meet, and each computes its four outputs from the deposited matrices. */
  static uint32_t dep_a[32][32][4], dep_b[32][32][2];
  const int lane = (int)(threadIdx.x & 31u), wp = (int)(threadIdx.x >> 5), g = lane >> 2, t = lane & 3;
  for (int i = 0; i < 4; i++) dep_a[wp][lane][i] = a[i];
  for (int i = 0; i < 2; i++) dep_b[wp][lane][i] = b[i];
  __syncwarp();
  for (int e = 0; e < 4; e++) {
    const int row = g + ((e & 2) ? 8 : 0), col = 2 * t + (e & 1);
    int32_t acc = 0;
    for (int k = 0; k < 32; k++) {
      const uint32_t aw = dep_a[wp][4 * (row & 7) + ((k & 15) >> 2)][(row >> 3) + ((k >> 4) << 1)];
      const uint32_t bw = dep_b[wp][4 * col + ((k & 15) >> 2)][k >> 4];
      const int32_t av = AS ? (int32_t)(int8_t)(aw >> (8 * (k & 3))) : (int32_t)((aw >> (8 * (k & 3))) & 0xFFu);
      const int32_t bv = BS ? (int32_t)(int8_t)(bw >> (8 * (k & 3))) : (int32_t)((bw >> (8 * (k & 3))) & 0xFFu);
      acc += av * bv;
    }
    c[e] += acc;
  }
  __syncwarp();
#endif
}

/* one CTA (256 threads) per block x channel; NL = number of 8-bit limbs (2 or 3) */
template <int NL>
__device__ __forceinline__ void ltm_block(const int32_t* __restrict__ src, uint32_t n, unsigned char* T,
                                          unsigned long long* Rsm, uint32_t tid)
{
  const uint32_t nframes = (n + LTM_M - 1u) / LTM_M;                  /* <= 256 */
  const uint32_t ksteps = (nframes + 31u) / 32u;
  /* ---- limb planes T[limb][j][f], zero beyond the block ---- */
  for (uint32_t i = tid; i < (uint32_t)NL * LTM_M * LTM_FP / 4u; i += 256u) reinterpret_cast<uint32_t*>(T)[i] = 0u;
  for (uint32_t i = tid; i < 264u; i += 256u) Rsm[i] = 0ull;
  __syncthreads();
  for (uint32_t i = tid; i < n; i += 256u) {
    const int32_t v = src[i];
    const uint32_t j = i & (LTM_M - 1u), f = i / LTM_M;
#pragma unroll
    for (int l = 0; l < NL; l++) T[((uint32_t)l * LTM_M + j) * LTM_FP + f] = (unsigned char)((uint32_t)v >> (8 * l));
  }
  __syncthreads();
  const uint32_t lane = tid & 31u, warp = tid >> 5, g = lane >> 2, t4 = lane & 3u;
  for (uint32_t strip = warp; strip < LTM_STRIPS; strip += 8u) {
    const uint32_t d = 8u * strip;                                    /* n-tile * 8 - m-tile * 16 */
    int32_t c[NL][NL][4];
#pragma unroll
    for (int x = 0; x < NL; x++)
#pragma unroll
      for (int y2 = 0; y2 < NL; y2++)
#pragma unroll
        for (int e = 0; e < 4; e++) c[x][y2][e] = 0;
    for (uint32_t mt = 0; mt < 4u; mt++) {
      const uint32_t col0 = d + 16u * mt;                             /* first column j' of the tile */
      if (col0 / 8u >= LTM_NT) break;
      const uint32_t jb = col0 + g;                                   /* this lane's B column */
      const uint32_t brow = jb & (LTM_M - 1u), bshift = jb / LTM_M;   /* row of T, frames ahead */
      const uint32_t arow = 16u * mt + g;
      for (uint32_t ks = 0; ks < ksteps; ks++) {
        const uint32_t f0 = 32u * ks + 4u * t4;
        uint32_t a[NL][4], b[NL][2];
#pragma unroll
        for (int l = 0; l < NL; l++) {
          const unsigned char* Ta = T + ((uint32_t)l * LTM_M + arow) * LTM_FP + f0;
          a[l][0] = *reinterpret_cast<const uint32_t*>(Ta);
          a[l][1] = *reinterpret_cast<const uint32_t*>(Ta + 8u * LTM_FP);
          a[l][2] = *reinterpret_cast<const uint32_t*>(Ta + 16u);
          a[l][3] = *reinterpret_cast<const uint32_t*>(Ta + 8u * LTM_FP + 16u);
          /* frames f0 + bshift .. + 3 of row brow: unaligned by bshift mod 4 bytes */
          const unsigned char* Tb = T + ((uint32_t)l * LTM_M + brow) * LTM_FP + ((f0 + bshift) & ~3u);
          const uint32_t sel = 0x3210u + 0x1111u * ((f0 + bshift) & 3u);
          const uint32_t w0 = *reinterpret_cast<const uint32_t*>(Tb), w1 = *reinterpret_cast<const uint32_t*>(Tb + 4u);
          const uint32_t w2 = *reinterpret_cast<const uint32_t*>(Tb + 16u), w3 = *reinterpret_cast<const uint32_t*>(Tb + 20u);
          b[l][0] = __byte_perm(w0, w1, sel);
          b[l][1] = __byte_perm(w2, w3, sel);
        }
#pragma unroll
        for (int x = 0; x < NL; x++)
#pragma unroll
          for (int y2 = 0; y2 < NL; y2++) {
            /* the top limb is signed, the lower ones are unsigned */
            if (x == NL - 1 && y2 == NL - 1) ltm_mma<true, true>(c[x][y2], a[x], b[y2]);
            else if (x == NL - 1) ltm_mma<true, false>(c[x][y2], a[x], b[y2]);
            else if (y2 == NL - 1) ltm_mma<false, true>(c[x][y2], a[x], b[y2]);
            else ltm_mma<false, false>(c[x][y2], a[x], b[y2]);
          }
      }
    }
    /* element e of the accumulator tile: row g (+8 for e >= 2), column 2 t4 (+1 for odd e): lag = d + col - row */
#pragma unroll
    for (int e = 0; e < 4; e++) {
      long long v = 0;
#pragma unroll
      for (int x = 0; x < NL; x++)
#pragma unroll
        for (int y2 = 0; y2 < NL; y2++) v += (long long)c[x][y2][e] * (1ll << (8 * (x + y2)));
      const int lag = (int)d + (int)(2u * t4 + (uint32_t)(e & 1)) - (int)(g + ((e & 2) ? 8u : 0u));
      if (lag >= 0 && lag < (int)SLAB_NUM_LTLAGS && v != 0) atomicAdd(&Rsm[lag], (unsigned long long)v);
    }
  }
  __syncthreads();
}

#define LTM_SMEM(NL) ((size_t)(NL) * LTM_M * LTM_FP + 16u)

/* E6a, tensor-core form.  Blocks whose residual does not fit three limbs are flagged in `wide` for the scalar
 * kernel.  The epilogue (scaling, risk detection for the faithful-FFT fallback) is the scalar kernel's. */
__global__ void __launch_bounds__(256) k_enc_ltcorr_mma(EncShape sh,
    const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ blk_len,
    const uint32_t* __restrict__ blk_type, const int32_t* __restrict__ r1, double* __restrict__ ac_out,
    uint32_t* __restrict__ risk_list, uint32_t* __restrict__ risk_count, uint32_t* __restrict__ wide,
    uint32_t* __restrict__ wide_count)
{
  SLAB_DYN_SMEM(unsigned char, T);
  __shared__ unsigned long long Rsm[264];
  __shared__ double lagv[264];
  __shared__ uint32_t red_u[8];
  __shared__ int s_risk;
  __shared__ uint32_t s_npeaks;
  __shared__ uint16_t s_peaks[264];
  const uint32_t bc = blockIdx.x, b = bc / sh.nch, c = bc - b * sh.nch, tid = threadIdx.x;
  if (tid == 0) wide[bc] = 0u;
  if (blk_type[b] != SLAB_BLOCK_COMPRESS) return;
  const uint32_t n = blk_len[b];
  const int32_t* src = r1 + (size_t)c * sh.NP + blk_start[b];        /* blk_start = padded starts here */
  uint32_t maxabs = 0;
  for (uint32_t i = tid; i < n; i += 256u) {
    const int32_t v = src[i];
    const uint32_t a = (v < 0) ? (0u - (uint32_t)v) : (uint32_t)v;
    maxabs = a > maxabs ? a : maxabs;
  }
#pragma unroll
  for (int dd = 16; dd > 0; dd >>= 1) { const uint32_t o = __shfl_xor_sync(SLAB_FULL_MASK, maxabs, dd); maxabs = o > maxabs ? o : maxabs; }
  if ((tid & 31u) == 0) red_u[tid >> 5] = maxabs;
  __syncthreads();
  maxabs = 0;
  for (uint32_t w = 0; w < 8u; w++) maxabs = red_u[w] > maxabs ? red_u[w] : maxabs;
  if (maxabs >= (1u << 23) || n > LTM_M * LTM_MAXF) {               /* the scalar kernel takes this one */
    if (tid == 0) { wide[bc] = 1u; atomicAdd(wide_count, 1u); }
    return;
  }
  if (maxabs < (1u << 15)) ltm_block<2>(src, n, T, Rsm, tid);
  else ltm_block<3>(src, n, T, Rsm, tid);
  if (tid == 0) s_risk = 0;
  for (uint32_t t = tid; t < SLAB_NUM_LTLAGS; t += 256u) {
    const double v = (double)(long long)Rsm[t];
    ac_out[(size_t)bc * 264u + t] = v * sh.ac_scale;
    lagv[t] = v;
  }
  __syncthreads();
  if (risk_list != nullptr) {
    /* see k_enc_ltcorr: which decisions of the pitch picker could the reference's FFT round-off turn? */
    const double tol = fabs(lagv[0]) * 1e-11;
    if (tid == 0) s_npeaks = 0;
    __syncthreads();
    if (fabs(lagv[0]) > 0.0) {
      for (uint32_t t = tid; t < 258u; t += 256u) {
        const double v = lagv[t];
        if (fabs(v) <= tol || fabs(v - lagv[t + 1u]) <= tol) s_risk = 1;
        if (t >= 1u && t < 257u && v > 0.0 && v > lagv[t - 1u] && v > lagv[t + 1u]) s_peaks[atomicAdd(&s_npeaks, 1u)] = (uint16_t)t;
      }
    }
    __syncthreads();
    const uint32_t np = s_npeaks;
    for (uint32_t i = tid; i + 1u < np; i += 256u) {
      const double v = lagv[s_peaks[i]];
      for (uint32_t j = i + 1u; j < np; j++) if (fabs(v - lagv[s_peaks[j]]) <= tol) s_risk = 1;
    }
    __syncthreads();
    if (tid == 0 && s_risk) risk_list[atomicAdd(risk_count, 1u)] = bc;
  }
}

#endif
