/*
 * sla_oracle.c - TEST INFRASTRUCTURE ONLY (see sla_oracle.h).
 *
 * A from-scratch, single-threaded restatement of the SLA codec's block path in plain C99.
 * Every function names the reference location whose arithmetic it restates (paths relative to the
 * reference repository root).  Floating point follows the reference's operation order exactly and
 * is compiled with -ffp-contract=off, so on the same libm the doubles are bit-identical to the
 * reference (checked in tests/test_oracle_vs_ref.py).
 */
#include "sla_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORA_PI          3.1415926535897932384626433832795029   /* src/include/private/SLAUtility.h:13 */
#define ORA_HEADER_SIZE 43u
#define ORA_MIN_BLOCK   2048u                                   /* SLAInternal.h:15 */
#define ORA_GRID        1024u                                   /* SLAInternal.h:16 */
#define ORA_BIGWEIGHT   ((double)(1UL << 24))                   /* SLAPredictor.c:16 */
#define ORA_MAX_PERIOD  256u                                    /* SLAInternal.h:9 */

static int32_t sra(int32_t v, uint32_t s) { return v >> s; }   /* arithmetic on every target we use */
static uint32_t zigzag(int32_t s) { return (s < 0) ? (uint32_t)(-(s << 1)) - 1u : (uint32_t)(s << 1); }
static int32_t unzigzag(uint32_t u) { return (int32_t)(u >> 1) ^ -(int32_t)(u & 1u); }
static uint32_t nlz32(uint32_t x) { return x ? (uint32_t)__builtin_clz(x) : 32u; }
static uint32_t log2ceil(uint32_t x) { return 32u - nlz32(x - 1u); }          /* SLAUtility.h:53 */
static int32_t sgn(int32_t v) { return (v > 0) - (v < 0); }
static double round_half_away(double d) { return (d >= 0.0) ? floor(d + 0.5) : -floor(-d + 0.5); } /* SLAUtility.c:436 */
static double log2_via_ln(double x) { return log(x) * 1.4426950408889634; }                         /* SLAUtility.c:442 */

/* ---------------------------------------------------------------- CRC-16/IBM, SLAUtility.c:322 */
uint16_t ora_crc16(const uint8_t* data, uint64_t size)
{
  uint16_t crc = 0;
  while (size--) {
    crc ^= *data++;
    for (int b = 0; b < 8; b++) crc = (uint16_t)((crc & 1u) ? (crc >> 1) ^ 0xA001u : (crc >> 1));
  }
  return crc;
}

/* --------------------------------------------------- offset_lshift, SLAEncoder.c:425-455 */
uint32_t ora_lshift_offset(const int32_t* const* input, uint32_t nch, uint32_t n, uint32_t bps)
{
  uint32_t mask = 0;
  for (uint32_t c = 0; c < nch; c++)
    for (uint32_t i = 0; i < n; i++) mask |= (uint32_t)input[c][i];
  if (mask == 0) return 0;
  uint32_t ntz = (uint32_t)__builtin_ctz(mask);
  return bps - (32u - ntz);
}

/* ------------------------------------------------------------ windows, SLAUtility.c:88-189 */
void ora_make_window(uint32_t type, double* w, uint32_t n)
{
  if (type == 0 || n == 1) { for (uint32_t i = 0; i < n; i++) w[i] = 1.0; return; }
  for (uint32_t i = 0; i < n; i++) {
    double x = (double)i / (n - 1);
    switch (type) {
      case 1: w[i] = sin(ORA_PI * x); break;
      case 2: w[i] = 0.5f - 0.5f * cos(2.0f * ORA_PI * x); break;
      case 3: w[i] = 0.42f - 0.5f * cos(2.0f * ORA_PI * x) + 0.08f * cos(4.0f * ORA_PI * x); break;
      default: w[i] = sin((ORA_PI / 2.0f) * sin(ORA_PI * x) * sin(ORA_PI * x)); break;
    }
  }
}

/* -------------------------------------------- folded autocorrelation, SLAPredictor.c:331-388 */
void ora_autocorr(const double* x, uint32_t n, double* r, uint32_t nlags)
{
  for (uint32_t k = 0; k < nlags; k++) r[k] = 0.0;   /* (the reference leaves lags >= n untouched) */
  if (nlags > n) nlags = n;
  if (nlags == 0) return;
  double acc = 0.0;
  for (uint32_t i = 0; i < n; i++) acc += x[i] * x[i];
  r[0] = acc;
  for (uint32_t lag = 1; lag < nlags; lag++) {
    const uint32_t two = lag << 1;
    const uint32_t groups = (3 * lag < n) ? 1 + (n - 3 * lag) / two : 0;
    const uint32_t span = groups * two;
    acc = 0.0;
    for (uint32_t i = 0; i < lag; i++)
      for (uint32_t l = 0; l < span; l += two)
        acc += x[l + lag + i] * (x[l + i] + x[l + two + i]);
    for (uint32_t i = 0; i < n - span - lag; i++)
      acc += x[span + lag + i] * x[span + i];
    r[lag] = acc;
  }
}

/* ------------------------ Levinson-Durbin to PARCOR, SLAPredictor.c:217-328 (u/v vectors folded) */
void ora_parcor_double(const double* x, uint32_t n, double* parcor, uint32_t order)
{
  double r[ORA_MAX_ORD + 2], a[ORA_MAX_ORD + 2], t[ORA_MAX_ORD + 2], e;
  ora_autocorr(x, n, r, order + 1);
  for (uint32_t i = 0; i <= order; i++) parcor[i] = 0.0;
  if (n < order) return;                              /* :234 */
  if (fabs(r[0]) < FLT_EPSILON) return;               /* :274 */
  for (uint32_t i = 0; i < order + 2; i++) a[i] = 0.0;
  a[0] = 1.0;
  a[1] = -r[1] / r[0];
  parcor[1] = r[1] / r[0];
  e = r[0] + r[1] * a[1];
  for (uint32_t d = 1; d < order; d++) {
    double g = 0.0;
    for (uint32_t i = 0; i < d + 1; i++) g += a[i] * r[d + 1 - i];
    g /= (-e);
    e = (1.0 - g * g) * e;
    /* a_new[i] = u[i] + g*v[i] with u = (1,a1..ad,0), v = (0,ad..a1,1) */
    for (uint32_t i = 0; i < d + 2; i++) t[i] = a[i];
    a[0] = 1.0 + g * 0.0;
    for (uint32_t i = 1; i <= d; i++) a[i] = t[i] + g * t[d + 1 - i];
    a[d + 1] = 0.0 + g * 1.0;
    parcor[d + 1] = -g;
  }
}

/* ---------------------------------------- code-length estimate, SLAPredictor.c:416-468 */
double ora_code_length(const double* x, uint32_t n, uint32_t bps, const double* parcor, uint32_t order)
{
  double pw = 0.0, vr = 0.0, len;
  for (uint32_t i = 0; i < n; i++) pw += x[i] * x[i];
  pw *= pow(2, (double)(2 * (bps - 1)));
  if (fabs(pw) <= FLT_MIN) return 0.0;
  pw = log2_via_ln(pw) - log2_via_ln((double)n);
  for (uint32_t k = 1; k <= order; k++) vr += log2_via_ln(1.0 - parcor[k] * parcor[k]);
  len = 1.9426950408889634 + 0.5 * (pw + vr);
  len /= 8;
  if (len <= 0) len = 1.0 / 8;
  return len;
}

/* --------- real FFT: the Numerical Recipes four1/realft pair the reference vendors,
 * SLAUtility.c:220-319; same butterfly order and the same trigonometric recurrences. */
static void cfft_nr(double* d /* 1-based view */, unsigned long nn, int isign)
{
  unsigned long n = nn << 1, j = 1;
  for (unsigned long i = 1; i < n; i += 2) {
    if (j > i) {
      double s = d[j]; d[j] = d[i]; d[i] = s;
      s = d[j + 1]; d[j + 1] = d[i + 1]; d[i + 1] = s;
    }
    unsigned long m = n >> 1;
    while (m >= 2 && j > m) { j -= m; m >>= 1; }
    j += m;
  }
  for (unsigned long mmax = 2; n > mmax; ) {
    unsigned long step = mmax << 1;
    double theta = isign * (6.28318530717959 / (double)mmax);
    double wt = sin(0.5 * theta), wpr = -2.0 * wt * wt, wpi = sin(theta), wr = 1.0, wi = 0.0;
    for (unsigned long m = 1; m < mmax; m += 2) {
      for (unsigned long i = m; i <= n; i += step) {
        unsigned long k = i + mmax;
        double tr = wr * d[k] - wi * d[k + 1];
        double ti = wr * d[k + 1] + wi * d[k];
        d[k] = d[i] - tr; d[k + 1] = d[i + 1] - ti;
        d[i] += tr; d[i + 1] += ti;
      }
      wt = wr;
      wr = wt * wpr - wi * wpi + wr;
      wi = wi * wpr + wt * wpi + wi;
    }
    mmax = step;
  }
}

void ora_realft(double* data, uint32_t n, int sign)
{
  double* d = data - 1;
  double c1 = 0.5, c2, theta = 3.141592653589793 / (double)(n >> 1);
  if (sign == 1) { c2 = -0.5; cfft_nr(d, n >> 1, 1); } else { c2 = 0.5; theta = -theta; }
  double wt = sin(0.5 * theta), wpr = -2.0 * wt * wt, wpi = sin(theta), wr = 1.0 + wpr, wi = wpi;
  unsigned long np3 = n + 3;
  for (unsigned long i = 2; i <= (n >> 2); i++) {
    unsigned long i1 = i + i - 1, i2 = i1 + 1, i3 = np3 - i2, i4 = i3 + 1;
    double h1r = c1 * (d[i1] + d[i3]), h1i = c1 * (d[i2] - d[i4]);
    double h2r = -c2 * (d[i2] + d[i4]), h2i = c2 * (d[i1] - d[i3]);
    d[i1] = h1r + wr * h2r - wi * h2i;
    d[i2] = h1i + wr * h2i + wi * h2r;
    d[i3] = h1r - wr * h2r + wi * h2i;
    d[i4] = -h1i + wr * h2i + wi * h2r;
    wt = wr;
    wr = wt * wpr - wi * wpi + wr;
    wi = wi * wpr + wt * wpi + wi;
  }
  if (sign == 1) {
    double h = d[1];
    d[1] = h + d[2]; d[2] = h - d[2];
  } else {
    double h = d[1];
    d[1] = c1 * (h + d[2]); d[2] = c1 * (h - d[2]);
    cfft_nr(d, n >> 1, -1);
  }
}

/* ---- small dense solver: Crout LU, implicit-scaled partial pivoting, iterative refinement with a
 * long-double residual; SLAUtility.c:487-674 (including its one-way row_scale overwrite). */
static int lu_solve(double A[ORA_MAX_TAPS][ORA_MAX_TAPS], double* b, uint32_t dim, uint32_t refinements)
{
  double LU[ORA_MAX_TAPS][ORA_MAX_TAPS], scale[ORA_MAX_TAPS], x[ORA_MAX_TAPS], err[ORA_MAX_TAPS];
  uint32_t piv[ORA_MAX_TAPS];
  memcpy(LU, A, sizeof(LU));
  memcpy(x, b, sizeof(double) * dim);
  for (uint32_t r = 0; r < dim; r++) {
    double big = 0.0;
    for (uint32_t c = 0; c < dim; c++) if (fabs(LU[r][c]) > big) big = fabs(LU[r][c]);
    if (fabs(big) <= FLT_EPSILON) return -1;
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
      for (uint32_t k = 0; k < dim; k++) { double t = LU[best][k]; LU[best][k] = LU[c][k]; LU[c][k] = t; }
      scale[best] = scale[c];
    }
    piv[c] = best;
    if (fabs(LU[c][c]) <= FLT_EPSILON) return -1;
    if (c != dim - 1) {
      double inv = 1.0 / LU[c][c];
      for (r = c + 1; r < dim; r++) LU[r][c] *= inv;
    }
  }
  for (uint32_t pass = 0; pass <= refinements; pass++) {
    double* v = (pass == 0) ? x : err;
    if (pass > 0) {
      for (uint32_t r = 0; r < dim; r++) {
        long double e = -b[r];
        for (uint32_t c = 0; c < dim; c++) e += A[r][c] * x[c];
        err[r] = (double)e;
      }
    }
    /* forward / back substitution, SLAUtility.c:579-617 */
    uint32_t first = 0;
    for (uint32_t r = 0; r < dim; r++) {
      uint32_t p = piv[r];
      double s = v[p];
      v[p] = v[r];
      if (first != 0) { for (uint32_t c = first; c < r; c++) s -= LU[r][c] * v[c]; }
      else if (s != 0.0) { first = r; }
      v[r] = s;
    }
    for (uint32_t r = dim; r-- > 0; ) {
      double s = v[r];
      for (uint32_t c = r + 1; c < dim; c++) s -= LU[r][c] * v[c];
      v[r] = s / LU[r][r];
    }
    if (pass > 0) for (uint32_t r = 0; r < dim; r++) x[r] -= err[r];
  }
  memcpy(b, x, sizeof(double) * dim);
  return 0;
}

/* ------------------------------------- long-term (pitch) analysis, SLAPredictor.c:791-980 */
int ora_longterm_analyse(const int32_t* res, uint32_t n, uint32_t fft_size, uint32_t taps,
                         uint32_t* pitch, double* coef)
{
  double* ac = (double*)malloc(sizeof(double) * fft_size);
  uint32_t cand[ORA_MAX_PERIOD], ncand = 0, i, chosen;
  double peak_max = 0.0;
  int rc = 0;

  for (i = 0; i < fft_size; i++) ac[i] = (i < n) ? (double)res[i] * pow(2.0f, -31.0f) : 0.0;
  ora_realft(ac, fft_size, 1);
  ac[0] *= ac[0];
  ac[1] *= ac[1];
  for (i = 1; i < fft_size / 2; i++) {
    double re = ac[2 * i], im = ac[2 * i + 1];
    ac[2 * i] = re * re + im * im;
    ac[2 * i + 1] = 0.0;
  }
  ora_realft(ac, fft_size, -1);

  if (fabs(ac[0]) <= FLT_MIN) {
    *pitch = 0;
    for (i = 0; i < taps; i++) coef[i] = 0.0;
    goto out;
  }
  i = 1;
  while (i < ORA_MAX_PERIOD && ncand < ORA_MAX_PERIOD) {
    uint32_t start, end, j, at = 0;
    double best = 0.0;
    for (start = i; start < ORA_MAX_PERIOD; start++)
      if (ac[start - 1] < 0.0 && ac[start] > 0.0) break;
    for (end = start + 1; end < ORA_MAX_PERIOD; end++)
      if (ac[end] > 0.0 && ac[end + 1] < 0.0) break;
    for (j = start; j <= end; j++)
      if (ac[j] > ac[j - 1] && ac[j] > ac[j + 1] && ac[j] > best) { at = j; best = ac[j]; }
    if (at != 0) {
      cand[ncand++] = at;
      if (best > peak_max) peak_max = best;
    }
    i = end + 1;
  }
  if (ncand == 0) { rc = 1; goto out; }
  for (i = 0; i < ncand; i++) if (ac[cand[i]] >= 1.0f * peak_max) break;
  chosen = cand[i];
  if (chosen < taps / 2 + 1) { rc = 1; goto out; }
  {
    double R[ORA_MAX_TAPS][ORA_MAX_TAPS], v[ORA_MAX_TAPS], mag = 0.0;
    uint32_t j, k;
    memset(R, 0, sizeof(R));
    for (j = 0; j < taps; j++)
      for (k = 0; k < taps; k++) R[j][k] = ac[(j >= k) ? (j - k) : (k - j)];
    for (j = 0; j < taps; j++) v[j] = ac[j + chosen - taps / 2];
    if (lu_solve(R, v, taps, 2) != 0) { rc = 1; goto out; }
    for (j = 0; j < taps; j++) mag += fabs(v[j]);
    if (mag >= 1.0f) {
      for (j = 0; j < taps; j++) v[j] = 0.0;
      v[taps / 2] = ac[chosen] / ac[0];
    }
    *pitch = chosen;
    for (j = 0; j < taps; j++) coef[j] = v[j];
  }
out:
  free(ac);
  return rc;
}

/* ---------------------------------- Dijkstra on the dense edge matrix, SLAPredictor.c:1521-1581 */
int ora_dijkstra(const double* adj, uint32_t stride, uint32_t nnodes, uint32_t* path, double* cost)
{
  uint8_t done[64];
  uint32_t i, cur = 0;
  if (nnodes > 64) return -1;
  for (i = 0; i < nnodes; i++) { done[i] = 0; path[i] = 0xFFFFFFFFu; cost[i] = ORA_BIGWEIGHT; }
  cost[0] = 0.0;
  for (;;) {
    double best = ORA_BIGWEIGHT;
    for (i = 0; i < nnodes; i++) if (!done[i] && cost[i] < best) { best = cost[i]; cur = i; }
    if (cur == nnodes - 1) break;
    for (i = 0; i < nnodes; i++)
      if (cost[i] > adj[cur * stride + i] + cost[cur]) { cost[i] = adj[cur * stride + i] + cost[cur]; path[i] = cur; }
    done[cur] = 1;
  }
  return 0;
}

/* --------------------------------------------- stereo mid/side, SLAUtility.c:391-433 */
void ora_ms_forward(int32_t* l, int32_t* r, uint32_t n)
{
  for (uint32_t i = 0; i < n; i++) { int32_t m = (l[i] + r[i]) >> 1, s = l[i] - r[i]; l[i] = m; r[i] = s; }
}
void ora_ms_inverse(int32_t* m, int32_t* s, uint32_t n)
{
  for (uint32_t i = 0; i < n; i++) {
    int32_t side = s[i], mid = (int32_t)(((uint32_t)m[i] << 1) | ((uint32_t)side & 1u));
    m[i] = (mid + side) >> 1; s[i] = (mid - side) >> 1;
  }
}

/* ------------------------------------------ 31/32 emphasis, SLAPredictor.c:1741-1791 */
void ora_preemphasis(int32_t* x, uint32_t n)
{
  int32_t prev = 0;
  for (uint32_t i = 0; i < n; i++) { int32_t cur = x[i]; x[i] -= sra((int32_t)((uint32_t)prev * 31u), 5); prev = cur; }
}
void ora_deemphasis(int32_t* x, uint32_t n)
{
  int32_t prev = 0;
  for (uint32_t i = 0; i < n; i++) { x[i] += sra((int32_t)((uint32_t)prev * 31u), 5); prev = x[i]; }
}

/* ------------------------------------------ PARCOR lattice, SLAPredictor.c:557-607, 722-736 */
static int32_t lat_mul(int32_t k, int32_t v) { return sra((int32_t)((uint32_t)k * (uint32_t)v + (1u << 14)), 15); }

void ora_parcor_predict(const int32_t* x, uint32_t n, const int32_t* coef, uint32_t order, int32_t* res)
{
  int32_t f[ORA_MAX_ORD + 1], b[ORA_MAX_ORD + 1];
  memset(b, 0, sizeof(b));
  for (uint32_t s = 0; s < n; s++) {
    f[0] = x[s];
    for (uint32_t m = 1; m <= order; m++) f[m] = f[m - 1] - lat_mul(coef[m], b[m - 1]);
    for (uint32_t m = order; m >= 1; m--) b[m] = b[m - 1] - lat_mul(coef[m], f[m - 1]);
    b[0] = x[s];
    res[s] = f[order];
  }
}

void ora_parcor_synth(const int32_t* res, uint32_t n, const int32_t* coef, uint32_t order, int32_t* out)
{
  int32_t b[ORA_MAX_ORD + 1];
  memset(b, 0, sizeof(b));
  for (uint32_t s = 0; s < n; s++) {
    int32_t f = res[s];
    for (uint32_t m = order; m >= 1; m--) {
      f += lat_mul(coef[m], b[m - 1]);
      b[m] = b[m - 1] - lat_mul(coef[m], f);
    }
    out[s] = f;
    b[0] = f;
  }
}

/* ------- long-term FIR / recursive filter, SLAPredictor.c:1031-1108. The reference's mirrored ring
 * buffer resolves to: tap j reads the signal at n - (pitch + taps/2 - j); the first pitch+taps/2
 * samples pass through.  "signal" is the input when predicting and the output when synthesising. */
void ora_longterm_filter(const int32_t* in, uint32_t n, uint32_t pitch, const int32_t* coef_q31,
                         uint32_t taps, int32_t* out, int synth)
{
  const uint32_t delay = pitch + (taps >> 1);
  if (out != in) memmove(out, in, sizeof(int32_t) * n);
  if (pitch == 0) return;
  if (!synth) {
    /* FIR on the input: walk backwards so that in-place use is safe */
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
    memcpy(tmp, in, sizeof(int32_t) * n);
    for (uint32_t s = delay; s < n; s++) {
      int64_t acc = (int64_t)1 << 30;
      for (uint32_t j = 0; j < taps; j++) acc += (int64_t)coef_q31[j] * tmp[s - delay + j];
      out[s] = tmp[s] - (int32_t)(acc >> 31);
    }
    free(tmp);
  } else {
    for (uint32_t s = delay; s < n; s++) {
      int64_t acc = (int64_t)1 << 30;
      for (uint32_t j = 0; j < taps; j++) acc += (int64_t)coef_q31[j] * out[s - delay + j];
      out[s] += (int32_t)(acc >> 31);
    }
  }
}

/* ---- cascaded FIR + "IIR" sign-LMS, SLAPredictor.c:121-145 (step table), 1202-1331, 1334-1463.
 * hist_x = past inputs (predict) / outputs (synth); hist_p = past predictions, both pre-filled with
 * the first `order` samples, which pass through.  All products wrap in 32 bits. */
void ora_lms_filter(const int32_t* in, uint32_t n, uint32_t order, int32_t* out, int synth)
{
  int32_t cx[64], cp[64], hx[64], hp[64];   /* h*[i] = value i+1 samples ago */
  if (order > 64) return;
  memset(cx, 0, sizeof(cx)); memset(cp, 0, sizeof(cp));
  if (out != in) memmove(out, in, sizeof(int32_t) * n);
  if (n <= order) return;
  for (uint32_t i = 0; i < order; i++) hx[i] = hp[i] = in[order - 1 - i];
  for (uint32_t s = order; s < n; s++) {
    uint32_t acc = 1u << 9;
    for (uint32_t i = 0; i < order; i++) {
      acc += (uint32_t)cx[i] * (uint32_t)hx[i];
      acc += (uint32_t)cp[i] * (uint32_t)hp[i];
    }
    int32_t pred = sra((int32_t)acc, 10);
    int32_t resid, value;
    if (synth) { resid = in[s]; value = (int32_t)((uint32_t)resid + (uint32_t)pred); out[s] = value; }
    else       { value = in[s]; resid = (int32_t)((uint32_t)value - (uint32_t)pred); out[s] = resid; }
    /* delta = sign(resid) * (bitlength(|resid|) >> 1) */
    uint32_t mag = (resid > 0) ? (uint32_t)resid : (uint32_t)(-(int64_t)resid);
    int32_t step = sgn(resid) * (int32_t)((32u - nlz32(mag)) >> 1);
    for (uint32_t i = 0; i < order; i++) { cx[i] += step * sgn(hx[i]); cp[i] += step * sgn(hp[i]); }
    for (uint32_t i = order - 1; i > 0; i--) { hx[i] = hx[i - 1]; hp[i] = hp[i - 1]; }
    hx[0] = value; hp[0] = pred;
  }
}

/* ============================================================ MSB-first bit I/O (SLABitStream.h) */
typedef struct { uint8_t* p; uint32_t cap; uint64_t bitpos; int overflow; } BitW;

static void bw_put(BitW* w, uint32_t val, uint32_t nbits)
{
  for (uint32_t i = nbits; i-- > 0; ) {
    uint64_t byte = w->bitpos >> 3;
    if (byte >= w->cap) { w->overflow = 1; w->bitpos++; continue; }
    if ((val >> i) & 1u) w->p[byte] |= (uint8_t)(0x80u >> (w->bitpos & 7));
    w->bitpos++;
  }
}
static void bw_zeros(BitW* w, uint32_t n) { w->bitpos += n; if ((w->bitpos + 7) / 8 > w->cap) w->overflow = 1; }
static void bw_align(BitW* w) { w->bitpos = (w->bitpos + 7) & ~(uint64_t)7; }

typedef struct { const uint8_t* p; uint64_t nbits; uint64_t bitpos; } BitR;

static uint32_t br_bit(BitR* r)
{
  uint32_t b = 0;
  if (r->bitpos < r->nbits) b = (r->p[r->bitpos >> 3] >> (7 - (r->bitpos & 7))) & 1u;
  r->bitpos++;
  return b;
}
static uint64_t br_get(BitR* r, uint32_t n) { uint64_t v = 0; while (n--) v = (v << 1) | br_bit(r); return v; }
static uint32_t br_zero_run(BitR* r)
{
  uint32_t run = 0;
  while (r->bitpos < r->nbits && br_bit(r) == 0) run++;
  return run;
}
static void br_align(BitR* r) { r->bitpos = (r->bitpos + 7) & ~(uint64_t)7; }

/* ================================================== entropy coder (SLACoder.c) */
/* Q8 running mean -> Rice modulus, SLACoder.c:30-31 */
static uint32_t rice_modulus(uint64_t p)
{
  uint32_t m = (uint32_t)(((p >> 1) + 128u) >> 8);
  if (m < 1) m = 1;
  return 1u << log2ceil(m);
}
/* SLACoder.c:26-28 - note the 32-bit wrap of (v << 8) and of 9 * (...) */
static uint64_t rice_update(uint64_t p, uint32_t v)
{
  uint32_t w = 9u * (uint32_t)(v << 8);
  return (uint64_t)(119u * p + w + 64u) >> 7;
}
static uint32_t rice_param_get(uint64_t p) { uint32_t m = (uint32_t)((p + 128u) >> 8); return m < 1 ? 1 : m; }

static void put_unary(BitW* w, uint32_t q) { bw_zeros(w, q); bw_put(w, 1, 1); }

/* SLACoder.c:120-138 */
static void put_gamma(BitW* w, uint32_t v)
{
  if (v == 0) { bw_put(w, 1, 1); return; }
  uint32_t nd = log2ceil(v + 2);
  bw_zeros(w, nd - 1);
  bw_put(w, v + 1, nd);
}
static uint32_t get_gamma(BitR* r)
{
  uint32_t nd = br_zero_run(r) + 1;
  if (nd == 1) return 0;
  return (uint32_t)((1UL << (nd - 1)) + br_get(r, nd - 1) - 1);
}

/* SLACoder.c:224-270, two parameters */
static void put_recursive_rice(BitW* w, uint64_t* p, uint32_t v)
{
  uint32_t m0 = rice_modulus(p[0]);
  if (v < m0) {
    put_unary(w, 0);
    if (m0 != 1) bw_put(w, v & (m0 - 1), log2ceil(m0));
    p[0] = rice_update(p[0], v);
    return;
  }
  p[0] = rice_update(p[0], v);
  v -= m0;
  uint32_t m1 = rice_modulus(p[1]);
  uint32_t q = 1 + v / m1;
  if (q < 16) put_unary(w, q);
  else { put_unary(w, 16); put_gamma(w, q - 16); }
  if (m1 != 1) bw_put(w, v & (m1 - 1), log2ceil(m1));
  p[1] = rice_update(p[1], v);
}

/* SLACoder.c:273-318 */
static uint32_t get_recursive_rice(BitR* r, uint64_t* p)
{
  uint32_t q = br_zero_run(r), v, m0 = rice_modulus(p[0]);
  if (q == 0) {
    v = (m0 != 1) ? (uint32_t)br_get(r, log2ceil(m0)) : 0;
    p[0] = rice_update(p[0], v);
    return v;
  }
  uint32_t m1 = rice_modulus(p[1]);
  if (q == 16) q += get_gamma(r);
  uint32_t tail = m1 * (q - 1) + ((m1 != 1) ? (uint32_t)br_get(r, log2ceil(m1)) : 0);
  v = m0 + tail;
  p[0] = rice_update(p[0], v);
  p[1] = rice_update(p[1], tail);
  return v;
}

/* SLACoder.c:45-82 / 85-117 */
static void put_golomb(BitW* w, uint32_t m, uint32_t v)
{
  uint32_t q = v / m, rest = v % m;
  put_unary(w, q);
  if ((m & (m - 1)) == 0) { if (m > 1) bw_put(w, rest, log2ceil(m)); return; }
  uint32_t b = log2ceil(m), cut = (1u << b) - m;
  if (rest < cut) bw_put(w, rest, b - 1); else bw_put(w, rest + cut, b);
}
static uint32_t get_golomb(BitR* r, uint32_t m)
{
  uint32_t q = br_zero_run(r);
  if ((m & (m - 1)) == 0) return (uint32_t)(q * m + br_get(r, log2ceil(m)));
  uint32_t b = log2ceil(m), cut = (1u << b) - m;
  uint64_t rest = br_get(r, b - 1);
  if (rest < cut) return (uint32_t)(q * m + rest);
  rest = (rest << 1) + br_get(r, 1);
  return (uint32_t)(q * m + rest - cut);
}

/* ============================================================ partition search */
static void load_segment(const OraParams* p, const int32_t* const* input, uint32_t n, uint32_t shift,
                         double** xd, int32_t** xi)
{
  for (uint32_t c = 0; c < p->num_channels; c++)
    for (uint32_t i = 0; i < n; i++) {
      xd[c][i] = (double)input[c][i] * pow(2, -31);
      xi[c][i] = sra(input[c][i], shift);
    }
  if (p->ch_process == 1) {
    for (uint32_t i = 0; i < n; i++) {                      /* SLAUtility.c:370-388 */
      double m = (xd[0][i] + xd[1][i]) / 2, s = xd[0][i] - xd[1][i];
      xd[0][i] = m; xd[1][i] = s;
    }
    ora_ms_forward(xi[0], xi[1], n);
  }
}

/* SLAEncoder.c:356-422 + SLAPredictor.c:1584-1705 */
int ora_search_partitions(const OraParams* p, const int32_t* const* input, uint32_t seg,
                          uint32_t min_block, uint32_t* nparts, uint32_t* parts)
{
  const uint32_t nch = p->num_channels;
  double* xd[ORA_MAX_CH]; int32_t* xi[ORA_MAX_CH];
  double adj[64 * 64], cost[64], parcor[ORA_MAX_ORD + 1];
  uint32_t path[64], first_nz, nnodes, c, i, j, hops, node;

  for (c = 0; c < nch; c++) { xd[c] = malloc(sizeof(double) * seg); xi[c] = malloc(sizeof(int32_t) * seg); }
  load_segment(p, input, seg, 32 - p->bits_per_sample, xd, xi);

  for (first_nz = 0; first_nz < seg; first_nz++) {
    int any = 0;
    for (c = 0; c < nch; c++) any |= (xi[c][first_nz] != 0);
    if (any) break;
  }
  if (first_nz >= min_block) { *nparts = 1; parts[0] = first_nz; goto done; }

  nnodes = (seg + ORA_GRID - 1) / ORA_GRID + 1;
  for (i = 0; i < nnodes; i++)
    for (j = 0; j < nnodes; j++) {
      double total = 0.0;
      adj[i * 64 + j] = ORA_BIGWEIGHT;
      if (j <= i) continue;
      uint32_t off = i * ORA_GRID, len = (j - i) * ORA_GRID;
      if (len > seg - off) len = seg - off;
      if (len < min_block || len > seg) continue;
      for (c = 0; c < nch; c++) {
        ora_parcor_double(xd[c] + off, len, parcor, p->parcor_order);
        total += len * ora_code_length(xd[c] + off, len, p->bits_per_sample, parcor, p->parcor_order);
      }
      total += 50;      /* SLAPredictor.c:20  */
      total += 300;     /* SLAInternal.h:29   */
      adj[i * 64 + j] = total;
    }
  ora_dijkstra(adj, 64, nnodes, path, cost);
  for (hops = 0, node = nnodes - 1; node != 0; node = path[node]) hops++;
  for (i = 0, node = nnodes - 1; i < hops; i++, node = path[node]) {
    uint32_t off = path[node] * ORA_GRID, len = (node - path[node]) * ORA_GRID;
    if (len > seg - off) len = seg - off;
    parts[hops - 1 - i] = len;
  }
  *nparts = hops;
done:
  for (c = 0; c < nch; c++) { free(xd[c]); free(xi[c]); }
  return 0;
}

/* ============================================================ block encode, SLAEncoder.c:458-801 */
uint32_t ora_encode_block(const OraParams* p, uint32_t lshift, const int32_t* const* input,
                          uint32_t n, uint8_t* out, uint32_t cap, OraBlock* info,
                          int32_t* const* residual_out)
{
  const uint32_t nch = p->num_channels, P = p->parcor_order, T = p->longterm_order;
  double* xd[ORA_MAX_CH]; int32_t* xi[ORA_MAX_CH]; int32_t* res[ORA_MAX_CH];
  double* win = malloc(sizeof(double) * n);
  double parcor[ORA_MAX_CH][ORA_MAX_ORD + 1], lt[ORA_MAX_CH][ORA_MAX_TAPS];
  int32_t code[ORA_MAX_CH][ORA_MAX_ORD + 1], kq[ORA_MAX_CH][ORA_MAX_ORD + 1], ltq[ORA_MAX_CH][ORA_MAX_TAPS];
  uint32_t rshift[ORA_MAX_CH], pitch[ORA_MAX_CH], init[ORA_MAX_CH];
  uint64_t rp[ORA_MAX_CH][2];
  uint32_t c, i, k, type = ORA_BLOCK_SILENT, size;
  BitW w;

  memset(parcor, 0, sizeof(parcor)); memset(lt, 0, sizeof(lt)); memset(code, 0, sizeof(code));
  memset(kq, 0, sizeof(kq)); memset(ltq, 0, sizeof(ltq)); memset(rshift, 0, sizeof(rshift));
  memset(pitch, 0, sizeof(pitch)); memset(init, 0, sizeof(init));
  for (c = 0; c < nch; c++) {
    xd[c] = malloc(sizeof(double) * n); xi[c] = malloc(sizeof(int32_t) * n); res[c] = calloc(n, sizeof(int32_t));
  }
  ora_make_window(p->window_type, win, n);
  load_segment(p, input, n, 32 - p->bits_per_sample + lshift, xd, xi);
  for (c = 0; c < nch; c++) for (i = 0; i < n; i++) if (xi[c][i] != 0) type = ORA_BLOCK_COMPRESS;

  for (c = 0; c < nch && type == ORA_BLOCK_COMPRESS; c++) {
    double prev = 0.0, est;
    uint32_t peak = 0, bw;
    for (i = 0; i < n; i++) xd[c][i] *= win[i];
    for (i = 0; i < n; i++) {                                  /* SLAPredictor.c:1794-1813 */
      double cur = xd[c][i];
      xd[c][i] -= prev * ((pow(2.0f, 5.0) - 1.0f) * pow(2.0f, -5.0));
      prev = cur;
    }
    ora_parcor_double(xd[c], n, parcor[c], P);
    est = ora_code_length(xd[c], n, p->bits_per_sample, parcor[c], P);
    est = (8 * est) / p->bits_per_sample;
    if (est >= 0.95f) { type = ORA_BLOCK_RAW; break; }
    for (i = 0; i < n; i++) {                                  /* SLAUtility.c:677-696 */
      uint32_t a = (xi[c][i] > 0) ? (uint32_t)xi[c][i] : (uint32_t)(-xi[c][i]);
      if (a > peak) peak = a;
    }
    bw = (peak > 0) ? log2ceil(peak) + 1 : 1;
    rshift[c] = (bw > 16) ? bw - 16 : 0;
    for (k = 1; k <= P; k++) {                                 /* SLAEncoder.c:573-589 */
      uint32_t qb = (k < 4) ? 16 : 8;
      int32_t lim = 1 << (qb - 1);
      int32_t q = (int32_t)round_half_away(parcor[c][k] * pow(2.0f, (double)(qb - 1)));
      if (q < -lim) q = -lim;
      if (q > lim - 1) q = lim - 1;
      code[c][k] = q;
      kq[c][k] = sra((int32_t)((uint32_t)q << (16u - qb)), rshift[c]);
    }
    memcpy(res[c], xi[c], sizeof(int32_t) * n);
    ora_preemphasis(res[c], n);
    ora_parcor_predict(res[c], n, kq[c], P, res[c]);
    if (ora_longterm_analyse(res[c], n, p->fft_size, T, &pitch[c], lt[c]) != 0 || pitch[c] >= ORA_MAX_PERIOD)
      pitch[c] = 0;
    for (k = 0; k < T; k++)
      ltq[c][k] = (int32_t)((uint32_t)(int32_t)round_half_away(lt[c][k] * pow(2.0f, 15)) << 16);
    if (pitch[c] >= 3) ora_longterm_filter(res[c], n, pitch[c], ltq[c], T, res[c], 0);
    ora_lms_filter(res[c], n, p->lms_order, res[c], 0);
  }

  /* initial Rice parameter = mean of the zig-zagged residual, SLACoder.c:361-385 */
  for (c = 0; c < nch; c++) {
    uint64_t sum = 0;
    for (i = 0; i < n; i++) sum += zigzag(res[c][i]);
    sum /= n;
    init[c] = (uint32_t)(sum > 1 ? sum : 1);
    rp[c][0] = rp[c][1] = (uint32_t)(init[c] << 8);
  }

  memset(out, 0, cap);
  w.p = out; w.cap = cap; w.bitpos = 0; w.overflow = 0;
  bw_put(&w, 0xFFFF, 16); bw_put(&w, 0, 32); bw_put(&w, 0, 16);
  bw_put(&w, n, 16); bw_put(&w, type, 2);
  for (c = 0; c < nch && type == ORA_BLOCK_COMPRESS; c++) {
    bw_put(&w, rshift[c], 4);
    for (k = 1; k <= P; k++) bw_put(&w, zigzag(code[c][k]), (k < 4) ? 16 : 8);
    if (pitch[c] >= 3) {
      bw_put(&w, 1, 1); bw_put(&w, pitch[c], 10);
      for (k = 0; k < T; k++) bw_put(&w, zigzag(sra(ltq[c][k], 16)), 16);
    } else bw_put(&w, 0, 1);
    bw_put(&w, rice_param_get(rp[c][0]), p->bits_per_sample);
  }
  bw_align(&w);
  if (type == ORA_BLOCK_RAW) {
    for (i = 0; i < n; i++)
      for (c = 0; c < nch; c++)
        bw_put(&w, zigzag(xi[c][i]), p->bits_per_sample - lshift + ((c == 1 && p->ch_process == 1) ? 1 : 0));
  } else if (type == ORA_BLOCK_COMPRESS) {
    uint64_t avg = 0;
    for (c = 0; c < nch; c++) avg += rice_param_get(rp[c][0]);
    avg /= nch;
    if (avg > 8) {                                              /* SLACoder.c:449-456 */
      for (i = 0; i < n; i++) for (c = 0; c < nch; c++) put_recursive_rice(&w, rp[c], zigzag(res[c][i]));
    } else {
      for (i = 0; i < n; i++) for (c = 0; c < nch; c++) put_golomb(&w, rice_param_get(rp[c][0]), zigzag(res[c][i]));
    }
  }
  bw_align(&w);
  size = (uint32_t)(w.bitpos >> 3);
  if (w.overflow || size > cap) size = 0;
  else {
    uint32_t field = size - 6;
    uint16_t crc = ora_crc16(out + 8, size - 8);
    out[2] = (uint8_t)(field >> 24); out[3] = (uint8_t)(field >> 16); out[4] = (uint8_t)(field >> 8); out[5] = (uint8_t)field;
    out[6] = (uint8_t)(crc >> 8); out[7] = (uint8_t)crc;
  }
  if (info) {
    info->num_samples = n; info->block_type = type; info->block_size = size;
    for (c = 0; c < nch; c++) {
      info->rshift[c] = rshift[c]; info->pitch[c] = pitch[c]; info->rice_init[c] = (uint32_t)(init[c] << 8);
      memcpy(info->parcor[c], parcor[c], sizeof(parcor[c])); memcpy(info->parcor_code[c], code[c], sizeof(code[c]));
      memcpy(info->lt[c], lt[c], sizeof(lt[c])); memcpy(info->lt_q31[c], ltq[c], sizeof(ltq[c]));
    }
  }
  if (residual_out && type == ORA_BLOCK_COMPRESS)
    for (c = 0; c < nch; c++) memcpy(residual_out[c], res[c], sizeof(int32_t) * n);
  for (c = 0; c < nch; c++) { free(xd[c]); free(xi[c]); free(res[c]); }
  free(win);
  return size;
}

/* ============================================================ file header, SLAEncoder.c:227-292 */
static void put_be(uint8_t** p, uint32_t v, int bytes) { while (bytes--) *(*p)++ = (uint8_t)(v >> (8 * bytes)); }
static uint32_t get_be(const uint8_t** p, int bytes) { uint32_t v = 0; while (bytes--) v = (v << 8) | *(*p)++; return v; }

static void write_header(const OraParams* p, uint32_t lshift, uint32_t nsamples, uint32_t nblocks,
                         uint32_t max_block_size, uint32_t max_bps, uint8_t* out)
{
  uint8_t* q = out;
  put_be(&q, 'S', 1); put_be(&q, 'L', 1); put_be(&q, '*', 1); put_be(&q, 1, 1);
  put_be(&q, ORA_HEADER_SIZE - 8, 4); put_be(&q, 0, 2); put_be(&q, 1, 4);
  put_be(&q, p->num_channels, 1); put_be(&q, nsamples, 4); put_be(&q, p->sampling_rate, 4);
  put_be(&q, p->bits_per_sample, 1); put_be(&q, lshift, 1); put_be(&q, p->parcor_order, 1);
  put_be(&q, p->longterm_order, 1); put_be(&q, p->lms_order, 1); put_be(&q, p->ch_process, 1);
  put_be(&q, nblocks, 4); put_be(&q, p->max_block_samples, 2); put_be(&q, max_block_size, 4);
  put_be(&q, max_bps, 4);
  uint16_t crc = ora_crc16(out + 10, ORA_HEADER_SIZE - 10);
  out[8] = (uint8_t)(crc >> 8); out[9] = (uint8_t)crc;
}

int ora_decode_header(const uint8_t* data, uint32_t size, OraHeader* h)
{
  const uint8_t* q = data;
  int rc = 0;
  if (size < ORA_HEADER_SIZE) return 3;
  if (q[0] != 'S' || q[1] != 'L' || q[2] != '*' || q[3] != 1) return 1;
  q += 8;
  if (get_be(&q, 2) != ora_crc16(data + 10, ORA_HEADER_SIZE - 10)) rc = 2;
  if (get_be(&q, 4) != 1) return 1;
  h->num_channels = get_be(&q, 1); h->num_samples = get_be(&q, 4); h->sampling_rate = get_be(&q, 4);
  h->bits_per_sample = get_be(&q, 1); h->offset_lshift = get_be(&q, 1); h->parcor_order = get_be(&q, 1);
  h->longterm_order = get_be(&q, 1); h->lms_order = get_be(&q, 1); h->ch_process = get_be(&q, 1);
  h->num_blocks = get_be(&q, 4); h->max_block_samples = get_be(&q, 2); h->max_block_size = get_be(&q, 4);
  h->max_bit_per_second = get_be(&q, 4);
  return rc;
}

/* ============================================================ whole-file encode, SLAEncoder.c:804-932 */
int ora_encode_whole(const OraParams* p, const int32_t* const* input, uint32_t num_samples,
                     uint8_t* out, uint32_t cap, uint32_t* out_size,
                     OraBlock* blocks, uint32_t max_blocks, uint32_t* num_blocks,
                     int32_t* const* residual_out)
{
  const uint32_t nch = p->num_channels;
  const int32_t* ptr[ORA_MAX_CH]; int32_t* rptr[ORA_MAX_CH];
  uint32_t parts[64], nparts, pos = 0, nblk = 0, cur = ORA_HEADER_SIZE, biggest = 0, peak_bps = 0, c, k;
  if (cap < ORA_HEADER_SIZE) return -4;
  uint32_t lshift = ora_lshift_offset(input, nch, num_samples, p->bits_per_sample);
  while (pos < num_samples) {
    uint32_t left = num_samples - pos;
    uint32_t seg = left < p->max_block_samples ? left : p->max_block_samples;
    uint32_t minb = left < ORA_MIN_BLOCK ? left : ORA_MIN_BLOCK;
    if (cur >= cap) return -4;
    for (c = 0; c < nch; c++) ptr[c] = input[c] + pos;
    ora_search_partitions(p, ptr, seg, minb, &nparts, parts);
    for (k = 0; k < nparts; k++) {
      OraBlock tmp, *info = (blocks && nblk < max_blocks) ? &blocks[nblk] : &tmp;
      memset(info, 0, sizeof(*info));
      for (c = 0; c < nch; c++) { ptr[c] = input[c] + pos; rptr[c] = residual_out ? residual_out[c] + pos : NULL; }
      uint32_t sz = ora_encode_block(p, lshift, ptr, parts[k], out + cur, cap - cur, info,
                                     residual_out ? rptr : NULL);
      if (sz == 0) return -4;
      info->sample_offset = pos; info->byte_offset = cur;
      cur += sz; pos += parts[k]; nblk++;
      if (sz > biggest) biggest = sz;
      uint32_t bps = (8u * sz * p->sampling_rate) / parts[k];     /* uint32 wrap is part of the format */
      if (bps > peak_bps) peak_bps = bps;
    }
  }
  write_header(p, lshift, num_samples, nblk, biggest, peak_bps, out);
  *out_size = cur;
  if (num_blocks) *num_blocks = nblk;
  return 0;
}

/* ============================================================ whole-file decode, SLADecoder.c:309-732 */
int ora_decode_whole(const uint8_t* data, uint32_t size, int check_crc,
                     int32_t* const* out, uint32_t out_capacity, uint32_t* out_samples,
                     OraBlock* blocks, uint32_t max_blocks, uint32_t* num_blocks)
{
  OraHeader h;
  uint32_t pos = 0, off = ORA_HEADER_SIZE, nblk = 0, c, i, k;
  int rc = ora_decode_header(data, size, &h);
  if (rc == 1) return 1;
  if (rc == 2) return 2;
  if (rc == 3) return 12;
  const uint32_t nch = h.num_channels, P = h.parcor_order, T = h.longterm_order;
  const uint32_t up = 32 - h.bits_per_sample + h.offset_lshift;
  int32_t* buf[ORA_MAX_CH];
  for (c = 0; c < nch; c++) buf[c] = malloc(sizeof(int32_t) * 65536);

  rc = 0;
  while (pos < h.num_samples) {
    if (off > size) { rc = 12; break; }
    const uint8_t* b = data + off;
    uint32_t avail = size - off, bsize, n, type;
    int32_t kq[ORA_MAX_CH][ORA_MAX_ORD + 1], ltq[ORA_MAX_CH][ORA_MAX_TAPS];
    uint32_t pitch[ORA_MAX_CH], rsh[ORA_MAX_CH];
    uint64_t rp[ORA_MAX_CH][2];
    BitR r;
    if (avail < 11) { rc = 12; break; }
    r.p = b; r.nbits = (uint64_t)avail * 8; r.bitpos = 0;
    if (br_get(&r, 16) != 0xFFFF) { rc = 10; break; }
    bsize = (uint32_t)br_get(&r, 32) + 6;
    uint16_t crc = (uint16_t)br_get(&r, 16);
    if (check_crc && avail >= bsize && ora_crc16(b + 8, bsize - 8) != crc) { rc = 11; break; }
    n = (uint32_t)br_get(&r, 16);
    type = (uint32_t)br_get(&r, 2);
    memset(pitch, 0, sizeof(pitch)); memset(rsh, 0, sizeof(rsh)); memset(kq, 0, sizeof(kq)); memset(ltq, 0, sizeof(ltq));
    for (c = 0; c < nch && type == ORA_BLOCK_COMPRESS; c++) {
      rsh[c] = (uint32_t)br_get(&r, 4);
      for (k = 1; k <= P; k++) {
        uint32_t qb = (k < 4) ? 16 : 8;
        int32_t q = unzigzag((uint32_t)br_get(&r, qb));
        kq[c][k] = sra((int32_t)((uint32_t)q << (16u - qb)), rsh[c]);
      }
      if (br_get(&r, 1)) {
        pitch[c] = (uint32_t)br_get(&r, 10);
        for (k = 0; k < T; k++) ltq[c][k] = (int32_t)((uint32_t)unzigzag((uint32_t)br_get(&r, 16)) << 16);
      }
      rp[c][0] = rp[c][1] = (uint32_t)((uint32_t)br_get(&r, h.bits_per_sample) << 8);
    }
    br_align(&r);
    if (bsize > avail) { rc = 12; break; }
    if (n > out_capacity - pos) { rc = 13; break; }

    if (type == ORA_BLOCK_SILENT) {
      for (c = 0; c < nch; c++) memset(buf[c], 0, sizeof(int32_t) * n);
    } else if (type == ORA_BLOCK_RAW) {
      for (i = 0; i < n; i++)
        for (c = 0; c < nch; c++)
          buf[c][i] = unzigzag((uint32_t)br_get(&r, h.bits_per_sample - h.offset_lshift + ((c == 1 && h.ch_process == 1) ? 1 : 0)));
    } else {
      uint64_t avg = 0;
      for (c = 0; c < nch; c++) avg += rice_param_get(rp[c][0]);
      avg /= nch;
      if (avg > 8) { for (i = 0; i < n; i++) for (c = 0; c < nch; c++) buf[c][i] = unzigzag(get_recursive_rice(&r, rp[c])); }
      else { for (i = 0; i < n; i++) for (c = 0; c < nch; c++) buf[c][i] = unzigzag(get_golomb(&r, rice_param_get(rp[c][0]))); }
      for (c = 0; c < nch; c++) {
        ora_lms_filter(buf[c], n, h.lms_order, buf[c], 1);
        if (pitch[c] != 0) ora_longterm_filter(buf[c], n, pitch[c], ltq[c], T, buf[c], 1);
        ora_parcor_synth(buf[c], n, kq[c], P, buf[c]);
        ora_deemphasis(buf[c], n);
      }
    }
    if (h.ch_process == 1) ora_ms_inverse(buf[0], buf[1], n);
    for (c = 0; c < nch; c++) for (i = 0; i < n; i++) out[c][pos + i] = (int32_t)((uint32_t)buf[c][i] << up);
    if (blocks && nblk < max_blocks) {
      OraBlock* info = &blocks[nblk];
      memset(info, 0, sizeof(*info));
      info->sample_offset = pos; info->num_samples = n; info->block_type = type;
      info->block_size = bsize; info->byte_offset = off;
      for (c = 0; c < nch; c++) { info->rshift[c] = rsh[c]; info->pitch[c] = pitch[c]; }
    }
    /* the reference advances by the bytes the bit reader consumed (SLADecoder.c:651,717) */
    off += (uint32_t)((r.bitpos + 7) >> 3);
    pos += n; nblk++;
  }
  for (c = 0; c < nch; c++) free(buf[c]);
  *out_samples = pos;
  if (num_blocks) *num_blocks = nblk;
  return rc;
}
