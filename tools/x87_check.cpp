#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#define X87_FN static inline
#define X87_CLZ(x) __builtin_clzll(x)
#define __host__
#define __device__
#define __forceinline__ inline
#include <stdint.h>
/* the accumulator lives in sla_b200/csrc/slab_common.cuh; this check pulls in just that part */
#include "x87_part.h"
int main() {
  srand(1);
  long bad = 0, n = 0;
  for (int it = 0; it < 3000000; it++) {
    int terms = 1 + rand() % 7;
    double b = ldexp((double)rand() / RAND_MAX - 0.5, rand() % 40 - 20);
    long double e = -b;
    SlabX87 s = slab_x87_from_double(-b);
    for (int k = 0; k < terms; k++) {
      double a = ldexp((double)rand() / RAND_MAX - 0.5, rand() % 40 - 20) * (1.0 + 1e-9 * (rand() % 1000));
      double x = ldexp((double)rand() / RAND_MAX - 0.5, rand() % 30 - 15);
      if (it % 5 == 0 && k == terms - 1) { a = b; x = 1.0; }         /* cancellation */
      if (it % 7 == 0) { a = (double)(rand() % 1000 - 500); x = (double)(rand() % 1000 - 500); }
      volatile double p = a * x;
      e += p;
      s = slab_x87_add(s, slab_x87_from_double(p));
    }
    double want = (double)e, got = slab_x87_to_double(s);
    n++;
    if (memcmp(&want, &got, 8) != 0 && !(want == 0.0 && got == 0.0)) { if (bad < 10) printf("mismatch %a vs %a\n", want, got); bad++; }
  }
  printf("%ld / %ld mismatches\n", bad, n);
  return bad != 0;
}
