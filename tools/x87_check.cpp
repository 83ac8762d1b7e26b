#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
// Check of the software x87 accumulator in sla_b200/csrc/slab_common.cuh against native long double.
// build + run (x86-64):  python tools/x87_check.py
#define __host__
#define __device__
#define __forceinline__ inline
#include "x87_part.h"     // the SlabX87 section of slab_common.cuh, cut out by tools/x87_check.py
int main() {
  srand(1);
  long bad = 0, n = 0;
  for (int it = 0; it < 4000000; it++) {
    int terms = 1 + rand() % 7;
    double b = ldexp((double)rand() / RAND_MAX - 0.5, rand() % 40 - 20);
    long double e = -b;
    SlabX87 s = slab_x87_from_double(-b);
    for (int k = 0; k < terms; k++) {
      double a = ldexp((double)rand() / RAND_MAX - 0.5, rand() % 40 - 20) * (1.0 + 1e-9 * (rand() % 1000));
      double x = ldexp((double)rand() / RAND_MAX - 0.5, rand() % 30 - 15);
      if (it % 5 == 0 && k == terms - 1) { a = b; x = 1.0; }         /* cancellation */
      if (it % 5 == 1 && k == terms - 1) { a = b * (1.0 + ldexp(1.0, -(rand() % 60))); x = 1.0; }   /* near cancellation */
      if (it % 7 == 0) { a = (double)(rand() % 1000 - 500); x = (double)(rand() % 1000 - 500); }
      if (it % 11 == 0) { a = ldexp(a, rand() % 200 - 100); }        /* far apart exponents */
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
