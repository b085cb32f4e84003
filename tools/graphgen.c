/*
 * graphgen.c — deterministic synthetic CSR generator for the bench/test workloads.
 * Not part of the product and not part of the oracle: it only manufactures inputs.
 *
 * gen_columns: for every row r draw deg(r) = rowptr[r+1]-rowptr[r] DISTINCT column indices uniformly
 * from [0, K) and write them sorted ascending.  Each row owns a splitmix64 stream seeded by
 * (seed, r), so the result does not depend on the number of threads.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static int cmp_int(const void *a, const void *b) {
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

void gen_columns(int M, int K, const int64_t *rowptr, uint64_t seed, int *col) {
#pragma omp parallel
  {
    size_t words = ((size_t)K + 63) / 64;
    uint64_t *bits = (uint64_t *)calloc(words, sizeof(uint64_t));
#pragma omp for schedule(dynamic, 256)
    for (int r = 0; r < M; r++) {
      int64_t lo = rowptr[r];
      int d = (int)(rowptr[r + 1] - lo);
      if (d <= 0) continue;
      if (d > K) d = K;
      uint64_t s = seed * 0xD1342543DE82EF95ull + (uint64_t)r * 0x2545F4914F6CDD1Dull + 1;
      int *out = col + lo;
      int complement = d > K / 2;
      int want = complement ? K - d : d;
      int got = 0;
      /* mark `want` distinct columns */
      while (got < want) {
        uint32_t c = (uint32_t)(((splitmix64(&s) >> 32) * (uint64_t)K) >> 32);
        uint64_t m = 1ull << (c & 63);
        if (!(bits[c >> 6] & m)) {
          bits[c >> 6] |= m;
          if (!complement) out[got] = (int)c;
          got++;
        }
      }
      if (!complement) {
        qsort(out, (size_t)d, sizeof(int), cmp_int);
        for (int i = 0; i < d; i++) bits[out[i] >> 6] = 0;
      } else {
        int n = 0;
        for (int c = 0; c < K; c++)
          if (!(bits[c >> 6] & (1ull << (c & 63)))) out[n++] = c;
        memset(bits, 0, words * sizeof(uint64_t));
      }
    }
    free(bits);
  }
}

/* x[i] = uniform float in [lo, hi) from a counter-based stream: deterministic, thread-count independent. */
void gen_uniform(int64_t n, uint64_t seed, float lo, float hi, float *x) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + (uint64_t)i;
    uint32_t u = (uint32_t)(splitmix64(&s) >> 40); /* 24 bits */
    x[i] = lo + (hi - lo) * ((float)u * (1.0f / 16777216.0f));
  }
}

/* gen_columns_community: the same, with planted community structure — rows [c*S, (c+1)*S) form community c and every column
 * of a row is drawn from the row's own community with probability p_intra (per mille), else uniformly from [0, K).  At most
 * S/2 intra-community columns per row (a hub row spills over to the whole graph).  Same degree sequence, sorted distinct columns. */
void gen_columns_community(int M, int K, const int64_t *rowptr, uint64_t seed, int S, int p_intra, int *col) {
#pragma omp parallel
  {
    size_t words = ((size_t)K + 63) / 64;
    uint64_t *bits = (uint64_t *)calloc(words, sizeof(uint64_t));
#pragma omp for schedule(dynamic, 256)
    for (int r = 0; r < M; r++) {
      int64_t lo = rowptr[r];
      int d = (int)(rowptr[r + 1] - lo);
      if (d <= 0) continue;
      if (d > K) d = K;
      uint64_t s = seed * 0xD1342543DE82EF95ull + (uint64_t)r * 0x2545F4914F6CDD1Dull + 1;
      int *out = col + lo;
      int c0 = (r / S) * S, cs = (c0 + S <= K) ? S : (K - c0 > 0 ? K - c0 : 0);
      int got = 0, intra = 0;
      while (got < d) {
        uint32_t c;
        int want_intra = cs > 0 && intra < cs / 2 && (int)((splitmix64(&s) >> 40) % 1000) < p_intra;
        if (want_intra) c = (uint32_t)c0 + (uint32_t)(((splitmix64(&s) >> 32) * (uint64_t)cs) >> 32);
        else c = (uint32_t)(((splitmix64(&s) >> 32) * (uint64_t)K) >> 32);
        uint64_t m = 1ull << (c & 63);
        if (!(bits[c >> 6] & m)) {
          bits[c >> 6] |= m;
          out[got++] = (int)c;
          if (want_intra) intra++;
        } else if (d > K / 2 && !want_intra) {
          /* dense row: walk to the next free column instead of redrawing forever */
          uint32_t cc = c;
          do { cc = (cc + 1 == (uint32_t)K) ? 0 : cc + 1; } while (bits[cc >> 6] & (1ull << (cc & 63)));
          bits[cc >> 6] |= 1ull << (cc & 63);
          out[got++] = (int)cc;
        }
      }
      qsort(out, (size_t)d, sizeof(int), cmp_int);
      for (int i = 0; i < d; i++) bits[out[i] >> 6] &= ~(1ull << (out[i] & 63));
    }
    free(bits);
  }
}
