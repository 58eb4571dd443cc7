/* A torch-free, Python-free caller of the C ABI (include/ia_b200.h): what a CPU-side host program binds.
 *
 *   gcc -O2 -I include examples/pair_score_host.c -o /tmp/pair_score_host \
 *       -L item_alignment_b200 -lia_b200 -Wl,-rpath,$PWD/item_alignment_b200 -lm
 *   /tmp/pair_score_host cosine 1000 64
 *
 * Fills two [n, d] fp32 matrices from a fixed linear congruential generator, scores the pairs through
 * ia_pair_score_host (host buffers in, host buffers out), and prints one line per pair group so that a test can
 * recompute the same numbers independently (tests/test_gpu_c_example.py). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ia_b200.h"

static unsigned lcg_state = 20221009u;
static float lcg_unit(void) { /* uniform in [-1, 1), exactly representable arithmetic */
  lcg_state = lcg_state * 1664525u + 1013904223u;
  return (float)(lcg_state >> 8) * (1.0f / 8388608.0f) - 1.0f;
}

int main(int argc, char** argv) {
  const char* name = argc > 1 ? argv[1] : "cosine";
  const long n = argc > 2 ? atol(argv[2]) : 1000;
  const long d = argc > 3 ? atol(argv[3]) : 64;
  int measure;
  if (!strcmp(name, "inner_product")) measure = IA_INNER;
  else if (!strcmp(name, "cosine")) measure = IA_COSINE;
  else if (!strcmp(name, "l1")) measure = IA_L1;
  else if (!strcmp(name, "l2")) measure = IA_L2;
  else { fprintf(stderr, "Unsupported similarty measure: %s\n", name); return 2; }

  float* x = malloc(sizeof(float) * n * d);
  float* y = malloc(sizeof(float) * n * d);
  float* sim = malloc(sizeof(float) * n);
  float* probs = malloc(sizeof(float) * n);
  unsigned char* labels = malloc(n);
  if (!x || !y || !sim || !probs || !labels) return 3;
  for (long i = 0; i < n * d; ++i) x[i] = lcg_unit();
  for (long i = 0; i < n * d; ++i) y[i] = (i % 3 == 0) ? x[i] : lcg_unit();

  const int rc = ia_pair_score_host(measure, IA_F32, x, y, n, d, sim, probs, 0.5, labels, 0);
  if (rc != IA_OK) { fprintf(stderr, "ia_pair_score_host failed (%d): %s\n", rc, ia_last_error()); return 1; }

  long positives = 0;
  double sum_sim = 0.0, sum_probs = 0.0;
  for (long i = 0; i < n; ++i) { positives += labels[i]; sum_sim += sim[i]; sum_probs += probs[i]; }
  printf("%s n=%ld d=%ld positives=%ld sum_sim=%.9g sum_probs=%.9g launches=%lld\n", ia_version(), n, d, positives, sum_sim,
         sum_probs, (long long)ia_launch_count());
  for (long i = 0; i < n && i < 8; ++i) printf("pair %ld sim=%.9g probs=%.9g label=%d\n", i, sim[i], probs[i], (int)labels[i]);
  free(x); free(y); free(sim); free(probs); free(labels);
  return 0;
}
