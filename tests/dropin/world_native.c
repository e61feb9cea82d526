/* world_native.c -- a whole run driven from plain C through include/mcb200.h: what the
 * reference's `main` + Worker::spin + gather_weights_absorbed do (src/main.cpp:15-96,
 * src/worker_sync.cpp:24-135, src/worker.cpp:183-216), with K ranks in ONE process (one host
 * thread drives the GPUs; ranks share a GPU when there are fewer GPUs than ranks).
 *
 *   world_native K nb_particles [nb_cells] > tally.txt
 *
 * stdout: one line "events scatters n_left n_right n_dead" and then nb_cells lines "%.17g" of
 * weights_absorbed.  Test infrastructure (tests/test_gpu_dropin.py compares with the oracle). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcb200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != MCB200_OK) {                                                      \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, mcb200_last_error());        \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

int main(int argc, char **argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 3;
  const long long n = argc > 2 ? atoll(argv[2]) : 20000;
  const int cells = argc > 3 ? atoi(argv[3]) : 1000;
  const int ndev = mcb200_device_count();
  if (ndev <= 0) {
    fprintf(stderr, "no CUDA device (there is no CPU fallback)\n");
    return 2;
  }
  mcb200_world **w = calloc((size_t)K, sizeof *w);
  mcb200_world_result *res = calloc((size_t)K, sizeof *res);
  double *tally = calloc((size_t)cells, sizeof *tally);
  for (int r = 0; r < K; ++r) {
    mcb200_world_desc d;
    memset(&d, 0, sizeof d);
    d.abi_version = MCB200_ABI_VERSION;
    d.device = r % ndev;
    d.rank = r;
    d.world_size = K;
    d.x_min = 0.0f;                        /* config.yaml:1-10 */
    d.x_max = 1.0f;
    d.x_ini = sqrtf(2.0f) / 2.0f;
    d.nb_cells = cells;
    d.particle_min_weight = 9.99999996e-13f;
    d.max_ctas = K > ndev ? 592 / ((K + ndev - 1) / ndev) : 0;   /* ranks sharing a GPU must all be resident */
    CHECK(mcb200_world_create(&d, &w[r]));
    CHECK(mcb200_world_set_option(w[r], "max_run_ms", 60000));
  }
  for (int a = 0; a < K; ++a)
    for (int b = 0; b < K; ++b)
      if (a != b) CHECK(mcb200_world_connect_local(w[a], w[b]));
  CHECK(mcb200_world_run(w, K, n, 5127801ull, res));       /* Worker::spin, every rank */
  CHECK(mcb200_world_gather_tally_f64(w, K, tally));       /* Worker::gather_weights_absorbed */
  long long ev = 0, sc = 0, nl = 0, nr = 0, nd = 0;
  for (int r = 0; r < K; ++r) {
    ev += res[r].events;
    sc += res[r].scatters;
    nl += res[r].n_left;
    nr += res[r].n_right;
    nd += res[r].n_dead;
  }
  printf("%lld %lld %lld %lld %lld\n", ev, sc, nl, nr, nd);
  for (int c = 0; c < cells; ++c) printf("%.17g\n", tally[c]);
  for (int r = 0; r < K; ++r) CHECK(mcb200_world_disconnect(w[r]));
  for (int r = 0; r < K; ++r) mcb200_world_destroy(w[r]);
  free(w);
  free(res);
  free(tally);
  return 0;
}
