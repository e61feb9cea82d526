// culayer_warm.cu -- times `cusimulate` (include/culayer/culayer.hpp:6-13) WARM: the reference's
// own harness (src/test_culayer.cu) times its first and only call, i.e. mostly CUDA context
// creation.  Same configuration (1000 cells, 1e6 histories, source particles from the reference's
// CPU decompose_domain); the operator is called three times on fresh copies and the best of the
// last two is printed.  Linked twice by the Makefile: with this repository's cusimulate and with
// the reference's prototype (src/culayer.cu + src/culayer_kernel.cu).  Test infrastructure.
#include <sys/time.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "culayer.hpp"
#include "layer.hpp"
#include "particle.hpp"

static double now() {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000000;
  const int cells = 1000;
  Layer layer(decompose_domain(0.0f, 1.0f, sqrtf(2.0f) / 2.0f, 1, 0, cells, n, 0.0f));
  layer.create_particles(layer.nb_particles_create);   // src/layer.cpp:89-121: materialise the source
  std::vector<Particle> src(layer.particles);
  double best = 1e30;
  double sum0 = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    std::vector<Particle> p(src);
    std::vector<float> w(cells, 0.0f);
    const double t0 = now();
    cusimulate((int)p.size(), p.data(), layer.sigs.data(), layer.absorption_rates.data(), w.data(), 0,
               cells, layer.dx);
    const double dt = now() - t0;
    double s = 0.0;
    for (int j = 0; j < cells; ++j) s += w[j];
    if (rep == 0) sum0 = s;
    else if (dt < best) best = dt;
    printf("call %d: %.6f seconds, sum(weights_absorbed) = %.6f\n", rep, dt, s);
  }
  printf("GPU warm = %.6f seconds (%d histories, tally %.6f)\n", best, n, sum0);
  return 0;
}
