/* Plain-C use of the drop-in boundary (include/ptf_b200.h): the same calls a Julia `ccall` wrapper makes.
 *   gcc -std=c99 -Iinclude examples/c_api_example.c -Lpassivetracerflows.jl_b200 -lptf_b200 -lm \
 *       -Wl,-rpath,$PWD/passivetracerflows.jl_b200 -o /tmp/c_api_example && /tmp/c_api_example
 * 2-D steady cellular flow (examples/cellularflow.jl of the reference), nx = ny = 256, RK4, 100 steps; prints the tracer
 * mean and variance before and after.  Without a CUDA device it prints the library's error and exits with status 2:
 * there is no CPU fallback. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptf_b200.h"

#define CHECK(call)                                                                      \
  do {                                                                                   \
    int32_t rc_ = (call);                                                                \
    if (rc_ != PTF_OK) {                                                                 \
      fprintf(stderr, "%s -> %s: %s\n", #call, ptf_error_string(rc_), ptf_last_error(h)); \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

int main(void) {
  const int64_t n = 256;
  const double L = 2.0 * 3.14159265358979323846, psi0 = 0.2;
  ptf_handle* h = NULL;
  ptf_desc d;
  ptf_desc_init(&d); /* reference defaults: 2-D, 128^2, 2*pi, kappa 0.1, dt 0.01, RK4 */
  d.n[0] = d.n[1] = n;
  d.kappa[0] = d.kappa[1] = 0.002;
  d.dt = 0.005;
  d.flow_kind = PTF_FLOW_STEADY;
  CHECK(ptf_create(&d, &h));

  double* u = (double*)malloc(sizeof(double) * n * n);
  double* v = (double*)malloc(sizeof(double) * n * n);
  double* c = (double*)malloc(sizeof(double) * n * n);
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < n; ++i) {
      const double x = -L / 2 + (L / n) * i, y = -L / 2 + (L / n) * j; /* FourierFlows grid origin */
      u[j * n + i] = psi0 * cos(x) * sin(y);                           /* examples/cellularflow.jl:52-58 */
      v[j * n + i] = -psi0 * sin(x) * cos(y);
      c[j * n + i] = 0.5 * exp(-((x - 0.2 * L) * (x - 0.2 * L) + y * y) / (2 * 0.15 * 0.15));
    }
  CHECK(ptf_set_velocity(h, 0, u, n * n));
  CHECK(ptf_set_velocity(h, 1, v, n * n));
  CHECK(ptf_set_c(h, c, 0));

  double mean, var, smax, t;
  int64_t step;
  CHECK(ptf_diag(h, &mean, &var, &smax));
  printf("t = 0:      mean(c) = %.12f  var(c) = %.12f\n", mean, var);
  CHECK(ptf_step(h, 100)); /* stepforward!(prob, 100) */
  CHECK(ptf_get_c(h, c));  /* updatevars!(prob); prob.vars.c */
  CHECK(ptf_get_clock(h, &t, &step, NULL));
  CHECK(ptf_diag(h, &mean, &var, &smax));
  printf("t = %.3f:  mean(c) = %.12f  var(c) = %.12f  (step %lld, c[0] = %.6e)\n", t, mean, var, (long long)step, c[0]);
  ptf_destroy(h);
  free(u);
  free(v);
  free(c);
  return 0;
}
