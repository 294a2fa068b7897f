/* A C client of include/b200mc.h: prices through the C ABI with no Python in between.
 *
 *   price_client <device>
 *
 * Creates an engine, runs (1) one European call with three spot-bumped scenarios on common random numbers (the latency
 * path: one option), (2) a 5-option batch of up-and-out barrier calls (parameter block staged through HBM), prints the
 * raw moments as hexadecimal doubles (exact), then destroys the engine.  tests/test_abi_exports.py compares the lines bit
 * for bit with what the Python pricer path (ctypes -> the same entry points) returns. */
#include <stdio.h>
#include <string.h>

#include "b200mc.h"

static int fail(b200mc_engine_t* eng, const char* what, int rc) {
  fprintf(stderr, "%s failed (%d): %s\n", what, rc, b200mc_last_error(eng));
  if (eng) b200mc_destroy(eng);
  return 1;
}

int main(int argc, char** argv) {
  int device = 0, rc, i;
  b200mc_engine_t* eng = NULL;
  b200mc_spec_t spec;
  b200mc_params_t p[5];
  b200mc_moments_t m[5];
  if (argc > 1) sscanf(argv[1], "%d", &device);
  if (b200mc_abi_version() != B200MC_ABI_VERSION) return fail(NULL, "ABI version check", b200mc_abi_version());
  if ((rc = b200mc_create(&eng, device)) != 0) return fail(NULL, "b200mc_create", rc);

  memset(&spec, 0, sizeof spec);
  spec.kind = B200MC_EUROPEAN, spec.antithetic = 1, spec.n_steps = 50;
  memset(p, 0, sizeof p);
  for (i = 0; i < 3; ++i) {
    p[i].S = 100.0 + 1e-4 * (1 - i), p[i].K = 100.0, p[i].T = 1.0, p[i].r = 0.05, p[i].sigma = 0.2, p[i].q = 0.01;
  }
  if ((rc = b200mc_simulate(eng, &spec, p, 1, 3, 42u, 0u, 0u, 100000u, m)) != 0) return fail(eng, "b200mc_simulate", rc);
  for (i = 0; i < 3; ++i) printf("european %d %a %a %a\n", i, m[i].sum, m[i].sum_sq, m[i].n);

  memset(&spec, 0, sizeof spec);
  spec.kind = B200MC_BARRIER, spec.n_steps = 64;
  for (i = 0; i < 5; ++i) {
    p[i].S = 100.0, p[i].K = 90.0 + 5.0 * i, p[i].T = 0.5, p[i].r = 0.03, p[i].sigma = 0.25, p[i].q = 0.0, p[i].barrier = 125.0;
  }
  if ((rc = b200mc_simulate(eng, &spec, p, 5, 1, 7u, 3u, 1000u, 200001u, m)) != 0) return fail(eng, "b200mc_simulate (batch)", rc);
  for (i = 0; i < 5; ++i) printf("barrier %d %a %a %a\n", i, m[i].sum, m[i].sum_sq, m[i].n);

  /* error path: the message comes back through the handle, nothing aborts */
  spec.n_steps = 0;
  rc = b200mc_simulate(eng, &spec, p, 5, 1, 7u, 0u, 0u, 10u, m);
  printf("invalid %d %s\n", rc, b200mc_last_error(eng));
  printf("launches %llu\n", (unsigned long long)b200mc_kernel_launches(eng));
  b200mc_destroy(eng);
  return 0;
}
