/* lut_shim.c -- see lut_shim.h.  Plain C, compiled by gcc; everything CUDA is behind
 * include/sboxgates_b200.h. */
#define _POSIX_C_SOURCE 200809L
#include "lut_shim.h"

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sboxgates_b200.h"

_Static_assert(sizeof(sbg_ttable) == 32, "ttable must be 32 bytes (state.h:64-68)");
_Static_assert(sizeof(sbg_gate) == 64, "gate must be 64 bytes (state.h:72-79)");
_Static_assert(offsetof(sbg_gate, type) == 32 && offsetof(sbg_gate, in1) == 36
    && offsetof(sbg_gate, function) == 42, "gate field offsets (state.h:72-79)");
_Static_assert(offsetof(sbg_state, num_gates) == 10 && offsetof(sbg_state, gates) == 32,
    "state field offsets (state.h:81-88)");
_Static_assert(sizeof(sbg_state) == 32032, "state must be 32,032 bytes (state.h:81-88)");

static sbg_handle *g_handle = NULL;
static uint64_t g_calls[2] = {0, 0};
static double g_seconds[2] = {0.0, 0.0};
static double g_kernel_ms[4] = {0.0, 0.0, 0.0, 0.0}; /* search5, filter7, sort, decomp7 */

static void shim_exit(void) {
  if (g_handle != NULL) {
    if (getenv("SBG_SHIM_STATS") != NULL) {
      fprintf(stderr, "[sbg] search_5lut: %llu calls %.3f s; search_7lut: %llu calls %.3f s; "
          "%llu kernel launches; kernel time: search5 %.3f s, filter7 %.3f s, sort %.3f s, "
          "decomp7 %.3f s\n", (unsigned long long)g_calls[0], g_seconds[0],
          (unsigned long long)g_calls[1], g_seconds[1],
          (unsigned long long)sbg_launch_count(g_handle), 1e-3 * g_kernel_ms[0],
          1e-3 * g_kernel_ms[1], 1e-3 * g_kernel_ms[2], 1e-3 * g_kernel_ms[3]);
    }
    sbg_destroy(g_handle);
    g_handle = NULL;
  }
}

/* Errors are fatal, as in the reference, whose internal inconsistencies are assert()s
   (lut.c:118-119, 201, 452; sboxgates.h:31-44). */
static void die(const char *what, int rc) {
  fprintf(stderr, "sboxgates_b200: %s failed (%d): %s\n", what, rc,
      g_handle != NULL ? sbg_last_error(g_handle) : "no handle");
  abort();
}

static sbg_handle *handle(void) {
  if (g_handle == NULL) {
    const char *dev = getenv("SBG_DEVICE");
    int rc = sbg_create(&g_handle, dev != NULL ? atoi(dev) : 0);
    if (rc != SBG_OK) die("sbg_create", rc);
    atexit(shim_exit);
  }
  return g_handle;
}

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void load(sbg_handle *h, const sbg_state *st, const sbg_ttable target, const sbg_ttable mask,
    const int8_t *inbits) {
  static uint64_t flat[SBG_SHIM_MAX_GATES * 4];
  uint64_t t[4], m[4];
  const int n = st->num_gates;
  for (int i = 0; i < n; i++) {
    memcpy(flat + 4 * i, &st->gates[i].table, 32);
  }
  memcpy(t, &target, 32);
  memcpy(m, &mask, 32);
  int rc = sbg_load_problem(h, flat, n, t, m, inbits);
  if (rc != SBG_OK) die("sbg_load_problem", rc);
}

/* get_lut_function's random fill of never-constrained LUT bits (lut.c:104-106): one draw iff some
   inner cell was not seen under the mask. */
static uint8_t fill_dont_cares(const sbg_result *res) {
  uint8_t fi = res->func_inner;
  if (res->inner_seen != 0xff) {
    fi |= (uint8_t)(~res->inner_seen & (uint8_t)xorshift1024());
  }
  return fi;
}

bool search_5lut(const sbg_state st, const sbg_ttable target, const sbg_ttable mask,
    const int8_t *inbits, uint16_t *ret, int verbosity) {
  if (ret == NULL || st.num_gates < 5) abort(); /* lut.c:118-119 */
  const double t0 = now();
  sbg_handle *h = handle();

  uint8_t func_order[256]; /* lut.c:125-135: 256 draws, always */
  for (int i = 0; i < 256; i++) func_order[i] = (uint8_t)i;
  for (int i = 0; i < 256; i++) {
    const uint64_t j = xorshift1024() % (uint64_t)(i + 1);
    const uint8_t t = func_order[i];
    func_order[i] = func_order[j];
    func_order[j] = t;
  }
  memset(ret, 0, sizeof(uint16_t) * 10); /* lut.c:171 */

  load(h, &st, target, mask, inbits);
  sbg_result res;
  int rc = sbg_search5(h, func_order, &res);
  if (rc != SBG_OK) die("sbg_search5", rc);
  if (res.found) {
    ret[0] = res.func_outer;
    ret[1] = fill_dont_cares(&res);
    for (int i = 0; i < 5; i++) ret[2 + i] = res.gates[i];
    if (verbosity >= 1) { /* lut.c:219-222 */
      printf("[% 4d] Found 5LUT: %02x %02x    %3d %3d %3d %3d %3d\n", 0, ret[0], ret[1], ret[2],
          ret[3], ret[4], ret[5], ret[6]);
    }
  }
  g_calls[0]++;
  g_seconds[0] += now() - t0;
  g_kernel_ms[0] += sbg_last_kernel_ms(h, 0);
  return res.found != 0;
}

bool search_7lut(const sbg_state st, const sbg_ttable target, const sbg_ttable mask,
    const int8_t *inbits, uint16_t *ret, int verbosity) {
  if (ret == NULL || st.num_gates < 7) abort(); /* lut.c:258-259 */
  const double t0 = now();
  sbg_handle *h = handle();

  /* lut.c:362-378 draws these after phase 1; phase 1 draws nothing, so the stream is the same. */
  uint8_t outer_order[256], middle_order[256];
  for (int i = 0; i < 256; i++) outer_order[i] = middle_order[i] = (uint8_t)i;
  for (int i = 0; i < 256; i++) {
    const uint64_t oj = xorshift1024() % (uint64_t)(i + 1);
    const uint64_t mj = xorshift1024() % (uint64_t)(i + 1);
    const uint8_t ot = outer_order[i];
    const uint8_t mt = middle_order[i];
    outer_order[i] = outer_order[oj];
    middle_order[i] = middle_order[mj];
    outer_order[oj] = ot;
    middle_order[mj] = mt;
  }
  memset(ret, 0, sizeof(uint16_t) * 10); /* lut.c:383 */

  load(h, &st, target, mask, inbits);
  sbg_result res;
  int rc = sbg_search7(h, outer_order, middle_order, &res);
  if (rc != SBG_OK) die("sbg_search7", rc);
  if (res.found) {
    ret[0] = res.func_outer;
    ret[1] = res.func_middle;
    ret[2] = fill_dont_cares(&res);
    for (int i = 0; i < 7; i++) ret[3 + i] = res.gates[i];
    if (verbosity >= 1) { /* lut.c:470-473 */
      printf("[% 4d] Found 7LUT: %02x %02x %02x %3d %3d %3d %3d %3d %3d %3d\n", 0, ret[0], ret[1],
          ret[2], ret[3], ret[4], ret[5], ret[6], ret[7], ret[8], ret[9]);
    }
  }
  g_calls[1]++;
  g_seconds[1] += now() - t0;
  for (int i = 1; i < 4; i++) g_kernel_ms[i] += sbg_last_kernel_ms(h, i);
  return res.found != 0;
}

void sbg_shim_stats(uint64_t *calls5, uint64_t *calls7, double *seconds5, double *seconds7) {
  if (calls5 != NULL) *calls5 = g_calls[0];
  if (calls7 != NULL) *calls7 = g_calls[1];
  if (seconds5 != NULL) *seconds5 = g_seconds[0];
  if (seconds7 != NULL) *seconds7 = g_seconds[1];
}
