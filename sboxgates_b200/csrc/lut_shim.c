/* lut_shim.c -- see lut_shim.h.  Plain C, compiled by gcc; everything CUDA is behind
 * include/sboxgates_b200.h. */
#define _POSIX_C_SOURCE 200809L
#include "lut_shim.h"

#include <pthread.h>
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

/* Devices: SBG_GPUS=N (default 1) drives CUDA devices SBG_DEVICE .. SBG_DEVICE+N-1 from this one
   process.  Searches above the size thresholds are sharded over them (one host thread per device
   around the sbg_*_part calls, minimum key / concatenated hit lists merged here -- the in-process
   counterpart of the all-reduce(MIN) / all-gather that sboxgates_b200/distributed.py does over
   NCCL); smaller ones run on the first device only. */
#define SBG_SHIM_MAX_GPUS 8
static sbg_handle *g_handles[SBG_SHIM_MAX_GPUS];
static int g_ngpus = 0;
static double g_shard_min5 = 5e7, g_shard_min7 = 2e8;
static int g_shard_min_list = 8192;
static uint64_t g_sharded_calls = 0;
#define g_handle (g_handles[0])
static uint64_t g_calls[2] = {0, 0};
static double g_seconds[2] = {0.0, 0.0};
static double g_kernel_ms[4] = {0.0, 0.0, 0.0, 0.0}; /* search5, filter7, sort, decomp7 */

static double g_init_seconds = 0.0;   /* sbg_create: CUDA start-up + buffers, once */

static void shim_exit(void) {
  if (g_handle != NULL) {
    if (getenv("SBG_SHIM_STATS") != NULL) {
      fprintf(stderr, "[sbg] start-up (sbg_create) %.3f s, inside the first call; "
          "search_5lut: %llu calls %.3f s; search_7lut: %llu calls %.3f s; "
          "%llu kernel launches; kernel time: search5 %.3f s, filter7 %.3f s, sort %.3f s, "
          "decomp7 %.3f s\n", g_init_seconds, (unsigned long long)g_calls[0], g_seconds[0],
          (unsigned long long)g_calls[1], g_seconds[1],
          (unsigned long long)sbg_launch_count(g_handle), 1e-3 * g_kernel_ms[0],
          1e-3 * g_kernel_ms[1], 1e-3 * g_kernel_ms[2], 1e-3 * g_kernel_ms[3]);
    }
    if (g_ngpus > 1 && getenv("SBG_SHIM_STATS") != NULL) {
      fprintf(stderr, "[sbg] %d devices, %llu sharded search phases\n", g_ngpus,
          (unsigned long long)g_sharded_calls);
    }
    for (int i = 0; i < g_ngpus; i++) {
      sbg_destroy(g_handles[i]);
      g_handles[i] = NULL;
    }
  }
}

/* Errors are fatal, as in the reference, whose internal inconsistencies are assert()s
   (lut.c:118-119, 201, 452; sboxgates.h:31-44). */
static void die(const char *what, int rc) {
  fprintf(stderr, "sboxgates_b200: %s failed (%d): %s\n", what, rc,
      g_handle != NULL ? sbg_last_error(g_handle) : "no handle");
  abort();
}

static double now(void);

static sbg_handle *handle(void) {
  if (g_handle == NULL) {
    const double t_init = now();
    const char *dev = getenv("SBG_DEVICE");
    const char *gpus = getenv("SBG_GPUS");
    const int first = dev != NULL ? atoi(dev) : 0;
    int want = gpus != NULL ? atoi(gpus) : 1;
    if (want < 1) want = 1;
    if (want > SBG_SHIM_MAX_GPUS) want = SBG_SHIM_MAX_GPUS;
    for (int i = 0; i < want; i++) {
      int rc = sbg_create(&g_handles[i], first + i);
      if (rc != SBG_OK) {
        fprintf(stderr, "sboxgates_b200: sbg_create(device %d) failed (%d): %s\n", first + i, rc,
            g_handles[i] != NULL ? sbg_last_error(g_handles[i]) : "no handle");
        abort();
      }
      g_ngpus = i + 1;
    }
    if (getenv("SBG_SHARD_MIN5") != NULL) g_shard_min5 = atof(getenv("SBG_SHARD_MIN5"));
    if (getenv("SBG_SHARD_MIN7") != NULL) g_shard_min7 = atof(getenv("SBG_SHARD_MIN7"));
    if (getenv("SBG_SHARD_MIN_LIST") != NULL) g_shard_min_list = atoi(getenv("SBG_SHARD_MIN_LIST"));
    atexit(shim_exit);
    g_init_seconds = now() - t_init;
  }
  return g_handle;
}

static double n_choose(int n, int k) {
  double r = 1.0;
  for (int i = 1; i <= k; i++) r = r * (double)(n - i + 1) / (double)i;
  return r;
}

/* ---- one host thread per device for the sharded phases --------------------------------------- */

typedef struct {
  int part;
  int phase;                 /* 5: search5_part; 71: load + filter7_part; 72: set_list7 + decomp7_part */
  const uint64_t *flat;      /* problem (phases 5 and 71; device 0 is loaded by the caller) */
  int n;
  const uint64_t *target, *mask;
  const int8_t *inbits;
  const uint8_t *order_a, *order_b;
  uint64_t *list;            /* 71: out (SBG_LIST_CAP entries); 72: in (merged) */
  int count;                 /* 71: out; 72: in */
  uint64_t key;              /* 5, 72: out */
  int rc;
} shard_job;

static void *shard_main(void *arg) {
  shard_job *j = (shard_job *)arg;
  sbg_handle *h = g_handles[j->part];
  j->rc = SBG_OK;
  if (j->part != 0 && (j->phase == 5 || j->phase == 71)) {
    j->rc = sbg_load_problem(h, j->flat, j->n, j->target, j->mask, j->inbits);
    if (j->rc != SBG_OK) return NULL;
  }
  if (j->phase == 5) {
    j->rc = sbg_search5_part(h, j->part, g_ngpus, j->order_a, &j->key);
  } else if (j->phase == 71) {
    j->rc = sbg_filter7_part(h, j->part, g_ngpus, j->list, &j->count);
  } else {
    j->rc = sbg_set_list7(h, j->list, j->count);
    if (j->rc == SBG_OK) {
      j->rc = sbg_decomp7_part(h, j->part, g_ngpus, j->order_a, j->order_b, &j->key);
    }
  }
  return NULL;
}

static void run_shards(shard_job *jobs) {
  pthread_t tid[SBG_SHIM_MAX_GPUS];
  for (int i = 1; i < g_ngpus; i++) {
    if (pthread_create(&tid[i], NULL, shard_main, &jobs[i]) != 0) abort();
  }
  shard_main(&jobs[0]);
  for (int i = 1; i < g_ngpus; i++) pthread_join(tid[i], NULL);
  for (int i = 0; i < g_ngpus; i++) {
    if (jobs[i].rc != SBG_OK) {
      fprintf(stderr, "sboxgates_b200: sharded phase %d failed on device %d (%d): %s\n",
          jobs[i].phase, i, jobs[i].rc, sbg_last_error(g_handles[i]));
      abort();
    }
  }
  g_sharded_calls++;
}

static int cmp_u64(const void *a, const void *b) {
  const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : x > y;
}

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static uint64_t g_flat[SBG_SHIM_MAX_GATES * 4];
static uint64_t g_t[4], g_m[4];

static void load(sbg_handle *h, const sbg_state *st, const sbg_ttable target, const sbg_ttable mask,
    const int8_t *inbits) {
  uint64_t *flat = g_flat, *t = g_t, *m = g_m;
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
  int rc;
  if (g_ngpus > 1 && n_choose(st.num_gates, 5) >= g_shard_min5) {
    shard_job jobs[SBG_SHIM_MAX_GPUS];
    uint64_t key = SBG_KEY_NONE;
    for (int i = 0; i < g_ngpus; i++) {
      jobs[i] = (shard_job){.part = i, .phase = 5, .flat = g_flat, .n = st.num_gates,
          .target = g_t, .mask = g_m, .inbits = inbits, .order_a = func_order};
    }
    run_shards(jobs);
    for (int i = 0; i < g_ngpus; i++) key = jobs[i].key < key ? jobs[i].key : key;
    rc = sbg_finish5(h, key, func_order, &res);
  } else {
    rc = sbg_search5(h, func_order, &res);
  }
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
  int rc;
  if (g_ngpus > 1 && n_choose(st.num_gates, 7) >= g_shard_min7) {
    /* phase 1 sharded: per-device sorted lists -> merged, sorted, cut at SBG_LIST_CAP */
    static uint64_t *lists = NULL;
    if (lists == NULL) lists = malloc(sizeof(uint64_t) * SBG_LIST_CAP * SBG_SHIM_MAX_GPUS);
    if (lists == NULL) abort();
    shard_job jobs[SBG_SHIM_MAX_GPUS];
    for (int i = 0; i < g_ngpus; i++) {
      jobs[i] = (shard_job){.part = i, .phase = 71, .flat = g_flat, .n = st.num_gates,
          .target = g_t, .mask = g_m, .inbits = inbits,
          .list = lists + (size_t)i * SBG_LIST_CAP};
    }
    run_shards(jobs);
    int total = jobs[0].count;
    for (int i = 1; i < g_ngpus; i++) {
      memmove(lists + total, lists + (size_t)i * SBG_LIST_CAP, sizeof(uint64_t) * jobs[i].count);
      total += jobs[i].count;
    }
    qsort(lists, (size_t)total, sizeof(uint64_t), cmp_u64);
    if (total > SBG_LIST_CAP) total = SBG_LIST_CAP;
    uint64_t key = SBG_KEY_NONE;
    if (total >= g_shard_min_list) {
      for (int i = 0; i < g_ngpus; i++) {
        jobs[i] = (shard_job){.part = i, .phase = 72, .order_a = outer_order,
            .order_b = middle_order, .list = lists, .count = total};
      }
      run_shards(jobs);
      for (int i = 0; i < g_ngpus; i++) key = jobs[i].key < key ? jobs[i].key : key;
    } else {
      rc = sbg_set_list7(h, lists, total);
      if (rc != SBG_OK) die("sbg_set_list7", rc);
      rc = sbg_decomp7_part(h, 0, 1, outer_order, middle_order, &key);
      if (rc != SBG_OK) die("sbg_decomp7_part", rc);
    }
    rc = sbg_finish7(h, key, outer_order, middle_order, &res);
  } else {
    rc = sbg_search7(h, outer_order, middle_order, &res);
  }
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
