/* oracle/ref_glue.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Glue that is linked next to the UNMODIFIED reference objects (compiled in place from
 * /root/reference by oracle/Makefile, outputs only under oracle/_ref/).  It never re-implements
 * any search logic; it only
 *   - injects a fixed RNG seed (the reference seeds xorshift1024 from /dev/urandom,
 *     sboxgates.c:246-260),
 *   - exposes the reference's by-value/vector-typed functions (lut.h:28-58) through plain
 *     pointer signatures that ctypes can call, and
 *   - records every search_5lut/search_7lut call of a real run as a replayable fixture.
 *
 * Sections are selected with -D flags so that each _ref target links only what it needs:
 *   SBGREF_WRAP_FOPEN  __wrap_fopen: "/dev/urandom" -> $SBG_SEEDFILE            (sboxgates_ref)
 *   SBGREF_RNG         strong xorshift1024 with inspectable state (the reference's own definition
 *                      is weakened with objcopy); same public xorshift1024* algorithm, and
 *                      tests/test_oracle_ref.py checks both give the same graph for the same seed
 *   SBGREF_API         pointer-based entry points                                (libsbgref.so)
 *   SBGREF_RECORDER    strong search_5lut/search_7lut that forward to the renamed reference
 *                      functions and append a record to $SBG_RECORD              (sboxgates_rec)
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

int sbgref_fake_rank = 0;
int sbgref_fake_size = 1;

void sbgref_set_fake_rank(int rank, int size) {
  sbgref_fake_rank = rank;
  sbgref_fake_size = size;
}

#ifdef SBGREF_WRAP_FOPEN
FILE *__real_fopen(const char *path, const char *mode);
FILE *__wrap_fopen(const char *path, const char *mode) {
  if (strcmp(path, "/dev/urandom") == 0) {
    const char *seed = getenv("SBG_SEEDFILE");
    if (seed != NULL && seed[0] != '\0') {
      return __real_fopen(seed, mode);
    }
  }
  return __real_fopen(path, mode);
}
#endif

#ifdef SBGREF_RNG
#include <stdint.h>
/* xorshift1024* (Vigna 2014), the generator sboxgates.c:246-268 uses.  State is exposed so a test
   can start a reference search from any RNG state and count the draws it made. */
static uint64_t rng_s[16];
static int rng_p = 0;
static int rng_ready = 0;
static uint64_t rng_draws = 0;

void sbgref_rng_set(const uint64_t *s16, int p) {
  memcpy(rng_s, s16, sizeof(rng_s));
  rng_p = p & 15;
  rng_ready = 1;
  rng_draws = 0;
}

void sbgref_rng_get(uint64_t *s16, int *p, uint64_t *draws) {
  memcpy(s16, rng_s, sizeof(rng_s));
  *p = rng_p;
  *draws = rng_draws;
}

uint64_t xorshift1024(void) {
  if (!rng_ready) {
    const char *seed = getenv("SBG_SEEDFILE");
    FILE *fp = fopen(seed != NULL && seed[0] != '\0' ? seed : "/dev/urandom", "r");
    if (fp == NULL || fread(rng_s, sizeof(rng_s), 1, fp) != 1) {
      fprintf(stderr, "ref_glue: cannot read RNG seed\n");
      abort();
    }
    fclose(fp);
    rng_ready = 1;
  }
  rng_draws++;
  const uint64_t a = rng_s[rng_p];
  rng_p = (rng_p + 1) & 15;
  uint64_t b = rng_s[rng_p];
  b ^= b << 31;
  rng_s[rng_p] = b ^ a ^ (b >> 11) ^ (a >> 30);
  return rng_s[rng_p] * UINT64_C(1181783497276652981);
}
#endif /* SBGREF_RNG */

#if defined(SBGREF_API) || defined(SBGREF_RECORDER)
#include "lut.h" /* the reference's own header, from -I/root/reference */

static void fill_state(state *st, const uint64_t *tables, int n) {
  memset(st, 0, sizeof(*st));
  st->max_gates = MAX_GATES;
  st->num_gates = (gatenum)n;
  for (int i = 0; i < 8; i++) {
    st->outputs[i] = NO_GATE;
  }
  for (int i = 0; i < n; i++) {
    memcpy(&st->gates[i].table, tables + 4 * i, 32);
    st->gates[i].type = LUT;
    st->gates[i].in1 = st->gates[i].in2 = st->gates[i].in3 = NO_GATE;
  }
}

static ttable load_tt(const uint64_t *w) {
  ttable t;
  memcpy(&t, w, 32);
  return t;
}
#endif

#ifdef SBGREF_API
int sbgref_sizes(int *sz_ttable, int *sz_gate, int *sz_state, int *off_gates) {
  *sz_ttable = (int)sizeof(ttable);
  *sz_gate = (int)sizeof(gate);
  *sz_state = (int)sizeof(state);
  *off_gates = (int)__builtin_offsetof(state, gates);
  return 0;
}

/* sboxgates.h:49-66, boolfunc.h:28-40: what the node-level shim assumes about `options`. */
int sbgref_options_layout(int *off_randomize, int *off_lut_graph, int *off_verbosity, int *sz_options,
    int *sz_boolfunc) {
  *off_randomize = (int)__builtin_offsetof(options, randomize);
  *off_lut_graph = (int)__builtin_offsetof(options, lut_graph);
  *off_verbosity = (int)__builtin_offsetof(options, verbosity);
  *sz_options = (int)sizeof(options);
  *sz_boolfunc = (int)sizeof(boolfunc);
  return 0;
}

int sbgref_check_n_lut_possible(int num, const uint64_t *target, const uint64_t *mask,
    const uint64_t *tables /* num x 4 */) {
  ttable tt[7];
  for (int i = 0; i < num; i++) {
    tt[i] = load_tt(tables + 4 * i);
  }
  return check_n_lut_possible(num, load_tt(target), load_tt(mask), tt) ? 1 : 0;
}

int sbgref_get_lut_function(const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    const uint64_t *target, const uint64_t *mask, int randomize, uint8_t *func) {
  return get_lut_function(load_tt(in1), load_tt(in2), load_tt(in3), load_tt(target), load_tt(mask),
      randomize != 0, func) ? 1 : 0;
}

void sbgref_generate_lut_ttable(int func, const uint64_t *in1, const uint64_t *in2,
    const uint64_t *in3, uint64_t *out) {
  ttable t = generate_lut_ttable((uint8_t)func, load_tt(in1), load_tt(in2), load_tt(in3));
  memcpy(out, &t, 32);
}

/* lut.h:46-47.  ret must hold 10 entries. */
int sbgref_search_5lut(const uint64_t *tables, int n, const uint64_t *target, const uint64_t *mask,
    const int8_t *inbits, uint16_t *ret) {
  static state st; /* 32 KB: keep it off ctypes' stack */
  fill_state(&st, tables, n);
  return search_5lut(st, load_tt(target), load_tt(mask), inbits, ret, 0) ? 1 : 0;
}

/* lut.h:54-55. */
int sbgref_search_7lut(const uint64_t *tables, int n, const uint64_t *target, const uint64_t *mask,
    const int8_t *inbits, uint16_t *ret) {
  static state st;
  fill_state(&st, tables, n);
  return search_7lut(st, load_tt(target), load_tt(mask), inbits, ret, 0) ? 1 : 0;
}
#endif /* SBGREF_API */

#ifdef SBGREF_RECORDER
/* Built against lut.c compiled with -Dsearch_5lut=ref_search_5lut -Dsearch_7lut=ref_search_7lut
   (genuine reference code under another name); a second, normally named copy of lut.c has these two
   symbols weakened so that lut_search (lut.c:553,593) binds to the strong definitions below. */
bool ref_search_5lut(const state st, const ttable target, const ttable mask, const int8_t *inbits,
    uint16_t *ret, int verbosity);
bool ref_search_7lut(const state st, const ttable target, const ttable mask, const int8_t *inbits,
    uint16_t *ret, int verbosity);

void sbgref_rng_get(uint64_t *s16, int *p, uint64_t *draws);

static FILE *rec_fp = NULL;
static long rec_limit = -1;
static long rec_count = 0;

static void rec_open(void) {
  if (rec_fp != NULL) return;
  const char *path = getenv("SBG_RECORD");
  if (path == NULL || path[0] == '\0') return;
  rec_fp = fopen(path, "wb");
  const char *lim = getenv("SBG_RECORD_LIMIT");
  if (lim != NULL) rec_limit = atol(lim);
}

/* Record layout (little endian):
     u32 magic ("SBG5"/"SBG7")  u32 n
     u64 tables[n][4]  u64 target[4]  u64 mask[4]  i8 inbits[8]
     u64 rng_state[16]  u32 rng_p            (state BEFORE the call)
     u32 found  u16 ret[10]  u64 draws       (draws = xorshift1024 calls made by the search)
     u64 nanoseconds                          (reference wall time of the call)            */
static bool record_call(int which, const state st, const ttable target, const ttable mask,
    const int8_t *inbits, uint16_t *ret, int verbosity) {
  rec_open();
  uint64_t s0[16], s1[16], d0, d1;
  int p0, p1;
  sbgref_rng_get(s0, &p0, &d0);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  bool found = which == 5 ? ref_search_5lut(st, target, mask, inbits, ret, verbosity)
                          : ref_search_7lut(st, target, mask, inbits, ret, verbosity);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  sbgref_rng_get(s1, &p1, &d1);
  if (rec_fp != NULL && (rec_limit < 0 || rec_count < rec_limit)) {
    uint32_t magic = which == 5 ? 0x35474253u : 0x37474253u;
    uint32_t n = st.num_gates;
    fwrite(&magic, 4, 1, rec_fp);
    fwrite(&n, 4, 1, rec_fp);
    for (uint32_t i = 0; i < n; i++) {
      fwrite(&st.gates[i].table, 32, 1, rec_fp);
    }
    fwrite(&target, 32, 1, rec_fp);
    fwrite(&mask, 32, 1, rec_fp);
    fwrite(inbits, 1, 8, rec_fp);
    fwrite(s0, 8, 16, rec_fp);
    uint32_t p = (uint32_t)p0;
    fwrite(&p, 4, 1, rec_fp);
    uint32_t f = found ? 1 : 0;
    fwrite(&f, 4, 1, rec_fp);
    fwrite(ret, 2, 10, rec_fp);
    uint64_t draws = d1 - d0;
    fwrite(&draws, 8, 1, rec_fp);
    uint64_t ns = (uint64_t)(t1.tv_sec - t0.tv_sec) * 1000000000ull
        + (uint64_t)(t1.tv_nsec - t0.tv_nsec);
    fwrite(&ns, 8, 1, rec_fp);
    fflush(rec_fp);
    rec_count++;
  }
  if (rec_limit >= 0 && rec_count >= rec_limit && getenv("SBG_RECORD_EXIT") != NULL) {
    exit(0);
  }
  return found;
}

bool search_5lut(const state st, const ttable target, const ttable mask, const int8_t *inbits,
    uint16_t *ret, int verbosity) {
  return record_call(5, st, target, mask, inbits, ret, verbosity);
}

bool search_7lut(const state st, const ttable target, const ttable mask, const int8_t *inbits,
    uint16_t *ret, int verbosity) {
  return record_call(7, st, target, mask, inbits, ret, verbosity);
}
#endif /* SBGREF_RECORDER */
