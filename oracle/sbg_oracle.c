/* oracle/sbg_oracle.c -- TEST INFRASTRUCTURE, not product code.  See sbg_oracle.h.
 *
 * Every function names the reference lines it restates.  The code is written for obviousness, not
 * speed: plain loops over positions and cells.  The one shortcut, orc_solve_inner(), is the
 * word-parallel closed form of get_lut_function and is itself checked against the bit-serial
 * restatement (orc_get_lut_function) and against the reference's object code in the tests.
 */
#include "sbg_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* RNG: xorshift1024* exactly as consumed at sboxgates.c:262-267.                                */

uint64_t orc_rng_next(orc_rng *rng) {
  const uint64_t s0 = rng->s[rng->p];
  rng->p = (rng->p + 1) & 15;
  uint64_t s1 = rng->s[rng->p];
  s1 ^= s1 << 31;
  rng->s[rng->p] = s1 ^ s0 ^ (s1 >> 11) ^ (s0 >> 30);
  rng->draws++;
  return rng->s[rng->p] * UINT64_C(1181783497276652981);
}

/* ------------------------------------------------------------------------------------------ */
/* Combinations in lexicographic order (lut.c:635-662, 743-770).                                */

int64_t orc_n_choose_k(int n, int k) {
  if (k < 0 || k > n) return 0;
  int64_t r = 1;
  for (int i = 1; i <= k; i++) {
    r *= (n - i + 1);
    r /= i;
  }
  return r;
}

/* lut.c:635-662: the rank-th t-subset of {0..n-1}, smallest element first. */
void orc_nth_combination(int64_t rank, int n, int t, uint16_t *out) {
  int next = 0;
  for (int pos = 0; pos < t; pos++) {
    for (;; next++) {
      const int64_t with_next = orc_n_choose_k(n - next - 1, t - pos - 1);
      if (rank < with_next) break;
      rank -= with_next;
    }
    out[pos] = (uint16_t)next++;
  }
}

int64_t orc_combination_rank(int n, int t, const uint16_t *comb) {
  int64_t rank = 0;
  int next = 0;
  for (int pos = 0; pos < t; pos++) {
    for (; next < comb[pos]; next++) {
      rank += orc_n_choose_k(n - next - 1, t - pos - 1);
    }
    next++;
  }
  return rank;
}

/* lut.c:743-758: successor; the last combination is left unchanged. */
void orc_next_combination(uint16_t *comb, int t, int n) {
  int i = t - 1;
  while (i >= 0 && comb[i] + t - i >= n) i--;
  if (i < 0) return;
  comb[i]++;
  for (int k = i + 1; k < t; k++) comb[k] = comb[k - 1] + 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Truth-table primitives.                                                                      */

static int tt_any(const uint64_t *a) { return (a[0] | a[1] | a[2] | a[3]) != 0; }

static int tt_bit(const uint64_t *a, int pos) { return (int)((a[pos >> 6] >> (pos & 63)) & 1); }

/* state.c:202-230: output bit p = bit (in1_p<<2 | in2_p<<1 | in3_p) of func. */
void orc_lut_ttable(uint8_t func, const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    uint64_t *out) {
  for (int v = 0; v < 4; v++) {
    uint64_t r = 0;
    for (int m = 0; m < 8; m++) {
      if (!((func >> m) & 1)) continue;
      const uint64_t a = (m & 4) ? in1[v] : ~in1[v];
      const uint64_t b = (m & 2) ? in2[v] : ~in2[v];
      const uint64_t c = (m & 1) ? in3[v] : ~in3[v];
      r |= a & b & c;
    }
    out[v] = r;
  }
}

/* lut.c:34-66.  The recursion there splits on tables[0] first and at each of the 2^num leaves
   rejects iff the leaf cell holds both a masked position with target 1 and one with target 0
   (lut.c:38-42); the final comparison at lut.c:65 can then never fail.  Here the leaves are visited
   directly. */
int orc_check_n_lut_possible(int num, const uint64_t *target, const uint64_t *mask,
    const uint64_t *tables) {
  for (int cell = 0; cell < (1 << num); cell++) {
    uint64_t ones[4], zeros[4];
    for (int v = 0; v < 4; v++) {
      uint64_t tt = mask[v];
      for (int i = 0; i < num; i++) {
        const uint64_t t = tables[4 * i + v];
        tt &= ((cell >> (num - 1 - i)) & 1) ? t : ~t;
      }
      ones[v] = tt & target[v];
      zeros[v] = tt & ~target[v];
    }
    if (tt_any(ones) && tt_any(zeros)) return 0;
  }
  return 1;
}

/* lut.c:79-109, bit-serial like the original: walk the masked positions, fix the LUT bit of the
   cell each one falls in, fail on the first contradiction; then fill the never-seen cells from one
   RNG draw (lut.c:104-106).  The visiting order does not influence the outcome. */
int orc_get_lut_function(const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    const uint64_t *target, const uint64_t *mask, int randomize, orc_rng *rng, uint8_t *func) {
  uint8_t f = 0, seen = 0;
  *func = 0;
  for (int pos = 0; pos < 256; pos++) {
    if (!tt_bit(mask, pos)) continue;
    const int cell = tt_bit(in1, pos) << 2 | tt_bit(in2, pos) << 1 | tt_bit(in3, pos);
    const int want = tt_bit(target, pos);
    if (!((seen >> cell) & 1)) {
      seen |= (uint8_t)(1 << cell);
      f |= (uint8_t)(want << cell);
    } else if (((f >> cell) & 1) != want) {
      return 0;
    }
  }
  if (randomize && seen != 0xff) {
    f |= (uint8_t)(~seen & (uint8_t)orc_rng_next(rng));
  }
  *func = f;
  return 1;
}

/* Closed form of the above without the random fill: per inner cell, func bit = "some masked
   position of the cell has target 1", seen bit = "the cell has a masked position"; conflict iff a
   cell has both a masked 1 and a masked 0. */
int orc_solve_inner(const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    const uint64_t *target, const uint64_t *mask, uint8_t *func, uint8_t *seen) {
  uint8_t f = 0, s = 0;
  for (int cell = 0; cell < 8; cell++) {
    uint64_t ones = 0, zeros = 0;
    for (int v = 0; v < 4; v++) {
      const uint64_t a = (cell & 4) ? in1[v] : ~in1[v];
      const uint64_t b = (cell & 2) ? in2[v] : ~in2[v];
      const uint64_t c = (cell & 1) ? in3[v] : ~in3[v];
      const uint64_t in_cell = a & b & c & mask[v];
      ones |= in_cell & target[v];
      zeros |= in_cell & ~target[v];
    }
    if (ones && zeros) return 0;
    if (ones) f |= (uint8_t)(1 << cell);
    if (ones || zeros) s |= (uint8_t)(1 << cell);
  }
  *func = f;
  *seen = s;
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Orderings.                                                                                   */

/* lut.c:189,224-229: outer inputs = k-th 3-subset of positions {0..4} in lexicographic order,
   inner inputs = the two remaining positions ascending. */
void orc_order5_row(int k, int *row5) {
  uint16_t outer[3] = {0, 1, 2};
  for (int i = 0; i < k; i++) orc_next_combination(outer, 3, 5);
  int used = 0, w = 3;
  for (int i = 0; i < 3; i++) {
    row5[i] = outer[i];
    used |= 1 << outer[i];
  }
  for (int i = 0; i < 5; i++) {
    if (!((used >> i) & 1)) row5[w++] = i;
  }
}

/* lut.c:396-415 is a literal 70-row table.  Its rule: outer = 3-subsets of
   {0..6} in lexicographic order; middle = 3-subsets of the remaining four, lexicographic; keep the
   row iff min(outer) < min(middle); the last entry is the leftover position.  The tests compare
   this generator with the literal table. */
void orc_order7_row(int k, int *row7) {
  int count = 0;
  for (int a = 0; a < 7; a++) for (int b = a + 1; b < 7; b++) for (int c = b + 1; c < 7; c++) {
    int rest[4], r = 0;
    for (int i = 0; i < 7; i++) {
      if (i != a && i != b && i != c) rest[r++] = i;
    }
    for (int skip = 3; skip >= 0; skip--) { /* leaving out rest[3] first = lexicographic middles */
      int mid[3], m = 0;
      for (int i = 0; i < 4; i++) {
        if (i != skip) mid[m++] = rest[i];
      }
      if (a >= mid[0]) continue;
      if (count == k) {
        row7[0] = a; row7[1] = b; row7[2] = c;
        row7[3] = mid[0]; row7[4] = mid[1]; row7[5] = mid[2];
        row7[6] = rest[skip];
        return;
      }
      count++;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */

static int rejected_by_inbits(const uint16_t *comb, int t, const int8_t *inbits) {
  for (int k = 0; k < 8 && inbits[k] != -1; k++) { /* lut.c:177-185, 297-305 */
    for (int m = 0; m < t; m++) {
      if (comb[m] == (uint16_t)inbits[k]) return 1;
    }
  }
  return 0;
}

static void shuffle_identity(uint8_t *perm) {
  for (int i = 0; i < 256; i++) perm[i] = (uint8_t)i;
}

/* lut.c:116-249 at size == 1. */
int orc_search_5lut(const uint64_t *tables, int n, const uint64_t *target, const uint64_t *mask,
    const int8_t *inbits, orc_rng *rng, uint16_t *ret, orc_stats *stats) {
  uint8_t func_order[256];
  shuffle_identity(func_order);
  for (int i = 0; i < 256; i++) { /* lut.c:130-135: 256 draws, always */
    const uint64_t j = orc_rng_next(rng) % (uint64_t)(i + 1);
    const uint8_t t = func_order[i];
    func_order[i] = func_order[j];
    func_order[j] = t;
  }
  memset(ret, 0, 10 * sizeof(uint16_t)); /* lut.c:171 */
  if (stats) memset(stats, 0, sizeof(*stats));

  int rows[10][5];
  for (int k = 0; k < 10; k++) orc_order5_row(k, rows[k]);

  const int64_t total = orc_n_choose_k(n, 5);
  uint16_t nums[5] = {0, 1, 2, 3, 4};
  for (int64_t r = 0; r < total; r++, orc_next_combination(nums, 5, n)) {
    if (stats) stats->tuples_filtered++;
    if (rejected_by_inbits(nums, 5, inbits)) continue;
    uint64_t tt[5 * 4];
    for (int m = 0; m < 5; m++) memcpy(tt + 4 * m, tables + 4 * nums[m], 32);
    if (!orc_check_n_lut_possible(5, target, mask, tt)) continue; /* lut.c:187 */
    if (stats) stats->tuples_feasible++;
    for (int k = 0; k < 10; k++) {
      const int *o = rows[k];
      for (int fo = 0; fo < 256; fo++) {
        if (stats) stats->candidates++;
        const uint8_t func_outer = func_order[fo];
        uint64_t t_outer[4];
        orc_lut_ttable(func_outer, tt + 4 * o[0], tt + 4 * o[1], tt + 4 * o[2], t_outer);
        uint8_t fi, seen;
        if (!orc_solve_inner(t_outer, tt + 4 * o[3], tt + 4 * o[4], target, mask, &fi, &seen)) {
          continue;
        }
        if (seen != 0xff) fi |= (uint8_t)(~seen & (uint8_t)orc_rng_next(rng)); /* lut.c:104-106 */
        ret[0] = func_outer;
        ret[1] = fi;
        for (int m = 0; m < 5; m++) ret[2 + m] = nums[o[m]];
        return 1;
      }
    }
  }
  return 0;
}

/* lut.c:290-327 at size == 1. */
int orc_filter_7lut(const uint64_t *tables, int n, const uint64_t *target, const uint64_t *mask,
    const int8_t *inbits, uint16_t *list, int cap, orc_stats *stats) {
  const int64_t total = orc_n_choose_k(n, 7);
  uint16_t nums[7] = {0, 1, 2, 3, 4, 5, 6};
  int count = 0;
  for (int64_t r = 0; r < total; r++, orc_next_combination(nums, 7, n)) {
    if (stats) stats->tuples_filtered++;
    if (!rejected_by_inbits(nums, 7, inbits)) {
      uint64_t tt[7 * 4];
      for (int m = 0; m < 7; m++) memcpy(tt + 4 * m, tables + 4 * nums[m], 32);
      if (orc_check_n_lut_possible(7, target, mask, tt)) {
        memcpy(list + 7 * count, nums, 7 * sizeof(uint16_t));
        count++;
      }
    }
    if (count >= cap) break; /* lut.c:316-318 */
  }
  if (stats) stats->tuples_feasible = (uint64_t)count;
  return count;
}

/* The reference keeps the 256 LUT outputs of the current outer and middle triple in two caches
   keyed by `int outer_cache_set` / `middle_cache_set` (lut.c:379-380,432-439).  The key it compares
   against is the 48-bit value a<<32|b<<16|c, so the stored int only ever holds b<<16|c, and the
   comparison succeeds -- no regeneration -- exactly when a == 0 and (b,c) equal the (b,c) of the
   last regeneration, whatever that regeneration's first gate was.  This struct reproduces that
   rule so that rows evaluated with a stale cache behave as in the reference. */
typedef struct {
  uint32_t low_key;  /* b<<16|c of the last regeneration; 0 initially (lut.c:379-380) */
  uint16_t gates[3]; /* the triple whose LUT outputs are actually in the cache */
} lut_cache;

static int cache_lookup(lut_cache *c, uint16_t a, uint16_t b, uint16_t cc) {
  const uint32_t low = (uint32_t)b << 16 | cc;
  const int hit = (a == 0) && (low == c->low_key);
  if (!hit) {
    c->low_key = low;
    c->gates[0] = a;
    c->gates[1] = b;
    c->gates[2] = cc;
    return 0;
  }
  return !(c->gates[0] == a && c->gates[1] == b && c->gates[2] == cc); /* 1 = stale */
}

/* lut.c:256-487 at size == 1. */
int orc_search_7lut(const uint64_t *tables, int n, const uint64_t *target, const uint64_t *mask,
    const int8_t *inbits, orc_rng *rng, uint16_t *ret, orc_stats *stats) {
  if (stats) memset(stats, 0, sizeof(*stats));
  uint16_t *list = malloc(sizeof(uint16_t) * 7 * ORC_LIST_CAP);
  if (list == NULL) abort();
  const int count = orc_filter_7lut(tables, n, target, mask, inbits, list, ORC_LIST_CAP, stats);

  uint8_t outer_order[256], middle_order[256];
  shuffle_identity(outer_order);
  shuffle_identity(middle_order);
  for (int i = 0; i < 256; i++) { /* lut.c:369-378: 512 interleaved draws, after phase 1 */
    const uint64_t oj = orc_rng_next(rng) % (uint64_t)(i + 1);
    const uint64_t mj = orc_rng_next(rng) % (uint64_t)(i + 1);
    const uint8_t ot = outer_order[i];
    const uint8_t mt = middle_order[i];
    outer_order[i] = outer_order[oj];
    middle_order[i] = middle_order[mj];
    outer_order[oj] = ot;
    middle_order[mj] = mt;
  }
  memset(ret, 0, 10 * sizeof(uint16_t)); /* lut.c:383 */

  int rows[70][7];
  for (int k = 0; k < 70; k++) orc_order7_row(k, rows[k]);

  lut_cache outer_cache = {0, {0, 0, 0}}, middle_cache = {0, {0, 0, 0}};
  uint64_t (*t_outer)[4] = malloc(256 * 32);
  uint64_t (*t_middle)[4] = malloc(256 * 32);
  if (t_outer == NULL || t_middle == NULL) abort();

  int found = 0;
  for (int i = 0; i < count && !found; i++) {
    const uint16_t *tuple = list + 7 * i;
    for (int k = 0; k < 70 && !found; k++) {
      uint16_t g[7];
      for (int m = 0; m < 7; m++) g[m] = tuple[rows[k][m]];
      const int stale_o = cache_lookup(&outer_cache, g[0], g[1], g[2]);
      const int stale_m = cache_lookup(&middle_cache, g[3], g[4], g[5]);
      if (stats && (stale_o || stale_m)) stats->stale_cache_rows++;
      const int row_is_stale = stale_o || stale_m;
      for (int f = 0; f < 256; f++) { /* lut.c:70-74, from the triple actually cached */
        orc_lut_ttable((uint8_t)f, tables + 4 * outer_cache.gates[0],
            tables + 4 * outer_cache.gates[1], tables + 4 * outer_cache.gates[2], t_outer[f]);
        orc_lut_ttable((uint8_t)f, tables + 4 * middle_cache.gates[0],
            tables + 4 * middle_cache.gates[1], tables + 4 * middle_cache.gates[2], t_middle[f]);
      }
      const uint64_t *tg = tables + 4 * g[6];
      for (int fo = 0; fo < 256 && !found; fo++) {
        const uint8_t func_outer = outer_order[fo];
        for (int fm = 0; fm < 256; fm++) {
          if (stats) stats->candidates++;
          const uint8_t func_middle = middle_order[fm];
          uint8_t fi, seen;
          if (!orc_solve_inner(t_outer[func_outer], t_middle[func_middle], tg, target, mask, &fi,
              &seen)) {
            continue;
          }
          if (seen != 0xff) fi |= (uint8_t)(~seen & (uint8_t)orc_rng_next(rng));
          ret[0] = func_outer; /* lut.c:453-462 */
          ret[1] = func_middle;
          ret[2] = fi;
          for (int m = 0; m < 7; m++) ret[3 + m] = g[m];
          if (stats) stats->stale_hit = (uint64_t)row_is_stale;
          found = 1;
          break;
        }
      }
    }
  }
  free(t_outer);
  free(t_middle);
  free(list);
  return found;
}

/* ------------------------------------------------------------------------------------------ */
/* "Minimum key over a share" forms (see sbg_oracle.h).  Same loops as above, restricted to a share
   and reporting where the first match sits instead of building ret[]. */

uint64_t orc_search5_key(const uint64_t *tables, int n, const uint64_t *target, const uint64_t *mask,
    const int8_t *inbits, const uint8_t *func_order, int part, int nparts) {
  int rows[10][5];
  for (int k = 0; k < 10; k++) orc_order5_row(k, rows[k]);
  const int64_t total = orc_n_choose_k(n, 5);
  uint16_t nums[5] = {0, 1, 2, 3, 4};
  for (int64_t r = 0; r < total; r++, orc_next_combination(nums, 5, n)) {
    if (r % nparts != part) continue;
    if (rejected_by_inbits(nums, 5, inbits)) continue;
    uint64_t tt[5 * 4];
    for (int m = 0; m < 5; m++) memcpy(tt + 4 * m, tables + 4 * nums[m], 32);
    if (!orc_check_n_lut_possible(5, target, mask, tt)) continue;
    for (int k = 0; k < 10; k++) {
      const int *o = rows[k];
      for (int pos = 0; pos < 256; pos++) {
        uint64_t t_outer[4];
        uint8_t fi, seen;
        orc_lut_ttable(func_order[pos], tt + 4 * o[0], tt + 4 * o[1], tt + 4 * o[2], t_outer);
        if (orc_solve_inner(t_outer, tt + 4 * o[3], tt + 4 * o[4], target, mask, &fi, &seen)) {
          return (uint64_t)r << 12 | (uint64_t)k << 8 | (uint64_t)pos;
        }
      }
    }
  }
  return UINT64_MAX;
}

uint64_t orc_decomp7_key(const uint64_t *tables, const uint64_t *target, const uint64_t *mask,
    const uint16_t *list, int count, const uint8_t *outer_order, const uint8_t *middle_order,
    int part, int nparts) {
  int rows[70][7];
  for (int k = 0; k < 70; k++) orc_order7_row(k, rows[k]);
  uint64_t (*t_outer)[4] = malloc(256 * 32);
  uint64_t (*t_middle)[4] = malloc(256 * 32);
  if (t_outer == NULL || t_middle == NULL) abort();
  uint64_t key = UINT64_MAX;
  for (int i = part; i < count && key == UINT64_MAX; i += nparts) {
    const uint16_t *tuple = list + 7 * i;
    /* State of the reference's outer cache on entry to tuple i: left by row 69 of tuple i-1
       (lut.c:432-435), wherever that tuple was processed. */
    lut_cache outer_cache = {0, {0, 0, 0}}, middle_cache = {0, {0, 0, 0}};
    if (i > 0) {
      const uint16_t *prev = list + 7 * (i - 1);
      cache_lookup(&outer_cache, prev[rows[69][0]], prev[rows[69][1]], prev[rows[69][2]]);
      cache_lookup(&middle_cache, prev[rows[69][3]], prev[rows[69][4]], prev[rows[69][5]]);
    }
    for (int k = 0; k < 70 && key == UINT64_MAX; k++) {
      uint16_t g[7];
      for (int m = 0; m < 7; m++) g[m] = tuple[rows[k][m]];
      cache_lookup(&outer_cache, g[0], g[1], g[2]);
      cache_lookup(&middle_cache, g[3], g[4], g[5]);
      for (int f = 0; f < 256; f++) {
        orc_lut_ttable((uint8_t)f, tables + 4 * outer_cache.gates[0],
            tables + 4 * outer_cache.gates[1], tables + 4 * outer_cache.gates[2], t_outer[f]);
        orc_lut_ttable((uint8_t)f, tables + 4 * middle_cache.gates[0],
            tables + 4 * middle_cache.gates[1], tables + 4 * middle_cache.gates[2], t_middle[f]);
      }
      for (int po = 0; po < 256 && key == UINT64_MAX; po++) {
        for (int pm = 0; pm < 256; pm++) {
          uint8_t fi, seen;
          if (orc_solve_inner(t_outer[outer_order[po]], t_middle[middle_order[pm]],
              tables + 4 * g[6], target, mask, &fi, &seen)) {
            key = (uint64_t)i << 23 | (uint64_t)k << 16 | (uint64_t)po << 8 | (uint64_t)pm;
            break;
          }
        }
      }
    }
  }
  free(t_outer);
  free(t_middle);
  return key;
}
