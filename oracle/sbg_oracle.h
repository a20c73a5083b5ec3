/* oracle/sbg_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement of the reference's 3-LUT search path at MPI size == 1 (lut.c, state.c:202-230,
 * sboxgates.c:246-268).  It is the checker the GPU path is compared against.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it; the product
 * (sboxgates_b200/) never does.
 *
 * Pinning: tests/test_oracle_golden.py replays tests/golden/ (fixtures produced by the reference's
 * own object code, oracle/_ref/libsbgref.so and oracle/_ref/sboxgates_rec) through every function
 * below; tests/test_oracle_ref.py does the same live when oracle/_ref/ exists.
 *
 * Truth tables are 4 x uint64_t, lane v / bit b = S-box input 64*v + b (state.h:64-68,
 * state.c:232-250).
 */
#ifndef SBG_ORACLE_H
#define SBG_ORACLE_H

#include <stdint.h>

#define ORC_MAX_GATES 500
#define ORC_LIST_CAP 100000 /* lut.c:291,316-318 */

typedef struct {
  uint64_t s[16];
  int32_t p;
  uint64_t draws;
} orc_rng;

/* Work counters, in the units SURVEY.md section 8d defines. */
typedef struct {
  uint64_t tuples_filtered;  /* T-units: combinations put through inbits-reject + feasibility */
  uint64_t tuples_feasible;  /* feasible combinations (5-LUT: tried; 7-LUT: length of the list) */
  uint64_t candidates;       /* C-units: (tuple, ordering, fo[, fm]) candidates decided */
  uint64_t stale_cache_rows; /* 7-LUT rows evaluated with the reference's stale outer cache */
  uint64_t stale_hit;        /* 1 if the returned match came from such a row */
} orc_stats;

uint64_t orc_rng_next(orc_rng *rng);
int64_t orc_n_choose_k(int n, int k);
void orc_nth_combination(int64_t rank, int n, int t, uint16_t *out);
int64_t orc_combination_rank(int n, int t, const uint16_t *comb);
void orc_next_combination(uint16_t *comb, int t, int n);

void orc_lut_ttable(uint8_t func, const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    uint64_t *out);
int orc_check_n_lut_possible(int num, const uint64_t *target, const uint64_t *mask,
    const uint64_t *tables);
int orc_get_lut_function(const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    const uint64_t *target, const uint64_t *mask, int randomize, orc_rng *rng, uint8_t *func);
int orc_solve_inner(const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    const uint64_t *target, const uint64_t *mask, uint8_t *func, uint8_t *seen);

int orc_search_5lut(const uint64_t *tables, int n, const uint64_t *target, const uint64_t *mask,
    const int8_t *inbits, orc_rng *rng, uint16_t *ret, orc_stats *stats);
int orc_search_7lut(const uint64_t *tables, int n, const uint64_t *target, const uint64_t *mask,
    const int8_t *inbits, orc_rng *rng, uint16_t *ret, orc_stats *stats);

/* Phase 1 of search_7lut alone (lut.c:290-327): writes up to cap feasible 7-combinations
   (7 x uint16 each) in lexicographic order, returns how many. */
int orc_filter_7lut(const uint64_t *tables, int n, const uint64_t *target, const uint64_t *mask,
    const int8_t *inbits, uint16_t *list, int cap, orc_stats *stats);

/* The same searches expressed as "minimum key over a share of the work", which is how the sharded
   implementations (GPU parts, ranks) are specified.  part/nparts select the combinations (5-LUT)
   or list entries (7-LUT) whose rank / index is congruent to part modulo nparts; the orders are the
   already shuffled function orders; no RNG is involved.  Keys: 5-LUT rank<<12 | k<<8 | pos;
   7-LUT idx<<23 | k<<16 | pos_outer<<8 | pos_middle; UINT64_MAX if the share holds no match. */
uint64_t orc_search5_key(const uint64_t *tables, int n, const uint64_t *target, const uint64_t *mask,
    const int8_t *inbits, const uint8_t *func_order, int part, int nparts);
uint64_t orc_decomp7_key(const uint64_t *tables, const uint64_t *target, const uint64_t *mask,
    const uint16_t *list, int count, const uint8_t *outer_order, const uint8_t *middle_order,
    int part, int nparts);

/* Row k (0..69) of the ordering table at lut.c:396-415, regenerated from its rule. */
void orc_order7_row(int k, int *row7);
/* Ordering k (0..9) of search_5lut (lut.c:189,224-229). */
void orc_order5_row(int k, int *row5);

#endif
