/* lut_shim.h -- the reference-facing boundary: the LUT search functions with the exact signatures
 * of the reference's lut.h, so an unmodified sboxgates host (sboxgates.c, state.c, the rest of
 * lut.c) links against this shim + libsboxgates_b200.so instead of its own definitions.
 *
 * Two granularities, chosen at link time:
 *   libsbg_lutshim.a       search_5lut / search_7lut (lut.h:46-55): the reference's own lut_search
 *                          keeps calling them, one device round trip per call;
 *   libsbg_lutshim_node.a  additionally lut_search (lut.h:57-58, lut.c:489-631): the 3-LUT scan,
 *                          search_5lut and search_7lut of a node become ONE device call chain.
 *                          Because the shuffles of the later stages have to be known before the
 *                          earlier stages' outcomes are, this variant also interposes
 *                          xorshift1024() with a look-ahead buffer in front of the HOST's own
 *                          generator (which it calls under the name sbg_host_xorshift1024): the
 *                          sequence of values every caller sees is unchanged.
 *
 * The types below are layout twins of state.h:64-88, boolfunc.h:28-40 and sboxgates.h:49-66,
 * re-declared here so that this translation unit does not need the reference's headers (and nvcc
 * never sees a GCC vector type).  The layout is asserted at compile time in lut_shim.c and checked
 * against the reference's object code by tests/test_oracle_ref.py.  Compile with the same -m flags
 * as the host objects: `ttable` is passed in a YMM register only when AVX is enabled
 * (SURVEY.md section 8b).
 */
#ifndef SBG_LUT_SHIM_H
#define SBG_LUT_SHIM_H

#include <stdbool.h>
#include <stdint.h>

#define SBG_SHIM_MAX_GATES 500
#define SBG_SHIM_NO_GATE ((uint16_t)-1) /* state.h:30 */

typedef uint64_t sbg_ttable __attribute__((aligned(32))) __attribute__((vector_size(32)));

typedef struct {
  sbg_ttable table;
  int32_t type;      /* gate_type enum */
  uint16_t in1;
  uint16_t in2;
  uint16_t in3;
  uint8_t function;
} sbg_gate;

typedef struct {
  int32_t max_sat_metric;
  int32_t sat_metric;
  uint16_t max_gates;
  uint16_t num_gates;
  uint16_t outputs[8];
  sbg_gate gates[SBG_SHIM_MAX_GATES];
} sbg_state;

/* boolfunc.h:28-40 */
typedef struct {
  int32_t num_inputs;
  uint8_t fun;
  int32_t fun1, fun2;
  bool not_a, not_b, not_c, not_out, ab_commutative, ac_commutative, bc_commutative;
} sbg_boolfunc;

/* sboxgates.h:49-66 */
typedef struct {
  char fname[1000];
  char gfname[1000];
  int32_t iterations;
  int32_t oneoutput;
  int32_t permute;
  int32_t metric;
  bool output_c, output_dot, lut_graph, randomize, try_nots;
  sbg_boolfunc avail_gates[17];
  sbg_boolfunc avail_not[49];
  sbg_boolfunc avail_3[256];
  int32_t num_avail_3;
  int32_t verbosity;
} sbg_options;

/* lut.h:46-47 */
bool search_5lut(const sbg_state st, const sbg_ttable target, const sbg_ttable mask,
    const int8_t *inbits, uint16_t *ret, int verbosity);
/* lut.h:54-55 */
bool search_7lut(const sbg_state st, const sbg_ttable target, const sbg_ttable mask,
    const int8_t *inbits, uint16_t *ret, int verbosity);
#ifdef SBG_SHIM_NODE
/* lut.h:57-58 */
uint16_t lut_search(sbg_state *st, const sbg_ttable target, const sbg_ttable mask,
    const int8_t *inbits, const uint16_t *gate_order, const sbg_options *opt);
/* The host's generator under its link-time alias (see the header comment). */
uint64_t sbg_host_xorshift1024(void);
#endif

/* Supplied by the host program (sboxgates.h:113, sboxgates.c:246-268): the shim must draw from the
   host's generator so that the host's later shuffles are unchanged.  (Node variant: defined by
   the shim itself as the look-ahead front of sbg_host_xorshift1024.) */
uint64_t xorshift1024(void);

/* Optional: totals over the life of the process (calls, seconds inside the search functions). */
void sbg_shim_stats(uint64_t *calls5, uint64_t *calls7, double *seconds5, double *seconds7);

#endif
