/* lut_shim.h -- the reference-facing boundary: search_5lut / search_7lut with the exact
 * signatures of the reference's lut.h:46-55, so an unmodified sboxgates host (sboxgates.c,
 * state.c, the rest of lut.c) links against libsbg_lutshim.a + libsboxgates_b200.so instead of its
 * own two functions.
 *
 * The types below are layout twins of state.h:64-88, re-declared here so that this translation
 * unit does not need the reference's headers (and nvcc never sees a GCC vector type).  The layout
 * is asserted at compile time in lut_shim.c: ttable 32 B / 32-aligned, gate 64 B, gates[] at
 * offset 32, state 32,032 B.  Compile with the same -m flags as the host objects: `ttable` is
 * passed in a YMM register only when AVX is enabled.
 */
#ifndef SBG_LUT_SHIM_H
#define SBG_LUT_SHIM_H

#include <stdbool.h>
#include <stdint.h>

#define SBG_SHIM_MAX_GATES 500

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

/* lut.h:46-47 */
bool search_5lut(const sbg_state st, const sbg_ttable target, const sbg_ttable mask,
    const int8_t *inbits, uint16_t *ret, int verbosity);
/* lut.h:54-55 */
bool search_7lut(const sbg_state st, const sbg_ttable target, const sbg_ttable mask,
    const int8_t *inbits, uint16_t *ret, int verbosity);

/* Supplied by the host program (sboxgates.h:113, sboxgates.c:246-268): the shim must draw from the
   host's generator so that the host's later shuffles are unchanged. */
uint64_t xorshift1024(void);

/* Optional: totals over the life of the process (calls, seconds inside the two functions). */
void sbg_shim_stats(uint64_t *calls5, uint64_t *calls7, double *seconds5, double *seconds7);

#endif
