/* sboxgates_b200.h -- C ABI of the B200-native 3-LUT exhaustive search.
 *
 * This is the drop-in boundary for the `--lut` path of dansarie/sboxgates: plain pointers and
 * sizes, no torch / C++ / vector types.  Each entry point names the reference interface it
 * replaces (paths are relative to the reference checkout).
 *
 *   reference                                      here
 *   ---------------------------------------------  ----------------------------------------------
 *   search_5lut(st,target,mask,inbits,ret,v)       sbg_search5()         (lut.h:46-47, lut.c:116-249)
 *   search_7lut(st,target,mask,inbits,ret,v)       sbg_search7()         (lut.h:54-55, lut.c:256-487)
 *   MPI_Bcast(&work,...) of the search state       sbg_load_problem()    (lut.c:533-540,
 *                                                                         sboxgates.c:625-627)
 *   the per-rank slice of C(n,5) / C(n,7)          sbg_search5_part(), sbg_filter7_part()
 *                                                                        (lut.c:137-149, 265-277)
 *   MPI_Allgather(v) of the 7-LUT hit lists        sbg_set_list7()       (lut.c:329-349)
 *   the per-rank slice of the hit list             sbg_decomp7_part()    (lut.c:351-360, 416-484)
 *   get_search_result(): first finder wins         min over ranks of the 64-bit keys the *_part
 *                                                  calls return (one all-reduce(MIN) per phase),
 *                                                  then sbg_finish5()/sbg_finish7()  (lut.c:665-740)
 *
 * Semantics are those of the reference at MPI size == 1: the result is the FIRST match in the
 * reference's enumeration order (combination in lexicographic order, then ordering, then position
 * in the shuffled function order(s)), whatever the number of GPUs.  The shuffled function orders
 * are inputs because the caller owns the RNG (lut.c:125-135, 362-378 consume the host's
 * xorshift1024); likewise the random fill of don't-care LUT bits (lut.c:104-106) is left to the
 * caller: results carry the solved inner function and its `seen` mask.
 *
 * Truth tables are 4 x uint64_t per gate, gate-major: bit b of word v = value at S-box input
 * 64*v+b -- the memory image of the reference's `ttable` (state.h:64-68).
 *
 * All functions return SBG_OK (0) or a negative error code; sbg_last_error() gives the text.
 * A handle is bound to one CUDA device and is not thread-safe (the reference's functions are not
 * re-entrant either, lut.h / SURVEY.md section 8b).
 */
#ifndef SBOXGATES_B200_H
#define SBOXGATES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBG_OK 0
#define SBG_ERR_ARG (-1)       /* bad argument (n out of range, null pointer, ...) */
#define SBG_ERR_CUDA (-2)      /* a CUDA call failed; see sbg_last_error() */
#define SBG_ERR_OVERFLOW (-3)  /* hit buffer too small even for the serial retry */
#define SBG_ERR_STATE (-4)     /* call sequence error (e.g. no problem loaded) */

#define SBG_MAX_GATES 500      /* state.h:26 */
#define SBG_LIST_CAP 100000    /* lut.c:291,316-318: at most this many feasible 7-tuples are tried */
#define SBG_KEY_NONE UINT64_MAX
#define SBG_PROBLEM_SLOTS 64    /* device-resident search states per handle */

typedef struct sbg_handle sbg_handle;

/* What a search returns; the caller turns it into the reference's ret[10] (lut.c:202-211,
   453-462) after applying the random don't-care fill. */
typedef struct {
  int32_t found;
  int32_t ordering;        /* k: 0..9 (lut.c:189-229) or 0..69 (lut.c:396-415) */
  int32_t pos_outer;       /* position of func_outer in the shuffled order */
  int32_t pos_middle;      /* 7-LUT only */
  uint8_t func_outer;
  uint8_t func_middle;     /* 7-LUT only */
  uint8_t func_inner;      /* solved bits only; don't-care bits are 0 */
  uint8_t inner_seen;      /* bit c set = inner cell c occurs under the mask (0xff: no fill) */
  uint16_t gates[7];       /* LUT inputs in reference order: a,b,c,d,e[,f,g] */
  uint16_t stale_outer;    /* 7-LUT: 1 if the reference would have evaluated this row with its
                              stale outer cache (lut.c:432-435); func_inner then follows suit */
  uint64_t index;          /* 5-LUT: lexicographic rank of the combination; 7-LUT: list index */
  uint64_t key;            /* the packed minimum key (SBG_KEY_NONE if nothing matched) */
  uint64_t tuples_feasible;/* 5-LUT: feasible combinations met; 7-LUT: length of the hit list */
  uint64_t tuples_swept;   /* combinations put through the feasibility test by this device */
} sbg_result;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int sbg_create(sbg_handle **out, int device);
void sbg_destroy(sbg_handle *h);
const char *sbg_last_error(const sbg_handle *h);
/* Run on an externally owned stream (a cudaStream_t, e.g. torch's current stream); NULL restores
   the handle's own stream. */
int sbg_set_stream(sbg_handle *h, void *cuda_stream);
/* Number of this library's own kernels the handle has launched so far (bench.py reports it;
   CUB's radix-sort kernels are not counted). */
uint64_t sbg_launch_count(const sbg_handle *h);
/* CUDA-event time (ms) spent in the named kernel family by the last search call:
   0 = search5, 1 = filter7, 2 = sort, 3 = decomp7. */
float sbg_last_kernel_ms(const sbg_handle *h, int which);

/* ---- problem -------------------------------------------------------------------------------- */
/* Uploads one search state: n gate tables, target, mask, and the -1 terminated list of input-bit
   gates already used as multiplexer selectors (lut.c:177-185).  Host pointers. */
int sbg_load_problem(sbg_handle *h, const uint64_t *tables, int n, const uint64_t *target,
    const uint64_t *mask, const int8_t *inbits);
/* The same in two steps, for callers that keep several search states resident in HBM: stage a
   state into slot 0..SBG_PROBLEM_SLOTS-1 (upload), later make a staged slot the current problem
   (no transfer).  sbg_load_problem() = stage + use of slot 0.  This is the device-resident
   replacement of the `state` array the reference re-broadcasts per call (lut.c:533-540). */
int sbg_stage_problem(sbg_handle *h, int slot, const uint64_t *tables, int n,
    const uint64_t *target, const uint64_t *mask, const int8_t *inbits);
int sbg_use_problem(sbg_handle *h, int slot);

/* ---- whole searches on one device (host buffers in, result out) ----------------------------- */
int sbg_search5(sbg_handle *h, const uint8_t *func_order /*256*/, sbg_result *res);
int sbg_search7(sbg_handle *h, const uint8_t *outer_order /*256*/, const uint8_t *middle_order
    /*256*/, sbg_result *res);

/* ---- sharded building blocks (one process per GPU; part = rank, nparts = world size) -------- */
/* 5-LUT: this part's share of C(n,5); *key = local minimum key or SBG_KEY_NONE. */
int sbg_search5_part(sbg_handle *h, int part, int nparts, const uint8_t *func_order,
    uint64_t *key);
int sbg_finish5(sbg_handle *h, uint64_t key, const uint8_t *func_order, sbg_result *res);

/* 7-LUT phase 1: this part's share of C(n,7).  Writes this part's feasible combinations, sorted,
   at most SBG_LIST_CAP of them, as packed 63-bit words (9 bits per gate, first gate in the most
   significant position, so integer order = lexicographic order) to `list` (host memory, room for
   SBG_LIST_CAP entries) and their number to *count.  `list` may be NULL (no copy).  With
   nparts == 1 the device-resident result is installed as the list, so sbg_decomp7_part() may follow
   directly. */
int sbg_filter7_part(sbg_handle *h, int part, int nparts, uint64_t *list, int *count);
/* Installs the merged hit list (any order, duplicates not allowed; it is sorted and truncated to
   SBG_LIST_CAP here, which reproduces lut.c:316-318 for size == 1). */
int sbg_set_list7(sbg_handle *h, const uint64_t *list, int count);
/* 7-LUT phase 2 over list indices congruent to part modulo nparts. */
int sbg_decomp7_part(sbg_handle *h, int part, int nparts, const uint8_t *outer_order,
    const uint8_t *middle_order, uint64_t *key);
int sbg_finish7(sbg_handle *h, uint64_t key, const uint8_t *outer_order,
    const uint8_t *middle_order, sbg_result *res);

/* ---- helpers shared with the host side ------------------------------------------------------ */
/* Test hook, no device needed: how a sweep's work is cut into tickets ("Dense
   states").  width = 7 with prefix_gates = 4 or 5 (search_7lut phase 1), width = 5 with
   prefix_gates = 3 (search_5lut, fused kernel); n >= 8 gates, `excluded` = bit g set for an excluded
   gate g < 8 (lut.c:177-185); mode 0 = whole-prefix tickets only, 1 = a head of `waves` waves of
   (prefix, chunk) tickets over the allowed gates, 2 = chunk tickets throughout.
   out[0] = chunk tickets' items, out[1] = chunks per prefix, out[2] = lexicographic rank, among all
   prefixes, of the first prefix left to the whole-prefix tickets (= out[3] if none is left),
   out[3] = number of prefixes. */
int sbg_plan_tickets(int width, int prefix_gates, int n, uint32_t excluded, int mode,
                     uint64_t waves, uint64_t *out);

/* Row k of the ordering tables (lut.c:189-229 for width 5, lut.c:396-415 for width 7). */
int sbg_ordering_row(int width, int k, int *row);
/* Closed form of get_lut_function without the random fill (lut.c:79-103): returns 1 and the
   solved function / seen mask, or 0 on conflict. */
int sbg_solve_inner(const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    const uint64_t *target, const uint64_t *mask, uint8_t *func, uint8_t *seen);
/* generate_lut_ttable (state.c:202-230). */
void sbg_lut_table(uint8_t func, const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* SBOXGATES_B200_H */
