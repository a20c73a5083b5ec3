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
#define SBG_LANES 8             /* searches of one sbg_search_batch() call that run concurrently */

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
/* Number of kernels the handle has launched so far (all of them this library's own; bench.py
   reports it). */
uint64_t sbg_launch_count(const sbg_handle *h);
/* Kernel timing is off by default (an event between two kernels of a chain forbids their
   overlap); sbg_set_timing(h, 1) or SBG_TIMING=1 turns it on.  sbg_last_kernel_ms() then gives the
   CUDA-event time (ms) of the named kernel family in the last call (summed over the searches of a
   batch): 0 = search5, 1 = filter7 (phase 1), 2 = ordering the hit list, 3 = decomp7. */
int sbg_set_timing(sbg_handle *h, int on);
float sbg_last_kernel_ms(const sbg_handle *h, int which);
/* Bytes this handle has shipped so far: out[0] = problem data host->device (gate tables that
   changed, target, mask; as kernel arguments or copies), out[1] = results device->host (mapped
   memory and copies), out[2..4] = state changes that took a bulk copy / travelled as kernel
   arguments / were no change at all. */
int sbg_transfer_stats(const sbg_handle *h, uint64_t *out /*5*/);
/* Host seconds sbg_search_node() has spent so far: out[0] = enqueueing the chains (launch calls),
   out[1] = waiting for and decoding their results; out[2..4] = the waiting alone, per stage
   (3-LUT scan, search_5lut, search_7lut; all calls of the handle). */
int sbg_host_seconds(const sbg_handle *h, double *out /*5*/);
/* Measures the device's LOP3 issue rate (warp instructions per second, whole chip): the ceiling
   the search kernels are bound by. */
int sbg_alu_peak(sbg_handle *h, double *warp_instr_per_s);

/* ---- problem -------------------------------------------------------------------------------- */
/* Makes one search state the current problem: n gate tables, target, mask, and the -1 terminated
   list of input-bit gates already used as multiplexer selectors (lut.c:177-185).  Host pointers.
   The gate tables stay resident on the device between calls (the replacement of the `state`
   array the reference re-broadcasts per call): only the gates that differ from the previous state
   of the slot are shipped -- a graph build only ever appends gates (state.h:87) or replaces a
   suffix when it backtracks -- together with target and mask, as arguments of the next search's
   first kernel; compression to the masked positions happens on the device. */
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

/* ---- one call per node, batches of independent nodes ----------------------------------------- */
/* A job is what lut_search() does for one node (lut.c:489-631): the 3-LUT scan over the caller's
   shuffled gate order (lut.c:501-523), search_5lut (lut.c:553) and search_7lut (lut.c:593), each
   stage only if the earlier ones found nothing -- as ONE launch chain on the device, the stages
   predicated there, no host round trip in between.  The caller says which stages it wants
   (lut_search skips search_5lut / search_7lut when the gate budget forbids two / three more gates,
   lut.c:525-527, 582-586). */
#define SBG_DO_SCAN3 1
#define SBG_DO_SEARCH5 2
#define SBG_DO_SEARCH7 4
typedef struct {
  int32_t slot;                /* staged problem (sbg_stage_problem / sbg_load_problem = slot 0) */
  int32_t flags;               /* SBG_DO_* */
  const uint8_t *order5;       /* 256: shuffled function order of search_5lut (lut.c:125-135) */
  const uint8_t *outer7;       /* 256 + 256: the two orders of search_7lut (lut.c:362-378) */
  const uint8_t *middle7;
  const uint16_t *gate_order;  /* n: create_circuit's shuffled gate order (sboxgates.c:285-299) */
} sbg_job;
typedef struct {
  int32_t found_stage;         /* 0 nothing, else 3 / 5 / 7 */
  uint16_t gates3[3];          /* stage 3: LUT inputs in gate_order order (gi, gk, gm) */
  uint8_t func3, seen3;        /* solved function bits / cells seen under the mask (fill as below) */
  uint64_t key3;               /* the position triple in gate_order, i << 18 | k << 9 | m, or
                                  SBG_KEY_NONE */
  sbg_result r5;               /* as sbg_search5 (found = 0 if the stage did not run) */
  sbg_result r7;               /* as sbg_search7 */
} sbg_node_result;
/* One node; returns as soon as a stage has matched. */
int sbg_search_node(sbg_handle *h, const sbg_job *job, sbg_node_result *result);
/* Independent nodes (the first children of a create_circuit node, sboxgates.c:458-607; the output
   bits of generate_graph, sboxgates.c:701-788; the -i iterations): their chains run concurrently
   on up to SBG_LANES streams, results are read once per job. */
int sbg_search_batch(sbg_handle *h, int njobs, const sbg_job *jobs, sbg_node_result *results);

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
/* Installs the merged hit list: `list` is the concatenation of the parts' ascending lists (at most
   64 ascending runs, no duplicates); the runs are merged on the device and cut at SBG_LIST_CAP,
   which reproduces lut.c:316-349 for size == 1. */
int sbg_set_list7(sbg_handle *h, const uint64_t *list, int count);
/* The same without touching the host: this part's ordered list as a device pointer (valid until the
   next search on the handle), and the merge of `nruns` ascending runs that already sit in device
   memory, run r at runs + r * stride with counts[r] entries (what an all-gather of the parts'
   lists into one buffer gives). */
int sbg_list7_device(sbg_handle *h, const uint64_t **list, int *count);
int sbg_set_list7_device(sbg_handle *h, const uint64_t *runs, uint64_t stride, const int *counts,
    int nruns);
/* One process driving several devices (handles hs[0..nh-1], one per device, each holding its part's
   ordered list from sbg_filter7_part(part = i, nparts = nh)): gathers every part's list onto every
   device (peer copies over NVLink) and merges them there; afterwards every handle has the same
   installed list.  *total = its length.  The in-process counterpart of the all-gather
   (lut.c:329-349). */
int sbg_allgather_merge7(sbg_handle *const *hs, int nh, int *total);
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

/* Test hook, no device needed: the ticket tables of search_7lut phase 1's weighted form (4-gate
   prefixes cut into groups of `group_pairs` (e,f) pairs, tickets numbered in lexicographic order of
   (prefix, group); 7 <= n <= 72).  out[0] = tickets of the whole sweep, out[1 + 76 * (r-1) + x] =
   tickets of all ways to choose r more prefix gates the first of which is >= x (r = 1..4). */
#define SBG_WEIGHTED_ROW 76
int sbg_weighted_tickets(int n, uint32_t group_pairs, uint32_t *out);

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
