// sbg_api.cu -- host side of libsboxgates_b200.so: the C ABI declared in
// include/sboxgates_b200.h.  No search logic runs on the CPU here except the O(256) decode of the
// winning key into the reference's ret[] vocabulary (sbg_finish5 / sbg_finish7); there is no CPU
// fallback -- without a CUDA device sbg_create() fails.
#include "sbg_device.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <utility>

#include "../../include/sboxgates_b200.h"

using namespace sbg;

namespace {

// ---- combinatorics and ordering tables (host copies) -------------------------------------------

uint64_t h_binom[501][8];
int h_rows7[70][7];
int h_rows5[10][5];
bool g_tables_ready = false;

void build_host_tables() {
  if (g_tables_ready) return;
  for (int m = 0; m <= 500; m++) {
    for (int r = 0; r < 8; r++) {
      if (r > m) {
        h_binom[m][r] = 0;
      } else if (r == 0) {
        h_binom[m][r] = 1;
      } else {
        h_binom[m][r] = h_binom[m - 1][r - 1] + (r <= m - 1 ? h_binom[m - 1][r] : 0);
      }
    }
  }
  // lut.c:189,224-229: outer = 3-subsets of {0..4} in lexicographic order, rest ascending.
  int k = 0;
  for (int a = 0; a < 5; a++) for (int b = a + 1; b < 5; b++) for (int c = b + 1; c < 5; c++) {
    int w = 3;
    h_rows5[k][0] = a; h_rows5[k][1] = b; h_rows5[k][2] = c;
    for (int i = 0; i < 5; i++) {
      if (i != a && i != b && i != c) h_rows5[k][w++] = i;
    }
    k++;
  }
  // lut.c:396-415: outer = 3-subsets of {0..6} (lexicographic), middle = 3-subsets of the other
  // four (lexicographic), kept iff min(outer) < min(middle); last = the leftover position.
  k = 0;
  for (int a = 0; a < 7; a++) for (int b = a + 1; b < 7; b++) for (int c = b + 1; c < 7; c++) {
    int rest[4], r = 0;
    for (int i = 0; i < 7; i++) {
      if (i != a && i != b && i != c) rest[r++] = i;
    }
    for (int skip = 3; skip >= 0; skip--) {
      int mid[3], m = 0;
      for (int i = 0; i < 4; i++) {
        if (i != skip) mid[m++] = rest[i];
      }
      if (a >= mid[0]) continue;
      h_rows7[k][0] = a; h_rows7[k][1] = b; h_rows7[k][2] = c;
      h_rows7[k][3] = mid[0]; h_rows7[k][4] = mid[1]; h_rows7[k][5] = mid[2];
      h_rows7[k][6] = rest[skip];
      k++;
    }
  }
  g_tables_ready = true;
}

void unrank_combination(uint64_t rank, int n, int t, uint16_t *out) {
  int x = 0;
  for (int pos = 0; pos < t; pos++) {
    for (;; x++) {
      const uint64_t cnt = h_binom[n - x - 1][t - pos - 1];
      if (rank < cnt) break;
      rank -= cnt;
    }
    out[pos] = (uint16_t)x++;
  }
}

int popcount256(const uint64_t *m) {
  return __builtin_popcountll(m[0]) + __builtin_popcountll(m[1]) + __builtin_popcountll(m[2])
      + __builtin_popcountll(m[3]);
}

// Gathers the bits of `t` at the set positions of `mask` into the low bits of out[0..7].
void compress_table(const uint64_t *t, const uint64_t *mask, uint32_t *out) {
  uint64_t acc[4] = {0, 0, 0, 0};
  int fill = 0;
  for (int v = 0; v < 4; v++) {
    uint64_t m = mask[v];
    const uint64_t x = t[v];
#if defined(__BMI2__)
    const uint64_t bits = __builtin_ia32_pext_di(x, m);
    const int cnt = __builtin_popcountll(m);
#else
    uint64_t bits = 0;
    int cnt = 0;
    while (m != 0) {
      const int b = __builtin_ctzll(m);
      m &= m - 1;
      bits |= ((x >> b) & 1ull) << cnt;
      cnt++;
    }
#endif
    if (cnt == 0) continue;
    acc[fill >> 6] |= bits << (fill & 63);
    if ((fill & 63) + cnt > 64) acc[(fill >> 6) + 1] |= bits >> (64 - (fill & 63));
    fill += cnt;
  }
  for (int v = 0; v < 4; v++) {
    out[2 * v] = (uint32_t)acc[v];
    out[2 * v + 1] = (uint32_t)(acc[v] >> 32);
  }
}

}  // namespace

// ---- handle ----------------------------------------------------------------------------------

struct sbg_handle {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;

  DevProblem *d_prob = nullptr;  // the problem in use (points into d_slots)
  DevProblem *d_slots = nullptr; // kSlots device-resident problems
  DevProblem *h_prob = nullptr;  // pinned staging copy
  // per-call block: control words + 7-LUT parameters, contiguous so one copy uploads both
  struct DevCall {
    DevCtl ctl;
    DevParams7 par;
  };
  DevCall *d_call = nullptr;
  DevCall *h_call = nullptr;     // pinned
  DevCtl *d_ctl = nullptr;       // header of d_sorted_block
  DevCtl *h_ctl = nullptr;       // = &h_call->ctl
  DevParams7 *d_par7 = nullptr;  // = &d_call->par
  DevParams7 *h_par7 = nullptr;  // = &h_call->par
  DevCtl *h_ctl_out = nullptr;   // pinned: control words read back
  uint64_t *h_head = nullptr;    // pinned: first kHeadEntries of the sorted list, read back with them
  DevTables *d_tab = nullptr;    // lane-indexed ordering tables
  uint8_t *d_pos5 = nullptr;     // written by k_begin5 from its argument

  uint64_t *d_hits = nullptr;    // unordered feasible tuples of this device
  char *d_sorted_block = nullptr;  // [control words, 128 B][sorted list]
  uint64_t *d_sorted = nullptr;  // sorted copy
  uint64_t *d_list = nullptr;    // installed list (points into d_sorted or d_hits)
  size_t hits_cap = 0;
  void *d_cub = nullptr;
  size_t cub_bytes = 0;
  uint32_t list_count = 0;
  bool list_ready = false;

  // host copies of the staged problems (for sbg_finish*); the one in use is mirrored below
  struct HostProblem {
    uint64_t tables[SBG_MAX_GATES][4];
    uint64_t target[4];
    uint64_t mask[4];
    int n = 0;
    int nw = 0;
    uint32_t inmask = 0;
    bool ready = false;
    bool rows_ready = false;   // DevProblem::xr built on the device
  };
  HostProblem *slots = nullptr;  // kSlots entries
  uint64_t (*tables)[4] = nullptr;
  uint64_t *target = nullptr;
  uint64_t *mask = nullptr;
  int n = 0;
  int nw = 0;
  int cur_slot = 0;
  uint32_t inmask = 0;
  bool problem_ready = false;

  uint64_t swept = 0;
  uint64_t feasible = 0;
  std::map<std::pair<const void *, size_t>, int> occupancy;  // grid_for's cache
  // tuning knobs, read from the environment when the handle is created (tests create handles under
  // different settings to cross-check the alternative kernels against each other)
  int opt_batch = 0;        // SBG_BATCH: prefixes per ticket batch (0 = automatic)
  int opt_pm_prefix = 0;    // SBG_PM_PREFIX: 4 or 5 (0 = by n)
  int opt_filter = 0;       // SBG_FILTER: 0 position-major, 1 bitmap sweep
  int opt_search5 = 0;      // SBG_SEARCH5: 0 by size, 1 fused, 2 two kernels
  int opt_head = -1;        // SBG_HEAD: chunked phase of the 7-LUT filter, 0 none, 1 first prefixes,
                            // 2 everything (-1 = by n and mask size)
  int opt_shift = -1;       // SBG_SHIFT: phase-1 shifted single-word windows, 0 never, 1 whenever n <= 63
  int opt_head_waves = 0;   // SBG_HEAD_WAVES: size of the chunked head in waves of warps (0 = default)
  uint64_t launches = 0;      // our kernels
  uint64_t lib_launches = 0;  // CUB radix-sort kernels
  float ms[4] = {0, 0, 0, 0};
  bool sort_pending = false;
  cudaEvent_t ev[8];
  char err[512] = {0};
};

namespace {

constexpr int kSlots = SBG_PROBLEM_SLOTS;
constexpr size_t kPerPrefixMax = SBG_LIST_CAP + 32 * 512;  // hits one prefix can emit
constexpr size_t kHeadEntries = 1024;  // list entries read back together with the control words

int fail(sbg_handle *h, int code, const char *fmt, ...) {
  if (h != nullptr) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define SBG_CUDA(h, call)                                                                  \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      return fail((h), SBG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
          __FILE__, __LINE__);                                                             \
    }                                                                                      \
  } while (0)

template <int NW, int P>
size_t sweep_smem(int n) {
  const int npad = (n + 3) & ~3;
  return sizeof(uint32_t) * (size_t)(NW * npad + kWarpsPerCta * (1 << P) * 2 * NW);
}

template <int NW>
size_t decomp_smem(int n) {
  const int npad = (n + 3) & ~3;
  return sizeof(uint32_t) * (size_t)(NW * npad);
}

// Persistent grid: as many CTAs as are resident at once, but no more than there is work for.
template <typename Kernel>
int grid_for(sbg_handle *h, Kernel kernel, size_t smem, uint64_t work_items_in_warps) {
  // the occupancy query costs a few microseconds; a search makes several launches and a graph
  // build makes tens of thousands of searches, so remember the answer per (kernel, smem size)
  auto &cache = h->occupancy;   // per handle: handles may be driven from different threads
  const auto key = std::make_pair(reinterpret_cast<const void *>(kernel), smem);
  int per_sm = 1;
  auto it = cache.find(key);
  if (it != cache.end()) {
    per_sm = it->second;
  } else {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem);
    cache[key] = per_sm;
  }
  if (per_sm < 1) per_sm = 1;
  uint64_t want = (work_items_in_warps + kWarpsPerCta - 1) / kWarpsPerCta;
  uint64_t cap = (uint64_t)per_sm * (uint64_t)h->sm_count;
  if (want < 1) want = 1;
  return (int)std::min(want, cap);
}

// Prefixes per ticket batch.  One batch costs one global atomic; batches should hold enough pairs
// to hide its latency (~64 chunks of 32), leave several batches per resident warp, and never hold
// more work than a warp's fair share (prefixes are dealt heaviest first, so the tail evens out).
//
// The batch size also defines how work is dealt to the parts of a sharded search (part p takes
// batches p, p + nparts, ...), so it must be the same on every rank: it is a function of the
// problem and a nominal warp count only, never of the device a rank happens to run on.
constexpr uint64_t kNominalWarps = 148 * 2 * kWarpsPerCta;

constexpr int kSinglePrefixMaxGates = 72;
// Shifted single-word windows in phase 1 up to here (SBG_SHIFT=0|1 overrides).  Measured against
// the aligned two-word windows, full mask (scripts/sweep_shift.sh): n = 32 / 40 / 48 / 56 / 63:
// 0.111 / 0.338 / 1.045 / 2.146 / 4.369 ms -> 0.086 / 0.287 / 0.869 / 1.958 / 4.301 ms.
constexpr int kShiftMaxGates = 60;
uint64_t pick_batch(const sbg_handle *h, uint64_t tickets, int n, int P) {
  const uint64_t warps = kNominalWarps;
  if (h->opt_batch > 0) {
    uint64_t b = std::max<uint64_t>(1, std::min<uint64_t>(16, (uint64_t)h->opt_batch));
    while (b & (b - 1)) b &= b - 1;
    return b;
  }
  // work per ticket in lane-items: (f,g) pairs for the sweeps (P = 3, 5), (e,f) pairs for the
  // position-major kernel with 4-gate prefixes (P = 4), single f for its 5-gate form (P = 6)
  const bool pm = P == 4 || P == 6;
  // position-major kernel, 4-gate prefixes: single prefixes up to n = 72 (measured, scripts/
  // sweep_head.sh / sweep_batch.sh: n = 48 / 64 full mask 1.07 / 7.08 ms against 1.26 / 8.99 ms with
  // the formula below; from n = 80 on the formula's 4 is as good or better)
  if (P == 4 && n <= kSinglePrefixMaxGates) return 1;
  const uint64_t total = pm ? h_binom[n - 1][6] : h_binom[n][P + 2];
  const uint64_t avg_pairs = std::max<uint64_t>(1, total / std::max<uint64_t>(1, tickets));
  const uint64_t qmax = P == 6 ? (uint64_t)std::max(1, n - 7) : h_binom[n - P - (P == 4 ? 1 : 0)][2];
  uint64_t b = (64 * 32 + avg_pairs - 1) / avg_pairs;
  b = std::min<uint64_t>(b, std::max<uint64_t>(1, tickets / (warps * 4)));
  b = std::min<uint64_t>(b, std::max<uint64_t>(1, total / (warps * std::max<uint64_t>(1, qmax))));
  b = std::max<uint64_t>(1, std::min<uint64_t>(16, b));
  while (b & (b - 1)) b &= b - 1;   // power of two: a batch must not straddle two deal blocks
  return b;
}

// Chunked phase of phase 1 (see k_filter7_pm): a head of kHeadWaves waves of (prefix, chunk) items
// in front of the prefix form.  A list that fills early -- small masks make most combinations
// feasible -- then costs microseconds instead of the first wave of whole-prefix batches, which at
// large n is enormous (measured before: n = 200 / 300 / 500 with 16-32 masked positions took 2.6 /
// 20 / 258 s through the overflow retry, now 0.1-0.35 ms; profiles/r01_dense_cases.md).  Sweeping
// EVERYTHING in chunk items is much slower where the sweep has to cover the space (n = 128, 64
// positions: 3.8 s against 0.25 s), so that form is kept for the overflow retry.  Below
// kHeadAlwaysMinGates the head is used for small masks only, below kHeadMinGates never (measured:
// the rijndael -o 0 run and bench.py at n = 40 are indifferent to it).
constexpr int kHeadAlwaysMinGates = 128;
constexpr int kHeadMaxPositions = 64;
constexpr int kHeadMinGates = 48;
constexpr uint64_t kHeadWaves = 64;
constexpr uint64_t kHeadWaves5 = 16;   // search_5lut: the first match is what ends it, a short head does
constexpr size_t kPerChunkMax = 32 * 512;   // hits one (prefix, chunk) item can emit

struct ChunkPlan {
  unsigned long long items = 0;     // (prefix, chunk) items of the chunked phase
  int chunks = 0;                   // chunks per prefix
  unsigned long long t_offset = 0;  // rank of the first prefix left to the prefix form
  bool all = false;                 // the chunked phase covers everything
};

// K = size of the combinations, P = gates per prefix, mode: 0 none, 1 a head of `waves` waves of chunk
// tickets, 2 everything in chunk tickets.
template <int P, int K>
ChunkPlan plan_chunks_mode(int n, uint32_t inmask, int mode, uint64_t waves, uint64_t qmax) {
  ChunkPlan pl;
  const int na = n - __builtin_popcount(inmask & 0xffu);   // allowed gates
  const uint64_t total_c = na >= K ? h_binom[na - (K - P)][P] : 0;
  if (mode == 0 || total_c == 0) return pl;
  // qmax = lane items of the first (largest) prefix
  pl.chunks = (int)std::max<uint64_t>(1, (qmax + 31) / 32);
  uint64_t prefixes = total_c;
  if (mode == 1) {
    prefixes = std::min<uint64_t>(total_c,
        std::max<uint64_t>(1, waves * kNominalWarps / (uint64_t)pl.chunks));
  }
  pl.items = prefixes * (uint64_t)pl.chunks;
  pl.all = prefixes == total_c;
  if (!pl.all) {
    // the first allowed prefix not covered: index `prefixes` among the P-subsets of the allowed
    // gates, as gate numbers, ranked among the P-subsets of all gates
    int c[P];
    uint64_t t = prefixes;
    const int np = na - (K - P);
    int x = 0;
    for (int pos = 0; pos < P; pos++) {
      for (;; x++) {
        const uint64_t cnt = h_binom[np - x - 1][P - pos - 1];
        if (t < cnt) break;
        t -= cnt;
      }
      c[pos] = x++;
    }
    for (int i = 0; i < P; i++) {
      int g = c[i];
      for (int bit = 0; bit < 8; bit++) g += (((inmask >> bit) & 1u) != 0 && bit <= g) ? 1 : 0;
      c[i] = g;
    }
    const int nr = n - (K - P);
    uint64_t rank = 0;
    int prev = -1;
    for (int pos = 0; pos < P; pos++) {
      for (int y = prev + 1; y < c[pos]; y++) rank += h_binom[nr - y - 1][P - pos - 1];
      prev = c[pos];
    }
    pl.t_offset = rank;
  }
  return pl;
}

template <int P>
int launch_sweep(sbg_handle *h, int part, int nparts, int max_warps, bool emit5 = false) {
  const int n = h->n;
  const uint64_t total = h_binom[n - 2][P];
  const unsigned long long cap = h->hits_cap;
  // search_5lut on a large state (fused kernel): a head of chunk tickets, so that on a dense state
  // the feasible-but-not-decomposable tuples in front of the first match are spread over the
  // machine instead of being decomposed by the one warp that owns their prefix
  ChunkPlan pl;
  if (P == 3 && !emit5 && max_warps == 0 && h->opt_head != 0 && n >= kHeadAlwaysMinGates) {
    pl = plan_chunks_mode<P, P + 2>(n, h->inmask, 1, kHeadWaves5, h_binom[n - 3][2]);
  }
  const uint64_t tickets = pl.all ? 0 : (total - pl.t_offset + nparts - 1) / nparts;
  const uint64_t chunk_tickets = (pl.items + kDeal * nparts - 1) / (kDeal * nparts) * kDeal;
#define SBG_LAUNCH_SWEEP(NWV)                                                                  \
  {                                                                                            \
    const size_t smem = sweep_smem<NWV, P>(n);                                                 \
    int grid = grid_for(h, k_sweep<NWV, P>, smem, tickets + chunk_tickets);                    \
    if (max_warps > 0) grid = std::min(grid, (max_warps + kWarpsPerCta - 1) / kWarpsPerCta);   \
    uint64_t bsz = pick_batch(h, tickets, n, P);                                               \
    if (max_warps > 0) bsz = 1;                                                                \
    k_sweep<NWV, P><<<grid, kThreads, smem, h->stream>>>(h->d_prob, h->d_ctl, h->d_pos5,       \
        h->d_hits, cap, part, nparts, (unsigned long long)SBG_LIST_CAP, (int)bsz, max_warps,   \
        emit5, h->d_tab, pl.all ? (unsigned long long)total : pl.t_offset, pl.items,           \
        std::max(1, pl.chunks), chunk_tickets);                                                \
  }
  switch (h->nw) {
    case 1: SBG_LAUNCH_SWEEP(1) break;
    case 2: SBG_LAUNCH_SWEEP(2) break;
    case 4: SBG_LAUNCH_SWEEP(4) break;
    default: SBG_LAUNCH_SWEEP(8) break;
  }
#undef SBG_LAUNCH_SWEEP
  h->launches++;
  SBG_CUDA(h, cudaGetLastError());
  return SBG_OK;
}

template <int NW, int P>
size_t filter_pm_smem(int n, int m, bool shifted = false) {
  const int npad = (n + 3) & ~3;
  const int ngw = (((n + 31) >> 5) + 1) & ~1;
  return sizeof(uint32_t) * (size_t)(NW * npad + ((m * ngw + 3) & ~3) + kWarpsPerCta * (1 << P) * NW
      + (shifted ? kWarpsPerCta * m : 0));
}

// Position-major phase 1 (k_filter7_pm): work items are 4- or 5-gate prefixes.  The 5-gate form does
// half the work per visited position but keeps only n-7-ish lanes of a warp busy, so it is used
// from n = kPm5MinGates on (measured cross-over, profiles/); SBG_PM_PREFIX=4|5 overrides.
constexpr int kPm5MinGates = 128;
template <int P>
ChunkPlan plan_chunks(const sbg_handle *h, int m, bool retry) {
  const int n = h->n;
  int mode = h->opt_head;
  if (mode < 0) {
    mode = n >= kHeadAlwaysMinGates || (m <= kHeadMaxPositions && n >= kHeadMinGates) ? 1 : 0;
  }
  // overflow retry: one form for everything, so that the bound on the hits in flight is simple --
  // chunk items where a prefix is large, whole prefixes otherwise
  if (retry) mode = n >= kHeadAlwaysMinGates ? 2 : 0;
  // lane items: (e,f) pairs out of the n-5 gates that leave room for g; single f for 5-gate prefixes
  const uint64_t qmax = P == 4 ? h_binom[n - 5][2] : (uint64_t)(n - 6);
  return plan_chunks_mode<P, 7>(n, h->inmask, mode,
      h->opt_head_waves > 0 ? (uint64_t)h->opt_head_waves : kHeadWaves, qmax);
}

// retry: the hit buffer overflowed; run again with the number of working warps bounded so that it
// cannot (tickets taken synchronously, see the kernel).
template <int P>
int launch_filter7_pm_p(sbg_handle *h, int part, int nparts, bool retry) {
  const int n = h->n;
  const int m = popcount256(h->mask);
  const uint64_t total = h_binom[n - (7 - P)][P];
  const unsigned long long cap = h->hits_cap;
  const ChunkPlan pl = plan_chunks<P>(h, m, retry);
  // No item is handed out once the list cap is reached, so with w warps at work the buffer holds
  // fewer than cap + w x (hits one item can emit) entries; a whole prefix stops by itself after
  // cap + one chunk.
  const int max_warps = !retry ? 0 : (int)std::max<size_t>(1, pl.all
      ? (h->hits_cap - SBG_LIST_CAP) / kPerChunkMax : h->hits_cap / kPerPrefixMax - 1);
  const uint64_t tickets = pl.all ? 0 : (total - pl.t_offset + nparts - 1) / nparts;
  // chunk tickets of one part: whole deal blocks, the same count for every part
  const uint64_t chunk_tickets = (pl.items + kDeal * nparts - 1) / (kDeal * nparts) * kDeal;
#define SBG_LAUNCH_PM(NWV, WV, FSV, SHV)                                                       \
  {                                                                                            \
    const size_t smem = filter_pm_smem<NWV, P>(n, m, SHV);                                     \
    int grid = grid_for(h, k_filter7_pm<NWV, WV, P, FSV, SHV>, smem, tickets + chunk_tickets); \
    if (max_warps > 0) grid = std::min(grid, (max_warps + kWarpsPerCta - 1) / kWarpsPerCta);   \
    uint64_t bsz = pl.all ? 1 : pick_batch(h, tickets, n, P == 4 ? 4 : 6);                     \
    if (max_warps > 0) bsz = 1;                                                                \
    k_filter7_pm<NWV, WV, P, FSV, SHV><<<grid, kThreads, smem, h->stream>>>(h->d_prob,         \
        h->d_ctl,                                                                              \
        h->d_hits, cap, part, nparts, (unsigned long long)SBG_LIST_CAP, (int)bsz, max_warps,   \
        pl.all ? (unsigned long long)total : pl.t_offset, pl.items, std::max(1, pl.chunks),    \
        chunk_tickets);                                                                        \
    h->launches++;                                                                             \
  }
  const bool shifted = P == 4 && (h->opt_shift >= 0 ? h->opt_shift != 0 && n <= 63
                                                     : n <= kShiftMaxGates);
  if constexpr (P == 4) {
    if (shifted) {         // one word of 31 candidate gates from the first possible g on
      switch (h->nw) {
        case 1: SBG_LAUNCH_PM(1, 1, true, true) break;
        case 2: SBG_LAUNCH_PM(2, 1, true, true) break;
        case 4: SBG_LAUNCH_PM(4, 1, true, true) break;
        default: SBG_LAUNCH_PM(8, 1, true, true) break;
      }
    }
  }
  if (shifted) {
  } else if (n <= 31) {  // one word of candidate gates per pass, its top bit free
    switch (h->nw) {
      case 1: SBG_LAUNCH_PM(1, 1, true, false) break;
      case 2: SBG_LAUNCH_PM(2, 1, true, false) break;
      case 4: SBG_LAUNCH_PM(4, 1, true, false) break;
      default: SBG_LAUNCH_PM(8, 1, true, false) break;
    }
  } else if (n <= 63) {  // two words, one pass, top bit free
    switch (h->nw) {
      case 1: SBG_LAUNCH_PM(1, 2, true, false) break;
      case 2: SBG_LAUNCH_PM(2, 2, true, false) break;
      case 4: SBG_LAUNCH_PM(4, 2, true, false) break;
      default: SBG_LAUNCH_PM(8, 2, true, false) break;
    }
  } else {
    switch (h->nw) {
      case 1: SBG_LAUNCH_PM(1, 2, false, false) break;
      case 2: SBG_LAUNCH_PM(2, 2, false, false) break;
      case 4: SBG_LAUNCH_PM(4, 2, false, false) break;
      default: SBG_LAUNCH_PM(8, 2, false, false) break;
    }
  }
#undef SBG_LAUNCH_PM
  SBG_CUDA(h, cudaGetLastError());
  return SBG_OK;
}

// Position-major rows of the problem in use, built on the device the first time phase 1 needs them.
int ensure_rows(sbg_handle *h) {
  sbg_handle::HostProblem &hp = h->slots[h->cur_slot];
  if (hp.rows_ready) return SBG_OK;
  const int m = popcount256(h->mask);
  if (m > 0) {
    k_build_rows<<<(m * 16 + 255) / 256, 256, 0, h->stream>>>(h->d_slots + h->cur_slot);
    h->launches++;
    SBG_CUDA(h, cudaGetLastError());
  }
  hp.rows_ready = true;
  return SBG_OK;
}

int launch_filter7_pm(sbg_handle *h, int part, int nparts, bool retry) {
  int rc = ensure_rows(h);
  if (rc != SBG_OK) return rc;
  const bool five = h->opt_pm_prefix != 0 ? h->opt_pm_prefix == 5 : h->n >= kPm5MinGates;
  return five ? launch_filter7_pm_p<5>(h, part, nparts, retry)
              : launch_filter7_pm_p<4>(h, part, nparts, retry);
}

// Which phase-1 kernel: the position-major one unless SBG_FILTER=sweep asks for the bitmap sweep.
bool use_position_major(const sbg_handle *h) {
  return h->opt_filter == 0;
}

// count_on_device: the list length is ctl->list_count (written by k_sort_small); the grid is then
// sized for the longest list that kernel sorts, surplus CTAs return before staging anything.
int launch_decomp7(sbg_handle *h, int part, int nparts, bool count_on_device = false) {
  const int n = h->n;
  const uint64_t items = count_on_device ? (uint64_t)kSmallSort
                                         : (h->list_count + nparts - 1) / nparts;
  const unsigned int count_arg = count_on_device ? 0xffffffffu : h->list_count;
#define SBG_LAUNCH_DECOMP(NWV)                                                                 \
  {                                                                                            \
    const size_t smem = decomp_smem<NWV>(n);                                                   \
    const int grid = grid_for(h, k_decomp7<NWV>, smem, items);                                 \
    k_decomp7<NWV><<<grid, kThreads, smem, h->stream>>>(h->d_prob, h->d_ctl, h->d_par7,        \
        h->d_list, count_arg, part, nparts, h->d_tab);                                         \
  }
  switch (h->nw) {
    case 1: SBG_LAUNCH_DECOMP(1) break;
    case 2: SBG_LAUNCH_DECOMP(2) break;
    case 4: SBG_LAUNCH_DECOMP(4) break;
    default: SBG_LAUNCH_DECOMP(8) break;
  }
#undef SBG_LAUNCH_DECOMP
  h->launches++;
  SBG_CUDA(h, cudaGetLastError());
  return SBG_OK;
}

int reset_ctl(sbg_handle *h) {
  DevCtl *c = h->h_ctl;
  memset(c, 0, sizeof(*c));
  c->best = ~0ull;
  c->stop_ticket = ~0ull;
  SBG_CUDA(h, cudaMemcpyAsync(h->d_ctl, c, sizeof(DevCtl), cudaMemcpyHostToDevice, h->stream));
  return SBG_OK;
}

// The control words sit in a 128-byte header in front of the sorted list, so that one copy brings
// back both them and (with_head) the first kHeadEntries list entries.
constexpr size_t kCtlHeaderBytes = 128;
static_assert(sizeof(DevCtl) <= kCtlHeaderBytes, "control words must fit the header");

int fetch_ctl(sbg_handle *h, bool with_head = false) {
  SBG_CUDA(h, cudaMemcpyAsync(h->h_ctl_out, h->d_ctl,
      with_head ? kCtlHeaderBytes + kHeadEntries * sizeof(uint64_t) : sizeof(DevCtl),
      cudaMemcpyDeviceToHost, h->stream));
  SBG_CUDA(h, cudaStreamSynchronize(h->stream));
  *h->h_ctl = *h->h_ctl_out;
  return SBG_OK;
}

float elapsed(sbg_handle *h, int a, int b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]) != cudaSuccess) {
    (void)cudaGetLastError();  // do not leave a sticky "last error" behind
    return 0.f;
  }
  return ms;
}

int sort_hits(sbg_handle *h, uint64_t *d_in, uint64_t *d_out, size_t count) {
  size_t need = h->cub_bytes;
  SBG_CUDA(h, cub::DeviceRadixSort::SortKeys(h->d_cub, need, d_in, d_out, (int)count, 0, 63,
      h->stream));
  h->lib_launches += 3;  // CUB's histogram + onesweep passes: library kernels, not ours
  return SBG_OK;
}

// Phase 1 on this device: leaves the sorted local list (<= SBG_LIST_CAP) in d_sorted.
int run_filter7(sbg_handle *h, int part, int nparts, uint32_t *count_out) {
  int rc;
  int max_warps = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    if ((rc = reset_ctl(h)) != SBG_OK) return rc;
    cudaEventRecord(h->ev[0], h->stream);
    if (use_position_major(h)) {
      if ((rc = launch_filter7_pm(h, part, nparts, attempt > 0)) != SBG_OK) return rc;
    } else if ((rc = launch_sweep<5>(h, part, nparts, max_warps)) != SBG_OK) {
      return rc;
    }
    cudaEventRecord(h->ev[1], h->stream);
    if ((rc = fetch_ctl(h)) != SBG_OK) return rc;
    h->ms[1] = elapsed(h, 0, 1);
    if (!h->h_ctl->overflow) break;
    if (attempt == 1) {
      return fail(h, SBG_ERR_OVERFLOW, "7-LUT hit buffer (%zu entries) overflowed", h->hits_cap);
    }
    // A prefix stops contributing once it has emitted SBG_LIST_CAP hits (checked between chunks of
    // 32 lanes x <= 500 gates), and prefixes are handed out in order: with w warps in flight the
    // buffer needs at most (w + 1) * kPerPrefixMax entries.
    max_warps = (int)std::max<size_t>(1, h->hits_cap / kPerPrefixMax - 1);
  }
  h->swept = h->h_ctl->swept;
  const size_t total = (size_t)h->h_ctl->hit_count;
  uint32_t keep = 0;
  h->ms[2] = 0.f;
  h->sort_pending = false;
  if (total > 0) {
    cudaEventRecord(h->ev[2], h->stream);
    if ((rc = sort_hits(h, h->d_hits, h->d_sorted, total)) != SBG_OK) return rc;
    cudaEventRecord(h->ev[3], h->stream);
    h->sort_pending = true;
    keep = (uint32_t)std::min<size_t>(total, SBG_LIST_CAP);
  }
  *count_out = keep;
  return SBG_OK;
}

void build_params7(sbg_handle *h, const uint8_t *outer_order, const uint8_t *middle_order) {
  DevParams7 *p = h->h_par7;
  for (int pos = 0; pos < 256; pos++) {
    p->pos_outer[outer_order[pos]] = (uint8_t)pos;
    p->pos_middle[middle_order[pos]] = (uint8_t)pos;
  }
}


// h->h_par7 holds the two inverse permutations (build_params7); they go to the device as kernel
// arguments.  reset: also reset the control words (first kernel of a call).
int launch_prepare7(sbg_handle *h, bool reset) {
  Pos512 pos;
  memcpy(pos.outer, h->h_par7->pos_outer, 256);
  memcpy(pos.middle, h->h_par7->pos_middle, 256);
  k_prepare7<<<1, 1024, 0, h->stream>>>(h->d_par7, h->d_ctl, pos, reset ? 1 : 0);
  h->launches++;
  SBG_CUDA(h, cudaGetLastError());
  return SBG_OK;
}

int run_decomp7(sbg_handle *h, int part, int nparts, const uint8_t *outer_order,
    const uint8_t *middle_order, uint64_t *key) {
  int rc;
  *key = SBG_KEY_NONE;
  h->ms[3] = 0.f;
  if (h->list_count == 0) return SBG_OK;
  build_params7(h, outer_order, middle_order);
  if ((rc = launch_prepare7(h, true)) != SBG_OK) return rc;
  cudaEventRecord(h->ev[4], h->stream);
  if ((rc = launch_decomp7(h, part, nparts)) != SBG_OK) return rc;
  cudaEventRecord(h->ev[5], h->stream);
  if ((rc = fetch_ctl(h)) != SBG_OK) return rc;
  h->ms[3] = elapsed(h, 4, 5);
  *key = h->h_ctl->best;
  return SBG_OK;
}

int launch_decomp5(sbg_handle *h) {
  const int n = h->n;
#define SBG_LAUNCH_D5(NWV)                                                                     \
  {                                                                                            \
    const size_t smem = decomp_smem<NWV>(n);                                                   \
    k_decomp5<NWV><<<2 * h->sm_count, kThreads, smem, h->stream>>>(h->d_prob, h->d_ctl,        \
        h->d_pos5, h->d_hits, h->d_tab);                                                       \
  }
  switch (h->nw) {
    case 1: SBG_LAUNCH_D5(1) break;
    case 2: SBG_LAUNCH_D5(2) break;
    case 4: SBG_LAUNCH_D5(4) break;
    default: SBG_LAUNCH_D5(8) break;
  }
#undef SBG_LAUNCH_D5
  h->launches++;
  SBG_CUDA(h, cudaGetLastError());
  return SBG_OK;
}

// search_5lut on this part.  Small searches (the bulk of a real run) use two kernels -- sweep that
// only records feasible tuples, then one warp per recorded tuple -- so that the decomposition of
// several feasible tuples met by one warp is not serialised; large ones use the fused kernel, whose
// ordered early exit matters there.  Either way one host synchronisation.
int run_search5(sbg_handle *h, int part, int nparts, const uint8_t *func_order, uint64_t *key) {
  int rc;
  const uint64_t two_kernel_max = 4000000;  // C(n,5) up to n = 52
  bool two = h->opt_search5 != 0 ? h->opt_search5 == 2 : h_binom[h->n][5] <= two_kernel_max;
  Pos256 pos5;
  for (int pos = 0; pos < 256; pos++) pos5.b[func_order[pos]] = (uint8_t)pos;
  for (;;) {
    k_begin5<<<1, 256, 0, h->stream>>>(h->d_ctl, h->d_pos5, pos5);
    h->launches++;
    cudaEventRecord(h->ev[6], h->stream);
    if ((rc = launch_sweep<3>(h, part, nparts, 0, two)) != SBG_OK) return rc;
    if (two && (rc = launch_decomp5(h)) != SBG_OK) return rc;
    cudaEventRecord(h->ev[7], h->stream);
    if ((rc = fetch_ctl(h)) != SBG_OK) return rc;
    if (two && h->h_ctl->overflow != 0) {
      two = false;   // more feasible tuples than the buffer holds: let the fused kernel do it
      continue;
    }
    break;
  }
  h->ms[0] = elapsed(h, 6, 7);
  h->swept = h->h_ctl->swept;
  h->feasible = two ? h->h_ctl->hit_count : h->h_ctl->feasible;
  *key = h->h_ctl->best;
  return SBG_OK;
}

bool valid_order(const uint8_t *order) {
  if (order == nullptr) return false;
  bool seen[256] = {false};
  for (int i = 0; i < 256; i++) {
    if (seen[order[i]]) return false;
    seen[order[i]] = true;
  }
  return true;
}

}  // namespace

// ---- C ABI -------------------------------------------------------------------------------------

extern "C" {

int sbg_plan_tickets(int width, int prefix_gates, int n, uint32_t excluded, int mode,
    uint64_t waves, uint64_t *out) {
  build_host_tables();
  if (out == nullptr || n < width || n > SBG_MAX_GATES) return SBG_ERR_ARG;
  ChunkPlan pl;
  uint64_t total = 0;
  if (width == 7 && prefix_gates == 4 && n >= 8) {
    pl = plan_chunks_mode<4, 7>(n, excluded, mode, waves, h_binom[n - 5][2]);
    total = h_binom[n - 3][4];
  } else if (width == 7 && prefix_gates == 5 && n >= 8) {
    pl = plan_chunks_mode<5, 7>(n, excluded, mode, waves, (uint64_t)(n - 6));
    total = h_binom[n - 2][5];
  } else if (width == 5 && prefix_gates == 3 && n >= 8) {
    pl = plan_chunks_mode<3, 5>(n, excluded, mode, waves, h_binom[n - 3][2]);
    total = h_binom[n - 2][3];
  } else {
    return SBG_ERR_ARG;
  }
  out[0] = pl.items;
  out[1] = (uint64_t)pl.chunks;
  out[2] = pl.items == 0 ? 0 : (pl.all ? total : pl.t_offset);
  out[3] = total;
  return SBG_OK;
}

int sbg_ordering_row(int width, int k, int *row) {
  build_host_tables();
  if (row == nullptr) return SBG_ERR_ARG;
  if (width == 5 && k >= 0 && k < 10) {
    for (int i = 0; i < 5; i++) row[i] = h_rows5[k][i];
    return SBG_OK;
  }
  if (width == 7 && k >= 0 && k < 70) {
    for (int i = 0; i < 7; i++) row[i] = h_rows7[k][i];
    return SBG_OK;
  }
  return SBG_ERR_ARG;
}

void sbg_lut_table(uint8_t func, const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    uint64_t *out) {
  for (int v = 0; v < 4; v++) {
    uint64_t r = 0;
    for (int m = 0; m < 8; m++) {
      if ((func >> m) & 1) {
        r |= ((m & 4) ? in1[v] : ~in1[v]) & ((m & 2) ? in2[v] : ~in2[v])
            & ((m & 1) ? in3[v] : ~in3[v]);
      }
    }
    out[v] = r;
  }
}

int sbg_solve_inner(const uint64_t *in1, const uint64_t *in2, const uint64_t *in3,
    const uint64_t *target, const uint64_t *mask, uint8_t *func, uint8_t *seen) {
  uint8_t f = 0, s = 0;
  for (int cell = 0; cell < 8; cell++) {
    uint64_t ones = 0, zeros = 0;
    for (int v = 0; v < 4; v++) {
      const uint64_t in_cell = ((cell & 4) ? in1[v] : ~in1[v]) & ((cell & 2) ? in2[v] : ~in2[v])
          & ((cell & 1) ? in3[v] : ~in3[v]) & mask[v];
      ones |= in_cell & target[v];
      zeros |= in_cell & ~target[v];
    }
    if (ones != 0 && zeros != 0) return 0;
    if (ones != 0) f |= (uint8_t)(1u << cell);
    if ((ones | zeros) != 0) s |= (uint8_t)(1u << cell);
  }
  *func = f;
  *seen = s;
  return 1;
}

int sbg_create(sbg_handle **out, int device) {
  if (out == nullptr) return SBG_ERR_ARG;
  *out = nullptr;
  build_host_tables();
  sbg_handle *h = new sbg_handle();
  *out = h;  // returned even on failure so the caller can read the error text
  h->device = device;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    return fail(h, SBG_ERR_CUDA, "no CUDA device available (%s); sboxgates_b200 has no CPU path",
        e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (device < 0 || device >= ndev) return fail(h, SBG_ERR_ARG, "device %d out of range", device);
  const bool timing = getenv("SBG_DEBUG_TIMING") != nullptr;
  auto stamp = [&](const char *what) {
    static double last = 0.0;
    if (!timing) return;
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    const double t = ts.tv_sec + 1e-9 * ts.tv_nsec;
    if (last != 0.0) fprintf(stderr, "[sbg_create] %-28s %.1f ms\n", what, 1e3 * (t - last));
    last = t;
  };
  stamp("start");
  SBG_CUDA(h, cudaSetDevice(device));
  SBG_CUDA(h, cudaFree(nullptr));
  stamp("context");
  cudaDeviceProp prop;
  SBG_CUDA(h, cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  SBG_CUDA(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  for (int i = 0; i < 8; i++) SBG_CUDA(h, cudaEventCreate(&h->ev[i]));

  stamp("stream+events");
  SBG_CUDA(h, cudaMemcpyToSymbol(c_binom, h_binom, sizeof(h_binom)));
  stamp("first symbol (module load)");
  {
    // search5: lane = u<<2 | v2, canonical cell bit of slot s is 4-s.
    uint8_t src5[10][32];
    for (int k = 0; k < 10; k++) {
      const int *o = h_rows5[k];
      for (int lane = 0; lane < 32; lane++) {
        const int u = lane >> 2, v = lane & 3;
        int c = 0;
        c |= ((u >> 2) & 1) << (4 - o[0]);
        c |= ((u >> 1) & 1) << (4 - o[1]);
        c |= (u & 1) << (4 - o[2]);
        c |= ((v >> 1) & 1) << (4 - o[3]);
        c |= (v & 1) << (4 - o[4]);
        src5[k][lane] = (uint8_t)c;
      }
    }
    DevTables host_tab;
    memcpy(host_tab.src5, src5, sizeof(src5));

    // decomp7: group the 70 rows by outer triple; canonical cell bit of slot s (tuple_summary):
    // a..e -> 4..0, f -> 6, g -> 5.
    static const int cb[7] = {4, 3, 2, 1, 0, 6, 5};
    uint32_t src7[25][32];
    uint8_t first_k[25], nrows[25], row_b[70];
    int nj = 0;
    for (int k = 0; k < 70;) {
      const int *o = h_rows7[k];
      int rows = 1;
      while (k + rows < 70 && h_rows7[k + rows][0] == o[0] && h_rows7[k + rows][1] == o[1]
          && h_rows7[k + rows][2] == o[2]) {
        rows++;
      }
      int rest[4], r = 0;
      for (int s = 0; s < 7; s++) {
        if (s != o[0] && s != o[1] && s != o[2]) rest[r++] = s;
      }
      for (int i = 0; i < rows; i++) {
        const int gslot = h_rows7[k + i][6];
        int m = 0;
        while (rest[m] != gslot) m++;
        row_b[k + i] = (uint8_t)(3 - m);
      }
      for (int lane = 0; lane < 32; lane++) {
        const int u0 = lane >> 4, v4 = lane & 15;
        uint32_t packed = 0;
        for (int t4 = 0; t4 < 4; t4++) {
          int c = 0;
          c |= ((t4 >> 1) & 1) << cb[o[0]];
          c |= (t4 & 1) << cb[o[1]];
          c |= u0 << cb[o[2]];
          for (int m = 0; m < 4; m++) c |= ((v4 >> (3 - m)) & 1) << cb[rest[m]];
          packed |= (uint32_t)c << (8 * t4);
        }
        src7[nj][lane] = packed;
      }
      first_k[nj] = (uint8_t)k;
      nrows[nj] = (uint8_t)rows;
      nj++;
      k += rows;
    }
    if (nj != 25) return fail(h, SBG_ERR_STATE, "internal: %d outer triples (expected 25)", nj);
    memcpy(host_tab.src7, src7, sizeof(src7));
    SBG_CUDA(h, cudaMalloc(&h->d_tab, sizeof(DevTables)));
    SBG_CUDA(h, cudaMemcpy(h->d_tab, &host_tab, sizeof(DevTables), cudaMemcpyHostToDevice));
    SBG_CUDA(h, cudaMemcpyToSymbol(c_j_first_k, first_k, sizeof(first_k)));
    SBG_CUDA(h, cudaMemcpyToSymbol(c_j_rows, nrows, sizeof(nrows)));
    SBG_CUDA(h, cudaMemcpyToSymbol(c_row_b, row_b, sizeof(row_b)));
  }

  stamp("constant tables");
  SBG_CUDA(h, cudaMalloc(&h->d_slots, sizeof(DevProblem) * kSlots));
  h->d_prob = h->d_slots;
  h->slots = new sbg_handle::HostProblem[kSlots];
  SBG_CUDA(h, cudaMallocHost(&h->h_prob, sizeof(DevProblem)));
  SBG_CUDA(h, cudaMalloc(&h->d_call, sizeof(sbg_handle::DevCall)));
  SBG_CUDA(h, cudaMallocHost(&h->h_call, sizeof(sbg_handle::DevCall)));
  h->h_ctl = &h->h_call->ctl;
  h->d_par7 = &h->d_call->par;
  h->h_par7 = &h->h_call->par;
  SBG_CUDA(h, cudaMallocHost(&h->h_ctl_out, kCtlHeaderBytes + kHeadEntries * sizeof(uint64_t)));
  h->h_head = reinterpret_cast<uint64_t *>(reinterpret_cast<char *>(h->h_ctl_out) + kCtlHeaderBytes);
  SBG_CUDA(h, cudaMalloc(&h->d_pos5, 256));
  stamp("small buffers + pinned");
  if (getenv("SBG_BATCH") != nullptr) h->opt_batch = atoi(getenv("SBG_BATCH"));
  if (getenv("SBG_PM_PREFIX") != nullptr) h->opt_pm_prefix = atoi(getenv("SBG_PM_PREFIX"));
  if (getenv("SBG_FILTER") != nullptr) h->opt_filter = strcmp(getenv("SBG_FILTER"), "sweep") == 0;
  if (getenv("SBG_HEAD") != nullptr) h->opt_head = std::max(0, std::min(2, atoi(getenv("SBG_HEAD"))));
  if (getenv("SBG_SHIFT") != nullptr) h->opt_shift = atoi(getenv("SBG_SHIFT")) != 0;
  if (getenv("SBG_HEAD_WAVES") != nullptr) h->opt_head_waves = atoi(getenv("SBG_HEAD_WAVES"));
  if (getenv("SBG_SEARCH5") != nullptr) {
    h->opt_search5 = strcmp(getenv("SBG_SEARCH5"), "two") == 0 ? 2 : 1;
  }
  const char *cap_env = getenv("SBG_HITS_CAP");
  h->hits_cap = cap_env != nullptr ? (size_t)strtoull(cap_env, nullptr, 10) : ((size_t)32 << 20);
  if (h->hits_cap < 3 * kPerPrefixMax) h->hits_cap = 3 * kPerPrefixMax;
  SBG_CUDA(h, cudaMalloc(&h->d_hits, h->hits_cap * sizeof(uint64_t)));
  SBG_CUDA(h, cudaMalloc(&h->d_sorted_block, kCtlHeaderBytes + h->hits_cap * sizeof(uint64_t)));
  h->d_ctl = reinterpret_cast<DevCtl *>(h->d_sorted_block);
  h->d_sorted = reinterpret_cast<uint64_t *>(h->d_sorted_block + kCtlHeaderBytes);
  h->cub_bytes = 0;
  SBG_CUDA(h, cub::DeviceRadixSort::SortKeys(nullptr, h->cub_bytes, h->d_hits, h->d_sorted,
      (int)h->hits_cap, 0, 63, h->stream));
  SBG_CUDA(h, cudaMalloc(&h->d_cub, h->cub_bytes));
  SBG_CUDA(h, cudaStreamSynchronize(h->stream));
  stamp("hit buffers + cub temp");
  return SBG_OK;
}

void sbg_destroy(sbg_handle *h) {
  if (h == nullptr) return;
  if (h->own_stream != nullptr) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->d_slots); cudaFreeHost(h->h_prob);
    cudaFree(h->d_call); cudaFreeHost(h->h_call);
    cudaFreeHost(h->h_ctl_out);
    cudaFree(h->d_pos5);
    cudaFree(h->d_tab);
    cudaFree(h->d_hits); cudaFree(h->d_sorted_block); cudaFree(h->d_cub);
    for (int i = 0; i < 8; i++) cudaEventDestroy(h->ev[i]);
    cudaStreamDestroy(h->own_stream);
  }
  delete[] h->slots;
  delete h;
}

const char *sbg_last_error(const sbg_handle *h) { return h != nullptr ? h->err : "null handle"; }

int sbg_set_stream(sbg_handle *h, void *cuda_stream) {
  if (h == nullptr) return SBG_ERR_ARG;
  h->stream = cuda_stream != nullptr ? (cudaStream_t)cuda_stream : h->own_stream;
  return SBG_OK;
}

uint64_t sbg_launch_count(const sbg_handle *h) { return h != nullptr ? h->launches : 0; }

float sbg_last_kernel_ms(const sbg_handle *h, int which) {
  if (h == nullptr || which < 0 || which > 3) return 0.f;
  if (which == 2 && h->sort_pending) {
    // The sort is timed lazily: its end event has completed by the time any result was fetched.
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) != cudaSuccess) {
      (void)cudaGetLastError();
      ms = 0.f;
    }
    const_cast<sbg_handle *>(h)->ms[2] = ms;
    const_cast<sbg_handle *>(h)->sort_pending = false;
  }
  return h->ms[which];
}

int sbg_use_problem(sbg_handle *h, int slot) {
  if (h == nullptr) return SBG_ERR_ARG;
  if (slot < 0 || slot >= kSlots) return fail(h, SBG_ERR_ARG, "slot %d out of range", slot);
  sbg_handle::HostProblem &hp = h->slots[slot];
  if (!hp.ready) return fail(h, SBG_ERR_STATE, "slot %d holds no problem", slot);
  h->d_prob = h->d_slots + slot;
  h->cur_slot = slot;
  h->tables = hp.tables;
  h->target = hp.target;
  h->mask = hp.mask;
  h->n = hp.n;
  h->nw = hp.nw;
  h->inmask = hp.inmask;
  h->problem_ready = true;
  h->list_ready = false;
  h->list_count = 0;
  return SBG_OK;
}

int sbg_stage_problem(sbg_handle *h, int slot, const uint64_t *tables, int n,
    const uint64_t *target, const uint64_t *mask, const int8_t *inbits) {
  if (h == nullptr) return SBG_ERR_ARG;
  if (slot < 0 || slot >= kSlots) return fail(h, SBG_ERR_ARG, "slot %d out of range", slot);
  if (tables == nullptr || target == nullptr || mask == nullptr || inbits == nullptr) {
    return fail(h, SBG_ERR_ARG, "null argument");
  }
  if (n < 1 || n > SBG_MAX_GATES) return fail(h, SBG_ERR_ARG, "n = %d out of range", n);
  SBG_CUDA(h, cudaSetDevice(h->device));
  sbg_handle::HostProblem &hp = h->slots[slot];
  uint32_t inmask = 0;
  for (int k = 0; k < 8 && inbits[k] != -1; k++) {
    if (inbits[k] >= 0 && inbits[k] < 8) inmask |= 1u << inbits[k];
  }
  // lut_search calls search_5lut and then search_7lut on the same state (lut.c:553,593): the second
  // upload is skipped when nothing changed
  if (hp.ready && hp.n == n && hp.inmask == inmask && memcmp(hp.target, target, 32) == 0
      && memcmp(hp.mask, mask, 32) == 0 && memcmp(hp.tables, tables, (size_t)n * 32) == 0) {
    return SBG_OK;
  }
  // The pinned staging block may still be in flight from the previous upload.
  SBG_CUDA(h, cudaStreamSynchronize(h->stream));
  hp.inmask = inmask;
  memcpy(hp.tables, tables, (size_t)n * 32);
  memcpy(hp.target, target, 32);
  memcpy(hp.mask, mask, 32);
  hp.n = n;
  const int m = popcount256(mask);
  hp.nw = m <= 32 ? 1 : m <= 64 ? 2 : m <= 128 ? 4 : 8;
  hp.ready = true;

  DevProblem *p = h->h_prob;
  memset(p, 0, offsetof(DevProblem, xr));
  p->n = n;
  p->nw = hp.nw;
  uint32_t cm[8], ct[8], tmp[8];
  const uint64_t ones[4] = {~0ull, ~0ull, ~0ull, ~0ull};
  compress_table(ones, mask, cm);
  compress_table(target, mask, ct);
  for (int w = 0; w < 8; w++) {
    p->M[w] = cm[w];
    p->T[w] = ct[w] & cm[w];
  }
  for (int g = 0; g < n; g++) {
    compress_table(tables + 4 * g, mask, tmp);
    for (int w = 0; w < 8; w++) p->tabs[w][g] = tmp[w] & cm[w];
  }
  p->inmask = inmask;
  p->m = m;
  // The position-major rows (DevProblem::xr) are derived on the device, and only when a 7-LUT
  // search asks for them (ensure_rows): most states of a run never get that far.
  hp.rows_ready = false;
  SBG_CUDA(h, cudaMemcpyAsync(h->d_slots + slot, p, offsetof(DevProblem, xr), cudaMemcpyHostToDevice,
      h->stream));
  return SBG_OK;
}

int sbg_load_problem(sbg_handle *h, const uint64_t *tables, int n, const uint64_t *target,
    const uint64_t *mask, const int8_t *inbits) {
  int rc = sbg_stage_problem(h, 0, tables, n, target, mask, inbits);
  if (rc != SBG_OK) return rc;
  return sbg_use_problem(h, 0);
}

int sbg_search5_part(sbg_handle *h, int part, int nparts, const uint8_t *func_order,
    uint64_t *key) {
  if (h == nullptr || key == nullptr) return SBG_ERR_ARG;
  if (!h->problem_ready) return fail(h, SBG_ERR_STATE, "no problem loaded");
  if (h->n < 5) return fail(h, SBG_ERR_ARG, "search_5lut needs n >= 5 (lut.c:119)");
  if (nparts < 1 || part < 0 || part >= nparts) return fail(h, SBG_ERR_ARG, "bad part %d/%d", part, nparts);
  if (!valid_order(func_order)) return fail(h, SBG_ERR_ARG, "func_order is not a permutation");
  SBG_CUDA(h, cudaSetDevice(h->device));
  return run_search5(h, part, nparts, func_order, key);
}

int sbg_finish5(sbg_handle *h, uint64_t key, const uint8_t *func_order, sbg_result *res) {
  if (h == nullptr || res == nullptr || func_order == nullptr) return SBG_ERR_ARG;
  memset(res, 0, sizeof(*res));
  res->key = key;
  res->tuples_feasible = h->feasible;
  res->tuples_swept = h->swept;
  if (key == SBG_KEY_NONE) return SBG_OK;
  const uint64_t rank = key >> 12;
  const int k = (int)((key >> 8) & 0xf);
  const int pos = (int)(key & 0xff);
  if (k >= 10 || rank >= h_binom[h->n][5]) return fail(h, SBG_ERR_STATE, "corrupt 5-LUT key");
  uint16_t comb[5];
  unrank_combination(rank, h->n, 5, comb);
  const int *o = h_rows5[k];
  for (int i = 0; i < 5; i++) res->gates[i] = comb[o[i]];
  res->found = 1;
  res->ordering = k;
  res->pos_outer = pos;
  res->func_outer = func_order[pos];
  res->index = rank;
  uint64_t t_outer[4];
  sbg_lut_table(res->func_outer, h->tables[res->gates[0]], h->tables[res->gates[1]],
      h->tables[res->gates[2]], t_outer);
  if (!sbg_solve_inner(t_outer, h->tables[res->gates[3]], h->tables[res->gates[4]], h->target,
      h->mask, &res->func_inner, &res->inner_seen)) {
    return fail(h, SBG_ERR_STATE, "internal: winning 5-LUT key does not decompose");
  }
  return SBG_OK;
}

int sbg_search5(sbg_handle *h, const uint8_t *func_order, sbg_result *res) {
  uint64_t key = SBG_KEY_NONE;
  int rc = sbg_search5_part(h, 0, 1, func_order, &key);
  if (rc != SBG_OK) return rc;
  return sbg_finish5(h, key, func_order, res);
}

int sbg_filter7_part(sbg_handle *h, int part, int nparts, uint64_t *list, int *count) {
  if (h == nullptr || count == nullptr) return SBG_ERR_ARG;
  if (!h->problem_ready) return fail(h, SBG_ERR_STATE, "no problem loaded");
  if (h->n < 7) return fail(h, SBG_ERR_ARG, "search_7lut needs n >= 7 (lut.c:259)");
  if (nparts < 1 || part < 0 || part >= nparts) return fail(h, SBG_ERR_ARG, "bad part %d/%d", part, nparts);
  SBG_CUDA(h, cudaSetDevice(h->device));
  uint32_t keep = 0;
  int rc = run_filter7(h, part, nparts, &keep);
  if (rc != SBG_OK) return rc;
  *count = (int)keep;
  h->list_ready = false;
  if (list != nullptr && keep > 0) {
    SBG_CUDA(h, cudaMemcpyAsync(list, h->d_sorted, (size_t)keep * sizeof(uint64_t),
        cudaMemcpyDeviceToHost, h->stream));
    SBG_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  // The part's own sorted list stays on the device; when it is the whole space (nparts == 1) it
  // IS the list, and phase 2 may follow without sbg_set_list7().
  h->d_list = h->d_sorted;
  h->list_count = keep;
  h->list_ready = nparts == 1;
  return SBG_OK;
}

int sbg_set_list7(sbg_handle *h, const uint64_t *list, int count) {
  if (h == nullptr || count < 0 || (count > 0 && list == nullptr)) return SBG_ERR_ARG;
  if ((size_t)count > h->hits_cap) return fail(h, SBG_ERR_ARG, "list of %d entries too long", count);
  SBG_CUDA(h, cudaSetDevice(h->device));
  h->list_count = 0;
  if (count > 0) {
    SBG_CUDA(h, cudaMemcpyAsync(h->d_hits, list, (size_t)count * sizeof(uint64_t),
        cudaMemcpyHostToDevice, h->stream));
    int rc = sort_hits(h, h->d_hits, h->d_sorted, (size_t)count);
    if (rc != SBG_OK) return rc;
    SBG_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->d_list = h->d_sorted;
  h->list_count = (uint32_t)std::min<int>(count, SBG_LIST_CAP);
  h->list_ready = true;
  return SBG_OK;
}

int sbg_decomp7_part(sbg_handle *h, int part, int nparts, const uint8_t *outer_order,
    const uint8_t *middle_order, uint64_t *key) {
  if (h == nullptr || key == nullptr) return SBG_ERR_ARG;
  if (!h->problem_ready) return fail(h, SBG_ERR_STATE, "no problem loaded");
  if (!h->list_ready) return fail(h, SBG_ERR_STATE, "no 7-LUT list installed");
  if (nparts < 1 || part < 0 || part >= nparts) return fail(h, SBG_ERR_ARG, "bad part %d/%d", part, nparts);
  if (!valid_order(outer_order) || !valid_order(middle_order)) {
    return fail(h, SBG_ERR_ARG, "function order is not a permutation");
  }
  SBG_CUDA(h, cudaSetDevice(h->device));
  return run_decomp7(h, part, nparts, outer_order, middle_order, key);
}

static int finish7_impl(sbg_handle *h, uint64_t key, const uint8_t *outer_order,
    const uint8_t *middle_order, sbg_result *res, const uint64_t *head, size_t head_len) {
  if (h == nullptr || res == nullptr || outer_order == nullptr || middle_order == nullptr) {
    return SBG_ERR_ARG;
  }
  memset(res, 0, sizeof(*res));
  res->key = key;
  res->tuples_feasible = h->list_count;
  res->tuples_swept = h->swept;
  if (key == SBG_KEY_NONE) return SBG_OK;
  const uint64_t idx = key >> 23;
  const int k = (int)((key >> 16) & 0x7f);
  const int po = (int)((key >> 8) & 0xff);
  const int pm = (int)(key & 0xff);
  if (idx >= h->list_count || k >= 70) return fail(h, SBG_ERR_STATE, "corrupt 7-LUT key");
  SBG_CUDA(h, cudaSetDevice(h->device));
  uint64_t pair[2] = {0, 0};
  const size_t first = idx > 0 ? idx - 1 : 0;
  if (head != nullptr && idx < head_len) {   // already on the host
    pair[0] = head[first];
    pair[1] = head[idx];
  } else {
    SBG_CUDA(h, cudaMemcpyAsync(pair, h->d_list + first, (idx > 0 ? 2 : 1) * sizeof(uint64_t),
        cudaMemcpyDeviceToHost, h->stream));
    SBG_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  const uint64_t cur = idx > 0 ? pair[1] : pair[0];
  uint16_t t[7];
  for (int i = 0; i < 7; i++) t[i] = (uint16_t)((cur >> (9 * (6 - i))) & 0x1ff);
  const int *o = h_rows7[k];
  for (int i = 0; i < 7; i++) res->gates[i] = t[o[i]];
  res->found = 1;
  res->ordering = k;
  res->pos_outer = po;
  res->pos_middle = pm;
  res->func_outer = outer_order[po];
  res->func_middle = middle_order[pm];
  res->index = idx;
  // lut.c:432-435 quirk (see k_decomp7): rows 0-3 may have been evaluated with the previous
  // tuple's outer tables; the solved inner function must come from the same tables.
  uint16_t outer_a = res->gates[0];
  if (idx > 0 && k < 4 && t[0] == 0) {
    const uint64_t prev = pair[0];
    if ((uint16_t)((prev >> 9) & 0x1ff) == t[1] && (uint16_t)(prev & 0x1ff) == t[2]) {
      outer_a = (uint16_t)((prev >> 45) & 0x1ff);
      res->stale_outer = 1;
    }
  }
  uint64_t t_outer[4], t_middle[4];
  sbg_lut_table(res->func_outer, h->tables[outer_a], h->tables[res->gates[1]],
      h->tables[res->gates[2]], t_outer);
  sbg_lut_table(res->func_middle, h->tables[res->gates[3]], h->tables[res->gates[4]],
      h->tables[res->gates[5]], t_middle);
  if (!sbg_solve_inner(t_outer, t_middle, h->tables[res->gates[6]], h->target, h->mask,
      &res->func_inner, &res->inner_seen)) {
    return fail(h, SBG_ERR_STATE, "internal: winning 7-LUT key does not decompose");
  }
  return SBG_OK;
}

int sbg_finish7(sbg_handle *h, uint64_t key, const uint8_t *outer_order,
    const uint8_t *middle_order, sbg_result *res) {
  return finish7_impl(h, key, outer_order, middle_order, res, nullptr, 0);
}

// Whole search_7lut on one device.  Fast path: upload (control words + parameters, one copy) ->
// phase 1 -> on-device sort of a short hit list -> phase 2 -> one read-back, i.e. a single host
// synchronisation per call.  Long lists (more than kSmallSort hits) take the step-by-step path
// with CUB's radix sort.
int sbg_search7(sbg_handle *h, const uint8_t *outer_order, const uint8_t *middle_order,
    sbg_result *res) {
  if (h == nullptr || res == nullptr) return SBG_ERR_ARG;
  if (!h->problem_ready) return fail(h, SBG_ERR_STATE, "no problem loaded");
  if (h->n < 7) return fail(h, SBG_ERR_ARG, "search_7lut needs n >= 7 (lut.c:259)");
  if (!valid_order(outer_order) || !valid_order(middle_order)) {
    return fail(h, SBG_ERR_ARG, "function order is not a permutation");
  }
  SBG_CUDA(h, cudaSetDevice(h->device));
  int rc;
  build_params7(h, outer_order, middle_order);
  if ((rc = launch_prepare7(h, true)) != SBG_OK) return rc;
  cudaEventRecord(h->ev[0], h->stream);
  if (use_position_major(h)) {
    if ((rc = launch_filter7_pm(h, 0, 1, false)) != SBG_OK) return rc;
  } else if ((rc = launch_sweep<5>(h, 0, 1, 0)) != SBG_OK) {
    return rc;
  }
  cudaEventRecord(h->ev[1], h->stream);
  k_sort_small<<<1, 1024, 0, h->stream>>>(h->d_hits, h->d_sorted, h->d_ctl,
      (unsigned int)SBG_LIST_CAP);
  h->launches++;
  cudaEventRecord(h->ev[4], h->stream);
  h->d_list = h->d_sorted;
  if ((rc = launch_decomp7(h, 0, 1, true)) != SBG_OK) return rc;
  cudaEventRecord(h->ev[5], h->stream);
  if ((rc = fetch_ctl(h, true)) != SBG_OK) return rc;
  h->ms[1] = elapsed(h, 0, 1);
  h->ms[2] = elapsed(h, 1, 4);
  h->ms[3] = elapsed(h, 4, 5);
  h->sort_pending = false;
  h->swept = h->h_ctl->swept;

  if (h->h_ctl->overflow == 0 && h->h_ctl->sorted_ok != 0) {
    h->list_count = h->h_ctl->list_count;
    h->list_ready = true;
    return finish7_impl(h, h->h_ctl->best, outer_order, middle_order, res, h->h_head,
        std::min<size_t>(kHeadEntries, h->list_count));
  }
  uint32_t keep = 0;
  if (h->h_ctl->overflow != 0) {
    // hit buffer overflowed: redo phase 1 with the bounded-parallelism retry
    if ((rc = run_filter7(h, 0, 1, &keep)) != SBG_OK) return rc;
  } else {
    // long list: phase 1 is done, sort its hits with CUB
    const size_t total = (size_t)h->h_ctl->hit_count;
    cudaEventRecord(h->ev[2], h->stream);
    if ((rc = sort_hits(h, h->d_hits, h->d_sorted, total)) != SBG_OK) return rc;
    cudaEventRecord(h->ev[3], h->stream);
    h->sort_pending = true;
    keep = (uint32_t)std::min<size_t>(total, SBG_LIST_CAP);
  }
  h->d_list = h->d_sorted;
  h->list_count = keep;
  h->list_ready = true;
  uint64_t key = SBG_KEY_NONE;
  if ((rc = run_decomp7(h, 0, 1, outer_order, middle_order, &key)) != SBG_OK) return rc;
  return sbg_finish7(h, key, outer_order, middle_order, res);
}

}  // extern "C"
