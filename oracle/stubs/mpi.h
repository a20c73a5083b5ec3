/* oracle/stubs/mpi.h -- TEST INFRASTRUCTURE, not product code.
 *
 * A single-rank stand-in for <mpi.h> so that the UNMODIFIED reference sources
 * under /root/reference compile with plain gcc (no MPI runtime exists in this
 * image).  It implements exactly the calls the reference makes
 * (sboxgates.c:619-642,791-868,1044-1174; lut.c:121-159,212-238,329-347,
 * 385-393,463-482,664-740) with communicator size 1 / rank 0 semantics:
 * broadcasts and barriers are no-ops, gathers are memcpy, nothing ever arrives
 * on the early-termination channel.  "size == 1, fixed seed" is the
 * well-defined behaviour our parity target is pinned to.
 *
 * An optional fake (rank,size) can be injected with sbgref_set_fake_rank() so
 * the bench can time the slice lut.c:137-149 would hand to rank r of R without
 * any communication.
 */
#ifndef SBG_STUB_MPI_H
#define SBG_STUB_MPI_H

#include <stddef.h>
#include <string.h>

typedef int MPI_Comm;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef int MPI_Datatype;   /* value = element size in bytes (0 for derived types) */
typedef struct { int unused; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_ANY_SOURCE (-1)
#define MPI_REQUEST_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

#define MPI_INT 4
#define MPI_C_BOOL 1
#define MPI_UINT8_T 1
#define MPI_UINT16_T 2
#define MPI_UINT64_T 8

extern int sbgref_fake_rank;
extern int sbgref_fake_size;

static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = sbgref_fake_rank; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = sbgref_fake_size; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
  (void)buf; (void)count; (void)t; (void)root; (void)c; return MPI_SUCCESS;
}
static inline int MPI_Irecv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c,
    MPI_Request *req) {
  (void)buf; (void)count; (void)t; (void)src; (void)tag; (void)c; *req = 1; return MPI_SUCCESS;
}
static inline int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dst, int tag, MPI_Comm c,
    MPI_Request *req) {
  (void)buf; (void)count; (void)t; (void)dst; (void)tag; (void)c; *req = 1; return MPI_SUCCESS;
}
static inline int MPI_Recv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c,
    MPI_Status *st) {
  (void)buf; (void)count; (void)t; (void)src; (void)tag; (void)c; (void)st; return MPI_SUCCESS;
}
static inline int MPI_Test(MPI_Request *req, int *flag, MPI_Status *st) {
  (void)req; (void)st; *flag = 0; return MPI_SUCCESS;
}
static inline int MPI_Iprobe(int src, int tag, MPI_Comm c, int *flag, MPI_Status *st) {
  (void)src; (void)tag; (void)c; (void)st; *flag = 0; return MPI_SUCCESS;
}
static inline int MPI_Cancel(MPI_Request *req) { (void)req; return MPI_SUCCESS; }
static inline int MPI_Wait(MPI_Request *req, MPI_Status *st) {
  (void)st; *req = MPI_REQUEST_NULL; return MPI_SUCCESS;
}
static inline int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *st) {
  (void)st; for (int i = 0; i < n; i++) reqs[i] = MPI_REQUEST_NULL; return MPI_SUCCESS;
}
/* With a faked size > 1 the gathers only ever see this rank's contribution (slot `rank`
   of the count array, offset 0 of the data) -- good enough for timing a slice, never used
   for parity. */
static inline int MPI_Allgather(const void *sb, int sc, MPI_Datatype st, void *rb, int rc,
    MPI_Datatype rt, MPI_Comm c) {
  (void)rc; (void)rt; (void)c;
  if (sbgref_fake_size > 1) {
    memset(rb, 0, (size_t)sc * (size_t)st * (size_t)sbgref_fake_size);
  }
  memcpy(rb, sb, (size_t)sc * (size_t)st);
  return MPI_SUCCESS;
}
static inline int MPI_Allgatherv(const void *sb, int sc, MPI_Datatype st, void *rb, const int *rcs,
    const int *displs, MPI_Datatype rt, MPI_Comm c) {
  (void)rcs; (void)displs; (void)rt; (void)c;
  memcpy(rb, sb, (size_t)sc * (size_t)st);
  return MPI_SUCCESS;
}
static inline int MPI_Type_create_struct(int n, const int *bl, const MPI_Aint *d,
    const MPI_Datatype *t, MPI_Datatype *out) {
  (void)n; (void)bl; (void)d; (void)t; *out = 0; return MPI_SUCCESS;
}
static inline int MPI_Type_create_resized(MPI_Datatype in, MPI_Aint lb, MPI_Aint extent,
    MPI_Datatype *out) {
  (void)in; (void)lb; (void)extent; *out = 0; return MPI_SUCCESS;
}
static inline int MPI_Type_commit(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }

#endif /* SBG_STUB_MPI_H */
