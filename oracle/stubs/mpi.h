/* oracle/stubs/mpi.h -- TEST INFRASTRUCTURE ONLY: a one-process stand-in for the MPI calls the reference's
 * constraint path makes (MPI is absent from this image), so that the reference's own sources compile in place.
 * Every collective on the single rank is a copy of the send buffer into the receive buffer. */
#ifndef ALENS_ORACLE_MPI_STUB_H
#define ALENS_ORACLE_MPI_STUB_H
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int MPI_Comm;
typedef int MPI_Datatype; /* = size in bytes */
typedef int MPI_Op;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0
#define MPI_IN_PLACE ((void *)-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_SIGNED_CHAR 1
#define MPI_UNSIGNED_CHAR 1
#define MPI_SHORT 2
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_LONG_LONG 8
#define MPI_LONG_LONG_INT 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_DOUBLE 8
#define MPI_C_BOOL 1
#define MPI_CXX_BOOL 1
#define MPI_DATATYPE_NULL 0
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_LOR 3
#define MPI_LAND 4
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_MULTIPLE 3
#define MPI_MAX_PROCESSOR_NAME 64

static inline void alens_mpi_copy_(const void *s, void *r, int n, MPI_Datatype t) {
    if (s != MPI_IN_PLACE && s != r && n > 0) memcpy(r, s, (size_t)n * (size_t)t);
}
static inline int MPI_Init(int *a, char ***b) { (void)a; (void)b; return 0; }
static inline int MPI_Init_thread(int *a, char ***b, int req, int *prov) { (void)a; (void)b; if (prov) *prov = req; return 0; }
static inline int MPI_Initialized(int *f) { *f = 1; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline double MPI_Wtime(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static inline int MPI_Get_processor_name(char *n, int *l) { strcpy(n, "localhost"); *l = 9; return 0; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) { (void)o; (void)c; alens_mpi_copy_(s, r, n, t); return 0; }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, int root, MPI_Comm c) { (void)o; (void)c; (void)root; alens_mpi_copy_(s, r, n, t); return 0; }
static inline int MPI_Scan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) { (void)o; (void)c; alens_mpi_copy_(s, r, n, t); return 0; }
static inline int MPI_Exscan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) { (void)s; (void)r; (void)n; (void)t; (void)o; (void)c; return 0; } /* rank 0: recvbuf undefined */
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
static inline int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c) { (void)nr; (void)tr; (void)c; alens_mpi_copy_(s, r, ns, ts); return 0; }
static inline int MPI_Gather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c) { (void)nr; (void)tr; (void)c; (void)root; alens_mpi_copy_(s, r, ns, ts); return 0; }
static inline int MPI_Allgatherv(const void *s, int ns, MPI_Datatype ts, void *r, const int *nr, const int *displ, MPI_Datatype tr, MPI_Comm c) { (void)nr; (void)c; if (s != MPI_IN_PLACE) memcpy((char *)r + (size_t)displ[0] * tr, s, (size_t)ns * ts); return 0; }
static inline int MPI_Gatherv(const void *s, int ns, MPI_Datatype ts, void *r, const int *nr, const int *displ, MPI_Datatype tr, int root, MPI_Comm c) { (void)root; return MPI_Allgatherv(s, ns, ts, r, nr, displ, tr, c); }
static inline int MPI_Scatterv(const void *s, const int *ns, const int *displ, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c) { (void)nr; (void)tr; (void)root; (void)c; memcpy(r, (const char *)s + (size_t)displ[0] * ts, (size_t)ns[0] * ts); return 0; }
static inline int MPI_Scatter(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c) { (void)nr; (void)tr; (void)root; (void)c; alens_mpi_copy_(s, r, ns, ts); return 0; }
static inline int MPI_Alltoall(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c) { (void)nr; (void)tr; (void)c; alens_mpi_copy_(s, r, ns, ts); return 0; }
static inline int MPI_Alltoallv(const void *s, const int *ns, const int *sd, MPI_Datatype ts, void *r, const int *nr, const int *rd, MPI_Datatype tr, MPI_Comm c) { (void)nr; (void)c; memcpy((char *)r + (size_t)rd[0] * tr, (const char *)s + (size_t)sd[0] * ts, (size_t)ns[0] * ts); return 0; }
static inline int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype *out) { *out = n * t; return 0; }
static inline int MPI_Type_commit(MPI_Datatype *t) { (void)t; return 0; }
static inline int MPI_Type_free(MPI_Datatype *t) { (void)t; return 0; }
static inline int MPI_Waitall(int n, MPI_Request *r, MPI_Status *s) { (void)n; (void)r; (void)s; return 0; }
static inline int MPI_Wait(MPI_Request *r, MPI_Status *s) { (void)r; (void)s; return 0; }
/* point-to-point never happens on one rank */
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *q) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)q; abort(); return 0; }
static inline int MPI_Irecv(void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *q) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)q; abort(); return 0; }
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; abort(); return 0; }
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Status *s) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)s; abort(); return 0; }
static inline int MPI_Get_count(const MPI_Status *s, MPI_Datatype t, int *n) { (void)t; *n = s ? s->count_ : 0; return 0; }
#endif
