/*
 * mpi.h -- minimal stand-in for <mpi.h>, TEST INFRASTRUCTURE ONLY.
 *
 * The reference (nullike/photoNs-2.0) is an MPI program, and this image has no
 * MPI.  This header plus mpi_shim.c provide exactly the 17 MPI entry points the
 * reference's short-range path uses (list: SURVEY.md Appendix C), implemented
 * over fork() + socketpair() so the UNMODIFIED reference sources can be run at
 * NP = 1..N ranks on one host.  Nothing here is part of the product.
 *
 * Datatypes are encoded as their size in bytes (the reference only ever builds
 * contiguous byte types: src/initial.c:233-243).
 */
#ifndef PN_ORACLE_MPI_SHIM_H
#define PN_ORACLE_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_BYTE   ((MPI_Datatype)1)
#define MPI_CHAR   ((MPI_Datatype)1)
#define MPI_INT    ((MPI_Datatype)4)
#define MPI_FLOAT  ((MPI_Datatype)4)
#define MPI_DOUBLE ((MPI_Datatype)8)
#define MPI_LONG   ((MPI_Datatype)8)
#define MPI_SUM ((MPI_Op)1)
#define MPI_MAX ((MPI_Op)2)
#define MPI_MIN ((MPI_Op)3)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_ANY_TAG (-1)

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm c, int code);
int MPI_Comm_size(MPI_Comm c, int *size);
int MPI_Comm_rank(MPI_Comm c, int *rank);
int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out);
int MPI_Barrier(MPI_Comm c);
int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *t);
int MPI_Type_free(MPI_Datatype *t);
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request *req);
int MPI_Recv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st);
int MPI_Wait(MPI_Request *req, MPI_Status *st);
int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount, MPI_Datatype rt, MPI_Comm c);
int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount, MPI_Datatype rt, MPI_Comm c);
int MPI_Alltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype st,
                  void *rbuf, const int *rcounts, const int *rdispls, MPI_Datatype rt, MPI_Comm c);
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c);

/* shim-only introspection used by the harness (not MPI) */
extern long pn_shim_recv_seq[256];    /* number of MPI_Recv calls seen per tag (tag < 256) */
extern long pn_shim_recv_bytes[256];  /* bytes delivered by the latest MPI_Recv per tag   */
int pn_shim_world_size(void);
int pn_shim_world_rank(void);

#ifdef __cplusplus
}
#endif
#endif
