/*
  oracle/shim/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).

  A minimal in-process stand-in for the MPI-1 subset that the reference hot
  path touches (src/TMRBase.cpp:42-108, src/TMROctForest.cpp:1824,1941,2018,
  2048,2059,2415,2482,2499,2505,4167,6740).  "Ranks" are threads of one
  process: shim_run(R, fn, arg) starts R threads, each seeing its own rank
  through MPI_Comm_rank.  Collectives rendezvous on a barrier and copy through
  a shared slot table; MPI_Isend is an eager copy into a per-(dst,src) FIFO and
  MPI_Recv blocks on that FIFO.

  This lets the UNMODIFIED reference sources compile and run in a container
  that has no MPI installation.
*/
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef struct {
  int unused;
} MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0

/* builtin datatype handle == its byte size */
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_INT32_T 4
#define MPI_INT16_T 2
#define MPI_CHAR 1
#define MPI_BYTE 1

#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
double MPI_Wtime(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Type_create_struct(int count, const int lens[], const MPI_Aint disp[],
                           const MPI_Datatype types[], MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *t);
int MPI_Type_free(MPI_Datatype *t);
int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf,
                  int rcount, MPI_Datatype rtype, MPI_Comm comm);
int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf,
                 int rcount, MPI_Datatype rtype, MPI_Comm comm);
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag,
              MPI_Comm comm, MPI_Request *req);
int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag,
             MPI_Comm comm, MPI_Status *status);
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status stats[]);

/* Run fn(rank, arg) on nranks threads-as-ranks and join them. */
typedef void (*shim_rank_fn)(int rank, void *arg);
void shim_run(int nranks, shim_rank_fn fn, void *arg);
/* Alternative to shim_run for callers that own their threads (e.g. Python
   threads through ctypes): open a world of nranks, have each thread attach as
   one rank, close the world when all ranks are done. */
void shim_world_begin(int nranks);
void shim_attach(int rank);
void shim_world_end(void);
/* Barrier usable from inside rank bodies. */
void shim_barrier(void);
int shim_rank(void);
int shim_size(void);

#ifdef __cplusplus
}
#endif
#endif
