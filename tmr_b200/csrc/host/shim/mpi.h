/*
  shim/mpi.h -- stand-in for <mpi.h> when the drop-in is built without an MPI
  installation (this image has none).  TMROctForest keeps its MPI_Comm
  constructor argument (reference src/TMROctForest.h:51); in the B200 build
  the communicator only carries (rank, size): one process per GPU, launched by
  torchrun, with the data path on NCCL inside the CUDA layer.  When a real MPI
  is available, drop this directory from the include path.
*/
#ifndef TMR_B200_MPI_SHIM_H
#define TMR_B200_MPI_SHIM_H

#include <math.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0

/* process-wide (rank, size), set once by the launcher glue */
void tmr_b200_set_world(int rank, int size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);

#ifdef __cplusplus
}
#endif
#endif
