/*
  tmr_b200_ext.h -- B200-only entry points next to the flat class binding
  (include/tmr_capi.h).  They have no counterpart in the reference: they give
  a caller that keeps its data on the GPU (or wants whole arrays at once)
  access to what the drop-in TMROctForest holds.  Implemented in
  tmr_b200/csrc/host/tmr_b200_ext.cpp and TMROctant.cpp.
*/
#ifndef TMR_B200_EXT_H
#define TMR_B200_EXT_H

#include "tmrgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the process-wide CUDA context of the drop-in classes (created on first use;
   NULL when no CUDA device is available -- there is no CPU fallback) */
tmrgpu_ctx *tmr_b200_context(void);
/* the device forest behind a tmrc_forest handle (TMROctForest*): the handle
   every tmrgpu_* call of include/tmrgpu.h takes */
tmrgpu_forest *tmr_b200_device_forest(void *forest);
/* attach this process (or thread) as `rank` of `size`: what MPI_Comm_rank/size
   report to the drop-in classes, and the NCCL communicator named by `id`
   (tmrgpu_comm_unique_id).  Collective. */
int tmr_b200_init_world(int rank, int size, const void *id);
/* createInterpolation as ONE hand-off: the whole prolongation in CSR form
   (rows in the order the reference emits its addInterp calls,
   src/TMROctForest.cpp:6683,6775) instead of one virtual call per row --
   the bulk ingest a TACSBVecInterp-like consumer needs.  The
   pointers are borrowed from the fine forest.  Returns the number of rows. */
int tmr_b200_create_interpolation_csr(void *fine, void *coarse, const int **rows,
                                      const int **rowp, const int **cols,
                                      const double **vals, int *nnz);

#ifdef __cplusplus
}
#endif
#endif
