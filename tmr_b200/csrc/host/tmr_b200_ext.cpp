/*
  tmr_b200_ext.cpp -- B200-only extensions next to the flat C binding: access
  to the device forest behind a TMROctForest so that measurement code can drive
  the CUDA layer (include/tmrgpu.h) on device-resident inputs.
*/
#include "TMROctForest.h"
#include "TMRTrilinearVolume.h"
#include "tmr_capi.h"
#include "tmr_b200_ext.h"
#include "tmrgpu.h"

extern "C" {

/* tmrgpu_forest behind a tmrc_forest handle (TMROctForest*) */
tmrgpu_forest *tmr_b200_device_forest(void *forest) {
  return static_cast<TMROctForest *>(forest)->getDeviceForest();
}

/* attach this process (thread) to a world of `size` ranks: sets what
   MPI_Comm_rank/size report to the drop-in classes and connects the CUDA
   context to the NCCL communicator identified by `id`
   (tmrgpu_comm_unique_id).  Collective over the world. */
int tmr_b200_init_world(int rank, int size, const void *id) {
  tmr_b200_set_world(rank, size);
  tmrgpu_ctx *ctx = tmr_b200_context();
  if (!ctx) return 1;
  return tmrgpu_ctx_init_comm(ctx, rank, size, id);
}

/* the whole prolongation at once (TMROctForest::createInterpolationCSR):
   borrowed pointers into arrays owned by the fine forest; returns the rows */
int tmr_b200_create_interpolation_csr(void *fine, void *coarse, const int **rows,
                                      const int **rowp, const int **cols,
                                      const double **vals, int *nnz) {
  const int *rp = NULL;
  const int n = static_cast<TMROctForest *>(fine)->createInterpolationCSR(
      static_cast<TMROctForest *>(coarse), rows, &rp, cols, vals);
  if (rowp) *rowp = rp;
  if (nnz) *nnz = (rp && n > 0) ? rp[n] : 0;
  return n;
}

/* declared in tmr_capi.h: the drop-in's version builds a TMRTrilinearTopology */
int tmrc_set_trilinear_topology(tmrc_forest f, int num_nodes, const int *conn,
                                int num_blocks, const double *xpts) {
  TMROctForest *forest = static_cast<TMROctForest *>(f);
  /* edge and face numbering exactly as setConnectivity derives it */
  TMROctForest *tmp = new TMROctForest(MPI_COMM_SELF);
  tmp->incref();
  tmp->setConnectivity(num_nodes, conn, num_blocks);
  int nb, nf, ne, nn;
  const int *bc, *bfc, *bec, *ids;
  tmp->getConnectivity(&nb, &nf, &ne, &nn, &bc, &bfc, &bec, &ids);
  TMRTrilinearTopology *topo =
      new TMRTrilinearTopology(nn, ne, nf, nb, bc, bec, bfc, xpts);
  tmp->decref();
  forest->setTopology(topo);
  return 0;
}

/* declared in tmr_capi.h */
int tmrc_set_entity_name(tmrc_forest f, int kind, int index, const char *name) {
  TMROctForest *forest = static_cast<TMROctForest *>(f);
  TMRTrilinearTopology *topo =
      dynamic_cast<TMRTrilinearTopology *>(forest->getTopology());
  TMREntity *e = topo ? topo->entity(kind, index) : NULL;
  if (!e) return 1;
  e->setName(name);
  return 0;
}

}  // extern "C"
