/*
  shim/TMRTopology.h -- minimal declaration of the CAD topology class the
  forest can be attached to (reference src/TMRTopology.h:381-400).  The CAD
  layer is outside the hot path; in the full TMR tree the real header is used
  instead of this one.
*/
#ifndef TMR_B200_TOPOLOGY_SHIM_H
#define TMR_B200_TOPOLOGY_SHIM_H

#include "TMRBase.h"

class TMRTopology : public TMREntity {
 public:
  virtual ~TMRTopology() {}
  virtual void getConnectivity(int *nnodes, int *nedges, int *nfaces,
                               int *nvolumes, const int **volume_nodes,
                               const int **volume_edges,
                               const int **volume_faces) = 0;
};

#endif
