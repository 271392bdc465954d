/*
  shim/TMRTopology.h -- minimal declarations of the CAD topology classes the
  forest can be attached to: the part of reference src/TMRTopology.h the hot
  path touches (TMRVolume::evalPoint :261, TMRTopology::getVolume :392,
  getConnectivity :397-400) plus what the name queries ask of it: an entity
  per vertex / edge / face that carries a name (TMREntity::getName;
  TMRTopology::getVertex / getEdge / getFace :367-391).  The CAD layer is
  outside the hot path; in the full TMR tree the real header is used instead
  of this one.
*/
#ifndef TMR_B200_TOPOLOGY_SHIM_H
#define TMR_B200_TOPOLOGY_SHIM_H

#include "TMRBase.h"

/* nameable stand-ins of reference src/TMRTopology.h:47-75, 79-141, 168-245:
   the geometric evaluations of the real classes are not part of this build */
class TMRVertex : public TMREntity {
 public:
  virtual ~TMRVertex() {}
};
class TMREdge : public TMREntity {
 public:
  virtual ~TMREdge() {}
};
class TMRFace : public TMREntity {
 public:
  virtual ~TMRFace() {}
};

/* reference src/TMRTopology.h:251-280 */
class TMRVolume : public TMREntity {
 public:
  TMRVolume(int /*nfaces*/, TMRFace ** /*faces*/) {}
  virtual ~TMRVolume() {}
  virtual void getRange(double *umin, double *vmin, double *wmin, double *umax,
                        double *vmax, double *wmax) {
    *umin = *vmin = *wmin = 0.0;
    *umax = *vmax = *wmax = 0.0;
  }
  /* parametric point (u,v,w) in [0,1]^3 -> physical location */
  virtual int evalPoint(double, double, double, TMRPoint *X) {
    X->zero();
    return 1;
  }
};

class TMRTopology : public TMREntity {
 public:
  virtual ~TMRTopology() {}
  virtual void getVolume(int vol_num, TMRVolume **volume) = 0;
  /* a topology without nameable boundary entities answers NULL */
  virtual void getFace(int /*face_num*/, TMRFace **face) { *face = NULL; }
  virtual void getEdge(int /*edge_num*/, TMREdge **edge) { *edge = NULL; }
  virtual void getVertex(int /*vertex_num*/, TMRVertex **vertex) { *vertex = NULL; }
  virtual void getConnectivity(int *nnodes, int *nedges, int *nfaces,
                               int *nvolumes, const int **volume_nodes,
                               const int **volume_edges,
                               const int **volume_faces) = 0;
};

#endif
