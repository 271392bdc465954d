/*
  TMRTrilinearVolume.h -- B200 extension: a TMRVolume that is the trilinear
  image of the unit cube through 8 corner points (tree corner c has bit0 = x,
  bit1 = y, bit2 = z, as everywhere in the forest), and a topology made of
  such volumes on a hexahedral super-mesh.  It gives the forest a geometry
  without the CAD layer; TMROctForest::evaluateNodeLocations
  recognises these volumes and evaluates all node locations on the GPU, any
  other TMRVolume goes through its virtual evalPoint on the host exactly as in
  the reference (src/TMROctForest.cpp:5524-5675).
*/
#ifndef TMR_TRILINEAR_VOLUME_H
#define TMR_TRILINEAR_VOLUME_H

#include <vector>

#include "TMRTopology.h"
#include "common.h"

class TMRTrilinearVolume : public TMRVolume {
 public:
  explicit TMRTrilinearVolume(const double *corners /* 8 x (x,y,z) */)
      : TMRVolume(0, NULL) {
    for (int i = 0; i < 24; i++) X[i] = corners[i];
  }
  int evalPoint(double u, double v, double w, TMRPoint *P) {
    double p[3];
    tmrgpu::trilinear_point(X, u, v, w, p);
    P->x = p[0];
    P->y = p[1];
    P->z = p[2];
    return 0;
  }
  const double *corners() const { return X; }

 private:
  double X[24];
};

/* hexahedral super-mesh with one trilinear volume per tree */
class TMRTrilinearTopology : public TMRTopology {
 public:
  /* edge / face numbering as TMROctForest::setConnectivity derives it */
  TMRTrilinearTopology(int num_nodes, int num_edges, int num_faces,
                       int num_blocks, const int *block_conn,
                       const int *block_edge_conn, const int *block_face_conn,
                       const double *xpts /* num_nodes x (x,y,z) */)
      : nn(num_nodes), ne(num_edges), nf(num_faces), nb(num_blocks),
        bc(block_conn, block_conn + 8 * num_blocks),
        bec(block_edge_conn, block_edge_conn + 12 * num_blocks),
        bfc(block_face_conn, block_face_conn + 6 * num_blocks) {
    for (int b = 0; b < nb; b++) {
      double c[24];
      for (int k = 0; k < 8; k++) {
        for (int a = 0; a < 3; a++) c[3 * k + a] = xpts[3 * bc[8 * b + k] + a];
      }
      TMRTrilinearVolume *v = new TMRTrilinearVolume(c);
      v->incref();
      vols.push_back(v);
    }
    /* one nameable entity per super-mesh vertex / edge / face (name queries) */
    for (int i = 0; i < nn; i++) verts.push_back(held(new TMRVertex()));
    for (int i = 0; i < ne; i++) edges.push_back(held(new TMREdge()));
    for (int i = 0; i < nf; i++) faces.push_back(held(new TMRFace()));
  }
  ~TMRTrilinearTopology() {
    for (size_t i = 0; i < vols.size(); i++) vols[i]->decref();
    for (size_t i = 0; i < verts.size(); i++) verts[i]->decref();
    for (size_t i = 0; i < edges.size(); i++) edges[i]->decref();
    for (size_t i = 0; i < faces.size(); i++) faces[i]->decref();
  }
  void getVolume(int vol_num, TMRVolume **volume) { *volume = vols[vol_num]; }
  void getFace(int face_num, TMRFace **face) { *face = faces[face_num]; }
  void getEdge(int edge_num, TMREdge **edge) { *edge = edges[edge_num]; }
  void getVertex(int vertex_num, TMRVertex **vertex) { *vertex = verts[vertex_num]; }
  /* kind 0 vertex, 1 edge, 2 face, 3 volume; NULL if out of range */
  TMREntity *entity(int kind, int index) {
    if (index < 0) return NULL;
    if (kind == 0 && index < nn) return verts[index];
    if (kind == 1 && index < ne) return edges[index];
    if (kind == 2 && index < nf) return faces[index];
    if (kind == 3 && index < nb) return vols[index];
    return NULL;
  }
  void getConnectivity(int *nnodes, int *nedges, int *nfaces, int *nvolumes,
                       const int **volume_nodes, const int **volume_edges,
                       const int **volume_faces) {
    *nnodes = nn;
    *nedges = ne;
    *nfaces = nf;
    *nvolumes = nb;
    *volume_nodes = bc.data();
    *volume_edges = bec.data();
    *volume_faces = bfc.data();
  }

 private:
  int nn, ne, nf, nb;
  std::vector<int> bc, bec, bfc;
  std::vector<TMRTrilinearVolume *> vols;
  std::vector<TMRVertex *> verts;
  std::vector<TMREdge *> edges;
  std::vector<TMRFace *> faces;
  template <class T>
  static T *held(T *e) {
    e->incref();
    return e;
  }
};

#endif
