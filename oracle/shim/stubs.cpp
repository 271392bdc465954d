/*
  oracle/shim/stubs.cpp -- TEST INFRASTRUCTURE ONLY.
  Link definitions for the CAD-layer symbols the reference forest sources
  reference (reference src/TMRTopology.h:251-280,382-400; their own
  definitions live in src/TMRTopology.cpp, which needs the whole geometry
  layer and is not compiled here) and two LAPACK routines (reference
  src/tmrlapack.h:28-47, Bernstein+topology branch of evaluateNodeLocations
  only, never reached by the tests).

  TMRTopology is served by a stand-in: a hexahedral super-mesh whose volumes
  are trilinear images of the unit cube (tmrc_set_trilinear_topology,
  include/tmr_capi.h), so that the UNMODIFIED reference
  TMROctForest::evaluateNodeLocations (src/TMROctForest.cpp:5524-5675) can be
  run as the oracle of the node-location path.  The stand-in also owns one
  nameable vertex / edge / face per super-mesh entity (tmrc_set_entity_name),
  so that the UNMODIFIED getOctsWithName / getNodesWithName
  (src/TMROctForest.cpp:5747-5862, :5882-6203) are the oracle of the name
  queries.
*/
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <vector>

#include "TMROctForest.h"
#include "TMRTopology.h"

static void unreachable(const char *what) {
  fprintf(stderr, "oracle stub reached: %s\n", what);
  abort();
}

/* ---- TMRVolume: the base-class members of reference src/TMRTopology.cpp:910-968
   the vtable needs ---- */
TMRVolume::TMRVolume(int, TMRFace **) {
  num_faces = 0;
  faces = NULL;
  mesh = NULL;
}
TMRVolume::~TMRVolume() {}
void TMRVolume::getRange(double *umin, double *vmin, double *wmin, double *umax,
                         double *vmax, double *wmax) {
  *umin = *vmin = *wmin = 0.0;
  *umax = *vmax = *wmax = 0.0;
}
int TMRVolume::evalPoint(double, double, double, TMRPoint *X) {
  X->zero();
  return 1;
}

namespace {

/* trilinear image of (u,v,w) through 8 corners (corner c: bit0 x, bit1 y, bit2
   z); the same expression, in the same order, as tmrgpu::trilinear_point
   (tmr_b200/csrc/gpu/common.h) -- the geometry is test input, both sides must
   evaluate it identically for the comparison to be about the forest */
class TestTrilinearVolume : public TMRVolume {
 public:
  explicit TestTrilinearVolume(const double *c) : TMRVolume(0, NULL) {
    for (int i = 0; i < 24; i++) X[i] = c[i];
  }
  int evalPoint(double u, double v, double w, TMRPoint *P) {
    const double a[2] = {1.0 - u, u}, b[2] = {1.0 - v, v}, c[2] = {1.0 - w, w};
    double p[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < 8; k++) {
      const double n = (a[k & 1] * b[(k >> 1) & 1]) * c[k >> 2];
      p[0] += n * X[3 * k];
      p[1] += n * X[3 * k + 1];
      p[2] += n * X[3 * k + 2];
    }
    P->x = p[0];
    P->y = p[1];
    P->z = p[2];
    return 0;
  }
  double X[24];
};

}  // namespace

/* ---- TMRVertex / TMREdge / TMRFace: base-class members (reference
   src/TMRTopology.cpp) reduced to what a nameable entity needs; the geometric
   evaluations are never asked for by the forest */
TMRVertex::TMRVertex() {
  var = -1;
  copy = NULL;
}
TMRVertex::~TMRVertex() {}
int TMRVertex::getParamOnEdge(TMREdge *, double *) {
  unreachable("TMRVertex::getParamOnEdge");
  return 1;
}
int TMRVertex::getParamsOnFace(TMRFace *, double *, double *) {
  unreachable("TMRVertex::getParamsOnFace");
  return 1;
}
TMREdge::TMREdge() {
  v1 = v2 = NULL;
  mesh = NULL;
  source = NULL;
  copy = NULL;
}
TMREdge::~TMREdge() {}
int TMREdge::invEvalPoint(TMRPoint, double *) {
  unreachable("TMREdge::invEvalPoint");
  return 1;
}
int TMREdge::evalDeriv(double, TMRPoint *, TMRPoint *) {
  unreachable("TMREdge::evalDeriv");
  return 1;
}
int TMREdge::eval2ndDeriv(double, TMRPoint *, TMRPoint *, TMRPoint *) {
  unreachable("TMREdge::eval2ndDeriv");
  return 1;
}
int TMREdge::getParamsOnFace(TMRFace *, double, int, double *, double *) {
  unreachable("TMREdge::getParamsOnFace");
  return 1;
}
TMRFace::TMRFace(int _orientation) {
  orientation = _orientation;
  mesh = NULL;
  source_volume = NULL;
  source = NULL;
  copy_orient = 0;
  copy = NULL;
  num_loops = max_num_loops = 0;
  loops = NULL;
  loop_dirs = NULL;
}
TMRFace::~TMRFace() {}
int TMRFace::getOrientation() { return orientation; }
int TMRFace::invEvalPoint(TMRPoint, double *, double *) {
  unreachable("TMRFace::invEvalPoint");
  return 1;
}
int TMRFace::evalDeriv(double, double, TMRPoint *, TMRPoint *, TMRPoint *) {
  unreachable("TMRFace::evalDeriv");
  return 1;
}
int TMRFace::eval2ndDeriv(double, double, TMRPoint *, TMRPoint *, TMRPoint *,
                          TMRPoint *, TMRPoint *, TMRPoint *) {
  unreachable("TMRFace::eval2ndDeriv");
  return 1;
}

namespace {
class TestVertex : public TMRVertex {
 public:
  int evalPoint(TMRPoint *p) {
    p->zero();
    return 0;
  }
};
class TestEdge : public TMREdge {
 public:
  void getRange(double *tmin, double *tmax) {
    *tmin = 0.0;
    *tmax = 1.0;
  }
  int evalPoint(double, TMRPoint *X) {
    X->zero();
    return 0;
  }
};
class TestFace : public TMRFace {
 public:
  void getRange(double *umin, double *vmin, double *umax, double *vmax) {
    *umin = *vmin = 0.0;
    *umax = *vmax = 1.0;
  }
  int evalPoint(double, double, TMRPoint *X) {
    X->zero();
    return 0;
  }
};

struct TopoData {
  int nn, ne, nf, nb;
  std::vector<int> bc, bec, bfc;
  std::vector<TMRVolume *> vols;
  std::vector<TMRVertex *> verts;
  std::vector<TMREdge *> edges;
  std::vector<TMRFace *> faces;
};
std::map<const TMRTopology *, TopoData *> g_topos;

TopoData *data_of(const TMRTopology *t) {
  std::map<const TMRTopology *, TopoData *>::iterator it = g_topos.find(t);
  if (it == g_topos.end()) unreachable("TMRTopology without stand-in data");
  return it->second;
}

}  // namespace

/* the stand-in never touches the real class's members (geo, the volume / face
   / edge maps): every getter below answers from TopoData */
TMRTopology::TMRTopology(MPI_Comm, TMRModel *) {}
TMRTopology::~TMRTopology() {
  std::map<const TMRTopology *, TopoData *>::iterator it = g_topos.find(this);
  if (it != g_topos.end()) {
    TopoData *d = it->second;
    for (size_t i = 0; i < d->vols.size(); i++) d->vols[i]->decref();
    for (size_t i = 0; i < d->verts.size(); i++) d->verts[i]->decref();
    for (size_t i = 0; i < d->edges.size(); i++) d->edges[i]->decref();
    for (size_t i = 0; i < d->faces.size(); i++) d->faces[i]->decref();
    delete it->second;
    g_topos.erase(it);
  }
}
void TMRTopology::getVolume(int i, TMRVolume **v) { *v = data_of(this)->vols[i]; }
void TMRTopology::getFace(int i, TMRFace **f) { *f = data_of(this)->faces[i]; }
void TMRTopology::getEdge(int i, TMREdge **e) { *e = data_of(this)->edges[i]; }
void TMRTopology::getVertex(int i, TMRVertex **v) { *v = data_of(this)->verts[i]; }
void TMRTopology::getConnectivity(int *nnodes, int *nedges, int *nfaces, int *nvolumes,
                                  const int **volume_nodes, const int **volume_edges,
                                  const int **volume_faces) {
  TopoData *d = data_of(this);
  *nnodes = d->nn;
  *nedges = d->ne;
  *nfaces = d->nf;
  *nvolumes = d->nb;
  *volume_nodes = d->bc.data();
  *volume_edges = d->bec.data();
  *volume_faces = d->bfc.data();
}

extern "C" {
/* include/tmr_capi.h */
int tmrc_set_trilinear_topology(void *f, int num_nodes, const int *conn,
                                int num_blocks, const double *xpts) {
  TMROctForest *forest = static_cast<TMROctForest *>(f);
  /* edge and face numbering as the reference's own setConnectivity derives it */
  TMROctForest *tmp = new TMROctForest(MPI_COMM_SELF);
  tmp->incref();
  tmp->setConnectivity(num_nodes, conn, num_blocks);
  int nb, nf, ne, nn;
  const int *bc, *bfc, *bec, *ids;
  tmp->getConnectivity(&nb, &nf, &ne, &nn, &bc, &bfc, &bec, &ids);
  TopoData *d = new TopoData();
  d->nn = nn;
  d->ne = ne;
  d->nf = nf;
  d->nb = nb;
  d->bc.assign(bc, bc + 8 * nb);
  d->bec.assign(bec, bec + 12 * nb);
  d->bfc.assign(bfc, bfc + 6 * nb);
  for (int b = 0; b < nb; b++) {
    double c[24];
    for (int k = 0; k < 8; k++) {
      for (int a = 0; a < 3; a++) c[3 * k + a] = xpts[3 * conn[8 * b + k] + a];
    }
    TMRVolume *v = new TestTrilinearVolume(c);
    v->incref();
    d->vols.push_back(v);
  }
  for (int i = 0; i < nn; i++) {
    d->verts.push_back(new TestVertex());
    d->verts.back()->incref();
  }
  for (int i = 0; i < ne; i++) {
    d->edges.push_back(new TestEdge());
    d->edges.back()->incref();
  }
  for (int i = 0; i < nf; i++) {
    d->faces.push_back(new TestFace());
    d->faces.back()->incref();
  }
  tmp->decref();
  TMRTopology *topo = new TMRTopology(MPI_COMM_SELF, NULL);
  g_topos[topo] = d;
  forest->setTopology(topo);
  return 0;
}

/* include/tmr_capi.h: kind 0 vertex, 1 edge, 2 face, 3 volume */
int tmrc_set_entity_name(void *f, int kind, int index, const char *name) {
  TMROctForest *forest = static_cast<TMROctForest *>(f);
  TMRTopology *topo = forest->getTopology();
  if (!topo) return 1;
  TopoData *d = data_of(topo);
  TMREntity *e = NULL;
  if (kind == 0 && index >= 0 && index < d->nn) e = d->verts[index];
  if (kind == 1 && index >= 0 && index < d->ne) e = d->edges[index];
  if (kind == 2 && index >= 0 && index < d->nf) e = d->faces[index];
  if (kind == 3 && index >= 0 && index < d->nb) e = d->vols[index];
  if (!e) return 1;
  e->setName(name);
  return 0;
}

void dgetrf_(int *, int *, double *, int *, int *, int *) {
  unreachable("dgetrf_");
}
void dgetrs_(const char *, int *, int *, double *, int *, int *, double *,
             int *, int *) {
  unreachable("dgetrs_");
}
}
