/*
  oracle/shim/stubs.cpp -- TEST INFRASTRUCTURE ONLY.
  Link stubs for symbols the reference forest sources reference but never
  reach when the forest is driven through setConnectivity (no CAD topology):
  TMRTopology getters (declared reference src/TMRTopology.h:383-400) and two
  LAPACK routines (reference src/tmrlapack.h:28-47, Bernstein+topology branch
  of evaluateNodeLocations only).
*/
#include <stdio.h>
#include <stdlib.h>

#include "TMRTopology.h"

static void unreachable(const char *what) {
  fprintf(stderr, "oracle stub reached: %s\n", what);
  abort();
}

void TMRTopology::getVolume(int, TMRVolume **) { unreachable("getVolume"); }
void TMRTopology::getFace(int, TMRFace **) { unreachable("getFace"); }
void TMRTopology::getEdge(int, TMREdge **) { unreachable("getEdge"); }
void TMRTopology::getVertex(int, TMRVertex **) { unreachable("getVertex"); }
void TMRTopology::getConnectivity(int *, int *, int *, int *, const int **,
                                  const int **, const int **) {
  unreachable("getConnectivity");
}

extern "C" {
void dgetrf_(int *, int *, double *, int *, int *, int *) {
  unreachable("dgetrf_");
}
void dgetrs_(const char *, int *, int *, double *, int *, int *, double *,
             int *, int *) {
  unreachable("dgetrs_");
}
}
