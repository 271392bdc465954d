/*
  TMRBase.cpp -- runtime pieces behind TMRBase.h for the B200 drop-in.
  (Behaviour of reference src/TMRBase.cpp:42-175; the MPI struct datatypes the
  reference registers there are not needed: octants travel as 8-byte keys over
  NCCL inside the CUDA layer.)
*/
#include "TMRBase.h"

#include <stdio.h>

static int tmr_initialized = 0;
/* thread_local: one process normally drives one GPU, but the test-only
   emulation runs several ranks as threads of one process */
static thread_local int world_rank = 0;
static thread_local int world_size = 1;

extern "C" {
void tmr_b200_set_world(int rank, int size) {
  world_rank = rank;
  world_size = size < 1 ? 1 : size;
}
int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  *rank = (comm == MPI_COMM_SELF) ? 0 : world_rank;
  return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm comm, int *size) {
  *size = (comm == MPI_COMM_SELF) ? 1 : world_size;
  return MPI_SUCCESS;
}
}

void TMRInitialize() { tmr_initialized = 1; }
int TMRIsInitialized() { return tmr_initialized; }
void TMRFinalize() { tmr_initialized = 0; }

double TMREntity::eps_dist = 1e-6;
double TMREntity::eps_cosine = 1e-6;
int TMREntity::entity_id_count = 0;

TMREntity::TMREntity() : ref_count(0), name(NULL), entity_id(entity_id_count) {
  entity_id_count++;
}

TMREntity::~TMREntity() {
  if (name) delete[] name;
}

void TMREntity::incref() { ref_count++; }

void TMREntity::decref() {
  ref_count--;
  if (ref_count == 0) {
    delete this;
  }
}

void TMREntity::setName(const char *_name) {
  if (name) delete[] name;
  name = NULL;
  if (_name) {
    name = new char[strlen(_name) + 1];
    strcpy(name, _name);
  }
}

const char *TMREntity::getName() const { return name; }

void TMREntity::setTolerances(double _eps_dist, double _eps_cosine) {
  eps_dist = _eps_dist;
  eps_cosine = _eps_cosine;
}

void TMREntity::getTolerances(double *_eps_dist, double *_eps_cosine) {
  *_eps_dist = eps_dist;
  *_eps_cosine = eps_cosine;
}
