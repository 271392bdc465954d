/*
  TMROctant.h -- octant record and octant containers of the B200 drop-in.

  Same public interface as reference src/TMROctant.h:36-151 (TMROctant's field
  layout is byte-identical: Cython and the FE layer index raw TMROctant
  arrays).  TMROctantArray::sort()/contains() run on the GPU through
  include/tmrgpu.h; TMROctantQueue/TMROctantHash remain small host containers
  for callers that still build octant lists by hand (e.g. reference
  src/topology/TMR_TACSTopoCreator.cpp:146) -- the forest itself no longer uses
  them.
*/
#ifndef TMR_OCTANT_H
#define TMR_OCTANT_H

#include <stdlib.h>

#include "TMRBase.h"

class TMROctant {
 public:
  int childId();
  void getSibling(int id, TMROctant *sib);
  void parent(TMROctant *parent);
  void faceNeighbor(int face, TMROctant *neighbor);
  void edgeNeighbor(int edge, TMROctant *neighbor);
  void cornerNeighbor(int corner, TMROctant *neighbor);
  int compare(const TMROctant *oct) const;
  int comparePosition(const TMROctant *oct) const;
  int compareNode(const TMROctant *oct) const;
  int contains(TMROctant *oct);

  int32_t block;    // tree (block) index
  int32_t x, y, z;  // anchor coordinates
  int32_t tag;      // user / bookkeeping tag
  int16_t level;    // refinement level
  int16_t info;     // extra information
};

class TMROctantArray {
 public:
  /* takes ownership of a new[]-allocated array */
  TMROctantArray(TMROctant *array, int size, int _use_node_index = 0);
  ~TMROctantArray();

  TMROctantArray *duplicate();
  void getArray(TMROctant **_array, int *_size);
  void sort();
  TMROctant *contains(TMROctant *q, int use_nodes = 0);
  void merge(TMROctantArray *list);

 private:
  int use_node_index;
  int is_sorted;
  int size, max_size;
  TMROctant *array;
};

class TMROctantQueue {
 public:
  TMROctantQueue();
  ~TMROctantQueue();
  int length();
  void push(TMROctant *oct);
  TMROctant pop();
  TMROctantArray *toArray();

 private:
  struct Store;
  Store *store;
};

class TMROctantHash {
 public:
  TMROctantHash(int _use_node_index = 0);
  ~TMROctantHash();
  TMROctantArray *toArray();
  int addOctant(TMROctant *oct);

 private:
  struct Store;
  Store *store;
  int use_node_index;
};

/* process-wide CUDA context used by the drop-in classes (created on first
   use on device $LOCAL_RANK or $TMR_B200_DEVICE); returns NULL and prints an
   error when no GPU is available -- there is no CPU fallback */
struct tmrgpu_ctx;
extern "C" tmrgpu_ctx *tmr_b200_context(void);
/* run all subsequent forest work on this cudaStream_t (call before first use) */
extern "C" void tmr_b200_use_stream(void *stream);

#endif  // TMR_OCTANT_H
