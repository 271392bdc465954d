/*
  octant_cycle.cpp -- the adaptation cycle written against the TMROctForest
  class API only, the way an application of the reference is written.  The SAME
  source builds against the reference's headers (src/TMROctForest.h) and
  against this repository's drop-in headers (tmr_b200/csrc/host/), and prints
  the same lines: that is what "drop-in behind the class" means.
  tests/test_zz_cpp_dropin.py compiles it both ways and compares the output.

    # drop-in, CUDA library
    g++ -std=c++14 -Itmr_b200/csrc/host -Itmr_b200/csrc/host/shim -Iinclude \
        examples/octant_cycle.cpp -Ltmr_b200/lib -ltmr_b200 \
        -Wl,-rpath,$PWD/tmr_b200/lib -o octant_cycle

  Super-mesh: 7 trees around a central one, every face orientation (the layout
  of reference examples/parallel/octant_test.cpp:52-55).
*/
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "TACSBVecInterp.h"
#include "TMROctForest.h"

static const int kNumNodes = 16, kNumTrees = 7;
static const int kConn[7 * 8] = {
    0, 1, 2,  3,  4, 5,  6, 7,  8,  10, 0,  1, 9,  11, 4,  5,  5, 11, 1,
    10, 7, 15, 3,  14, 7, 15, 3,  14, 6,  13, 2, 12, 9,  13, 4,  6,  8, 12,
    0,  2, 10, 14, 8,  12, 1, 3,  0,  2,  4,  5, 6,  7,  9,  11, 13, 15};

/* splitmix64: refinement flags and checksums must not depend on any library */
static uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}

static uint64_t fold_ints(const int *a, long n) {
  uint64_t h = 0;
  for (long i = 0; i < n; i++) h = mix(h ^ (uint64_t)(uint32_t)a[i]);
  return h;
}

static uint64_t fold_doubles(const double *a, long n) {
  uint64_t h = 0;
  for (long i = 0; i < n; i++) {
    uint64_t bits;
    memcpy(&bits, &a[i], 8);
    h = mix(h ^ bits);
  }
  return h;
}

static void report(const char *what, TMROctForest *forest) {
  TMROctantArray *octants;
  forest->getOctants(&octants);
  TMROctant *array;
  int size;
  octants->getArray(&array, &size);
  uint64_t h = 0;
  for (int i = 0; i < size; i++) {
    h = mix(h ^ (uint64_t)(uint32_t)array[i].block);
    h = mix(h ^ (uint64_t)(uint32_t)array[i].x);
    h = mix(h ^ (uint64_t)(uint32_t)array[i].y);
    h = mix(h ^ (uint64_t)(uint32_t)array[i].z);
    h = mix(h ^ (uint64_t)(uint32_t)array[i].level);
  }
  printf("%-22s %8d octants  %016llx\n", what, size, (unsigned long long)h);
}

int main(int argc, char *argv[]) {
  const int order = argc > 1 ? atoi(argv[1]) : 2;
  TMROctForest *forest = new TMROctForest(MPI_COMM_SELF, order);
  forest->incref();
  forest->setConnectivity(kNumNodes, kConn, kNumTrees);
  forest->createTrees(1);
  report("createTrees(1)", forest);

  for (int pass = 0; pass < 3; pass++) {
    TMROctantArray *octants;
    forest->getOctants(&octants);
    TMROctant *array;
    int size;
    octants->getArray(&array, &size);
    int *flags = new int[size];
    for (int i = 0; i < size; i++) {
      const uint64_t key = mix(mix(array[i].block) ^ mix((uint64_t)array[i].x * 3 + pass) ^
                               mix((uint64_t)array[i].y * 5) ^ mix((uint64_t)array[i].z * 7));
      flags[i] = (key % 100 < 30) ? 1 : ((key % 100 > 95) ? -1 : 0);
    }
    forest->refine(flags);
    delete[] flags;
    report("refine", forest);
    forest->balance(pass % 2);
    report("balance", forest);
  }

  forest->createNodes();
  const int *conn, *dep_ptr, *dep_conn, *numbers;
  const double *dep_weights;
  int num_elements, num_owned;
  forest->getNodeConn(&conn, &num_elements, &num_owned);
  const int num_dep = forest->getDepNodeConn(&dep_ptr, &dep_conn, &dep_weights);
  const int num_local = forest->getNodeNumbers(&numbers);
  printf("createNodes  order %d: %d elements, %d owned, %d local, %d dependent nodes\n",
         order, num_elements, num_owned, num_local, num_dep);
  printf("  conn        %016llx\n",
         (unsigned long long)fold_ints(conn, (long)num_elements * order * order * order));
  printf("  numbers     %016llx\n", (unsigned long long)fold_ints(numbers, num_local));
  printf("  dep_ptr     %016llx\n", (unsigned long long)fold_ints(dep_ptr, num_dep + 1));
  printf("  dep_conn    %016llx\n", (unsigned long long)fold_ints(dep_conn, dep_ptr[num_dep]));
  printf("  dep_weights %016llx\n",
         (unsigned long long)fold_doubles(dep_weights, dep_ptr[num_dep]));

  /* one multigrid level down, as tmr/TopOptUtils.py builds its hierarchy */
  TMROctForest *coarse = forest->coarsen();
  coarse->incref();
  coarse->balance(1);
  coarse->createNodes();
  report("coarsen + balance", coarse);
  TACSBVecInterp *interp = new TACSBVecInterp();
  forest->createInterpolation(coarse, interp);
  double wsum = 0.0;
  for (size_t i = 0; i < interp->vals.size(); i++) wsum += interp->vals[i];
  printf("createInterpolation: %d rows, %d entries, weight sum %.12f\n",
         (int)interp->rows.size(), (int)interp->cols.size(), wsum);
  printf("  rows        %016llx\n",
         (unsigned long long)fold_ints(interp->rows.data(), (long)interp->rows.size()));
  printf("  cols        %016llx\n",
         (unsigned long long)fold_ints(interp->cols.data(), (long)interp->cols.size()));
  delete interp;
  coarse->decref();
  forest->decref();
  return 0;
}
