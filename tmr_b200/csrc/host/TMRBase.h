/*
  TMRBase.h -- base definitions of the B200 drop-in: the constants, small value
  types and the intrusive reference count that the forest API exposes
  (interface of reference src/TMRBase.h:37-188; only what the octree hot path
  and its callers use).
*/
#ifndef TMR_BASE_H
#define TMR_BASE_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mpi.h"

#define TMR_EXTERN_C_BEGIN extern "C" {
#define TMR_EXTERN_C_END }

/* deepest octant level; coordinates are integers in [0, 2^30) */
static const int TMR_MAX_LEVEL = 30;

enum TMRInterpolationType {
  TMR_UNIFORM_POINTS,
  TMR_GAUSS_LOBATTO_POINTS,
  TMR_BERNSTEIN_POINTS
};

class TMRPoint {
 public:
  TMRPoint() {}
  TMRPoint(const TMRPoint &p) : x(p.x), y(p.y), z(p.z) {}
  TMRPoint(double _x, double _y, double _z) : x(_x), y(_y), z(_z) {}
  TMRPoint &operator=(const TMRPoint &p) {
    x = p.x;
    y = p.y;
    z = p.z;
    return *this;
  }
  inline void zero() { x = y = z = 0.0; }
  inline double dot(const TMRPoint &p) const {
    return x * p.x + y * p.y + z * p.z;
  }
  inline double dot(const TMRPoint *p) const {
    return x * p->x + y * p->y + z * p->z;
  }
  double x, y, z;
};

/* (index, weight) pair used when assembling interpolation rows */
class TMRIndexWeight {
 public:
  int index;
  double weight;

  /* sort by index and merge equal indices by summing their weights; returns
     the merged length */
  static int uniqueSort(TMRIndexWeight *array, int size) {
    qsort(array, size, sizeof(TMRIndexWeight), by_index);
    int out = 0;
    for (int i = 0; i < size; i++) {
      if (out > 0 && array[out - 1].index == array[i].index) {
        array[out - 1].weight += array[i].weight;
      } else {
        array[out++] = array[i];
      }
    }
    return out;
  }

 private:
  static int by_index(const void *a, const void *b) {
    return static_cast<const TMRIndexWeight *>(a)->index -
           static_cast<const TMRIndexWeight *>(b)->index;
  }
};

void TMRInitialize();
int TMRIsInitialized();
void TMRFinalize();

/* intrusive reference counting: objects start at 0 and are deleted when a
   decref() brings the count back to 0 */
class TMREntity {
 public:
  TMREntity();
  virtual ~TMREntity();
  void setName(const char *name);
  const char *getName() const;
  void incref();
  void decref();
  static void setTolerances(double _eps_dist, double _eps_cosine);
  static void getTolerances(double *_eps_dist, double *_eps_cosine);
  int getEntityId() const { return entity_id; }

 protected:
  static double eps_dist;
  static double eps_cosine;

 private:
  int ref_count;
  char *name;
  const int entity_id;
  static int entity_id_count;
};

#endif  // TMR_BASE_H
