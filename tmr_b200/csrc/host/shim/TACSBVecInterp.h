/*
  shim/TACSBVecInterp.h -- stand-in for TACS's TACSBVecInterp (package
  smdogroup/tacs, not available in this image).  createInterpolation only calls
  addInterp (reference src/TMROctForest.cpp:6683,6775); this class records the
  rows it is handed.  With a real TACS on the include path this file is unused.
*/
#ifndef TMR_B200_TACS_BVEC_INTERP_SHIM_H
#define TMR_B200_TACS_BVEC_INTERP_SHIM_H

#include <vector>

class TACSBVecInterp {
 public:
  TACSBVecInterp() { rowp.push_back(0); }
  void addInterp(int row, const double w[], const int vars[], int n) {
    rows.push_back(row);
    cols.insert(cols.end(), vars, vars + n);
    vals.insert(vals.end(), w, w + n);
    rowp.push_back((int)cols.size());
  }
  std::vector<int> rows, rowp, cols;
  std::vector<double> vals;
};

#endif
