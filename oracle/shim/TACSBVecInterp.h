/*
  oracle/shim/TACSBVecInterp.h -- TEST INFRASTRUCTURE ONLY.

  Recording stand-in for the one TACS symbol the hot path uses:
  TACSBVecInterp::addInterp (call sites: reference src/TMROctForest.cpp:6683
  and :6775).  TACS (smdogroup/tacs, unpinned -- the reference CI clones HEAD)
  is not available in this image.  The real class only stores the rows it is
  handed, so parity is defined on the argument stream recorded here.
*/
#ifndef ORACLE_SHIM_TACS_BVEC_INTERP_H
#define ORACLE_SHIM_TACS_BVEC_INTERP_H

#include <math.h>
#include <stdio.h>

#include <vector>

class TACSBVecInterp {
 public:
  TACSBVecInterp() { rowp.push_back(0); }
  void addInterp(int row, const double w[], const int cols_in[], int n) {
    rows.push_back(row);
    for (int k = 0; k < n; k++) {
      cols.push_back(cols_in[k]);
      vals.push_back(w[k]);
    }
    rowp.push_back((int)cols.size());
  }
  void incref() {}
  void decref() {}
  std::vector<int> rows;   // row ids in call order
  std::vector<int> rowp;   // CSR pointer into cols/vals, size rows+1
  std::vector<int> cols;
  std::vector<double> vals;
};

#endif
