/*
  tmr_capi.cpp -- implementation of include/tmr_capi.h on top of whatever
  "TMROctForest.h" is on the include path (the reference's, or this repo's
  drop-in).  Only the public class API is used (reference
  src/TMROctForest.h:46-181, src/TMROctant.h:36-81).
*/
#include "tmr_capi.h"

#include <string.h>

#include "TMROctForest.h"

#ifndef TMRC_BACKEND_NAME
#define TMRC_BACKEND_NAME "reference-cpu"
#endif

static inline TMROctForest *F(tmrc_forest f) {
  return static_cast<TMROctForest *>(f);
}

extern "C" {

const char *tmrc_backend(void) { return TMRC_BACKEND_NAME; }

tmrc_forest tmrc_forest_create(int mesh_order, int interp_type) {
  TMROctForest *forest = new TMROctForest(
      MPI_COMM_WORLD, mesh_order, (TMRInterpolationType)interp_type);
  forest->incref();
  return forest;
}

tmrc_forest tmrc_forest_create_self(int mesh_order, int interp_type) {
  TMROctForest *forest = new TMROctForest(
      MPI_COMM_SELF, mesh_order, (TMRInterpolationType)interp_type);
  forest->incref();
  return forest;
}

void tmrc_forest_destroy(tmrc_forest f) {
  if (f) F(f)->decref();
}

void tmrc_set_connectivity(tmrc_forest f, int num_nodes, const int *block_conn,
                           int num_blocks) {
  F(f)->setConnectivity(num_nodes, block_conn, num_blocks);
}

void tmrc_set_mesh_order(tmrc_forest f, int mesh_order, int interp_type) {
  F(f)->setMeshOrder(mesh_order, (TMRInterpolationType)interp_type);
}

int tmrc_get_mesh_order(tmrc_forest f) { return F(f)->getMeshOrder(); }
int tmrc_get_interp_type(tmrc_forest f) { return (int)F(f)->getInterpType(); }
void tmrc_repartition(tmrc_forest f, int max_rank) {
  F(f)->repartition(max_rank);
}
void tmrc_create_trees(tmrc_forest f, int refine_level) {
  F(f)->createTrees(refine_level);
}
void tmrc_create_random_trees(tmrc_forest f, int nrand, int min_level,
                              int max_level) {
  F(f)->createRandomTrees(nrand, min_level, max_level);
}

tmrc_forest tmrc_duplicate(tmrc_forest f) {
  TMROctForest *dup = F(f)->duplicate();
  dup->incref();
  return dup;
}

tmrc_forest tmrc_coarsen(tmrc_forest f) {
  TMROctForest *c = F(f)->coarsen();
  c->incref();
  return c;
}

void tmrc_refine(tmrc_forest f, const int *refinement, int min_level,
                 int max_level) {
  F(f)->refine(refinement, min_level, max_level);
}

void tmrc_balance(tmrc_forest f, int balance_corner) {
  F(f)->balance(balance_corner);
}

void tmrc_create_nodes(tmrc_forest f) { F(f)->createNodes(); }

int tmrc_num_octants(tmrc_forest f) {
  TMROctantArray *arr = NULL;
  F(f)->getOctants(&arr);
  int size = 0;
  if (arr) arr->getArray(NULL, &size);
  return size;
}

void tmrc_get_octants(tmrc_forest f, tmrc_octant *out) {
  TMROctantArray *arr = NULL;
  F(f)->getOctants(&arr);
  if (!arr) return;
  int size = 0;
  TMROctant *a = NULL;
  arr->getArray(&a, &size);
  memcpy(out, a, (size_t)size * sizeof(TMROctant));
}

void tmrc_write_octants(tmrc_forest f, const tmrc_octant *in, int n) {
  TMROctantArray *arr = NULL;
  F(f)->getOctants(&arr);
  if (!arr) return;
  int size = 0;
  TMROctant *a = NULL;
  arr->getArray(&a, &size);
  if (n > size) n = size;
  memcpy(a, in, (size_t)n * sizeof(TMROctant));
}

void tmrc_get_node_conn(tmrc_forest f, const int **conn, int *num_elements,
                        int *num_owned_nodes) {
  F(f)->getNodeConn(conn, num_elements, num_owned_nodes);
}

int tmrc_get_dep_node_conn(tmrc_forest f, const int **ptr, const int **conn,
                           const double **weights) {
  return F(f)->getDepNodeConn(ptr, conn, weights);
}

int tmrc_get_node_numbers(tmrc_forest f, const int **node_numbers) {
  return F(f)->getNodeNumbers(node_numbers);
}

int tmrc_get_owned_node_range(tmrc_forest f, const int **node_range) {
  return F(f)->getOwnedNodeRange(node_range);
}

int tmrc_get_ext_pre_offset(tmrc_forest f) { return F(f)->getExtPreOffset(); }

int tmrc_get_local_node_number(tmrc_forest f, int node) {
  return F(f)->getLocalNodeNumber(node);
}

int tmrc_get_interp_knots(tmrc_forest f, const double **knots) {
  return F(f)->getInterpKnots(knots);
}

void tmrc_eval_interp(tmrc_forest f, const double *pt, double *N, double *N1,
                      double *N2, double *N3, double *N11, double *N22,
                      double *N33, double *N23, double *N13, double *N12) {
  if (N11) {
    F(f)->evalInterp(pt, N, N1, N2, N3, N11, N22, N33, N23, N13, N12);
  } else if (N1) {
    F(f)->evalInterp(pt, N, N1, N2, N3);
  } else {
    F(f)->evalInterp(pt, N);
  }
}

void tmrc_get_connectivity(tmrc_forest f, int *nblocks, int *nfaces,
                           int *nedges, int *nnodes, const int **block_conn,
                           const int **block_face_conn,
                           const int **block_edge_conn,
                           const int **block_face_ids) {
  F(f)->getConnectivity(nblocks, nfaces, nedges, nnodes, block_conn,
                        block_face_conn, block_edge_conn, block_face_ids);
}

void tmrc_get_inverse_connectivity(tmrc_forest f, const int **node_block_conn,
                                   const int **node_block_ptr,
                                   const int **edge_block_conn,
                                   const int **edge_block_ptr,
                                   const int **face_block_conn,
                                   const int **face_block_ptr) {
  F(f)->getInverseConnectivity(node_block_conn, node_block_ptr, edge_block_conn,
                               edge_block_ptr, face_block_conn, face_block_ptr);
}

void tmrc_transform_nodes(tmrc_forest f, tmrc_octant *nodes, int n,
                          int edge_dir, int *edge_reversed, int *src_face_id) {
  for (int i = 0; i < n; i++) {
    int rev = 0, fid = 0;
    F(f)->transformNode(reinterpret_cast<TMROctant *>(&nodes[i]), edge_dir,
                        &rev, &fid);
    if (edge_reversed) edge_reversed[i] = rev;
    if (src_face_id) src_face_id[i] = fid;
  }
}

void tmrc_find_enclosing(tmrc_forest f, int order, const double *knots,
                         const tmrc_octant *nodes, int n, int *out_index,
                         int *out_owner) {
  TMROctantArray *arr = NULL;
  F(f)->getOctants(&arr);
  TMROctant *base = NULL;
  int size = 0;
  if (arr) arr->getArray(&base, &size);
  for (int i = 0; i < n; i++) {
    TMROctant node;
    memcpy(&node, &nodes[i], sizeof(TMROctant));
    int owner = 0;
    TMROctant *t = F(f)->findEnclosing(order, knots, &node, &owner);
    out_index[i] = t ? (int)(t - base) : -1;
    if (out_owner) out_owner[i] = owner;
  }
}

static TMROctantArray *wrap_list(const tmrc_octant *list, int n, int use_node) {
  TMROctant *copy = new TMROctant[n > 0 ? n : 1];
  if (n > 0) memcpy(copy, list, (size_t)n * sizeof(TMROctant));
  return new TMROctantArray(copy, n, use_node);
}

static int copy_out(TMROctantArray *arr, tmrc_octant *out, int cap) {
  TMROctant *a = NULL;
  int size = 0;
  arr->getArray(&a, &size);
  const int ncopy = size < cap ? size : cap;
  if (ncopy > 0) memcpy(out, a, (size_t)ncopy * sizeof(TMROctant));
  delete arr;
  return size;
}

int tmrc_distribute_octants(tmrc_forest f, const tmrc_octant *list, int n,
                            int use_tags, int include_local, int use_node_index,
                            tmrc_octant *out, int cap, int *oct_ptr,
                            int *recv_ptr) {
  TMROctantArray *arr = wrap_list(list, n, use_node_index);
  int *p = NULL, *rp = NULL;
  TMROctantArray *got = F(f)->distributeOctants(arr, use_tags, &p, &rp,
                                                include_local, use_node_index);
  delete arr;
  int size = 1;
  MPI_Comm_size(F(f)->getMPIComm(), &size);
  for (int i = 0; i <= size; i++) {
    if (oct_ptr) oct_ptr[i] = p[i];
    if (recv_ptr) recv_ptr[i] = rp[i];
  }
  delete[] p;
  delete[] rp;
  return copy_out(got, out, cap);
}

int tmrc_send_octants(tmrc_forest f, const tmrc_octant *list, int n,
                      const int *oct_ptr, const int *recv_ptr,
                      int use_node_index, tmrc_octant *out, int cap) {
  TMROctantArray *arr = wrap_list(list, n, use_node_index);
  TMROctantArray *got = F(f)->sendOctants(arr, oct_ptr, recv_ptr, use_node_index);
  delete arr;
  return copy_out(got, out, cap);
}

tmrc_interp tmrc_interp_create(void) { return new TACSBVecInterp(); }

void tmrc_interp_destroy(tmrc_interp p) {
  delete static_cast<TACSBVecInterp *>(p);
}

void tmrc_create_interpolation(tmrc_forest fine, tmrc_forest coarse,
                               tmrc_interp p) {
  F(fine)->createInterpolation(F(coarse), static_cast<TACSBVecInterp *>(p));
}

void tmrc_interp_get(tmrc_interp p, int *nrows, int *nnz, const int **rows,
                     const int **rowp, const int **cols, const double **vals) {
  TACSBVecInterp *I = static_cast<TACSBVecInterp *>(p);
  if (nrows) *nrows = (int)I->rows.size();
  if (nnz) *nnz = (int)I->cols.size();
  if (rows) *rows = I->rows.data();
  if (rowp) *rowp = I->rowp.data();
  if (cols) *cols = I->cols.data();
  if (vals) *vals = I->vals.data();
}

int tmrc_array_sort(tmrc_octant *array, int n, int use_node_index) {
  TMROctant *copy = new TMROctant[n > 0 ? n : 1];
  memcpy(copy, array, (size_t)n * sizeof(TMROctant));
  TMROctantArray arr(copy, n, use_node_index);  // takes ownership of copy
  arr.sort();
  TMROctant *sorted = NULL;
  int size = 0;
  arr.getArray(&sorted, &size);
  memcpy(array, sorted, (size_t)size * sizeof(TMROctant));
  return size;
}

void tmrc_array_contains(tmrc_octant *array, int n, int use_node_index,
                         const tmrc_octant *queries, int nq, int use_position,
                         int *out_index) {
  TMROctant *copy = new TMROctant[n > 0 ? n : 1];
  memcpy(copy, array, (size_t)n * sizeof(TMROctant));
  TMROctantArray arr(copy, n, use_node_index);
  arr.sort();
  TMROctant *sorted = NULL;
  int size = 0;
  arr.getArray(&sorted, &size);
  for (int i = 0; i < nq; i++) {
    TMROctant q;
    memcpy(&q, &queries[i], sizeof(TMROctant));
    TMROctant *t = arr.contains(&q, use_position);
    out_index[i] = t ? (int)(t - sorted) : -1;
  }
}


static TMROctantArray *copy_to_array(const tmrc_octant *a, int n, int use_node_index) {
  TMROctant *copy = new TMROctant[n > 0 ? n : 1];
  if (n > 0) memcpy(copy, a, (size_t)n * sizeof(TMROctant));
  return new TMROctantArray(copy, n, use_node_index);
}

int tmrc_array_merge(const tmrc_octant *a, int na, const tmrc_octant *b, int nb,
                     int use_node_index, tmrc_octant *out, int cap) {
  TMROctantArray *A = copy_to_array(a, na, use_node_index);
  TMROctantArray *B = copy_to_array(b, nb, use_node_index);
  A->merge(B);
  TMROctant *arr = NULL;
  int size = 0;
  A->getArray(&arr, &size);
  if (out && size <= cap && size > 0) memcpy(out, arr, (size_t)size * sizeof(TMROctant));
  delete A;
  delete B;
  return size;
}

int tmrc_queue_exercise(const tmrc_octant *in, int n, int npop,
                        tmrc_octant *popped, tmrc_octant *rest) {
  TMROctantQueue q;
  for (int i = 0; i < n; i++) {
    TMROctant o;
    memcpy(&o, &in[i], sizeof(TMROctant));
    q.push(&o);
  }
  for (int i = 0; i < npop && q.length() > 0; i++) {
    TMROctant o = q.pop();
    memcpy(&popped[i], &o, sizeof(TMROctant));
  }
  const int left = q.length();
  TMROctantArray *arr = q.toArray();
  TMROctant *a = NULL;
  int size = 0;
  arr->getArray(&a, &size);
  if (size > 0) memcpy(rest, a, (size_t)size * sizeof(TMROctant));
  delete arr;
  return left;
}

int tmrc_hash_exercise(const tmrc_octant *in, int n, int use_node_index,
                       int *added, tmrc_octant *out, int cap) {
  TMROctantHash h(use_node_index);
  for (int i = 0; i < n; i++) {
    TMROctant o;
    memcpy(&o, &in[i], sizeof(TMROctant));
    added[i] = h.addOctant(&o);
  }
  TMROctantArray *arr = h.toArray();
  TMROctant *a = NULL;
  int size = 0;
  arr->getArray(&a, &size);
  if (out && size <= cap && size > 0) memcpy(out, a, (size_t)size * sizeof(TMROctant));
  delete arr;
  return size;
}

int tmrc_get_points(tmrc_forest f, const double **xyz) {
  TMRPoint *X = NULL;
  const int n = F(f)->getPoints(&X);
  if (xyz) *xyz = reinterpret_cast<const double *>(X);
  return X ? n : 0;
}

int tmrc_get_octs_with_name(tmrc_forest f, const char *name, tmrc_octant *out,
                            int cap) {
  TMROctantArray *list = F(f)->getOctsWithName(name);
  if (!list) return -1;
  TMROctant *a;
  int n;
  list->getArray(&a, &n);
  for (int i = 0; i < n && i < cap; i++) {
    memcpy(&out[i], &a[i], sizeof(tmrc_octant));
  }
  delete list;
  return n;
}

int tmrc_get_nodes_with_name(tmrc_forest f, const char *name, int *out, int cap) {
  int *nodes = NULL;
  const int n = F(f)->getNodesWithName(name, &nodes);
  if (!nodes) return -1;
  for (int i = 0; i < n && i < cap; i++) out[i] = nodes[i];
  delete[] nodes;
  return n;
}

void tmrc_write(tmrc_forest f, int which, const char *filename) {
  if (which == 0) F(f)->writeToVTK(filename);
  if (which == 1) F(f)->writeToTecplot(filename);
  if (which == 2) F(f)->writeForestToVTK(filename);
}

}  // extern "C"
