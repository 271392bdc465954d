/*
  tmr_capi.h -- flat C view of the TMROctForest / TMROctantArray C++ classes.

  The reference exposes its forest to Python through Cython
  (reference tmr/TMR.pyx:3270-3790, tmr/cpp_headers/TMR.pxd:370-423), which
  needs mpi4py/tacs/paropt/egads4py and cannot be built in this image.  This
  header is the ctypes-friendly equivalent of that binding: one C function per
  public method the Cython layer calls, nothing else.  The SAME source
  (tmr_b200/csrc/capi/tmr_capi.cpp) is compiled twice:
    * against the reference's own headers/sources  -> oracle/_ref/libtmr_ref.so
    * against this repo's drop-in TMROctForest      -> tmr_b200/lib/libtmr_b200.so
  which is the source-level proof that the drop-in keeps the class API.

  All handles are opaque; arrays are plain pointers + sizes; octant records use
  the reference's 24-byte layout (reference src/TMROctant.h:49-53).
*/
#ifndef TMR_CAPI_H
#define TMR_CAPI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* byte-identical to TMROctant (reference src/TMROctant.h:49-53) */
typedef struct {
  int32_t block, x, y, z, tag;
  int16_t level, info;
} tmrc_octant;

typedef void *tmrc_forest; /* TMROctForest* */
typedef void *tmrc_interp; /* TACSBVecInterp* (recording stand-in) */

/* ---- library identity ------------------------------------------------- */
const char *tmrc_backend(void); /* "reference-cpu" or "b200-cuda" */

/* ---- TMROctForest (reference src/TMROctForest.h:46-181) ---------------- */
tmrc_forest tmrc_forest_create(int mesh_order, int interp_type);
void tmrc_forest_destroy(tmrc_forest f);
void tmrc_set_connectivity(tmrc_forest f, int num_nodes, const int *block_conn,
                           int num_blocks);
void tmrc_set_mesh_order(tmrc_forest f, int mesh_order, int interp_type);
int tmrc_get_mesh_order(tmrc_forest f);
int tmrc_get_interp_type(tmrc_forest f);
void tmrc_repartition(tmrc_forest f, int max_rank);
void tmrc_create_trees(tmrc_forest f, int refine_level);
void tmrc_create_random_trees(tmrc_forest f, int nrand, int min_level,
                              int max_level);
tmrc_forest tmrc_duplicate(tmrc_forest f);
tmrc_forest tmrc_coarsen(tmrc_forest f);
/* refinement == NULL refines every octant by one level */
void tmrc_refine(tmrc_forest f, const int *refinement, int min_level,
                 int max_level);
void tmrc_balance(tmrc_forest f, int balance_corner);
void tmrc_create_nodes(tmrc_forest f);

/* getOctants()+getArray(): number of local octants / copy of the records */
int tmrc_num_octants(tmrc_forest f);
void tmrc_get_octants(tmrc_forest f, tmrc_octant *out);
/* Python's OctantArray.__setitem__ (reference tmr/TMR.pyx:3303-3317) writes
   through the borrowed array; this is the bulk form of that write. */
void tmrc_write_octants(tmrc_forest f, const tmrc_octant *in, int n);

/* getNodeConn: borrowed pointer, valid until the next mutating call */
void tmrc_get_node_conn(tmrc_forest f, const int **conn, int *num_elements,
                        int *num_owned_nodes);
int tmrc_get_dep_node_conn(tmrc_forest f, const int **ptr, const int **conn,
                           const double **weights);
int tmrc_get_node_numbers(tmrc_forest f, const int **node_numbers);
int tmrc_get_owned_node_range(tmrc_forest f, const int **node_range);
int tmrc_get_ext_pre_offset(tmrc_forest f);
int tmrc_get_local_node_number(tmrc_forest f, int node);
int tmrc_get_interp_knots(tmrc_forest f, const double **knots);
/* TMROctForest::evalInterp (reference src/TMROctForest.cpp:1508-1620): shape
   functions and, where the pointers are non-NULL, their first (3) and second
   (6: 11,22,33,23,13,12) derivatives at pt[3]; each array holds order^3
   values */
void tmrc_eval_interp(tmrc_forest f, const double *pt, double *N, double *N1,
                      double *N2, double *N3, double *N11, double *N22,
                      double *N33, double *N23, double *N13, double *N12);

void tmrc_get_connectivity(tmrc_forest f, int *nblocks, int *nfaces,
                           int *nedges, int *nnodes, const int **block_conn,
                           const int **block_face_conn,
                           const int **block_edge_conn,
                           const int **block_face_ids);
void tmrc_get_inverse_connectivity(tmrc_forest f, const int **node_block_conn,
                                   const int **node_block_ptr,
                                   const int **edge_block_conn,
                                   const int **edge_block_ptr,
                                   const int **face_block_conn,
                                   const int **face_block_ptr);

/* transformNode on n records in place (edge_dir = -1 for none); the two
   optional outputs receive the edge_reversed / src_face_id of each call */
void tmrc_transform_nodes(tmrc_forest f, tmrc_octant *nodes, int n,
                          int edge_dir, int *edge_reversed, int *src_face_id);

/* findEnclosing for n node-octants (info = local node index); out_index[i] is
   the index of the enclosing element in the local array or -1; out_owner the
   mpi owner estimate */
void tmrc_find_enclosing(tmrc_forest f, int order, const double *knots,
                         const tmrc_octant *nodes, int n, int *out_index,
                         int *out_owner);

/* distributeOctants / sendOctants (reference src/TMROctForest.h:166-176) on a
   caller-provided list; the received records are copied into `out` (capacity
   `cap`); returns the received count.  oct_ptr / recv_ptr: size()+1 ints. */
int tmrc_distribute_octants(tmrc_forest f, const tmrc_octant *list, int n,
                            int use_tags, int include_local, int use_node_index,
                            tmrc_octant *out, int cap, int *oct_ptr,
                            int *recv_ptr);
int tmrc_send_octants(tmrc_forest f, const tmrc_octant *list, int n,
                      const int *oct_ptr, const int *recv_ptr,
                      int use_node_index, tmrc_octant *out, int cap);

/* createInterpolation into a recording interp object */
tmrc_interp tmrc_interp_create(void);
void tmrc_interp_destroy(tmrc_interp p);
void tmrc_create_interpolation(tmrc_forest fine, tmrc_forest coarse,
                               tmrc_interp p);
/* rows in call order; rowp has nrows+1 entries */
void tmrc_interp_get(tmrc_interp p, int *nrows, int *nnz, const int **rows,
                     const int **rowp, const int **cols, const double **vals);

/* ---- TMROctantArray (reference src/TMROctant.h:65-81) ------------------ */
/* sort()+uniq in place; returns the new size */
int tmrc_array_sort(tmrc_octant *array, int n, int use_node_index);
/* contains() for nq queries against a (sorted) array; out_index = position
   of the match or -1 */
void tmrc_array_contains(tmrc_octant *array, int n, int use_node_index,
                         const tmrc_octant *queries, int nq, int use_position,
                         int *out_index);

/* merge(): set union of two arrays, `a` keeps its entry on ties (reference
   src/TMROctant.cpp:429-509); both are sorted first if needed.  Returns the
   merged size (out may be NULL to query it). */
int tmrc_array_merge(const tmrc_octant *a, int na, const tmrc_octant *b, int nb,
                     int use_node_index, tmrc_octant *out, int cap);

/* ---- TMROctantQueue (reference src/TMROctant.h:89-113, .cpp:514-590) ---- */
/* push `n` octants, pop `npop` (into popped[]), toArray() the rest into rest[];
   returns length() after the pops */
int tmrc_queue_exercise(const tmrc_octant *in, int n, int npop,
                        tmrc_octant *popped, tmrc_octant *rest);

/* ---- TMROctantHash (reference src/TMROctant.h:121-151, .cpp:599-798) ---- */
/* addOctant() for every input (added[i] = its return value), then toArray();
   returns the number of unique octants.  The ORDER of toArray() is a hash
   detail callers never rely on (they sort): compare as sets. */
int tmrc_hash_exercise(const tmrc_octant *in, int n, int use_node_index,
                       int *added, tmrc_octant *out, int cap);

/* a forest on MPI_COMM_SELF: not partitioned even inside a multi-rank job
   (reference src/TMROctForest.cpp:331-337) */
tmrc_forest tmrc_forest_create_self(int mesh_order, int interp_type);

/* ---- geometry without the CAD layer (node-location tests) --------------------
   Attach a topology whose volumes are trilinear hexahedra through the given
   corner points (xpts: num_nodes x (x,y,z)); edges and faces are numbered as
   setConnectivity numbers them.  Implemented per backend: with the drop-in's
   TMRTrilinearTopology (tmr_b200/csrc/host/tmr_b200_ext.cpp), and for the
   reference with a stand-in topology in oracle/shim/stubs.cpp whose volumes
   evaluate the same trilinear expression.  Returns 0 on success. */
int tmrc_set_trilinear_topology(tmrc_forest f, int num_nodes, const int *conn,
                                int num_blocks, const double *xpts);
/* getPoints (reference :1476-1481): borrowed pointer to num x (x,y,z) */
int tmrc_get_points(tmrc_forest f, const double **xyz);

/* ---- name queries (boundary conditions are applied through these) -----------
   Name one entity of the topology attached by tmrc_set_trilinear_topology:
   kind 0 vertex, 1 edge, 2 face, 3 volume; index in the numbering of
   tmrc_get_connectivity; name NULL clears it.  Per backend, like the topology. */
int tmrc_set_entity_name(tmrc_forest f, int kind, int index, const char *name);
/* getOctsWithName (reference :5747-5862): local octants of a named volume, or
   touching a named face (info = the face index).  Writes min(count, cap)
   records, returns count (-1: no topology / no octants). */
int tmrc_get_octs_with_name(tmrc_forest f, const char *name, tmrc_octant *out,
                            int cap);
/* getNodesWithName (reference :5882-6203): sorted unique numbers of the local
   nodes on named vertices / edges / faces.  Same return convention. */
int tmrc_get_nodes_with_name(tmrc_forest f, const char *name, int *out, int cap);

/* ---- text writers (reference :1149-1384) --------------------------------------
   which 0: writeToVTK (super-mesh), 1: writeToTecplot (super-mesh),
   2: writeForestToVTK (every local octant as a brick) */
void tmrc_write(tmrc_forest f, int which, const char *filename);

#ifdef __cplusplus
}
#endif
#endif
