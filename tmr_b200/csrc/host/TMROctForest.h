/*
  TMROctForest.h -- the forest-of-octrees class of the B200 drop-in.

  Public interface = reference src/TMROctForest.h:46-181 (same names,
  signatures and defaults, so tmr/TMR.pyx and the FE layer recompile
  unchanged).  The private half is new: the octants, node numbering and
  dependent-node data live in GPU memory behind include/tmrgpu.h; the host
  arrays the getters hand out are mirrors, materialised on first request and
  valid until the next mutating call (the reference's borrowing rule).
*/
#ifndef TMR_OCTANT_FOREST_H
#define TMR_OCTANT_FOREST_H

#include <vector>

#include "TACSBVecInterp.h"
#include "TMROctant.h"
#include "TMRTopology.h"

struct tmrgpu_forest;

class TMROctForest : public TMREntity {
 public:
  static const int MAX_ORDER = 16;

  TMROctForest(MPI_Comm _comm, int mesh_order = 2,
               TMRInterpolationType interp_type = TMR_GAUSS_LOBATTO_POINTS);
  ~TMROctForest();

  MPI_Comm getMPIComm() { return comm; }

  void setTopology(TMRTopology *_topo);
  TMRTopology *getTopology();

  void setConnectivity(int _num_nodes, const int *_block_conn, int _num_blocks);
  void setFullConnectivity(int _num_nodes, int _num_edges, int _num_faces,
                           int _num_blocks, const int *_block_conn,
                           const int *_block_edge_conn,
                           const int *_block_face_conn);

  void setMeshOrder(int mesh_order, TMRInterpolationType interp_type =
                                        TMR_GAUSS_LOBATTO_POINTS);
  int getMeshOrder();
  TMRInterpolationType getInterpType();

  void repartition(int max_rank = -1);

  void createTrees(int refine_level);
  void createRandomTrees(int nrand = 10, int min_level = 0, int max_level = 8);

  TMROctForest *duplicate();
  TMROctForest *coarsen();

  void refine(const int refinement[] = NULL, int min_level = 0,
              int max_level = TMR_MAX_LEVEL);

  void balance(int balance_corner = 0);

  void createNodes();

  void getNodeConn(const int **_conn = NULL, int *_num_elements = NULL,
                   int *_num_owned_nodes = NULL, int *_num_local_nodes = NULL);
  int getDepNodeConn(const int **_ptr = NULL, const int **_conn = NULL,
                     const double **_weights = NULL);

  void createInterpolation(TMROctForest *coarse, TACSBVecInterp *interp);

  TMROctantArray *getOctsWithName(const char *name);
  int getNodesWithName(const char *name, int **_nodes);

  int getOwnedNodeRange(const int **_node_range);

  void getOctants(TMROctantArray **_octants);
  int getNodeNumbers(const int **_node_numbers);
  int getExtPreOffset();
  int getPoints(TMRPoint **_X);
  int getLocalNodeNumber(int node);
  int getInterpKnots(const double **_knots);
  void evalInterp(const double pt[], double N[]);
  void evalInterp(const double pt[], double N[], double Nxi[], double Neta[],
                  double Nzeta[]);
  void evalInterp(const double pt[], double N[], double N1[], double N2[],
                  double N3[], double N11[], double N22[], double N33[],
                  double N23[], double N13[], double N12[]);

  void getConnectivity(int *_nblocks, int *_nfaces, int *_nedges, int *_nnodes,
                       const int **_block_conn, const int **_block_face_conn,
                       const int **_block_edge_conn,
                       const int **_block_face_ids);
  void getInverseConnectivity(const int **_node_block_conn,
                              const int **_node_block_ptr,
                              const int **_edge_block_conn,
                              const int **_edge_block_ptr,
                              const int **_face_block_conn,
                              const int **_face_block_ptr);

  TMROctant *findEnclosing(const int order, const double *knots,
                           TMROctant *node, int *mpi_owner = NULL);

  void transformNode(TMROctant *oct, int edge_dir = -1,
                     int *edge_reversed = NULL, int *src_face_id = NULL);

  TMROctantArray *distributeOctants(TMROctantArray *list, int use_tags = 0,
                                    int **oct_ptr = NULL,
                                    int **oct_recv_ptr = NULL,
                                    int include_local = 0,
                                    int use_node_index = 0);
  TMROctantArray *sendOctants(TMROctantArray *list, const int *oct_ptr,
                              const int *oct_recv_ptr, int use_node_index = 0);

  void writeToVTK(const char *filename);
  void writeToTecplot(const char *filename);
  void writeForestToVTK(const char *filename);

  /* B200 extension (not in the reference): the device forest behind this
     object, for callers that keep node data on the GPU */
  tmrgpu_forest *getDeviceForest() { return dev; }
  /* B200 extension: the whole prolongation in ONE hand-off instead of one
     TACSBVecInterp::addInterp call per row (reference :6683, :6775): CSR
     arrays owned by this forest, rows in exactly the order
     createInterpolation() would emit them; returns the number of rows.
     createInterpolation(coarse, interp) is this followed by the addInterp loop. */
  int createInterpolationCSR(TMROctForest *coarse, const int **rows,
                             const int **rowp, const int **cols,
                             const double **vals);

 private:
  /* super-mesh connectivity shared between a forest and its duplicates */
  class BlockTables : public TMREntity {
   public:
    BlockTables();
    ~BlockTables();
    void nodesToBlocks();
    void edgesFromNodes();
    void facesFromNodes();
    void edgesToBlocks();
    void facesToBlocks();
    void entityOwners();
    int num_nodes, num_edges, num_faces, num_blocks;
    int *block_conn, *block_face_conn, *block_edge_conn, *block_face_ids;
    int *node_block_ptr, *node_block_conn;
    int *edge_block_ptr, *edge_block_conn;
    int *face_block_ptr, *face_block_conn;
    int *face_block_owners, *edge_block_owners, *node_block_owners;
  };

  void dropTables();
  void superMeshPoints(std::vector<TMRPoint> &X);
  void dropMeshData(int drop_octants, int drop_owners);
  void dropHostNodeMirrors();
  void pushTablesToDevice();
  /* make the device array current if the host mirror was handed out */
  int syncOctantsToDevice();
  void octantsReplacedOnDevice();
  void fetchNodeData();
  void fetchNodeNumbers();
  void evaluateNodeLocations();
  int ensureDevice();

  /* createInterpolationCSR results */
  int *interp_rows, *interp_rowp, *interp_cols;
  double *interp_vals;

  MPI_Comm comm;
  int mpi_rank, mpi_size;

  TMRInterpolationType interp_type;
  double *interp_knots;
  int mesh_order;

  TMRTopology *topo;
  BlockTables *tables;

  tmrgpu_forest *dev;

  /* host mirrors */
  TMROctantArray *octants; /* NULL until requested */
  int octants_exposed;     /* mirror may have been modified by the caller */
  TMROctant *owners;

  int *conn, *node_numbers, *node_range;
  int num_local_nodes, num_dep_nodes, num_owned_nodes, ext_pre_offset;
  int num_elements_nodes; /* element count the node data was built for */
  int *dep_ptr, *dep_conn;
  double *dep_weights;
  TMRPoint *X;
  int nodes_on_host; /* mirrors of the node data are current */
  int nodes_exist;   /* createNodes() has run since the last invalidation */
};

#endif  // TMR_OCTANT_FOREST_H
