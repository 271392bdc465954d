/*
  TMROctForest.cpp -- host half of the B200 drop-in forest.

  This file owns what is cheap and serial (the O(nblocks) super-mesh tables,
  argument checking, lazily materialised host mirrors) and forwards every
  data-parallel step to the CUDA layer through include/tmrgpu.h.  Each public
  method cites the reference code whose behaviour it keeps.
*/
#include "TMROctForest.h"

#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "TMRTrilinearVolume.h"
#include "common.h" /* shared host/device geometry: transform_node, tables */
#include "tmrgpu.h"

/* ---- small helpers -------------------------------------------------------- */
namespace {

/* corners of block face f in (u,v) order, k = u + 2v */
inline int face_corner(int f, int k) {
  tmrgpu::i32 x, y, z;
  tmrgpu::face_place(f, f & 1, k & 1, k >> 1, &x, &y, &z);
  return x + 2 * y + 4 * z;
}

/* owner face node k coincides with node orient(id,k) of a face whose
   orientation id relative to the owner is `id` */
inline int orient(int id, int k) {
  tmrgpu::i32 a, b;
  tmrgpu::owner_to_face(id, 1, k & 1, k >> 1, &a, &b);
  return a + 2 * b;
}

inline int *copy_ints(const int *src, int n) {
  int *dst = new int[n > 0 ? n : 1];
  if (n > 0) memcpy(dst, src, (size_t)n * sizeof(int));
  return dst;
}

/* turn counts stored at ptr[i+1] into CSR offsets */
inline void counts_to_offsets(int *ptr, int n) {
  ptr[0] = 0;
  for (int i = 0; i < n; i++) ptr[i + 1] += ptr[i];
}

inline bool same_pair(int a0, int a1, int b0, int b1) {
  return (a0 == b0 && a1 == b1) || (a0 == b1 && a1 == b0);
}

}  // namespace

/* ---- BlockTables (reference src/TMROctForest.cpp:558-1143) ------------------ */
TMROctForest::BlockTables::BlockTables() {
  num_nodes = num_edges = num_faces = num_blocks = 0;
  block_conn = block_face_conn = block_edge_conn = block_face_ids = NULL;
  node_block_ptr = node_block_conn = NULL;
  edge_block_ptr = edge_block_conn = NULL;
  face_block_ptr = face_block_conn = NULL;
  face_block_owners = edge_block_owners = node_block_owners = NULL;
}

TMROctForest::BlockTables::~BlockTables() {
  delete[] block_conn;
  delete[] block_face_conn;
  delete[] block_edge_conn;
  delete[] block_face_ids;
  delete[] node_block_ptr;
  delete[] node_block_conn;
  delete[] edge_block_ptr;
  delete[] edge_block_conn;
  delete[] face_block_ptr;
  delete[] face_block_conn;
  delete[] face_block_owners;
  delete[] edge_block_owners;
  delete[] node_block_owners;
}

/* node -> (8*block + corner), blocks ascending (reference :643-693) */
void TMROctForest::BlockTables::nodesToBlocks() {
  node_block_ptr = new int[num_nodes + 1];
  std::fill(node_block_ptr, node_block_ptr + num_nodes + 1, 0);
  for (int i = 0; i < 8 * num_blocks; i++) node_block_ptr[block_conn[i] + 1]++;
  counts_to_offsets(node_block_ptr, num_nodes);
  node_block_conn = new int[node_block_ptr[num_nodes] > 0
                                ? node_block_ptr[num_nodes]
                                : 1];
  std::vector<int> fill(node_block_ptr, node_block_ptr + num_nodes);
  for (int b = 0; b < num_blocks; b++) {
    for (int c = 0; c < 8; c++) {
      const int node = block_conn[8 * b + c];
      /* the reference stores the FIRST corner of b that equals this node */
      int first = 0;
      while (block_conn[8 * b + first] != node) first++;
      node_block_conn[fill[node]++] = 8 * b + first;
    }
  }
}

/* unique edge numbers in first-encounter order (reference :771-842) */
void TMROctForest::BlockTables::edgesFromNodes() {
  block_edge_conn = new int[12 * num_blocks > 0 ? 12 * num_blocks : 1];
  std::fill(block_edge_conn, block_edge_conn + 12 * num_blocks, -1);
  int next = 0;
  for (int b = 0; b < num_blocks; b++) {
    for (int e = 0; e < 12; e++) {
      if (block_edge_conn[12 * b + e] >= 0) continue;
      const int n1 = block_conn[8 * b + tmrgpu::edge_corner(e, 0)];
      const int n2 = block_conn[8 * b + tmrgpu::edge_corner(e, 1)];
      std::vector<int> unnumbered(1, 12 * b + e);
      int number = -1;
      for (int ip = node_block_ptr[n1]; ip < node_block_ptr[n1 + 1]; ip++) {
        const int ob = node_block_conn[ip] / 8;
        for (int oe = 0; oe < 12; oe++) {
          const int m1 = block_conn[8 * ob + tmrgpu::edge_corner(oe, 0)];
          const int m2 = block_conn[8 * ob + tmrgpu::edge_corner(oe, 1)];
          if (!same_pair(n1, n2, m1, m2)) continue;
          if (block_edge_conn[12 * ob + oe] >= 0) {
            number = block_edge_conn[12 * ob + oe];
          } else if (unnumbered.size() < 128) {
            unnumbered.push_back(12 * ob + oe);
          }
        }
      }
      if (number < 0) number = next++;
      for (size_t k = 0; k < unnumbered.size(); k++) {
        block_edge_conn[unnumbered[k]] = number;
      }
    }
  }
  num_edges = next;
}

/* unique face numbers in first-encounter order (reference :847-945) */
void TMROctForest::BlockTables::facesFromNodes() {
  block_face_conn = new int[6 * num_blocks > 0 ? 6 * num_blocks : 1];
  std::fill(block_face_conn, block_face_conn + 6 * num_blocks, -1);
  int next = 0;
  for (int b = 0; b < num_blocks; b++) {
    for (int f = 0; f < 6; f++) {
      if (block_face_conn[6 * b + f] >= 0) continue;
      int fn[4];
      for (int k = 0; k < 4; k++) fn[k] = block_conn[8 * b + face_corner(f, k)];
      std::vector<int> unnumbered(1, 6 * b + f);
      int number = -1;
      const int pivot = fn[0];
      for (int ip = node_block_ptr[pivot]; ip < node_block_ptr[pivot + 1];
           ip++) {
        const int ob = node_block_conn[ip] / 8;
        if (ob == b) continue;
        for (int of = 0; of < 6; of++) {
          int on[4];
          for (int k = 0; k < 4; k++) {
            on[k] = block_conn[8 * ob + face_corner(of, k)];
          }
          bool match = false;
          for (int id = 0; id < 8 && !match; id++) {
            match = (fn[0] == on[orient(id, 0)] && fn[1] == on[orient(id, 1)] &&
                     fn[2] == on[orient(id, 2)] && fn[3] == on[orient(id, 3)]);
          }
          if (!match) continue;
          if (block_face_conn[6 * ob + of] >= 0) {
            number = block_face_conn[6 * ob + of];
          } else if (unnumbered.size() < 128) {
            unnumbered.push_back(6 * ob + of);
          }
        }
      }
      if (number < 0) number = next++;
      for (size_t k = 0; k < unnumbered.size(); k++) {
        block_face_conn[unnumbered[k]] = number;
      }
    }
  }
  num_faces = next;
}

/* edge -> (12*block + local edge), owner (lowest block) first
   (reference :698-766) */
void TMROctForest::BlockTables::edgesToBlocks() {
  edge_block_ptr = new int[num_edges + 1];
  std::fill(edge_block_ptr, edge_block_ptr + num_edges + 1, 0);
  for (int i = 0; i < 12 * num_blocks; i++) {
    edge_block_ptr[block_edge_conn[i] + 1]++;
  }
  counts_to_offsets(edge_block_ptr, num_edges);
  edge_block_conn =
      new int[edge_block_ptr[num_edges] > 0 ? edge_block_ptr[num_edges] : 1];
  std::vector<int> fill(edge_block_ptr, edge_block_ptr + num_edges);
  for (int b = 0; b < num_blocks; b++) {
    for (int e = 0; e < 12; e++) {
      edge_block_conn[fill[block_edge_conn[12 * b + e]]++] = b;
    }
  }
  for (int edge = 0; edge < num_edges; edge++) {
    const int begin = edge_block_ptr[edge], end = edge_block_ptr[edge + 1];
    int owner = num_blocks;
    for (int ip = begin; ip < end; ip++) {
      owner = std::min(owner, edge_block_conn[ip]);
    }
    if (owner == num_blocks) continue;
    int oe = 0;
    while (oe < 12 && block_edge_conn[12 * owner + oe] != edge) oe++;
    const int n1 = block_conn[8 * owner + tmrgpu::edge_corner(oe, 0)];
    const int n2 = block_conn[8 * owner + tmrgpu::edge_corner(oe, 1)];
    for (int ip = begin; ip < end; ip++) {
      const int b = edge_block_conn[ip];
      for (int e = 0; e < 12; e++) {
        const int m1 = block_conn[8 * b + tmrgpu::edge_corner(e, 0)];
        const int m2 = block_conn[8 * b + tmrgpu::edge_corner(e, 1)];
        if (same_pair(n1, n2, m1, m2)) {
          edge_block_conn[ip] = 12 * b + e;
          break;
        }
      }
    }
  }
}

/* face -> (6*block + local face) and the orientation id of every block face
   relative to its owner face (reference :950-1085) */
void TMROctForest::BlockTables::facesToBlocks() {
  face_block_ptr = new int[num_faces + 1];
  std::fill(face_block_ptr, face_block_ptr + num_faces + 1, 0);
  for (int i = 0; i < 6 * num_blocks; i++) {
    face_block_ptr[block_face_conn[i] + 1]++;
  }
  counts_to_offsets(face_block_ptr, num_faces);
  face_block_conn =
      new int[face_block_ptr[num_faces] > 0 ? face_block_ptr[num_faces] : 1];
  std::vector<int> fill(face_block_ptr, face_block_ptr + num_faces);
  for (int b = 0; b < num_blocks; b++) {
    for (int f = 0; f < 6; f++) {
      face_block_conn[fill[block_face_conn[6 * b + f]]++] = b;
    }
  }
  block_face_ids = new int[6 * num_blocks > 0 ? 6 * num_blocks : 1];
  std::fill(block_face_ids, block_face_ids + 6 * num_blocks, 0);
  for (int face = 0; face < num_faces; face++) {
    const int begin = face_block_ptr[face], end = face_block_ptr[face + 1];
    int owner = num_blocks;
    for (int ip = begin; ip < end; ip++) {
      owner = std::min(owner, face_block_conn[ip]);
    }
    if (owner == num_blocks) continue;
    int of = 0;
    while (of < 6 && block_face_conn[6 * owner + of] != face) of++;
    int on[4];
    for (int k = 0; k < 4; k++) on[k] = block_conn[8 * owner + face_corner(of, k)];
    for (int ip = begin; ip < end; ip++) {
      const int b = face_block_conn[ip];
      int f = 0;
      for (; f < 6; f++) {
        if (block_face_conn[6 * b + f] != face) continue;
        int an[4];
        for (int k = 0; k < 4; k++) an[k] = block_conn[8 * b + face_corner(f, k)];
        bool match = false;
        for (int id = 0; id < 8; id++) {
          match = (on[0] == an[orient(id, 0)] && on[1] == an[orient(id, 1)] &&
                   on[2] == an[orient(id, 2)] && on[3] == an[orient(id, 3)]);
          if (match) {
            block_face_ids[6 * b + f] = id;
            break;
          }
        }
        if (match) break;
      }
      face_block_conn[ip] = 6 * b + f;
    }
  }
}

/* owner = lowest block touching the entity (reference :1091-1143) */
void TMROctForest::BlockTables::entityOwners() {
  face_block_owners = new int[num_faces > 0 ? num_faces : 1];
  edge_block_owners = new int[num_edges > 0 ? num_edges : 1];
  node_block_owners = new int[num_nodes > 0 ? num_nodes : 1];
  for (int f = 0; f < num_faces; f++) {
    int o = num_blocks;
    for (int ip = face_block_ptr[f]; ip < face_block_ptr[f + 1]; ip++) {
      o = std::min(o, face_block_conn[ip] / 6);
    }
    face_block_owners[f] = o;
  }
  for (int e = 0; e < num_edges; e++) {
    int o = num_blocks;
    for (int ip = edge_block_ptr[e]; ip < edge_block_ptr[e + 1]; ip++) {
      o = std::min(o, edge_block_conn[ip] / 12);
    }
    edge_block_owners[e] = o;
  }
  for (int n = 0; n < num_nodes; n++) {
    int o = num_blocks;
    for (int ip = node_block_ptr[n]; ip < node_block_ptr[n + 1]; ip++) {
      o = std::min(o, node_block_conn[ip] / 8);
    }
    node_block_owners[n] = o;
  }
}

/* ---- construction / teardown ------------------------------------------------- */
TMROctForest::TMROctForest(MPI_Comm _comm, int _mesh_order,
                           TMRInterpolationType _interp_type) {
  if (!TMRIsInitialized()) TMRInitialize();
  comm = _comm;
  MPI_Comm_rank(comm, &mpi_rank);
  MPI_Comm_size(comm, &mpi_size);
  mesh_order = 2;
  interp_knots = NULL;
  topo = NULL;
  tables = NULL;
  dev = NULL;
  octants = NULL;
  octants_exposed = 0;
  owners = NULL;
  conn = node_numbers = node_range = NULL;
  dep_ptr = dep_conn = NULL;
  interp_rows = interp_rowp = interp_cols = NULL;
  interp_vals = NULL;
  dep_weights = NULL;
  X = NULL;
  num_local_nodes = num_dep_nodes = num_owned_nodes = ext_pre_offset = 0;
  num_elements_nodes = 0;
  nodes_on_host = nodes_exist = 0;
  setMeshOrder(_mesh_order, _interp_type);
}

TMROctForest::~TMROctForest() {
  if (topo) topo->decref();
  dropTables();
  dropMeshData(1, 1);
  if (dev) tmrgpu_forest_destroy(dev);
  delete[] interp_knots;
  delete[] interp_rows;
  delete[] interp_rowp;
  delete[] interp_cols;
  delete[] interp_vals;
}

int TMROctForest::ensureDevice() {
  if (dev) return 0;
  tmrgpu_ctx *ctx = tmr_b200_context();
  if (!ctx) return 1;
  if (tmrgpu_forest_create(ctx, &dev)) return 1;
  /* a one-rank communicator (MPI_COMM_SELF) inside a multi-rank job: this
     forest is not partitioned, whatever communicator the context holds */
  if (mpi_size == 1) tmrgpu_forest_set_serial(dev, 1);
  return 0;
}

void TMROctForest::dropTables() {
  if (tables) tables->decref();
  tables = NULL;
}

/* The host copies of the node arrays are page-locked buffers owned by the
   device-side forest (tmrgpu_node_mirror): each getter copies only the array
   it hands out, and the pointers stay valid until the node data is freed. */
void TMROctForest::dropHostNodeMirrors() {
  delete[] node_range;
  delete[] X;
  conn = node_numbers = node_range = NULL;
  dep_ptr = dep_conn = NULL;
  dep_weights = NULL;
  X = NULL;
  num_local_nodes = num_dep_nodes = num_owned_nodes = ext_pre_offset = 0;
  nodes_on_host = 0;
}

/* freeMeshData (reference :426-483): node data always goes; octants / owners
   on request */
void TMROctForest::dropMeshData(int drop_octants, int drop_owners) {
  if (drop_owners) {
    delete[] owners;
    owners = NULL;
  }
  if (drop_octants) {
    delete octants;
    octants = NULL;
    octants_exposed = 0;
  }
  dropHostNodeMirrors();
  nodes_exist = 0;
  if (dev) tmrgpu_free_nodes(dev);
}

void TMROctForest::pushTablesToDevice() {
  if (ensureDevice()) return;
  BlockTables *t = tables;
  tmrgpu_set_connectivity(
      dev, t->num_blocks, t->num_nodes, t->num_edges, t->num_faces,
      t->block_conn, t->block_edge_conn, t->block_face_conn, t->block_face_ids,
      t->node_block_ptr, t->node_block_conn, t->edge_block_ptr,
      t->edge_block_conn, t->face_block_ptr, t->face_block_conn,
      t->node_block_owners, t->edge_block_owners, t->face_block_owners);
}

/* ---- topology / connectivity -------------------------------------------------- */
void TMROctForest::setTopology(TMRTopology *_topo) {
  dropTables();
  dropMeshData(1, 1);
  if (_topo) {
    _topo->incref();
    if (topo) topo->decref();
    topo = _topo;
    int nn, ne, nf, nb;
    const int *bc, *bec, *bfc;
    topo->getConnectivity(&nn, &ne, &nf, &nb, &bc, &bec, &bfc);
    setFullConnectivity(nn, ne, nf, nb, bc, bec, bfc);
  }
}

TMRTopology *TMROctForest::getTopology() { return topo; }

void TMROctForest::setConnectivity(int _num_nodes, const int *_block_conn,
                                   int _num_blocks) {
  dropTables();
  dropMeshData(1, 1);
  tables = new BlockTables();
  tables->incref();
  tables->num_nodes = _num_nodes;
  tables->num_blocks = _num_blocks;
  tables->block_conn = copy_ints(_block_conn, 8 * _num_blocks);
  tables->nodesToBlocks();
  tables->edgesFromNodes();
  tables->edgesToBlocks();
  tables->facesFromNodes();
  tables->facesToBlocks();
  tables->entityOwners();
  pushTablesToDevice();
}

void TMROctForest::setFullConnectivity(int _num_nodes, int _num_edges,
                                       int _num_faces, int _num_blocks,
                                       const int *_block_conn,
                                       const int *_block_edge_conn,
                                       const int *_block_face_conn) {
  dropTables();
  dropMeshData(1, 1);
  tables = new BlockTables();
  tables->incref();
  tables->num_nodes = _num_nodes;
  tables->num_edges = _num_edges;
  tables->num_faces = _num_faces;
  tables->num_blocks = _num_blocks;
  tables->block_conn = copy_ints(_block_conn, 8 * _num_blocks);
  tables->nodesToBlocks();
  tables->block_edge_conn = copy_ints(_block_edge_conn, 12 * _num_blocks);
  tables->edgesToBlocks();
  tables->block_face_conn = copy_ints(_block_face_conn, 6 * _num_blocks);
  tables->facesToBlocks();
  tables->entityOwners();
  pushTablesToDevice();
}

/* ---- order / knots (reference :1389-1427) -------------------------------------- */
void TMROctForest::setMeshOrder(int _mesh_order,
                                TMRInterpolationType _interp_type) {
  dropMeshData(0, 0);
  delete[] interp_knots;
  mesh_order = _mesh_order;
  if (mesh_order < 2) mesh_order = 2;
  if (mesh_order > MAX_ORDER) mesh_order = MAX_ORDER;
  interp_type = _interp_type;
  interp_knots = new double[2 * mesh_order];
  std::fill(interp_knots, interp_knots + 2 * mesh_order, 0.0);
  interp_knots[0] = -1.0;
  interp_knots[mesh_order - 1] = 1.0;
  for (int i = 1; i < mesh_order - 1; i++) {
    if (interp_type == TMR_UNIFORM_POINTS) {
      interp_knots[i] = -1.0 + 2.0 * i / (mesh_order - 1);
    } else {
      interp_knots[i] = -cos(M_PI * i / (mesh_order - 1));
    }
  }
}

int TMROctForest::getMeshOrder() { return mesh_order; }
TMRInterpolationType TMROctForest::getInterpType() { return interp_type; }

int TMROctForest::getInterpKnots(const double **_knots) {
  if (_knots) *_knots = interp_knots;
  return mesh_order;
}

/* ---- octant mirror management -------------------------------------------------- */
int TMROctForest::syncOctantsToDevice() {
  if (ensureDevice()) return 1;
  if (octants && octants_exposed) {
    /* the caller holds a mutable pointer into the mirror (Python writes
       through it, reference tmr/TMR.pyx:3303-3317): re-upload */
    TMROctant *a;
    int n;
    octants->getArray(&a, &n);
    const int rc = tmrgpu_upload_octants(
        dev, reinterpret_cast<const tmrgpu_octant *>(a), n);
    octants_exposed = 0;
    return rc;
  }
  return 0;
}

void TMROctForest::octantsReplacedOnDevice() {
  delete octants;
  octants = NULL;
  octants_exposed = 0;
}

void TMROctForest::getOctants(TMROctantArray **_octants) {
  if (!octants && dev && tables) {
    const int n = (int)tmrgpu_count(dev);
    TMROctant *a = new TMROctant[n > 0 ? n : 1];
    tmrgpu_download_octants(dev, reinterpret_cast<tmrgpu_octant *>(a));
    octants = new TMROctantArray(a, n);
  }
  if (octants) octants_exposed = 1;
  if (_octants) *_octants = octants;
}

/* ---- tree creation -------------------------------------------------------------- */
static void rank_block_range(int num_blocks, int rank, int size, int *start,
                             int *end) {
  /* contiguous deal of trees to ranks (reference :1758-1770) */
  const int base = num_blocks / size, extra = num_blocks % size;
  *start = rank * base + std::min(rank, extra);
  *end = *start + base + (rank < extra ? 1 : 0);
}

void TMROctForest::createTrees(int refine_level) {
  if (!tables || ensureDevice()) {
    fprintf(stderr,
            "TMROctForest Error: Cannot call createTrees(), no connectivity "
            "has been set\n");
    return;
  }
  dropMeshData(1, 1);
  int start, end;
  rank_block_range(tables->num_blocks, mpi_rank, mpi_size, &start, &end);
  tmrgpu_create_trees(dev, refine_level, start, end);
}

void TMROctForest::createRandomTrees(int nrand, int min_level, int max_level) {
  if (!tables || ensureDevice()) {
    fprintf(stderr,
            "TMROctForest Error: Cannot call createRandomTrees(), no "
            "connectivity has been set\n");
    return;
  }
  dropMeshData(1, 1);
  int start, end;
  rank_block_range(tables->num_blocks, mpi_rank, mpi_size, &start, &end);
  /* same rand() draw order as the reference (:1863-1881), so the same libc
     seed gives the same forest */
  const int size = nrand * (end - start);
  std::vector<tmrgpu_octant> recs(size > 0 ? size : 1);
  for (int count = 0, block = start; block < end; block++) {
    for (int i = 0; i < nrand; i++, count++) {
      const int32_t level = min_level + (rand() % (max_level - min_level + 1));
      const int32_t h = 1 << (TMR_MAX_LEVEL - level);
      tmrgpu_octant &o = recs[count];
      o.x = h * (rand() % (1 << level));
      o.y = h * (rand() % (1 << level));
      o.z = h * (rand() % (1 << level));
      o.tag = 0;
      o.block = block;
      o.level = (int16_t)level;
      o.info = 0;
    }
  }
  if (tmrgpu_upload_octants(dev, recs.data(), size) == 0) {
    tmrgpu_sort_unique(dev);
  }
}

/* ---- repartition (reference :1922-2088) ------------------------------------------ */
void TMROctForest::repartition(int max_rank) {
  if (!tables || !dev) {
    fprintf(stderr,
            "TMROctForest Error: Cannot call repartition(), no octants have "
            "been created\n");
    return;
  }
  dropMeshData(0, 1);
  if (syncOctantsToDevice()) return;
  if (mpi_size > 1) {
    /* equal-count re-split over NCCL; on one rank nothing moves */
    if (tmrgpu_repartition(dev, max_rank) == 0) octantsReplacedOnDevice();
  }
}

/* ---- duplicate / coarsen (reference :2097-2164) ----------------------------------- */
TMROctForest *TMROctForest::duplicate() {
  TMROctForest *dup = new TMROctForest(comm, mesh_order, interp_type);
  if (tables && dev) {
    syncOctantsToDevice();
    tables->incref();
    dup->tables = tables;
    dup->topo = topo;
    if (topo) topo->incref();
    if (dup->ensureDevice() == 0) tmrgpu_duplicate(dev, dup->dev);
  }
  return dup;
}

TMROctForest *TMROctForest::coarsen() {
  TMROctForest *coarse = new TMROctForest(comm, mesh_order, interp_type);
  if (tables && dev) {
    syncOctantsToDevice();
    tables->incref();
    coarse->tables = tables;
    coarse->topo = topo;
    if (topo) topo->incref();
    if (coarse->ensureDevice() == 0) tmrgpu_coarsen(dev, coarse->dev);
  }
  return coarse;
}

/* ---- refine / balance -------------------------------------------------------------- */
void TMROctForest::refine(const int refinement[], int min_level,
                          int max_level) {
  if (!tables || !dev) {
    fprintf(stderr,
            "TMROctForest Error: Cannot call refine(), no octants have been "
            "created\n");
    return;
  }
  dropMeshData(0, 0);
  if (syncOctantsToDevice()) return;
  if (tmrgpu_refine(dev, refinement, min_level, max_level) == 0) {
    octantsReplacedOnDevice();
  }
}

void TMROctForest::balance(int balance_corner) {
  if (!tables || !dev) {
    fprintf(stderr,
            "TMROctForest Error: Cannot call balance(), no octants have been "
            "created\n");
    return;
  }
  /* note: like the reference (:2917), balance() does NOT invalidate node data */
  if (syncOctantsToDevice()) return;
  if (tmrgpu_balance(dev, balance_corner) == 0) {
    octantsReplacedOnDevice();
  }
}

/* ---- nodes ---------------------------------------------------------------------------- */
void TMROctForest::createNodes() {
  if (!tables || !dev) {
    fprintf(stderr,
            "TMROctForest Error: Cannot call createNodes(), no octants have "
            "been created\n");
    return;
  }
  if (nodes_exist) return; /* reference :4071-4075 */
  if (syncOctantsToDevice()) return;
  if (tmrgpu_create_nodes(dev, mesh_order, (int)interp_type, interp_knots)) {
    return;
  }
  nodes_exist = 1;
  nodes_on_host = 0;
  /* createNodes rewrites the info field of every octant
     (computeDepFacesAndEdges, reference :3619); keep a live mirror in step */
  if (octants) {
    TMROctant *a;
    int n;
    octants->getArray(&a, &n);
    std::vector<int16_t> info(n > 0 ? n : 1);
    tmrgpu_download_info(dev, info.data());
    for (int i = 0; i < n; i++) a[i].info = info[i];
  }
}

/* sizes and the owned-node ranges: a few integers, no array is copied */
void TMROctForest::fetchNodeData() {
  if (nodes_on_host || !nodes_exist || !dev) return;
  int64_t s[6];
  if (tmrgpu_node_sizes(dev, s)) return;
  dropHostNodeMirrors();
  num_elements_nodes = (int)s[0];
  num_local_nodes = (int)s[1];
  num_dep_nodes = (int)s[2];
  num_owned_nodes = (int)s[3];
  /* node_range: owned-node prefix over ranks (reference :4165-4172) */
  node_range = new int[mpi_size + 1];
  tmrgpu_node_range(dev, node_range);
  ext_pre_offset = -1; /* needs the sorted node numbers: see fetchNodeNumbers */
  nodes_on_host = 1;
}

template <class T>
static T *fetch_mirror(tmrgpu_forest *dev, int which) {
  const void *p = NULL;
  if (!dev || tmrgpu_node_mirror(dev, which, &p)) return NULL;
  return const_cast<T *>(static_cast<const T *>(p));
}

void TMROctForest::fetchNodeNumbers() {
  fetchNodeData();
  if (node_numbers || !nodes_on_host) return;
  /* the reference hands out node_numbers sorted ascending (:4246) */
  node_numbers = fetch_mirror<int>(dev, 1);
  if (node_numbers) {
    int *item = std::lower_bound(node_numbers, node_numbers + num_local_nodes,
                                 node_range[mpi_rank]);
    ext_pre_offset = (int)(item - node_numbers);
  }
}

void TMROctForest::getNodeConn(const int **_conn, int *_num_elements,
                               int *_num_owned_nodes, int *_num_local_nodes) {
  fetchNodeData();
  if (nodes_on_host && !conn) conn = fetch_mirror<int>(dev, 0);
  (void)_num_local_nodes; /* never written by the reference (:5686-5704) */
  int nelems = 0;
  if (dev) nelems = (int)tmrgpu_count(dev);
  if (_conn) *_conn = conn;
  if (_num_elements) *_num_elements = nelems;
  if (_num_owned_nodes) *_num_owned_nodes = num_owned_nodes;
}

int TMROctForest::getDepNodeConn(const int **_ptr, const int **_conn,
                                 const double **_weights) {
  fetchNodeData();
  if (nodes_on_host && !dep_ptr) {
    dep_ptr = fetch_mirror<int>(dev, 2);
    dep_conn = fetch_mirror<int>(dev, 3);
    dep_weights = fetch_mirror<double>(dev, 4);
  }
  if (_ptr) *_ptr = dep_ptr;
  if (_conn) *_conn = dep_conn;
  if (_weights) *_weights = dep_weights;
  return num_dep_nodes;
}

int TMROctForest::getOwnedNodeRange(const int **_node_range) {
  fetchNodeData();
  if (_node_range) *_node_range = node_range;
  return mpi_size;
}

int TMROctForest::getNodeNumbers(const int **_node_numbers) {
  fetchNodeNumbers();
  if (_node_numbers) *_node_numbers = node_numbers;
  return num_local_nodes;
}

int TMROctForest::getExtPreOffset() {
  fetchNodeNumbers();
  return ext_pre_offset;
}

/* Node locations (reference evaluateNodeLocations :5524-5675).  Without a
   topology they are zero, as in the reference.  With one, every local node is
   evaluated through the first element (in element order) and local slot that
   reference it, at u + 0.5 d (1 + knot[i]): on the GPU when every volume is a
   TMRTrilinearVolume, else through the volumes' virtual evalPoint on the host,
   the reference's own loop. */
void TMROctForest::evaluateNodeLocations() {
  X = new TMRPoint[num_local_nodes + 1];
  for (int i = 0; i <= num_local_nodes; i++) X[i].zero();
  if (!topo || num_local_nodes == 0 || !tables) return;
  const int nb = tables->num_blocks;
  std::vector<TMRVolume *> vols(nb, (TMRVolume *)NULL);
  bool trilinear = true;
  for (int b = 0; b < nb; b++) {
    topo->getVolume(b, &vols[b]);
    trilinear = trilinear && dynamic_cast<TMRTrilinearVolume *>(vols[b]) != NULL;
  }
  if (trilinear && !(interp_type == TMR_BERNSTEIN_POINTS && mesh_order > 2)) {
    std::vector<double> corners((size_t)nb * 24);
    for (int b = 0; b < nb; b++) {
      const double *c = static_cast<TMRTrilinearVolume *>(vols[b])->corners();
      std::copy(c, c + 24, corners.begin() + (size_t)b * 24);
    }
    if (tmrgpu_eval_trilinear_points(dev, corners.data(),
                                     reinterpret_cast<double *>(X)) == 0) {
      return;
    }
    fprintf(stderr, "TMROctForest Error: node locations could not be "
                    "evaluated on the device\n");
    return;
  }
  /* general volumes: the reference's host loop */
  TMROctant *octs;
  int num_elements;
  TMROctantArray *elems = NULL;
  getOctants(&elems); /* materialises the host mirror of the elements */
  if (!elems) return;
  elems->getArray(&octs, &num_elements);
  const int *c0;
  getNodeConn(&c0);
  fetchNodeNumbers();
  if (!c0 || !node_numbers) return;
  const int p = mesh_order, size = p * p * p;
  const double *knots = interp_knots;
  std::vector<char> flags(num_local_nodes, 0);
  const bool bern = (interp_type == TMR_BERNSTEIN_POINTS && p > 2);
  std::vector<double> inverse;
  if (bern) {
    /* interpolation matrix of the Bernstein basis at the knot points and its
       inverse (reference :5539-5577, there through LAPACK dgetrf/dgetrs) */
    std::vector<double> A((size_t)size * size);
    for (int i = 0; i < size; i++) {
      double pt[3];
      pt[0] = knots[i % p];
      pt[1] = knots[(i % p * p) / p]; /* as in the reference (:5546) */
      pt[2] = knots[i / (p * p)];
      evalInterp(pt, &A[(size_t)size * i]);
    }
    inverse.assign((size_t)size * size, 0.0);
    for (int i = 0; i < size; i++) inverse[(size_t)(size + 1) * i] = 1.0;
    /* column-major view of the row-major matrix = its transpose, as the
       reference passes it: solve A^T Y = I by Gaussian elimination with
       partial pivoting */
    std::vector<double> M((size_t)size * size);
    for (int r = 0; r < size; r++) {
      for (int c = 0; c < size; c++) M[(size_t)r * size + c] = A[(size_t)c * size + r];
    }
    for (int k = 0; k < size; k++) {
      int piv = k;
      for (int r = k + 1; r < size; r++) {
        if (fabs(M[(size_t)r * size + k]) > fabs(M[(size_t)piv * size + k])) piv = r;
      }
      if (piv != k) {
        for (int c = 0; c < size; c++) {
          std::swap(M[(size_t)k * size + c], M[(size_t)piv * size + c]);
          std::swap(inverse[(size_t)c * size + k], inverse[(size_t)c * size + piv]);
        }
      }
      for (int r = k + 1; r < size; r++) {
        const double fct = M[(size_t)r * size + k] / M[(size_t)k * size + k];
        if (fct == 0.0) continue;
        for (int c = k; c < size; c++) M[(size_t)r * size + c] -= fct * M[(size_t)k * size + c];
        for (int c = 0; c < size; c++) {
          inverse[(size_t)c * size + r] -= fct * inverse[(size_t)c * size + k];
        }
      }
    }
    for (int c = 0; c < size; c++) { /* back substitution, column c of Y */
      for (int r = size - 1; r >= 0; r--) {
        double v = inverse[(size_t)c * size + r];
        for (int q = r + 1; q < size; q++) v -= M[(size_t)r * size + q] * inverse[(size_t)c * size + q];
        inverse[(size_t)c * size + r] = v / M[(size_t)r * size + r];
      }
    }
  }
  std::vector<TMRPoint> Xtmp(size);
  for (int i = 0; i < num_elements; i++) {
    TMRVolume *vol = vols[octs[i].block];
    const int32_t h = 1 << (TMR_MAX_LEVEL - octs[i].level);
    const double d = tmrgpu::param_coordinate(h);
    const double u = tmrgpu::param_coordinate(octs[i].x);
    const double v = tmrgpu::param_coordinate(octs[i].y);
    const double w = tmrgpu::param_coordinate(octs[i].z);
    const int *c = &c0[(size_t)size * i];
    if (bern) {
      for (int kk = 0; kk < p; kk++) {
        for (int jj = 0; jj < p; jj++) {
          for (int ii = 0; ii < p; ii++) {
            vol->evalPoint(u + 0.5 * d * (1.0 + knots[ii]), v + 0.5 * d * (1.0 + knots[jj]),
                           w + 0.5 * d * (1.0 + knots[kk]), &Xtmp[ii + jj * p + kk * p * p]);
          }
        }
      }
    }
    for (int kk = 0; kk < p; kk++) {
      for (int jj = 0; jj < p; jj++) {
        for (int ii = 0; ii < p; ii++) {
          const int local = ii + jj * p + kk * p * p;
          const int index = getLocalNodeNumber(c[local]);
          if (index < 0 || flags[index]) continue;
          flags[index] = 1;
          if (bern) {
            X[index].zero();
            for (int j = 0; j < size; j++) {
              X[index].x += inverse[(size_t)local * size + j] * Xtmp[j].x;
              X[index].y += inverse[(size_t)local * size + j] * Xtmp[j].y;
              X[index].z += inverse[(size_t)local * size + j] * Xtmp[j].z;
            }
          } else {
            vol->evalPoint(u + 0.5 * d * (1.0 + knots[ii]), v + 0.5 * d * (1.0 + knots[jj]),
                           w + 0.5 * d * (1.0 + knots[kk]), &X[index]);
          }
        }
      }
    }
  }
}

int TMROctForest::getPoints(TMRPoint **_X) {
  fetchNodeData();
  if (!X && nodes_on_host) evaluateNodeLocations();
  if (_X) *_X = X;
  return num_local_nodes;
}

int TMROctForest::getLocalNodeNumber(int node) {
  fetchNodeNumbers();
  if (node_numbers) {
    int *end = node_numbers + num_local_nodes;
    int *item = std::lower_bound(node_numbers, end, node);
    if (item != end && *item == node) return (int)(item - node_numbers);
  }
  return -1;
}

/* ---- interpolation ---------------------------------------------------------------------- */
int TMROctForest::createInterpolationCSR(TMROctForest *coarse, const int **rows,
                                         const int **rowp, const int **cols,
                                         const double **vals) {
  if (rows) *rows = NULL;
  if (rowp) *rowp = NULL;
  if (cols) *cols = NULL;
  if (vals) *vals = NULL;
  createNodes();
  coarse->createNodes();
  if (!dev || !coarse->dev || !nodes_exist || !coarse->nodes_exist) return 0;
  int64_t nrows = 0, nnz = 0;
  if (tmrgpu_create_interp(dev, coarse->dev, &nrows, &nnz)) return 0;
  delete[] interp_rows;
  delete[] interp_rowp;
  delete[] interp_cols;
  delete[] interp_vals;
  interp_rows = new int[nrows + 1];
  interp_rowp = new int[nrows + 2];
  interp_cols = new int[nnz + 1];
  interp_vals = new double[nnz + 1];
  /* fresh pages: fault them in from several threads before the copy lands
     (a 1.5e9-entry prolongation is 18 GB of first touches) */
  {
    const int T = 8;
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) {
      th.push_back(std::thread([=]() {
        const int64_t a = nnz * t / T, b = nnz * (t + 1) / T;
        memset(interp_cols + a, 0, (size_t)(b - a) * sizeof(int));
        memset(interp_vals + a, 0, (size_t)(b - a) * sizeof(double));
        const int64_t c = nrows * t / T, d = nrows * (t + 1) / T;
        memset(interp_rows + c, 0, (size_t)(d - c) * sizeof(int));
        memset(interp_rowp + c, 0, (size_t)(d - c) * sizeof(int));
      }));
    }
    for (size_t k = 0; k < th.size(); k++) th[k].join();
  }
  if (tmrgpu_download_interp(dev, interp_rows, interp_rowp, interp_cols,
                             interp_vals)) {
    return 0;
  }
  if (rows) *rows = interp_rows;
  if (rowp) *rowp = interp_rowp;
  if (cols) *cols = interp_cols;
  if (vals) *vals = interp_vals;
  return (int)nrows;
}

void TMROctForest::createInterpolation(TMROctForest *coarse,
                                       TACSBVecInterp *interp) {
  const int *rows, *rowp, *cols;
  const double *vals;
  const int nrows = createInterpolationCSR(coarse, &rows, &rowp, &cols, &vals);
  /* same call stream as the reference's loop (:6683): one addInterp per owned
     fine node, in first-touch order */
  for (int r = 0; r < nrows; r++) {
    interp->addInterp(rows[r], &vals[rowp[r]], &cols[rowp[r]],
                      rowp[r + 1] - rowp[r]);
  }
}

TMROctant *TMROctForest::findEnclosing(const int order, const double *knots,
                                       TMROctant *node, int *mpi_owner) {
  if (mpi_owner) *mpi_owner = mpi_rank;
  if (!dev || syncOctantsToDevice()) return NULL;
  int index = -1;
  if (tmrgpu_find_enclosing(dev, order, knots,
                            reinterpret_cast<const tmrgpu_octant *>(node), 1,
                            &index)) {
    return NULL;
  }
  if (index < 0) {
    /* not on this rank: name the rank that owns the node's position
       (reference :6348-6372; src/topology/TMR_TACSTopoCreator.cpp:166-217
       routes on this value) */
    if (mpi_owner && mpi_size > 1) {
      const int32_t hmax = 1 << TMR_MAX_LEVEL;
      const int32_t h = 1 << (TMR_MAX_LEVEL - node->level);
      const int idx[3] = {node->info % order, (node->info % (order * order)) / order,
                          node->info / (order * order)};
      const int32_t base[3] = {node->x, node->y, node->z};
      int32_t c[3];
      for (int a = 0; a < 3; a++) {
        int32_t ci = -1;
        if (idx[a] == 0 || idx[a] == order - 1) {
          ci = base[a] + (idx[a] / (order - 1)) * h;
        } else if (order % 2 == 1 && idx[a] == order / 2) {
          ci = base[a] + h / 2;
        }
        const double cd = base[a] + 0.5 * h * (1.0 + knots[idx[a]]);
        c[a] = ci < 0 ? (int)cd : ci;
        if (c[a] == 0) {
          c[a] += 1;
        } else if (c[a] == hmax) {
          c[a] -= 1;
        }
      }
      TMROctant n;
      n.block = node->block;
      n.x = c[0];
      n.y = c[1];
      n.z = c[2];
      if (!owners) {
        owners = new TMROctant[mpi_size];
        tmrgpu_get_owners(dev, reinterpret_cast<tmrgpu_octant *>(owners));
      }
      int rank = 0; /* getOctantMPIOwner (:2334-2343) */
      while (rank < mpi_size - 1 && owners[rank + 1].comparePosition(&n) <= 0) rank++;
      *mpi_owner = rank;
    }
    return NULL;
  }
  TMROctantArray *arr;
  getOctants(&arr); /* a pointer into the mirror escapes: marks it exposed */
  TMROctant *a;
  arr->getArray(&a, NULL);
  return &a[index];
}

void TMROctForest::transformNode(TMROctant *oct, int edge_dir,
                                 int *edge_reversed, int *src_face_id) {
  if (!tables) return;
  tmrgpu::ConnTables t;
  t.nblocks = tables->num_blocks;
  t.nnodes = tables->num_nodes;
  t.nedges = tables->num_edges;
  t.nfaces = tables->num_faces;
  t.block_conn = tables->block_conn;
  t.block_edge_conn = tables->block_edge_conn;
  t.block_face_conn = tables->block_face_conn;
  t.block_face_ids = tables->block_face_ids;
  t.node_block_ptr = tables->node_block_ptr;
  t.node_block_conn = tables->node_block_conn;
  t.edge_block_ptr = tables->edge_block_ptr;
  t.edge_block_conn = tables->edge_block_conn;
  t.face_block_ptr = tables->face_block_ptr;
  t.face_block_conn = tables->face_block_conn;
  t.node_block_owners = tables->node_block_owners;
  t.edge_block_owners = tables->edge_block_owners;
  t.face_block_owners = tables->face_block_owners;
  tmrgpu::transform_node(t, &oct->block, &oct->x, &oct->y, &oct->z, edge_dir,
                         edge_reversed, src_face_id);
}

/* ---- evalInterp (reference :1508-1620) ---------------------------------------------------- */
namespace {
/* value, first and second derivative of the 1-D Lagrange basis */
void lagrange_all(int order, double u, const double *knots, double *N,
                  double *Nd, double *Ndd) {
  for (int i = 0; i < order; i++) {
    double v = 1.0, d1 = 0.0, d2 = 0.0;
    for (int j = 0; j < order; j++) {
      if (j == i) continue;
      v *= (u - knots[j]) / (knots[i] - knots[j]);
    }
    for (int j = 0; j < order; j++) {
      if (j == i) continue;
      double t = 1.0 / (knots[i] - knots[j]);
      for (int k = 0; k < order; k++) {
        if (k == i || k == j) continue;
        t *= (u - knots[k]) / (knots[i] - knots[k]);
      }
      d1 += t;
      for (int k = 0; k < order; k++) {
        if (k == i || k == j) continue;
        double s = 1.0 / ((knots[i] - knots[j]) * (knots[i] - knots[k]));
        for (int m = 0; m < order; m++) {
          if (m == i || m == j || m == k) continue;
          s *= (u - knots[m]) / (knots[i] - knots[m]);
        }
        d2 += s;
      }
    }
    N[i] = v;
    if (Nd) Nd[i] = d1;
    if (Ndd) Ndd[i] = d2;
  }
}
/* value, first and second derivative of the Bernstein basis of degree
   order-1 on [-1,1]: B' = p/2 (B_{j-1}^{p-1} - B_j^{p-1}) and once more for B''
   (reference src/TMRInterpolation.h:164-300) */
void bernstein_all(int order, double u, double *N, double *Nd, double *Ndd) {
  const int p = order - 1;
  tmrgpu::bernstein_basis(order, u, N);
  if (Nd) {
    double lo[TMROctForest::MAX_ORDER];
    if (p >= 1) tmrgpu::bernstein_basis(order - 1, u, lo);
    for (int j = 0; j < order; j++) {
      const double a = (p >= 1 && j >= 1) ? lo[j - 1] : 0.0;
      const double b = (p >= 1 && j <= p - 1) ? lo[j] : 0.0;
      Nd[j] = 0.5 * p * (a - b);
    }
  }
  if (Ndd) {
    double lo[TMROctForest::MAX_ORDER];
    if (p >= 2) tmrgpu::bernstein_basis(order - 2, u, lo);
    for (int j = 0; j < order; j++) {
      const double a = (p >= 2 && j >= 2) ? lo[j - 2] : 0.0;
      const double b = (p >= 2 && j >= 1 && j <= p - 1) ? lo[j - 1] : 0.0;
      const double c = (p >= 2 && j <= p - 2) ? lo[j] : 0.0;
      Ndd[j] = 0.25 * p * (p - 1) * (a - 2.0 * b + c);
    }
  }
}

void basis_all(int interp_type, int order, double u, const double *knots,
               double *N, double *Nd, double *Ndd) {
  if (interp_type == TMR_BERNSTEIN_POINTS) {
    bernstein_all(order, u, N, Nd, Ndd);
  } else {
    lagrange_all(order, u, knots, N, Nd, Ndd);
  }
}
}  // namespace

void TMROctForest::evalInterp(const double pt[], double N[]) {
  double a[3][MAX_ORDER];
  for (int d = 0; d < 3; d++) {
    if (interp_type == TMR_BERNSTEIN_POINTS) {
      tmrgpu::bernstein_basis(mesh_order, pt[d], a[d]);
    } else {
      tmrgpu::lagrange_basis(mesh_order, pt[d], interp_knots, a[d]);
    }
  }
  for (int k = 0; k < mesh_order; k++) {
    for (int j = 0; j < mesh_order; j++) {
      for (int i = 0; i < mesh_order; i++) *N++ = a[0][i] * a[1][j] * a[2][k];
    }
  }
}

void TMROctForest::evalInterp(const double pt[], double N[], double Nxi[],
                              double Neta[], double Nzeta[]) {
  double a[3][MAX_ORDER], d[3][MAX_ORDER];
  for (int c = 0; c < 3; c++) {
    basis_all(interp_type, mesh_order, pt[c], interp_knots, a[c], d[c], NULL);
  }
  for (int k = 0; k < mesh_order; k++) {
    for (int j = 0; j < mesh_order; j++) {
      for (int i = 0; i < mesh_order; i++) {
        *N++ = a[0][i] * a[1][j] * a[2][k];
        *Nxi++ = d[0][i] * a[1][j] * a[2][k];
        *Neta++ = a[0][i] * d[1][j] * a[2][k];
        *Nzeta++ = a[0][i] * a[1][j] * d[2][k];
      }
    }
  }
}

void TMROctForest::evalInterp(const double pt[], double N[], double N1[],
                              double N2[], double N3[], double N11[],
                              double N22[], double N33[], double N23[],
                              double N13[], double N12[]) {
  double a[3][MAX_ORDER], d[3][MAX_ORDER], s[3][MAX_ORDER];
  for (int c = 0; c < 3; c++) {
    basis_all(interp_type, mesh_order, pt[c], interp_knots, a[c], d[c], s[c]);
  }
  for (int k = 0; k < mesh_order; k++) {
    for (int j = 0; j < mesh_order; j++) {
      for (int i = 0; i < mesh_order; i++) {
        *N++ = a[0][i] * a[1][j] * a[2][k];
        *N1++ = d[0][i] * a[1][j] * a[2][k];
        *N2++ = a[0][i] * d[1][j] * a[2][k];
        *N3++ = a[0][i] * a[1][j] * d[2][k];
        *N11++ = s[0][i] * a[1][j] * a[2][k];
        *N22++ = a[0][i] * s[1][j] * a[2][k];
        *N23++ = a[0][i] * d[1][j] * d[2][k];
        *N13++ = d[0][i] * a[1][j] * d[2][k];
        /* as the reference does (src/TMROctForest.cpp:1611-1613): the mixed
           xy derivative lands in N33 and N12 is never written -- kept for
           drop-in behaviour, not corrected */
        *N33++ = d[0][i] * d[1][j] * a[2][k];
      }
    }
  }
}

/* ---- table getters (reference :1625-1739) -------------------------------------------------- */
void TMROctForest::getConnectivity(int *_nblocks, int *_nfaces, int *_nedges,
                                   int *_nnodes, const int **_block_conn,
                                   const int **_block_face_conn,
                                   const int **_block_edge_conn,
                                   const int **_block_face_ids) {
  BlockTables *t = tables;
  if (_nblocks) *_nblocks = t ? t->num_blocks : 0;
  if (_nfaces) *_nfaces = t ? t->num_faces : 0;
  if (_nedges) *_nedges = t ? t->num_edges : 0;
  if (_nnodes) *_nnodes = t ? t->num_nodes : 0;
  if (_block_conn) *_block_conn = t ? t->block_conn : NULL;
  if (_block_face_conn) *_block_face_conn = t ? t->block_face_conn : NULL;
  if (_block_edge_conn) *_block_edge_conn = t ? t->block_edge_conn : NULL;
  if (_block_face_ids) *_block_face_ids = t ? t->block_face_ids : NULL;
}

void TMROctForest::getInverseConnectivity(const int **_node_block_conn,
                                          const int **_node_block_ptr,
                                          const int **_edge_block_conn,
                                          const int **_edge_block_ptr,
                                          const int **_face_block_conn,
                                          const int **_face_block_ptr) {
  BlockTables *t = tables;
  if (_node_block_conn) *_node_block_conn = t ? t->node_block_conn : NULL;
  if (_node_block_ptr) *_node_block_ptr = t ? t->node_block_ptr : NULL;
  if (_edge_block_conn) *_edge_block_conn = t ? t->edge_block_conn : NULL;
  if (_edge_block_ptr) *_edge_block_ptr = t ? t->edge_block_ptr : NULL;
  if (_face_block_conn) *_face_block_conn = t ? t->face_block_conn : NULL;
  if (_face_block_ptr) *_face_block_ptr = t ? t->face_block_ptr : NULL;
}

/* ---- exchange plumbing (reference :2379-2509) ---------------------------------------------------
   Public entry points for callers that route their own octant lists (e.g.
   reference src/topology/TMR_TACSTopoCreator.cpp:166-217).  The interval
   matching is the reference's (a scan of the sorted list against owners[]);
   the records travel GPU-to-GPU over NCCL. */
TMROctantArray *TMROctForest::distributeOctants(TMROctantArray *list,
                                                int use_tags, int **_oct_ptr,
                                                int **_oct_recv_ptr,
                                                int include_local,
                                                int use_node_index) {
  int size;
  TMROctant *array;
  list->getArray(&array, &size);
  int *oct_ptr = new int[mpi_size + 1];
  int *oct_recv_ptr = new int[mpi_size + 1];
  oct_ptr[0] = 0;
  if (use_tags) {
    /* matchTagIntervals (:2365-2374): the list is sorted by destination tag */
    for (int i = 0, rank = 0; rank < mpi_size; rank++) {
      while (i < size && array[i].tag <= rank) i++;
      oct_ptr[rank + 1] = i;
    }
  } else {
    /* matchOctantIntervals (:2348-2360): sorted by position, cut at owners[] */
    if (mpi_size > 1 && !owners && dev) {
      owners = new TMROctant[mpi_size];
      tmrgpu_get_owners(dev, reinterpret_cast<tmrgpu_octant *>(owners));
    }
    int index = 0;
    for (int rank = 0; rank < mpi_size - 1; rank++) {
      while (index < size && owners &&
             owners[rank + 1].comparePosition(&array[index]) > 0) {
        index++;
      }
      oct_ptr[rank + 1] = index;
    }
  }
  oct_ptr[mpi_size] = size;
  std::vector<int> counts(mpi_size), recv_counts(mpi_size);
  for (int i = 0; i < mpi_size; i++) {
    counts[i] = (!include_local && i == mpi_rank) ? 0 : oct_ptr[i + 1] - oct_ptr[i];
  }
  tmrgpu_ctx *ctx = tmr_b200_context();
  if (mpi_size == 1) {
    recv_counts[0] = counts[0];
  } else if (ctx) {
    tmrgpu_exchange_counts(ctx, counts.data(), recv_counts.data());
  }
  oct_recv_ptr[0] = 0;
  for (int i = 0; i < mpi_size; i++) {
    oct_recv_ptr[i + 1] = oct_recv_ptr[i] + recv_counts[i];
  }
  TMROctantArray *dist = sendOctants(list, oct_ptr, oct_recv_ptr, use_node_index);
  if (_oct_ptr) {
    *_oct_ptr = oct_ptr;
  } else {
    delete[] oct_ptr;
  }
  if (_oct_recv_ptr) {
    *_oct_recv_ptr = oct_recv_ptr;
  } else {
    delete[] oct_recv_ptr;
  }
  return dist;
}

TMROctantArray *TMROctForest::sendOctants(TMROctantArray *list,
                                          const int *oct_ptr,
                                          const int *oct_recv_ptr,
                                          int use_node_index) {
  int size;
  TMROctant *array;
  list->getArray(&array, &size);
  const int recv_size = oct_recv_ptr[mpi_size];
  TMROctant *recv = new TMROctant[recv_size > 0 ? recv_size : 1];
  /* the local segment only moves when both sides agree on its length
     (reference :2489-2494); the others always do */
  std::vector<int> sp(oct_ptr, oct_ptr + mpi_size + 1);
  std::vector<int> send_ptr(mpi_size + 1, 0), send_counts(mpi_size);
  const int r = mpi_rank;
  const int local_recv = oct_recv_ptr[r + 1] - oct_recv_ptr[r];
  const int local_send = oct_ptr[r + 1] - oct_ptr[r];
  /* pack the segments that actually travel into one staging array */
  std::vector<TMROctant> stage;
  stage.reserve(size);
  for (int i = 0; i < mpi_size; i++) {
    int cnt = oct_ptr[i + 1] - oct_ptr[i];
    if (i == r) cnt = (local_recv > 0 && local_recv == local_send) ? local_send : 0;
    send_ptr[i + 1] = send_ptr[i] + cnt;
    for (int k = 0; k < cnt; k++) stage.push_back(array[oct_ptr[i] + k]);
  }
  std::vector<int> recv_ptr(oct_recv_ptr, oct_recv_ptr + mpi_size + 1);
  if (!(local_recv > 0 && local_recv == local_send)) {
    /* nothing arrives in the local slot: shift the later segments down */
    for (int i = r + 1; i <= mpi_size; i++) recv_ptr[i] -= local_recv;
  }
  tmrgpu_ctx *ctx = tmr_b200_context();
  if (mpi_size == 1) {
    if (recv_ptr[1] > 0) {
      memcpy(&recv[oct_recv_ptr[0]], stage.data(), (size_t)recv_ptr[1] * sizeof(TMROctant));
    }
  } else if (ctx) {
    std::vector<TMROctant> tmp(recv_ptr[mpi_size] > 0 ? recv_ptr[mpi_size] : 1);
    tmrgpu_exchange_records(
        ctx, reinterpret_cast<const tmrgpu_octant *>(stage.data()),
        send_ptr.data(), reinterpret_cast<tmrgpu_octant *>(tmp.data()),
        recv_ptr.data());
    for (int i = 0; i < mpi_size; i++) {
      const int cnt = recv_ptr[i + 1] - recv_ptr[i];
      if (cnt > 0) {
        memcpy(&recv[oct_recv_ptr[i]], &tmp[recv_ptr[i]],
               (size_t)cnt * sizeof(TMROctant));
      }
    }
  }
  return new TMROctantArray(recv, recv_size, use_node_index);
}

/* ---- name queries -------------------------------------------------------------
   Boundary conditions and material regions are attached through the names of the
   topology's volumes, faces, edges and vertices.  Both queries are host loops
   over the octant mirror (and, for nodes, the conn mirror) like the reference's;
   what an octant touches follows from its coordinates alone. */
namespace {
/* the reference's test: a NULL query matches unnamed entities */
inline bool name_matches(const TMREntity *e, const char *name) {
  const char *n = e ? e->getName() : NULL;
  return name ? (n && strcmp(n, name) == 0) : (n == NULL);
}
/* which sides of its tree an octant touches: lo[a] / hi[a] per axis */
struct TreeSides {
  bool lo[3], hi[3];
  TreeSides(const TMROctant &o) {
    const int32_t hmax = 1 << TMR_MAX_LEVEL;
    const int32_t h = 1 << (TMR_MAX_LEVEL - o.level);
    const int32_t p[3] = {o.x, o.y, o.z};
    for (int a = 0; a < 3; a++) {
      lo[a] = (p[a] == 0);
      hi[a] = (p[a] + h == hmax);
    }
  }
  bool on(int a) const { return lo[a] || hi[a]; }
  bool side(int a, int s) const { return s ? hi[a] : lo[a]; }
};
}  // namespace

/* reference :5747-5862: octants of a named volume as they are; otherwise one
   copy per named tree face the octant touches, info = the face index */
TMROctantArray *TMROctForest::getOctsWithName(const char *name) {
  if (!topo) {
    fprintf(stderr,
            "TMROctForest Error: Must define topology to use "
            "getOctsWithName()\n");
    return NULL;
  }
  TMROctantArray *local = NULL;
  if (tables) getOctants(&local);
  if (!local) {
    fprintf(stderr,
            "TMROctForest: Must create octants to use getOctsWithName()\n");
    return NULL;
  }
  TMROctant *array;
  int size;
  local->getArray(&array, &size);
  std::vector<TMROctant> found;
  for (int i = 0; i < size; i++) {
    const TMROctant &o = array[i];
    TMRVolume *vol = NULL;
    topo->getVolume(o.block, &vol);
    if (name_matches(vol, name)) {
      found.push_back(o);
      continue;
    }
    const TreeSides t(o);
    for (int a = 0; a < 3; a++) {
      for (int s = 0; s < 2; s++) {
        /* a root octant touches both sides; below the root the reference
           looks at the low side first and never at both (:5808-5853) */
        if (!t.side(a, s) || (o.level > 0 && s == 1 && t.lo[a])) continue;
        const int face_index = 2 * a + s;
        TMRFace *face = NULL;
        topo->getFace(tables->block_face_conn[6 * o.block + face_index], &face);
        if (name_matches(face, name)) {
          found.push_back(o);
          found.back().info = face_index;
        }
      }
    }
  }
  TMROctant *out = new TMROctant[found.empty() ? 1 : found.size()];
  if (!found.empty()) memcpy(out, found.data(), found.size() * sizeof(TMROctant));
  return new TMROctantArray(out, (int)found.size());
}

/* reference :5882-6203: sorted unique numbers of the nodes of every local
   element that lie on a named vertex, edge or face of the element's tree.  (A
   matching VOLUME name adds nothing there -- the loop at :6174-6181 never
   advances its counter -- and nothing here.) */
int TMROctForest::getNodesWithName(const char *name, int **_nodes) {
  if (_nodes) *_nodes = NULL;
  if (!topo) {
    fprintf(stderr,
            "TMROctForest Error: Must define topology to use "
            "getNodesWithName()\n");
    return 0;
  }
  const int *c_all = NULL;
  int nelems = 0;
  if (nodes_exist) getNodeConn(&c_all, &nelems);
  if (!c_all) {
    fprintf(stderr,
            "TMROctForest Error: Nodes must be created before calling "
            "getNodesWithName()\n");
    return 0;
  }
  TMROctantArray *local = NULL;
  getOctants(&local);
  TMROctant *octs;
  int size;
  local->getArray(&octs, &size);
  const int p = mesh_order, p2 = p * p, npe = p * p * p;
  std::vector<int> list;
  for (int i = 0; i < size; i++) {
    const TMROctant &o = octs[i];
    const TreeSides t(o);
    const int non = (t.on(0) ? 1 : 0) + (t.on(1) ? 1 : 0) + (t.on(2) ? 1 : 0);
    if (non == 0) continue;
    const int *c = &c_all[(size_t)npe * o.tag];
    const int stride[3] = {1, p, p2};
    if (non == 3) {
      for (int v = 0; v < 8; v++) {
        if (!(t.side(0, v & 1) && t.side(1, (v >> 1) & 1) && t.side(2, v >> 2))) continue;
        TMRVertex *vert = NULL;
        topo->getVertex(tables->block_conn[8 * o.block + v], &vert);
        if (name_matches(vert, name)) {
          list.push_back(c[(p - 1) * ((v & 1) + p * ((v >> 1) & 1) + p2 * (v >> 2))]);
        }
      }
    }
    if (non >= 2) {
      for (int e = 0; e < 12; e++) {
        /* edge e runs along axis e/4; (e & 1, (e >> 1) & 1) are the sides of the
           other two axes in ascending axis order (reference edge numbering) */
        const int along = e / 4;
        const int a1 = (along == 0) ? 1 : 0, a2 = (along == 2) ? 1 : 2;
        const int s1 = e & 1, s2 = (e >> 1) & 1;
        if (!(t.side(a1, s1) && t.side(a2, s2))) continue;
        TMREdge *edge = NULL;
        topo->getEdge(tables->block_edge_conn[12 * o.block + e], &edge);
        if (!name_matches(edge, name)) continue;
        const int base = (p - 1) * (s1 * stride[a1] + s2 * stride[a2]);
        for (int q = 0; q < p; q++) list.push_back(c[base + q * stride[along]]);
      }
    }
    for (int f = 0; f < 6; f++) {
      const int a = f / 2;
      if (!t.side(a, f & 1)) continue;
      TMRFace *face = NULL;
      topo->getFace(tables->block_face_conn[6 * o.block + f], &face);
      if (!name_matches(face, name)) continue;
      const int a1 = (a == 0) ? 1 : 0, a2 = (a == 2) ? 1 : 2;
      const int base = (p - 1) * (f & 1) * stride[a];
      for (int r = 0; r < p; r++) {
        for (int q = 0; q < p; q++) list.push_back(c[base + q * stride[a1] + r * stride[a2]]);
      }
    }
  }
  std::sort(list.begin(), list.end());
  list.erase(std::unique(list.begin(), list.end()), list.end());
  int *out = new int[list.empty() ? 1 : list.size()];
  if (!list.empty()) memcpy(out, list.data(), list.size() * sizeof(int));
  if (_nodes) {
    *_nodes = out;
  } else {
    delete[] out;
  }
  return (int)list.size();
}

/* ---- writers (reference :1149-1384): plain-text dumps of the super-mesh and of
   the forest, for visual checks.  They need nothing from the CAD layer beyond
   TMRVolume::evalPoint, so they work with any topology; the text is the
   reference's, byte for byte. */
namespace {
/* VTK / Tecplot brick order of a tree's (or an octant's) 8 corners */
const int kBrick[8] = {0, 1, 3, 2, 4, 5, 7, 6};

void vtk_cells(FILE *fp, int ncells, const int *conn /* NULL: 8 k + c */, int base) {
  fprintf(fp, "\nCELLS %d %d\n", ncells, 9 * ncells);
  for (int k = 0; k < ncells; k++) {
    fprintf(fp, "8");
    for (int c = 0; c < 8; c++) {
      fprintf(fp, " %d", (conn ? conn[8 * k + kBrick[c]] : 8 * k + kBrick[c]) + base);
    }
    fprintf(fp, "\n");
  }
  fprintf(fp, "\nCELL_TYPES %d\n", ncells);
  for (int k = 0; k < ncells; k++) fprintf(fp, "12\n");
  fprintf(fp, "CELL_DATA %d\n", ncells);
  fprintf(fp, "SCALARS entity_index float 1\n");
  fprintf(fp, "LOOKUP_TABLE default\n");
}
}  // namespace

/* location of every super-mesh node: the owning tree's volume evaluated at the
   corner the node sits on (reference :1164-1196) */
void TMROctForest::superMeshPoints(std::vector<TMRPoint> &X) {
  X.assign(tables->num_nodes > 0 ? tables->num_nodes : 1, TMRPoint());
  for (int k = 0; k < tables->num_nodes; k++) {
    const int block = tables->node_block_owners[k];
    int corner = 0;
    while (corner < 8 && tables->block_conn[8 * block + corner] != k) corner++;
    TMRVolume *vol = NULL;
    topo->getVolume(block, &vol);
    X[k].zero();
    if (vol) vol->evalPoint(corner & 1 ? 1.0 : 0.0, corner & 2 ? 1.0 : 0.0,
                            corner & 4 ? 1.0 : 0.0, &X[k]);
  }
}

void TMROctForest::writeToVTK(const char *filename) {
  if (mpi_rank != 0 || !topo || !tables) return;
  FILE *fp = fopen(filename, "w");
  if (!fp) return;
  std::vector<TMRPoint> X;
  superMeshPoints(X);
  fprintf(fp, "# vtk DataFile Version 3.0\n");
  fprintf(fp, "vtk output\nASCII\n");
  fprintf(fp, "DATASET UNSTRUCTURED_GRID\n");
  fprintf(fp, "POINTS %d float\n", tables->num_nodes);
  for (int k = 0; k < tables->num_nodes; k++) {
    fprintf(fp, "%e %e %e\n", X[k].x, X[k].y, X[k].z);
  }
  vtk_cells(fp, tables->num_blocks, tables->block_conn, 0);
  for (int k = 0; k < tables->num_blocks; k++) fprintf(fp, "%e\n", 1.0 * k);
  fclose(fp);
}

void TMROctForest::writeToTecplot(const char *filename) {
  if (mpi_rank != 0 || !topo || !tables) return;
  FILE *fp = fopen(filename, "w");
  if (!fp) return;
  const int nn = tables->num_nodes, nb = tables->num_blocks;
  std::vector<TMRPoint> X;
  superMeshPoints(X);
  fprintf(fp, "Variables = X,Y,Z,block\n");
  fprintf(fp, "Zone N = %d E = %d ", nn, nb);
  fprintf(fp, "DATAPACKING=BLOCK, ZONETYPE=FEBRICK\n");
  fprintf(fp, "VARLOCATION = ([4]=CELLCENTERED)\n");
  for (int k = 0; k < nn; k++) fprintf(fp, "%e\n", X[k].x);
  for (int k = 0; k < nn; k++) fprintf(fp, "%e\n", X[k].y);
  for (int k = 0; k < nn; k++) fprintf(fp, "%e\n", X[k].z);
  for (int k = 0; k < nb; k++) fprintf(fp, "%e\n", 1.0 * k);
  for (int k = 0; k < nb; k++) {
    for (int c = 0; c < 8; c++) {
      fprintf(fp, c ? " %d" : "%d", tables->block_conn[8 * k + kBrick[c]] + 1);
    }
    fprintf(fp, "\n");
  }
  fclose(fp);
}

/* every local octant as its own brick, coloured by tree (reference :1317-1384;
   each rank writes the file it is given) */
void TMROctForest::writeForestToVTK(const char *filename) {
  if (!topo || !tables) return;
  TMROctantArray *local = NULL;
  getOctants(&local);
  if (!local) return;
  FILE *fp = fopen(filename, "w");
  if (!fp) return;
  TMROctant *octs;
  int size;
  local->getArray(&octs, &size);
  fprintf(fp, "# vtk DataFile Version 3.0\n");
  fprintf(fp, "vtk output\nASCII\n");
  fprintf(fp, "DATASET UNSTRUCTURED_GRID\n");
  fprintf(fp, "POINTS %d float\n", 8 * size);
  const int32_t hmax = 1 << TMR_MAX_LEVEL;
  for (int k = 0; k < size; k++) {
    const int32_t h = 1 << (TMR_MAX_LEVEL - octs[k].level);
    TMRVolume *vol = NULL;
    topo->getVolume(octs[k].block, &vol);
    for (int c = 0; c < 8; c++) {
      TMRPoint p;
      p.zero();
      vol->evalPoint(1.0 * (octs[k].x + (c & 1) * h) / hmax,
                     1.0 * (octs[k].y + ((c >> 1) & 1) * h) / hmax,
                     1.0 * (octs[k].z + (c >> 2) * h) / hmax, &p);
      fprintf(fp, "%e %e %e\n", p.x, p.y, p.z);
    }
  }
  vtk_cells(fp, size, NULL, 0);
  for (int k = 0; k < size; k++) fprintf(fp, "%e\n", 1.0 * octs[k].block);
  fclose(fp);
}
