/*
  common.h -- types, key formats and per-octant geometry shared by every
  kernel body of the forest hot path.  Everything here is TMR_HD (host+device)
  so that the host-side drop-in class (transformNode) and the kernels run the
  same integer arithmetic.

  Conventions restated from the reference (cited per function):
    * coordinates are int32 in [0, 2^30); h(level) = 1 << (30 - level)
      (reference src/TMRBase.h:37, src/TMROctant.cpp:28-78)
    * element order = (block, Morton with x most significant then y then z,
      level)  (reference src/TMROctant.cpp:171-204)
*/
#ifndef TMRGPU_COMMON_H
#define TMRGPU_COMMON_H

#include <stdint.h>

#if defined(__CUDACC__)
#define TMR_HD __host__ __device__ __forceinline__
#define TMR_UNROLL _Pragma("unroll")
#else
#define TMR_HD inline
#define TMR_UNROLL
#endif

/* Atomics used inside kernel bodies.  In the device pass they are hardware
   atomics; in a host pass (the host-side class sharing these headers, or the
   test-only serial emulation) they degrade to plain read-modify-write. */
#if defined(__CUDA_ARCH__)
#define TMR_ATOMIC_MAX_I32(p, v) atomicMax((int *)(p), (int)(v))
#define TMR_ATOMIC_MIN_I32(p, v) atomicMin((int *)(p), (int)(v))
#define TMR_ATOMIC_OR_I32(p, v) atomicOr((int *)(p), (int)(v))
#define TMR_ATOMIC_OR_U64(p, v) \
  atomicOr((unsigned long long *)(p), (unsigned long long)(v))
#define TMR_ATOMIC_MAX_U64(p, v) \
  atomicMax((unsigned long long *)(p), (unsigned long long)(v))
#define TMR_ATOMIC_MIN_U64(p, v) \
  atomicMin((unsigned long long *)(p), (unsigned long long)(v))
#define TMR_ATOMIC_ADD_U64(p, v) \
  atomicAdd((unsigned long long *)(p), (unsigned long long)(v))
#else
#define TMR_ATOMIC_MAX_I32(p, v) \
  do {                           \
    if (*(p) < (v)) *(p) = (v);  \
  } while (0)
#define TMR_ATOMIC_MIN_I32(p, v) \
  do {                           \
    if (*(p) > (v)) *(p) = (v);  \
  } while (0)
#define TMR_ATOMIC_OR_I32(p, v) \
  do {                          \
    *(p) |= (v);                \
  } while (0)
#define TMR_ATOMIC_OR_U64(p, v) \
  do {                          \
    *(p) |= (v);                \
  } while (0)
#define TMR_ATOMIC_MAX_U64(p, v) \
  do {                           \
    if (*(p) < (v)) *(p) = (v);  \
  } while (0)
#define TMR_ATOMIC_MIN_U64(p, v) \
  do {                           \
    if (*(p) > (v)) *(p) = (v);  \
  } while (0)
#define TMR_ATOMIC_ADD_U64(p, v) \
  do {                           \
    *(p) += (v);                 \
  } while (0)
#endif

namespace tmrgpu {

/* fetch-and-add returning the previous value */
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
unsigned long long fetch_add_u64(unsigned long long *p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  const unsigned long long old = *p;
  *p = old + v;
  return old;
#endif
}

/* append `value` to a list whose length lives at *count: on the device the
   lanes of a warp that append together take ONE atomic (the lowest active lane
   reserves the range), so a list fed by millions of threads does not serialise
   on its counter */
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
void append_u32(unsigned long long *count, uint32_t *list, long long cap, uint32_t value) {
#if defined(__CUDA_ARCH__)
  const unsigned m = __activemask();
  const int lane = (int)(threadIdx.x & 31);
  const int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
  base = __shfl_sync(m, base, leader);
  const unsigned long long pos = base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
  if ((long long)pos < cap) list[pos] = value;
#else
  const unsigned long long pos = (*count)++;
  if ((long long)pos < cap) list[pos] = value;
#endif
}

/* fetch-and-or returning the previous value */
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
int fetch_or_i32(int *p, int v) {
#if defined(__CUDA_ARCH__)
  return atomicOr(p, v);
#else
  const int old = *p;
  *p = old | v;
  return old;
#endif
}

/* fetch-and-add on an int (shared-memory task counters) */
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
int fetch_add_i32(int *p, int v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  const int old = *p;
  *p = old + v;
  return old;
#endif
}

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;
typedef int32_t i32;

static const int kMaxLevel = 30;
static const i32 kHmax = 1 << 30;

/* 24-byte octant record, byte-identical to TMROctant
   (reference src/TMROctant.h:49-53). */
struct Oct24 {
  i32 block, x, y, z, tag;
  int16_t level, info;
};

/* ---- bit interleaving -------------------------------------------------- */
/* spread the low 21 bits of v so that bit i lands on bit 3*i */
TMR_HD u64 spread3(u32 v) {
  u64 x = v & 0x1fffffu;
  x = (x | (x << 32)) & 0x001f00000000ffffULL;
  x = (x | (x << 16)) & 0x001f0000ff0000ffULL;
  x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
  x = (x | (x << 2)) & 0x1249249249249249ULL;
  return x;
}

TMR_HD u32 compact3(u64 x) {
  x &= 0x1249249249249249ULL;
  x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ULL;
  x = (x ^ (x >> 4)) & 0x100f00f00f00f00fULL;
  x = (x ^ (x >> 8)) & 0x001f0000ff0000ffULL;
  x = (x ^ (x >> 16)) & 0x001f00000000ffffULL;
  x = (x ^ (x >> 32)) & 0x1fffffULL;
  return (u32)x;
}

/* x-major Morton code: the comparator in reference src/TMROctant.cpp:189-196
   gives x priority over y over z at equal bit position. */
TMR_HD u64 morton3(u32 x, u32 y, u32 z) {
  return (spread3(x) << 2) | (spread3(y) << 1) | spread3(z);
}

TMR_HD void unmorton3(u64 m, u32 *x, u32 *y, u32 *z) {
  *x = compact3(m >> 2);
  *y = compact3(m >> 1);
  *z = compact3(m);
}

/* ---- element keys -------------------------------------------------------
   key = [ block | 3*D-bit Morton of (x,y,z) >> (30-D) | 5-bit level ]
   D = deepest level present in the array, so every anchor is a multiple of
   h(D) and the truncated Morton code orders exactly like the reference's
   30-bit comparison.  Sorting keys ascending == qsort(compare_octants).   */
struct KeyFmt {
  int D;      /* Morton depth (levels) */
  int bbits;  /* bits reserved for the block index */

  TMR_HD int total_bits() const { return bbits + 3 * D + 5; }
  TMR_HD u64 encode(i32 block, i32 x, i32 y, i32 z, int level) const {
    const int s = kMaxLevel - D;
    u64 m = morton3((u32)x >> s, (u32)y >> s, (u32)z >> s);
    return ((u64)(u32)block << (3 * D + 5)) | (m << 5) | (u64)level;
  }
  TMR_HD void decode(u64 key, i32 *block, i32 *x, i32 *y, i32 *z,
                     int *level) const {
    const int s = kMaxLevel - D;
    *level = (int)(key & 31);
    u64 m = (key >> 5) & ((D > 0) ? ((1ULL << (3 * D)) - 1) : 0ULL);
    u32 ux, uy, uz;
    unmorton3(m, &ux, &uy, &uz);
    *x = (i32)(ux << s);
    *y = (i32)(uy << s);
    *z = (i32)(uz << s);
    *block = (i32)(key >> (3 * D + 5));
  }
  /* key with the level field cleared: equal for octants sharing an anchor */
  TMR_HD u64 position(u64 key) const { return key >> 5; }
};

/* Re-express a key in a format with a different Morton depth (same bbits).
   Only valid when the octant's level <= the new depth. */
TMR_HD u64 rekey(u64 key, int D_old, int D_new) {
  const u64 level = key & 31;
  u64 rest = key >> 5; /* [block | morton_Dold] */
  if (D_new >= D_old) {
    /* block field moves up by 3*(D_new-D_old), morton gets trailing zeros */
    const u64 mmask = (D_old > 0) ? ((1ULL << (3 * D_old)) - 1) : 0ULL;
    u64 m = rest & mmask;
    u64 b = rest >> (3 * D_old);
    return (b << (3 * D_new + 5)) | (m << (3 * (D_new - D_old) + 5)) | level;
  } else {
    const u64 mmask = (1ULL << (3 * D_old)) - 1;
    u64 m = (rest & mmask) >> (3 * (D_old - D_new));
    u64 b = rest >> (3 * D_old);
    return (b << (3 * D_new + 5)) | (m << 5) | level;
  }
}

/* ---- node keys ----------------------------------------------------------
   A node coordinate is a multiple of h(Dn) in [0, 2^30], except that the
   reference clamps 2^30 to 2^30-1 AFTER the inter-tree transform
   (reference src/TMROctForest.cpp:4029-4037).  2^30-1 is all ones, every
   other admissible coordinate has (30-Dn) trailing zeros, so collapsing the
   trailing run into ONE bit preserves the reference's Morton order:
     c' = (c >> (30-Dn)) << 1          for c < 2^30
     c' = 2^(Dn+1) - 1                 for c == 2^30 (before the clamp) or
                                       c == 2^30-1 (after it)
   key = [ block | Morton of (c'x, c'y, c'z), Dn+1 bits per axis ]          */
struct NodeFmt {
  int Dn;
  int bbits;
  /* label bits below the Morton code: 0, or 2 when nodes at one position can
     differ by label (order-3 Bernstein points: node / edge / face / block
     labels, reference initLabel src/TMROctForest.cpp:6798-6811).  The
     reference's compareNode orders by position, then label
     (src/TMROctant.cpp:245-274). */
  int lbits;
  TMR_HD int pos_bits() const { return 3 * (Dn + 1) + lbits; }
  TMR_HD int total_bits() const { return bbits + pos_bits(); }
  TMR_HD u32 squeeze(i32 c) const {
    if (c >= kHmax - 1) return (1u << (Dn + 1)) - 1u;
    return ((u32)c >> (kMaxLevel - Dn)) << 1;
  }
  TMR_HD u64 encode(i32 block, i32 x, i32 y, i32 z, int label = 0) const {
    u64 m = morton3(squeeze(x), squeeze(y), squeeze(z));
    return ((((u64)(u32)block << (3 * (Dn + 1))) | m) << lbits) | (u64)label;
  }
};

/* label of element-local node slot (i,j,k) of an order-3 element when labels
   are in use: 0 corner, 1 edge midpoint, 2 face centre, 3 body centre */
TMR_HD int slot_label(int order, int i, int j, int k) {
  const int hi = order - 1;
  const int ext = ((i == 0 || i == hi) ? 1 : 0) + ((j == 0 || j == hi) ? 1 : 0) +
                  ((k == 0 || k == hi) ? 1 : 0);
  return 3 - ext;
}

/* parametric coordinate of an integer tree coordinate (reference
   convert_to_coordinate, src/TMROctForest.cpp:298-308) */
TMR_HD double param_coordinate(i32 x) {
  if (x == 0) return 0.0;
  if (x == kHmax - 1) return 1.0;
  return 1.0 * x / kHmax;
}

/* trilinear image of (u,v,w) in [0,1]^3 through 8 corner points (corner c:
   bit0 = x, bit1 = y, bit2 = z).  One fixed operation order, so the host class
   (TMRTrilinearVolume::evalPoint), the test oracle's stand-in and the device
   kernel produce the same bits (the library is built with -fmad=false). */
TMR_HD void trilinear_point(const double *X, double u, double v, double w, double *p) {
  const double a[2] = {1.0 - u, u}, b[2] = {1.0 - v, v}, c[2] = {1.0 - w, w};
  p[0] = p[1] = p[2] = 0.0;
  for (int k = 0; k < 8; k++) {
    const double n = (a[k & 1] * b[(k >> 1) & 1]) * c[k >> 2];
    p[0] += n * X[3 * k];
    p[1] += n * X[3 * k + 1];
    p[2] += n * X[3 * k + 2];
  }
}

/* ---- super-mesh connectivity tables (device or host pointers) -----------
   Layout and meaning follow reference src/TMROctForest.h:323-410: inverse
   maps store 8*block+corner / 12*block+edge / 6*block+face with the owner
   (lowest block) first. */
struct ConnTables {
  int nblocks, nnodes, nedges, nfaces;
  const int *block_conn;      /* 8 per block  */
  const int *block_edge_conn; /* 12 per block */
  const int *block_face_conn; /* 6 per block  */
  const int *block_face_ids;  /* 6 per block, orientation id 0..7 vs owner */
  const int *node_block_ptr, *node_block_conn;
  const int *edge_block_ptr, *edge_block_conn;
  const int *face_block_ptr, *face_block_conn;
  const int *node_block_owners, *edge_block_owners, *face_block_owners;
};

/* local corner pairs of the 12 block edges: edges 0-3 run along x, 4-7 along
   y, 8-11 along z (reference src/TMROctForest.cpp:29-31) */
TMR_HD int edge_corner(int edge, int end) {
  /* along x: corners (2*s, 2*s+1); along y: base pattern {0,1,4,5} +{0,2};
     along z: corner s and s+4 */
  if (edge < 4) return 2 * edge + end;
  if (edge < 8) {
    const int s = edge - 4;
    return (s & 1) + 4 * (s >> 1) + 2 * end;
  }
  return (edge - 8) + 4 * end;
}

/* ---- face orientation transforms ----------------------------------------
   Orientation ids 0..7 of a tree face relative to the owner face
   (reference src/TMROctForest.cpp:69-143).  Each id is "optionally swap the
   two in-face axes, then optionally mirror each": packed as three bit masks
   indexed by id.  M is the mirror length: hmax-h for octant anchors,
   hmax for nodes, 0 for direction vectors. */
TMR_HD void face_to_owner(int id, i32 M, i32 a, i32 b, i32 *u, i32 *v) {
  const int swap = (0x5A >> id) & 1;
  const int fu = (0xC6 >> id) & 1;
  const int fv = (0x6C >> id) & 1;
  const i32 p = swap ? b : a;
  const i32 q = swap ? a : b;
  *u = fu ? M - p : p;
  *v = fv ? M - q : q;
}

TMR_HD void owner_to_face(int id, i32 M, i32 u, i32 v, i32 *a, i32 *b) {
  const int swap = (0x5A >> id) & 1;
  const int fu = (0xC6 >> id) & 1;
  const int fv = (0x6C >> id) & 1;
  const i32 p = fu ? M - u : u;
  const i32 q = fv ? M - v : v;
  *a = swap ? q : p;
  *b = swap ? p : q;
}

/* pick / place the two in-face coordinates of face `f` (0,1: x-faces use
   (y,z); 2,3: y-faces use (x,z); 4,5: z-faces use (x,y)) */
TMR_HD void face_pick(int f, i32 x, i32 y, i32 z, i32 *a, i32 *b) {
  if (f < 2) {
    *a = y;
    *b = z;
  } else if (f < 4) {
    *a = x;
    *b = z;
  } else {
    *a = x;
    *b = y;
  }
}

TMR_HD void face_place(int f, i32 n, i32 a, i32 b, i32 *x, i32 *y, i32 *z) {
  if (f < 2) {
    *x = n;
    *y = a;
    *z = b;
  } else if (f < 4) {
    *x = a;
    *y = n;
    *z = b;
  } else {
    *x = a;
    *y = b;
    *z = n;
  }
}

/* place a point on block edge `e`: u along the edge, the two transverse
   coordinates at 0 or T according to the edge's side bits */
TMR_HD void edge_place(int e, i32 u, i32 T, i32 *x, i32 *y, i32 *z) {
  if (e < 4) {
    *x = u;
    *y = T * (e & 1);
    *z = T * (e >> 1);
  } else if (e < 8) {
    *x = T * (e & 1);
    *y = u;
    *z = T * ((e - 4) >> 1);
  } else {
    *x = T * (e & 1);
    *y = T * ((e - 8) >> 1);
    *z = u;
  }
}

/* do blocks b0 (local edge e0) and b1 (local edge e1) traverse their shared
   edge in opposite directions?  (reference src/TMROctForest.cpp:2660-2667) */
TMR_HD int edge_is_reversed(const ConnTables &t, int b0, int e0, int b1,
                            int e1) {
  const int n1 = t.block_conn[8 * b0 + edge_corner(e0, 0)];
  const int n2 = t.block_conn[8 * b0 + edge_corner(e0, 1)];
  const int m1 = t.block_conn[8 * b1 + edge_corner(e1, 0)];
  const int m2 = t.block_conn[8 * b1 + edge_corner(e1, 1)];
  return (n1 == m2 && n2 == m1);
}

/* ---- transformNode -------------------------------------------------------
   Canonicalise a node that lies on a tree corner / edge / face into the frame
   of the owner tree (lowest block sharing it), then clamp hmax -> hmax-1.
   Semantics of reference src/TMROctForest.cpp:3847-4039.  Coordinates are in
   [0, hmax].  Optional outputs as in the reference. */
TMR_HD void transform_node(const ConnTables &t, i32 *block, i32 *x, i32 *y,
                           i32 *z, int edge_dir, int *edge_reversed,
                           int *src_face_id) {
  const i32 H = kHmax;
  const int lo_x = (*x == 0), lo_y = (*y == 0), lo_z = (*z == 0);
  const int on_x = lo_x || (*x == H);
  const int on_y = lo_y || (*y == H);
  const int on_z = lo_z || (*z == H);
  if (edge_reversed) *edge_reversed = 0;
  if (src_face_id) *src_face_id = 0;
  const int nb = on_x + on_y + on_z;
  if (nb == 0) return;

  const int b = *block;
  if (nb == 3) {
    const int corner = (lo_x ? 0 : 1) + (lo_y ? 0 : 2) + (lo_z ? 0 : 4);
    const int node = t.block_conn[8 * b + corner];
    if (b != t.node_block_owners[node]) {
      const int first = t.node_block_conn[t.node_block_ptr[node]];
      const int ob = first / 8, oc = first % 8;
      *block = ob;
      *x = H * (oc & 1);
      *y = H * ((oc >> 1) & 1);
      *z = H * (oc >> 2);
    }
  } else if (nb == 2) {
    int e;
    i32 u;
    if (on_y && on_z) {
      e = (lo_y ? 0 : 1) + (lo_z ? 0 : 2);
      u = *x;
    } else if (on_x && on_z) {
      e = (lo_x ? 4 : 5) + (lo_z ? 0 : 2);
      u = *y;
    } else {
      e = (lo_x ? 8 : 9) + (lo_y ? 0 : 2);
      u = *z;
    }
    const int edge = t.block_edge_conn[12 * b + e];
    if (b != t.edge_block_owners[edge]) {
      const int first = t.edge_block_conn[t.edge_block_ptr[edge]];
      const int ob = first / 12, oe = first % 12;
      const int rev = edge_is_reversed(t, b, e, ob, oe);
      if (edge_reversed) *edge_reversed = rev;
      const i32 uo = rev ? H - u : u;
      *block = ob;
      edge_place(oe, uo, H, x, y, z);
    }
  } else {
    const int f = on_x * (lo_x ? 0 : 1) + on_y * (lo_y ? 2 : 3) +
                  on_z * (lo_z ? 4 : 5);
    const int face = t.block_face_conn[6 * b + f];
    if (b != t.face_block_owners[face]) {
      const int id = t.block_face_ids[6 * b + f];
      if (src_face_id) *src_face_id = id;
      i32 a, c, u, v;
      face_pick(f, *x, *y, *z, &a, &c);
      face_to_owner(id, H, a, c, &u, &v);
      if (edge_reversed) {
        /* direction test: push the unit vector of edge_dir through the same
           orientation with mirror length 0; a negative component means the
           edge runs backwards on the owner face */
        i32 d[3] = {0, 0, 0};
        if (edge_dir >= 0 && edge_dir < 3) d[edge_dir] = 1;
        i32 da, dc, du, dv;
        face_pick(f, d[0], d[1], d[2], &da, &dc);
        face_to_owner(id, 0, da, dc, &du, &dv);
        *edge_reversed = (du < 0 || dv < 0) ? 1 : 0;
      }
      const int first = t.face_block_conn[t.face_block_ptr[face]];
      const int ob = first / 6, of = first % 6;
      *block = ob;
      face_place(of, H * (of & 1), u, v, x, y, z);
    }
  }
  if (*x == H) *x = H - 1;
  if (*y == H) *y = H - 1;
  if (*z == H) *z = H - 1;
}

/* child-id -> the three parent faces / edges that touch the parent corner the
   child sits at (reference src/TMROctForest.cpp:49-64).  Child id bit0=x,
   bit1=y, bit2=z. */
TMR_HD int child_face(int id, int k) {
  /* k=0: x-face on the child's x side; k=1: y-face; k=2: z-face */
  return 2 * k + ((id >> k) & 1);
}

TMR_HD int child_edge(int id, int k) {
  /* k=0: the x-parallel edge (sides given by the y,z bits), k=1: y-parallel
     (x,z bits), k=2: z-parallel (x,y bits) */
  const int bx = id & 1, by = (id >> 1) & 1, bz = id >> 2;
  if (k == 0) return by + 2 * bz;
  if (k == 1) return 4 + bx + 2 * bz;
  return 8 + bx + 2 * by;
}

/* the two element edges of hanging face k (k as in child_face) that lie on
   the parent's own edges (reference src/TMROctForest.cpp:44-48) */
TMR_HD int child_face_edge(int id, int k, int j) {
  /* face k is normal to axis k; its two parent-aligned edges run along the
     other two axes: returned in the reference's order (higher axis first) */
  const int a_hi = (k == 2) ? 1 : 2; /* z for x/y faces, y for z faces */
  const int a_lo = (k == 0) ? 1 : 0; /* y for x faces, x for y/z faces */
  return child_edge(id, j == 0 ? a_hi : a_lo);
}

/* the four edges bounding block face f (reference src/TMROctForest.cpp:38-42) */
TMR_HD int face_edge(int f, int k) {
  const int side = f & 1;
  if (f < 2) {
    /* x-face: z-edges 8+side, 10+side; y-edges 4+side, 6+side */
    return (k < 2 ? 8 : 4) + side + 2 * (k & 1);
  } else if (f < 4) {
    /* y-face: z-edges 8+2s, 9+2s; x-edges 0+s, 2+s */
    return (k < 2) ? (8 + 2 * side + (k & 1)) : (side + 2 * (k & 1));
  }
  /* z-face: y-edges 4+2s, 5+2s; x-edges 0+2s, 1+2s */
  return (k < 2) ? (4 + 2 * side + (k & 1)) : (2 * side + (k & 1));
}

/* 1-D Lagrange basis on `order` knots evaluated at u
   (reference src/TMRInterpolation.h:40-52; same operation order so the fp64
   results agree to the last bit) */
TMR_HD void lagrange_basis(int order, double u, const double *knots,
                           double *N) {
  for (int i = 0; i < order; i++) {
    double v = 1.0;
    for (int j = 0; j < order; j++) {
      if (i != j) {
        double d = 1.0 / (knots[i] - knots[j]);
        v *= (u - knots[j]) * d;
      }
    }
    N[i] = v;
  }
}

/* ---- Bernstein points (reference src/TMRInterpolation.h:164-183,309-560) --
   Bernstein basis of degree order-1 at u in [-1,1] by the usual recurrence
   B_j^n = (1-t) B_j^(n-1) + t B_(j-1)^(n-1), t = (1+u)/2. */
TMR_HD void bernstein_basis(int order, double u, double *N) {
  const double u1 = 0.5 * (1.0 - u);
  const double u2 = 0.5 * (u + 1.0);
  N[0] = 1.0;
  for (int j = 1; j < order; j++) {
    double s = 0.0;
    for (int k = 0; k < j; k++) {
      const double t = N[k];
      N[k] = s + u1 * t;
      s = u2 * t;
    }
    N[j] = s;
  }
}

TMR_HD double binomial_small(int n, int k) {
  double c = 1.0;
  for (int i = 0; i < k; i++) c = c * (double)(n - i) / (double)(i + 1);
  return c;
}

/* Control point u of the two half-length children of a degree-p Bezier edge
   (p = order-1; u = -p..0 indexes the left child, 0..p the right one) as a
   combination of the parent's control points: rows of the midpoint
   subdivision (de Casteljau) matrices, C(k,j)/2^k.  All values are dyadic, so
   they equal the reference's tabulated eval_bernstein_weights exactly. */
TMR_HD void bernstein_subdivision_weights(int order, int u, double *N) {
  const int p = order - 1;
  for (int j = 0; j < order; j++) N[j] = 0.0;
  if (u <= 0) {
    const int k = u + p;
    const double scale = 1.0 / (double)(1 << k);
    for (int j = 0; j <= k; j++) N[j] = binomial_small(k, j) * scale;
  } else {
    const int k = u;
    const double scale = 1.0 / (double)(1 << (p - k));
    for (int j = k; j <= p; j++) N[j] = binomial_small(p - k, j - k) * scale;
  }
}

/* Control point i of a degree-(p+1) curve expressed in the degree-p control
   points of the same curve (degree elevation), p = coarse_order-1: weights
   i/(p+1) on point i-1 and (p+1-i)/(p+1) on point i (reference
   eval_bernstein_interp_weights, src/TMRInterpolation.h:456-560, tabulated for
   coarse orders 2..5).  The reference's table for coarse order 4 has the rows
   of fine points 2 and 3 exchanged; a drop-in has to return what the
   reference returns, so the exchange is reproduced here. */
TMR_HD void bernstein_elevation_weights(int coarse_order, int i, double *N) {
  const int p = coarse_order - 1;
  for (int j = 0; j < coarse_order; j++) N[j] = 0.0;
  if (coarse_order == 4 && (i == 2 || i == 3)) i = 5 - i;
  if (i > 0) N[i - 1] = (double)i / (double)(p + 1);
  if (i <= p) N[i] = (double)(p + 1 - i) / (double)(p + 1);
}

TMR_HD int popc32(u32 v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}

TMR_HD int popc64(u64 v) {
#if defined(__CUDA_ARCH__)
  return __popcll(v);
#else
  return __builtin_popcountll(v);
#endif
}

TMR_HD int ctz64(u64 v) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}
TMR_HD int ctz32(u32 v) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}

/* first index i in [0,n) with a[i] >= key */
TMR_HD i64 lower_bound_u64(const u64 *a, i64 n, u64 key) {
  i64 lo = 0, hi = n;
  while (lo < hi) {
    const i64 mid = lo + ((hi - lo) >> 1);
    if (a[mid] < key) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo;
}

/* Radix-indexed search: table[p] = first index whose key >> shift >= p, so a
   probe is one table read (the table is sized to stay L2-resident) plus a
   binary search inside one short bucket instead of log2(n) steps over the
   whole array. */
struct KeyIndex {
  const u32 *table; /* [nprefix + 1] */
  u64 nprefix;
  int shift;
  TMR_HD i64 find(const u64 *a, u64 key) const {
    u64 p = key >> shift;
    if (p >= nprefix) return -1;
    i64 lo = table[p], hi = table[p + 1];
    while (lo < hi) {
      const i64 mid = lo + ((hi - lo) >> 1);
      if (a[mid] < key) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    return (lo < (i64)table[p + 1] && a[lo] == key) ? lo : -1;
  }
  /* index of the last key <= `key` (-1 if none): the element whose Morton
     range starts at or before a position */
  TMR_HD i64 pred(const u64 *a, u64 key) const {
    u64 p = key >> shift;
    if (p >= nprefix) p = nprefix - 1;
    i64 lo = table[p], hi = table[p + 1];
    while (lo < hi) {
      const i64 mid = lo + ((hi - lo) >> 1);
      if (a[mid] <= key) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    return lo - 1;
  }
};

/* One bit per (level, tree, cell): set where a leaf of exactly that level
   sits.  An exact-leaf probe (the 6 parent-level neighbour tests per element
   of computeDepFacesAndEdges, reference src/TMROctForest.cpp:3619-3702) then
   costs one L2-resident word read instead of a table read plus a binary
   search with 64-bit compares (measured: 377 warp instructions per probe).
   Levels 0..lmax are covered, lmax chosen so the map stays within a byte
   budget; probes at deeper levels fall back to the key search. */
struct LeafMap {
  const u32 *bits; /* NULL = no map */
  int lmax;
  i32 block0, nblk; /* trees [block0, block0 + nblk): this rank's slice */
  u64 word_off[kMaxLevel + 1];
  TMR_HD bool covers(int level) const { return bits && level <= lmax; }
  /* leaves of other trees are not on this rank */
  TMR_HD bool in_range(i32 block) const {
    return block >= block0 && block < block0 + nblk;
  }
  TMR_HD u64 index(i32 block, u64 cell, int level) const {
    return ((u64)(u32)(block - block0) << (3 * level)) + cell;
  }
  TMR_HD bool test(i32 block, u64 cell, int level) const {
    if (!in_range(block)) return false;
    const u64 idx = index(block, cell, level);
    return (bits[word_off[level] + (idx >> 5)] >> (idx & 31)) & 1u;
  }
};

TMR_HD i64 find_u64(const u64 *a, i64 n, u64 key) {
  const i64 i = lower_bound_u64(a, n, key);
  return (i < n && a[i] == key) ? i : -1;
}

/* splitmix64 finaliser + the record hash used for synthetic refinement flags
   and order-independent checksums */
TMR_HD u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

TMR_HD u64 record_hash(u64 seed, i32 block, i32 x, i32 y, i32 z, int level) {
  u64 k = splitmix64(seed ^ (u64)(u32)block);
  k = splitmix64(k ^ (u64)(u32)x);
  k = splitmix64(k ^ (u64)(u32)y);
  k = splitmix64(k ^ (u64)(u32)z);
  k = splitmix64(k ^ (u64)(uint16_t)level);
  return k;
}

}  // namespace tmrgpu

#endif
