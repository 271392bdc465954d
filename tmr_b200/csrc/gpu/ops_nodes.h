/*
  ops_nodes.h -- createNodes on the device, one rank or several.

  Replaces reference src/TMROctForest.cpp:4064-4268 and its helpers
  computeDepFacesAndEdges :3619-3702, createLocalNodes :4290-4640,
  createLocalConn :4660-4867, labelDependentNodes :3711-3832,
  createDependentConn :5157-5519.

    hanging info      per element the 3 face + 3 edge exact-leaf probes of its
                      parent's neighbours against a per-level leaf bitmap
                      (HangingFn); the probes that leave the tree run in a second
                      dense launch over the compacted tree-face elements
                      (HangingBoundaryFn); on several ranks a miss that belongs
                      to another rank becomes a query answered by its owner
    nodes + conn      order 2: (leaf, slot) naming, scans, no sort
                      (ops_nodes_slots.h, build_nodes_slots; dependents are
                      labelled and everything numbered in slot space, the
                      connectivity comes out as final node numbers).
                      General path (orders >= 3, label bits, unbalanced input):
                      order^3 canonical node keys per element (transformNode)
                      with payload (element, slot) -> radix sort -> run heads =
                      unique nodes, conn[payload] = run index; dep_label, scans,
                      numbering, remap
    dep winner        per dependent node the LAST (element, edge|face) in the
                      reference's loop order that writes its stencil (atomicMax;
                      fused into the slot resolve on the order-2 path)
    dep fill          one thread per dependent node: the parent edge / face
                      node numbers (from the siblings' connectivity rows, or by
                      position) and the tabulated fp64 weight rows
    host mirrors      node_mirror_*: conn / dep_conn by DMA, dep_ptr + dep_weights
                      as stencil codes rebuilt by host threads, node numbers as
                      ranges
*/
#ifndef TMRGPU_OPS_NODES_H
#define TMRGPU_OPS_NODES_H

#include <sched.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "ops_balance.h"
#include "ops_route.h"
#include "ops_nodes_slots.h"

namespace tmrgpu {

/* build table[p] = lower_bound(keys, p << shift) for p in [0, nprefix] */
struct BuildIndexFn {
  const u64 *keys;
  i64 n;
  int shift;
  u64 nprefix;
  u32 *table;
  TMR_HD void operator()(i64 p) const {
    table[p] = ((u64)p >= nprefix) ? (u32)n
                                   : (u32)lower_bound_u64(keys, n, (u64)p << shift);
  }
};

/* index over a sorted key array whose keys are < key_limit; at most 2^22
   buckets (16 MB of u32, L2-resident) */
inline KeyIndex build_key_index(Ctx &ctx, const u64 *keys, i64 n, u64 key_limit,
                                DBuf<u32> &store, int max_bits = 22) {
  int shift = 0;
  while ((key_limit >> shift) > (1ULL << max_bits)) shift++;
  KeyIndex ix;
  ix.shift = shift;
  ix.nprefix = (key_limit >> shift) + 1;
  store.alloc(ctx, (i64)ix.nprefix + 1);
  BuildIndexFn b = {keys, n, shift, ix.nprefix, store.get()};
  launch(ctx, (i64)ix.nprefix + 1, b, "build_key_index");
  ix.table = store.get();
  return ix;
}

/* leaf map over the levels that fit a 128 MB budget (see LeafMap) */
inline LeafMap plan_leaf_map(int D, i32 block0, int nblocks, u64 *total_words) {
  LeafMap map;
  map.bits = NULL;
  map.lmax = -1;
  map.block0 = block0;
  map.nblk = nblocks;
  const u64 budget = 1ULL << 25; /* 32-bit words */
  u64 words = 0;
  for (int l = 0; l <= kMaxLevel; l++) map.word_off[l] = 0;
  for (int l = 0; l < D; l++) {
    const u64 w = ((((u64)nblocks << (3 * l)) + 31) >> 5) + 1;
    if (words + w > budget) break;
    map.word_off[l] = words;
    words += w;
    map.lmax = l;
  }
  *total_words = words;
  return map;
}

struct ElemView {
  const u64 *keys;
  i64 n;
  KeyFmt fmt;
  ConnTables t;
  KeyIndex ix;
  LeafMap map;
  /* multi-rank: probes that land in another rank's range become queries
     (key, owner, element*8+bit) answered by that rank -- a pull-style ghost
     exchange: only elements on the partition surface generate traffic */
  OwnerMap om;
  int me;
  int multi;
  /* trees strictly between these two lie entirely in this rank's range: a
     probe inside such a tree can never be another rank's leaf */
  i32 own_tree_lo, own_tree_hi;
  u64 *fq_key;
  u32 *fq_dest;
  u64 *fq_code;
  unsigned long long *fq_count;
  i64 fq_cap;

  /* does this rank hold the leaf `key` (exact anchor and level)? */
  TMR_HD bool local_leaf(u64 key) const {
    const int level = (int)(key & 31);
    if (map.covers(level)) {
      const u64 pos = key >> 5;
      const u64 m = fmt.D > 0 ? (pos & ((1ULL << (3 * fmt.D)) - 1)) : 0ULL;
      return map.test((i32)(pos >> (3 * fmt.D)), m >> (3 * (fmt.D - level)), level);
    }
    return ix.find(keys, key) >= 0;
  }
  /* multi-rank: a leaf that would live on another rank becomes a query */
  TMR_HD bool ask_owner(u64 key, u64 code) const {
    const int o = om.owner(key >> 5);
    if (o == me) return false;
    const unsigned long long slot = fetch_add_u64(fq_count, 1ULL);
    if ((i64)slot < fq_cap) {
      fq_key[slot] = key;
      fq_dest[slot] = (u32)o;
      fq_code[slot] = code;
    }
    return true;
  }
  TMR_HD bool probe_key(u64 key, u64 code) const {
    if (multi && ask_owner(key, code)) return false;
    return local_leaf(key);
  }
  TMR_HD bool probe(i32 block, i32 x, i32 y, i32 z, int level, u64 code) const {
    return probe_key(fmt.encode(block, x, y, z, level), code);
  }
  TMR_HD bool leaf_exists(i32 block, i32 x, i32 y, i32 z, int level) const {
    return local_leaf(fmt.encode(block, x, y, z, level));
  }

  /* is there a level-`level` leaf in another tree that is the image of the
     out-of-tree octant (x,y,z) across face f?  (reference checkAdjacentFaces
     src/TMROctForest.cpp:3465-3521) */
  TMR_HD bool across_face(int f, i32 block, i32 x, i32 y, i32 z, int level,
                          u64 code) const {
    const i32 h = 1 << (kMaxLevel - level);
    const i32 M = kHmax - h;
    const int face = t.block_face_conn[6 * block + f];
    i32 a, b, u, v;
    face_pick(f, x, y, z, &a, &b);
    face_to_owner(t.block_face_ids[6 * block + f], M, a, b, &u, &v);
    for (int ip = t.face_block_ptr[face]; ip < t.face_block_ptr[face + 1];
         ip++) {
      const int adj = t.face_block_conn[ip] / 6;
      if (adj == block) continue;
      const int af = t.face_block_conn[ip] % 6;
      i32 a1, b1, X, Y, Z;
      owner_to_face(t.block_face_ids[6 * adj + af], M, u, v, &a1, &b1);
      face_place(af, M * (af & 1), a1, b1, &X, &Y, &Z);
      if (probe(adj, X, Y, Z, level, code)) return true;
    }
    return false;
  }
  /* same across tree edge e (reference checkAdjacentEdges :3532-3605) */
  TMR_HD bool across_edge(int e, i32 block, i32 x, i32 y, i32 z, int level,
                          u64 code) const {
    const i32 h = 1 << (kMaxLevel - level);
    const i32 M = kHmax - h;
    const i32 u = (e < 4) ? x : (e < 8 ? y : z);
    const int edge = t.block_edge_conn[12 * block + e];
    for (int ip = t.edge_block_ptr[edge]; ip < t.edge_block_ptr[edge + 1];
         ip++) {
      const int adj = t.edge_block_conn[ip] / 12;
      if (adj == block) continue;
      const int ae = t.edge_block_conn[ip] % 12;
      const i32 uu = edge_is_reversed(t, block, e, adj, ae) ? M - u : u;
      i32 X, Y, Z;
      edge_place(ae, uu, M, &X, &Y, &Z);
      if (probe(adj, X, Y, Z, level, code)) return true;
    }
    return false;
  }
};

TMR_HD int child_id_of(i32 x, i32 y, i32 z, int level) {
  const i32 h = 1 << (kMaxLevel - level);
  return ((x & h) ? 1 : 0) | ((y & h) ? 2 : 0) | ((z & h) ? 4 : 0);
}

/* offset of the same-level neighbour across block edge e, in units of h */
TMR_HD void edge_dir(int e, int *dx, int *dy, int *dz) {
  const int s = e & 3;
  const int a = (s & 1) ? 1 : -1, b = (s >> 1) ? 1 : -1;
  if (e < 4) {
    *dx = 0; *dy = a; *dz = b;
  } else if (e < 8) {
    *dx = a; *dy = 0; *dz = b;
  } else {
    *dx = a; *dy = b; *dz = 0;
  }
}

/* set the map bit of every leaf whose level the map covers */
struct LeafMapBuildFn {
  const u64 *keys;
  KeyFmt fmt;
  LeafMap map;
  u32 *bits;
  TMR_HD void operator()(i64 i) const {
    const u64 key = keys[i];
    const int level = (int)(key & 31);
    if (level > map.lmax) return;
    const u64 pos = key >> 5;
    const u64 m = fmt.D > 0 ? (pos & ((1ULL << (3 * fmt.D)) - 1)) : 0ULL;
    const i32 block = (i32)(pos >> (3 * fmt.D));
    const u64 idx = map.index(block, m >> (3 * (fmt.D - level)), level);
    TMR_ATOMIC_OR_I32(&bits[map.word_off[level] + (idx >> 5)], 1u << (idx & 31));
  }
};

/* 6-bit hanging info of every element
   (reference computeDepFacesAndEdges src/TMROctForest.cpp:3619-3702).

   The parent-level neighbours of the element's PARENT across the 3 faces and 3
   edges meeting at the element's corner are addressed in Morton space: the
   parent's cell index is the element's Morton code shifted down, a step of one
   cell along an axis is a dilated increment/decrement on that axis' bits
   (x-major code: axis k of the child id sits at bit 2-k of every triple).  No
   coordinates are decoded unless the step leaves the tree; only those probes
   take the inter-tree path with its orientation transforms. */
struct HangingFn {
  ElemView ev;
  int *info32; /* 32-bit accumulator per element (bits 0-5 info) */
  /* one rank: nothing is patched in afterwards, the 6 bits go straight into
     the forest's int16 info array (no accumulator, no split pass) */
  int16_t *info16;
  TMR_HD void store_info(i64 i, int bits) const {
    if (info16) {
      info16[i] = (int16_t)bits;
    } else {
      info32[i] = bits;
    }
  }
  /* elements with a probe that leaves their tree: the inter-tree path (face /
     edge tables, orientation transforms) runs in a second, dense launch over
     this list -- inline it made every warp with one such element execute it
     (ncu: 17 of 32 lanes active on average) */
  u32 *slow_list;
  unsigned long long *slow_count;

  /* exact-leaf probe of a level-pl cell of this tree at levels the map does
     not cover */
  TMR_HD bool cell_probe(i32 block, u64 cell, int pl, u64 code) const {
    const int D = ev.fmt.D;
    return ev.probe_key(((u64)(u32)block << (3 * D + 5)) |
                            (cell << (3 * (D - pl) + 5)) | (u64)pl,
                        code);
  }

  /* probes that leave the tree (boundary elements only) */
  TMR_HD bool outside_face(int id, int k, i32 block, i32 px, i32 py, i32 pz,
                                i32 hp, int pl, u64 code) const {
    const int f = child_face(id, k);
    const i32 d = (f & 1) ? hp : -hp;
    return ev.across_face(f, block, px + (k == 0 ? d : 0), py + (k == 1 ? d : 0),
                          pz + (k == 2 ? d : 0), pl, code);
  }
  TMR_HD bool outside_edge(int id, int k, i32 block, i32 px, i32 py, i32 pz,
                                i32 hp, int pl, u64 code) const {
    const int e = child_edge(id, k);
    int dx, dy, dz;
    edge_dir(e, &dx, &dy, &dz);
    const i32 nx = px + dx * hp, ny = py + dy * hp, nz = pz + dz * hp;
    const int ox = (nx < 0 || nx >= kHmax);
    const int oy = (ny < 0 || ny >= kHmax);
    const int oz = (nz < 0 || nz >= kHmax);
    if (ox + oy + oz >= 2) return ev.across_edge(e, block, nx, ny, nz, pl, code);
    const int f = ox * (nx < 0 ? 0 : 1) + oy * (ny < 0 ? 2 : 3) +
                  oz * (nz < 0 ? 4 : 5);
    return ev.across_face(f, block, nx, ny, nz, pl, code);
  }

  TMR_HD void operator()(i64 i) const {
    const u64 key = ev.keys[i];
    const int level = (int)(key & 31);
    if (level == 0) {
      store_info(i, 0);
      return;
    }
    const int D = ev.fmt.D;
    const u64 pos = key >> 5;
    const u64 m = pos & ((1ULL << (3 * D)) - 1);
    const i32 block = (i32)(pos >> (3 * D));
    const int sh = 3 * (D - level);
    const int md = (int)((m >> sh) & 7); /* x-major digit: 4x + 2y + z */
    const int id = ((md >> 2) & 1) | (md & 2) | ((md & 1) << 2);
    const int pl = level - 1;
    const u64 mp = m >> (sh + 3); /* parent's cell index at level pl */
    const u64 lmask = pl > 0 ? ((1ULL << (3 * pl)) - 1) : 0ULL;
    u64 am[3], nc[3];
    bool out[3];
    TMR_UNROLL
    for (int k = 0; k < 3; k++) {
      am[k] = (0x1249249249249249ULL << (2 - k)) & lmask;
      const u64 comp = mp & am[k];
      if ((id >> k) & 1) {
        out[k] = comp == am[k];
        nc[k] = ((comp | ~am[k]) + 1) & am[k];
      } else {
        out[k] = comp == 0;
        nc[k] = (comp - 1) & am[k];
      }
    }
    /* the 6 neighbour cells: faces k = 0..2, then edges parallel to axis k */
    u64 cell[6];
    bool inside[6];
    TMR_UNROLL
    for (int k = 0; k < 3; k++) {
      const int a = (k == 0) ? 1 : 0, b = (k == 2) ? 1 : 2; /* the other axes */
      cell[k] = (mp & ~am[k]) | nc[k];
      inside[k] = !out[k];
      cell[k + 3] = (mp & ~(am[a] | am[b])) | nc[a] | nc[b];
      inside[k + 3] = !out[a] && !out[b];
    }
    int bits = 0;
    if (ev.map.covers(pl)) {
      /* all map words are requested before any is tested: six independent
         L2 reads in flight per thread instead of a chain of probe-and-branch */
      u32 word[6];
      TMR_UNROLL
      for (int q = 0; q < 6; q++) {
        const u64 idx = ev.map.index(block, cell[q], pl);
        word[q] = inside[q] ? ev.map.bits[ev.map.word_off[pl] + (idx >> 5)] : 0u;
      }
      TMR_UNROLL
      for (int q = 0; q < 6; q++) {
        if (!inside[q]) continue;
        const u64 idx = ev.map.index(block, cell[q], pl);
        if ((word[q] >> (idx & 31)) & 1u) {
          bits |= 1 << q;
        } else if (ev.multi && !(block > ev.own_tree_lo && block < ev.own_tree_hi)) {
          /* only a miss can be another rank's leaf */
          ev.ask_owner(((u64)(u32)block << (3 * D + 5)) |
                           (cell[q] << (3 * (D - pl) + 5)) | (u64)pl,
                       ((u64)i << 3) | (u64)q);
        }
      }
    } else {
      TMR_UNROLL
      for (int q = 0; q < 6; q++) {
        if (inside[q] && cell_probe(block, cell[q], pl, ((u64)i << 3) | (u64)q)) {
          bits |= 1 << q;
        }
      }
    }
    store_info(i, bits);
    /* probes that leave the tree: second launch (HangingBoundaryFn) */
    if (out[0] || out[1] || out[2]) append_u32(slow_count, slow_list, ev.n, (u32)i);
  }
};

/* second pass over the elements whose parent touches a face of its tree: the
   probes that leave the tree (reference checkAdjacentFaces / Edges through
   the face and edge transforms, :3465-3605) */
struct HangingBoundaryFn {
  HangingFn h;
  TMR_HD void operator()(i64 q) const {
    const ElemView &ev = h.ev;
    const i64 i = (i64)h.slow_list[q];
    const u64 key = ev.keys[i];
    const int level = (int)(key & 31);
    const int D = ev.fmt.D;
    const u64 pos = key >> 5;
    const u64 m = pos & ((1ULL << (3 * D)) - 1);
    const i32 block = (i32)(pos >> (3 * D));
    const int sh = 3 * (D - level);
    const int md = (int)((m >> sh) & 7);
    const int id = ((md >> 2) & 1) | (md & 2) | ((md & 1) << 2);
    const int pl = level - 1;
    const u64 mp = m >> (sh + 3);
    const u64 lmask = pl > 0 ? ((1ULL << (3 * pl)) - 1) : 0ULL;
    bool out[3];
    TMR_UNROLL
    for (int k = 0; k < 3; k++) {
      const u64 am = (0x1249249249249249ULL << (2 - k)) & lmask;
      const u64 comp = mp & am;
      out[k] = ((id >> k) & 1) ? (comp == am) : (comp == 0);
    }
    u32 ux, uy, uz;
    unmorton3(mp, &ux, &uy, &uz);
    const int s = kMaxLevel - pl;
    const i32 px = (i32)(ux << s), py = (i32)(uy << s), pz = (i32)(uz << s), hp = 1 << s;
    int bits = 0;
    TMR_UNROLL
    for (int k = 0; k < 3; k++) {
      const int a = (k == 0) ? 1 : 0, b = (k == 2) ? 1 : 2; /* the other axes */
      if (out[k] &&
          h.outside_face(id, k, block, px, py, pz, hp, pl, ((u64)i << 3) | (u64)k)) {
        bits |= 1 << k;
      }
      if ((out[a] || out[b]) &&
          h.outside_edge(id, k, block, px, py, pz, hp, pl,
                         ((u64)i << 3) | (u64)(k + 3))) {
        bits |= 1 << (k + 3);
      }
    }
    if (bits) {
      if (h.info16) {
        h.info16[i] = (int16_t)(h.info16[i] | bits);
      } else {
        h.info32[i] |= bits;
      }
    }
  }
};

/* owner side of the pull exchange: does each requested leaf exist here? */
struct AnswerProbeFn {
  const u64 *req;
  const u64 *keys;
  KeyIndex ix;
  unsigned char *ans;
  TMR_HD void operator()(i64 i) const { ans[i] = ix.find(keys, req[i]) >= 0 ? 1 : 0; }
};

/* requester side: set the info bit (and remember it came from another rank)
   for every probe that was answered "exists" */
struct PatchProbeFn {
  const u64 *code;
  const unsigned char *ans;
  int *info32; /* bits 0-5 info, bits 8-13 foreign mask */
  /* the elements that got a foreign bit, each once: the few whose parent
     edge / face nodes must be created explicitly (ParentNodeGen) */
  u32 *foreign_list;
  unsigned long long *foreign_count;
  i64 cap;
  TMR_HD void operator()(i64 i) const {
    if (ans[i]) {
      const u64 c = code[i];
      const int bit = (int)(c & 7);
      const int old = fetch_or_i32(&info32[c >> 3], (1 << bit) | (1 << (bit + 8)));
      if (!(old & 0x3f00)) append_u32(foreign_count, foreign_list, cap, (u32)(c >> 3));
    }
  }
};

struct Info32SplitFn {
  const int *info32;
  int16_t *info;
  unsigned char *fmask; /* optional */
  TMR_HD void operator()(i64 i) const {
    info[i] = (int16_t)(info32[i] & 63);
    if (fmask) fmask[i] = (unsigned char)((info32[i] >> 8) & 63);
  }
};

/* decode the 6-bit info into face / edge masks
   (reference decode_index_from_info src/TMROctForest.cpp:247-276) */
TMR_HD void decode_info(int id, int info, int *face_mask, int *edge_mask) {
  int fm = 0, em = 0;
  for (int k = 0; k < 3; k++) {
    if (info & (1 << k)) {
      fm |= 1 << child_face(id, k);
      em |= 1 << child_face_edge(id, k, 0);
      em |= 1 << child_face_edge(id, k, 1);
    }
    if (info & (1 << (k + 3))) em |= 1 << child_edge(id, k);
  }
  *face_mask = fm;
  *edge_mask = em;
}

TMR_HD int dep_corner_mask2(int id, int inf) {
  int fm, em;
  decode_info(id, inf, &fm, &em);
  for (int f = 0; f < 6; f++) {
    if (fm & (1 << f)) {
      for (int k = 0; k < 4; k++) em |= 1 << face_edge(f, k);
    }
  }
  int out = 0;
  TMR_UNROLL
  for (int c = 0; c < 8; c++) {
    const int ii = c & 1, jj = (c >> 1) & 1, kk = c >> 2;
    const bool dep = (((em >> (jj + 2 * kk)) & 1) && ii != (id & 1)) ||
                     (((em >> (4 + ii + 2 * kk)) & 1) && jj != ((id >> 1) & 1)) ||
                     (((em >> (8 + ii + 2 * jj)) & 1) && kk != (id >> 2));
    if (dep) out |= 1 << c;
  }
  return out;
}

/* element-local node offset of position p along block edge e */
TMR_HD int edge_node_offset(int order, int e, int p) {
  const int s = e & 3, hi = order - 1;
  const int a = hi * (s & 1), b = hi * (s >> 1);
  if (e < 4) return p + a * order + b * order * order;
  if (e < 8) return a + p * order + b * order * order;
  return a + b * order + p * order * order;
}

/* element-local node offset of in-face position (p,q) on block face f */
TMR_HD int face_node_offset(int order, int f, int p, int q) {
  const int n = (order - 1) * (f & 1);
  if (f < 2) return n + p * order + q * order * order;
  if (f < 4) return p + n * order + q * order * order;
  return p + q * order + n * order * order;
}

/* ---- node candidates ---------------------------------------------------------
   createLocalNodes + transformNode (reference :4290-4372, :3847-4039).

   Every element needs the index of each of its order^3 nodes; the sort of the
   candidate keys is the most expensive step of createNodes, so as few
   candidates as possible are emitted:
     * order 2, a COMPLETE family (8 sibling leaves, consecutive in the array):
       its leader emits the 27 distinct nodes of the 2x2x2 block once instead
       of 8x8 = 64; each candidate is tagged "shared" and the scatter fans it
       out to every sibling that has it as a corner;
     * everything else: order^3 candidates per element.
   The payload of a candidate is (element << 4) | code with code = corner
   (0..7, or slot at order 3 via the wide encoding) and bit 3 = "shared". */

struct NodeEmit {
  const u64 *keys;
  i64 E;
  KeyFmt fmt;
  NodeFmt nfmt;
  ConnTables t;
  int order;    /* geometry order: 2, or 3 for every mesh order >= 3 */
  int families; /* use family emission (order 2 only) */
  /* multi-rank: order-preserving dense ids of the trees this rank's node keys
     can name (NULL = tree index as is); shortens the sort key so that the
     conn-slot payload still fits beside it */
  const u32 *dense;
  /* check all 8 keys of a family, not only its first and last: needed when the
     forest is not known to be a complete tree (after a bare refine a sibling
     can be replaced by a deeper representative with the same anchor); the slot
     construction has verified completeness where it succeeded */
  int strict;
  TMR_HD u64 tree_id(i32 block) const {
    return dense ? (u64)dense[block] : (u64)(u32)block;
  }

  TMR_HD int digit_of(u64 k) const {
    const int L = (int)(k & 31);
    return L == 0 ? 0 : (int)((k >> (5 + 3 * (fmt.D - L))) & 7);
  }
  /* is e the first of 8 consecutive sibling leaves? */
  TMR_HD bool leader(i64 e) const {
    if (!families || e < 0 || e + 7 >= E) return false;
    const u64 k = keys[e];
    const int L = (int)(k & 31);
    if (L == 0 || digit_of(k) != 0) return false;
    const int s = 5 + 3 * (fmt.D - L);
    if (keys[e + 7] != k + (7ULL << s)) return false;
    if (strict) {
      for (int j = 1; j < 7; j++) {
        if (keys[e + j] != k + ((u64)j << s)) return false;
      }
    }
    return true;
  }
  /* is element e a member of a complete family? (m = its child digit) */
  TMR_HD bool in_family(i64 e, int *m) const {
    *m = digit_of(keys[e]);
    return families && leader(e - *m);
  }
  /* A member (a,b,c) of a complete family emits its node (i,j,k) unless some
     axis has the LAST index (order-1) with the member on the low side: that
     point is index 0 of the next sibling.  Per axis the two siblings emit
     order-1 and order positions, so the family emits (2 order - 1)^3 distinct
     nodes (27 at order 2, 125 at order 3) instead of 8 order^3, spread over
     its 8 threads. */
  TMR_HD u32 count(i64 e) const {
    const int n = order;
    if (!families) return (u32)(n * n * n);
    int m;
    if (!in_family(e, &m)) return (u32)(n * n * n);
    return (u32)((n - 1 + ((m >> 2) & 1)) * (n - 1 + ((m >> 1) & 1)) *
                 (n - 1 + (m & 1)));
  }

  /* payload encoding: (element, slot); family mode keeps the slot in the low
     3 (order 2) or 5 (order 3) bits */
  TMR_HD int slot_bits() const { return order == 2 ? 3 : 5; }
  TMR_HD u64 payload(i64 e, int code) const {
    return families ? (((u64)e << slot_bits()) | (u64)code)
                    : ((u64)e * (u64)(order * order * order) + (u64)code);
  }

  /* spread node coordinates of the element along one axis (axis bit a of each
     Morton triple: 2 = x, 1 = y, 0 = z), already shifted into node-key
     position; false if the element touches the tree boundary on this axis */
  template <int kOrder>
  TMR_HD bool axis_nodes(u64 mc, int a, int level, u64 *out) const {
    const int order = kOrder ? kOrder : this->order;
    const int D = fmt.D;
    const int sh = 3 * (D - level);
    const int up = 3 * (nfmt.Dn - D);
    const u64 am = 0x1249249249249249ULL << a;
    const u64 lvl = (level > 0 ? ((1ULL << (3 * level)) - 1) : 0ULL) & am;
    const u64 c = mc & am;
    /* one node step = 2^(Dn-level)/(order-1) units of depth Dn */
    const int stepbit = sh + up - (order == 3 ? 3 : 0) + a;
    u64 v = c << up;
    out[0] = v << 3;
    TMR_UNROLL
    for (int i = 1; i < order; i++) {
      v = ((v | ~am) + (1ULL << stepbit)) & am;
      out[i] = v << 3;
    }
    return c != 0 && ((mc >> sh) & lvl) != lvl;
  }

  /* kOrder = 2 or 3 (the geometry: 2x2x2 corners or 3x3x3 entity positions)
     unrolls the node loops */
  template <int kOrder, class Emit>
  TMR_HD void run_order(i64 e, Emit &emit) const {
    const int order = kOrder ? kOrder : this->order;
    int m = 0;
    const bool fam = families && in_family(e, &m);
    /* sibling bits in x-major order: m = 4*xbit + 2*ybit + zbit */
    const int bx = (m >> 2) & 1, by = (m >> 1) & 1, bz = m & 1;
    /* Interior elements (no node on a tree boundary) never leave Morton
       space: the per-axis components of the element's code ARE the spread
       node coordinates, a step of h/(order-1) along an axis is a dilated
       increment at bit 3(D-level) of that axis (the same bit for order 2 at
       depth D and order 3 at depth D+1), and the node format's trailing flag
       bit is one more shift by 3.  Only boundary elements decode coordinates
       and go through transform_node. */
    const u64 key0 = keys[e];
    const int level = (int)(key0 & 31);
    const int D = fmt.D;
    const u64 mc = D > 0 ? ((key0 >> 5) & ((1ULL << (3 * D)) - 1)) : 0ULL;
    i32 block = (i32)(key0 >> (3 * D + 5));
    const int np = order;
    bool interior = level > 0;
    u64 sx[kMaxOrder], sy[kMaxOrder], sz[kMaxOrder];
    interior = axis_nodes<kOrder>(mc, 2, level, sx) && interior;
    interior = axis_nodes<kOrder>(mc, 1, level, sy) && interior;
    interior = axis_nodes<kOrder>(mc, 0, level, sz) && interior;
    i32 x = 0, y = 0, z = 0, step = 0;
    if (!interior) {
      int lv;
      fmt.decode(key0, &block, &x, &y, &z, &lv);
      step = (1 << (kMaxLevel - level)) / (order - 1);
    }
    const u64 hi = tree_id(block) << (3 * (nfmt.Dn + 1));
    const int lb = nfmt.lbits;
    TMR_UNROLL
    for (int kk = 0; kk < np; kk++) {
      TMR_UNROLL
      for (int jj = 0; jj < np; jj++) {
        TMR_UNROLL
        for (int ii = 0; ii < np; ii++) {
          if (fam && ((ii == np - 1 && !bx) || (jj == np - 1 && !by) ||
                      (kk == np - 1 && !bz))) {
            continue;
          }
          const int label = lb ? slot_label(np, ii, jj, kk) : 0;
          u64 key;
          if (interior) {
            key = ((hi | sx[ii] | sy[jj] | sz[kk]) << lb) | (u64)label;
          } else {
            i32 b = block, nx = x + ii * step, ny = y + jj * step,
                nz = z + kk * step;
            transform_node(t, &b, &nx, &ny, &nz, -1, NULL, NULL);
            key = nfmt.encode((i32)tree_id(b), nx, ny, nz, label);
          }
          /* whether the slot stands for the whole family is re-derived from
             the keys at scatter time (in_family), not carried in the payload */
          emit(key, families ? (((u64)e << (kOrder == 2 ? 3 : 5)) |
                                (u64)(ii + np * jj + np * np * kk))
                             : payload(e, ii + np * jj + np * np * kk));
        }
      }
    }
  }
};

struct NodeEmitCountFn {
  NodeEmit g;
  TMR_HD u32 operator()(i64 e) const { return g.count(e); }
};

struct CandStore {
  u64 *k;
  u32 *v;     /* NULL in packed mode */
  int pshift; /* packed mode: payload << pshift */
  TMR_HD void operator()(u64 key, u64 payload) {
    if (v) {
      *k++ = key;
      *v++ = (u32)payload;
    } else {
      *k++ = key | (payload << pshift);
    }
  }
};

template <int kOrder>
struct NodeEmitFillFn {
  NodeEmit g;
  u64 *out_keys;
  u32 *out_vals;
  int pshift;
  TMR_HD void operator()(i64 e, u32 o) const {
    CandStore s = {out_keys + o, out_vals ? out_vals + o : (u32 *)0, pshift};
    g.template run_order<kOrder>(e, s);
  }
};

/* fixed order^3 candidates per element: offsets are e*npe, no scan needed */
template <int kOrder>
struct NodeEmitDenseFn {
  NodeEmit g;
  u64 *out_keys;
  u32 *out_vals;
  int pshift;
  TMR_HD void operator()(i64 e) const {
    const i64 o = e * (i64)(kOrder * kOrder * kOrder);
    CandStore s = {out_keys + o, out_vals ? out_vals + o : (u32 *)0, pshift};
    g.template run_order<kOrder>(e, s);
  }
};

struct RunHeadFn {
  const u64 *keys;
  TMR_HD u32 operator()(i64 i) const {
    return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
  }
};

/* run heads of keys whose high bits carry a payload */
struct RunHeadMaskedFn {
  const u64 *keys;
  u64 mask;
  TMR_HD u32 operator()(i64 i) const {
    return (i == 0 || (keys[i] & mask) != (keys[i - 1] & mask)) ? 1u : 0u;
  }
};

/* sorted candidates -> unique node keys + local connectivity */
static const u32 kNoSlot = 0xffffffffu; /* candidate that fills no conn slot */

template <int kNp> /* geometry order 2 or 3: folds the slot arithmetic */
struct NodeScatterFn {
  const u64 *keys;
  const u32 *vals; /* NULL: payload = keys[i] >> pshift */
  u64 mask;        /* node-key bits */
  int pshift;
  u64 no_slot;     /* payload value meaning "fills no conn slot" */
  u64 *node_keys;
  int *conn_local;
  unsigned char *created; /* optional: node is created by a local element */
  NodeEmit g;             /* payload decoding */
  const u32 *undense;     /* dense tree id -> tree index (NULL = identity) */
  int mbits;              /* Morton bits of a node key */
  /* heads_before = exclusive scan of run heads */
  TMR_HD void operator()(i64 i, u32 heads_before) const {
    const u64 k = keys[i] & mask;
    const bool head = (i == 0 || k != (keys[i - 1] & mask));
    const u32 run = heads_before + (head ? 1u : 0u) - 1u;
    if (head) {
      node_keys[run] = undense ? (((u64)undense[k >> mbits] << mbits) |
                                  (k & ((1ULL << mbits) - 1)))
                               : k;
    }
    const u64 p = vals ? (u64)vals[i] : (keys[i] >> pshift);
    if (p == no_slot) return;
    if (created) created[run] = 1;
    if (!g.families) {
      conn_local[p] = (int)run;
      return;
    }
    const int np = kNp, npe = np * np * np, sb = (kNp == 2) ? 3 : 5;
    const i64 e = (i64)(p >> sb);
    const int slot = (int)(p & ((1u << sb) - 1u));
    int m;
    if (!g.in_family(e, &m)) {
      conn_local[e * npe + slot] = (int)run;
      return;
    }
    /* node of a complete family: fan out to every sibling that holds it.  e is
       the emitting sibling, its child digit gives the family's first element
       and the node's position 0..2(np-1) on the family's grid; a sibling on
       side a of an axis holds grid position q at index q - a (np-1).
       (Measured alternative: one store into a per-family node table plus an
       expansion kernel -- scatter 5.8 -> 4.5 ms but the expansion cost 5.8 ms
       with one thread per family; an out-of-place expansion moves as many
       bytes as it saves.) */
    const i64 e0 = e - m;
    /* np is 2 or 3: constant divisors instead of run-time divisions */
    const int si = (np == 2) ? (slot & 1) : (slot % 3);
    const int sj = (np == 2) ? ((slot >> 1) & 1) : ((slot / 3) % 3);
    const int sk = (np == 2) ? (slot >> 2) : (slot / 9);
    const int qi = (np - 1) * ((m >> 2) & 1) + si;
    const int qj = (np - 1) * ((m >> 1) & 1) + sj;
    const int qk = (np - 1) * (m & 1) + sk;
    for (int a = 0; a < 2; a++) {
      const int ia = qi - a * (np - 1);
      if (ia < 0 || ia > np - 1) continue;
      for (int b = 0; b < 2; b++) {
        const int jb = qj - b * (np - 1);
        if (jb < 0 || jb > np - 1) continue;
        for (int c = 0; c < 2; c++) {
          const int kc = qk - c * (np - 1);
          if (kc < 0 || kc > np - 1) continue;
          conn_local[(e0 + 4 * a + 2 * b + c) * npe + ia + np * jb + np * np * kc] =
              (int)run;
        }
      }
    }
  }
};

/* nodes of the PARENT's hanging edges / faces (reference createLocalNodes
   :4378-4536): needed so that a dependent node's stencil can be numbered even
   when the coarse neighbour that owns those nodes lives on another rank */
struct ParentNodeGen {
  const u64 *keys;
  const unsigned char *fmask; /* info bits whose coarse neighbour is remote */
  KeyFmt fmt;
  NodeFmt nfmt;
  ConnTables t;
  int order;
  const u32 *dense; /* as NodeEmit::dense */
  TMR_HD i32 tree_id(i32 block) const {
    return dense ? (i32)dense[block] : block;
  }
  template <class Emit>
  TMR_HD void run(i64 e, Emit &emit) const {
    /* a hanging edge/face whose coarse neighbour is a LOCAL leaf needs nothing:
       that leaf creates the same nodes as its own corners */
    const int inf = fmask[e];
    if (!inf) return;
    i32 block, x, y, z;
    int level;
    fmt.decode(keys[e], &block, &x, &y, &z, &level);
    const int id = child_id_of(x, y, z, level);
    int fm, em;
    decode_info(id, inf, &fm, &em);
    const i32 h = 1 << (kMaxLevel - level);
    const i32 hp = 2 * h;
    const i32 px = x & ~h, py = y & ~h, pz = z & ~h;
    const i32 step = hp / (order - 1);
    for (int ed = 0; ed < 12; ed++) {
      if (!(em & (1 << ed))) continue;
      const int s = ed & 3;
      const i32 ta = hp * (s & 1), tb = hp * (s >> 1);
      for (int ii = 0; ii < order; ii++) {
        i32 b = block, nx, ny, nz;
        if (ed < 4) {
          nx = px + ii * step; ny = py + ta; nz = pz + tb;
        } else if (ed < 8) {
          nx = px + ta; ny = py + ii * step; nz = pz + tb;
        } else {
          nx = px + ta; ny = py + tb; nz = pz + ii * step;
        }
        transform_node(t, &b, &nx, &ny, &nz, -1, NULL, NULL);
        /* parent edge: end points are corners, the rest edge nodes */
        const int label =
            nfmt.lbits ? ((ii == 0 || ii == order - 1) ? 0 : 1) : 0;
        emit(b, nx, ny, nz, label);
      }
    }
    for (int f = 0; f < 6; f++) {
      if (!(fm & (1 << f))) continue;
      const i32 nn = hp * (f & 1);
      for (int q = 0; q < order; q++) {
        for (int p = 0; p < order; p++) {
          i32 b = block, nx, ny, nz;
          if (f < 2) {
            nx = px + nn; ny = py + p * step; nz = pz + q * step;
          } else if (f < 4) {
            nx = px + p * step; ny = py + nn; nz = pz + q * step;
          } else {
            nx = px + p * step; ny = py + q * step; nz = pz + nn;
          }
          transform_node(t, &b, &nx, &ny, &nz, -1, NULL, NULL);
          const int pe = (p == 0 || p == order - 1) ? 1 : 0;
          const int qe = (q == 0 || q == order - 1) ? 1 : 0;
          const int label = nfmt.lbits ? (2 - pe - qe) : 0;
          emit(b, nx, ny, nz, label);
        }
      }
    }
  }
};

/* trees whose index can appear in this rank's node keys: the trees of its own
   elements and the owner trees of their corners, edges and faces (the targets
   of transform_node) */
struct MarkTreesFn {
  const u64 *keys;
  i64 E;
  KeyFmt fmt;
  ConnTables t;
  u32 *used;
  TMR_HD void operator()(i64 b) const {
    const i64 first = (i64)(keys[0] >> (3 * fmt.D + 5));
    const i64 last = (i64)(keys[E - 1] >> (3 * fmt.D + 5));
    if (b < first || b > last) return;
    used[b] = 1;
    for (int c = 0; c < 8; c++) used[t.node_block_owners[t.block_conn[8 * b + c]]] = 1;
    for (int e = 0; e < 12; e++) {
      used[t.edge_block_owners[t.block_edge_conn[12 * b + e]]] = 1;
    }
    for (int f = 0; f < 6; f++) {
      used[t.face_block_owners[t.block_face_conn[6 * b + f]]] = 1;
    }
  }
};
struct UsedFlagFn {
  const u32 *used;
  TMR_HD u32 operator()(i64 b) const { return used[b]; }
};
struct UndenseFn {
  const u32 *used;
  const u32 *dense;
  u32 *undense;
  TMR_HD void operator()(i64 b) const {
    if (used[b]) undense[dense[b]] = (u32)b;
  }
};

struct CountKeyEmit {
  u32 n;
  TMR_HD void operator()(i32, i32, i32, i32, int) { n++; }
};
struct StoreKeyEmit {
  const ParentNodeGen *g;
  u64 *k;
  u32 *v; /* NULL in packed mode */
  u64 packed_no_slot; /* no_slot << pshift */
  TMR_HD void operator()(i32 b, i32 x, i32 y, i32 z, int label) {
    const u64 key = g->nfmt.encode(g->tree_id(b), x, y, z, label);
    if (v) {
      *k++ = key;
      *v++ = kNoSlot;
    } else {
      *k++ = key | packed_no_slot;
    }
  }
};
struct ParentNodeCountFn {
  ParentNodeGen g;
  TMR_HD u32 operator()(i64 e) const {
    CountKeyEmit c = {0};
    g.run(e, c);
    return c.n;
  }
};
struct ParentNodeFillFn {
  ParentNodeGen g;
  u64 *out_keys;
  u32 *out_vals;
  u64 packed_no_slot;
  TMR_HD void operator()(i64 e, u32 o) const {
    StoreKeyEmit s = {&g, out_keys + o, out_vals ? out_vals + o : (u32 *)0,
                      packed_no_slot};
    g.run(e, s);
  }
};



/* Edge / face membership of element-local node slot (i,j,k), order n per
   axis: which of the 12 element edges and 6 faces contain it, and where.
   Edges 0-3 run along x (index = side_y + 2 side_z), 4-7 along y, 8-11 along
   z; faces 0..5 = -x,+x,-y,+y,-z,+z with in-face axes as face_node_offset. */
struct SlotGeom {
  int ed[3], epos[3]; /* edge along axis a (or -1) and the position on it */
  int fc[3], fp[3], fq[3]; /* face normal to axis a (or -1) and (p,q) on it */
};
TMR_HD SlotGeom slot_geom(int n, int i, int j, int k) {
  SlotGeom g;
  const int hi = n - 1;
  const bool xi = (i == 0 || i == hi), yj = (j == 0 || j == hi),
             zk = (k == 0 || k == hi);
  const int sx = i ? 1 : 0, sy = j ? 1 : 0, sz = k ? 1 : 0;
  g.ed[0] = (yj && zk) ? (sy + 2 * sz) : -1;
  g.epos[0] = i;
  g.ed[1] = (xi && zk) ? (4 + sx + 2 * sz) : -1;
  g.epos[1] = j;
  g.ed[2] = (xi && yj) ? (8 + sx + 2 * sy) : -1;
  g.epos[2] = k;
  g.fc[0] = xi ? sx : -1;
  g.fp[0] = j;
  g.fq[0] = k;
  g.fc[1] = yj ? (2 + sy) : -1;
  g.fp[1] = i;
  g.fq[1] = k;
  g.fc[2] = zk ? (4 + sz) : -1;
  g.fp[2] = i;
  g.fq[2] = j;
  return g;
}

/* labelDependentNodes (reference :3711-3832), one pass over the element's
   node slots.  kOrder = 2 or 3 unrolls the slot loop; 0 = run-time order. */
template <int kOrder>
struct DepLabelFn {
  const u64 *keys;
  const int16_t *info;
  KeyFmt fmt;
  int order;
  int bernstein;
  const int *conn_local;
  unsigned char *dep_flag;
  TMR_HD void operator()(i64 e) const {
    const int inf = info[e];
    if (!inf) return;
    const int n = kOrder ? kOrder : order;
    const u64 key = keys[e];
    const int level = (int)(key & 31);
    const int md = level == 0 ? 0 : (int)((key >> (5 + 3 * (fmt.D - level))) & 7);
    const int id = ((md >> 2) & 1) | (md & 2) | ((md & 1) << 2);
    int fm, em;
    decode_info(id, inf, &fm, &em);
    for (int f = 0; f < 6; f++) {
      if (fm & (1 << f)) {
        for (int k = 0; k < 4; k++) em |= 1 << face_edge(f, k);
      }
    }
    /* dependent positions along an edge, by the child's side on that axis */
    const bool trim = (n == 3 && !bernstein);
    const int lo0 = 1, hi0 = trim ? n - 1 : n;     /* child on the low side */
    const int lo1 = trim ? 1 : 0, hi1 = n - 1;     /* child on the high side */
    const int *c = conn_local + e * (n * n * n);
    TMR_UNROLL
    for (int kk = 0; kk < n; kk++) {
      TMR_UNROLL
      for (int jj = 0; jj < n; jj++) {
        TMR_UNROLL
        for (int ii = 0; ii < n; ii++) {
          const SlotGeom g = slot_geom(n, ii, jj, kk);
          bool dep = false;
          TMR_UNROLL
          for (int a = 0; a < 3; a++) {
            if (g.ed[a] >= 0 && (em & (1 << g.ed[a]))) {
              const int bit = (id >> a) & 1;
              const int p = g.epos[a];
              dep = dep || (bit ? (p >= lo1 && p < hi1) : (p >= lo0 && p < hi0));
            }
            if (g.fc[a] >= 0 && (fm & (1 << g.fc[a]))) {
              dep = dep || (g.fp[a] >= 1 && g.fp[a] < n - 1 && g.fq[a] >= 1 &&
                            g.fq[a] < n - 1);
            }
          }
          if (dep) dep_flag[c[ii + n * jj + n * n * kk]] = 1;
        }
      }
    }
  }
};

struct DepFlagFn {
  const unsigned char *dep_flag;
  TMR_HD u32 operator()(i64 i) const { return dep_flag[i] ? 1u : 0u; }
};

/* node numbers in node order: dependents -1,-2,..; independents
   first_owned, first_owned+1, ..  (reference :4113-4181, single rank) */
struct NumberNodesFn {
  const unsigned char *dep_flag;
  const u32 *dep_before;
  int first_owned;
  int *node_num;
  int *dep_node; /* dependent index -> node index */
  TMR_HD void operator()(i64 i) const {
    const u32 d = dep_before[i];
    if (dep_flag[i]) {
      node_num[i] = -(int)d - 1;
      dep_node[d] = (int)i;
    } else {
      node_num[i] = first_owned + (int)(i - (i64)d);
    }
  }
};

/* createDependentConn passes 1+2 (reference :5183-5270): who writes the
   stencil of each dependent node, and with which length.  Codes are
   "1 + position in the reference's loop order" so that 0 means never; the
   reference's last writer is the maximum code.  One pass over the element's
   node slots: among the hanging edges (faces) containing a slot only the
   highest-numbered one can be this element's maximum, so each slot issues at
   most one edge and one face atomic.  Face positions on the two element edges
   that lie on the parent's own edges are skipped: decode_info put both edges
   into em, so a dependent node there has an edge winner from this element,
   and an edge winner takes precedence over any face winner (DepLenFn). */
template <int kOrder>
struct DepWinnerFn {
  const u64 *keys;
  const int16_t *info;
  KeyFmt fmt;
  int order;
  const int *conn; /* node NUMBERS (already remapped): dependent node d is
                      -d-1, so no gather is needed to classify a slot */
  u64 *win_edge;
  u64 *win_face;
  TMR_HD void operator()(i64 e) const {
    const int n = kOrder ? kOrder : order;
    run(e, conn + e * (n * n * n));
  }
  /* c: the element's node numbers (its conn row, possibly still in registers) */
  TMR_HD void run(i64 e, const int *c) const {
    const int inf = info[e];
    if (!inf) return;
    const int n = kOrder ? kOrder : order;
    const u64 key = keys[e];
    const int level = (int)(key & 31);
    const int md = level == 0 ? 0 : (int)((key >> (5 + 3 * (fmt.D - level))) & 7);
    const int id = ((md >> 2) & 1) | (md & 2) | ((md & 1) << 2);
    int fm, em;
    decode_info(id, inf, &fm, &em);
    const int bx = id & 1, by = (id >> 1) & 1, bz = id >> 2;
    TMR_UNROLL
    for (int kk = 0; kk < n; kk++) {
      TMR_UNROLL
      for (int jj = 0; jj < n; jj++) {
        TMR_UNROLL
        for (int ii = 0; ii < n; ii++) {
          const SlotGeom g = slot_geom(n, ii, jj, kk);
          int best_ed = -1, best_k = 0, best_f = -1, best_pos = 0;
          TMR_UNROLL
          for (int a = 0; a < 3; a++) { /* ascending edge / face numbers */
            if (g.ed[a] >= 0 && (em & (1 << g.ed[a]))) {
              best_ed = g.ed[a];
              best_k = g.epos[a];
            }
            if (g.fc[a] >= 0 && (fm & (1 << g.fc[a]))) {
              const int f = g.fc[a];
              const int p_skip = (n - 1) * ((f < 2) ? by : bx);
              const int q_skip = (n - 1) * ((f < 4) ? bz : by);
              if (g.fp[a] != p_skip && g.fq[a] != q_skip) {
                best_f = f;
                best_pos = g.fp[a] + g.fq[a] * n;
              }
            }
          }
          if (best_ed < 0 && best_f < 0) continue;
          const int num = c[ii + n * jj + n * n * kk];
          if (num >= 0) continue;
          if (best_ed >= 0) {
            const u64 code = (((u64)e * 12 + best_ed) << 4) + (u64)best_k + 1;
            TMR_ATOMIC_MAX_U64(&win_edge[-num - 1], code);
          }
          if (best_f >= 0) {
            const u64 code = (((u64)e * 6 + best_f) << 8) + (u64)best_pos + 1;
            TMR_ATOMIC_MAX_U64(&win_face[-num - 1], code);
          }
        }
      }
    }
  }
};

/* slot resolve with the winner pass of the dependent CSR fused in: the
   element's node numbers go from registers straight into DepWinnerFn, which
   saves a second 2.8 GB read of the connectivity */
struct SlotResolveWin2Fn {
  SlotResolve2Fn rs;
  DepWinnerFn<2> win;
  TMR_HD void operator()(i64 e) const {
    u32 leaf[8];
    rs.row(e, leaf);
    win.run(e, reinterpret_cast<const int *>(leaf));
  }
};
/* several ranks: elements with a corner outside the rank's range get their
   last numbers from the B pass and are revisited by PendingWinnerFn */
struct SlotResolveWin3Fn {
  SlotResolve3Fn rs;
  DepWinnerFn<2> win;
  unsigned char *pending;
  TMR_HD void operator()(i64 e) const {
    u32 leaf[8];
    if (rs.row(e, leaf)) {
      win.run(e, reinterpret_cast<const int *>(leaf));
    } else {
      pending[e] = 1;
    }
  }
};
struct PendingWinnerFn { /* after BPlaceFn: the elements with a B corner */
  const unsigned char *pending;
  DepWinnerFn<2> win;
  TMR_HD void operator()(i64 e) const {
    if (pending[e]) win(e);
  }
};

#if defined(__CUDACC__)
/* gather-latency-bound bodies: occupancy over registers */
template <>
struct LaunchMinBlocks<SlotResolveWin2Fn> {
  static const int value = 6;
};
template <>
struct LaunchMinBlocks<SlotResolveWin3Fn> {
  static const int value = 6;
};
#endif

struct DepLenFn {
  const u64 *win_edge;
  const u64 *win_face;
  int order;
  TMR_HD u32 operator()(i64 d) const {
    if (win_edge[d]) return (u32)order;
    if (win_face[d]) return (u32)(order * order);
    return 0;
  }
};

/* ---- orders >= 4: entities -> nodes ------------------------------------------
   The sorted unique keys are the 27-per-element ENTITIES (corner, edge, face,
   block, told apart by the label bits); an entity with label L owns
   (order-2)^L consecutive local nodes (reference createNodes :4090-4100). */
struct EntityMultFn {
  const u64 *ent_keys;
  int order;
  TMR_HD u32 operator()(i64 i) const {
    const int label = (int)(ent_keys[i] & 3);
    u32 m = 1;
    for (int k = 0; k < label; k++) m *= (u32)(order - 2);
    return m;
  }
};

/* local node id of the sub-node of an entity reached from a tree whose own
   frame may differ from the owner tree's: `b,x,y,z` = the entity's position in
   the tree it is seen from.  Edge nodes run backwards when the edge is
   reversed on the owner, face nodes go through the face orientation
   (reference createLocalConn :4716-4836, getEdgeNodes :4947-4964,
   getFaceNodes :5083-5141). */
struct EntityNodes {
  ConnTables t;
  int order;
  /* position p = 1..order-2 along the edge that runs in direction edge_dir */
  TMR_HD int edge_sub(i32 b, i32 x, i32 y, i32 z, int edge_dir, int p) const {
    int rev = 0;
    transform_node(t, &b, &x, &y, &z, edge_dir, &rev, NULL);
    return rev ? (order - 2 - p) : (p - 1);
  }
  TMR_HD int edge_reversed(i32 b, i32 x, i32 y, i32 z, int edge_dir) const {
    int rev = 0;
    transform_node(t, &b, &x, &y, &z, edge_dir, &rev, NULL);
    return rev;
  }
  TMR_HD int face_id(i32 b, i32 x, i32 y, i32 z) const {
    int id = 0;
    transform_node(t, &b, &x, &y, &z, -1, NULL, &id);
    return id;
  }
  /* in-face position (a, c), both 1..order-2, on a face with orientation id */
  TMR_HD int face_sub(int id, int a, int c) const {
    i32 u, v;
    face_to_owner(id, order - 1, a, c, &u, &v);
    return (u - 1) + (v - 1) * (order - 2);
  }
};

/* createLocalConn (reference :4660-4867) from the element's 27 entity indices */
struct ConnBuildFn {
  const u64 *keys;
  KeyFmt fmt;
  EntityNodes en;
  const int *ent_conn; /* [E][27] entity index per 3x3x3 position */
  const u32 *ent_off;  /* first local node of every entity */
  int *conn;           /* [E][order^3] local node ids */
  TMR_HD void operator()(i64 e) const {
    const int n = en.order;
    i32 block, x, y, z;
    int level;
    fmt.decode(keys[e], &block, &x, &y, &z, &level);
    const i32 h = 1 << (kMaxLevel - level - 1);
    const int *ec = ent_conn + e * 27;
    int *c = conn + e * (i64)(n * n * n);
    for (int k3 = 0; k3 < 3; k3++) {
      for (int j3 = 0; j3 < 3; j3++) {
        for (int i3 = 0; i3 < 3; i3++) {
          const int base = (int)ent_off[ec[i3 + 3 * j3 + 9 * k3]];
          const i32 ex = x + h * i3, ey = y + h * j3, ez = z + h * k3;
          const int mx = (i3 == 1), my = (j3 == 1), mz = (k3 == 1);
          /* slot index of an extreme (0 / order-1) coordinate */
          const int ci = (n - 1) * (i3 / 2), cj = (n - 1) * (j3 / 2),
                    ck = (n - 1) * (k3 / 2);
          if (mx + my + mz == 0) {
            c[ci + n * cj + n * n * ck] = base;
          } else if (mx + my + mz == 1) {
            const int dir = mx ? 0 : (my ? 1 : 2);
            const int rev = en.edge_reversed(block, ex, ey, ez, dir);
            for (int p = 1; p < n - 1; p++) {
              const int ii = mx ? p : ci, jj = my ? p : cj, kk = mz ? p : ck;
              c[ii + n * jj + n * n * kk] = base + (rev ? (n - 2 - p) : (p - 1));
            }
          } else if (mx + my + mz == 2) {
            const int id = en.face_id(block, ex, ey, ez);
            for (int q = 1; q < n - 1; q++) {
              for (int p = 1; p < n - 1; p++) {
                /* in-face axes in the reference's order: (y,z), (x,z), (x,y) */
                int ii, jj, kk;
                if (!mx) {
                  ii = ci; jj = p; kk = q;
                } else if (!my) {
                  ii = p; jj = cj; kk = q;
                } else {
                  ii = p; jj = q; kk = ck;
                }
                c[ii + n * jj + n * n * kk] = base + en.face_sub(id, p, q);
              }
            }
          } else {
            for (int kk = 1; kk < n - 1; kk++) {
              for (int jj = 1; jj < n - 1; jj++) {
                for (int ii = 1; ii < n - 1; ii++) {
                  c[ii + n * jj + n * n * kk] =
                      base + (ii - 1) + (jj - 1) * (n - 2) +
                      (kk - 1) * (n - 2) * (n - 2);
                }
              }
            }
          }
        }
      }
    }
  }
};

/* createDependentConn pass 3 (reference :5272-5508) */
/* rows of the weight table: [kind 0 edge / 1 face][side bit][position k] ->
   `order` weights, by exactly the expressions DepFillFn evaluates */
struct DepWeightTableFn {
  int order;
  int bernstein;
  double knots[kMaxOrder];
  double *table;
  TMR_HD void operator()(i64 item) const {
    const int k = (int)(item % order), b = (int)((item / order) & 1),
              kind = (int)(item / (2 * order));
    double *row = table + item * order;
    if (bernstein) {
      bernstein_subdivision_weights(order, (order - 1) * (b - 1) + k, row);
    } else if (kind == 0) {
      const double u = 1.0 * (b - 1) + 0.5 * (1.0 + knots[k]);
      lagrange_basis(order, u, knots, row);
    } else {
      double u = -1.0 + 0.5 * (1.0 + knots[k]);
      u += 1.0 * b;
      lagrange_basis(order, u, knots, row);
    }
  }
};

struct DepFillData {
  const u64 *keys;
  KeyFmt fmt;
  NodeFmt nfmt;
  ConnTables t;
  int order;
  double knots[kMaxOrder];
  const u64 *node_keys;
  i64 num_nodes;
  const int *node_num;
  const u64 *win_edge;
  const u64 *win_face;
  const int *dep_ptr;
  int *dep_conn;
  double *dep_weights;
  unsigned short *dep_code; /* compact form of ptr + weights (DepExpandJob) */
  const double *wtab;       /* 1-D weight rows (DepWeightTableFn) */
  SlotLookup sl; /* one rank, slot construction: node numbers by position */
  SlotLookupM slm; /* the same on several ranks */

  KeyIndex node_ix;
  /* order 2 shortcut: in a complete family the parent's corner c is corner c
     of sibling c, so the parent's edge/face nodes are read from the siblings'
     connectivity (node numbers, already remapped) instead of being searched
     by key */
  const int *conn;
  NodeEmit fam;

  int bernstein;
  /* orders >= 4: first local node of every entity (NULL below order 4, where
     entity index == node index) and the sub-node permutations */
  const u32 *ent_off;
  EntityNodes en;
};

/* kOrder = 2 or 3 compiles the label-free fast paths only (the order-2 kernel
   is one of the five largest of the cycle); 0 = run-time order incl. the
   entity look-ups of orders >= 4 */
template <int kOrder>
struct DepFillFn : DepFillData {

  /* entity at (block,x,y,z) with `label` -> first local node (or -1) */
  TMR_HD i64 entity_base(i32 block, i32 x, i32 y, i32 z, int label) const {
    transform_node(t, &block, &x, &y, &z, -1, NULL, NULL);
    const i64 idx = node_ix.find(node_keys, nfmt.encode(block, x, y, z, label));
    if (idx < 0) return -1;
    return ent_off ? (i64)ent_off[idx] : idx;
  }
  TMR_HD int num_at(i64 node) const { return node >= 0 ? node_num[node] : 0; }

  /* the `order` nodes of edge ed of the octant (block, px,py,pz) of size hp
     (reference getEdgeNodes :4886-4966) */
  TMR_HD void edge_nodes_general(i32 block, i32 px, i32 py, i32 pz, i32 hp,
                                 int ed, int *out) const {
    const int n = order;
    const i32 hh = hp / 2;
    const int s = ed & 3;
    const i32 ta = hp * (s & 1), tb = hp * (s >> 1);
    for (int g = 0; g < 3; g++) {
      i32 nx, ny, nz;
      if (ed < 4) {
        nx = px + g * hh; ny = py + ta; nz = pz + tb;
      } else if (ed < 8) {
        nx = px + ta; ny = py + g * hh; nz = pz + tb;
      } else {
        nx = px + ta; ny = py + tb; nz = pz + g * hh;
      }
      if (g != 1) {
        out[(g / 2) * (n - 1)] = num_at(entity_base(block, nx, ny, nz, 0));
      } else {
        const i64 base = entity_base(block, nx, ny, nz, 1);
        const int rev = en.edge_reversed(block, nx, ny, nz, ed >> 2);
        for (int k = 1; k < n - 1; k++) {
          out[k] = num_at(base < 0 ? -1 : base + (rev ? (n - 2 - k) : (k - 1)));
        }
      }
    }
  }
  /* the order^2 nodes of face f (reference getFaceNodes :4990-5146) */
  TMR_HD void face_nodes_general(i32 block, i32 px, i32 py, i32 pz, i32 hp,
                                 int f, int *out) const {
    const int n = order;
    const i32 hh = hp / 2;
    const i32 nn = hp * (f & 1);
    /* directions of the first / second in-face axis */
    const int ax1 = (f < 2) ? 1 : 0, ax2 = (f < 4) ? 2 : 1;
    for (int jj = 0; jj < 3; jj++) {
      for (int ii = 0; ii < 3; ii++) {
        i32 nx, ny, nz;
        if (f < 2) {
          nx = px + nn; ny = py + hh * ii; nz = pz + hh * jj;
        } else if (f < 4) {
          nx = px + hh * ii; ny = py + nn; nz = pz + hh * jj;
        } else {
          nx = px + hh * ii; ny = py + hh * jj; nz = pz + nn;
        }
        const int ie = (ii != 1), je = (jj != 1);
        if (ie && je) {
          out[(ii / 2) * (n - 1) + (jj / 2) * (n - 1) * n] =
              num_at(entity_base(block, nx, ny, nz, 0));
        } else if (ie || je) {
          /* an edge of the face: it runs along the axis whose index is 1 */
          const i64 base = entity_base(block, nx, ny, nz, 1);
          const int rev = en.edge_reversed(block, nx, ny, nz, ie ? ax2 : ax1);
          const int incr = ie ? n : 1;
          const int start = ie ? (ii / 2) * (n - 1) : (jj / 2) * (n - 1) * n;
          for (int k = 1; k < n - 1; k++) {
            out[start + k * incr] =
                num_at(base < 0 ? -1 : base + (rev ? (n - 2 - k) : (k - 1)));
          }
        } else {
          const i64 base = entity_base(block, nx, ny, nz, 2);
          const int id = en.face_id(block, nx, ny, nz);
          for (int k = 1; k < n - 1; k++) {
            for (int j = 1; j < n - 1; j++) {
              out[j + k * n] = num_at(base < 0 ? -1 : base + en.face_sub(id, j, k));
            }
          }
        }
      }
    }
  }

  TMR_HD int lookup(i32 block, i32 x, i32 y, i32 z, int label) const {
    transform_node(t, &block, &x, &y, &z, -1, NULL, NULL);
    if (sl.on) return sl.number(block, x, y, z);
    if (slm.on) return slm.number(block, x, y, z);
    const i64 idx = node_ix.find(node_keys, nfmt.encode(block, x, y, z, label));
    return idx >= 0 ? node_num[idx] : 0;
  }
  /* labels in use: end points of a parent edge are corner nodes, the rest
     edge nodes; on a face additionally the centre is a face node */
  TMR_HD int line_label(int i) const {
    return nfmt.lbits ? ((i == 0 || i == order - 1) ? 0 : 1) : 0;
  }
  /* first element of e's family if that family is complete, else -1 */
  TMR_HD i64 family_base(i64 e) const {
    if (kOrder != 2) return -1;
    int m;
    return fam.in_family(e, &m) ? e - m : -1;
  }
  TMR_HD int parent_corner_node(i64 e0, int c) const {
    const int m = 4 * (c & 1) + 2 * ((c >> 1) & 1) + (c >> 2);
    return conn[(e0 + m) * 8 + c];
  }

  /* Order 2: an edge stencil is the 2 end corners of the parent's edge, a face
     stencil the 4 corners of the parent's face.  Edge and face winners decode
     into ONE form -- (element, count, parent corners, weights) -- and share the
     gather loop, so the lanes of a warp do not take turns through two long
     branches (ncu before: 14 of 32 lanes active on average). */
  TMR_HD void fill_order2(i64 d) const {
    const u64 we = win_edge[d], wf = win_face[d];
    if (!we && !wf) {
      dep_code[d] = kDepCodeNone;
      return;
    }
    const bool is_edge = we != 0;
    const u64 code = (is_edge ? we : wf) - 1;
    const u64 ee = is_edge ? (code >> 4) : (code >> 8);
    const int sub = (int)(ee % (is_edge ? 12 : 6)); /* edge or face number */
    const i64 e = (i64)(ee / (is_edge ? 12 : 6));
    const int pos = (int)(code & (is_edge ? 15 : 255));
    const u64 key = keys[e];
    const int level = (int)(key & 31);
    const int md = level == 0 ? 0 : (int)((key >> (5 + 3 * (fmt.D - level))) & 7);
    const int id = ((md >> 2) & 1) | (md & 2) | ((md & 1) << 2);
    int cnt, c[4];
    double w[4];
    if (is_edge) {
      const int sa = sub & 1, sb = (sub >> 1) & 1, ax = sub >> 2;
      TMR_UNROLL
      for (int ii = 0; ii < 2; ii++) {
        c[ii] = ax == 0 ? (ii + 2 * sa + 4 * sb)
                        : (ax == 1 ? (sa + 2 * ii + 4 * sb) : (sa + 2 * sb + 4 * ii));
      }
      c[2] = c[3] = 0;
      const int bit = (id >> ax) & 1;
      const double *row = wtab + (size_t)(bit * 2 + pos) * 2;
      w[0] = row[0];
      w[1] = row[1];
      w[2] = w[3] = 0.0;
      cnt = 2;
      dep_code[d] = (unsigned short)(kDepCodeEdge | (bit << 4) | pos);
    } else {
      const int n1 = sub & 1, fa = sub >> 1;
      const int ii = pos & 1, jj = pos >> 1;
      const int bx = id & 1, by = (id >> 1) & 1, bz = id >> 2;
      const int b1 = (fa == 0) ? by : bx;
      const int b2 = (fa == 2) ? by : bz;
      const double *Nu = wtab + (size_t)((2 + b1) * 2 + ii) * 2;
      const double *Nv = wtab + (size_t)((2 + b2) * 2 + jj) * 2;
      TMR_UNROLL
      for (int q = 0; q < 2; q++) {
        TMR_UNROLL
        for (int p = 0; p < 2; p++) {
          c[p + 2 * q] = fa == 0 ? (n1 + 2 * p + 4 * q)
                                 : (fa == 1 ? (p + 2 * n1 + 4 * q) : (p + 2 * q + 4 * n1));
          w[p + 2 * q] = Nu[p] * Nv[q];
        }
      }
      cnt = 4;
      dep_code[d] = (unsigned short)((b1 << 9) | (ii << 5) | (b2 << 4) | jj);
    }
    const int ptr = dep_ptr[d];
    const i64 e0 = family_base(e);
    if (e0 >= 0) {
      int num[4];
      TMR_UNROLL
      for (int j = 0; j < 4; j++) num[j] = j < cnt ? parent_corner_node(e0, c[j]) : 0;
      TMR_UNROLL
      for (int j = 0; j < 4; j++) {
        if (j < cnt) {
          dep_conn[ptr + j] = num[j];
          dep_weights[ptr + j] = w[j];
        }
      }
      return;
    }
    /* incomplete family: the parent's corners by position */
    i32 block, x, y, z;
    int lv;
    fmt.decode(key, &block, &x, &y, &z, &lv);
    const i32 h = 1 << (kMaxLevel - lv), hp = 2 * h;
    const i32 px = x & ~h, py = y & ~h, pz = z & ~h;
    for (int j = 0; j < cnt; j++) {
      dep_conn[ptr + j] = lookup(block, px + hp * (c[j] & 1), py + hp * ((c[j] >> 1) & 1),
                                 pz + hp * (c[j] >> 2), 0);
      dep_weights[ptr + j] = w[j];
    }
  }

  TMR_HD void operator()(i64 d) const {
    if (kOrder == 2) {
      fill_order2(d);
      return;
    }
    const int order = kOrder ? kOrder : DepFillData::order;
    const bool general = (kOrder == 0) && ent_off != NULL;
    const int ptr = dep_ptr[d];
    if (!win_edge[d] && !win_face[d]) dep_code[d] = kDepCodeNone;
    if (win_edge[d]) {
      const u64 code = win_edge[d] - 1;
      const int k = (int)(code & 15);
      const u64 ee = code >> 4;
      const int ed = (int)(ee % 12);
      const i64 e = (i64)(ee / 12);
      i32 block, x, y, z;
      int level;
      fmt.decode(keys[e], &block, &x, &y, &z, &level);
      const int id = child_id_of(x, y, z, level);
      const i32 h = 1 << (kMaxLevel - level);
      const i32 hp = 2 * h;
      const i32 px = x & ~h, py = y & ~h, pz = z & ~h;
      const i32 step = hp / (order - 1);
      const int s = ed & 3;
      const i32 ta = hp * (s & 1), tb = hp * (s >> 1);
      const i64 e0 = family_base(e);
      if (general) edge_nodes_general(block, px, py, pz, hp, ed, dep_conn + ptr);
      for (int ii = 0; ii < order && !general; ii++) {
        if (e0 >= 0) {
          const int sa = s & 1, sb = s >> 1;
          const int c = ed < 4 ? (ii + 2 * sa + 4 * sb)
                               : (ed < 8 ? (sa + 2 * ii + 4 * sb)
                                         : (sa + 2 * sb + 4 * ii));
          dep_conn[ptr + ii] = parent_corner_node(e0, c);
          continue;
        }
        i32 nx, ny, nz;
        if (ed < 4) {
          nx = px + ii * step; ny = py + ta; nz = pz + tb;
        } else if (ed < 8) {
          nx = px + ta; ny = py + ii * step; nz = pz + tb;
        } else {
          nx = px + ta; ny = py + tb; nz = pz + ii * step;
        }
        dep_conn[ptr + ii] = lookup(block, nx, ny, nz, line_label(ii));
      }
      const int bit = (id >> (ed >> 2)) & 1;
      dep_code[d] = (unsigned short)(kDepCodeEdge | (bit << 4) | k);
      {
        /* the row was evaluated once by DepWeightTableFn (reference :5364-5374:
           Lagrange values at u = (bit - 1) + (1 + knot[k]) / 2, or the Bernstein
           subdivision row) */
        const double *row = wtab + (size_t)(bit * order + k) * order;
        for (int j = 0; j < order; j++) dep_weights[ptr + j] = row[j];
      }
    } else if (win_face[d]) {
      const u64 code = win_face[d] - 1;
      const int pos = (int)(code & 255);
      const u64 ff = code >> 8;
      const int f = (int)(ff % 6);
      const i64 e = (i64)(ff / 6);
      const int ii = pos % order, jj = pos / order;
      i32 block, x, y, z;
      int level;
      fmt.decode(keys[e], &block, &x, &y, &z, &level);
      const int id = child_id_of(x, y, z, level);
      const i32 h = 1 << (kMaxLevel - level);
      const i32 hp = 2 * h;
      const i32 px = x & ~h, py = y & ~h, pz = z & ~h;
      const i32 step = hp / (order - 1);
      const i32 nn = hp * (f & 1);
      const i64 e0 = family_base(e);
      if (general) face_nodes_general(block, px, py, pz, hp, f, dep_conn + ptr);
      for (int q = 0; q < order && !general; q++) {
        for (int p = 0; p < order; p++) {
          if (e0 >= 0) {
            const int n1 = f & 1;
            const int c = f < 2 ? (n1 + 2 * p + 4 * q)
                                : (f < 4 ? (p + 2 * n1 + 4 * q)
                                         : (p + 2 * q + 4 * n1));
            dep_conn[ptr + p + q * order] = parent_corner_node(e0, c);
            continue;
          }
          i32 nx, ny, nz;
          if (f < 2) {
            nx = px + nn; ny = py + p * step; nz = pz + q * step;
          } else if (f < 4) {
            nx = px + p * step; ny = py + nn; nz = pz + q * step;
          } else {
            nx = px + p * step; ny = py + q * step; nz = pz + nn;
          }
          dep_conn[ptr + p + q * order] =
              lookup(block, nx, ny, nz,
                     nfmt.lbits ? (line_label(p) + line_label(q)) : 0);
        }
      }
      /* child bits along the face's first / second in-face axis */
      const int bx = id & 1, by = (id >> 1) & 1, bz = id >> 2;
      const int b1 = (f < 2) ? by : bx;
      const int b2 = (f < 4) ? bz : by;
      dep_code[d] = (unsigned short)((b1 << 9) | (ii << 5) | (b2 << 4) | jj);
      /* tensor product of two tabulated rows (reference :5453-5473) */
      const double *Nu = wtab + (size_t)((2 + b1) * order + ii) * order;
      const double *Nv = wtab + (size_t)((2 + b2) * order + jj) * order;
      for (int q = 0; q < order; q++) {
        for (int p = 0; p < order; p++) {
          dep_weights[ptr + p + q * order] = Nu[p] * Nv[q];
        }
      }
    }
  }
};

#if defined(__CUDACC__)
template <>
struct LaunchMinBlocks<DepFillFn<2> > {
  static const int value = 5;
};
#endif

struct ConnRemapFn {
  const int *conn_local;
  const int *node_num;
  int *conn;
  TMR_HD void operator()(i64 j) const { conn[j] = node_num[conn_local[j]]; }
};

struct ConnRemapInPlaceFn {
  const int *node_num;
  int *conn;
  TMR_HD void operator()(i64 j) const { conn[j] = node_num[conn[j]]; }
};

/* node numbers as sortable unsigned keys and back */
struct NumToKeyFn {
  const int *num;
  u64 *keys;
  TMR_HD void operator()(i64 i) const {
    keys[i] = (u64)((u32)num[i] ^ 0x80000000u);
  }
};
struct KeyToNumFn {
  const u64 *keys;
  int *num;
  TMR_HD void operator()(i64 i) const {
    num[i] = (int)((u32)keys[i] ^ 0x80000000u);
  }
};

/* getNodeNumbers(): every local node number, ascending (reference :4246) */
struct NumberRangeFn {
  int first;
  int *out;
  TMR_HD void operator()(i64 i) const { out[i] = first + (int)i; }
};

inline int ensure_node_arrays(Forest &f);

/* every local node number, ascending, in a device array */
inline int build_sorted_numbers(Forest &f, DBuf<int> &out) {
  Ctx &ctx = *f.ctx;
  NodeData &nd = f.nodes;
  const i64 n = nd.num_local_nodes;
  if (!nd.valid) return 1;
  out.alloc(ctx, n);
  if (n == 0) return 0;
  if (!forest_comm(f)) {
    /* one rank: the numbers are -Nd..-1 (dependent) and 0..owned-1, every
       value once -- the sorted array is a range, no sort needed */
    NumberRangeFn r = {-(int)nd.num_dep_nodes, out.get()};
    launch(ctx, n, r, "nodes_number_range");
    return 0;
  }
  if (ensure_node_arrays(f)) return 1;
  DBuf<u64> k(ctx, n), k_alt(ctx, n);
  DBuf<u32> v0, v1;
  NumToKeyFn a = {nd.node_num.get(), k.get()};
  launch(ctx, n, a, "nodes_numbers_to_keys");
  radix_sort(ctx, k, k_alt, v0, v1, n, 0, 32);
  KeyToNumFn b = {k.get(), out.get()};
  launch(ctx, n, b, "nodes_keys_to_numbers");
  return 0;
}

inline int sorted_node_numbers(Forest &f, int *h_out) {
  Ctx &ctx = *f.ctx;
  DBuf<int> out;
  if (build_sorted_numbers(f, out)) return 1;
  copy_d2h(ctx, h_out, out.get(), (size_t)f.nodes.num_local_nodes * sizeof(int));
  return check_errors(ctx, "sorted_node_numbers");
}

/* ---- host mirrors of the node arrays -------------------------------------------
   Page-locked copies owned by the forest.  node_mirror_start enqueues the copy
   on the context's copy stream (ordered after the main stream's work so far),
   node_mirror_get waits for it.  createNodes starts the arrays named by
   NodeData::prefetch as soon as each is final, so the 5.9 GB read-back of the
   86 M-octant mesh overlaps the rest of createNodes instead of following it. */
/* number of host threads for the mirror expansion: the CPUs this process may
   run on (bench.py binds each rank to its GPU's NUMA node), at most 8 (measured on
   the C2 mesh: 2 / 4 / 8 / 16 threads -> 197 / 124 / 101 / 105 ms per end-to-end
   step; beyond 8 they take memory bandwidth from the DMA of conn) */
inline int mirror_threads() {
  int n = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
#endif
  if (const char *ev = getenv("TMR_B200_HOST_THREADS")) n = atoi(ev);
  return n < 1 ? 1 : (n > 8 ? 8 : n);
}

/* run body(t, T) on T threads (the caller is thread 0) */
template <class Body>
inline void host_parallel(int T, const Body &body) {
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.push_back(std::thread([&body, t, T]() { body(t, T); }));
  body(0, T);
  for (size_t k = 0; k < th.size(); k++) th[k].join();
}

/* dep_ptr and dep_weights from the stencil codes (DepExpandJob): lengths ->
   prefix sums -> weight rows, each of T threads on a contiguous slice.  The
   distinct codes of a mesh are few (4 edge rows and 16 face products at order
   2), so every code is first mapped to a dense class whose full weight row is
   tabulated once; the per-node work is then a table copy without
   data-dependent branches. */
inline void dep_expand_host(const unsigned short *codes, const double *table, i64 nd,
                            int order, int *ptr, double *w, int T) {
  const int n2 = order * order;
  /* classes: [0, 2 p) edge rows, [2 p, 2 p + 4 p^2) face products, last = none */
  const int nclass = 2 * order + 4 * n2 + 1;
  std::vector<unsigned short> cls(65536, (unsigned short)(nclass - 1));
  std::vector<int> len(nclass, 0);
  std::vector<double> rows((size_t)nclass * n2, 0.0);
  const double *face = table + (size_t)2 * order * order;
  for (int b = 0; b < 2; b++) {
    for (int k = 0; k < order; k++) {
      const int c = b * order + k;
      cls[kDepCodeEdge | (b << 4) | k] = (unsigned short)c;
      len[c] = order;
      for (int j = 0; j < order; j++) rows[(size_t)c * n2 + j] = table[(size_t)c * order + j];
    }
  }
  for (int b1 = 0; b1 < 2; b1++) {
    for (int i = 0; i < order; i++) {
      for (int b2 = 0; b2 < 2; b2++) {
        for (int j = 0; j < order; j++) {
          const int c = 2 * order + ((b1 * order + i) * 2 + b2) * order + j;
          cls[(b1 << 9) | (i << 5) | (b2 << 4) | j] = (unsigned short)c;
          len[c] = n2;
          const double *nu = face + (size_t)(b1 * order + i) * order;
          const double *nv = face + (size_t)(b2 * order + j) * order;
          for (int q = 0; q < n2; q++) rows[(size_t)c * n2 + q] = nu[q % order] * nv[q / order];
        }
      }
    }
  }
  std::vector<i64> sum(T + 1, 0);
  host_parallel(T, [&](int t, int TT) {
    const i64 a = nd * t / TT, b = nd * (t + 1) / TT;
    i64 s = 0;
    for (i64 d = a; d < b; d++) s += len[cls[codes[d]]];
    sum[t + 1] = s;
  });
  for (int t = 0; t < T; t++) sum[t + 1] += sum[t];
  host_parallel(T, [&](int t, int TT) {
    const i64 a = nd * t / TT, b = nd * (t + 1) / TT;
    i64 o = sum[t];
    i64 d = a;
    if (order == 2) {
      /* 4 doubles are stored for every node and the cursor advances by the
         true length (an edge row's two surplus values are overwritten by the
         next node); near the end of the slice's output range the exact path
         takes over so that nothing is written past it */
      const double *r4 = rows.data();
      const i64 end = sum[t + 1];
      for (; d < b && o + 4 <= end; d++) {
        const int c = cls[codes[d]];
        const double *r = r4 + 4 * c;
#if defined(__x86_64__)
        /* streaming stores: the arrays are written once and read by the
           caller much later -- no read-for-ownership traffic next to the DMA
           of conn that is landing in the same memory */
        _mm_stream_si32(ptr + d, (int)o);
        const long long *q = reinterpret_cast<const long long *>(r);
        long long *dst = reinterpret_cast<long long *>(w + o);
        _mm_stream_si64(dst, q[0]);
        _mm_stream_si64(dst + 1, q[1]);
        _mm_stream_si64(dst + 2, q[2]);
        _mm_stream_si64(dst + 3, q[3]);
#else
        ptr[d] = (int)o;
        w[o] = r[0];
        w[o + 1] = r[1];
        w[o + 2] = r[2];
        w[o + 3] = r[3];
#endif
        o += len[c];
      }
#if defined(__x86_64__)
      _mm_sfence();
#endif
    }
    for (; d < b; d++) {
      const int c = cls[codes[d]];
      ptr[d] = (int)o;
      const double *r = rows.data() + (size_t)c * n2;
      for (int j = 0; j < len[c]; j++) w[o + j] = r[j];
      o += len[c];
    }
    if (t == TT - 1) ptr[nd] = (int)o;
  });
}

/* dep_ptr + dep_weights mirrors: 2 bytes per dependent node over the bus
   (on their own copy lane, ahead of the large arrays), expanded by host
   threads in the background */
inline int dep_expand_start(Forest &f) {
  NodeData &nd = f.nodes;
  Ctx &ctx = *f.ctx;
  if (nd.mirror[kMirrorDepPtr].p) return 0;
  nd.mctx = &ctx;
  const i64 Nd = nd.num_dep_nodes;
  const int order = nd.order;
  int *ptr = static_cast<int *>(host_alloc(ctx, (size_t)(Nd + 1) * sizeof(int) + 16));
  double *w = static_cast<double *>(host_alloc(ctx, (size_t)nd.dep_nnz * sizeof(double) + 16));
  if (!ptr || !w) return 1;
  nd.mirror[kMirrorDepPtr].p = ptr;
  nd.mirror[kMirrorDepWeights].p = w;
  if (Nd == 0) {
    ptr[0] = 0;
    return 0;
  }
  DepExpandJob &job = nd.dep_job;
  const size_t tab_n = (size_t)4 * order * order;
  job.codes = host_alloc(ctx, (size_t)Nd * sizeof(unsigned short) + 16);
  job.table = static_cast<double *>(host_alloc(ctx, tab_n * sizeof(double)));
  if (!job.codes || !job.table) return 1;
  job.ev_table = copy_d2h_async(ctx, job.table, nd.dep_wtab.get(), tab_n * sizeof(double), 1);
  job.ev_codes = copy_d2h_async(ctx, job.codes, nd.dep_code.get(),
                                (size_t)Nd * sizeof(unsigned short), 1);
  Ctx *c = &ctx;
  void *ev_codes = job.ev_codes, *ev_table = job.ev_table;
  const unsigned short *codes = static_cast<const unsigned short *>(job.codes);
  const double *table = job.table;
  const int T = mirror_threads();
  job.thread = new std::thread([c, ev_codes, ev_table, codes, table, Nd, order, ptr, w, T]() {
    copy_sync(*c, ev_table);
    copy_sync(*c, ev_codes);
    dep_expand_host(codes, table, Nd, order, ptr, w, T);
  });
  return 0;
}

inline int node_mirror_start(Forest &f, int which) {
  NodeData &nd = f.nodes;
  Ctx &ctx = *f.ctx;
  if (which < 0 || which >= kNumMirrors) return 1;
  if (nd.mirror[which].p) return 0;
  if (which == kMirrorDepPtr || which == kMirrorDepWeights) return dep_expand_start(f);
  const void *src = NULL;
  size_t bytes = 0;
  const i64 npe = (i64)nd.order * nd.order * nd.order;
  switch (which) {
    case kMirrorConn:
      src = nd.conn.get();
      bytes = (size_t)(nd.num_elements * npe) * sizeof(int);
      break;
    case kMirrorNumbers:
      bytes = (size_t)nd.num_local_nodes * sizeof(int);
      if (!forest_comm(f)) {
        /* one rank: the numbers are -Nd..-1 (dependent) and 0..owned-1, every
           value once: the sorted array is a range, written by host threads
           (no sort, nothing over the bus) */
        int *p = static_cast<int *>(host_alloc(ctx, bytes + 16));
        if (!p) return 1;
        nd.mctx = &ctx;
        nd.mirror[which].p = p;
        const i64 n = nd.num_local_nodes;
        const int first = -(int)nd.num_dep_nodes;
        const int T = mirror_threads() > 4 ? 4 : mirror_threads();
        nd.mirror[which].thread = new std::thread([p, n, first, T]() {
          host_parallel(T, [p, n, first](int t, int TT) {
            const i64 a = n * t / TT, b = n * (t + 1) / TT;
            for (i64 i = a; i < b; i++) p[i] = first + (int)i;
          });
        });
        return 0;
      }
      if (nd.ext_numbers_valid) {
        /* several ranks, slot construction: dependents -Nd..-1, then this
           rank's owned range with the (few) numbers owned elsewhere merged in
           around it: only those cross the bus */
        const i64 nx = nd.ext_numbers.size();
        int *p = static_cast<int *>(host_alloc(ctx, bytes + 16));
        int *xs = static_cast<int *>(host_alloc(ctx, (size_t)nx * sizeof(int) + 16));
        if (!p || !xs) return 1;
        nd.mctx = &ctx;
        nd.mirror[which].p = p;
        void *ev = NULL;
        if (nx > 0) {
          DBuf<u64> k(ctx, nx), k_alt(ctx, nx);
          DBuf<u32> v0, v1;
          NumToKeyFn a = {nd.ext_numbers.get(), k.get()};
          launch(ctx, nx, a, "nodes_numbers_to_keys");
          radix_sort(ctx, k, k_alt, v0, v1, nx, 0, 32);
          nd.sorted_numbers.alloc(ctx, nx);
          KeyToNumFn b = {k.get(), nd.sorted_numbers.get()};
          launch(ctx, nx, b, "nodes_keys_to_numbers");
          ev = copy_d2h_async(ctx, xs, nd.sorted_numbers.get(), (size_t)nx * sizeof(int), 1);
        }
        Ctx *c = &ctx;
        const i64 Ndep = nd.num_dep_nodes, nown = nd.num_owned_nodes;
        const int start = nd.node_range_start;
        nd.mirror[which].ev = ev; /* released by node_mirror_get */
        nd.mirror[which].aux = xs;
        const int T = mirror_threads() > 4 ? 4 : mirror_threads();
        nd.mirror[which].thread = new std::thread([c, ev, p, xs, nx, Ndep, nown, start, T]() {
          copy_sync(*c, ev);
          i64 nlo = 0; /* externals owned by lower ranks */
          while (nlo < nx && xs[nlo] < start) nlo++;
          for (i64 q = 0; q < nlo; q++) p[Ndep + q] = xs[q];
          for (i64 q = nlo; q < nx; q++) p[Ndep + nown + q] = xs[q];
          int *pd = p, *po = p + Ndep + nlo;
          host_parallel(T, [pd, po, Ndep, nown, start](int t, int TT) {
            for (i64 i = Ndep * t / TT; i < Ndep * (t + 1) / TT; i++) pd[i] = (int)(i - Ndep);
            for (i64 i = nown * t / TT; i < nown * (t + 1) / TT; i++) po[i] = start + (int)i;
          });
        });
        return 0;
      }
      if (build_sorted_numbers(f, nd.sorted_numbers)) return 1;
      src = nd.sorted_numbers.get();
      break;
    default:
      src = nd.dep_conn.get();
      bytes = (size_t)nd.dep_nnz * sizeof(int);
      break;
  }
  nd.mctx = &ctx;
  void *p = host_alloc(ctx, bytes + 16);
  if (!p) return 1;
  nd.mirror[which].p = p;
  nd.mirror[which].ev = copy_d2h_async(ctx, p, src, bytes);
  return 0;
}

inline const void *node_mirror_get(Forest &f, int which) {
  NodeData &nd = f.nodes;
  if (!nd.valid || node_mirror_start(f, which)) return NULL;
  if (which == kMirrorDepPtr || which == kMirrorDepWeights) nd.finish_dep_job();
  nd.mirror[which].join();
  if (nd.mirror[which].ev) {
    copy_wait(*f.ctx, nd.mirror[which].ev);
    nd.mirror[which].ev = NULL;
    if (which == kMirrorNumbers) nd.sorted_numbers.reset();
  }
  return nd.mirror[which].p;
}

/* packed family emission through expand_u64: key | payload << pshift */
template <class Sink>
struct PackedEmit {
  Sink &s;
  int pshift;
  TMR_HD void operator()(u64 key, u64 payload) { s(key | (payload << pshift)); }
};
struct NodeEmitPackedFn {
  NodeEmit g;
  int pshift;
  template <class Sink>
  TMR_HD void operator()(i64 e, Sink &sink) const {
    PackedEmit<Sink> pe = {sink, pshift};
    g.template run_order<2>(e, pe); /* staged path: 8 outputs per item at most */
  }
};

template <int kOrder>
struct NodeEmitPlaceFn {
  NodeEmitFillFn<kOrder> fill;
  const u32 *off;
  TMR_HD void operator()(i64 e) const { fill(e, off[e]); }
};

struct ParentPlaceFn {
  ParentNodeFillFn fill;
  const u32 *off;
  TMR_HD void operator()(i64 e) const { fill(e, off[e]); }
};

struct U32DestFn {
  const u32 *dest;
  TMR_HD int operator()(i64 i) const { return (int)dest[i]; }
};

/* home rank of a node = owner of its position (the reference distributes the
   node array with matchOctantIntervals, :4545) */
struct NodeHomeFn {
  const u64 *node_keys;
  int Dn;
  int lbits;
  OwnerMap om; /* positions at depth Dn */
  TMR_HD int operator()(i64 i) const {
    const u64 k = node_keys[i] >> lbits;
    const int sh = 3 * (Dn + 1);
    const u64 block = k >> sh;
    /* halving every squeezed coordinate = shifting the interleaved code by 3 */
    const u64 m = (k & low_mask(sh)) >> 3;
    return om.owner((block << (3 * Dn)) | m);
  }
};

/* at the home rank: value of a received item = donating rank or INF */
struct DonorValueFn {
  const unsigned char *created; /* received flags */
  const i64 *recv_off;          /* device, R+1 */
  int R;
  u32 *val;
  u32 *idx;
  TMR_HD void operator()(i64 i) const {
    int src = 0;
    while (src < R - 1 && recv_off[src + 1] <= i) src++;
    val[i] = created[i] ? (u32)src : 0x7fffffffu;
    idx[i] = (u32)i;
  }
};

struct OwnerMinFn { /* scan_apply body over the key-sorted received items */
  const u64 *keys;
  const u32 *idx;
  const u32 *val; /* by original received index */
  int *owner_run;
  u32 *run_of;
  TMR_HD void operator()(i64 j, u32 heads_before) const {
    const bool head = (j == 0 || keys[j] != keys[j - 1]);
    const u32 run = heads_before + (head ? 1u : 0u) - 1u;
    run_of[j] = run;
    TMR_ATOMIC_MIN_I32(&owner_run[run], (int)val[idx[j]]);
  }
};

/* owner of a node whose home is this rank starts as "me if I create it" */
struct ForeignNodeCountFn {
  NodeHomeFn home;
  int me;
  TMR_HD u32 operator()(i64 i) const { return home(i) != me ? 1u : 0u; }
};

struct ForeignNodeFillFn {
  NodeHomeFn home;
  int me;
  const u64 *node_keys;
  const unsigned char *created;
  u64 *out_key;
  u32 *out_dest;
  u32 *out_index;
  unsigned char *out_created;
  int *owner; /* also initialised here: one pass over the nodes instead of two */
  TMR_HD void operator()(i64 i, u32 o) const {
    const int d = home(i);
    owner[i] = (d == me && created[i]) ? me : 0x7fffffff;
    if (d != me) {
      out_key[o] = node_keys[i];
      out_dest[o] = (u32)d;
      out_index[o] = (u32)i;
      out_created[o] = created[i];
    }
  }
};

/* home side: received donors lower the owner of my own copy of the node */
struct HomeFoldFn {
  const u64 *rkeys; /* received keys, sorted */
  const u32 *run_of;
  const int *owner_run;
  const u64 *node_keys;
  i64 n;
  int *owner;
  TMR_HD void operator()(i64 j) const {
    const i64 i = find_u64(node_keys, n, rkeys[j]);
    if (i >= 0) TMR_ATOMIC_MIN_I32(&owner[i], owner_run[run_of[j]]);
  }
};

struct HomeReplyFn {
  const u64 *rkeys;
  const u32 *idx;
  const u32 *run_of;
  const int *owner_run;
  const u64 *node_keys;
  i64 n;
  const int *owner;
  int *reply; /* by original received index */
  TMR_HD void operator()(i64 j) const {
    const i64 i = find_u64(node_keys, n, rkeys[j]);
    int o = owner_run[run_of[j]];
    if (i >= 0 && owner[i] < o) o = owner[i];
    reply[idx[j]] = o;
  }
};

struct StoreOwnerFn {
  const u32 *index;
  const int *value;
  int *owner;
  TMR_HD void operator()(i64 k) const { owner[index[k]] = value[k]; }
};

struct OwnerFinishFn {
  int *owner;
  TMR_HD void operator()(i64 i) const {
    if (owner[i] == 0x7fffffff) owner[i] = -1;
  }
};

struct FillIntFn {
  int *p;
  int v;
  TMR_HD void operator()(i64 i) const { p[i] = v; }
};

/* classification of local nodes: 0 dependent, 1 owned here, 2 external */
struct OwnedFlagFn {
  const unsigned char *dep_flag;
  const int *owner; /* NULL on a single rank */
  int me;
  TMR_HD u32 operator()(i64 i) const {
    return (!dep_flag[i] && (!owner || owner[i] == me)) ? 1u : 0u;
  }
};

struct ExternalCountFn {
  const unsigned char *dep_flag;
  const int *owner;
  int me;
  TMR_HD u32 operator()(i64 i) const {
    return (!dep_flag[i] && owner[i] != me) ? 1u : 0u;
  }
};

/* orders >= 4: node-level copies of the entity owners */
struct ExpandOwnerFn {
  EntityMultFn mult;
  const u32 *ent_off;
  const int *ent_owner;
  u32 *ent_of;
  int *node_owner;
  TMR_HD void operator()(i64 j) const {
    const u32 m = mult(j), o = ent_off[j];
    for (u32 k = 0; k < m; k++) {
      ent_of[o + k] = (u32)j;
      node_owner[o + k] = ent_owner[j];
    }
  }
};

/* orders >= 4: a request names the entity key and, in bits 56..63, which of
   its nodes is meant */
static const int kSubNodeShift = 56;

struct NumberNodesMultiFn {
  const unsigned char *dep_flag;
  const u32 *dep_before;
  const u32 *owned_before;
  const int *owner;
  int me;
  int first_owned;
  int *node_num;
  int *dep_node;
  TMR_HD void operator()(i64 i) const {
    if (dep_flag[i]) {
      node_num[i] = -(int)dep_before[i] - 1;
      dep_node[dep_before[i]] = (int)i;
    } else if (!owner || owner[i] == me) {
      node_num[i] = first_owned + (int)owned_before[i];
    } else {
      node_num[i] = 0; /* filled from the owner's reply */
    }
  }
};

/* lists the nodes owned elsewhere and, in the same pass, numbers every node
   (NumberNodesMultiFn) */
struct ExternalFillFn {
  ExternalCountFn c;
  NumberNodesMultiFn number;
  const u64 *node_keys;
  u64 *out_keys;
  u32 *out_dest;
  u32 *out_node;
  const u32 *ent_of;  /* NULL below order 4 */
  const u32 *ent_off;
  TMR_HD void operator()(i64 i, u32 o) const {
    number(i);
    if (c(i)) {
      if (ent_of) {
        const u32 e = ent_of[i];
        out_keys[o] = node_keys[e] | ((u64)((u32)i - ent_off[e]) << kSubNodeShift);
      } else {
        out_keys[o] = node_keys[i];
      }
      out_dest[o] = (u32)(c.owner[i] < 0 ? c.me : c.owner[i]);
      out_node[o] = (u32)i;
    }
  }
};


struct LookupNumberFn { /* owner side: number of each requested node key */
  const u64 *req;
  const u64 *node_keys;
  i64 n;
  const int *node_num;
  int *reply;
  const u32 *ent_off; /* NULL below order 4 */
  TMR_HD void operator()(i64 i) const {
    if (ent_off) {
      const u64 sub = req[i] >> kSubNodeShift;
      const i64 j = find_u64(node_keys, n, req[i] & low_mask(kSubNodeShift));
      reply[i] = j >= 0 ? node_num[ent_off[j] + sub] : -1;
    } else {
      const i64 j = find_u64(node_keys, n, req[i]);
      reply[i] = j >= 0 ? node_num[j] : -1;
    }
  }
};

struct StoreExternalFn {
  const u32 *ext_node;
  const int *number;
  int *node_num;
  TMR_HD void operator()(i64 i) const { node_num[ext_node[i]] = number[i]; }
};

/* ---- slot construction of nodes + connectivity (ops_nodes_slots.h) ---------- */
template <class M>
struct ParentSlotFn {
  ParentNodeGen g;
  SlotView v;
  NodeSlotFn<M> ns;
  const u32 *list; /* the elements with a foreign bit */
  const unsigned long long *count;
  TMR_HD void operator()(i64 q) const {
    if ((unsigned long long)q >= *count) return;
    SlotKeyEmit<NodeSlotFn<M> > em = {&v, &ns.nfmt, &ns};
    g.run((i64)list[q], em);
  }
};

/* the locate pass (and, on several ranks, the parent nodes of remote hanging
   faces), instantiated for the Morton word the depth needs */
template <class M>
inline void launch_slot_locate(Ctx &ctx, Forest &f, NodeData &nd, const SlotView &v,
                               unsigned char *slot8, const unsigned char *dep_table,
                               u64 *b_key, u32 *b_pay, unsigned long long *b_count,
                               i64 cap, const unsigned char *fmask,
                               const u32 *foreign_list,
                               const unsigned long long *foreign_count,
                               i64 foreign_cap) {
  NodeSlotFn<M> ns = {v, reinterpret_cast<u32 *>(nd.conn.get()), slot8, f.info.get(),
                      dep_table, nd.nfmt, b_key, b_pay, b_count, cap};
  if (!v.multi) {
    NodeSlotFn<M, false> ns1 = {v, reinterpret_cast<u32 *>(nd.conn.get()), slot8,
                                f.info.get(), dep_table, nd.nfmt, b_key, b_pay, b_count, cap};
    launch_block3(ctx, f.n, ns1, "nodes_slot_locate");
    return;
  }
  launch_block3(ctx, f.n, ns, "nodes_slot_locate");
  if (v.multi && fmask && foreign_cap > 0) {
    ParentNodeGen pg = {f.keys.get(), fmask, f.fmt, nd.nfmt, f.tables, 2, NULL};
    ParentSlotFn<M> ps = {pg, v, ns, foreign_list, foreign_count};
    launch(ctx, foreign_cap, ps, "nodes_slot_parents");
  }
}

/* B nodes (several ranks): sorted (key, payload) -> unique keys, run index */
struct BUniqueFn {
  const u64 *k;
  const u32 *pay;
  u64 *ukeys;
  unsigned char *ucreated;
  unsigned char *udep; /* some element labels the node dependent */
  u32 *run_of;
  TMR_HD void operator()(i64 i, u32 heads_before) const {
    const bool head = (i == 0 || k[i] != k[i - 1]);
    const u32 run = heads_before + (head ? 1u : 0u) - 1u;
    if (head) ukeys[run] = k[i];
    run_of[i] = run;
    if (pay[i] != kConnB) {
      ucreated[run] = 1;
      if (pay[i] & kPayDep) udep[run] = 1;
    }
  }
};
struct BLowCountFn {
  const u64 *ukeys;
  i64 n;
  u64 first_key; /* node-key image of this rank's first position */
  i64 *out;
  TMR_HD void operator()(i64) const { out[0] = lower_bound_u64(ukeys, n, first_key); }
};
/* connectivity entries of the corners outside the rank's range: the final
   numbers of their B nodes */
struct BPlaceFn {
  const u32 *pay;
  const u32 *run_of;
  const int *unum; /* number of every unique B node */
  int *conn;
  TMR_HD void operator()(i64 i) const {
    if (pay[i] != kConnB) conn[pay[i] & ~kPayDep] = unum[run_of[i]];
  }
};
struct OwnerFromReplyFn {
  const int *got;
  int *owner;
  TMR_HD void operator()(i64 r) const { owner[r] = got[r] == 0x7fffffff ? -1 : got[r]; }
};
struct BHomeDestFn {
  NodeHomeFn home;
  TMR_HD int operator()(i64 r) const { return home(r); }
};

/* 1 = done; 0 = the forest needs the general candidate sort (agreed on by
   all ranks); < 0 error.  Dependent nodes are labelled in slot space and the
   connectivity comes out as final node NUMBERS (node_keys, node_num, the
   counts and node_range of `nd` are filled too).  Several ranks: `om_n` maps
   node positions to their home rank. */
/* several ranks: the state of the slot construction node_keys / node_num are
   rebuilt from on request, and the dependent stencils look nodes up in */
struct SlotMulti {
  DBuf<u64> mc, slotinfo1, xref, b_ukeys;
  DBuf<SlotInfoM> slotinfo;
  DBuf<int> xnum, b_num;
  DBuf<RankEntry> rank_tab;
  DBuf<u32> rank_cells;
  SlotNumbers sn;
  SlotView v;
  i64 nbu, nlow, NA;
};

/* node_keys and node_num on request (slot construction) */
inline int ensure_node_arrays(Forest &f) {
  NodeData &nd = f.nodes;
  Ctx &ctx = *f.ctx;
  if (!nd.valid) return 1;
  if (nd.num_local_nodes == 0 || nd.node_num.size() == nd.num_local_nodes) return 0;
  if (nd.slot_multi) {
    SlotMulti &sm = *static_cast<SlotMulti *>(nd.slot_multi.get());
    nd.node_keys.alloc(ctx, nd.num_local_nodes);
    nd.node_num.alloc(ctx, nd.num_local_nodes);
    if (sm.slotinfo1.size() != f.n) {
      sm.slotinfo1.alloc(ctx, f.n);
      SlotCountMFn sc1 = {sm.mc.get()};
      SlotInfoMFn si1 = {sm.mc.get(), sm.slotinfo1.get()};
      scan_apply(ctx, f.n, sc1, si1, "nodes_slot_scan");
    }
    SlotKeys3Fn kf = {f.keys.get(), f.fmt, sm.slotinfo1.get(), sm.sn, nd.node_keys.get(),
                      nd.node_num.get(), sm.nlow};
    launch(ctx, f.n, kf, "nodes_slot_keys");
    BNodeArraysFn bf = {sm.b_ukeys.get(), sm.b_num.get(), sm.nlow, sm.NA,
                        nd.node_keys.get(), nd.node_num.get()};
    launch(ctx, sm.nbu, bf, "nodes_b_keys");
    return 0;
  }
  if (nd.slot_info.size() != f.n) return 1;
  nd.node_keys.alloc(ctx, nd.num_local_nodes);
  nd.node_num.alloc(ctx, nd.num_local_nodes);
  SlotKeys2Fn kf = {f.keys.get(), f.fmt, nd.slot_info.get(), nd.node_keys.get(),
                    nd.node_num.get()};
  launch(ctx, f.n, kf, "nodes_slot_keys");
  return 0;
}

struct SlotState { /* outlives build_nodes_slots: the dependent CSR looks nodes up in it */
  DBuf<RankEntry> rank_tab;
  DBuf<u32> rank_cells;
  SlotLookup lookup;
  SlotLookupM lookup_m;
  /* winners of the dependent stencils (DepWinnerFn), found while the
     connectivity was resolved */
  DBuf<u64> win_edge, win_face;
  int winner_done;
  /* in: elements with a remote coarse neighbour (PatchProbeFn) */
  const u32 *foreign_list;
  const unsigned long long *foreign_count;
  i64 foreign_cap;
  SlotState() : winner_done(0), foreign_list(NULL), foreign_count(NULL), foreign_cap(0) {
    lookup.on = 0;
    lookup_m.on = 0;
  }
};

inline int build_nodes_slots(Forest &f, NodeData &nd,
                             const unsigned char *fmask, u64 k_first, u64 k_last,
                             const OwnerMap &om_n, i64 *Nn_out, SlotState &st) {
  Ctx &ctx = *f.ctx;
  Comm *comm = forest_comm(f);
  const int me = comm ? comm->rank : 0;
  const i64 E = f.n;
  const int D = f.fmt.D;
  DBuf<u32> mask, dmask(ctx, E);
  DBuf<u64> mc;
  dev_zero(ctx, dmask.get(), (size_t)E * sizeof(u32));
  if (comm) {
    mc.alloc(ctx, E);
    dev_zero(ctx, mc.get(), (size_t)E * sizeof(u64));
  } else {
    mask.alloc(ctx, E);
    dev_zero(ctx, mask.get(), (size_t)E * sizeof(u32));
  }
  DBuf<unsigned long long> ctl(ctx, 2); /* [0] B count, [1] fail flag */
  DBuf<unsigned char> slot8(ctx, E * 8);
  DBuf<unsigned char> dep_table(ctx, 512);
  DepTableFn dt = {dep_table.get()};
  launch(ctx, 512, dt, "nodes_dep_table");
  /* this rank's range of positions and the rank index over it */
  const u64 pos_lo = E > 0 ? (k_first >> 5) : 0ULL;
  const u64 pos_hi =
      E > 0 ? (k_last >> 5) + (1ULL << (3 * (D - (int)(k_last & 31)))) : 0ULL;
  DBuf<RankEntry> &rank_tab = st.rank_tab;
  DBuf<u32> &rank_cells = st.rank_cells;
  size_t ix_budget = (size_t)8 * (size_t)E > ((size_t)64 << 20)
                         ? (size_t)8 * (size_t)E
                         : ((size_t)64 << 20);
  if (const char *ev = getenv("TMR_B200_RANK_BUDGET")) ix_budget = (size_t)atol(ev);
  RankIndex elem_ix = {NULL, NULL, 0, 0};
  if (E > 0) {
    elem_ix = build_rank_index(ctx, f.keys.get(), E, D, pos_lo, pos_hi, ix_budget,
                               rank_tab, rank_cells);
  }
  DBuf<u64> b_key;
  DBuf<u32> b_pay;
  i64 cap = comm ? (E / 2 + 65536) : 0, nb = 0;
  unsigned long long h_ctl[2] = {0, 0};
  SlotView v = {f.keys.get(), E, f.fmt, f.tables, elem_ix, comm ? 1 : 0, pos_lo,
                pos_hi, mask.get(), mc.get(), dmask.get(), NULL};
  for (int attempt = 0; attempt < 2; attempt++) {
    if (comm) {
      b_key.alloc(ctx, cap);
      b_pay.alloc(ctx, cap);
    }
    dev_zero(ctx, ctl.get(), 2 * sizeof(unsigned long long));
    v.fail = reinterpret_cast<int *>(ctl.get() + 1);
    if (E > 0) {
      if (3 * D <= 30) {
        launch_slot_locate<u32>(ctx, f, nd, v, slot8.get(), dep_table.get(),
                                b_key.get(), b_pay.get(), ctl.get(), cap, fmask,
                                st.foreign_list, st.foreign_count, st.foreign_cap);
      } else {
        launch_slot_locate<u64>(ctx, f, nd, v, slot8.get(), dep_table.get(),
                                b_key.get(), b_pay.get(), ctl.get(), cap, fmask,
                                st.foreign_list, st.foreign_count, st.foreign_cap);
      }
    }
    if (!comm) break;
    copy_d2h(ctx, h_ctl, ctl.get(), sizeof(h_ctl));
    nb = (i64)h_ctl[0];
    if (h_ctl[1] || nb <= cap) break;
    cap = nb; /* rare: more off-range corners than the guess; marks are idempotent */
  }
  if (!comm) {
    /* ---- one rank: scan nodes and dependents together, number directly ---- */
    DBuf<SlotInfo2> slotinfo(ctx, E);
    SlotCount2Fn sc = {mask.get(), dmask.get()};
    SlotInfo2Fn si = {mask.get(), dmask.get(), slotinfo.get()};
    const u64 tot = scan_apply(ctx, E, sc, si, "nodes_slot_scan");
    copy_d2h(ctx, h_ctl, ctl.get(), sizeof(h_ctl));
    if (h_ctl[1]) return 0;
    const i64 Nn = (i64)(tot & 0x7fffffffULL), Nd = (i64)(tot >> 31);
    SlotResolve2Fn rs = {slotinfo.get(), slot8.get(),
                         reinterpret_cast<u32 *>(nd.conn.get())};
    st.win_edge.alloc(ctx, Nd);
    st.win_face.alloc(ctx, Nd);
    dev_zero(ctx, st.win_edge.get(), (size_t)Nd * sizeof(u64));
    dev_zero(ctx, st.win_face.get(), (size_t)Nd * sizeof(u64));
    DepWinnerFn<2> win = {f.keys.get(), f.info.get(), f.fmt, 2, nd.conn.get(),
                          st.win_edge.get(), st.win_face.get()};
    SlotResolveWin2Fn rw = {rs, win};
    launch(ctx, E, rw, "nodes_slot_resolve");
    st.winner_done = 1;
    /* node_keys / node_num (1.5 GB at 86 M octants) are not needed by anything
       createNodes produces: they are written on first request from the
       per-leaf counts kept here (ensure_node_arrays) */
    nd.slot_info.swap(slotinfo);
    st.lookup.on = 1;
    st.lookup.v = v;
    st.lookup.v.fail = NULL;
    st.lookup.si = nd.slot_info.get();
    nd.num_candidates = 0;
    nd.num_dep_nodes = Nd;
    nd.num_owned_nodes = Nn - Nd;
    nd.node_range_start = 0;
    nd.node_range.assign(2, 0);
    nd.node_range[1] = (int)(Nn - Nd);
    *Nn_out = Nn;
    return 1;
  }
  /* ---- several ranks.  Every rank takes the same path from here on: the
     exchanges below are collective ---- */
  if (global_max(ctx, *comm, h_ctl[1] ? 1 : 0)) return 0;
  const int R = comm->size;
  /* B nodes: sort, unique */
  i64 nbu = 0, nlow = 0;
  DBuf<u64> b_ukeys;
  DBuf<unsigned char> b_ucreated, b_udep;
  DBuf<u32> b_run;
  if (nb > 0) {
    DBuf<u64> k_alt(ctx, nb);
    DBuf<u32> p_alt(ctx, nb);
    b_key.set_size(nb);
    b_pay.set_size(nb);
    radix_sort(ctx, b_key, k_alt, b_pay, p_alt, nb, 0, nd.nfmt.total_bits(), "nodes_b");
    b_ukeys.alloc(ctx, nb);
    b_ucreated.alloc(ctx, nb);
    b_udep.alloc(ctx, nb);
    b_run.alloc(ctx, nb);
    dev_zero(ctx, b_ucreated.get(), (size_t)nb);
    dev_zero(ctx, b_udep.get(), (size_t)nb);
    RunHeadFn rh = {b_key.get()};
    BUniqueFn bu = {b_key.get(), b_pay.get(), b_ukeys.get(), b_ucreated.get(),
                    b_udep.get(), b_run.get()};
    nbu = (i64)scan_apply(ctx, nb, rh, bu, "nodes_b_unique");
    if (E > 0) {
      DBuf<i64> d_low(ctx, 1);
      BLowCountFn lc = {b_ukeys.get(), nbu, pos_lo << 3, d_low.get()};
      launch(ctx, 1, lc, "nodes_b_low");
      copy_d2h(ctx, &nlow, d_low.get(), sizeof(i64));
    }
  }
  /* number of slot nodes; their positions in node order (a second array of 8
     bytes per leaf) are only needed for node_keys / node_num and are scanned
     when those are asked for (ensure_node_arrays) */
  SlotCountMFn sc1 = {mc.get()};
  SlotNoStoreFn none;
  const i64 NA = (i64)scan_apply(ctx, E, sc1, none, "nodes_slot_scan");
  const i64 Nn = NA + nbu;
  if (Nn >= (1LL << 31)) {
    fprintf(stderr, "TMROctForest Error: too many local nodes\n");
    return -1;
  }
  /* ---- ownership: the B nodes go to their home ranks (reference :4538-4637) ---- */
  DBuf<u32> lowmask(ctx, E);
  dev_zero(ctx, lowmask.get(), (size_t)E * sizeof(u32));
  DBuf<int> b_owner(ctx, nbu);
  DBuf<u64> xref, xref_alt;
  DBuf<u32> xdest, xdest_alt;
  i64 nx = 0;
  {
    NodeHomeFn home = {b_ukeys.get(), nd.nfmt.Dn, nd.nfmt.lbits, om_n};
    BHomeDestFn hdest = {home};
    RoutePlan plan;
    make_route(ctx, *comm, nbu, hdest, plan);
    DBuf<u64> rk;
    DBuf<unsigned char> rc;
    route_array(ctx, *comm, plan, b_ukeys.get(), rk);
    route_array(ctx, *comm, plan, b_ucreated.get(), rc);
    const i64 nr = plan.nrecv;
    DBuf<int> reply(ctx, nr);
    DBuf<unsigned long long> xcount(ctx, 1);
    dev_zero(ctx, xcount.get(), sizeof(unsigned long long));
    xref.alloc(ctx, nr);
    xdest.alloc(ctx, nr);
    if (nr > 0) {
      DBuf<i64> d_roff(ctx, R + 1);
      copy_h2d(ctx, d_roff.get(), plan.recv_off.data(), (size_t)(R + 1) * sizeof(i64));
      DBuf<u32> val(ctx, nr), idx(ctx, nr), idx_alt(ctx, nr);
      DonorValueFn dv = {rc.get(), d_roff.get(), R, val.get(), idx.get()};
      launch(ctx, nr, dv, "nodes_donor_value");
      DBuf<u64> rk_alt(ctx, nr);
      radix_sort(ctx, rk, rk_alt, idx, idx_alt, nr, 0, nd.nfmt.total_bits());
      DBuf<int> owner_run(ctx, nr);
      FillIntFn fi2 = {owner_run.get(), 0x7fffffff};
      launch(ctx, nr, fi2, "nodes_owner_run_init");
      DBuf<u32> run_of(ctx, nr);
      RunHeadFn rh = {rk.get()};
      OwnerMinFn omin = {rk.get(), idx.get(), val.get(), owner_run.get(), run_of.get()};
      scan_apply(ctx, nr, rh, omin, "nodes_owner_min");
      HomeSlotFn hs = {v,          rk.get(),    idx.get(),   run_of.get(), owner_run.get(),
                       me,         lowmask.get(), xref.get(), xdest.get(), xcount.get(),
                       reply.get()};
      launch(ctx, nr, hs, "nodes_owner_home");
      unsigned long long h_nx = 0;
      copy_d2h(ctx, &h_nx, xcount.get(), sizeof(h_nx));
      nx = (i64)h_nx;
    }
    DBuf<int> got;
    route_back(ctx, *comm, plan, reply.get(), got);
    OwnerFromReplyFn fr = {got.get(), b_owner.get()};
    launch(ctx, nbu, fr, "nodes_owner_store");
  }
  /* second scan: dependents and owned independents per leaf */
  DBuf<SlotInfoM> slotinfo(ctx, E);
  SlotCount3Fn sc3 = {mc.get(), dmask.get(), lowmask.get()};
  SlotInfo3Fn si3 = {sc3, slotinfo.get()};
  const u64 tot3 = scan_apply(ctx, E, sc3, si3, "nodes_slot_scan2");
  const i64 dep_slots = (i64)(tot3 & 0x7fffffffULL), own_slots = (i64)(tot3 >> 31);
  /* the same two counts over the B nodes */
  DBuf<u64> b_before(ctx, nbu + 1);
  BCountFn bc = {b_udep.get(), b_owner.get(), me};
  BStoreFn bs = {b_before.get()};
  const u64 totb = scan_apply(ctx, nbu, bc, bs, "nodes_b_scan");
  u64 lowb = 0;
  if (nlow >= nbu) {
    lowb = totb;
  } else {
    copy_d2h(ctx, &lowb, b_before.get() + nlow, sizeof(u64));
  }
  const i64 Nd = dep_slots + (i64)(totb & 0x7fffffffULL);
  const i64 nown = own_slots + (i64)(totb >> 31);
  std::vector<i64> all(R);
  comm->allgather_host(ctx, &nown, all.data(), sizeof(i64));
  i64 start = 0;
  for (int r = 0; r < me; r++) start += all[r];
  /* reference node_range (:4165-4172): kept so that the getters stay local */
  nd.node_range.assign(R + 1, 0);
  for (int r = 0; r < R; r++) nd.node_range[r + 1] = nd.node_range[r] + (int)all[r];
  nd.num_owned_nodes = nown;
  nd.node_range_start = (int)start;
  nd.num_dep_nodes = Nd;
  BNumbers bn = {b_before.get(), b_udep.get(), b_owner.get(), nlow, me, (int)start,
                 (int)dep_slots, (int)own_slots};
  DBuf<int> b_num(ctx, nbu);
  BNumberFn bnf = {bn, b_num.get()};
  launch(ctx, nbu, bnf, "nodes_b_number");
  /* ---- numbers of the nodes owned elsewhere (reference :4183-4235) ---- */
  if (nx > 1) {
    xref.set_size(nx);
    xdest.set_size(nx);
    xref_alt.alloc(ctx, nx);
    xdest_alt.alloc(ctx, nx);
    radix_sort(ctx, xref, xref_alt, xdest, xdest_alt, nx, 0, 37);
  }
  DBuf<int> xnum(ctx, nx);
  SlotNumbers sn = {slotinfo.get(), xref.get(), xnum.get(), nx,
                    (int)(lowb & 0x7fffffffULL), (int)start + (int)(lowb >> 31)};
  {
    DBuf<u64> qk(ctx, nx + nbu);
    DBuf<u32> qd(ctx, nx + nbu), qn(ctx, nbu);
    XrefKeyFn xk = {f.keys.get(), D, xref.get(), qk.get()};
    launch(ctx, nx, xk, "nodes_external_keys");
    if (nx > 0) copy_d2d(ctx, qd.get(), xdest.get(), (size_t)nx * sizeof(u32));
    BExternalCountFn xc = {b_udep.get(), b_owner.get(), me};
    BExternalFillFn xf = {xc, b_ukeys.get(), qk.get() + nx, qd.get() + nx, qn.get()};
    const i64 nxb = (i64)scan_apply(ctx, nbu, xc, xf, "nodes_external_list");
    U32DestFn xdest_fn = {qd.get()};
    RoutePlan plan;
    make_route(ctx, *comm, nx + nxb, xdest_fn, plan);
    DBuf<u64> req;
    route_array(ctx, *comm, plan, qk.get(), req);
    DBuf<int> rep(ctx, plan.nrecv);
    SlotNumbers sn_own = sn;
    sn_own.nx = 0; /* a node asked for here is owned here: never external */
    LookupSlotNumberFn lk = {req.get(), v, sn_own, b_ukeys.get(), nbu, bn, rep.get()};
    launch(ctx, plan.nrecv, lk, "nodes_external_lookup");
    DBuf<int> got;
    route_back(ctx, *comm, plan, rep.get(), got);
    if (nx > 0) copy_d2d(ctx, xnum.get(), got.get(), (size_t)nx * sizeof(int));
    StoreBExternalFn se = {qn.get(), got.get() + nx, b_num.get()};
    launch(ctx, nxb, se, "nodes_external_store");
    /* kept for getNodeNumbers(): with them the sorted list of local numbers
       is two ranges and a short merge (node_mirror_start) */
    nd.ext_numbers.swap(got);
    nd.ext_numbers.set_size(nx + nxb);
    nd.ext_numbers_valid = (Nd + nown + nx + nxb == Nn);
  }
  /* ---- connectivity (node_keys / node_num follow on request:
     ensure_node_arrays) ---- */
  SlotResolve3Fn rs = {sn, slot8.get(), reinterpret_cast<u32 *>(nd.conn.get())};
  st.win_edge.alloc(ctx, Nd);
  st.win_face.alloc(ctx, Nd);
  dev_zero(ctx, st.win_edge.get(), (size_t)Nd * sizeof(u64));
  dev_zero(ctx, st.win_face.get(), (size_t)Nd * sizeof(u64));
  DepWinnerFn<2> win = {f.keys.get(), f.info.get(), f.fmt, 2, nd.conn.get(),
                        st.win_edge.get(), st.win_face.get()};
  DBuf<unsigned char> pending(ctx, E);
  dev_zero(ctx, pending.get(), (size_t)E);
  SlotResolveWin3Fn rw = {rs, win, pending.get()};
  launch(ctx, E, rw, "nodes_slot_resolve");
  if (nb > 0) {
    BPlaceFn bp = {b_pay.get(), b_run.get(), b_num.get(), nd.conn.get()};
    launch(ctx, nb, bp, "nodes_b_place");
    PendingWinnerFn pw = {pending.get(), win};
    launch(ctx, E, pw, "nodes_dep_winner_b");
  }
  st.winner_done = 1;
  /* keep what the node arrays and the stencil look-ups are built from */
  {
    std::shared_ptr<SlotMulti> keep(new SlotMulti());
    SlotMulti &sm = *keep;
    sm.mc.swap(mc);
    sm.slotinfo.swap(slotinfo);
    sm.xref.swap(xref);
    sm.xnum.swap(xnum);
    sm.b_ukeys.swap(b_ukeys);
    sm.b_num.swap(b_num);
    sm.rank_tab.swap(st.rank_tab);
    sm.rank_cells.swap(st.rank_cells);
    sm.sn = sn;
    sm.v = v;
    sm.v.fail = NULL;
    sm.nbu = nbu;
    sm.nlow = nlow;
    sm.NA = NA;
    st.lookup_m.on = 1;
    st.lookup_m.v = sm.v;
    st.lookup_m.sn = sm.sn;
    st.lookup_m.nfmt = nd.nfmt;
    st.lookup_m.b_ukeys = sm.b_ukeys.get();
    st.lookup_m.nbu = nbu;
    st.lookup_m.b_num = sm.b_num.get();
    nd.slot_multi = keep;
  }
  nd.num_candidates = nb;
  *Nn_out = Nn;
  return 1;
}

inline int create_nodes(Forest &f, int order, int interp_type,
                        const double *knots) {
  Ctx &ctx = *f.ctx;
  Comm *comm = forest_comm(f);
  const int me = comm ? comm->rank : 0;
  NodeData &nd = f.nodes;
  if (nd.valid) return 0; /* reference :4071-4075 */
  /* Bernstein points (type 2) at order 2: same knots (-1, 1), node labels and
     label ranges as the Lagrange case, and eval_bernstein_weights(2, u) /
     bernstein_shape_functions(2, u) = {(1-u)/2, (1+u)/2} are the linear
     Lagrange functions evaluated at dyadic u -- bit-identical results
     (reference src/TMRInterpolation.h:164-183,309-322, :5364-5374).  Order-3
     Bernstein needs edge/face/block node labels (initLabel :6798-6811). */
  if (order < 2 || order > kMaxOrder) {
    fprintf(stderr,
            "TMROctForest Error: the CUDA createNodes() supports mesh orders 2 "
            "and 3 (order %d, type %d requested)\n",
            order, interp_type);
    return 1;
  }
  /* from order 4 on the sorted keys are entities, not nodes (see
     EntityMultFn) */
  const bool general = order > 3;
  const int gorder = general ? 3 : order; /* geometry: 2x2x2 or 3x3x3 positions */
  const int npe = order * order * order;
  const i64 E = f.n;
  if (E * npe >= (1LL << 31)) {
    fprintf(stderr,
            "TMROctForest Error: %lld elements of order %d overflow the int32 "
            "connectivity of the TMROctForest API\n",
            (long long)E, order);
    return 1;
  }
  if (comm && comm->size > 64) {
    fprintf(stderr, "TMROctForest Error: more than 64 ranks are not supported\n");
    return 1;
  }
  nd.clear();
  nd.order = order;
  nd.interp_type = interp_type;
  for (int i = 0; i < kMaxOrder; i++) nd.knots[i] = (i < order) ? knots[i] : 0.0;
  nd.num_elements = E;
  if (comm) {
    /* the key depth must agree on every rank */
    const int Dg = (int)global_max(ctx, *comm, f.fmt.D);
    if (Dg != f.fmt.D) {
      RekeyFn rk = {f.keys.get(), f.fmt.D, Dg};
      launch(ctx, f.n, rk, "rekey");
      f.fmt.D = Dg;
    }
  }
  nd.nfmt.Dn = f.fmt.D + (order > 2 ? 1 : 0);
  nd.nfmt.bbits = f.bbits;
  const int bernstein = (interp_type == 2) ? 1 : 0;
  /* labels: reference initLabel :6798-6811 */
  nd.nfmt.lbits = (general || (bernstein && order >= 3)) ? 2 : 0;
  if (nd.nfmt.total_bits() > (general && comm ? kSubNodeShift : 64) ||
      nd.nfmt.Dn + 1 > 21) {
    fprintf(stderr,
            "TMROctForest Error: node keys of %d trees at depth %d exceed the "
            "64-bit key budget of the CUDA path\n",
            f.nblocks, nd.nfmt.Dn);
    return 1;
  }
  if (E == 0 && !comm) {
    nd.valid = true;
    nd.node_range.assign(2, 0);
    nd.dep_ptr.alloc(ctx, 1);
    dev_zero(ctx, nd.dep_ptr.get(), sizeof(int));
    return 0;
  }
  trace_mark(ctx, NULL);

  DBuf<u64> own_store, own_store_n;
  OwnerMap om = {NULL, 1}, om_n = {NULL, 1};
  if (comm) {
    om = make_owner_map(f, f.fmt.D, own_store);
    om_n = make_owner_map(f, nd.nfmt.Dn, own_store_n);
  }

  /* 1. hanging faces / edges.  Probes that land in another rank's range are
     collected as queries, answered by the owning rank and patched in (the
     reference pushes a ghost layer instead, computeAdjacentOctants
     :3287-3451; the answers are the same exact-leaf tests) */
  if (!f.info.get()) f.info.alloc(ctx, E);
  DBuf<unsigned char> fmask;
  DBuf<u32> foreign_list; /* elements with a remote coarse neighbour (several ranks) */
  DBuf<unsigned long long> foreign_count;
  DBuf<u32> elem_index_store;
  int ix_bits = 22;
  if (const char *ev = getenv("TMR_B200_IXBITS")) ix_bits = atoi(ev);
  /* exact-leaf searches by key: only the levels the leaf bitmap does not
     cover, and the probes other ranks ask about, need them */
  KeyIndex elem_ix = KeyIndex();
  u64 k_first = 0, k_last = 0;
  {
    DBuf<u32> leaf_bits;
    u64 map_words = 0;
    /* the map spans the trees of this rank's own slice only, so its size per
       GPU stays constant under weak scaling */
    i32 map_b0 = 0, map_nb = f.nblocks;
    if (E > 0) {
      copy_d2h(ctx, &k_first, f.keys.get(), sizeof(u64));
      copy_d2h(ctx, &k_last, f.keys.get() + (E - 1), sizeof(u64));
    }
    if (comm && E > 0) {
      map_b0 = (i32)(k_first >> (3 * f.fmt.D + 5));
      map_nb = (i32)(k_last >> (3 * f.fmt.D + 5)) - map_b0 + 1;
    }
    LeafMap lmap = plan_leaf_map(f.fmt.D, map_b0, map_nb, &map_words);
    if (comm || lmap.lmax < f.fmt.D - 1) {
      elem_ix = build_key_index(ctx, f.keys.get(), E, (u64)f.nblocks << (3 * f.fmt.D + 5),
                                elem_index_store, ix_bits);
    }
    if (lmap.lmax >= 0) {
      leaf_bits.alloc(ctx, (i64)map_words);
      dev_zero(ctx, leaf_bits.get(), (size_t)map_words * sizeof(u32));
      LeafMapBuildFn lb = {f.keys.get(), f.fmt, lmap, leaf_bits.get()};
      launch(ctx, E, lb, "nodes_leaf_map");
      lmap.bits = leaf_bits.get();
    }
    DBuf<int> info32(ctx, comm ? E : 0);
    DBuf<u32> slow_list(ctx, E);
    DBuf<unsigned long long> slow_count(ctx, 1);
    DBuf<u64> fq_key, fq_code;
    DBuf<u32> fq_dest;
    DBuf<unsigned long long> fq_count(ctx, 1);
    i64 cap = comm ? (E / 8 + 4096) : 0;
    i64 nq = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
      if (comm) {
        fq_key.alloc(ctx, cap);
        fq_code.alloc(ctx, cap);
        fq_dest.alloc(ctx, cap);
        dev_zero(ctx, fq_count.get(), sizeof(unsigned long long));
      }
      ElemView ev = {f.keys.get(), E,           f.fmt,         f.tables,
                     elem_ix,      lmap,        om,            me,
                     comm ? 1 : 0,
                     E > 0 ? (i32)(k_first >> (3 * f.fmt.D + 5)) : 0,
                     E > 0 ? (i32)(k_last >> (3 * f.fmt.D + 5)) : 0,
                     fq_key.get(), fq_dest.get(), fq_code.get(), fq_count.get(),
                     cap};
      dev_zero(ctx, slow_count.get(), sizeof(unsigned long long));
      HangingFn hang = {ev, info32.get(), comm ? (int16_t *)NULL : f.info.get(),
                        slow_list.get(), slow_count.get()};
      launch(ctx, E, hang, "nodes_hanging_info");
      /* the launch covers the worst case (every element on a tree face); threads
         beyond the list's length return at once */
      unsigned long long h_slow = 0;
      copy_d2h(ctx, &h_slow, slow_count.get(), sizeof(h_slow));
      HangingBoundaryFn hb = {hang};
      launch(ctx, (i64)h_slow, hb, "nodes_hanging_boundary");
      if (!comm) break;
      unsigned long long h_count = 0;
      copy_d2h(ctx, &h_count, fq_count.get(), sizeof(h_count));
      nq = (i64)h_count;
      if (nq <= cap) break;
      cap = nq; /* rare: the partition surface was larger than the guess */
    }
    if (comm) {
      U32DestFn qdest = {fq_dest.get()};
      RoutePlan plan;
      make_route(ctx, *comm, nq, qdest, plan);
      DBuf<u64> req;
      route_array(ctx, *comm, plan, fq_key.get(), req);
      DBuf<unsigned char> ans(ctx, plan.nrecv);
      AnswerProbeFn ap = {req.get(), f.keys.get(), elem_ix, ans.get()};
      launch(ctx, plan.nrecv, ap, "nodes_probe_answer");
      DBuf<unsigned char> got;
      route_back(ctx, *comm, plan, ans.get(), got);
      foreign_list.alloc(ctx, nq + 1);
      foreign_count.alloc(ctx, 1);
      dev_zero(ctx, foreign_count.get(), sizeof(unsigned long long));
      PatchProbeFn pp = {fq_code.get(), got.get(),           info32.get(),
                         foreign_list.get(), foreign_count.get(), nq};
      launch(ctx, nq, pp, "nodes_probe_patch");
      fmask.alloc(ctx, E);
    }
    if (comm) {
      Info32SplitFn sp = {info32.get(), f.info.get(), fmask.get()};
      launch(ctx, E, sp, "nodes_info_split");
    }
  }
  trace_mark(ctx, "nodes: hanging info");

  /* 2. node candidates -> sort -> unique nodes + local connectivity */
  const i64 nc = E * npe;               /* connectivity entries */
  const i64 ngc = E * (i64)(gorder * gorder * gorder); /* entity slots */
  nd.conn.alloc(ctx, nc);
  /* orders >= 4: the scatter fills the entity table, ConnBuildFn the conn */
  DBuf<int> ent_conn;
  if (general) ent_conn.alloc(ctx, ngc);
  DBuf<u32> ent_off, ent_of; /* entity -> first node, node -> entity */
  i64 Nn = 0;
  DBuf<unsigned char> created;
  /* order 2 without labels: nodes named by (leaf, slot), no candidate sort
     (ops_nodes_slots.h); TMR_B200_NODES=sort forces the general path */
  int slots_done = 0, numbered = 0;
  SlotState slot_state;
  {
    /* the choice must be the same on every rank (the slot construction has
       its own exchanges): nothing rank-local enters the condition */
    const char *mode = getenv("TMR_B200_NODES");
    if (gorder == 2 && !general && nd.nfmt.lbits == 0 && (comm || E > 0) &&
        !(mode && strcmp(mode, "sort") == 0)) {
      slot_state.foreign_list = foreign_list.get();
      slot_state.foreign_count = foreign_count.get();
      slot_state.foreign_cap = foreign_list.size() > 0 ? foreign_list.size() - 1 : 0;
      slots_done = build_nodes_slots(f, nd, fmask.get(), k_first, k_last, om_n, &Nn,
                                     slot_state);
      if (slots_done < 0) return 1;
      numbered = slots_done;
      if (getenv("TMR_B200_NODES_VERBOSE")) {
        fprintf(stderr, "[tmr_b200] createNodes: %s path, %lld elements\n",
                slots_done ? "slot" : "slot->sort fallback", (long long)E);
      }
    }
  }
  if (!slots_done) {
    /* multi-rank: also the parent edge/face nodes of hanging elements */
    i64 nextra = 0;
    DBuf<u32> poff;
    /* multi-rank: dense, order-preserving ids for the trees this rank's node
       keys can name, so that key + payload fit one 64-bit word at any tree
       count */
    DBuf<u32> tree_used, tree_dense, tree_undense;
    int sort_bbits = nd.nfmt.bbits;
    if (comm && E > 0) {
      tree_used.alloc(ctx, f.nblocks);
      tree_dense.alloc(ctx, f.nblocks);
      dev_zero(ctx, tree_used.get(), (size_t)f.nblocks * sizeof(u32));
      MarkTreesFn mk = {f.keys.get(), E, f.fmt, f.tables, tree_used.get()};
      launch(ctx, f.nblocks, mk, "nodes_mark_trees");
      UsedFlagFn uf = {tree_used.get()};
      const i64 nused = (i64)scan_counts(ctx, f.nblocks, uf, tree_dense.get(),
                                         "nodes_dense_trees");
      tree_undense.alloc(ctx, nused);
      UndenseFn un = {tree_used.get(), tree_dense.get(), tree_undense.get()};
      launch(ctx, f.nblocks, un, "nodes_undense_trees");
      sort_bbits = bits_for((int)nused);
    }
    ParentNodeGen pg = {f.keys.get(), fmask.get(), f.fmt,           nd.nfmt,
                        f.tables,     gorder,      tree_dense.get()};
    if (comm) {
      poff.alloc(ctx, E);
      ParentNodeCountFn pc = {pg};
      nextra = (i64)scan_counts(ctx, E, pc, poff.get(), "nodes_parent_count");
    }
    /* candidate emission plan */
    NodeEmit emit_gen = {f.keys.get(), E,        f.fmt, nd.nfmt,
                         f.tables,     gorder,   1,     tree_dense.get(), 1};
    i64 nemit = ngc;
    DBuf<u32> eoff;
    if (emit_gen.families) {
      eoff.alloc(ctx, E);
      NodeEmitCountFn ec = {emit_gen};
      nemit = (i64)scan_counts(ctx, E, ec, eoff.get(), "nodes_emit_count");
    }
    const i64 ntot = nemit + nextra;
    nd.num_candidates = ntot;
    if (ntot >= (1LL << 32) - 1) {
      fprintf(stderr, "TMROctForest Error: too many node candidates\n");
      return 1;
    }
    /* packed mode: when node-key bits + payload bits fit in 64, the payload
       (conn slot) rides in the key's high bits and the sort is keys-only:
       16 B instead of 24 B of HBM traffic per candidate per pass */
    const int mbits = nd.nfmt.pos_bits(); /* Morton + label bits */
    const int nbits = sort_bbits + mbits;
    const u64 max_payload =
        emit_gen.families
            ? (((u64)E << emit_gen.slot_bits()) | ((1ULL << emit_gen.slot_bits()) - 1))
            : (u64)ngc;
    int pbits = 1;
    while ((1ULL << pbits) <= max_payload + 1) pbits++;
    const bool packed = nbits + pbits <= 64;
    if (!packed && max_payload >= 0xffffffffULL) {
      fprintf(stderr, "TMROctForest Error: too many elements for the 32-bit "
                      "node payload\n");
      return 1;
    }
    const u64 kmask = nbits >= 64 ? ~0ULL : ((1ULL << nbits) - 1);
    const u64 no_slot = packed ? ((1ULL << pbits) - 1) : (u64)kNoSlot;
    DBuf<u64> ck(ctx, ntot), ck_alt(ctx, ntot);
    DBuf<u32> cv, cv_alt;
    if (!packed) {
      cv.alloc(ctx, ntot);
      cv_alt.alloc(ctx, ntot);
    }
    if (emit_gen.families && packed && gorder == 2) {
      NodeEmitPackedFn ef = {emit_gen, nbits};
      expand_u64<8>(ctx, E, eoff.get(), (u64)nemit, ef, ck.get(), "nodes_candidates");
    } else if (emit_gen.families && gorder == 2) {
      NodeEmitFillFn<2> ef = {emit_gen, ck.get(), cv.get(), nbits};
      NodeEmitPlaceFn<2> ep = {ef, eoff.get()};
      launch(ctx, E, ep, "nodes_candidates");
    } else if (emit_gen.families) {
      NodeEmitFillFn<3> ef = {emit_gen, ck.get(), cv.get(), nbits};
      NodeEmitPlaceFn<3> ep = {ef, eoff.get()};
      launch(ctx, E, ep, "nodes_candidates");
    } else {
      if (gorder == 2) {
        NodeEmitDenseFn<2> ed = {emit_gen, ck.get(), cv.get(), nbits};
        launch(ctx, E, ed, "nodes_candidates");
      } else {
        NodeEmitDenseFn<3> ed = {emit_gen, ck.get(), cv.get(), nbits};
        launch(ctx, E, ed, "nodes_candidates");
      }
    }
    if (nextra) {
      ParentNodeFillFn pf = {pg, ck.get() + nemit,
                             packed ? (u32 *)0 : cv.get() + nemit,
                             no_slot << nbits};
      ParentPlaceFn pp = {pf, poff.get()};
      launch(ctx, E, pp, "nodes_parent_fill");
    }
    trace_mark(ctx, "nodes: candidates");
    radix_sort(ctx, ck, ck_alt, cv, cv_alt, ntot, 0, nbits, "nodes");
    trace_mark(ctx, "nodes: sort");
    /* the number of unique nodes is not known before the scan: node keys are
       written into the (now free) ping-pong buffer and trimmed afterwards */
    if (comm) {
      created.alloc(ctx, ntot);
      dev_zero(ctx, created.get(), (size_t)ntot);
    }
    RunHeadMaskedFn rh = {ck.get(), kmask};
    if (gorder == 2) {
      NodeScatterFn<2> sc = {ck.get(),     cv.get(),
                             kmask,        nbits,
                             no_slot,      ck_alt.get(),
                             general ? ent_conn.get() : nd.conn.get(),
                             created.get(), emit_gen,
                             tree_undense.get(), mbits};
      Nn = (i64)scan_apply(ctx, ntot, rh, sc, "nodes_unique_scatter_conn");
    } else {
      NodeScatterFn<3> sc = {ck.get(),     cv.get(),
                             kmask,        nbits,
                             no_slot,      ck_alt.get(),
                             general ? ent_conn.get() : nd.conn.get(),
                             created.get(), emit_gen,
                             tree_undense.get(), mbits};
      Nn = (i64)scan_apply(ctx, ntot, rh, sc, "nodes_unique_scatter_conn");
    }

    nd.node_keys.alloc(ctx, Nn);
    copy_d2d(ctx, nd.node_keys.get(), ck_alt.get(), (size_t)Nn * sizeof(u64));
  }
  nd.num_local_nodes = Nn;
  trace_mark(ctx, "nodes: unique+conn");

  /* 2b. node ownership (multi-rank): lowest rank that creates the node from an
     element (reference createLocalNodes :4538-4637).  The home rank of a node
     (owner of its position) decides.  Nodes whose home is this rank stay in
     place; only the few whose home is elsewhere travel, and the home matches
     what it receives against its own node array. */
  DBuf<int> owner;
  if (comm && !numbered) {
    owner.alloc(ctx, Nn);
    NodeHomeFn home = {nd.node_keys.get(), nd.nfmt.Dn, nd.nfmt.lbits, om_n};
    /* foreign-home nodes: (key, created, local index) */
    DBuf<u64> fk(ctx, Nn);
    DBuf<u32> fd(ctx, Nn), fi(ctx, Nn);
    DBuf<unsigned char> fcr(ctx, Nn);
    ForeignNodeCountFn fnc = {home, me};
    ForeignNodeFillFn fnf = {home,     me,       nd.node_keys.get(), created.get(),
                             fk.get(), fd.get(), fi.get(),           fcr.get(),
                             owner.get()};
    const i64 nf = (i64)scan_apply(ctx, Nn, fnc, fnf, "nodes_foreign_home_list");
    U32DestFn fdest = {fd.get()};
    RoutePlan plan;
    make_route(ctx, *comm, nf, fdest, plan);
    DBuf<u64> rk;
    DBuf<unsigned char> rc;
    route_array(ctx, *comm, plan, fk.get(), rk);
    route_array(ctx, *comm, plan, fcr.get(), rc);
    const i64 nr = plan.nrecv;
    DBuf<int> reply(ctx, nr);
    if (nr > 0) {
      DBuf<i64> d_roff(ctx, comm->size + 1);
      copy_h2d(ctx, d_roff.get(), plan.recv_off.data(),
               (size_t)(comm->size + 1) * sizeof(i64));
      DBuf<u32> val(ctx, nr), idx(ctx, nr), idx_alt(ctx, nr);
      DonorValueFn dv = {rc.get(), d_roff.get(), comm->size, val.get(), idx.get()};
      launch(ctx, nr, dv, "nodes_donor_value");
      DBuf<u64> rk_alt(ctx, nr);
      radix_sort(ctx, rk, rk_alt, idx, idx_alt, nr, 0, nd.nfmt.total_bits());
      DBuf<int> owner_run(ctx, nr);
      FillIntFn fi2 = {owner_run.get(), 0x7fffffff};
      launch(ctx, nr, fi2, "nodes_owner_run_init");
      DBuf<u32> run_of(ctx, nr);
      RunHeadFn rh = {rk.get()};
      OwnerMinFn omin = {rk.get(), idx.get(), val.get(), owner_run.get(), run_of.get()};
      scan_apply(ctx, nr, rh, omin, "nodes_owner_min");
      /* fold the received donors into my own nodes, then answer */
      HomeFoldFn hf = {rk.get(), run_of.get(), owner_run.get(), nd.node_keys.get(),
                       Nn, owner.get()};
      launch(ctx, nr, hf, "nodes_owner_fold");
      HomeReplyFn hr = {rk.get(), idx.get(), run_of.get(), owner_run.get(),
                        nd.node_keys.get(), Nn, owner.get(), reply.get()};
      launch(ctx, nr, hr, "nodes_owner_reply");
    }
    DBuf<int> got;
    route_back(ctx, *comm, plan, reply.get(), got);
    StoreOwnerFn so = {fi.get(), got.get(), owner.get()};
    launch(ctx, nf, so, "nodes_owner_store");
    OwnerFinishFn of2 = {owner.get()};
    launch(ctx, Nn, of2, "nodes_owner_finish");
    trace_mark(ctx, "nodes: ownership");
  }

  if (general) {
    /* entities -> nodes: Nn becomes the number of local NODES; node_keys
       keeps one key per entity (stencil look-ups go through ent_off) */
    const i64 nent = Nn;
    ent_off.alloc(ctx, nent);
    EntityMultFn em = {nd.node_keys.get(), order};
    Nn = (i64)scan_counts(ctx, nent, em, ent_off.get(), "nodes_entity_offsets");
    EntityNodes en = {f.tables, order};
    ConnBuildFn cb = {f.keys.get(), f.fmt, en, ent_conn.get(), ent_off.get(),
                      nd.conn.get()};
    launch(ctx, E, cb, "nodes_conn_build");
    ent_conn.reset();
    if (Nn >= (1LL << 31)) {
      fprintf(stderr, "TMROctForest Error: too many local nodes\n");
      return 1;
    }
    if (comm) {
      /* the owner of an entity owns all its nodes (reference :4139-4150) */
      ent_of.alloc(ctx, Nn);
      DBuf<int> owner_node(ctx, Nn);
      ExpandOwnerFn eo = {em, ent_off.get(), owner.get(), ent_of.get(),
                          owner_node.get()};
      launch(ctx, nent, eo, "nodes_entity_owner");
      owner.swap(owner_node);
    }
    nd.num_local_nodes = Nn;
  }

  /* 3. dependent labels and numbering (already done by the one-rank slot
     construction, which numbers in slot space) */
  i64 Nd = nd.num_dep_nodes;
  DBuf<int> dep_node;
  if (!numbered) {
  DBuf<unsigned char> dep_flag(ctx, Nn);
  dev_zero(ctx, dep_flag.get(), (size_t)Nn);
  if (order == 2) {
    DepLabelFn<2> lab = {f.keys.get(), f.info.get(), f.fmt,         order,
                         bernstein,    nd.conn.get(), dep_flag.get()};
    launch(ctx, E, lab, "nodes_dep_label");
  } else if (order == 3) {
    DepLabelFn<3> lab = {f.keys.get(), f.info.get(), f.fmt,         order,
                         bernstein,    nd.conn.get(), dep_flag.get()};
    launch(ctx, E, lab, "nodes_dep_label");
  } else {
    DepLabelFn<0> lab = {f.keys.get(), f.info.get(), f.fmt,         order,
                         bernstein,    nd.conn.get(), dep_flag.get()};
    launch(ctx, E, lab, "nodes_dep_label");
  }
  DBuf<u32> dep_before(ctx, Nn);
  DepFlagFn df = {dep_flag.get()};
  Nd = (i64)scan_counts(ctx, Nn, df, dep_before.get(), "nodes_dep_scan");
  nd.num_dep_nodes = Nd;
  nd.node_num.alloc(ctx, Nn);
  dep_node.alloc(ctx, Nd);
  if (!comm) {
    nd.num_owned_nodes = Nn - Nd;
    nd.node_range_start = 0;
    nd.node_range.assign(2, 0);
    nd.node_range[1] = (int)(Nn - Nd);
    NumberNodesFn num = {dep_flag.get(), dep_before.get(), nd.node_range_start,
                         nd.node_num.get(), dep_node.get()};
    launch(ctx, Nn, num, "nodes_number");
  } else {
    DBuf<u32> owned_before(ctx, Nn);
    OwnedFlagFn of = {dep_flag.get(), owner.get(), me};
    const i64 nown = (i64)scan_counts(ctx, Nn, of, owned_before.get(), "nodes_owned_scan");
    std::vector<i64> all(comm->size);
    comm->allgather_host(ctx, &nown, all.data(), sizeof(i64));
    i64 start = 0;
    for (int r = 0; r < me; r++) start += all[r];
    /* reference node_range (:4165-4172): kept so that the getters stay local */
    nd.node_range.assign(comm->size + 1, 0);
    for (int r = 0; r < comm->size; r++) nd.node_range[r + 1] = nd.node_range[r] + (int)all[r];
    nd.num_owned_nodes = nown;
    nd.node_range_start = (int)start;
    NumberNodesMultiFn num = {dep_flag.get(), dep_before.get(), owned_before.get(),
                              owner.get(),    me,               (int)start,
                              nd.node_num.get(), dep_node.get()};
    /* numbers of the nodes owned elsewhere (reference :4183-4235) */
    DBuf<u64> xk(ctx, Nn);
    DBuf<u32> xd(ctx, Nn), xn(ctx, Nn);
    ExternalCountFn xc = {dep_flag.get(), owner.get(), me};
    ExternalFillFn xf = {xc,       num,      nd.node_keys.get(), xk.get(),
                         xd.get(), xn.get(), ent_of.get(),       ent_off.get()};
    const i64 nx = (i64)scan_apply(ctx, Nn, xc, xf, "nodes_external_list");
    U32DestFn xdest = {xd.get()};
    RoutePlan plan;
    make_route(ctx, *comm, nx, xdest, plan);
    DBuf<u64> req;
    route_array(ctx, *comm, plan, xk.get(), req);
    DBuf<int> rep(ctx, plan.nrecv);
    LookupNumberFn lk = {req.get(),         nd.node_keys.get(),
                         nd.node_keys.size(), nd.node_num.get(),
                         rep.get(),         general ? ent_off.get() : NULL};
    launch(ctx, plan.nrecv, lk, "nodes_external_lookup");
    DBuf<int> got;
    route_back(ctx, *comm, plan, rep.get(), got);
    StoreExternalFn se = {xn.get(), got.get(), nd.node_num.get()};
    launch(ctx, nx, se, "nodes_external_store");
  }
  trace_mark(ctx, "nodes: label+number");

  /* 4. local -> global numbers in the connectivity; the dependent-node CSR
     below reads numbers straight from it */
  ConnRemapInPlaceFn rm = {nd.node_num.get(), nd.conn.get()};
  launch(ctx, nc, rm, "nodes_conn_remap");
  trace_mark(ctx, "nodes: conn remap");
  } /* !numbered */
  /* node_mirror_start needs the sizes; `valid` stays false until the end */
  const int prefetch = nd.prefetch;
  if (prefetch) {
    nd.valid = true;
    if (prefetch & 1) node_mirror_start(f, kMirrorConn);
    if ((prefetch & 2) && !comm) node_mirror_start(f, kMirrorNumbers);
    nd.valid = false;
  }

  /* 5. dependent-node CSR */
  nd.dep_ptr.alloc(ctx, Nd + 1);
  if (Nd > 0) {
    DBuf<u64> win_edge, win_face;
    if (slot_state.winner_done) {
      win_edge.swap(slot_state.win_edge);
      win_face.swap(slot_state.win_face);
    } else {
      win_edge.alloc(ctx, Nd);
      win_face.alloc(ctx, Nd);
      dev_zero(ctx, win_edge.get(), (size_t)Nd * sizeof(u64));
      dev_zero(ctx, win_face.get(), (size_t)Nd * sizeof(u64));
    }
    if (slot_state.winner_done) {
      /* done inside the slot resolve */
    } else if (order == 2) {
      DepWinnerFn<2> win = {f.keys.get(), f.info.get(),   f.fmt,         order,
                            nd.conn.get(), win_edge.get(), win_face.get()};
      launch(ctx, E, win, "nodes_dep_winner");
    } else if (order == 3) {
      DepWinnerFn<3> win = {f.keys.get(), f.info.get(),   f.fmt,         order,
                            nd.conn.get(), win_edge.get(), win_face.get()};
      launch(ctx, E, win, "nodes_dep_winner");
    } else {
      DepWinnerFn<0> win = {f.keys.get(), f.info.get(),   f.fmt,         order,
                            nd.conn.get(), win_edge.get(), win_face.get()};
      launch(ctx, E, win, "nodes_dep_winner");
    }
    /* stencil lengths -> dep_ptr, written by the scan itself */
    DepLenFn len = {win_edge.get(), win_face.get(), order};
    const u64 nnz = scan_counts(ctx, Nd, len, reinterpret_cast<u32 *>(nd.dep_ptr.get()),
                                "nodes_dep_len_scan");
    FillIntFn last = {nd.dep_ptr.get() + Nd, (int)nnz};
    launch(ctx, 1, last, "nodes_dep_ptr");
    nd.dep_nnz = (i64)nnz;
    nd.dep_conn.alloc(ctx, (i64)nnz);
    nd.dep_weights.alloc(ctx, (i64)nnz);
    DBuf<u32> node_index_store;
    DepFillData fill;
    fill.sl = slot_state.lookup;
    fill.slm = slot_state.lookup_m;
    if (!fill.sl.on && !fill.slm.on) {
      fill.node_ix = build_key_index(ctx, nd.node_keys.get(), nd.node_keys.size(),
                                     (u64)f.nblocks << nd.nfmt.pos_bits(),
                                     node_index_store);
    } else {
      fill.node_ix = KeyIndex();
    }
    fill.keys = f.keys.get();
    fill.fmt = f.fmt;
    fill.nfmt = nd.nfmt;
    fill.t = f.tables;
    fill.order = order;
    for (int i = 0; i < kMaxOrder; i++) fill.knots[i] = nd.knots[i];
    fill.node_keys = nd.node_keys.get();
    fill.num_nodes = nd.node_keys.size();
    fill.node_num = nd.node_num.get();
    fill.win_edge = win_edge.get();
    fill.win_face = win_face.get();
    fill.bernstein = bernstein;
    fill.ent_off = general ? ent_off.get() : NULL;
    {
      EntityNodes en = {f.tables, order};
      fill.en = en;
    }
    fill.conn = nd.conn.get();
    {
      NodeEmit fg = {f.keys.get(), E, f.fmt, nd.nfmt, f.tables, order, order == 2 ? 1 : 0,
                     NULL, numbered ? 0 : 1};
      fill.fam = fg;
    }
    fill.dep_ptr = nd.dep_ptr.get();
    fill.dep_conn = nd.dep_conn.get();
    fill.dep_weights = nd.dep_weights.get();
    nd.dep_code.alloc(ctx, Nd);
    fill.dep_code = nd.dep_code.get();
    /* the 1-D weight rows of the stencils (and of the codes' host-side
       expansion), evaluated once */
    nd.dep_wtab.alloc(ctx, (i64)4 * order * order);
    fill.wtab = nd.dep_wtab.get();
    {
      DepWeightTableFn wt = {order, bernstein, {0}, nd.dep_wtab.get()};
      for (int i = 0; i < kMaxOrder; i++) wt.knots[i] = nd.knots[i];
      launch(ctx, (i64)4 * order, wt, "nodes_dep_weight_table");
    }
    if (order == 2) {
      DepFillFn<2> k;
      static_cast<DepFillData &>(k) = fill;
      launch(ctx, Nd, k, "nodes_dep_fill");
    } else if (order == 3) {
      DepFillFn<3> k;
      static_cast<DepFillData &>(k) = fill;
      launch(ctx, Nd, k, "nodes_dep_fill");
    } else {
      DepFillFn<0> k;
      static_cast<DepFillData &>(k) = fill;
      launch(ctx, Nd, k, "nodes_dep_fill");
    }
  } else {
    dev_zero(ctx, nd.dep_ptr.get(), sizeof(int));
    nd.dep_nnz = 0;
  }
  trace_mark(ctx, "nodes: dep CSR");

  nd.valid = true;
  if (prefetch & 4) {
    node_mirror_start(f, kMirrorDepPtr);
    node_mirror_start(f, kMirrorDepConn);
    node_mirror_start(f, kMirrorDepWeights);
  }
  if ((prefetch & 2) && comm) node_mirror_start(f, kMirrorNumbers);
  return check_errors(ctx, "create_nodes");
}

}  // namespace tmrgpu

#endif
