/*
  ops_nodes_slots.h -- sort-free construction of the unique node set and the
  local connectivity of an order-2 mesh (replaces the candidate sort of
  ops_nodes.h for the common case; same results, reference createLocalNodes
  src/TMROctForest.cpp:4290-4640 and createLocalConn :4660-4867).

  Every canonical node position (after transformNode and its hmax -> hmax-1
  clamp, reference :3847-4039) lies in exactly one leaf of the Morton-sorted
  element array: the LAST element whose anchor is at or before the position.
  In a 2:1 balanced, complete forest the position relative to that leaf is,
  per axis, one of three things:
      0   the leaf's own anchor coordinate,
      1   half the leaf's size above it (a hanging node on a lower face / edge
          of the leaf, created by finer neighbours),
      2   the clamped coordinate 2^30-1 (the leaf touches the upper face of
          its tree and the node sits on that face).
  So a node is named by (leaf, slot) with 27 possible slots, the unique node
  array in the reference's compareNode order (src/TMROctant.cpp:245-274) is
  "for each leaf in order, its occupied slots in slot order", and the index of
  a node is an exclusive scan of popc(slot mask) -- no sort, no candidates in
  HBM.  Slot order: the squeezed node coordinate (NodeFmt) of slot index 0 / 1
  / 2 on a leaf of 2^k cells is A<<1, A<<1 | 1<<k, A<<1 | (2<<k)-1, so node
  keys inside a leaf compare by the 6-bit code (hx hy hz lx ly lz), h = index
  >= 1, l = index == 2, x before y before z; 27 of the 64 codes are valid and
  a slot's ordinal among them indexes a 32-bit mask.

  The connectivity needs, for each of the 8 corners of each element, the leaf
  containing the corner: one predecessor search in the L2-resident radix index
  of the element keys.  A complete family of 8 siblings shares the 27 points
  of its 3x3x3 grid; 8 of them are the siblings' own anchors, the other 19
  are located once and handed to the siblings through shared memory
  (launch_block2), 2.4 searches per element instead of 7.

  Anything the slots cannot name -- a corner at a quarter position of its
  leaf (unbalanced input), a position not covered by any leaf (incomplete
  forest after a bare refine) -- raises a flag and create_nodes falls back to
  the general candidate sort.  Positions outside this rank's own Morton range
  (several ranks) go through the candidate sort as a small side list ("B"
  nodes): they sort entirely before or after the range's own nodes.
*/
#ifndef TMRGPU_OPS_NODES_SLOTS_H
#define TMRGPU_OPS_NODES_SLOTS_H

#include "ops_route.h"

namespace tmrgpu {

static const u64 kSlotValid = 0xff5533110f050301ULL; /* codes with l <= h */
static const u64 kLocB = ~0ULL;        /* outside this rank's Morton range */
static const u64 kLocFail = ~0ULL - 1; /* not nameable: general path */
static const u32 kConnB = 0xffffffffu; /* conn entry filled by the B pass */

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

TMR_HD int slot_ord(int c6) { return popc64(kSlotValid & ((1ULL << c6) - 1)); }
TMR_HD int slot_code(int ord) {
  u64 v = kSlotValid;
  for (int r = 0; r < ord; r++) v &= v - 1;
  return ctz64(v);
}

/* add one at bit `bit` of the dilated axis-a component of a Morton code */
TMR_HD u64 morton_axis_add(u64 m, int a, int bit) {
  const u64 am = 0x1249249249249249ULL << a;
  const u64 nc = (((m & am) | ~am) + (1ULL << bit)) & am;
  return (m & ~am) | nc;
}

TMR_HD void store8_u32(u32 *p, const u32 *v) {
#if defined(__CUDA_ARCH__)
  uint4 *q = reinterpret_cast<uint4 *>(p);
  q[0] = make_uint4(v[0], v[1], v[2], v[3]);
  q[1] = make_uint4(v[4], v[5], v[6], v[7]);
#else
  for (int c = 0; c < 8; c++) p[c] = v[c];
#endif
}
TMR_HD void load8_u32(const u32 *p, u32 *v) {
#if defined(__CUDA_ARCH__)
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  const uint4 a = q[0], b = q[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#else
  for (int c = 0; c < 8; c++) v[c] = p[c];
#endif
}

struct SlotView {
  const u64 *keys;
  i64 E;
  KeyFmt fmt;
  ConnTables t;
  KeyIndex ix;
  /* several ranks: positions (depth D) of this rank's own range */
  int multi;
  u64 pos_lo, pos_hi;
  u32 *mask;  /* per leaf: occupied slots */
  u32 *cmask; /* several ranks: slots that are a corner of a local element */
  int *fail;

  /* leaf and slot of the canonical position (block, Morton m of the depth-D
     cell holding it, clamp bit a set where the coordinate is 2^30-1; a = 2 x,
     1 y, 0 z): (leaf << 5) | ordinal, kLocB or kLocFail */
  TMR_HD u64 locate(i32 block, u64 m, int clamp) const {
    const int D = fmt.D;
    const u64 pos = ((u64)(u32)block << (3 * D)) | m;
    if (multi && (pos < pos_lo || pos >= pos_hi)) return kLocB;
    const i64 j = ix.pred(keys, (pos << 5) | 31ULL);
    if (j < 0) return kLocFail;
    const u64 kl = keys[j];
    const int k = D - (int)(kl & 31);
    if (k < 0 || ((kl >> 5) >> (3 * k)) != (pos >> (3 * k))) return kLocFail;
    const u64 lm = k > 0 ? ((1ULL << (3 * k)) - 1) : 0ULL;
    const u64 d = pos & lm;
    int c6 = 0;
    TMR_UNROLL
    for (int a = 0; a < 3; a++) {
      const u64 am = (0x1249249249249249ULL << a) & lm;
      const u64 da = d & am;
      if ((clamp >> a) & 1) {
        if (da != am) return kLocFail;
        c6 |= (8 | 1) << a;
      } else if (da != 0) {
        if (da != (1ULL << (3 * (k - 1) + a))) return kLocFail;
        c6 |= 8 << a;
      }
    }
    return ((u64)j << 5) | (u64)slot_ord(c6);
  }
  /* canonical coordinates (after transform_node) */
  TMR_HD u64 locate_xyz(i32 block, i32 x, i32 y, i32 z) const {
    const int s = kMaxLevel - fmt.D;
    const int clamp = ((x == kHmax - 1) ? 4 : 0) | ((y == kHmax - 1) ? 2 : 0) |
                      ((z == kHmax - 1) ? 1 : 0);
    return locate(block, morton3((u32)x >> s, (u32)y >> s, (u32)z >> s), clamp);
  }
  TMR_HD void mark(u64 v, bool corner) const {
    if (v >= kLocFail) return;
    const i64 leaf = (i64)(v >> 5);
    const u32 bit = 1u << (int)(v & 31);
    if (!(mask[leaf] & bit)) TMR_ATOMIC_OR_I32(&mask[leaf], bit);
    if (cmask && corner && !(cmask[leaf] & bit)) TMR_ATOMIC_OR_I32(&cmask[leaf], bit);
  }
};

/* per element: (leaf, slot) of its 8 corners */
struct NodeSlotFn {
  SlotView v;
  u32 *conn_leaf;       /* [E][8] leaf index of every corner (kConnB: B list) */
  unsigned char *slot8; /* [E][8] slot ordinal */
  /* several ranks: corners whose position is not in this rank's range */
  NodeFmt nfmt;
  u64 *b_key;
  u32 *b_pay;
  unsigned long long *b_count;
  i64 b_cap;

  struct Shared {
    u64 val[kLaunchThreads * 4];
    short fam0[kLaunchThreads];
  };

  TMR_HD void append_b(u64 key, u32 payload) const {
    const unsigned long long s = fetch_add_u64(b_count, 1ULL);
    if ((i64)s < b_cap) {
      b_key[s] = key;
      b_pay[s] = payload;
    }
  }

  TMR_HD void stage(i64 i, int t, Shared &sh) const {
    const u64 key = v.keys[i];
    const int L = (int)(key & 31), D = v.fmt.D;
    const int s = 3 * (D - L);
    const u64 m = D > 0 ? ((key >> 5) & ((1ULL << (3 * D)) - 1)) : 0ULL;
    const i32 block = (i32)(key >> (3 * D + 5));
    sh.fam0[t] = -1;
    if (L == 0) return;
    /* the element's own anchor is a node in place unless it lies on a lower
       tree face (then it goes through transform_node like any other corner) */
    {
      const u64 lv = (1ULL << (3 * L)) - 1;
      const u64 mc = m >> s;
      if ((mc & 0x1249249249249249ULL & lv) && (mc & 0x2492492492492492ULL & lv) &&
          (mc & 0x4924924924924924ULL & lv)) {
        v.mark((u64)i << 5, true);
      }
    }
    const int md = (int)((m >> s) & 7);
    const i64 e0 = i - md, i0 = i - t;
    if (e0 < i0 || e0 + 7 >= i0 + kLaunchThreads || e0 + 7 >= v.E) return;
    const u64 k0 = v.keys[e0];
    if ((k0 & 31) != (u64)L || ((k0 >> (5 + s)) & 7) != 0 ||
        v.keys[e0 + 7] != k0 + (7ULL << (5 + s))) {
      return;
    }
    /* 8 consecutive keys from sibling 0 to sibling 7: in a valid leaf set
       they are the 8 siblings, and this element is number md of them */
    if (key != k0 + ((u64)md << (5 + s))) {
      *v.fail = 1;
      return;
    }
    /* interior family: the parent's cell touches no tree face */
    const int Lp = L - 1;
    const u64 lvp = Lp > 0 ? ((1ULL << (3 * Lp)) - 1) : 0ULL;
    const u64 mp = m >> (s + 3);
    TMR_UNROLL
    for (int a = 0; a < 3; a++) {
      const u64 am = (0x1249249249249249ULL << a) & lvp;
      const u64 c = mp & am;
      if (c == 0 || c == am) return;
    }
    sh.fam0[t] = (short)(e0 - i0);
    const u64 mP = (m >> (s + 3)) << (s + 3);
    TMR_UNROLL
    for (int r = 0; r < 4; r++) {
      const int tp = md + 8 * r;
      if (tp >= 27) break;
      const int g[3] = {tp / 9, (tp / 3) % 3, tp % 3}; /* axis a: 0 z, 1 y, 2 x */
      u64 val;
      if (g[0] < 2 && g[1] < 2 && g[2] < 2) {
        val = (u64)(e0 + 4 * g[2] + 2 * g[1] + g[0]) << 5;
      } else {
        u64 mq = mP;
        TMR_UNROLL
        for (int a = 0; a < 3; a++) {
          if (g[a] == 1) mq |= 1ULL << (s + a);
          if (g[a] == 2) mq = morton_axis_add(mq, a, s + 3 + a);
        }
        val = v.locate(block, mq, 0);
        v.mark(val, true);
      }
      sh.val[t * 4 + r] = val;
    }
  }

  TMR_HD void finish(i64 i, int t, Shared &sh) const {
    const u64 key = v.keys[i];
    const int L = (int)(key & 31), D = v.fmt.D;
    const int s = 3 * (D - L);
    const u64 m = D > 0 ? ((key >> 5) & ((1ULL << (3 * D)) - 1)) : 0ULL;
    const i32 block = (i32)(key >> (3 * D + 5));
    u64 val[8];
    const int f0 = sh.fam0[t];
    if (f0 >= 0) {
      const int md = (int)((m >> s) & 7);
      const int gz0 = md & 1, gy0 = (md >> 1) & 1, gx0 = (md >> 2) & 1;
      TMR_UNROLL
      for (int c = 0; c < 8; c++) {
        const int gx = gx0 + (c & 1), gy = gy0 + ((c >> 1) & 1), gz = gz0 + (c >> 2);
        const int tp = 9 * gz + 3 * gy + gx;
        val[c] = sh.val[(f0 + (tp & 7)) * 4 + (tp >> 3)];
      }
    } else {
      /* own corners.  Interior element: Morton space; else coordinates and
         transform_node */
      bool interior = L > 0;
      if (interior) {
        const u64 lv = (1ULL << (3 * L)) - 1;
        const u64 mc = m >> s;
        TMR_UNROLL
        for (int a = 0; a < 3; a++) {
          const u64 am = (0x1249249249249249ULL << a) & lv;
          const u64 c = mc & am;
          if (c == 0 || c == am) interior = false;
        }
      }
      if (interior) {
        TMR_UNROLL
        for (int c = 0; c < 8; c++) {
          if (c == 0) {
            val[c] = (u64)i << 5;
            continue;
          }
          u64 mq = m;
          if (c & 1) mq = morton_axis_add(mq, 2, s + 2);
          if (c & 2) mq = morton_axis_add(mq, 1, s + 1);
          if (c & 4) mq = morton_axis_add(mq, 0, s);
          val[c] = v.locate(block, mq, 0);
          v.mark(val[c], true);
        }
      } else {
        i32 b0, x, y, z;
        int lv;
        v.fmt.decode(key, &b0, &x, &y, &z, &lv);
        const i32 h = 1 << (kMaxLevel - L);
        for (int c = 0; c < 8; c++) {
          i32 b = b0, nx = x + (c & 1) * h, ny = y + ((c >> 1) & 1) * h,
              nz = z + (c >> 2) * h;
          transform_node(v.t, &b, &nx, &ny, &nz, -1, NULL, NULL);
          val[c] = v.locate_xyz(b, nx, ny, nz);
          v.mark(val[c], true);
        }
      }
    }
    u32 leaf[8];
    u64 ords = 0;
    TMR_UNROLL
    for (int c = 0; c < 8; c++) {
      if (val[c] == kLocFail) {
        *v.fail = 1;
        leaf[c] = 0;
      } else if (val[c] == kLocB) {
        i32 b, x, y, z;
        int lv;
        v.fmt.decode(key, &b, &x, &y, &z, &lv);
        const i32 h = 1 << (kMaxLevel - L);
        x += (c & 1) * h;
        y += ((c >> 1) & 1) * h;
        z += (c >> 2) * h;
        transform_node(v.t, &b, &x, &y, &z, -1, NULL, NULL);
        append_b(nfmt.encode(b, x, y, z, 0), (u32)(i * 8 + c));
        leaf[c] = kConnB;
      } else {
        leaf[c] = (u32)(val[c] >> 5);
        ords |= (val[c] & 31) << (8 * c);
      }
    }
    store8_u32(conn_leaf + i * 8, leaf);
    *reinterpret_cast<u64 *>(slot8 + i * 8) = ords;
  }
};

/* several ranks: the parent edge / face nodes of hanging elements whose coarse
   neighbour is remote (ParentNodeGen) must exist as nodes too */
struct SlotKeyEmit {
  const SlotView *v;
  const NodeFmt *nfmt;
  const NodeSlotFn *ns;
  TMR_HD void operator()(i32 b, i32 x, i32 y, i32 z, int) const {
    const u64 val = v->locate_xyz(b, x, y, z);
    if (val == kLocFail) {
      *v->fail = 1;
    } else if (val == kLocB) {
      ns->append_b(nfmt->encode(b, x, y, z, 0), kConnB);
    } else {
      v->mark(val, false);
    }
  }
};

struct SlotCountFn {
  const u32 *mask;
  TMR_HD u32 operator()(i64 i) const { return (u32)popc32(mask[i]); }
};
struct SlotInfoFn {
  const u32 *mask;
  u64 *slotinfo;
  TMR_HD void operator()(i64 i, u32 off) const {
    slotinfo[i] = ((u64)off << 32) | (u64)mask[i];
  }
};

/* node keys (NodeFmt at Dn = D) of every occupied slot, in order */
struct SlotKeysFn {
  const u64 *keys;
  KeyFmt fmt;
  const u64 *slotinfo;
  const u32 *cmask;
  u64 *node_keys;
  unsigned char *created;
  i64 base; /* index of the first slot node */
  TMR_HD void operator()(i64 i) const {
    const u64 si = slotinfo[i];
    u32 mk = (u32)si;
    if (!mk) return;
    const u64 key = keys[i];
    const int D = fmt.D;
    const int k = D - (int)(key & 31);
    const u64 m = D > 0 ? ((key >> 5) & ((1ULL << (3 * D)) - 1)) : 0ULL;
    const u64 block = key >> (3 * D + 5);
    const u64 nb = (block << (3 * (D + 1))) | (m << 3);
    const u64 lm = (1ULL << (3 * (k + 1))) - 1;
    i64 o = base + (i64)(si >> 32);
    const u32 cm = cmask ? cmask[i] : 0u;
    while (mk) {
      const int ord = ctz32(mk);
      mk &= mk - 1;
      const int c6 = slot_code(ord);
      u64 extra = 0;
      TMR_UNROLL
      for (int a = 0; a < 3; a++) {
        if ((c6 >> a) & 1) {
          extra |= (0x1249249249249249ULL << a) & lm;
        } else if ((c6 >> (3 + a)) & 1) {
          extra |= 1ULL << (3 * k + a);
        }
      }
      node_keys[o] = nb | extra;
      if (created) created[o] = (unsigned char)((cm >> ord) & 1u);
      o++;
    }
  }
};

/* (leaf, slot) -> local node index, in place */
struct SlotResolveFn {
  const u64 *slotinfo;
  const unsigned char *slot8;
  u32 *conn;
  u32 base;
  TMR_HD void operator()(i64 e) const {
    u32 leaf[8];
    load8_u32(conn + e * 8, leaf);
    const u64 ords = *reinterpret_cast<const u64 *>(slot8 + e * 8);
    TMR_UNROLL
    for (int c = 0; c < 8; c++) {
      if (leaf[c] == kConnB) continue;
      const u64 si = slotinfo[leaf[c]];
      const int ord = (int)((ords >> (8 * c)) & 31);
      leaf[c] = base + (u32)(si >> 32) + (u32)popc32((u32)si & ((1u << ord) - 1u));
    }
    store8_u32(conn + e * 8, leaf);
  }
};

}  // namespace tmrgpu

#endif
