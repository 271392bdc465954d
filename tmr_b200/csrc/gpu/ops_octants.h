/*
  ops_octants.h -- element-array operations: record <-> key conversion,
  createTrees, sort/uniq, refine, coarsen, duplicate.
  Kernel bodies are TMR_HD functors; orchestration is host code issuing
  launch / scan_counts / radix_sort on the context's stream.
*/
#ifndef TMRGPU_OPS_OCTANTS_H
#define TMRGPU_OPS_OCTANTS_H

#include "forest_state.h"

namespace tmrgpu {

/* ---- record <-> key -------------------------------------------------------- */
struct EncodeRecordsFn {
  const Oct24 *recs;
  u64 *keys;
  int16_t *info;
  KeyFmt fmt;
  TMR_HD void operator()(i64 i) const {
    const Oct24 r = recs[i];
    keys[i] = fmt.encode(r.block, r.x, r.y, r.z, r.level);
    if (info) info[i] = r.info;
  }
};

/* key -> 24-byte record with tag = local index (every mutating call of the
   reference ends with tag=i, e.g. reference src/TMROctForest.cpp:2325-2328) */
struct DecodeRecordsFn {
  const u64 *keys;
  const int16_t *info;
  Oct24 *recs;
  KeyFmt fmt;
  TMR_HD void operator()(i64 i) const {
    Oct24 r;
    int level;
    fmt.decode(keys[i], &r.block, &r.x, &r.y, &r.z, &level);
    r.level = (int16_t)level;
    r.info = info ? info[i] : (int16_t)0;
    r.tag = (i32)i;
    recs[i] = r;
  }
};

struct RekeyFn {
  u64 *keys;
  int D_old, D_new;
  TMR_HD void operator()(i64 i) const { keys[i] = rekey(keys[i], D_old, D_new); }
};

/* all 8^level octants of blocks [block_start, block_start+nb) in sorted order
   (reference src/TMROctForest.cpp:1744-1833) */
struct CreateTreesFn {
  u64 *keys;
  int level;
  int block_start;
  TMR_HD void operator()(i64 i) const {
    const int sh = 3 * level;
    const u64 m = (sh > 0) ? ((u64)i & ((1ULL << sh) - 1)) : 0ULL;
    const u64 b = (u64)block_start + ((u64)i >> sh);
    keys[i] = (b << (sh + 5)) | (m << 5) | (u64)level;
  }
};

/* position-strictly-increasing test for an element array */
struct CheckSortedFn {
  const u64 *keys;
  int *flag;
  TMR_HD void operator()(i64 i) const {
    if (i > 0 && (keys[i - 1] >> 5) >= (keys[i] >> 5)) {
      TMR_ATOMIC_OR_I32(flag, 1);
    }
  }
};

/* run-length dedup of a sorted array: keep the LAST entry of every run of
   equal (key >> shift)  (reference src/TMROctant.cpp:373-392: among octants
   sharing an anchor the finest survives; node mode: shift = 0) */
struct RunTailFn {
  const u64 *keys;
  i64 n;
  int shift;
  TMR_HD u32 operator()(i64 i) const {
    return (i == n - 1 || (keys[i] >> shift) != (keys[i + 1] >> shift)) ? 1u
                                                                          : 0u;
  }
};

struct CompactFn {
  const u64 *keys;
  const u32 *vals; /* optional */
  i64 n;
  int shift;
  u64 *out_keys;
  u32 *out_vals;
  TMR_HD void operator()(i64 i, u32 o) const {
    if (i == n - 1 || (keys[i] >> shift) != (keys[i + 1] >> shift)) {
      out_keys[o] = keys[i];
      if (vals) out_vals[o] = vals[i];
    }
  }
};

/* sorted (keys[, vals]) -> unique, in place (buffers swapped); returns count */
inline i64 unique_keep_last(Ctx &ctx, DBuf<u64> &keys, DBuf<u64> &keys_alt,
                            DBuf<u32> &vals, DBuf<u32> &vals_alt, i64 n,
                            int shift) {
  if (n <= 0) return 0;
  RunTailFn tail = {keys.get(), n, shift};
  CompactFn c = {keys.get(), vals.get(), n, shift, keys_alt.get(), vals_alt.get()};
  const i64 m = (i64)scan_apply(ctx, n, tail, c, "unique_compact");
  keys.swap(keys_alt);
  if (vals.get()) vals.swap(vals_alt);
  return m;
}

struct InfoToU32Fn {
  const int16_t *info;
  u32 *out;
  TMR_HD void operator()(i64 i) const { out[i] = (u32)(uint16_t)info[i]; }
};
struct U32ToInfoFn {
  const u32 *in;
  int16_t *info;
  TMR_HD void operator()(i64 i) const { info[i] = (int16_t)(uint16_t)in[i]; }
};

/* Sort + uniq the forest's element array (element mode). */
inline void sort_unique_elements(Forest &f) {
  Ctx &ctx = *f.ctx;
  if (f.n <= 1) return;
  DBuf<u64> kalt(ctx, f.n);
  DBuf<u32> v, valt;
  const bool has_info = f.info.get() != NULL;
  if (has_info) {
    v.alloc(ctx, f.n);
    valt.alloc(ctx, f.n);
    InfoToU32Fn a = {f.info.get(), v.get()};
    launch(ctx, f.n, a, "info_pack");
  }
  radix_sort(ctx, f.keys, kalt, v, valt, f.n, 0, f.fmt.total_bits());
  const i64 m = unique_keep_last(ctx, f.keys, kalt, v, valt, f.n, 5);
  if (has_info) {
    U32ToInfoFn b = {v.get(), f.info.get()};
    launch(ctx, m, b, "info_unpack");
  }
  f.n = m;
}

/* ---- upload / download ------------------------------------------------------ */
inline int upload_octants(Forest &f, const Oct24 *h_recs, i64 n) {
  Ctx &ctx = *f.ctx;
  int D = 0;
  bool any_info = false;
  for (i64 i = 0; i < n; i++) {
    if (h_recs[i].level > D) D = h_recs[i].level;
    if (h_recs[i].info) any_info = true;
  }
  if (!key_budget_ok(f, D)) {
    fprintf(stderr,
            "TMROctForest Error: %d trees at depth %d exceed the 64-bit key "
            "budget of the CUDA path\n",
            f.nblocks, D);
    return 1;
  }
  /* node data is deliberately left alone: writing octants through the
     borrowed array does not invalidate nodes in the reference either.  It was
     built at the key depth all ranks agreed on (create_nodes), which this
     rank's own deepest level may lie below: keep that depth, or the kept slot
     tables and the keys other ranks send would no longer match these keys */
  if (f.nodes.valid && f.fmt.D > D && key_budget_ok(f, f.fmt.D)) D = f.fmt.D;
  f.fmt.D = D;
  f.fmt.bbits = f.bbits;
  f.n = n;
  f.keys.alloc(ctx, n);
  if (any_info) {
    f.info.alloc(ctx, n);
  } else {
    f.info.reset();
  }
  if (n == 0) return 0;
  DBuf<Oct24> d_recs(ctx, n);
  copy_h2d(ctx, d_recs.get(), h_recs, (size_t)n * sizeof(Oct24));
  EncodeRecordsFn e = {d_recs.get(), f.keys.get(), f.info.get(), f.fmt};
  launch(ctx, n, e, "encode_records");
  return check_errors(ctx, "upload_octants");
}

inline int download_octants(Forest &f, Oct24 *h_recs) {
  Ctx &ctx = *f.ctx;
  if (f.n == 0) return 0;
  DBuf<Oct24> d_recs(ctx, f.n);
  DecodeRecordsFn d = {f.keys.get(), f.info.get(), d_recs.get(), f.fmt};
  launch(ctx, f.n, d, "decode_records");
  copy_d2h(ctx, h_recs, d_recs.get(), (size_t)f.n * sizeof(Oct24));
  return check_errors(ctx, "download_octants");
}

struct InfoOnlyFn {
  const int16_t *info;
  int16_t *out;
  TMR_HD void operator()(i64 i) const { out[i] = info ? info[i] : (int16_t)0; }
};

/* ---- createTrees ------------------------------------------------------------ */
inline int create_trees(Forest &f, int level, int block_start, int block_end) {
  Ctx &ctx = *f.ctx;
  if (level < 0) level = 0;
  if (level >= kMaxLevel) level = kMaxLevel - 1;
  if (!key_budget_ok(f, level)) {
    fprintf(stderr,
            "TMROctForest Error: createTrees(%d) on %d trees exceeds the "
            "64-bit key budget of the CUDA path\n",
            level, f.nblocks);
    return 1;
  }
  f.nodes.clear();
  f.info.reset();
  f.fmt.D = level;
  f.fmt.bbits = f.bbits;
  const i64 per = 1LL << (3 * level);
  f.n = per * (i64)(block_end - block_start);
  f.keys.alloc(ctx, f.n);
  CreateTreesFn c = {f.keys.get(), level, block_start};
  launch(ctx, f.n, c, "create_trees");
  return check_errors(ctx, "create_trees");
}

/* ---- refine (reference src/TMROctForest.cpp:2169-2329) ---------------------- */
struct RefinePlan {
  const u64 *keys;
  const int *flags; /* device, may be NULL = +1 everywhere */
  int min_level, max_level;
  /* new level, #outputs, kept-verbatim */
  TMR_HD void plan(i64 i, int *new_level, u32 *count, int *kept) const {
    const int level = (int)(keys[i] & 31);
    const int r = flags ? flags[i] : 1;
    *new_level = level;
    *count = 1;
    *kept = 1;
    if (r > 0 && level < max_level) {
      int nl = level + r;
      if (nl > max_level) nl = max_level;
      const int per_axis_log2 = nl - level - 1; /* (2^(nl-level-1))^3 reps */
      *new_level = nl;
      *kept = 0;
      *count = (per_axis_log2 >= 10) ? 0x40000000u
                                     : (1u << (3 * per_axis_log2));
    } else if (r < 0 && level > min_level) {
      int nl = level + r;
      if (nl < min_level) nl = min_level;
      *new_level = nl;
      *kept = 0;
    }
  }
};

struct RefineCountFn {
  RefinePlan p;
  int *max_new_level; /* device scalar */
  TMR_HD u32 operator()(i64 i) const {
    int nl, kept;
    u32 c;
    p.plan(i, &nl, &c, &kept);
    TMR_ATOMIC_MAX_I32(max_new_level, nl);
    return c;
  }
};

struct RefineFillFn {
  RefinePlan p;
  const int16_t *info_in;
  const u32 *offset;
  KeyFmt fmt_old, fmt_new;
  u64 *out_keys;
  int16_t *out_info; /* may be NULL when info_in is NULL */
  TMR_HD void operator()(i64 i) const {
    int nl, kept;
    u32 c;
    p.plan(i, &nl, &c, &kept);
    const u32 o = offset[i];
    i32 block, x, y, z;
    int level;
    fmt_old.decode(p.keys[i], &block, &x, &y, &z, &level);
    if (kept) {
      out_keys[o] = fmt_new.encode(block, x, y, z, level);
      if (out_info) out_info[o] = info_in ? info_in[i] : (int16_t)0;
      return;
    }
    const i32 h = 1 << (kMaxLevel - nl);
    /* truncate the anchor to the new level (only changes it when coarsening) */
    x &= ~(h - 1);
    y &= ~(h - 1);
    z &= ~(h - 1);
    for (u32 t = 0; t < c; t++) {
      u32 ii, jj, kk;
      unmorton3((u64)t, &ii, &jj, &kk);
      out_keys[o + t] = fmt_new.encode(block, x + 2 * (i32)ii * h,
                                       y + 2 * (i32)jj * h,
                                       z + 2 * (i32)kk * h, nl);
      if (out_info) out_info[o + t] = 0;
    }
  }
};

inline int refine(Forest &f, const int *d_flags, int min_level, int max_level) {
  Ctx &ctx = *f.ctx;
  if (min_level < 0) min_level = 0;
  if (max_level > kMaxLevel) max_level = kMaxLevel;
  if (min_level > max_level) min_level = max_level;
  f.nodes.clear();
  f.last_in = f.n;
  if (f.n == 0) return 0;

  trace_mark(ctx, NULL);
  DBuf<int> scalars(ctx, 2);
  dev_zero(ctx, scalars.get(), 2 * sizeof(int));
  DBuf<u32> offset(ctx, f.n);
  RefinePlan plan = {f.keys.get(), d_flags, min_level, max_level};
  RefineCountFn cnt = {plan, scalars.get()};
  const u64 total = scan_counts(ctx, f.n, cnt, offset.get(), "refine_count");
  int h_scalars[2];
  copy_d2h(ctx, h_scalars, scalars.get(), 2 * sizeof(int));
  const int D_new = h_scalars[0];
  if (total >= (1ULL << 31)) {
    fprintf(stderr,
            "TMROctForest Error: refine() would create %llu octants on one "
            "rank (int32 index limit of the TMROctForest API)\n",
            (unsigned long long)total);
    return 1;
  }
  if (!key_budget_ok(f, D_new)) {
    fprintf(stderr,
            "TMROctForest Error: refine() to depth %d on %d trees exceeds the "
            "64-bit key budget of the CUDA path\n",
            D_new, f.nblocks);
    return 1;
  }
  KeyFmt fmt_new = f.fmt;
  fmt_new.D = D_new;
  DBuf<u64> out(ctx, (i64)total);
  DBuf<int16_t> out_info;
  if (f.info.get()) out_info.alloc(ctx, (i64)total);
  RefineFillFn fill = {plan,    f.info.get(), offset.get(), f.fmt,
                       fmt_new, out.get(),    out_info.get()};
  launch(ctx, f.n, fill, "refine_fill");
  f.keys.swap(out);
  f.info.swap(out_info);
  f.n = (i64)total;
  f.fmt = fmt_new;

  /* outputs of a valid leaf array with non-negative flags are already in
     Morton order; anything else (coarsening, overlapping input) is sorted */
  CheckSortedFn chk = {f.keys.get(), scalars.get() + 1};
  launch(ctx, f.n, chk, "check_sorted");
  copy_d2h(ctx, h_scalars, scalars.get(), 2 * sizeof(int));
  if (h_scalars[1]) sort_unique_elements(f);
  f.last_mid = f.n;
  trace_mark(ctx, "refine");
  return check_errors(ctx, "refine");
}

/* ---- coarsen (reference src/TMROctForest.cpp:2119-2164) --------------------- */
struct CoarsenCountFn {
  const u64 *keys;
  KeyFmt fmt;
  TMR_HD u32 operator()(i64 i) const {
    const u64 k = keys[i];
    const int level = (int)(k & 31);
    if (level == 0) return 1;
    /* child id 0 <=> the octal digit of the anchor at `level` is zero */
    const u64 digit = (k >> (5 + 3 * (fmt.D - level))) & 7;
    return digit == 0 ? 1u : 0u;
  }
};

struct CoarsenFillFn {
  const u64 *keys;
  const int16_t *info;
  const u32 *offset;
  KeyFmt fmt;
  u64 *out_keys;
  int16_t *out_info;
  TMR_HD void operator()(i64 i) const {
    const u64 k = keys[i];
    const int level = (int)(k & 31);
    if (level == 0) {
      out_keys[offset[i]] = k;
      if (out_info) out_info[offset[i]] = info ? info[i] : (int16_t)0;
      return;
    }
    const u64 digit = (k >> (5 + 3 * (fmt.D - level))) & 7;
    if (digit == 0) {
      out_keys[offset[i]] = (k & ~31ULL) | (u64)(level - 1);
      if (out_info) out_info[offset[i]] = 0;
    }
  }
};

inline int coarsen_into(const Forest &src, Forest &dst) {
  Ctx &ctx = *src.ctx;
  dst.nodes.clear();
  dst.fmt = src.fmt;
  dst.n = 0;
  dst.keys.reset();
  dst.info.reset();
  if (src.n == 0) return 0;
  DBuf<u32> offset(ctx, src.n);
  CoarsenCountFn cnt = {src.keys.get(), src.fmt};
  const i64 m = (i64)scan_counts(ctx, src.n, cnt, offset.get(), "coarsen_count");
  dst.keys.alloc(ctx, m);
  if (src.info.get()) dst.info.alloc(ctx, m);
  CoarsenFillFn fill = {src.keys.get(), src.info.get(), offset.get(),
                        src.fmt,        dst.keys.get(), dst.info.get()};
  launch(ctx, src.n, fill, "coarsen_fill");
  dst.n = m;
  return check_errors(ctx, "coarsen");
}

inline int duplicate_into(const Forest &src, Forest &dst) {
  Ctx &ctx = *src.ctx;
  dst.nodes.clear();
  dst.fmt = src.fmt;
  dst.n = src.n;
  dst.keys.alloc(ctx, src.n);
  copy_d2d(ctx, dst.keys.get(), src.keys.get(), (size_t)src.n * sizeof(u64));
  if (src.info.get()) {
    dst.info.alloc(ctx, src.n);
    copy_d2d(ctx, dst.info.get(), src.info.get(),
             (size_t)src.n * sizeof(int16_t));
  } else {
    dst.info.reset();
  }
  return check_errors(ctx, "duplicate");
}

/* ---- synthetic refinement flags + checksum (bench / tests) ------------------ */
struct SynthFlagsFn {
  const u64 *keys;
  KeyFmt fmt;
  u64 seed;
  int pct;
  int *flags;
  TMR_HD void operator()(i64 i) const {
    i32 b, x, y, z;
    int level;
    fmt.decode(keys[i], &b, &x, &y, &z, &level);
    flags[i] = (record_hash(seed, b, x, y, z, level) % 100ULL) < (u64)pct ? 1 : 0;
  }
};

/* order-independent checksum: sum of record hashes mod 2^64, accumulated into
   kChecksumSlots partial sums (spreads the atomics over many L2 lines) */
static const int kChecksumSlots = 4096;
struct ChecksumFn {
  const u64 *keys;
  KeyFmt fmt;
  u64 *sums; /* [kChecksumSlots] */
  TMR_HD void operator()(i64 i) const {
    i32 b, x, y, z;
    int level;
    fmt.decode(keys[i], &b, &x, &y, &z, &level);
    TMR_ATOMIC_ADD_U64(&sums[(i >> 5) & (kChecksumSlots - 1)],
                       record_hash(0, b, x, y, z, level));
  }
};

inline u64 checksum(Forest &f) {
  Ctx &ctx = *f.ctx;
  DBuf<u64> sums(ctx, kChecksumSlots);
  dev_zero(ctx, sums.get(), kChecksumSlots * sizeof(u64));
  ChecksumFn c = {f.keys.get(), f.fmt, sums.get()};
  launch(ctx, f.n, c, "checksum");
  std::vector<u64> h(kChecksumSlots);
  copy_d2h(ctx, h.data(), sums.get(), kChecksumSlots * sizeof(u64));
  u64 s = 0;
  for (int i = 0; i < kChecksumSlots; i++) s += h[i];
  return s;
}

}  // namespace tmrgpu

#endif
