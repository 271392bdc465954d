/* tests/emu/prim_emu.cpp -- TEST INFRASTRUCTURE ONLY (see prim_emu.h). */
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>

#include "prim.h"

namespace tmrgpu {

void *dev_alloc(Ctx &, size_t bytes) { return malloc(bytes ? bytes : 16); }
void dev_free(Ctx &, void *p) { free(p); }
void dev_cache_destroy(Ctx &) {}
void *host_alloc(Ctx &, size_t bytes) { return malloc(bytes ? bytes : 16); }
void host_free(Ctx &, void *p) { free(p); }
void copy_h2d(Ctx &, void *d, const void *s, size_t b) { memcpy(d, s, b); }
void copy_d2h(Ctx &, void *d, const void *s, size_t b) { memcpy(d, s, b); }
void *copy_d2h_async(Ctx &, void *d, const void *s, size_t b, int) {
  memcpy(d, s, b);
  return NULL;
}
void copy_wait(Ctx &, void *) {}
void copy_sync(Ctx &, void *) {}
void copy_d2d(Ctx &, void *d, const void *s, size_t b) { memcpy(d, s, b); }
void dev_zero(Ctx &, void *p, size_t b) { memset(p, 0, b); }
void dev_fill_ff(Ctx &, void *p, size_t b) { memset(p, 0xff, b); }
void stream_sync(Ctx &) {}
int check_errors(Ctx &ctx, const char *) { return ctx.last_error.empty() ? 0 : 1; }
void prof_begin(Ctx &, const char *) {}
void prof_end(Ctx &) {}
void prof_resolve(Ctx &) {}
void trace_mark(Ctx &, const char *) {}

void radix_sort(Ctx &ctx, DBuf<u64> &keys, DBuf<u64> &, DBuf<u32> &vals,
                DBuf<u32> &, i64 n, int bit_lo, int bit_hi, const char *) {
  if (n <= 1 || bit_hi <= bit_lo) return;
  const u64 mask = (bit_hi - bit_lo >= 64) ? ~0ULL
                                           : (((1ULL << (bit_hi - bit_lo)) - 1) << bit_lo);
  std::vector<i64> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  const u64 *k = keys.get();
  std::stable_sort(idx.begin(), idx.end(), [k, mask](i64 a, i64 b) {
    return (k[a] & mask) < (k[b] & mask);
  });
  std::vector<u64> tk(n);
  for (i64 i = 0; i < n; i++) tk[i] = k[idx[i]];
  memcpy(keys.get(), tk.data(), n * sizeof(u64));
  if (vals.get()) {
    std::vector<u32> tv(n);
    for (i64 i = 0; i < n; i++) tv[i] = vals.get()[idx[i]];
    memcpy(vals.get(), tv.data(), n * sizeof(u32));
  }
  ctx.launch_count++;
}

}  // namespace tmrgpu
