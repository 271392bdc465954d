/*
  TMROctant.cpp -- host side of the octant primitives for the B200 drop-in.
  Scalar helpers follow the semantics of reference src/TMROctant.cpp:28-291;
  array sort/search are forwarded to the CUDA layer.
*/
#include "TMROctant.h"

#include <stdio.h>

#include <deque>
#include <unordered_set>
#include <vector>

#include "tmrgpu.h"

/* ---- process-wide context --------------------------------------------------- */
static thread_local tmrgpu_ctx *g_ctx = NULL;
static thread_local void *g_stream = NULL;
static thread_local int g_ctx_failed = 0;

extern "C" void tmr_b200_use_stream(void *stream) { g_stream = stream; }

extern "C" tmrgpu_ctx *tmr_b200_context(void) {
  if (!g_ctx && !g_ctx_failed) {
    int device = 0;
    const char *dev = getenv("TMR_B200_DEVICE");
    if (!dev) dev = getenv("LOCAL_RANK");
    if (dev) device = atoi(dev);
    if (tmrgpu_ctx_create(device, g_stream, &g_ctx) != 0) {
      g_ctx = NULL;
      g_ctx_failed = 1;
      fprintf(stderr,
              "TMROctForest Error: no usable CUDA device; the B200 drop-in has "
              "no CPU fallback\n");
    }
  }
  return g_ctx;
}

/* ---- TMROctant ---------------------------------------------------------------- */
static inline int32_t side_length(int level) {
  return 1 << (TMR_MAX_LEVEL - level);
}

int TMROctant::childId() {
  const int32_t h = side_length(level);
  return ((x & h) ? 1 : 0) | ((y & h) ? 2 : 0) | ((z & h) ? 4 : 0);
}

void TMROctant::getSibling(int id, TMROctant *sib) {
  const int32_t h = side_length(level);
  const int32_t x0 = (x & h) ? x - h : x;
  const int32_t y0 = (y & h) ? y - h : y;
  const int32_t z0 = (z & h) ? z - h : z;
  sib->block = block;
  sib->level = level;
  sib->info = 0;
  sib->x = x0 + ((id & 1) ? h : 0);
  sib->y = y0 + ((id & 2) ? h : 0);
  sib->z = z0 + ((id & 4) ? h : 0);
}

void TMROctant::parent(TMROctant *p) {
  p->block = block;
  p->info = 0;
  if (level > 0) {
    const int32_t h = side_length(level);
    p->level = level - 1;
    p->x = x & ~h;
    p->y = y & ~h;
    p->z = z & ~h;
  } else {
    p->level = 0;
    p->x = x;
    p->y = y;
    p->z = z;
  }
}

void TMROctant::faceNeighbor(int face, TMROctant *nb) {
  const int32_t h = side_length(level);
  const int32_t step = (face & 1) ? h : -h;
  nb->block = block;
  nb->level = level;
  nb->info = 0;
  nb->x = x + ((face >> 1) == 0 ? step : 0);
  nb->y = y + ((face >> 1) == 1 ? step : 0);
  nb->z = z + ((face >> 1) == 2 ? step : 0);
}

void TMROctant::edgeNeighbor(int edge, TMROctant *nb) {
  const int32_t h = side_length(level);
  const int s = edge & 3;
  const int32_t a = (s & 1) ? h : -h;
  const int32_t b = (s >> 1) ? h : -h;
  nb->block = block;
  nb->level = level;
  nb->info = 0;
  if (edge < 4) {
    nb->x = x;
    nb->y = y + a;
    nb->z = z + b;
  } else if (edge < 8) {
    nb->x = x + a;
    nb->y = y;
    nb->z = z + b;
  } else {
    nb->x = x + a;
    nb->y = y + b;
    nb->z = z;
  }
}

void TMROctant::cornerNeighbor(int corner, TMROctant *nb) {
  const int32_t h = side_length(level);
  nb->block = block;
  nb->level = level;
  nb->info = 0;
  nb->x = x + ((corner & 1) ? h : -h);
  nb->y = y + ((corner & 2) ? h : -h);
  nb->z = z + ((corner & 4) ? h : -h);
}

/* -1/0/+1 ordering of two positions on the x-major Morton curve */
static inline int morton_order(const TMROctant *a, const TMROctant *b) {
  if (a->block != b->block) return a->block - b->block;
  const uint32_t dx = a->x ^ b->x, dy = a->y ^ b->y, dz = a->z ^ b->z;
  const uint32_t any = dx | dy | dz;
  int32_t p, q;
  if (dx > (any ^ dx)) {
    p = a->x;
    q = b->x;
  } else if (dy > (any ^ dy)) {
    p = a->y;
    q = b->y;
  } else {
    p = a->z;
    q = b->z;
  }
  return (p > q) - (p < q);
}

int TMROctant::compare(const TMROctant *oct) const {
  const int c = morton_order(this, oct);
  if (c != 0) return c;
  return level - oct->level;
}

int TMROctant::comparePosition(const TMROctant *oct) const {
  return morton_order(this, oct);
}

int TMROctant::compareNode(const TMROctant *oct) const {
  const int c = morton_order(this, oct);
  if (c != 0) return c;
  return info - oct->info;
}

int TMROctant::contains(TMROctant *oct) {
  const int32_t h = side_length(level);
  return (oct->block == block && oct->x >= x && oct->x < x + h &&
          oct->y >= y && oct->y < y + h && oct->z >= z && oct->z < z + h)
             ? 1
             : 0;
}

/* ---- TMROctantArray ----------------------------------------------------------- */
TMROctantArray::TMROctantArray(TMROctant *_array, int _size,
                               int _use_node_index)
    : use_node_index(_use_node_index),
      is_sorted(0),
      size(_size),
      max_size(_size),
      array(_array) {}

TMROctantArray::~TMROctantArray() { delete[] array; }

TMROctantArray *TMROctantArray::duplicate() {
  TMROctant *copy = new TMROctant[size > 0 ? size : 1];
  if (size > 0) memcpy(copy, array, (size_t)size * sizeof(TMROctant));
  TMROctantArray *dup = new TMROctantArray(copy, size, use_node_index);
  dup->is_sorted = is_sorted;
  return dup;
}

void TMROctantArray::getArray(TMROctant **_array, int *_size) {
  if (_array) *_array = array;
  if (_size) *_size = size;
}

void TMROctantArray::sort() {
  if (size > 1) {
    tmrgpu_ctx *ctx = tmr_b200_context();
    if (!ctx) return;
    int64_t nout = size;
    tmrgpu_array_sort(ctx, reinterpret_cast<tmrgpu_octant *>(array), size,
                      use_node_index, &nout);
    size = (int)nout;
  }
  is_sorted = 1;
}

/* One query against the sorted array: a binary search on the host with the
   scalar comparators, as the reference's bsearch (src/TMROctant.cpp:404-424) --
   O(log n) and no traffic.  (Uploading the array for a single probe made
   per-octant callers O(n) each; batches go through tmrgpu_array_contains.) */
TMROctant *TMROctantArray::contains(TMROctant *q, int use_position) {
  if (!is_sorted) sort();
  int lo = 0, hi = size;
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    const int c = use_node_index
                      ? q->compareNode(&array[mid])
                      : (use_position ? q->comparePosition(&array[mid])
                                      : q->compare(&array[mid]));
    if (c == 0) return &array[mid];
    if (c < 0) {
      hi = mid;
    } else {
      lo = mid + 1;
    }
  }
  return NULL;
}

void TMROctantArray::merge(TMROctantArray *list) {
  if (!is_sorted) sort();
  if (!list->is_sorted) list->sort();
  /* two-pointer set union on the full (position, level) order */
  std::vector<TMROctant> out;
  out.reserve((size_t)size + list->size);
  int i = 0, j = 0;
  while (i < size && j < list->size) {
    const int c = array[i].compare(&list->array[j]);
    if (c < 0) {
      out.push_back(array[i++]);
    } else if (c > 0) {
      out.push_back(list->array[j++]);
    } else {
      out.push_back(array[i++]);
      j++;
    }
  }
  while (i < size) out.push_back(array[i++]);
  while (j < list->size) out.push_back(list->array[j++]);
  const int len = (int)out.size();
  if (len > max_size) {
    delete[] array;
    array = new TMROctant[len];
    max_size = len;
  }
  if (len > 0) memcpy(array, out.data(), (size_t)len * sizeof(TMROctant));
  size = len;
}

/* ---- TMROctantQueue ----------------------------------------------------------- */
struct TMROctantQueue::Store {
  std::deque<TMROctant> q;
};

TMROctantQueue::TMROctantQueue() : store(new Store()) {}
TMROctantQueue::~TMROctantQueue() { delete store; }
int TMROctantQueue::length() { return (int)store->q.size(); }
void TMROctantQueue::push(TMROctant *oct) { store->q.push_back(*oct); }

TMROctant TMROctantQueue::pop() {
  TMROctant t;
  memset(&t, 0, sizeof(t));
  if (!store->q.empty()) {
    t = store->q.front();
    store->q.pop_front();
  }
  return t;
}

TMROctantArray *TMROctantQueue::toArray() {
  const int n = (int)store->q.size();
  TMROctant *a = new TMROctant[n > 0 ? n : 1];
  for (int i = 0; i < n; i++) a[i] = store->q[i];
  return new TMROctantArray(a, n);
}

/* ---- TMROctantHash ------------------------------------------------------------ */
namespace {
struct OctKey {
  int32_t block, x, y, z;
  int32_t tail; /* level (element mode) or info (node mode) */
  bool operator==(const OctKey &o) const {
    return block == o.block && x == o.x && y == o.y && z == o.z &&
           tail == o.tail;
  }
};
struct OctKeyHash {
  size_t operator()(const OctKey &k) const {
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    const int32_t v[5] = {k.block, k.x, k.y, k.z, k.tail};
    for (int i = 0; i < 5; i++) {
      h ^= (uint64_t)(uint32_t)v[i] + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
    }
    return (size_t)h;
  }
};
}  // namespace

struct TMROctantHash::Store {
  std::unordered_set<OctKey, OctKeyHash> set;
  std::vector<TMROctant> order; /* insertion order */
};

TMROctantHash::TMROctantHash(int _use_node_index)
    : store(new Store()), use_node_index(_use_node_index) {}
TMROctantHash::~TMROctantHash() { delete store; }

int TMROctantHash::addOctant(TMROctant *oct) {
  OctKey k = {oct->block, oct->x, oct->y, oct->z,
              use_node_index ? (int32_t)oct->info : (int32_t)oct->level};
  if (!store->set.insert(k).second) return 0;
  store->order.push_back(*oct);
  return 1;
}

TMROctantArray *TMROctantHash::toArray() {
  const int n = (int)store->order.size();
  TMROctant *a = new TMROctant[n > 0 ? n : 1];
  for (int i = 0; i < n; i++) a[i] = store->order[i];
  return new TMROctantArray(a, n, use_node_index);
}
