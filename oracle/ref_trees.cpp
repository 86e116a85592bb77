// TEST INFRASTRUCTURE ONLY — never linked into, imported by or called from the product
// (cbird_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it.
//
// Compiles the REFERENCE'S OWN search headers (by include path; no source is copied):
//     /root/reference/src/hamm.h            hamm64()
//     /root/reference/src/tree/vptree.h     VpTree  (what DctHashIndex ships, dcttree.h:26 VPTREE)
//     /root/reference/src/tree/radix.h      RadixMap_t (what DctVideoIndex ships, dctvideoindex.h:58-60)
//     /root/reference/src/tree/hammingtree.h HammingTree_t (what DctFeaturesIndex ships, incl. its cache file IO)
// and exposes them through a flat C interface for ctypes.  The DctTree glue below restates
// src/tree/dcttree.h:103-138 (it cannot be included: it pulls in index.h → Qt).
#include "ref_shim/qt_shim.h"
#include "ref_shim/qt_io_stub.h"

#include <string.h>

#include "hamm.h"
#include "tree/vptree.h"
#include "tree/radix.h"
#include "tree/hammingtree.h"

#include <chrono>
#include <thread>

namespace {

// dcttree.h:104-112 — value carried through the VP tree, with the min()/max() the tree asks for
struct TreeValue {
  uint64_t hash;
  uint32_t id;
  TreeValue() : hash(0), id(0) {}
  TreeValue(uint64_t h, uint32_t i) : hash(h), id(i) {}
  static TreeValue min() { return TreeValue(0, 0); }
  static TreeValue max() { return TreeValue(UINT64_MAX, 0); }
};
inline int treeDistance(TreeValue a, TreeValue b) { return hamm64(a.hash, b.hash); }  // dcttree.h:113
typedef VpTree<TreeValue, int, treeDistance> RefVpTree;                               // dcttree.h:115

// dctvideoindex.h:36-48 — 48-bit packed (video index, frame number)
#pragma pack(1)
struct __attribute__((packed)) VideoTreeIndex {
  uint32_t idx : 24;
  int frame : 24;
};
#pragma pack()
static_assert(sizeof(VideoTreeIndex) == 6, "packing");
typedef RadixMap_t<VideoTreeIndex> RefRadix;

}  // namespace

extern "C" {

int ref_hamm64(uint64_t a, uint64_t b) { return hamm64(a, b); }

// ---- DctTree (VpTree) : dcttree.h:117-137 -------------------------------------------------------
void* ref_dcttree_create(const uint64_t* hashes, const uint32_t* ids, int n) {
  std::vector<TreeValue> values;
  for (int i = 0; i < n; ++i) values.push_back(TreeValue(hashes[i], ids[i]));
  RefVpTree* t = new RefVpTree;
  if (n > 0) t->create(values);  // DctHashIndex::buildTree only builds when _numHashes>0 (dcthashindex.cpp:61-68)
  return t;
}
void ref_dcttree_destroy(void* t) { delete static_cast<RefVpTree*>(t); }

// returns number of matches (may exceed cap; only cap are written). Order = the tree's own order.
int ref_dcttree_search(void* t, uint64_t target, int threshold, uint32_t* out_ids, int* out_dist, int cap) {
  std::vector<int> distances;
  std::vector<TreeValue> results;
  static_cast<RefVpTree*>(t)->search(TreeValue(target, 0), threshold, &results, &distances);
  int n = int(results.size());
  for (int i = 0; i < n && i < cap; ++i) {
    out_ids[i] = results[i].id;
    out_dist[i] = distances[i];
  }
  return n;
}

// Batch driver used for parity (all needles) and for the CPU baseline (threads = host cores,
// static chunking like QtConcurrent::map over the global pool, database.cpp:1400-1432).
// out triples (needle index, id, dist) are appended per needle in tree order; returns total count.
// If out_* are null only counts; *elapsed_ms receives wall time of the search phase.
long long ref_dcttree_search_batch(void* t, const uint64_t* needles, int nq, int threshold, int threads,
                                   int* out_q, uint32_t* out_ids, int* out_dist, long long cap,
                                   double* elapsed_ms) {
  RefVpTree* tree = static_cast<RefVpTree*>(t);
  if (threads < 1) threads = 1;
  std::vector<std::vector<int>> q(threads), d(threads);
  std::vector<std::vector<uint32_t>> id(threads);
  std::vector<long long> counts(threads, 0);
  const bool keep = out_q != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int w) {
    int lo = int((long long)nq * w / threads), hi = int((long long)nq * (w + 1) / threads);
    std::vector<int> distances;
    std::vector<TreeValue> results;
    for (int i = lo; i < hi; ++i) {
      tree->search(TreeValue(needles[i], 0), threshold, &results, &distances);
      counts[w] += (long long)results.size();
      if (keep)
        for (size_t k = 0; k < results.size(); ++k) {
          q[w].push_back(i);
          id[w].push_back(results[k].id);
          d[w].push_back(distances[k]);
        }
    }
  };
  std::vector<std::thread> pool;
  for (int w = 1; w < threads; ++w) pool.emplace_back(work, w);
  work(0);
  for (auto& th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  if (elapsed_ms) *elapsed_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  long long total = 0, pos = 0;
  for (int w = 0; w < threads; ++w) {
    total += counts[w];
    if (keep)
      for (size_t k = 0; k < q[w].size() && pos < cap; ++k, ++pos) {
        out_q[pos] = q[w][k];
        out_ids[pos] = id[w][k];
        out_dist[pos] = d[w][k];
      }
  }
  return total;
}

// ---- RadixMap_t<VideoTreeIndex> : radix.h ------------------------------------------------------
void* ref_radix_create(unsigned radix) { return new RefRadix(radix); }
void ref_radix_destroy(void* r) { delete static_cast<RefRadix*>(r); }
unsigned long long ref_radix_index_of(void* r, uint64_t hash) { return static_cast<RefRadix*>(r)->indexOf(hash); }

void ref_radix_insert(void* r, const uint32_t* idx, const int* frame, const uint64_t* hashes, int n) {
  std::vector<RefRadix::Value> values;
  values.reserve(n);
  for (int i = 0; i < n; ++i) {
    VideoTreeIndex ti;
    ti.idx = idx[i];
    ti.frame = frame[i];
    values.push_back(RefRadix::Value(ti, hashes[i]));
  }
  static_cast<RefRadix*>(r)->insert(values);
}

// one query; matches come back in bucket insertion order (radix.h:187-210)
int ref_radix_search(void* r, uint64_t hash, int threshold, uint32_t* out_idx, int* out_frame,
                     uint64_t* out_hash, int* out_dist, int cap) {
  std::vector<RefRadix::Match> matches;
  static_cast<RefRadix*>(r)->search(hash, RefRadix::distance_t(threshold), matches);
  int n = int(matches.size());
  for (int i = 0; i < n && i < cap; ++i) {
    out_idx[i] = matches[i].value.index.idx;
    out_frame[i] = matches[i].value.index.frame;
    out_hash[i] = matches[i].value.hash;
    out_dist[i] = matches[i].distance;
  }
  return n;
}

// brute-force timing leg: radix 0 == one bucket == the same work the GPU scan does
long long ref_radix_search_batch_count(void* r, const uint64_t* needles, int nq, int threshold, int threads,
                                       double* elapsed_ms) {
  RefRadix* radix = static_cast<RefRadix*>(r);
  if (threads < 1) threads = 1;
  std::vector<long long> counts(threads, 0);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int w) {
    int lo = int((long long)nq * w / threads), hi = int((long long)nq * (w + 1) / threads);
    std::vector<RefRadix::Match> matches;
    for (int i = lo; i < hi; ++i) {
      matches.clear();
      radix->search(needles[i], RefRadix::distance_t(threshold), matches);
      counts[w] += (long long)matches.size();
    }
  };
  std::vector<std::thread> pool;
  for (int w = 1; w < threads; ++w) pool.emplace_back(work, w);
  work(0);
  for (auto& th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  if (elapsed_ms) *elapsed_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  long long total = 0;
  for (auto c : counts) total += c;
  return total;
}


// ---- HammingTree_t<uint32_t> : hammingtree.h ----------------------------------------------------
typedef HammingTree_t<uint32_t> RefHammingTree;

void* ref_htree_create() { return new RefHammingTree; }
void ref_htree_destroy(void* t) { delete static_cast<RefHammingTree*>(t); }
void ref_htree_insert(void* t, const uint32_t* indices, const uint64_t* hashes, int n) {
  std::vector<RefHammingTree::Value> values;
  values.reserve(n);
  for (int i = 0; i < n; ++i) values.push_back(RefHammingTree::Value(indices[i], hashes[i]));
  static_cast<RefHammingTree*>(t)->insert(values);
}
void ref_htree_remove(void* t, const uint32_t* indices, int n) {
  std::unordered_set<uint32_t> set(indices, indices + n);
  static_cast<RefHammingTree*>(t)->remove(set);
}
// matches sorted by distance (std::sort, ties in unspecified order)
int ref_htree_search(void* t, uint64_t hash, int threshold, uint32_t* out_index, uint64_t* out_hash, int* out_dist,
                     int cap) {
  std::vector<RefHammingTree::Match> matches;
  static_cast<RefHammingTree*>(t)->search(hash, threshold, matches);
  int n = int(matches.size());
  for (int i = 0; i < n && i < cap; ++i) {
    out_index[i] = matches[i].value.index;
    out_hash[i] = matches[i].value.hash;
    out_dist[i] = matches[i].distance;
  }
  return n;
}
void ref_htree_stats(void* t, int* num_nodes, int* max_height, int* num_values) {
  RefHammingTree::Stats st = static_cast<RefHammingTree*>(t)->stats();
  *num_nodes = st.numNodes;
  *max_height = st.maxHeight;
  *num_values = st.numValues;
}
int ref_htree_write(void* t, const char* path) {
  FILE* fp = fopen(path, "wb");
  if (!fp) return -1;
  QFile f(fp);
  static_cast<RefHammingTree*>(t)->write(f);
  fclose(fp);
  return 0;
}
int ref_htree_read(void* t, const char* path) {
  FILE* fp = fopen(path, "rb");
  if (!fp) return -1;
  QFile f(fp);
  bool ok = static_cast<RefHammingTree*>(t)->read(f);
  fclose(fp);
  return ok ? 0 : -2;
}

}  // extern "C"
