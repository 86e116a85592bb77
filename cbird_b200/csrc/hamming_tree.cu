// HammingTree_t on the device — src/tree/hammingtree.h (the search tree of DctFeaturesIndex, named by
// the north_star).
//
// Reference: a binary trie on the hash bits 0,1,2,... whose leaves hold at most CLUSTER_SIZE/8 = 8192
// hashes; a leaf that would exceed that splits on bit `depth` (:366-425). search() follows ONE path
// by the needle's low bits and scans that leaf linearly (:244-293): an approximate search whose miss
// rate grows with the tree depth.  Because a node splits exactly when the number of hashes that ever
// reached it exceeds 8192, the final shape depends only on the multiset of hashes, not on insertion
// order: node (depth d, low bits p) is internal iff more than 8192 hashes have those d low bits.
//
// Here the hashes are laid out leaf-contiguously in HBM; a batch of needles is grouped by leaf on the
// host and one tile-list launch of the scan kernel (scan64.cu) tests every (leaf rows x leaf needles)
// block.  Result sets equal the reference's; within a distance the order is unspecified there
// (std::sort :103-108) and (distance, index, hash) here.  Divergence: leaves with fewer than 4
// hashes hit an unsigned underflow in the reference's unrolled loop (`count - 4`, :264); this build
// simply searches them.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <map>
#include <unordered_set>

#include "common.h"

namespace cbird {

namespace {
constexpr uint32_t kLeafCapacity = 64 * 1024 / 8;  // CLUSTER_SIZE / sizeof(hash_t), hammingtree.h:57,377
constexpr int kMaxSplitDepth = 30;                  // `1 << bit` is an int shift in the reference (:236,:246)
constexpr int kKeyBits = kMaxSplitDepth + 1;

// Pre-order position of a hash in the trie: bit 0 decides first, the set child ("left") comes before the
// clear child, so the key is the complement of the low 31 bits with bit 0 most significant.
__global__ void trie_keys_kernel(const uint64_t* __restrict__ hash, uint32_t n, uint32_t* __restrict__ key,
                                 uint32_t* __restrict__ val) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  key[i] = __brev(~uint32_t(hash[i])) >> 1;
  val[i] = i;
}

__global__ void trie_gather_kernel(const uint32_t* __restrict__ order, uint32_t n, const uint64_t* __restrict__ hash,
                                   const uint32_t* __restrict__ index, uint64_t* __restrict__ s_hash,
                                   uint32_t* __restrict__ s_index) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = order[i];
  s_hash[i] = hash[r];
  s_index[i] = index[r];
}
}  // namespace

struct HammingTree {
  // values in insertion order
  std::vector<uint32_t> index;
  std::vector<uint64_t> hash;

  // trie (built lazily): nodes in pre-order; leaf nodes own rows [row_begin, row_end) of the sorted arrays
  struct Node {
    int bit = -1;               // split bit for internal nodes
    int set_child = -1;         // child for hash bit == 1 ("left" in the reference), -1 for leaves
    int clear_child = -1;       // child for hash bit == 0 ("right")
    uint32_t row_begin = 0, row_end = 0;
  };
  std::vector<Node> nodes;
  int max_height = 0;
  bool built = false;
  std::vector<uint32_t> s_index;  // leaf-contiguous copies
  std::vector<uint64_t> s_hash;

  int device = 0;
  std::mutex mu;
  cudaStream_t stream = nullptr;
  DevBuf<uint64_t> d_hash, d_q;
  DevBuf<cb_scan_tile> d_tiles;
  DevBuf<cb_pair> d_pairs;
  DevBuf<unsigned long long> d_counts;
  unsigned long long* h_counts = nullptr;

  ~HammingTree() {
    if (h_counts) cudaFreeHost(h_counts);
    if (stream) cudaStreamDestroy(stream);
  }

  // node over sorted rows [lo, hi) that share their `depth` low bits: internal iff more than 8192 hashes
  // reached it (:377-405); the children are the runs with bit `depth` set / clear (a binary search, the rows
  // are sorted by trie position)
  int build_node(const std::vector<uint32_t>& keys, uint32_t lo, uint32_t hi, int depth) {
    const int id = int(nodes.size());
    nodes.push_back(Node());
    max_height = std::max(max_height, depth);
    if (hi - lo > kLeafCapacity && depth <= kMaxSplitDepth) {
      const uint32_t bit = 1u << (kKeyBits - 1 - depth);  // key bit of hash bit `depth`; 0 there = hash bit set
      const uint32_t split = uint32_t(std::partition_point(keys.begin() + lo, keys.begin() + hi,
                                                           [bit](uint32_t k) { return !(k & bit); }) - keys.begin());
      nodes[id].bit = depth;
      const int a = build_node(keys, lo, split, depth + 1);
      const int b = build_node(keys, split, hi, depth + 1);
      nodes[id].set_child = a;
      nodes[id].clear_child = b;
    } else {
      nodes[id].row_begin = lo;
      nodes[id].row_end = hi;
    }
    return id;
  }

  // The trie shape depends only on the multiset of hashes, so it is built from a sort: keys on the device, a
  // stable radix sort (a leaf keeps insertion order like the reference's append), a gather into the
  // leaf-contiguous layout, then the node table from binary searches over the sorted keys.
  int build() {
    if (built) return CB_OK;
    int rc = ensure_device();
    if (rc != CB_OK) return rc;
    if (!stream) {
      device = current_device();
      CB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    }
    if (!h_counts) CB_CUDA(cudaMallocHost(&h_counts, 2 * sizeof(unsigned long long)));
    if ((rc = d_counts.reserve(2)) != CB_OK) return rc;
    nodes.clear();
    s_hash.clear();
    s_index.clear();
    max_height = 0;
    const size_t n = hash.size();
    if (n > 0xFFFFF000ull) {
      set_error("hamming tree: %zu hashes exceed the 32-bit row index", n);
      return CB_ERR_UNSUPPORTED;
    }
    if (n) {
      CB_CUDA(cudaSetDevice(device));
      DevBuf<uint64_t> d_raw;
      DevBuf<uint32_t> d_index, d_sindex, d_key, d_key2, d_val, d_val2;
      DevBuf<unsigned char> d_temp;
      if ((rc = d_raw.reserve(n)) != CB_OK || (rc = d_index.reserve(n)) != CB_OK || (rc = d_sindex.reserve(n)) != CB_OK ||
          (rc = d_key.reserve(n)) != CB_OK || (rc = d_key2.reserve(n)) != CB_OK || (rc = d_val.reserve(n)) != CB_OK ||
          (rc = d_val2.reserve(n)) != CB_OK || (rc = d_hash.reserve(n + 2)) != CB_OK)
        return rc;
      CB_CUDA(cudaMemcpyAsync(d_raw.p, hash.data(), n * 8, cudaMemcpyHostToDevice, stream));
      CB_CUDA(cudaMemcpyAsync(d_index.p, index.data(), n * 4, cudaMemcpyHostToDevice, stream));
      const unsigned blocks = unsigned((n + 255) / 256);
      trie_keys_kernel<<<blocks, 256, 0, stream>>>(d_raw.p, uint32_t(n), d_key.p, d_val.p);
      CB_CUDA(cudaGetLastError());
      size_t tb = 0;
      CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_key.p, d_key2.p, d_val.p, d_val2.p, static_cast<long long>(n), 0,
                                              kKeyBits, stream));
      if ((rc = d_temp.reserve(tb + 16)) != CB_OK) return rc;
      CB_CUDA(cub::DeviceRadixSort::SortPairs(d_temp.p, tb, d_key.p, d_key2.p, d_val.p, d_val2.p, static_cast<long long>(n), 0,
                                              kKeyBits, stream));
      trie_gather_kernel<<<blocks, 256, 0, stream>>>(d_val2.p, uint32_t(n), d_raw.p, d_index.p, d_hash.p, d_sindex.p);
      CB_CUDA(cudaGetLastError());
      counters().launches += 2;
      std::vector<uint32_t> keys(n);
      s_hash.resize(n);
      s_index.resize(n);
      CB_CUDA(cudaMemcpyAsync(keys.data(), d_key2.p, n * 4, cudaMemcpyDeviceToHost, stream));
      CB_CUDA(cudaMemcpyAsync(s_hash.data(), d_hash.p, n * 8, cudaMemcpyDeviceToHost, stream));
      CB_CUDA(cudaMemcpyAsync(s_index.data(), d_sindex.p, n * 4, cudaMemcpyDeviceToHost, stream));
      CB_CUDA(cudaStreamSynchronize(stream));
      build_node(keys, 0, uint32_t(n), 0);
    }
    built = true;
    return CB_OK;
  }

  const Node* leaf_for(uint64_t h) const {
    if (nodes.empty()) return nullptr;
    const Node* n = &nodes[0];
    while (n->set_child >= 0) n = &nodes[((h >> n->bit) & 1) ? n->set_child : n->clear_child];  // :245-249
    return n;
  }

  // all (needle, row) pairs with distance < threshold inside the needle's leaf; rows index s_hash/s_index
  int search(const uint64_t* needles, int64_t nq, int threshold, std::vector<cb_pair>& out) {
    out.clear();
    int rc = build();
    if (rc != CB_OK) return rc;
    if (!nq || s_hash.empty() || threshold <= 0) return CB_OK;
    CB_CUDA(cudaSetDevice(device));
    // group needles by leaf
    std::vector<std::pair<const Node*, uint32_t>> byleaf(nq);
    for (int64_t i = 0; i < nq; ++i) byleaf[i] = {leaf_for(needles[i]), uint32_t(i)};
    std::stable_sort(byleaf.begin(), byleaf.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<uint64_t> sorted(nq);
    std::vector<cb_scan_tile> tiles;
    uint64_t pair_tests = 0;
    for (int64_t i = 0; i < nq;) {
      int64_t j = i;
      while (j < nq && byleaf[j].first == byleaf[i].first) ++j;
      const Node* leaf = byleaf[i].first;
      for (uint32_t r = leaf->row_begin; r < leaf->row_end; r += 2048)
        tiles.push_back({r, std::min<uint32_t>(2048, leaf->row_end - r), uint32_t(i), uint32_t(j - i)});
      pair_tests += uint64_t(leaf->row_end - leaf->row_begin) * uint64_t(j - i);
      i = j;
    }
    for (int64_t i = 0; i < nq; ++i) sorted[i] = needles[byleaf[i].second];
    if (tiles.empty()) return CB_OK;
    if ((rc = d_q.reserve(size_t(nq) + 2)) != CB_OK || (rc = d_tiles.reserve(tiles.size())) != CB_OK) return rc;
    CB_CUDA(cudaMemcpyAsync(d_q.p, sorted.data(), size_t(nq) * 8, cudaMemcpyHostToDevice, stream));
    CB_CUDA(cudaMemcpyAsync(d_tiles.p, tiles.data(), tiles.size() * sizeof(cb_scan_tile), cudaMemcpyHostToDevice, stream));
    unsigned long long cap = d_pairs.cap ? d_pairs.cap : (1ull << 18);
    for (int attempt = 0; attempt < 3; ++attempt) {
      if ((rc = d_pairs.reserve(cap)) != CB_OK) return rc;
      cap = d_pairs.cap;
      CB_CUDA(cudaMemsetAsync(d_counts.p, 0, 2 * sizeof(unsigned long long), stream));
      Scan64Launch L{d_hash.p, uint32_t(s_hash.size()), d_q.p, uint32_t(nq), threshold, 0, d_pairs.p, cap, d_counts.p};
      rc = scan64_tiles_launch(L, d_tiles.p, uint32_t(tiles.size()), pair_tests, stream);
      if (rc != CB_OK) return rc;
      CB_CUDA(cudaMemcpyAsync(h_counts, d_counts.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      CB_CUDA(cudaStreamSynchronize(stream));
      if (h_counts[0] <= cap) break;
      cap = h_counts[0] + h_counts[0] / 8 + 1024;
      if (attempt == 2) {
        set_error("hamming tree: hit list overflow persisted");
        return CB_ERR_CUDA;
      }
    }
    const size_t n = size_t(h_counts[0]);
    counters().hits += n;
    out.resize(n);
    if (n) {
      CB_CUDA(cudaMemcpyAsync(out.data(), d_pairs.p, n * sizeof(cb_pair), cudaMemcpyDeviceToHost, stream));
      CB_CUDA(cudaStreamSynchronize(stream));
    }
    for (cb_pair& p : out) p.b = byleaf[p.b].second;  // back to the caller's needle order
    return CB_OK;
  }
};

}  // namespace cbird

using namespace cbird;

struct cb_hamming_tree {
  HammingTree impl;
};

extern "C" {

cb_hamming_tree* cb_hamming_tree_create(void) { return new (std::nothrow) cb_hamming_tree; }

void cb_hamming_tree_destroy(cb_hamming_tree* t) {
  if (!t) return;
  if (t->impl.stream) cudaSetDevice(t->impl.device);
  delete t;
}

int cb_hamming_tree_insert(cb_hamming_tree* t, const uint32_t* indices, const uint64_t* hashes, int64_t n) {
  CB_API_BEGIN
  if (!t || n < 0 || (n && (!indices || !hashes))) {
    set_error("cb_hamming_tree_insert: invalid argument");
    return CB_ERR_INVALID;
  }
  HammingTree& T = t->impl;
  std::lock_guard<std::mutex> lock(T.mu);
  if (T.hash.size() + size_t(n) > 0xFFFFF000ull) {
    set_error("cb_hamming_tree_insert: too many hashes");
    return CB_ERR_UNSUPPORTED;
  }
  T.index.insert(T.index.end(), indices, indices + n);
  T.hash.insert(T.hash.end(), hashes, hashes + n);
  T.built = false;
  return CB_OK;
  CB_API_END
}

int cb_hamming_tree_remove(cb_hamming_tree* t, const uint32_t* indices, int64_t n) {
  CB_API_BEGIN
  if (!t || n < 0 || (n && !indices)) {
    set_error("cb_hamming_tree_remove: invalid argument");
    return CB_ERR_INVALID;
  }
  HammingTree& T = t->impl;
  std::lock_guard<std::mutex> lock(T.mu);
  std::unordered_set<uint32_t> gone(indices, indices + n);
  for (auto& ix : T.index)
    if (gone.count(ix)) ix = 0;  // the hash stays searchable with index 0 (:351-358)
  for (auto& ix : T.s_index)
    if (gone.count(ix)) ix = 0;
  return CB_OK;
  CB_API_END
}

int cb_hamming_tree_stats(cb_hamming_tree* t, int32_t* num_nodes, int32_t* max_height, int64_t* num_values) {
  CB_API_BEGIN
  if (!t) return CB_ERR_INVALID;
  HammingTree& T = t->impl;
  std::lock_guard<std::mutex> lock(T.mu);
  int rc = T.build();
  if (rc != CB_OK) return rc;
  if (num_nodes) *num_nodes = int32_t(T.nodes.size());
  if (max_height) *max_height = T.max_height;
  if (num_values) *num_values = int64_t(T.hash.size());
  return CB_OK;
  CB_API_END
}

// search(): every value of the needle's leaf with distance < threshold, sorted by distance (:99-108)
int cb_hamming_tree_search_batch_alloc(cb_hamming_tree* t, const uint64_t* needles, int64_t n_needles, int threshold,
                                       cb_tree_match** out, int64_t* n_out) {
  CB_API_BEGIN
  if (!t || !out || !n_out || n_needles < 0 || (n_needles && !needles)) {
    set_error("cb_hamming_tree_search_batch_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  HammingTree& T = t->impl;
  std::lock_guard<std::mutex> lock(T.mu);
  std::vector<cb_pair> pairs;
  int rc = T.search(needles, n_needles, threshold, pairs);
  if (rc != CB_OK) return rc;
  std::vector<cb_tree_match> m(pairs.size());
  for (size_t i = 0; i < pairs.size(); ++i)
    m[i] = cb_tree_match{pairs[i].b, T.s_index[pairs[i].a], int32_t(pairs[i].dist), 0, T.s_hash[pairs[i].a]};
  std::sort(m.begin(), m.end(), [](const cb_tree_match& x, const cb_tree_match& y) {
    if (x.needle != y.needle) return x.needle < y.needle;
    if (x.distance != y.distance) return x.distance < y.distance;
    if (x.index != y.index) return x.index < y.index;
    return x.hash < y.hash;
  });
  *n_out = int64_t(m.size());
  *out = static_cast<cb_tree_match*>(malloc(std::max<size_t>(1, m.size()) * sizeof(cb_tree_match)));
  if (!*out) {
    set_error("out of host memory");
    return CB_ERR_INVALID;
  }
  if (!m.empty()) memcpy(*out, m.data(), m.size() * sizeof(cb_tree_match));
  return CB_OK;
  CB_API_END
}

// DctFeaturesIndex::find (src/dctfeaturesindex.cpp:260-358) on top of the tree: every needle hash votes
// for the media of its 10 nearest matches; media with more votes score lower. needle hashes may be
// omitted (n == 0) when needle_id > 0: they are then collected from the tree (findIndex, :270-274).
// Ties at the 10-cut: the reference keeps whatever its unstable std::sort left first; here the order is
// (distance, index, hash).
int cb_hamming_tree_find_votes(cb_hamming_tree* t, const uint64_t* needle_hashes, int64_t n, uint32_t needle_id,
                               int threshold, cb_match* out, int64_t cap, int64_t* n_out) {
  CB_API_BEGIN
  if (!t || !n_out || n < 0 || (n && !needle_hashes)) {
    set_error("cb_hamming_tree_find_votes: invalid argument");
    return CB_ERR_INVALID;
  }
  *n_out = 0;
  HammingTree& T = t->impl;
  std::lock_guard<std::mutex> lock(T.mu);
  std::vector<uint64_t> own;
  if (n == 0) {
    if (needle_id > 0)
      for (size_t i = 0; i < T.index.size(); ++i)
        if (T.index[i] == needle_id) own.push_back(T.hash[i]);
    if (own.empty()) return CB_OK;  // "needle has no hashes" (:276-279)
    needle_hashes = own.data();
    n = int64_t(own.size());
  }
  std::vector<cb_pair> pairs;
  int rc = T.search(needle_hashes, n, threshold, pairs);
  if (rc != CB_OK) return rc;
  struct Cand {
    uint32_t needle, index;
    int32_t dist;
    uint64_t hash;
  };
  std::vector<Cand> cand(pairs.size());
  for (size_t i = 0; i < pairs.size(); ++i)
    cand[i] = Cand{pairs[i].b, T.s_index[pairs[i].a], int32_t(pairs[i].dist), T.s_hash[pairs[i].a]};
  std::sort(cand.begin(), cand.end(), [](const Cand& x, const Cand& y) {
    if (x.needle != y.needle) return x.needle < y.needle;
    if (x.dist != y.dist) return x.dist < y.dist;
    if (x.index != y.index) return x.index < y.index;
    return x.hash < y.hash;
  });
  std::map<uint32_t, uint32_t> matches;  // QMap<uint32_t, uint32_t> :293
  std::map<uint32_t, int> scores;
  uint32_t maxMatches = 0;
  for (size_t i = 0; i < cand.size();) {
    size_t j = i;
    while (j < cand.size() && cand[j].needle == cand[i].needle) ++j;
    for (size_t k = i; k < j && k < i + 10; ++k) {  // "take the first 10" :299
      const int index = int(cand[k].index);
      if (index <= 0) continue;  // deleted :305
      const uint32_t mediaId = uint32_t(index);
      matches[mediaId] += 1;
      scores[mediaId] += cand[k].dist;
      if (needle_id != mediaId) maxMatches = std::max(matches[mediaId], maxMatches);  // :320
    }
    i = j;
  }
  int64_t w = 0;
  for (auto& kv : matches) {  // :333-356
    cb_match m{kv.first, 0, -1, -1, 0};
    const float avgScore = float(scores[kv.first]) / float(kv.second);
    if (kv.first == needle_id) m.score = -1;
    else if (maxMatches == 1) m.score = int32_t(10 * avgScore);
    else m.score = int32_t(maxMatches - kv.second);
    if (w < cap) out[w] = m;
    ++w;
  }
  *n_out = w;
  return w > cap ? CB_ERR_CAPACITY : CB_OK;
  CB_API_END
}

// cache file v2 (:156-200, :472-521): "cbird hamming tree:2:<sizeof index>:8:65536\n" + pre-order nodes:
// bool isLeaf; inner: int bit, set-child ("left"), clear-child ("right"); leaf: u32 count, index[count], hash[count]
int cb_hamming_tree_write(cb_hamming_tree* t, const char* path) {
  CB_API_BEGIN
  if (!t || !path) return CB_ERR_INVALID;
  HammingTree& T = t->impl;
  std::lock_guard<std::mutex> lock(T.mu);
  int rc = T.build();
  if (rc != CB_OK) return rc;
  FILE* f = fopen(path, "wb");
  if (!f) {
    set_error("cannot write %s", path);
    return CB_ERR_INVALID;
  }
  fprintf(f, "cbird hamming tree:%d:%d:%d:%d\n", 2, int(sizeof(uint32_t)), int(sizeof(uint64_t)), 64 * 1024);
  for (const HammingTree::Node& n : T.nodes) {  // nodes are stored in pre-order, set-child first
    const bool leaf = n.set_child < 0;
    fwrite(&leaf, sizeof(bool), 1, f);
    if (!leaf) {
      fwrite(&n.bit, sizeof(int), 1, f);
    } else {
      const uint32_t count = n.row_end - n.row_begin;
      fwrite(&count, sizeof(uint32_t), 1, f);
      if (count) {
        fwrite(T.s_index.data() + n.row_begin, sizeof(uint32_t), count, f);
        fwrite(T.s_hash.data() + n.row_begin, sizeof(uint64_t), count, f);
      }
    }
  }
  fclose(f);
  return CB_OK;
  CB_API_END
}

// The cache file is untrusted input: every count is checked against the bytes left in the file, split bits
// against [0, 63] (leaf_for shifts by them) and the recursion against the deepest tree 64-bit hashes can
// make, so a corrupt or truncated file is rejected instead of exhausting the stack or the heap.
struct TreeReader {
  FILE* f;
  long long left;  // bytes not yet consumed
  bool take(void* dst, size_t bytes) {
    if ((long long)bytes > left) return false;
    if (bytes && fread(dst, 1, bytes, f) != bytes) return false;
    left -= (long long)bytes;
    return true;
  }
};

static bool read_node(TreeReader& R, HammingTree& T, int depth) {
  if (depth > 64) return false;
  uint8_t leaf = 0;
  if (!R.take(&leaf, 1) || leaf > 1) return false;
  const int id = int(T.nodes.size());
  T.nodes.push_back(HammingTree::Node());
  T.max_height = std::max(T.max_height, depth);
  if (!leaf) {
    int bit = 0;
    if (!R.take(&bit, sizeof(int)) || bit < 0 || bit > 63) return false;
    T.nodes[id].bit = bit;
    const int a = int(T.nodes.size());
    if (!read_node(R, T, depth + 1)) return false;
    const int b = int(T.nodes.size());
    if (!read_node(R, T, depth + 1)) return false;
    T.nodes[id].set_child = a;
    T.nodes[id].clear_child = b;
  } else {
    uint32_t count = 0;
    if (!R.take(&count, sizeof(uint32_t))) return false;
    if ((long long)count * 12 > R.left || T.s_hash.size() + count > 0xFFFFF000ull) return false;
    T.nodes[id].row_begin = uint32_t(T.s_hash.size());
    if (count) {
      const size_t at = T.s_hash.size();
      T.s_index.resize(at + count);
      T.s_hash.resize(at + count);
      if (!R.take(T.s_index.data() + at, sizeof(uint32_t) * size_t(count))) return false;
      if (!R.take(T.s_hash.data() + at, sizeof(uint64_t) * size_t(count))) return false;
    }
    T.nodes[id].row_end = uint32_t(T.s_hash.size());
  }
  return true;
}

int cb_hamming_tree_read(cb_hamming_tree* t, const char* path) {
  CB_API_BEGIN
  if (!t || !path) return CB_ERR_INVALID;
  HammingTree& T = t->impl;
  std::lock_guard<std::mutex> lock(T.mu);
  FILE* f = fopen(path, "rb");
  if (!f) {
    set_error("cannot open %s", path);
    return CB_ERR_INVALID;
  }
  char line[128] = {0};
  if (!fgets(line, sizeof(line), f) || strcmp(line, "cbird hamming tree:2:4:8:65536\n") != 0) {
    fclose(f);
    set_error("%s: incompatible hamming tree header", path);  // "old file format?" / "incompatible format" :172-188
    return CB_ERR_INVALID;
  }
  T.nodes.clear();
  T.s_hash.clear();
  T.s_index.clear();
  T.max_height = 0;
  const long at = ftell(f);
  fseek(f, 0, SEEK_END);
  TreeReader R{f, (long long)ftell(f) - at};
  fseek(f, at, SEEK_SET);
  bool ok = false;
  try {
    ok = read_node(R, T, 0);  // bytes after the root's subtree are ignored, as the reference's reader does
  } catch (...) {
    ok = false;
  }
  fclose(f);
  if (!ok) {
    T.nodes.clear();
    T.s_hash.clear();
    T.s_index.clear();
    T.hash.clear();
    T.index.clear();
    T.built = false;
    set_error("%s: truncated or corrupt hamming tree file", path);
    return CB_ERR_INVALID;
  }
  // the file's shape is authoritative until the next insert
  T.hash = T.s_hash;
  T.index = T.s_index;
  int rc = ensure_device();
  if (rc != CB_OK) return rc;
  if (!T.stream) {
    T.device = current_device();
    CB_CUDA(cudaStreamCreateWithFlags(&T.stream, cudaStreamNonBlocking));
  }
  if (!T.h_counts) CB_CUDA(cudaMallocHost(&T.h_counts, 2 * sizeof(unsigned long long)));
  if ((rc = T.d_counts.reserve(2)) != CB_OK || (rc = T.d_hash.reserve(T.s_hash.size() + 2)) != CB_OK) return rc;
  if (!T.s_hash.empty()) {
    CB_CUDA(cudaMemcpyAsync(T.d_hash.p, T.s_hash.data(), T.s_hash.size() * 8, cudaMemcpyHostToDevice, T.stream));
    CB_CUDA(cudaStreamSynchronize(T.stream));
  }
  T.built = true;
  return CB_OK;
  CB_API_END
}

}  // extern "C"
