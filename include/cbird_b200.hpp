// cbird_b200.hpp — header-only C++ host layer over the C ABI (cbird_b200.h), shaped like cbird's own
// plugin surface so that a cbird build can subclass `Index` (src/index.h:150-281) with one-line
// forwards.  Same method names, argument meaning and soft-error behaviour as the
// reference classes (warn + empty result, src/dcthashindex.cpp:197-205); std:: containers stand in
// for the Qt ones (QVector -> std::vector, QSet -> std::set) because this header must compile
// without Qt.  No CPU fallback: without a CUDA device every call reports through `onWarning`.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "cbird_b200.h"

namespace cbird_b200 {

/// where qWarning() output goes; cbird installs its own handler, the default prints to stderr
inline std::function<void(const std::string&)>& onWarning() {
  static std::function<void(const std::string&)> fn = [](const std::string& s) { fprintf(stderr, "cbird_b200: %s\n", s.c_str()); };
  return fn;
}
inline bool ok(int status, const char* what) {
  if (status == CB_OK) return true;
  onWarning()(std::string(what) + ": " + cb_last_error());
  return false;
}

/// MatchRange, src/media.h:62-78
struct MatchRange {
  int srcIn = -1, dstIn = -1, len = 0;
  MatchRange() {}
  MatchRange(int s, int d, int l) : srcIn(s), dstIn(d), len(l) {}
  bool operator<(const MatchRange& r) const { return srcIn < r.srcIn; }
};

/// Index::Match, src/index.h:157-167
struct Match {
  uint32_t mediaId = 0;
  int score = 0;
  MatchRange range;
  Match() {}
  Match(uint32_t id, int s) : mediaId(id), score(s) {}
};
inline bool operator<(const Match& a, const Match& b) { return a.score < b.score; }  // src/index.h:284-286

/// the SearchParams fields the path reads, src/index.h:74-121 (same names and defaults)
struct SearchParams {
  enum { AlgoDCT = 0, AlgoDCTFeatures = 1, AlgoCVFeatures = 2, AlgoColor = 3, AlgoVideo = 4 };
  int algo = AlgoDCT, dctThresh = 5, cvThresh = 25, minMatches = 1, maxMatches = 5, maxThresh = 0;
  int skipFrames = 300, minFramesMatched = 30, minFramesNear = 60, videoRadix = 10;
  bool filterSelf = true, verbose = false;
  uint32_t target = 0;
  cb_params c() const {
    cb_params p;
    cb_params_default(&p);
    p.algo = algo; p.dctThresh = dctThresh; p.cvThresh = cvThresh; p.minMatches = minMatches;
    p.maxMatches = maxMatches; p.maxThresh = maxThresh; p.skipFrames = skipFrames;
    p.minFramesMatched = minFramesMatched; p.minFramesNear = minFramesNear; p.videoRadix = videoRadix;
    p.filterSelf = filterSelf; p.verbose = verbose; p.target = target;
    return p;
  }
};

/// Multi-GPU: call once, before any index is created — where Engine::Engine makes the Index objects
/// (src/engine.cpp:38-45). Afterwards DctHashIndex::load / similar fan out over the devices by themselves.
inline bool init(const std::vector<int>& devices) { return ok(cb_init(devices.data(), int(devices.size())), "cbird::init"); }
inline void shutdown() { cb_shutdown(); }

/// the slice of Media the indexes read (src/media.h:243,253,300,417,536-539)
struct Media {
  enum { TypeImage = 1, TypeVideo = 2 };
  int id = 0, type = TypeImage;
  uint64_t dctHash = 0;
  std::vector<int32_t> frames;        // VideoIndex::frames
  std::vector<uint64_t> hashes;       // VideoIndex::hashes
  std::vector<uint8_t> descriptors;   // KeyPointDescriptors, rows x 32
  int matchRangeDstIn = -1;
  std::string path;
};

namespace detail {
inline std::vector<Match> toMatches(const std::vector<cb_match>& m, int64_t n) {
  std::vector<Match> out;
  for (int64_t i = 0; i < n; ++i) {
    Match r(m[size_t(i)].mediaId, m[size_t(i)].score);
    r.range = MatchRange(m[size_t(i)].srcIn, m[size_t(i)].dstIn, m[size_t(i)].len);
    out.push_back(r);
  }
  return out;
}
template <typename Call>
std::vector<Match> collect(const char* what, Call call) {  // grow-and-retry on CB_ERR_CAPACITY
  std::vector<cb_match> buf(256);
  int64_t n = 0;
  int rc = call(buf.data(), int64_t(buf.size()), &n);
  if (rc == CB_ERR_CAPACITY) {
    buf.resize(size_t(n));
    rc = call(buf.data(), int64_t(buf.size()), &n);
  }
  if (!ok(rc, what)) return {};
  return toMatches(buf, n);
}
}  // namespace detail

/// src/dcthashindex.{h,cpp}
class DctHashIndex {
  cb_dct_index* _ix;
  explicit DctHashIndex(cb_dct_index* ix) : _ix(ix) {}

 public:
  DctHashIndex() : _ix(cb_dct_index_create()) {}
  ~DctHashIndex() { cb_dct_index_destroy(_ix); }
  DctHashIndex(const DctHashIndex&) = delete;
  DctHashIndex& operator=(const DctHashIndex&) = delete;
  int id() const { return SearchParams::AlgoDCT; }
  cb_dct_index* handle() const { return _ix; }

  bool isLoaded() const { return cb_dct_index_is_loaded(_ix) != 0; }
  size_t memoryUsage() const { return cb_dct_index_memory_usage(_ix); }
  int count() const { return int(cb_dct_index_count(_ix)); }
  /// load(): the rows of "select id,phash_dct from media where type=1" (:89)
  void load(const std::vector<uint32_t>& ids, const std::vector<uint64_t>& hashes) {
    ok(cb_dct_index_load(_ix, ids.data(), hashes.data(), int64_t(ids.size())), "DctHashIndex::load");
  }
  void save() {}  // no-op like :116-120
  void add(const std::vector<Media>& media) {
    std::vector<uint32_t> ids;
    std::vector<uint64_t> hashes;
    for (const Media& m : media) {
      ids.push_back(uint32_t(m.id));
      hashes.push_back(m.dctHash);
    }
    ok(cb_dct_index_add(_ix, ids.data(), hashes.data(), int64_t(ids.size())), "DctHashIndex::add");
  }
  void remove(const std::vector<int>& ids) { ok(cb_dct_index_remove(_ix, ids.data(), int64_t(ids.size())), "DctHashIndex::remove"); }
  std::set<uint32_t> mediaIds() const {
    int64_t n = 0;
    cb_dct_index_media_ids(_ix, nullptr, 0, &n);
    std::vector<uint32_t> v(size_t(n) + 1);
    cb_dct_index_media_ids(_ix, v.data(), n, &n);
    return std::set<uint32_t>(v.begin(), v.begin() + n);
  }
  std::vector<Match> find(const Media& m, const SearchParams& p) {
    const cb_params c = p.c();
    return detail::collect("DctHashIndex::find", [&](cb_match* out, int64_t cap, int64_t* n) {
      return cb_dct_index_find(_ix, m.dctHash, &c, out, cap, n);
    });
  }
  /// caller deletes, like Index::slice (src/database.cpp:1331)
  DctHashIndex* slice(const std::set<uint32_t>& mediaIds) const {
    std::vector<uint32_t> v(mediaIds.begin(), mediaIds.end());
    cb_dct_index* h = cb_dct_index_slice(_ix, v.data(), int64_t(v.size()));
    if (!h) {
      onWarning()(std::string("DctHashIndex::slice: ") + cb_last_error());
      return nullptr;
    }
    return new DctHashIndex(h);
  }
  /// `-similar` batch hook (src/database.cpp:1400-1432 + searchIndex post step :1729-1737):
  /// result[row] = matches of the row-th indexed item
  std::vector<std::vector<Match>> similar(const SearchParams& p) {
    const cb_params c = p.c();
    int64_t* offsets = nullptr;
    cb_hit* hits = nullptr;
    int64_t n = 0;
    std::vector<std::vector<Match>> out(static_cast<size_t>(count()));
    if (!ok(cb_dct_index_similar_alloc(_ix, &c, &offsets, &hits, &n), "DctHashIndex::similar")) return out;
    for (size_t row = 0; row < out.size(); ++row)
      for (int64_t k = offsets[row]; k < offsets[row + 1]; ++k) out[row].push_back(Match(hits[k].mediaId, hits[k].score));
    cb_free(offsets);
    cb_free(hits);
    return out;
  }
};

/// src/dctvideoindex.{h,cpp}
class DctVideoIndex {
  cb_video_index* _ix;
  explicit DctVideoIndex(cb_video_index* ix) : _ix(ix) {}

 public:
  DctVideoIndex() : _ix(cb_video_index_create()) {}
  ~DctVideoIndex() { cb_video_index_destroy(_ix); }
  DctVideoIndex(const DctVideoIndex&) = delete;
  DctVideoIndex& operator=(const DctVideoIndex&) = delete;
  int id() const { return SearchParams::AlgoVideo; }
  cb_video_index* handle() const { return _ix; }

  bool isLoaded() const { return cb_video_index_is_loaded(_ix) != 0; }
  size_t memoryUsage() const { return cb_video_index_memory_usage(_ix); }
  int count() const { return int(cb_video_index_count(_ix)); }
  /// load(): "select id from media where type=2 order by id" (:181); dataPath holds <id>.vdx
  void load(const std::vector<uint32_t>& ids, const std::string& dataPath = std::string()) {
    if (!ok(cb_video_index_load(_ix, ids.data(), int64_t(ids.size())), "DctVideoIndex::load")) return;
    if (!dataPath.empty())
      for (uint32_t id : ids) cb_video_index_set_video_file(_ix, id, (dataPath + "/" + std::to_string(id) + ".vdx").c_str());
  }
  void setVideo(uint32_t id, const std::vector<int32_t>& frames, const std::vector<uint64_t>& hashes) {
    ok(cb_video_index_set_video(_ix, id, frames.data(), hashes.data(), int64_t(frames.size())), "DctVideoIndex::setVideo");
  }
  void save() {}
  void add(const std::vector<Media>& media) {
    std::vector<uint32_t> ids;
    for (const Media& m : media) ids.push_back(uint32_t(m.id));
    ok(cb_video_index_add(_ix, ids.data(), int64_t(ids.size())), "DctVideoIndex::add");
    for (const Media& m : media)
      if (!m.frames.empty()) setVideo(uint32_t(m.id), m.frames, m.hashes);
  }
  void remove(const std::vector<int>& ids) { ok(cb_video_index_remove(_ix, ids.data(), int64_t(ids.size())), "DctVideoIndex::remove"); }
  /// find(): image needle -> findFrame, video needle -> findVideo (:282-289)
  std::vector<Match> find(const Media& needle, const SearchParams& p) {
    const cb_params c = p.c();
    if (needle.type == Media::TypeImage)
      return detail::collect("DctVideoIndex::findFrame", [&](cb_match* out, int64_t cap, int64_t* n) {
        return cb_video_index_find_frame(_ix, needle.dctHash, needle.matchRangeDstIn, &c, out, cap, n);
      });
    if (needle.type == Media::TypeVideo)
      return detail::collect("DctVideoIndex::findVideo", [&](cb_match* out, int64_t cap, int64_t* n) {
        const bool own = !needle.frames.empty();
        return cb_video_index_find_video(_ix, own ? needle.frames.data() : nullptr, own ? needle.hashes.data() : nullptr,
                                         own ? int64_t(needle.frames.size()) : 0, uint32_t(needle.id), &c, out, cap, n);
      });
    return {};
  }
  DctVideoIndex* slice(const std::set<uint32_t>& mediaIds) const {
    std::vector<uint32_t> v(mediaIds.begin(), mediaIds.end());
    cb_video_index* h = cb_video_index_slice(_ix, v.data(), int64_t(v.size()));
    return h ? new DctVideoIndex(h) : nullptr;
  }
};

/// src/cvfeaturesindex.{h,cpp}
class CvFeaturesIndex {
  cb_orb_index* _ix;
  explicit CvFeaturesIndex(cb_orb_index* ix) : _ix(ix) {}

 public:
  CvFeaturesIndex() : _ix(cb_orb_index_create()) {}
  ~CvFeaturesIndex() { cb_orb_index_destroy(_ix); }
  CvFeaturesIndex(const CvFeaturesIndex&) = delete;
  CvFeaturesIndex& operator=(const CvFeaturesIndex&) = delete;
  int id() const { return SearchParams::AlgoCVFeatures; }
  cb_orb_index* handle() const { return _ix; }

  bool isLoaded() const { return cb_orb_index_is_loaded(_ix) != 0; }
  size_t memoryUsage() const { return cb_orb_index_memory_usage(_ix); }
  int count() const { return int(cb_orb_index_count(_ix)); }
  /// load(): media (ascending id) with their descriptor rows (rows x 32 bytes each)
  void load(const std::vector<Media>& media) { pack(media, true); }
  void add(const std::vector<Media>& media) { pack(media, false); }
  void loadCache(const std::string& cachePath) { ok(cb_orb_index_load_cache(_ix, cachePath.c_str()), "CvFeaturesIndex::loadIndex"); }
  void save(const std::string& cachePath) { ok(cb_orb_index_save_cache(_ix, cachePath.c_str()), "CvFeaturesIndex::saveIndex"); }
  void remove(const std::vector<int>& ids) { ok(cb_orb_index_remove(_ix, ids.data(), int64_t(ids.size())), "CvFeaturesIndex::remove"); }
  bool findIndexData(Media& m) const {
    int64_t n = 0;
    if (cb_orb_index_descriptors(_ix, uint32_t(m.id), nullptr, 0, &n) != CB_OK || n <= 0) return false;
    m.descriptors.resize(size_t(n) * 32);
    return cb_orb_index_descriptors(_ix, uint32_t(m.id), m.descriptors.data(), n, &n) == CB_OK;
  }
  std::vector<Match> find(const Media& needle, const SearchParams& p) {
    const cb_params c = p.c();
    return detail::collect("CvFeaturesIndex::find", [&](cb_match* out, int64_t cap, int64_t* n) {
      const int64_t rows = int64_t(needle.descriptors.size() / 32);
      return cb_orb_index_find(_ix, rows ? needle.descriptors.data() : nullptr, rows, uint32_t(needle.id), &c, out, cap, n);
    });
  }
  CvFeaturesIndex* slice(const std::set<uint32_t>& mediaIds) const {
    std::vector<uint32_t> v(mediaIds.begin(), mediaIds.end());
    cb_orb_index* h = cb_orb_index_slice(_ix, v.data(), int64_t(v.size()));
    return h ? new CvFeaturesIndex(h) : nullptr;
  }

 private:
  void pack(const std::vector<Media>& media, bool load) {
    std::vector<uint32_t> ids;
    std::vector<int64_t> offs{0};
    std::vector<uint8_t> flat;
    for (const Media& m : media) {
      ids.push_back(uint32_t(m.id));
      flat.insert(flat.end(), m.descriptors.begin(), m.descriptors.end());
      offs.push_back(int64_t(flat.size() / 32));
    }
    ok((load ? cb_orb_index_load : cb_orb_index_add)(_ix, ids.data(), offs.data(), flat.data(), int64_t(ids.size())),
       load ? "CvFeaturesIndex::load" : "CvFeaturesIndex::add");
  }
};

/// dctHash64(cvImg) for 8-bit gray images (src/cvutil.cpp:435-545); 0 on failure ("no hash")
inline uint64_t dctHash64(const uint8_t* gray, int cols, int rows, int64_t step) {
  uint64_t h = 0;
  return ok(cb_hash_batch(gray, 1, cols, rows, step, 0, &h), "dctHash64") ? h : 0;
}

/// dctHash64(cvImg) for decoded 8UC3 (BGR) / 8UC4 (BGRA) / 8UC1 images: grayscale() (:1265-1283) + hash.
/// grayMode: CB_GRAY_Q14 for a cbird built against OpenCV 2.4.x (the pinned 2.4.13.7), CB_GRAY_Q15 for 4.x.
inline uint64_t dctHash64(const uint8_t* pixels, int cols, int rows, int channels, int64_t step, int grayMode) {
  uint64_t h = 0;
  return ok(cb_hash_batch_color(pixels, 1, cols, rows, channels, step, 0, grayMode, &h), "dctHash64") ? h : 0;
}

/// cv::BFMatcher(NORM_HAMMING).radiusMatch(query, train, maxDistance) as TemplateMatcher::match uses it
/// (src/templatematcher.cpp:134-139,217-218): {a = trainIdx, b = queryIdx, dist <= maxDistance}, sorted by
/// (queryIdx, dist, trainIdx); empty on failure.
inline std::vector<cb_pair> radiusMatch(const uint8_t* query, int64_t nQuery, const uint8_t* train, int64_t nTrain,
                                        int maxDistance) {
  cb_pair* p = nullptr;
  int64_t n = 0;
  std::vector<cb_pair> out;
  if (ok(cb_orb_radius_match_alloc(train, nTrain, query, nQuery, maxDistance, &p, &n), "radiusMatch")) out.assign(p, p + n);
  cb_free(p);
  return out;
}

}  // namespace cbird_b200
