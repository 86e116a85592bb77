// Compiles the header-only C++ host layer (include/cbird_b200.hpp) with plain g++ and drives the three
// indexes the way cbird's Database drives its Index plugins. Exit code 0 + "OK ..." on success.
// Without a CUDA device every call must fail softly (warning + empty result): prints "NO_DEVICE OK".
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "cbird_b200.hpp"

using namespace cbird_b200;

static int g_warnings = 0;

int main() {
  onWarning() = [](const std::string& s) {
    ++g_warnings;
    if (g_warnings <= 2) fprintf(stderr, "warning: %s\n", s.c_str());
  };
  int ndev = 0;
  const bool have_gpu = cb_device_count(&ndev) == CB_OK && ndev > 0;

  std::mt19937_64 rng(7);
  const int n = 5000;
  std::vector<uint32_t> ids(n);
  std::vector<uint64_t> hashes(n);
  for (int i = 0; i < n; ++i) {
    ids[i] = uint32_t(i + 1);
    hashes[i] = rng() & ~1ull;
    if (i >= 4000) hashes[i] = hashes[rng() % 4000] ^ (1ull << (1 + rng() % 63)) ^ (1ull << (1 + rng() % 63));
  }
  DctHashIndex dct;
  assert(!dct.isLoaded() && dct.count() == 0 && dct.memoryUsage() == 0);  // baseTestDefaults
  dct.load(ids, hashes);

  if (!have_gpu) {
    SearchParams p;
    Media needle;
    needle.dctHash = hashes[0];
    assert(!dct.isLoaded());
    assert(dct.find(needle, p).empty());
    assert(dctHash64(reinterpret_cast<const uint8_t*>(hashes.data()), 32, 32, 32) == 0);
    assert(g_warnings >= 2);
    printf("NO_DEVICE OK (%d soft errors)\n", g_warnings);
    return 0;
  }

  assert(dct.isLoaded() && dct.count() == n && dct.memoryUsage() == size_t(12) * n);
  SearchParams p;
  p.filterSelf = false;
  long total = 0;
  for (int row : {0, 17, 4321, 4999}) {
    Media needle;
    needle.id = int(ids[row]);
    needle.dctHash = hashes[row];
    std::vector<Match> got = dct.find(needle, p);
    std::vector<std::pair<int, uint32_t>> want;
    for (int j = 0; j < n; ++j) {
      const int d = __builtin_popcountll(hashes[row] ^ hashes[j]);
      if (d < p.dctThresh) want.push_back({d, ids[j]});
    }
    std::sort(want.begin(), want.end());
    assert(got.size() == want.size());
    for (size_t k = 0; k < got.size(); ++k) assert(got[k].score == want[k].first && got[k].mediaId == want[k].second);
    total += long(got.size());
  }
  p.filterSelf = true;
  p.maxMatches = 3;
  std::vector<std::vector<Match>> groups = dct.similar(p);
  assert(int(groups.size()) == n);
  long grouped = 0;
  for (int row = 0; row < n; ++row) {
    assert(groups[row].size() <= 3);
    for (const Match& m : groups[row]) assert(m.mediaId != ids[row] && m.score < 5);
    grouped += long(groups[row].size());
  }
  assert(grouped > 500);
  std::unique_ptr<DctHashIndex> chunk(dct.slice({1u, 2u, 3u, 4500u}));
  assert(chunk && chunk->count() == 4);
  dct.remove({int(ids[17])});
  Media m17;
  m17.dctHash = hashes[17];
  for (const Match& m : dct.find(m17, p)) assert(m.mediaId != ids[17]);

  // video: every video matches exactly itself with the reference unit-test parameters
  DctVideoIndex vid;
  std::vector<uint32_t> vids{11, 12, 13};
  vid.load(vids);
  std::vector<std::vector<int32_t>> frames(3);
  std::vector<std::vector<uint64_t>> vh(3);
  for (int v = 0; v < 3; ++v) {
    uint64_t h = rng() & ~1ull;
    for (int f = 0; f < 200; ++f) {
      h ^= 1ull << (1 + rng() % 63);
      frames[v].push_back(f * 7);
      vh[v].push_back(h);
    }
    vid.setVideo(vids[v], frames[v], vh[v]);
  }
  SearchParams vp;
  vp.algo = SearchParams::AlgoVideo;
  vp.filterSelf = false;
  vp.dctThresh = 1;
  vp.minFramesMatched = 1;
  vp.minFramesNear = 1;
  vp.skipFrames = 0;
  vp.videoRadix = 0;
  for (int v = 0; v < 3; ++v) {
    Media needle;
    needle.type = Media::TypeVideo;
    needle.frames = frames[v];
    needle.hashes = vh[v];
    std::vector<Match> got = vid.find(needle, vp);
    assert(got.size() == 1 && got[0].mediaId == vids[v] && got[0].range.srcIn == 0 && got[0].range.dstIn == 0);
  }

  // orb: an indexed image finds itself with score 0
  CvFeaturesIndex orb;
  std::vector<Media> media(20);
  for (int i = 0; i < 20; ++i) {
    media[i].id = 100 + i;
    media[i].descriptors.resize(50 * 32);
    for (auto& b : media[i].descriptors) b = uint8_t(rng());
  }
  orb.load(media);
  assert(orb.isLoaded() && orb.count() == 1000 && orb.memoryUsage() == 64000);
  Media q;
  q.descriptors = media[7].descriptors;
  std::vector<Match> om = orb.find(q, SearchParams());
  assert(om.size() == 1 && om[0].mediaId == 107 && om[0].score == 0);
  Media stored;
  stored.id = 107;
  assert(orb.findIndexData(stored) && stored.descriptors == media[7].descriptors);

  // dctHash64: constant frame -> 1 (src/cvutil.cpp:542)
  std::vector<uint8_t> flat(64 * 64, 90);
  assert(dctHash64(flat.data(), 64, 64, 64) == 1);
  // a gray BGR image hashes like its gray plane (all three weights sum to 1 in both fixed-point variants)
  std::vector<uint8_t> gray(64 * 64), bgr(64 * 64 * 3);
  for (int i = 0; i < 64 * 64; ++i) {
    gray[i] = uint8_t((i * 7 + (i >> 6) * 13) & 255);
    bgr[3 * i] = bgr[3 * i + 1] = bgr[3 * i + 2] = gray[i];
  }
  const uint64_t hg = dctHash64(gray.data(), 64, 64, 64);
  assert(hg != 0 && dctHash64(bgr.data(), 64, 64, 3, 64 * 3, CB_GRAY_Q15) == hg);
  assert(dctHash64(bgr.data(), 64, 64, 3, 64 * 3, CB_GRAY_Q14) == hg);
  // TemplateMatcher's radius match: media[7]'s descriptors against themselves at radius 0 -> the diagonal
  const std::vector<uint8_t>& d7 = media[7].descriptors;
  std::vector<cb_pair> rm = radiusMatch(d7.data(), int64_t(d7.size() / 32), d7.data(), int64_t(d7.size() / 32), 0);
  assert(rm.size() >= d7.size() / 32);
  for (const cb_pair& pr : rm) assert(pr.dist == 0);
  printf("OK %ld matches, %ld grouped, %d warnings\n", total, grouped, g_warnings);
  return 0;
}
