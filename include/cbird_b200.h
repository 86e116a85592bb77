/* cbird_b200 — C ABI of the B200-native hot path of cbird (scrubbbbs/cbird).
 *
 * This is the drop-in boundary: every entry point below is what a cbird build would bind in place of
 * the CPU code cited beside it (paths are relative to the cbird source tree).  Plain pointers and
 * sizes only; no C++ / torch / Qt types.  All functions are thread-safe per handle (cbird calls
 * Index::find concurrently from its Qt pool under a read lock, src/database.cpp:1400,1698) and never
 * abort: they return CB_OK (0) or a negative cb_status; cb_last_error() gives the message for the
 * calling thread.  There is NO CPU fallback: without a CUDA device every compute call fails with
 * CB_ERR_NO_DEVICE.
 *
 * Pointers named d_* are device pointers, everything else is host memory.
 */
#ifndef CBIRD_B200_H
#define CBIRD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cb_status {
  CB_OK = 0,
  CB_ERR_NO_DEVICE = -1,   /* no CUDA device / driver */
  CB_ERR_CUDA = -2,        /* a CUDA call failed, see cb_last_error() */
  CB_ERR_INVALID = -3,     /* bad argument */
  CB_ERR_CAPACITY = -4,    /* caller buffer too small; *n_out holds the required size */
  CB_ERR_UNSUPPORTED = -5, /* geometry / option outside the implemented range */
  CB_ERR_NOT_LOADED = -6   /* index not loaded (Index::isLoaded() false) */
} cb_status;

/* Index::Match (src/index.h:157-167) with MatchRange (src/media.h:62-78) flattened. */
typedef struct cb_match {
  uint32_t mediaId; /* unique id of indexed media */
  int32_t score;    /* lower is better */
  int32_t srcIn;    /* MatchRange: needle position   (-1 when unused) */
  int32_t dstIn;    /* MatchRange: matched position  (-1 when unused) */
  int32_t len;      /* MatchRange: length            ( 0 when unused) */
} cb_match;

/* The SearchParams fields the path reads (src/index.h:74-121); same names, same defaults. */
typedef struct cb_params {
  int32_t algo;             /* AlgoDCT=0, AlgoCVFeatures=2, AlgoVideo=4 (src/index.h:43-50) */
  int32_t dctThresh;        /* 5   match iff distance <  dctThresh (strict, src/tree/vptree.h:239) */
  int32_t cvThresh;         /* 25  match iff distance <  cvThresh  (src/cvfeaturesindex.cpp:511) */
  int32_t minMatches;       /* 1 */
  int32_t maxMatches;       /* 5   applied by the caller-side post step (src/database.cpp:1737) */
  int32_t skipFrames;       /* 300 vtrim */
  int32_t minFramesMatched; /* 30  vfm */
  int32_t minFramesNear;    /* 60  vfn */
  int32_t videoRadix;       /* 10  vradix; 0 = one bucket = exact */
  int32_t maxThresh;        /* 0 */
  uint8_t filterSelf;       /* 1 */
  uint8_t verbose;          /* 0 */
  uint8_t pad_[2];
  uint32_t target;          /* 0 */
} cb_params;

/* one radius-search hit of a batched query: (needle index, media id, distance) */
typedef struct cb_hit {
  uint32_t needle; /* index into the needle array of the call */
  uint32_t mediaId;
  int32_t score;
} cb_hit;

/* raw device-side hit of the scan kernels: (index on the A side, index on the B side, distance) */
typedef struct cb_pair {
  uint32_t a;
  uint32_t b;
  uint32_t dist;
  uint32_t pad_;
} cb_pair;

/* one work item of the tile-list scan: A rows [a_begin, a_begin+a_count), a_count <= 2048, against
 * B rows [b_begin, b_begin+b_count) */
typedef struct cb_scan_tile {
  uint32_t a_begin, a_count, b_begin, b_count;
} cb_scan_tile;

/* HammingTree_t::Match (src/tree/hammingtree.h:75-87) of a batched search */
typedef struct cb_tree_match {
  uint32_t needle;  /* index into the needle array of the call */
  uint32_t index;   /* Value::index (0 = removed) */
  int32_t distance;
  uint32_t pad_;
  uint64_t hash;    /* Value::hash */
} cb_tree_match;

typedef struct cb_stats {
  uint64_t comparisons;     /* pair tests issued by scan kernels since cb_stats_reset */
  uint64_t hits;            /* pairs under threshold */
  uint64_t kernel_launches; /* launches of this library's own kernels */
  uint64_t frames_hashed;
} cb_stats;

/* ---- library -------------------------------------------------------------------------------- */
const char* cb_version(void);
const char* cb_last_error(void);
/* select the CUDA device used by handles created afterwards on this thread (default 0) */
int cb_set_device(int device);
int cb_device_count(int* n_out);
void cb_params_default(cb_params* p);           /* SearchParams() defaults, src/index.h:74-121 */
int cb_stats_get(cb_stats* out);
void cb_stats_reset(void);
void cb_free(void* p);                           /* frees buffers returned by *_alloc calls */
/* ---- multi-GPU (one box, NVLink): the indexes of this process are replicated / sharded over several GPUs and
 * the sharded calls (cb_dct_index_load, cb_dct_index_similar_alloc) fan out over them, exchanging hit lists
 * with NCCL. Call once, before any index is created; without it every handle lives on the device selected
 * with cb_set_device. Where a cbird build would call it: Engine::Engine (src/engine.cpp:38-45), before the
 * Index objects are made.
 *   cb_init            this process drives all the listed devices (one host thread per device inside the calls)
 *   cb_comm_unique_id + cb_comm_init_rank   one process per GPU (e.g. torchrun): rank 0 makes the id, the
 *                      launcher's transport carries it to the others, every process then joins with its rank.
 *                      In this mode sharded results cover the calling rank's rows only (cb_dct_index_shard_rows)
 *                      and load/add/remove must be called with the same data by every rank. */
int cb_init(const int* devices, int n_devices);
int cb_comm_unique_id(uint8_t* id_out, int cap);  /* returns the id size (128) or a negative status */
int cb_comm_init_rank(const uint8_t* id, int id_bytes, int rank, int world, int device);
int cb_comm_info(int* world, int* n_local, int* first_rank);
/* rows [row_begin, row_end) of an n-row index whose results rank `rank` of `world` owns (host only, no device) */
int cb_comm_shard_rows(int64_t n, int rank, int world, int64_t* row_begin, int64_t* row_end);
void cb_shutdown(void);                          /* destroys the communicator; indexes must be destroyed first */

/* Measurement aid: CUDA-event timing of the library's dominant kernels on the streams they are launched on
 * (bench.py's roofline numbers). Disabled by default; when enabled every bracketed launch records two events,
 * nothing is synchronised until cb_profile_get. Slots: 0 mih_bucket_kernel, 1 radix sort of the bucket keys,
 * 2 radix sort of the hit keys, 3 scan64_kernel, 4 dct_hash32_kernel, 5 bucket-key generation, 6 gather of the hashes
 * into bucket order + bucket bounds, 7 searchIndex post step (count + scan). */
#define CB_PROFILE_SLOTS 8
typedef struct cb_profile {
  double ms[CB_PROFILE_SLOTS];        /* summed durations */
  uint64_t launches[CB_PROFILE_SLOTS];
} cb_profile;
void cb_profile_enable(int on);
int cb_profile_get(cb_profile* out, int reset); /* waits for the recorded events; call after synchronising */

/* ---- kernel (a): dctHash64, replaces src/cvutil.cpp:435-545 (called from src/scanner.cpp:862 and
 * src/media.cpp:996) — batched: n 8-bit luma frames of w x h, rows `row_stride` bytes apart, frames
 * `frame_stride` bytes apart. out[i] is never 0 on success. w,h >= 32 (smaller frames: CB_ERR_UNSUPPORTED); frames
 * of any larger size are accepted (tested up to 4000 x 3000). ------------------------------------------------ */
int cb_hash_batch(const uint8_t* frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride,
                  uint64_t* out);
/* device-resident variant on a caller stream (cudaStream_t passed as void*); asynchronous */
int cb_hash_batch_dev(const uint8_t* d_frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride,
                      uint64_t* d_out, void* stream);
/* grayscale() (src/cvutil.cpp:1265-1283), the first step of dctHash64 for decoded images (src/scanner.cpp:862):
 * cv::cvtColor(BGR2GRAY | BGRA2GRAY) of interleaved 8-bit pixels, B first.  channels: 1 (pass-through), 3, 4;
 * anything else is CB_ERR_UNSUPPORTED (the reference qFatal()s).  Strides are in BYTES.  OpenCV's fixed-point
 * weights changed between the release the reference pins and current ones, so the caller says which one the
 * database was written with (a binding passes CV_MAJOR_VERSION < 3 ? CB_GRAY_Q14 : CB_GRAY_Q15):
 *   CB_GRAY_Q14  OpenCV 2.4.x (cbird.pri:148 pins 2.4.13.7): (1868 B + 9617 G + 4899 R + 2^13) >> 14
 *   CB_GRAY_Q15  OpenCV 4.x: (3735 B + 19235 G + 9798 R + 2^14) >> 15  (pinned against cv2 4.13, all 2^24 colours)
 * cb_gray_batch writes dense n x h x w gray frames to host memory; cb_hash_batch_color is grayscale + dctHash64
 * without the gray frames leaving the device. */
enum { CB_GRAY_Q14 = 0, CB_GRAY_Q15 = 1 };
int cb_gray_batch(const uint8_t* frames, int64_t n, int w, int h, int channels, int64_t row_stride,
                  int64_t frame_stride, int gray_mode, uint8_t* out);
int cb_hash_batch_color(const uint8_t* frames, int64_t n, int w, int h, int channels, int64_t row_stride,
                        int64_t frame_stride, int gray_mode, uint64_t* out);

/* autocrop(img, range) of every frame — src/cvutil.cpp:1285-1401 (de-letterbox before hashing video
 * frames, src/media.cpp:963,994): rects[4*i..] = {left, top, right, bottom}, right/bottom exclusive;
 * the full frame when no crop applies. */
int cb_autocrop_batch(const uint8_t* frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride,
                      int range, int32_t* rects);
/* dctHash64 of each frame's rectangle taken as a VIEW into the frame (what autocrop leaves in cvImg):
 * blur size by the rectangle's area, blur border pixels come from the parent frame. Rectangles smaller
 * than 32 px on a side give hash 0 ("no hash"). */
int cb_hash_batch_rects(const uint8_t* frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride,
                        const int32_t* rects, uint64_t* out);
/* near-frame compression window of Media::makeVideoIndex (src/media.cpp:958-1031) over the hashes of
 * consecutive frames 0..n-1; threshold = IndexParams vht (8). out arrays need n+1 entries. Host only. */
int cb_video_compress(const uint64_t* hashes, int64_t n, int threshold, int32_t* out_frames, uint64_t* out_hashes,
                      int64_t* n_out);
/* Media::makeVideoIndex (src/media.cpp:925-1037) on one video's decoded luma frames (the decoder hands
 * 128x128 gray, src/scanner.cpp:1043-1048): autocrop(20) -> dctHash64 -> compression. Library-allocated
 * VideoIndex{frames, hashes} (cb_free). */
int cb_make_video_index_alloc(const uint8_t* frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride,
                              int threshold, int32_t** out_frames, uint64_t** out_hashes, int64_t* n_out);
/* the f32 DCT basis rows 0..8 (9*32 floats) and the 81-entry zig-zag table the kernel uses */
void cb_hash_tables(float* basis_9x32, int32_t* zigzag81);

/* ---- kernel (b): 64-bit Hamming radius scan, replaces hamm64 (src/hamm.h:24-26) driven by
 * VpTree::search (src/tree/vptree.h:50-69,228-255) / RadixMap_t::search (src/tree/radix.h:187-210).
 * Emits every (a,b) with popcount(A[a]^B[b]) < threshold and, when radix_bits>0, equal radix bucket
 * ((h>>1) & (2^radix_bits-1), src/tree/radix.h:135-141).  *d_count receives the total number of hits
 * (may exceed cap: only the first cap are stored, unordered); it ACCUMULATES, the caller zeroes it.
 * Asynchronous on `stream`. ---------------------------------------------------------------------- */
int cb_scan64_dev(const uint64_t* d_a, uint32_t n_a, const uint64_t* d_b, uint32_t n_b, int threshold,
                  int radix_bits, cb_pair* d_out, uint64_t cap, unsigned long long* d_count, void* stream);
/* `-similar` self-scan: needles = rows [0,n) of d_hashes, searched rows = [row_begin,row_end); indices
 * in the hits are absolute rows. symmetric=1 halves the pair tests: d(a,b)==d(b,a), so only 2048-row
 * tiles on/above the diagonal are tested, rows >= row_end are left to the ranks that own them, and every
 * off-diagonal hit is emitted as (a,b) and (b,a); the union over a tile-aligned partition of the rows is
 * exactly the full -similar hit set (row_begin must be a multiple of 2048). */
int cb_scan64_self_dev(const uint64_t* d_hashes, uint32_t n, uint32_t row_begin, uint32_t row_end, int threshold,
                       int symmetric, cb_pair* d_out, uint64_t cap, unsigned long long* d_count, void* stream);
/* same inner loop over an explicit list of (A block, B range) work items, one CTA each — the
 * radix-bucket search of DctVideoIndex (src/tree/radix.h:187-210 scans one bucket per needle frame) */
int cb_scan64_tiles_dev(const uint64_t* d_a, uint32_t n_a, const uint64_t* d_b, uint32_t n_b,
                        const cb_scan_tile* d_tiles, uint32_t n_tiles, int threshold, cb_pair* d_out, uint64_t cap,
                        unsigned long long* d_count, void* stream);
/* The same hit set as cb_scan64_self_dev(row_begin = 0, row_end = n) -- every ORDERED pair (a, b), a == b included,
 * with hamm64 < threshold -- found by multi-index hashing instead of testing all n^2 pairs: the 63 usable bits
 * (bit 0 of a dct hash is always clear, src/cvutil.cpp:537-538) are cut into `threshold` chunks; two hashes
 * closer than the threshold agree exactly on at least one chunk, so only rows sharing a chunk bucket are
 * compared. threshold in [1, cb_scan64_mih_max_threshold()], n * threshold < 2^32 - 65536. Multi-GPU: buckets are
 * dealt to `n_parts` ranks; rank `part` reports the pairs whose first shared bucket is its own, so the per-rank
 * lists are disjoint and their union is the full set. *d_count (not reset here) receives the total, also
 * beyond `cap`. Exact on any input; fastest when the hashes spread over the buckets. */
int cb_scan64_self_mih_dev(const uint64_t* d_hashes, uint32_t n, int threshold, uint32_t part, uint32_t n_parts,
                           cb_pair* d_out, uint64_t cap, unsigned long long* d_count, void* stream);
int cb_scan64_mih_max_threshold(void);
/* pair tests issued by the last cb_scan64_self_mih_dev of the calling thread (synchronises `stream`) */
int cb_scan64_mih_last_tests(void* stream, uint64_t* tests_out);
/* measurement / parity aid: force the pre-filter of the bucket scan (1 OR-fold, 2 AND-fold of two rows, 3 AND of
 * two OR-folds; 0 = automatic) and the number of chunks a bucket key is built from (1 or 2; 0 = automatic).
 * Every combination reports the identical hit set. */
void cb_scan64_mih_force(int variant, int need); /* need = -1: no multi-index pass at all (brute-force scan) */
/* what a self-join over n rows at this threshold would use: returns 1 and fills *variant / *need, or 0 when the
 * brute-force scan runs instead (threshold > 10, fewer than 2^15 rows) */
int cb_scan64_mih_config(uint64_t n, int threshold, int* variant, int* need);
/* the bucket layout the self-join uses for a threshold (host only, no device needed): chunk c of a hash is
 * (h >> shifts[c]) & masks[c]; returns the number of chunks (== threshold) or a negative status */
int cb_scan64_mih_plan(int threshold, int32_t* shifts, uint32_t* masks);
/* the two-chunk layout (threshold + 1 chunks, a bucket = the values of a PAIR of chunks): fills the chunk table
 * (threshold + 1 entries) and the units (pairs c1 < c2 in lexicographic order); returns the number of units
 * (threshold + 1) * threshold / 2. Host only. */
int cb_scan64_mih_plan2(int threshold, int32_t* shifts, uint32_t* masks, int32_t* unit_c1, int32_t* unit_c2);
/* variant index actually used for a threshold: 0 exact (2 POPC/pair), 1 OR-fold prefilter (1 POPC/pair),
 * 2 AND-fold prefilter (0.5 POPC/pair); all three produce identical hit sets */
int cb_scan64_variant(int threshold);
/* variant the last dense scan of this process really ran: for jobs over 2^34 pair tests the default is stepped down
 * (2 -> 1 -> 0) when a sample of the job's own pairs shows that the cheaper pre-filter lets too many pairs through
 * (clustered or correlated hashes); -1 before any scan */
int cb_scan64_last_variant(void);
/* force a variant (-1 = automatic); for measurements and parity tests */
void cb_scan64_force_variant(int variant);

/* ---- DctHashIndex (src/dcthashindex.{h,cpp}) ---------------------------------------------------- */
typedef struct cb_dct_index cb_dct_index;
cb_dct_index* cb_dct_index_create(void);                   /* DctHashIndex()            :40-43   */
void cb_dct_index_destroy(cb_dct_index* ix);               /* ~DctHashIndex/unload      :53-65   */
/* load(): the (id, phash_dct) rows the reference reads from SQL            :70-114  */
/* n < 2^31 - 4096. The rows live on the device; a host copy is made only by the calls that need one
 * (slice, mediaIds). With a communicator every rank uploads its share and an all-gather replicates it. */
int cb_dct_index_load(cb_dct_index* ix, const uint32_t* ids, const uint64_t* hashes, int64_t n);
int cb_dct_index_is_loaded(const cb_dct_index* ix);        /* isLoaded()                         */
int64_t cb_dct_index_count(const cb_dct_index* ix);        /* count()                            */
size_t cb_dct_index_memory_usage(const cb_dct_index* ix);  /* memoryUsage() = 12*count  :56-59   */
int cb_dct_index_add(cb_dct_index* ix, const uint32_t* ids, const uint64_t* hashes, int64_t n); /* :158-173 */
int cb_dct_index_remove(cb_dct_index* ix, const int32_t* ids, int64_t n);                       /* :175-191 */
/* slice(): new index holding only the given media ids, caller destroys     :222-250 */
cb_dct_index* cb_dct_index_slice(const cb_dct_index* ix, const uint32_t* ids, int64_t n);
/* mediaIds() for a loaded index: ids whose hash != 0                        :128-132 */
int cb_dct_index_media_ids(const cb_dct_index* ix, uint32_t* out, int64_t cap, int64_t* n_out);
/* find(): all rows with distance < p->dctThresh, ascending (score, mediaId); needle hash 0 -> none.
 * srcIn/dstIn/len are -1/-1/0.                                               :193-220
 * The reference calls find() from every thread of its pool at once (src/database.cpp:1400-1432 under the read
 * lock of :1698). Concurrent callers are combined: whoever arrives while no launch is in flight leads a batch of
 * up to 128 waiting needles through ONE kernel launch (needles travel as kernel arguments, hits come back
 * through mapped host memory, no copy or stream synchronisation), hands the results out and passes the lead on.
 * Results are identical to calling one at a time. */
int cb_dct_index_find(cb_dct_index* ix, uint64_t needle_hash, const cb_params* p, cb_match* out, int64_t cap,
                      int64_t* n_out);
/* launches and needles served by the find() queue so far */
int cb_dct_index_find_queue_stats(const cb_dct_index* ix, uint64_t* batches, uint64_t* needles);
/* N independent find() calls in one launch; hits sorted by (needle, score, mediaId). Returns a
 * library-allocated array in *out (release with cb_free). */
int cb_dct_index_find_batch_alloc(cb_dct_index* ix, const uint64_t* needle_hashes, int64_t n_needles,
                                  const cb_params* p, cb_hit** out, int64_t* n_out);
/* `-similar`: every indexed row is a needle (src/database.cpp:1400-1432), then the searchIndex post
 * step per needle (sort by score, drop self when filterSelf, cut at maxMatches; :1729-1737).
 * Result: CSR over rows — offsets[count+1] and hits (needle = row index). Library-allocated. */
int cb_dct_index_similar_alloc(cb_dct_index* ix, const cb_params* p, int64_t** offsets_out, cb_hit** hits_out,
                               int64_t* n_hits_out);
/* With a communicator (cb_init / cb_comm_init_rank) the pass is sharded: hashes replicated on every rank, the
 * bucket scans dealt to the ranks, every hit sent to the rank owning its needle row (NCCL all-to-all), sort and
 * post step per rank. cb_init: the CSR covers all rows, like the single-GPU call. cb_comm_init_rank: it covers
 * the calling rank's rows [row_begin, row_end) (offsets has row_end - row_begin + 1 entries starting at 0,
 * hits carry global row numbers as `needle`). */
int cb_dct_index_shard_rows(const cb_dct_index* ix, int64_t* row_begin, int64_t* row_end);
/* the same pass without the result copy: kept hits of this process's rows and the pair tests its ranks issued
 * (device-resident timing; the lists stay on the device) */
int cb_dct_index_similar_count(cb_dct_index* ix, const cb_params* p, int64_t* n_hits_out, uint64_t* pair_tests_out);
/* multi-GPU sharding hook: `-similar` needles are all rows, but only rows [row_begin,row_end) of the
 * index are searched (this rank's shard); hits carry GLOBAL row indices as `needle`, no post step.
 * Sorted by (needle, score, mediaId). */
int cb_dct_index_similar_shard_alloc(cb_dct_index* ix, const cb_params* p, int64_t row_begin, int64_t row_end,
                                     cb_hit** hits_out, int64_t* n_hits_out);

/* ---- DctVideoIndex (src/dctvideoindex.{h,cpp}) --------------------------------------------------
 * The per-video frame-hash tables the reference reads from "<dataPath>/<mediaId>.vdx"
 * (src/dctvideoindex.cpp:64-72) are handed over with cb_video_index_set_video (or read from .vdx files
 * with cb_vdx_load below).  The bucket "tree" is built lazily by the first search and rebuilt whenever
 * videoRadix / skipFrames / the contents change. */
typedef struct cb_video_index cb_video_index;
cb_video_index* cb_video_index_create(void);
void cb_video_index_destroy(cb_video_index* ix);
/* load(): media ids of type=video ordered by id; at most 2^24 are kept       :172-211 */
int cb_video_index_load(cb_video_index* ix, const uint32_t* ids, int64_t n);
/* VideoIndex{frames[], hashes[]} of one video (src/videoindex.h:45-46) */
int cb_video_index_set_video(cb_video_index* ix, uint32_t media_id, const int32_t* frames, const uint64_t* hashes,
                             int64_t n);
int cb_video_index_is_loaded(const cb_video_index* ix);
int64_t cb_video_index_count(const cb_video_index* ix);          /* count() = number of videos :49-53 */
size_t cb_video_index_memory_usage(const cb_video_index* ix);    /* 0 until the tree is built  :57-59 */
int cb_video_index_add(cb_video_index* ix, const uint32_t* ids, int64_t n);     /* :256-260 */
int cb_video_index_remove(cb_video_index* ix, const int32_t* ids, int64_t n);   /* :262-280 */
cb_video_index* cb_video_index_slice(const cb_video_index* ix, const uint32_t* ids, int64_t n); /* :389-397 */
/* findVideo(): needle = one video's (frames, hashes); frames==NULL && needle_id!=0 uses the stored
 * table of needle_id.  One match per similar video: score = 100 - percentNear, range = (srcIn, dstIn,
 * len), ascending mediaId.                                                     :399-657 */
int cb_video_index_find_video(cb_video_index* ix, const int32_t* frames, const uint64_t* hashes, int64_t n,
                              uint32_t needle_id, const cb_params* p, cb_match* out, int64_t cap, int64_t* n_out);
/* many needle videos in one launch: needle k owns frames/hashes[needle_offsets[k] .. needle_offsets[k+1]);
 * result CSR (result_offsets[n_needles+1], matches) is library-allocated (cb_free). */
int cb_video_index_find_videos_alloc(cb_video_index* ix, const int64_t* needle_offsets, const int32_t* frames,
                                     const uint64_t* hashes, const uint32_t* needle_ids, int64_t n_needles,
                                     const cb_params* p, int64_t** result_offsets, cb_match** matches,
                                     int64_t* n_matches);
/* findFrame(): needle = one image hash; nearest frame per video: score = distance,
 * range = (needle_dst_in<0 ? 0 : needle_dst_in, matched frame, 1), ascending video index.  :291-387 */
int cb_video_index_find_frame(cb_video_index* ix, uint64_t hash, int32_t needle_dst_in, const cb_params* p,
                              cb_match* out, int64_t cap, int64_t* n_out);

/* read "<dataPath>/<mediaId>.vdx" like insertHashes does (src/dctvideoindex.cpp:64-72) and hand the table
 * to the index; a missing or invalid file leaves the video without frames (warning semantics) */
int cb_video_index_set_video_file(cb_video_index* ix, uint32_t media_id, const char* vdx_path);

/* ---- .vdx codec (src/videoindex.cpp), host only ------------------------------------------------
 * decode: v1 or v2 by the "cbird" magic (getVersion :41-50), incl. the v1 repairs (:462-533); any
 * failure returns CB_ERR_INVALID with empty outputs (VideoIndex::load clears the table, :84-87).
 * encode: format v2 (save_v2 :271-349); frames[0] must be 0, frames strictly increasing. */
int cb_vdx_decode_alloc(const uint8_t* data, int64_t size, int32_t** frames, uint64_t** hashes, int64_t* n,
                        int* version);
int cb_vdx_encode_alloc(const int32_t* frames, const uint64_t* hashes, int64_t n, const char* writer_version,
                        uint8_t** data, int64_t* size);
int cb_vdx_is_valid(const uint8_t* data, int64_t size);                      /* isValid :90-102 */
int cb_vdx_load_alloc(const char* path, int32_t** frames, uint64_t** hashes, int64_t* n, int* version);
int cb_vdx_save(const char* path, const int32_t* frames, const uint64_t* hashes, int64_t n,
                const char* writer_version);

/* ---- HammingTree_t<uint32_t> (src/tree/hammingtree.h) — the approximate LSB-trie search that
 * DctFeaturesIndex uses (leaves of <= 8192 hashes; a needle is compared only with the leaf its low bits
 * select).  Same result sets as the reference, computed with the tile-list scan kernel. ------------ */
typedef struct cb_hamming_tree cb_hamming_tree;
cb_hamming_tree* cb_hamming_tree_create(void);
void cb_hamming_tree_destroy(cb_hamming_tree* t);
int cb_hamming_tree_insert(cb_hamming_tree* t, const uint32_t* indices, const uint64_t* hashes, int64_t n); /* :127-134 */
int cb_hamming_tree_remove(cb_hamming_tree* t, const uint32_t* indices, int64_t n);   /* index -> 0, :137-139 */
int cb_hamming_tree_stats(cb_hamming_tree* t, int32_t* num_nodes, int32_t* max_height, int64_t* num_values); /* :152-154 */
/* search() for every needle: matches sorted by (needle, distance, index, hash); library-allocated */
int cb_hamming_tree_search_batch_alloc(cb_hamming_tree* t, const uint64_t* needles, int64_t n_needles, int threshold,
                                       cb_tree_match** out, int64_t* n_out);
/* DctFeaturesIndex::find (src/dctfeaturesindex.cpp:260-358) over the tree: each needle hash votes for the
 * media (index) of its 10 nearest matches; score = -1 for the needle itself, 10*avg distance when nothing
 * has more than one vote, else maxVotes - votes; ascending mediaId. n == 0 && needle_id > 0: the needle's
 * hashes are taken from the tree. */
int cb_hamming_tree_find_votes(cb_hamming_tree* t, const uint64_t* needle_hashes, int64_t n, uint32_t needle_id,
                               int threshold, cb_match* out, int64_t cap, int64_t* n_out);
/* cache file format v2 (:156-200,:472-521), byte-compatible with the reference's reader/writer */
int cb_hamming_tree_write(cb_hamming_tree* t, const char* path);
int cb_hamming_tree_read(cb_hamming_tree* t, const char* path);

/* ---- CvFeaturesIndex (src/cvfeaturesindex.{h,cpp}): 256-bit ORB descriptors, kernel (c) ----------
 * Descriptors are rows of 32 bytes (cv::Mat N x 32 CV_8U); media m owns rows
 * [row_offsets[m], row_offsets[m+1]) of `desc`.  The reference's flann LSH knnSearch(k=10) (:497) is
 * replaced by an exact search; everything else (row->media maps, removal, threshold, median score)
 * follows the cited lines. */
typedef struct cb_orb_index cb_orb_index;
cb_orb_index* cb_orb_index_create(void);
void cb_orb_index_destroy(cb_orb_index* ix);
/* load(): media in ascending id order, empty or out-of-order entries are ignored      :167-250 */
int cb_orb_index_load(cb_orb_index* ix, const uint32_t* media_ids, const int64_t* row_offsets, const uint8_t* desc,
                      int64_t n_media);
int cb_orb_index_add(cb_orb_index* ix, const uint32_t* media_ids, const int64_t* row_offsets, const uint8_t* desc,
                     int64_t n_media);                                          /* :122-152 */
int cb_orb_index_remove(cb_orb_index* ix, const int32_t* ids, int64_t n);       /* :154-165 */
int cb_orb_index_is_loaded(const cb_orb_index* ix);                              /* :102     */
int64_t cb_orb_index_count(const cb_orb_index* ix);       /* count() = number of descriptors :104 */
size_t cb_orb_index_memory_usage(const cb_orb_index* ix); /* 2 x rows x 32            :106-120 */
cb_orb_index* cb_orb_index_slice(const cb_orb_index* ix, const uint32_t* ids, int64_t n);   /* :285-309 */
/* descriptorsForMediaId() / findIndexData(): rows of one media                     :421-436 */
int cb_orb_index_descriptors(const cb_orb_index* ix, uint32_t media_id, uint8_t* out, int64_t cap_rows,
                             int64_t* n_rows);
/* find(): needle = n_rows descriptors (or desc==NULL: the indexed descriptors of needle_id).
 * One match per media: score = median(distances) * 1000 / count, ascending mediaId.   :438-604 */
int cb_orb_index_find(cb_orb_index* ix, const uint8_t* desc, int64_t n_rows, uint32_t needle_id, const cb_params* p,
                      cb_match* out, int64_t cap, int64_t* n_out);
/* saveIndex()/loadIndex() cache files in `cache_dir` (cvfeatures.mat, cvfeatures_idmap.map,
 * cvfeatures_indexmap.map, cvfeatures.touch), byte-compatible with the reference    :387-419 */
int cb_orb_index_save_cache(cb_orb_index* ix, const char* cache_dir);
int cb_orb_index_load_cache(cb_orb_index* ix, const char* cache_dir);
/* the exact k nearest rows with distance < threshold of every needle row: cb_pair{a = row, b = needle
 * row, dist, pad_ = media id of the row (0 = removed)}, sorted by (needle row, dist, row); k<=0 = all.
 * Building block for sharded search: per-shard lists are merged by (dist, row) and cut at k. */
int cb_orb_index_knn_alloc(cb_orb_index* ix, const uint8_t* desc, int64_t n_rows, int k, int threshold,
                           cb_pair** out, int64_t* n_out);
/* TemplateMatcher's descriptor match (src/templatematcher.cpp:134-139,217-218):
 * cv::BFMatcher(NORM_HAMMING).radiusMatch(query, train, maxDistance) -- every (query row, train row) with
 * distance <= max_distance (OpenCV's radius is inclusive, unlike the strict `<` of find()).
 * cb_pair{a = trainIdx, b = queryIdx, dist}, sorted by (queryIdx, dist, trainIdx).  Stateless: both
 * descriptor sets (rows x 32 bytes, host memory) are given per call, as the reference rebuilds its
 * matcher per template. */
int cb_orb_radius_match_alloc(const uint8_t* train, int64_t n_train, const uint8_t* query, int64_t n_query,
                              int max_distance, cb_pair** out, int64_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* CBIRD_B200_H */
