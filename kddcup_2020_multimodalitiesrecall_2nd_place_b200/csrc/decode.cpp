// Host-side record decoder for the competition TSV files: what the reference does per line
// in Python -- split on tabs, base64-decode boxes / 2048-d features / class labels, pad to the box budget
// (imagebert_zk/load_data_v4.py:133-163 + seq_padding_2 at :91-102 and :380-383; lxmert/src/utils.py:23-36) -- done
// by a pool of C++ threads straight into caller-owned (pinned) batch arrays, so that the host can feed the GPU scorer
// (>60 k pairs/s = 18 GB/s of fp32 features) instead of ~2 k lines/s of numpy work.  No GPU code in this file.
//
// Columns of a line: product_id, image_h, image_w, num_boxes, b64(boxes f32 [nb,4]), b64(features f32 [nb,2048]),
// b64(class_labels i64 [nb]), query, query_id.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <unistd.h>

#include "../../include/mmrecall.h"

namespace mmr {
mmr_status fail(mmr_status code, const char* fmt, ...);   // common.cu
}

namespace {

// Four pre-shifted lookup tables: a valid quantum is t0[a] | t1[b] | t2[c] | t3[d] (24 bits); any invalid character
// sets bit 31.
struct B64Tables {
  uint32_t t[4][256];
  B64Tables() {
    const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int k = 0; k < 4; ++k)
      for (int i = 0; i < 256; ++i) t[k][i] = 0x80000000u;
    for (uint32_t i = 0; i < 64; ++i) {
      t[0][uint8_t(a[i])] = i << 18;
      t[1][uint8_t(a[i])] = i << 12;
      t[2][uint8_t(a[i])] = i << 6;
      t[3][uint8_t(a[i])] = i;
    }
  }
};
const B64Tables kB64;

// Decodes [p, p+n) into dst (capacity cap bytes); returns the number of bytes written or -1 on a malformed field.
// Same acceptance as base64.b64decode on well-formed input: '=' padding closes the field.
long b64_decode(const char* p, size_t n, uint8_t* dst, size_t cap) {
  while (n > 0 && p[n - 1] == '=') --n;
  const size_t full = n / 4, rem = n % 4;
  if (rem == 1) return -1;
  const size_t out = full * 3 + (rem ? rem - 1 : 0);
  if (out > cap) return -1;
  const uint8_t* q = reinterpret_cast<const uint8_t*>(p);
  uint32_t bad = 0;
  size_t i = 0;
  // bulk: 4 characters -> one 32-bit big-endian store of which 3 bytes count (the 4th is overwritten by the next
  // quantum); the last quantum is written byte by byte so that nothing lands past `out`
  for (; i + 1 < full; ++i) {
    const uint32_t w = kB64.t[0][q[4 * i]] | kB64.t[1][q[4 * i + 1]] | kB64.t[2][q[4 * i + 2]] | kB64.t[3][q[4 * i + 3]];
    bad |= w;
    const uint32_t be = __builtin_bswap32(w << 8);
    memcpy(dst + 3 * i, &be, 4);
  }
  for (; i < full; ++i) {
    const uint32_t w = kB64.t[0][q[4 * i]] | kB64.t[1][q[4 * i + 1]] | kB64.t[2][q[4 * i + 2]] | kB64.t[3][q[4 * i + 3]];
    bad |= w;
    dst[3 * i] = uint8_t(w >> 16);
    dst[3 * i + 1] = uint8_t(w >> 8);
    dst[3 * i + 2] = uint8_t(w);
  }
  if (rem) {
    uint32_t w = kB64.t[0][q[4 * full]] | kB64.t[1][q[4 * full + 1]];
    if (rem == 3) w |= kB64.t[2][q[4 * full + 2]];
    bad |= w;
    dst[3 * full] = uint8_t(w >> 16);
    if (rem == 3) dst[3 * full + 1] = uint8_t(w >> 8);
  }
  return (bad & 0x80000000u) ? -1 : long(out);
}

bool parse_i64(const char* p, size_t n, int64_t* out) {
  if (n == 0 || n > 20) return false;
  char buf[24];
  memcpy(buf, p, n);
  buf[n] = 0;
  char* end = nullptr;
  const long long v = strtoll(buf, &end, 10);
  if (end != buf + n) return false;
  *out = v;
  return true;
}

// error codes stored per line
enum { kOk = 0, kColumns = 1, kNumber = 2, kBase64 = 3, kTooManyBoxes = 4, kQueryOverflow = 5 };

// `dirty` (or null): in/out count of leading box slots of record i that may hold non-zero data from an earlier decode
// into the same arrays; with it, only the slots [keep, dirty) are zeroed instead of all of [keep, R).
int decode_line(const char* line, size_t len, const mmr_decode_out& o, int64_t i, std::atomic<size_t>& qcursor,
                int32_t* dirty) {
  while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r' || line[len - 1] == ' ')) --len;   // .strip()
  while (len > 0 && (*line == ' ' || *line == '\n')) { ++line; --len; }
  const char* f[9];
  size_t fl[9];
  int nf = 0;
  const char* s = line;
  for (size_t k = 0; k <= len; ++k) {
    if (k == len || line[k] == '\t') {
      if (nf < 9) { f[nf] = s; fl[nf] = size_t(line + k - s); }
      ++nf;
      s = line + k + 1;
    }
  }
  if (nf < 9) return kColumns;
  int64_t pid, h, w, nb, qid;
  if (!parse_i64(f[0], fl[0], &pid) || !parse_i64(f[1], fl[1], &h) || !parse_i64(f[2], fl[2], &w) ||
      !parse_i64(f[3], fl[3], &nb) || !parse_i64(f[8], fl[8], &qid))
    return kNumber;
  if (nb < 0) return kNumber;
  const int R = o.max_boxes, F = o.feat_dim;
  // the reference truncates to the box budget AFTER decoding (seq_padding_2: x[:maxlen]); decode straight into place
  // when everything fits, else through a scratch buffer
  const int keep = int(std::min<int64_t>(nb, R));
  if (dirty && dirty[i] >= 0 && dirty[i] < keep) dirty[i] = keep;   // (stays valid if this record fails half-way)
  float* boxes = o.boxes4 + size_t(i) * R * 4;
  float* feats = o.feats + size_t(i) * R * F;
  int64_t* labels = o.class_labels + size_t(i) * R;
  if (nb <= R) {
    if (b64_decode(f[4], fl[4], reinterpret_cast<uint8_t*>(boxes), size_t(nb) * 16) != long(nb) * 16) return kBase64;
    if (b64_decode(f[5], fl[5], reinterpret_cast<uint8_t*>(feats), size_t(nb) * F * 4) != long(nb) * F * 4) return kBase64;
    if (b64_decode(f[6], fl[6], reinterpret_cast<uint8_t*>(labels), size_t(nb) * 8) != long(nb) * 8) return kBase64;
  } else {
    if (nb > 4096) return kTooManyBoxes;
    std::vector<uint8_t> tmp(size_t(nb) * F * 4);
    if (b64_decode(f[4], fl[4], tmp.data(), size_t(nb) * 16) != long(nb) * 16) return kBase64;
    memcpy(boxes, tmp.data(), size_t(keep) * 16);
    if (b64_decode(f[5], fl[5], tmp.data(), size_t(nb) * F * 4) != long(nb) * F * 4) return kBase64;
    memcpy(feats, tmp.data(), size_t(keep) * F * 4);
    if (b64_decode(f[6], fl[6], tmp.data(), size_t(nb) * 8) != long(nb) * 8) return kBase64;
    memcpy(labels, tmp.data(), size_t(keep) * 8);
  }
  // zero padding of the unused box slots (seq_padding_2(..., padding_value=0)); at 36 slots of 8 KB and 4 boxes per
  // record the padding is 90 % of the bytes a record occupies, so a reused array is only cleared where it is not zero
  const int stale = dirty ? std::max(keep, std::min<int>(R, dirty[i] < 0 ? R : dirty[i])) : R;
  memset(boxes + size_t(keep) * 4, 0, size_t(stale - keep) * 16);
  memset(feats + size_t(keep) * F, 0, size_t(stale - keep) * F * 4);
  memset(labels + keep, 0, size_t(stale - keep) * 8);
  if (dirty) dirty[i] = keep;
  o.product_id[i] = pid;
  o.image_h[i] = int32_t(h);
  o.image_w[i] = int32_t(w);
  o.num_boxes[i] = int32_t(nb);
  o.query_id[i] = qid;
  const size_t off = qcursor.fetch_add(fl[7]);
  if (off + fl[7] > o.query_cap) return kQueryOverflow;
  memcpy(o.query_text + off, f[7], fl[7]);
  o.query_off[2 * i] = int64_t(off);
  o.query_off[2 * i + 1] = int64_t(fl[7]);
  return kOk;
}



// Worker threads that outlive a call: a 256-line batch decodes in about a millisecond on 16 threads, which is what
// creating and joining 15 threads costs.  One job at a time (calls are serialised); a forked child starts its own pool.
class DecodePool {
 public:
  void run(int helpers, const std::function<void()>& job) {
    std::lock_guard<std::mutex> serial(run_mu_);
    std::unique_lock<std::mutex> lk(mu_);
    if (pid_ != getpid()) {            // after fork() the parent's threads do not exist here: forget their handles
      threads_ = new std::vector<std::thread>();
      pid_ = getpid();
    }
    while (int(threads_->size()) < helpers) threads_->emplace_back([this, id = int(threads_->size())] { loop(id); });
    job_ = &job;
    want_ = helpers;
    pending_ = helpers;
    ++gen_;
    lk.unlock();
    cv_work_.notify_all();
    job();
    lk.lock();
    cv_done_.wait(lk, [this] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  void loop(int id) {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_work_.wait(lk, [&] { return stop_ || gen_ != seen; });
      if (stop_) return;
      seen = gen_;
      if (id >= want_) continue;       // this job asked for fewer threads
      const std::function<void()>* job = job_;
      lk.unlock();
      (*job)();
      lk.lock();
      if (--pending_ == 0) cv_done_.notify_one();
    }
  }
  std::mutex run_mu_, mu_;
  std::condition_variable cv_work_, cv_done_;
  std::vector<std::thread>* threads_ = new std::vector<std::thread>();
  const std::function<void()>* job_ = nullptr;
  uint64_t gen_ = 0;
  int want_ = 0, pending_ = 0;
  bool stop_ = false;
  pid_t pid_ = getpid();
};

DecodePool& decode_pool() {
  static DecodePool* pool = new DecodePool();    // never destroyed: its threads sleep on a condition variable until exit()
  return *pool;
}

mmr_status decode_tsv(const char* const* lines, const size_t* line_len, int64_t n_lines, const mmr_decode_out* out,
                      int32_t* dirty, int n_threads) {
  if (!lines || !line_len || !out || n_lines < 0)
    return mmr::fail(MMR_ERR_INVALID, "mmr_decode_tsv: null argument");
  const mmr_decode_out& o = *out;
  if (!o.product_id || !o.image_h || !o.image_w || !o.num_boxes || !o.boxes4 || !o.feats || !o.class_labels ||
      !o.query_id || !o.query_off || !o.query_text || o.max_boxes <= 0 || o.feat_dim <= 0)
    return mmr::fail(MMR_ERR_INVALID, "mmr_decode_tsv: incomplete output descriptor");
  if (n_threads <= 0) n_threads = int(std::max(1u, std::thread::hardware_concurrency()));
  n_threads = int(std::min<int64_t>(std::min(n_threads, 256), std::max<int64_t>(n_lines, 1)));
  std::atomic<int64_t> next(0);
  std::atomic<size_t> qcursor(0);
  std::atomic<int64_t> first_bad(-1);
  std::atomic<int> bad_code(0);
  const std::function<void()> work = [&]() {
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n_lines) return;
      const int rc = decode_line(lines[i], line_len[i], o, i, qcursor, dirty);
      if (rc != kOk) {
        int64_t expect = -1;
        if (first_bad.compare_exchange_strong(expect, i)) bad_code.store(rc);
      }
    }
  };
  if (n_threads > 1) decode_pool().run(n_threads - 1, work); else work();
  if (first_bad.load() >= 0) {
    static const char* what[] = {"", "fewer than 9 tab-separated columns", "malformed integer field",
                                 "malformed or wrong-sized base64 field", "more than 4096 boxes",
                                 "query text buffer too small"};
    return mmr::fail(MMR_ERR_INVALID, "mmr_decode_tsv: line %lld: %s", (long long)first_bad.load(), what[bad_code.load()]);
  }
  return MMR_OK;
}

}  // namespace

extern "C" mmr_status mmr_decode_tsv(const char* const* lines, const size_t* line_len, int64_t n_lines,
                                     const mmr_decode_out* out, int n_threads) {
  return decode_tsv(lines, line_len, n_lines, out, nullptr, n_threads);
}

extern "C" mmr_status mmr_decode_tsv_reuse(const char* const* lines, const size_t* line_len, int64_t n_lines,
                                           const mmr_decode_out* out, int32_t* dirty_boxes, int n_threads) {
  if (!dirty_boxes) return mmr::fail(MMR_ERR_INVALID, "mmr_decode_tsv_reuse: null dirty_boxes");
  return decode_tsv(lines, line_len, n_lines, out, dirty_boxes, n_threads);
}

// CRC-32C (Castagnoli) of a host buffer, continuing from `crc` (0 to start): the checksum TensorFlow checkpoints carry
// per index block and per tensor (tf_bundle.py verifies / writes them; evaluate_normal.py:204-212 restores such files).
// Host utility, no GPU: hardware crc32 instruction where the compiler targets SSE4.2, slicing-by-8 tables otherwise.
#if defined(__SSE4_2__)
#include <nmmintrin.h>
#endif
extern "C" uint32_t mmr_crc32c(const void* data, size_t n, uint32_t crc) {
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~crc;
#if defined(__SSE4_2__)
  uint64_t c64 = c;
  while (n >= 8) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    c64 = _mm_crc32_u64(c64, v);
    p += 8;
    n -= 8;
  }
  c = static_cast<uint32_t>(c64);
  while (n--) c = _mm_crc32_u8(c, *p++);
#else
  static uint32_t tab[8][256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t r = i;
      for (int k = 0; k < 8; ++k) r = (r & 1u) ? (r >> 1) ^ 0x82F63B78u : r >> 1;
      tab[0][i] = r;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) tab[t][i] = (tab[t - 1][i] >> 8) ^ tab[0][tab[t - 1][i] & 0xffu];
    init = true;
  }
  while (n >= 8) {
    uint32_t lo, hi;
    std::memcpy(&lo, p, 4);
    std::memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = tab[7][lo & 0xffu] ^ tab[6][(lo >> 8) & 0xffu] ^ tab[5][(lo >> 16) & 0xffu] ^ tab[4][lo >> 24] ^
        tab[3][hi & 0xffu] ^ tab[2][(hi >> 8) & 0xffu] ^ tab[1][(hi >> 16) & 0xffu] ^ tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = tab[0][(c ^ *p++) & 0xffu] ^ (c >> 8);
#endif
  return ~c;
}
