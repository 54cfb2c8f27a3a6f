// Native tensoriser for `.c2v` text (reference path_context_reader.py:184-228 + filter :153-177):
// a chunk of complete lines -> int32 index rows + float32 mask rows written straight into
// caller-provided (pinned) host buffers, multi-threaded over lines, with the three vocabularies
// held in open-addressing hash tables.  Replaces tf.data's CsvDataset -> string_split ->
// StaticHashTable.lookup chain; string<->index work stays on the host, as in the reference.
//
// C ABI (bound with ctypes in path_context_reader.py):
//   c2v_vocab_create / c2v_vocab_destroy
//   c2v_parse_chunk
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

namespace {

inline uint64_t hash_bytes(const char* p, size_t n) {        // FNV-1a 64
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
  return h ? h : 1;
}

struct Vocab {
  std::vector<char> bytes;            // all words, concatenated
  struct Slot { uint64_t h; int64_t off; int32_t len; int32_t idx; };
  std::vector<Slot> slots;            // h == 0: empty
  uint64_t mask = 0;
  int32_t oov = 0, pad = 0;

  int32_t lookup(const char* p, size_t n) const {
    const uint64_t h = hash_bytes(p, n);
    for (uint64_t i = h & mask;; i = (i + 1) & mask) {
      const Slot& s = slots[i];
      if (s.h == 0) return oov;
      if (s.h == h && (size_t)s.len == n && memcmp(bytes.data() + s.off, p, n) == 0) return s.idx;
    }
  }
};

struct ParseJob {
  const char* text;
  const int64_t* line_off;     // [n_lines + 1]
  int64_t n_lines;
  int C;
  const Vocab *tok, *pth, *tgt;
  int32_t *src, *path, *dst, *target;
  float* mask;
  uint8_t* keep;               // per line: 1 = passes the row filter
  int64_t* tgt_off;            // per line: offset / length of the target field (for evaluation strings)
  int32_t* tgt_len;
  int mode;                    // 0 train, 1 evaluate, 2 predict (no filter)
  std::atomic<int64_t> bad_line{-1};
  std::atomic<int> bad_kind{0};
};

// One pending vocabulary lookup of a line: the bytes of a context part, its hash, and -- once probed --
// the slot that may hold it.
struct Part {
  const char* p;
  uint32_t len;
  int32_t result;              // filled in by resolve()
  uint64_t h;
  const Vocab::Slot* cand;     // first slot with the same hash and length (nullptr: an empty slot came first)
};

// Lookups are latency-bound (two dependent cache misses each: the slot, then the word's bytes; the tables
// of a java14m-sized vocabulary are ~100 MB), so a line's ~600 lookups are issued as a software pipeline:
// hash every part and prefetch its slot, then probe all slots and prefetch the candidates' bytes, then compare.
inline void hash_and_prefetch(const Vocab& v, Part& q) {
  q.h = hash_bytes(q.p, q.len);
  __builtin_prefetch(&v.slots[q.h & v.mask]);
}

inline void probe(const Vocab& v, Part& q) {
  q.cand = nullptr;
  for (uint64_t i = q.h & v.mask;; i = (i + 1) & v.mask) {
    const Vocab::Slot& s = v.slots[i];
    if (s.h == 0) return;
    if (s.h == q.h && (uint32_t)s.len == q.len) {
      q.cand = &s;
      __builtin_prefetch(v.bytes.data() + s.off);
      return;
    }
  }
}

inline void resolve(const Vocab& v, Part& q) {
  if (!q.cand) { q.result = v.oov; return; }
  if (memcmp(v.bytes.data() + q.cand->off, q.p, q.len) == 0) { q.result = q.cand->idx; return; }
  q.result = v.lookup(q.p, q.len);        // same hash and length, different word: take the plain path
}

// one line -> one row.  Returns 0 ok, 1 wrong field count, 2 context with more than 3 parts.
// `parts` is per-thread scratch with room for 3 * C entries.
int parse_line(const ParseJob& J, int64_t li, Part* parts) {
  const char* p = J.text + J.line_off[li];
  const char* end = J.text + J.line_off[li + 1];
  while (end > p && (end[-1] == '\n' || end[-1] == '\r')) --end;
  const int C = J.C;
  int32_t* src = J.src + li * C;
  int32_t* pth = J.path + li * C;
  int32_t* dst = J.dst + li * C;
  float* msk = J.mask + li * C;
  const int32_t tpad = J.tok->pad, ppad = J.pth->pad;
  // field 0: target name
  const char* f = p;
  const char* sp = (const char*)memchr(f, ' ', end - f);
  const char* fe = sp ? sp : end;
  J.tgt_off[li] = f - J.text;
  J.tgt_len[li] = (int32_t)(fe - f);
  const int32_t ty = (fe == f) ? J.tgt->oov : J.tgt->lookup(f, fe - f);
  J.target[li] = ty;
  // pass 1: split the contexts; a part that is absent keeps len == UINT32_MAX and resolves to PAD
  constexpr uint32_t kAbsent = UINT32_MAX;
  int nfields = 1;
  int c = 0;
  while (sp) {
    f = sp + 1;
    sp = (const char*)memchr(f, ' ', end - f);
    fe = sp ? sp : end;
    ++nfields;
    if (c >= C) continue;       // keep counting fields for the error check
    Part* q = parts + 3 * c;
    q[0].len = q[1].len = q[2].len = kAbsent;
    if (fe != f) {
      const char* c1 = (const char*)memchr(f, ',', fe - f);
      if (!c1) {
        q[0].p = f; q[0].len = (uint32_t)(fe - f);                      // missing parts stay PAD
      } else {
        q[0].p = f; q[0].len = (uint32_t)(c1 - f);
        const char* c2 = (const char*)memchr(c1 + 1, ',', fe - (c1 + 1));
        if (!c2) {
          q[1].p = c1 + 1; q[1].len = (uint32_t)(fe - (c1 + 1));
        } else {
          q[1].p = c1 + 1; q[1].len = (uint32_t)(c2 - (c1 + 1));
          if (memchr(c2 + 1, ',', fe - (c2 + 1))) return 2;
          q[2].p = c2 + 1; q[2].len = (uint32_t)(fe - (c2 + 1));
        }
      }
      if (q[0].len != kAbsent) hash_and_prefetch(*J.tok, q[0]);
      if (q[1].len != kAbsent) hash_and_prefetch(*J.pth, q[1]);
      if (q[2].len != kAbsent) hash_and_prefetch(*J.tok, q[2]);
    }
    ++c;
  }
  if (nfields != C + 1) return 1;
  // pass 2: probe the (prefetched) slots, prefetch the candidates' bytes; pass 3: compare
  for (int i = 0; i < c; ++i) {
    Part* q = parts + 3 * i;
    if (q[0].len != kAbsent) probe(*J.tok, q[0]);
    if (q[1].len != kAbsent) probe(*J.pth, q[1]);
    if (q[2].len != kAbsent) probe(*J.tok, q[2]);
  }
  int32_t max_s = INT32_MIN, max_t = INT32_MIN, max_p = INT32_MIN;
  for (int i = 0; i < c; ++i) {
    Part* q = parts + 3 * i;
    int32_t s = tpad, k = ppad, t = tpad;
    if (q[0].len != kAbsent) { resolve(*J.tok, q[0]); s = q[0].result; }
    if (q[1].len != kAbsent) { resolve(*J.pth, q[1]); k = q[1].result; }
    if (q[2].len != kAbsent) { resolve(*J.tok, q[2]); t = q[2].result; }
    src[i] = s; pth[i] = k; dst[i] = t;
    msk[i] = (s != tpad || t != tpad || k != ppad) ? 1.0f : 0.0f;
    if (s > max_s) max_s = s;
    if (t > max_t) max_t = t;
    if (k > max_p) max_p = k;
  }
  // row filter (path_context_reader.py:153-177): reduce_max(indices) != PAD index, target > OOV (train)
  const bool any_valid = (max_s != tpad) || (max_t != tpad) || (max_p != ppad);
  uint8_t keep = 1;
  if (J.mode == 0) keep = (any_valid && ty > J.tgt->oov) ? 1 : 0;
  else if (J.mode == 1) keep = any_valid ? 1 : 0;
  J.keep[li] = keep;
  return 0;
}

}  // namespace

extern "C" {

// words: concatenated bytes; offsets: [n + 1]; indices: [n].  Duplicate words: the last one wins
// (as a Python dict built in order would).
void* c2v_vocab_create(const char* words, const int64_t* offsets, const int32_t* indices, int64_t n, int32_t oov, int32_t pad) {
  Vocab* v = new Vocab();
  v->oov = oov; v->pad = pad;
  v->bytes.assign(words, words + offsets[n]);
  uint64_t cap = 16;
  while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
  v->slots.assign(cap, Vocab::Slot{0, 0, 0, 0});
  v->mask = cap - 1;
  for (int64_t i = 0; i < n; ++i) {
    const char* p = words + offsets[i];
    const size_t len = (size_t)(offsets[i + 1] - offsets[i]);
    const uint64_t h = hash_bytes(p, len);
    for (uint64_t j = h & v->mask;; j = (j + 1) & v->mask) {
      Vocab::Slot& s = v->slots[j];
      if (s.h == 0) { s = Vocab::Slot{h, offsets[i], (int32_t)len, indices[i]}; break; }
      if (s.h == h && (size_t)s.len == len && memcmp(v->bytes.data() + s.off, p, len) == 0) { s.idx = indices[i]; break; }
    }
  }
  return v;
}

void c2v_vocab_destroy(void* v) { delete (Vocab*)v; }

int32_t c2v_vocab_lookup(const void* v, const char* word, int64_t len) { return ((const Vocab*)v)->lookup(word, (size_t)len); }

// Parses the complete lines in text[0, len) (the last line may lack a trailing newline).
// Outputs hold one row per line, in line order: src/path/dst [n, C] int32, mask [n, C] float32,
// target [n] int32, keep [n] uint8 (row filter), tgt_off/tgt_len [n] (target field location).
// Returns the number of lines (<= capacity), or -(line number + 1) on a malformed line with
// *err_kind = 1 (field count) / 2 (context with > 3 parts), or INT64_MIN if capacity is too small.
// Blank lines are skipped (they still count in the reported line number).
int64_t c2v_parse_chunk(const char* text, int64_t len, int32_t max_contexts, const void* tok, const void* pth, const void* tgt,
                        int32_t mode, int32_t n_threads, int64_t capacity, int32_t* src, int32_t* path, int32_t* dst,
                        float* mask, int32_t* target, uint8_t* keep, int64_t* tgt_off, int32_t* tgt_len, int32_t* err_kind) {
  // blank lines ("" / "\n") are not records: the Python statement of the reader skips them too.  A record's span
  // runs to the next record's start, so blank lines after it are trimmed with its own newline.
  std::vector<int64_t> off, lineno;
  off.reserve(1024);
  lineno.reserve(1024);
  int64_t pos = 0, line = 0;
  while (pos < len) {
    const char* nl = (const char*)memchr(text + pos, '\n', len - pos);
    const int64_t next = nl ? (nl - text) + 1 : len;
    if (!(nl && nl == text + pos)) { off.push_back(pos); lineno.push_back(line); }
    pos = next;
    ++line;
  }
  off.push_back(len);
  const int64_t n = (int64_t)off.size() - 1;
  if (n > capacity) return INT64_MIN;
  ParseJob J;
  J.text = text; J.line_off = off.data(); J.n_lines = n; J.C = max_contexts;
  J.tok = (const Vocab*)tok; J.pth = (const Vocab*)pth; J.tgt = (const Vocab*)tgt;
  J.src = src; J.path = path; J.dst = dst; J.mask = mask; J.target = target; J.keep = keep;
  J.tgt_off = tgt_off; J.tgt_len = tgt_len; J.mode = mode;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  if ((int64_t)n_threads > n) n_threads = n > 0 ? (int)n : 1;
  auto work = [&](int t) {
    const int64_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
    std::vector<Part> parts((size_t)3 * (max_contexts > 0 ? max_contexts : 1));
    for (int64_t i = lo; i < hi; ++i) {
      const int rc = parse_line(J, i, parts.data());
      if (rc) {
        int64_t expect = -1;
        if (J.bad_line.compare_exchange_strong(expect, i)) J.bad_kind.store(rc);
        return;
      }
    }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  if (J.bad_line.load() >= 0) {
    if (err_kind) *err_kind = J.bad_kind.load();
    return -(lineno[(size_t)J.bad_line.load()] + 1);
  }
  return n;
}

}  // extern "C"
