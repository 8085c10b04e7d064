// Host input pipeline, first stage (no GPU work in this file): a data file tokenised into the flat columns the batcher consumes.
// Replaces, for a whole file at once, the reference's per-line Python parser
//   io/sequential_iterator.py:195-268  parser_one_line (train line: 6 tab-separated columns; eval line: 11)
//   io/sequential_iterator.py:270-322  the comma-separated history columns, vocabulary lookup with 0 for unknown tokens,
//                                      play time ms -> s
//   io/sequential_iterator.py:324-332  parse_file
// Bit-exactness is by construction on a STRICT subset of what Python accepts, and by refusal elsewhere: plain decimal tokens are
// converted with std::from_chars (correctly rounded, like Python's float()); anything Python might treat differently - non-ASCII
// bytes, a lone '\r', whitespace or '_' inside a numeric token, inf/nan/hex spellings, out-of-range values, ragged or missing
// columns - makes the call return PAMREC_TOK_FALLBACK, and the caller runs the Python parser, which then also raises the
// reference's own exceptions for malformed lines.
#include <charconv>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#include "../../include/pamrec_b200.h"

namespace {

inline bool py_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }   // str.strip(), ASCII

std::string_view strip(std::string_view s) {
  size_t a = 0, b = s.size();
  while (a < b && py_space((unsigned char)s[a])) ++a;
  while (b > a && py_space((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

// token -> index, open addressing over views into the caller's key bytes
struct Vocab {
  std::vector<int64_t> slot;     // key index + 1, 0 = empty
  const PamrecVocab* v = nullptr;
  uint64_t mask = 0;
  static uint64_t hash(std::string_view s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h ^ (h >> 29);
  }
  std::string_view key(int64_t i) const { return std::string_view(v->bytes + v->offsets[i], (size_t)(v->offsets[i + 1] - v->offsets[i])); }
  void build(const PamrecVocab* vocab) {
    v = vocab;
    uint64_t cap = 16;
    while (cap < (uint64_t)(2 * v->n + 1)) cap <<= 1;
    slot.assign(cap, 0);
    mask = cap - 1;
    for (int64_t i = 0; i < v->n; ++i) {
      uint64_t p = hash(key(i)) & mask;
      while (slot[p]) p = (p + 1) & mask;
      slot[p] = i + 1;
    }
  }
  int32_t get(std::string_view s) const {                       // dict.get(token, 0)
    uint64_t p = hash(s) & mask;
    while (slot[p]) {
      if (key(slot[p] - 1) == s) return v->values[slot[p] - 1];
      p = (p + 1) & mask;
    }
    return 0;
  }
};

// [+-]? ( digits [. digits*] | . digits ) ( [eE] [+-]? digits )?   ->   correctly rounded double; false = leave it to Python
bool parse_float(std::string_view s, double* out) {
  size_t i = 0, n = s.size();
  bool neg = false;
  if (i < n && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; ++i; }
  const size_t start = i;
  size_t nd = 0, nfrac = 0;
  uint64_t mant = 0;
  while (i < n && s[i] >= '0' && s[i] <= '9') { mant = mant * 10 + (uint64_t)(s[i] - '0'); ++i; ++nd; }
  if (i < n && s[i] == '.') { ++i; while (i < n && s[i] >= '0' && s[i] <= '9') { mant = mant * 10 + (uint64_t)(s[i] - '0'); ++i; ++nd; ++nfrac; } }
  if (nd == 0) return false;
  if (i == n && nd <= 15) {
    // at most 15 digits and no exponent: the digits are an exact integer below 2^53 and 10^nfrac is exact, so ONE correctly
    // rounded division gives the correctly rounded value of the decimal (Clinger's fast path) - what float() returns
    static const double p10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    const double v = (double)mant / p10[nfrac];
    *out = neg ? -v : v;
    return true;
  }
  if (i < n && (s[i] == 'e' || s[i] == 'E')) {
    ++i;
    if (i < n && (s[i] == '+' || s[i] == '-')) ++i;
    size_t ne = 0;
    while (i < n && s[i] >= '0' && s[i] <= '9') { ++i; ++ne; }
    if (ne == 0) return false;
  }
  if (i != n) return false;
  double v = 0;
  auto r = std::from_chars(s.data() + start, s.data() + n, v, std::chars_format::general);
  if (r.ec != std::errc() || r.ptr != s.data() + n) return false;
  *out = neg ? -v : v;
  return true;
}

bool parse_int(std::string_view s, double* out) {              // int(token) for labels
  size_t i = 0, n = s.size();
  bool neg = false;
  if (i < n && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; ++i; }
  if (i == n || n - i > 15) return false;
  int64_t v = 0;
  for (; i < n; ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (s[i] - '0');
  }
  *out = (double)(neg ? -v : v);
  return true;
}

struct Part {                                                  // what one thread produced, in line order
  std::vector<int64_t> lens;
  std::vector<int32_t> items, cates, user_ids, tgt_item, tgt_cate;
  std::vector<double> durs, sats, plays, label_sat, label_play, tgt_dur;
  bool ok = true;
};

template <typename F>
bool for_tokens(std::string_view col, F f) {                   // column.strip().split(",")
  col = strip(col);
  size_t a = 0;
  for (;;) {
    size_t b = col.find(',', a);
    const std::string_view tok = col.substr(a, b == std::string_view::npos ? std::string_view::npos : b - a);
    if (!f(tok)) return false;
    if (b == std::string_view::npos) return true;
    a = b + 1;
  }
}

}  // namespace

struct PamrecTokens_ {
  std::vector<Part> parts;
  int64_t n_lines = 0, n_tokens = 0;
  bool train = false;
};

namespace {

bool parse_line(std::string_view line, bool train, const Vocab& users, const Vocab& items, const Vocab& cates, Part& out) {
  line = strip(line);                                          // words = line.strip().split("\t")
  std::string_view w[11];
  const int need = train ? 6 : 11;
  int nw = 0;
  size_t a = 0;
  while (nw < need) {
    size_t b = line.find('\t', a);
    w[nw++] = line.substr(a, b == std::string_view::npos ? std::string_view::npos : b - a);
    if (b == std::string_view::npos) break;
    a = b + 1;
  }
  if (nw < need) return false;                                 // IndexError in the reference
  const int h = train ? 1 : 6;
  if (train) {
    out.user_ids.push_back(users.get(w[0]));
  } else {
    double ls, lp, td;
    if (!parse_int(w[0], &ls) || !parse_float(w[1], &lp) || !parse_float(w[5], &td)) return false;
    out.label_sat.push_back(ls);
    out.label_play.push_back(lp / 1000.0);
    out.user_ids.push_back(users.get(w[2]));
    out.tgt_item.push_back(items.get(w[3]));
    out.tgt_cate.push_back(cates.get(w[4]));
    out.tgt_dur.push_back(td);
  }
  const size_t n0 = out.items.size();
  for_tokens(w[h], [&](std::string_view t) { out.items.push_back(items.get(t)); return true; });
  const size_t n = out.items.size() - n0;
  for_tokens(w[h + 1], [&](std::string_view t) { out.cates.push_back(cates.get(t)); return true; });
  auto floats = [&](std::string_view col, std::vector<double>& dst, double div) {
    return for_tokens(col, [&](std::string_view t) {
      double v;
      if (!parse_float(t, &v)) return false;
      dst.push_back(div == 1.0 ? v : v / div);
      return true;
    });
  };
  if (!floats(w[h + 2], out.durs, 1.0) || !floats(w[h + 3], out.sats, 1.0) || !floats(w[h + 4], out.plays, 1000.0)) return false;
  if (out.cates.size() - n0 != n || out.durs.size() - n0 != n || out.sats.size() - n0 != n || out.plays.size() - n0 != n)
    return false;                                              // ragged columns: the reference zips them, Python handles that
  out.lens.push_back((int64_t)n);
  return true;
}

}  // namespace

extern "C" {

int pamrec_tokenize_file(const char* path, int train, const PamrecVocab* users, const PamrecVocab* items, const PamrecVocab* cates,
                         int n_threads, PamrecTokens* out, int64_t* n_lines, int64_t* n_tokens) {
  if (!path || !users || !items || !cates || !out || !n_lines || !n_tokens) return -1;
  FILE* f = std::fopen(path, "rb");
  if (!f) return -2;
  std::string buf;
  {
    long size = -1;
    if (std::fseek(f, 0, SEEK_END) == 0) size = std::ftell(f);
    if (size < 0 || std::fseek(f, 0, SEEK_SET) != 0) { std::fclose(f); return -2; }
    buf.resize((size_t)size);
    const size_t got = size ? std::fread(&buf[0], 1, (size_t)size, f) : 0;
    const bool bad = std::ferror(f) != 0 || got != (size_t)size;
    std::fclose(f);
    if (bad) return -2;
  }
  // line starts (text mode with universal newlines: "\n" and "\r\n" end a line here, a lone '\r' is left to Python)
  std::vector<size_t> starts;
  {
    const size_t n = buf.size();
    size_t a = 0;
    for (size_t i = 0; i < n; ++i) {
      const unsigned char c = (unsigned char)buf[i];
      if (c >= 0x80 || c == 0) return PAMREC_TOK_FALLBACK;
      if (c == '\r' && (i + 1 >= n || buf[i + 1] != '\n')) return PAMREC_TOK_FALLBACK;
      if (c == '\n') { starts.push_back(a); a = i + 1; }
    }
    if (a < n) starts.push_back(a);
    starts.push_back(n);                                       // sentinel; a line's text may include its "\n" (stripped later)
  }
  const int64_t nl = (int64_t)starts.size() - 1;
  Vocab vu, vi, vc;
  vu.build(users); vi.build(items); vc.build(cates);
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 32) nt = 32;
  if ((int64_t)nt > nl) nt = nl > 0 ? (int)nl : 1;
  PamrecTokens t = new PamrecTokens_();
  t->train = train != 0;
  t->parts.resize(nt);
  auto work = [&](int k) {
    Part& p = t->parts[k];
    const int64_t lo = nl * k / nt, hi = nl * (k + 1) / nt;
    {                                                          // one allocation per column: every history token but the last of a
      size_t commas = 0;                                       // column is followed by a comma, and there are five columns
      for (const char* c = buf.data() + starts[lo], *e = buf.data() + starts[hi]; (c = (const char*)std::memchr(c, ',', e - c)); ++c) ++commas;
      const size_t cap = commas / 5 + 2 * (size_t)(hi - lo) + 16;
      p.items.reserve(cap); p.cates.reserve(cap); p.durs.reserve(cap); p.sats.reserve(cap); p.plays.reserve(cap);
      p.lens.reserve(hi - lo); p.user_ids.reserve(hi - lo);
    }
    for (int64_t i = lo; i < hi && p.ok; ++i)
      p.ok = parse_line(std::string_view(buf.data() + starts[i], starts[i + 1] - starts[i]), train != 0, vu, vi, vc, p);
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int k = 0; k < nt; ++k) th.emplace_back(work, k);
    for (auto& x : th) x.join();
  }
  for (const Part& p : t->parts) {
    if (!p.ok) { delete t; return PAMREC_TOK_FALLBACK; }
    t->n_lines += (int64_t)p.lens.size();
    t->n_tokens += (int64_t)p.items.size();
  }
  *out = t;
  *n_lines = t->n_lines;
  *n_tokens = t->n_tokens;
  return 0;
}

int pamrec_tokens_read(PamrecTokens t, const PamrecLines* dst) {
  if (!t || !dst || dst->n_lines != t->n_lines || !dst->offsets || !dst->user_ids) return -1;
  if (t->n_tokens && (!dst->items || !dst->cates || !dst->durs || !dst->sats || !dst->plays)) return -1;
  if (!t->train && t->n_lines && (!dst->label_sat || !dst->label_play || !dst->tgt_item || !dst->tgt_cate || !dst->tgt_dur)) return -1;
  int64_t line = 0, tok = 0;
  int64_t* off = const_cast<int64_t*>(dst->offsets);
  off[0] = 0;
  auto put = [](const auto& v, const auto* base, int64_t at) {
    using T = typename std::decay<decltype(v)>::type::value_type;
    if (!v.empty()) std::memcpy(const_cast<T*>(base) + at, v.data(), v.size() * sizeof(T));
  };
  for (const Part& p : t->parts) {
    put(p.items, dst->items, tok); put(p.cates, dst->cates, tok);
    put(p.durs, dst->durs, tok); put(p.sats, dst->sats, tok); put(p.plays, dst->plays, tok);
    put(p.user_ids, dst->user_ids, line);
    if (!t->train) {
      put(p.label_sat, dst->label_sat, line); put(p.label_play, dst->label_play, line);
      put(p.tgt_item, dst->tgt_item, line); put(p.tgt_cate, dst->tgt_cate, line); put(p.tgt_dur, dst->tgt_dur, line);
    }
    for (size_t i = 0; i < p.lens.size(); ++i) { off[line + 1] = off[line] + p.lens[i]; ++line; }
    tok += (int64_t)p.items.size();
  }
  return 0;
}

int pamrec_tokens_free(PamrecTokens t) {
  delete t;
  return 0;
}

}  // extern "C"
