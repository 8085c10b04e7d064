// Host input pipeline (no GPU work in this file): the batching of the reference's SequentialIterator over a file that
// has been tokenised once into flat columns.  Replaces the per-element Python loops of
//   io/sequential_iterator.py:475-763  (train: per-user history state machine, listwise groups of 5, round-robin passes)
//   io/sequential_iterator.py:375-474  (eval: one impression per line)
//   io/sequential_iterator.py:1009-1141 (_convert_data: padding, masks, play-ratio buckets, satisfied-only compaction)
// and reproduces their output bit for bit (tests/test_iterator_and_metrics.py compares SHA-256 digests of all 19 feed arrays
// with the reference's own iterator).  Python's `random` stays in Python: the caller passes the shuffled user order and the
// per-user warm-up lengths, so the RNG call sequence of the reference (IT:545, IT:622) is untouched.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/pamrec_b200.h"

namespace {

constexpr int kGroup = PAMREC_GROUP;          // BEGIN_HISTORY_LEN_MAX, IT:98
constexpr int kMaxSequence = 100;             // IT:476
constexpr double kValidThreshold = 8.0;       // seconds, IT:96

struct UserState {
  int64_t line = 0, cursor = 0;
  int32_t user = 0;
  std::vector<int32_t> items, cates;
  std::vector<double> durs, sats, plays;
  std::vector<int> nsi;                       // positions of not-satisfied entries, ascending
};

}  // namespace

struct PamrecBatcher_ {
  PamrecLines L;
  std::vector<double> borders;
  int T = 0;
  // train pass
  std::vector<UserState> users;
  std::vector<int64_t> alive, next_alive;
  size_t alive_pos = 0;
  // eval pass
  int64_t eval_line = 0;
  int min_seq = 1;
  bool training = false, active = false;
  std::vector<int64_t> eval_jobs;
  int n_threads = 1;                          // threads that fill the history arrays of an evaluation batch (PAMREC_BATCHER_THREADS)
  int n_cores = 1;

  // io/sequential_iterator.py:43-53 with numpy.searchsorted(side="right") semantics (NaN sorts last)
  int lisan(double x) const {
    int idx;
    if (std::isnan(x)) idx = (int)borders.size();
    else idx = (int)(std::upper_bound(borders.begin(), borders.end(), x) - borders.begin());
    idx -= 1;
    return idx < 0 ? 0 : idx;
  }

  // add_a_item_to_hist, IT:479-522
  void push(UserState& u, int64_t k) const {
    const double play = L.plays[k], sat = L.sats[k];
    if (play < kValidThreshold && sat != 1.0) return;
    u.items.push_back(L.items[k]); u.cates.push_back(L.cates[k]);
    u.durs.push_back(L.durs[k]); u.sats.push_back(sat); u.plays.push_back(play);
    if (sat == 0.0) u.nsi.push_back((int)u.items.size() - 1);
    const int len = (int)u.items.size();
    if (len > kMaxSequence) {
      const int k_evict = len - kMaxSequence;
      if (u.nsi.empty()) {
        u.items.erase(u.items.begin(), u.items.begin() + k_evict); u.cates.erase(u.cates.begin(), u.cates.begin() + k_evict);
        u.durs.erase(u.durs.begin(), u.durs.begin() + k_evict); u.sats.erase(u.sats.begin(), u.sats.begin() + k_evict);
        u.plays.erase(u.plays.begin(), u.plays.begin() + k_evict);
      } else {
        const int n_pop = std::min<int>(k_evict, (int)u.nsi.size());     // nsi[:k]
        for (int j = 0; j < n_pop; ++j) {                                // evict the oldest unsatisfied first
          const int idx = u.nsi[j] - j;
          u.items.erase(u.items.begin() + idx); u.cates.erase(u.cates.begin() + idx);
          u.durs.erase(u.durs.begin() + idx); u.sats.erase(u.sats.begin() + idx); u.plays.erase(u.plays.begin() + idx);
        }
        std::vector<int> rest;
        for (size_t q = (size_t)n_pop; q < u.nsi.size(); ++q) rest.push_back(u.nsi[q] - k_evict);
        u.nsi.swap(rest);
      }
    }
  }
};

namespace {

struct Out {
  float *labels_satisfied, *labels_play, *plays, *users;
  int32_t *items, *cates;
  float* durations;
  int32_t *item_history, *item_cate_history;
  float *item_duration_history, *mask, *item_satisfied_value_history, *item_play_value_history, *item_loop_times_history;
  int32_t *satisfied_item_history, *satisfied_cate_history;
  float *satisfied_duration_history, *satisfied_play_history, *satisfied_mask;
};

// one history (last min(len, T) entries, left aligned) into `rep` consecutive rows starting at row r  (IT:1052-1103)
template <typename I, typename C, typename D>
void write_history(const PamrecBatcher_& b, const Out& o, int64_t r, int rep, int64_t len, I item_at, C cate_at, D dbl_at) {
  const int T = b.T;
  const int64_t Lw = std::min<int64_t>(len, T), first = len - Lw;
  int32_t* ih = o.item_history + r * T;
  int32_t* ch = o.item_cate_history + r * T;
  float* dh = o.item_duration_history + r * T;
  float* sh = o.item_satisfied_value_history + r * T;
  float* ph = o.item_play_value_history + r * T;
  float* lh = o.item_loop_times_history + r * T;
  float* mk = o.mask + r * T;
  int32_t* si = o.satisfied_item_history + r * T;
  int32_t* sc = o.satisfied_cate_history + r * T;
  float* sd = o.satisfied_duration_history + r * T;
  float* sp = o.satisfied_play_history + r * T;
  float* sm = o.satisfied_mask + r * T;
  int cnt = 0;
  for (int t = 0; t < T; ++t) {
    if (t < Lw) {
      const int64_t k = first + t;
      const double dur = dbl_at(0, k), sat = dbl_at(1, k), play = dbl_at(2, k);
      ih[t] = item_at(k); ch[t] = cate_at(k);
      dh[t] = (float)dur; sh[t] = (float)sat; ph[t] = (float)play;
      lh[t] = (float)b.lisan(play / dur);                       // float64 ratio, IT:421 / IT:682
      mk[t] = 1.0f;
    } else {
      ih[t] = 0; ch[t] = 0; dh[t] = 0.f; sh[t] = 0.f; ph[t] = 0.f; lh[t] = 0.f; mk[t] = 0.f;
    }
    si[t] = 0; sc[t] = 0; sd[t] = 0.f; sp[t] = 0.f; sm[t] = 0.f;
  }
  for (int t = 0; t < Lw; ++t) {                                // satisfied-only compaction, bucket from the float32 values
    if (sh[t] == 1.0f) {
      si[cnt] = ih[t]; sc[cnt] = ch[t]; sd[cnt] = dh[t];
      sp[cnt] = (float)b.lisan((double)(ph[t] / dh[t]));
      sm[cnt] = 1.0f;
      ++cnt;
    }
  }
  for (int q = 1; q < rep; ++q) {
    const size_t bi = sizeof(int32_t) * T, bf = sizeof(float) * T;
    memcpy(ih + q * T, ih, bi); memcpy(ch + q * T, ch, bi); memcpy(dh + q * T, dh, bf); memcpy(sh + q * T, sh, bf);
    memcpy(ph + q * T, ph, bf); memcpy(lh + q * T, lh, bf); memcpy(mk + q * T, mk, bf);
    memcpy(si + q * T, si, bi); memcpy(sc + q * T, sc, bi); memcpy(sd + q * T, sd, bf); memcpy(sp + q * T, sp, bf);
    memcpy(sm + q * T, sm, bf);
  }
}

}  // namespace

extern "C" {

int pamrec_batcher_create(const PamrecLines* lines, const double* borders, int n_borders, int max_seq_len, PamrecBatcher* out) {
  if (!lines || !borders || n_borders < 1 || max_seq_len < 1 || !out) return -1;
  PamrecBatcher b = new PamrecBatcher_();
  b->L = *lines;
  b->borders.assign(borders, borders + n_borders);
  b->T = max_seq_len;
  {
    const char* e = getenv("PAMREC_BATCHER_THREADS");
    const int hw = (int)std::thread::hardware_concurrency();
    int nt = e ? atoi(e) : hw;
    b->n_threads = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
    b->n_cores = e ? 1 << 20 : (hw < 1 ? 1 : hw);             // an explicit thread count is taken as given
  }
  *out = b;
  return 0;
}

int pamrec_batcher_destroy(PamrecBatcher b) {
  delete b;
  return 0;
}

int pamrec_batcher_begin_train(PamrecBatcher b, const int64_t* order, const int32_t* begin_loc, int64_t n) {
  if (!b || (n > 0 && (!order || !begin_loc))) return -1;
  b->users.assign((size_t)n, UserState());
  b->alive.resize((size_t)n);
  for (int64_t k = 0; k < n; ++k) {
    UserState& u = b->users[(size_t)k];
    u.line = order[k];
    if (u.line < 0 || u.line >= b->L.n_lines) return -2;
    u.user = b->L.user_ids[u.line];
    const int64_t lo = b->L.offsets[u.line], hi = b->L.offsets[u.line + 1];
    int64_t i = lo;
    while (i < hi && b->L.sats[i] != 1.0) { b->push(u, i); ++i; }       // IT:556-570
    for (int q = 0; q < begin_loc[k] && i < hi; ++q) { b->push(u, i); ++i; }   // IT:573-588
    u.cursor = i;
    b->alive[(size_t)k] = k;
  }
  b->next_alive.clear();
  b->alive_pos = 0;
  b->training = true; b->active = true;
  return 0;
}

int pamrec_batcher_begin_eval(PamrecBatcher b, int min_seq_length) {
  if (!b || !b->L.label_sat) return -1;
  b->eval_line = 0; b->min_seq = min_seq_length;
  b->training = false; b->active = true;
  return 0;
}

}  // extern "C"

namespace {

// One GLOBAL batch of up to batch_size rows is drawn from the pass; the units rank `rank` of `world` owns (listwise groups
// g = rank, rank + world, ... of a training batch, rows likewise of an eval batch: pamrec_b200/dist.py split_feed) are written,
// compacted, into the arrays.  The history state of every user advances whether or not its group is written: the expensive
// part - padding, bucketing and compacting T entries for 12 arrays - is done for owned units only.
int next_impl(PamrecBatcher b, int batch_size, int world, int rank, void* const* arrays, float* g_labels, float* g_users,
              int* global_rows) {
  if (!b || !arrays || batch_size < 1 || !b->active || world < 1 || rank < 0 || rank >= world) return -1;
  Out o;
  static_assert(sizeof(Out) == 19 * sizeof(void*), "19 feed arrays");
  memcpy(&o, arrays, sizeof o);
  const PamrecLines& L = b->L;
  int64_t rows = 0, local = 0;                                // rows of the global batch so far / rows written here
  if (b->training) {
    if (batch_size % kGroup) return -3;
    while (rows < batch_size) {
      if (b->alive_pos == b->alive.size()) {                   // next round-robin pass
        if (b->next_alive.empty()) { b->alive.clear(); b->alive_pos = 0; break; }
        b->alive.swap(b->next_alive);
        b->next_alive.clear();
        b->alive_pos = 0;
        continue;
      }
      const int64_t ind = b->alive[b->alive_pos++];
      UserState& u = b->users[(size_t)ind];
      const int64_t hi = L.offsets[u.line + 1];
      const int64_t future = hi - u.cursor;
      if (future < kGroup) continue;
      const bool mine = (rows / kGroup) % world == rank;
      for (int q = 0; q < kGroup; ++q) {                       // what the metrics need of every row of the global batch
        if (g_labels) g_labels[rows + q] = (float)L.sats[u.cursor + q];
        if (g_users) g_users[rows + q] = (float)u.user;
      }
      for (int q = 0; mine && q < kGroup; ++q) {               // IT:645-676
        const int64_t k = u.cursor + q;
        const int64_t r = local + q;
        o.labels_satisfied[r] = (float)L.sats[k];
        o.labels_play[r] = L.plays[k] >= kValidThreshold ? 1.0f : 0.0f;
        o.plays[r] = (float)b->lisan(L.plays[k] / L.durs[k]);
        o.users[r] = (float)u.user;                            // IT:1115 builds users as float32
        o.items[r] = L.items[k]; o.cates[r] = L.cates[k];
        o.durations[r] = (float)L.durs[k];
      }
      if (mine) {
        write_history(*b, o, local, kGroup, (int64_t)u.items.size(), [&](int64_t k) { return u.items[(size_t)k]; },
                      [&](int64_t k) { return u.cates[(size_t)k]; },
                      [&](int which, int64_t k) { return which == 0 ? u.durs[(size_t)k] : (which == 1 ? u.sats[(size_t)k] : u.plays[(size_t)k]); });
        local += kGroup;
      }
      rows += kGroup;
      if (future > kGroup) {                                   // IT:719-741
        for (int q = 0; q < kGroup; ++q) b->push(u, u.cursor + q);
        u.cursor += kGroup;
        b->next_alive.push_back(ind);
      }
    }
  } else {
    // The lines of an evaluation pass are independent (no history state, IT:375-474): the per-line scalars are written here, the
    // T-long history arrays of the batch's rows - the expensive part - by a few threads over disjoint rows afterwards.
    std::vector<int64_t>& jobs = b->eval_jobs;                 // line of the k-th row written by this call
    jobs.clear();
    while (rows < batch_size && b->eval_line < L.n_lines) {
      const int64_t ln = b->eval_line++;
      const int64_t lo = L.offsets[ln], hi = L.offsets[ln + 1];
      if (hi - lo < b->min_seq) continue;                      // IT:406-407
      if (g_labels) g_labels[rows] = (float)L.label_sat[ln];
      if (g_users) g_users[rows] = (float)L.user_ids[ln];
      if (rows % world != rank) { rows += 1; continue; }
      const int64_t r = local;
      o.labels_satisfied[r] = (float)L.label_sat[ln];
      o.labels_play[r] = L.label_play[ln] >= 10.0 ? 1.0f : 0.0f;   // IT:410 (10 s at eval, 8 s at train)
      o.plays[r] = (float)L.label_play[ln];                    // IT:411 seconds, not a bucket
      o.users[r] = (float)L.user_ids[ln];
      o.items[r] = L.tgt_item[ln]; o.cates[r] = L.tgt_cate[ln];
      o.durations[r] = (float)L.tgt_dur[ln];
      jobs.push_back(ln);
      rows += 1;
      local += 1;
    }
    auto run = [&](int64_t r0, int64_t r1) {
      for (int64_t r = r0; r < r1; ++r) {
        const int64_t lo = L.offsets[jobs[(size_t)r]], hi = L.offsets[jobs[(size_t)r] + 1];
        write_history(*b, o, r, 1, hi - lo, [&](int64_t k) { return L.items[lo + k]; }, [&](int64_t k) { return L.cates[lo + k]; },
                      [&](int which, int64_t k) { return which == 0 ? L.durs[lo + k] : (which == 1 ? L.sats[lo + k] : L.plays[lo + k]); });
      }
    };
    const int64_t n = (int64_t)jobs.size();
    int nt = b->n_threads;
    if (world > 1 && b->n_cores / world < nt) nt = b->n_cores / world;   // one process per GPU: the ranks share the host's cores
    if (n < 512 || nt < 1) nt = 1;                             // not worth the thread start-up
    if (nt <= 1) run(0, n);
    else {
      std::vector<std::thread> pool;
      for (int i = 1; i < nt; ++i) pool.emplace_back(run, n * i / nt, n * (i + 1) / nt);
      run(0, n / nt);
      for (auto& th : pool) th.join();
    }
  }
  if (rows == 0) b->active = false;
  if (global_rows) *global_rows = (int)rows;
  return (int)local;
}

}  // namespace

extern "C" {

int pamrec_batcher_next(PamrecBatcher b, int batch_size, void* const* arrays) {
  return next_impl(b, batch_size, 1, 0, arrays, nullptr, nullptr, nullptr);
}

int pamrec_batcher_next_shard(PamrecBatcher b, int batch_size, int world, int rank, void* const* arrays,
                              float* global_labels_satisfied, float* global_users, int* global_rows) {
  if (!global_rows) return -1;
  return next_impl(b, batch_size, world, rank, arrays, global_labels_satisfied, global_users, global_rows);
}

}  // extern "C"
