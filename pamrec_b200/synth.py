"""Seeded synthetic WeChat-Channels / MX-TakaTak-shaped data in the reference's on-disk formats
(SURVEY.md Appendix C, section 8d).  The real datasets are not available offline.

Files written under ``<root>/<dataset>/`` (names fixed by example/00_quick_start/sequential.py:454-460):
  train_data                       one line per USER, 6 tab-separated columns (io/sequential_iterator.py:208-230)
  valid_data, test_data            one line per IMPRESSION, 11 columns (io/sequential_iterator.py:231-268)
  user_vocab.pkl, item_vocab.pkl, category_vocab.pkl   dict[str -> int], index 0 reserved
  <dataset>_business_recommenders.csv                  item \t cate \t duration (io/sequential_iterator.py:72-86)
"""
import os
import pickle

import numpy as np

from .sequential_iterator import bar_border_list, takatak_bar_border_list_dict

SHAPES = {
    "wechat": dict(n_users=20000, n_items=100000, n_cates=500, mean_len=150, zipf=1.1),
    "takatak": dict(n_users=50000, n_items=30000, n_cates=50, mean_len=120, zipf=1.1),
}


def _zipf(rng, a, size, n):
    r = rng.zipf(a, size=size).astype(np.int64)
    return 1 + (r - 1) % n


def generate(root, dataset="wechat", n_users=None, n_items=None, n_cates=None, mean_len=None, seed=20231,
             eval_per_user=2, min_len=12, max_len=600, n_neg=0):
    """Returns the dataset directory.  ``n_neg`` extra negative lines follow each positive eval line."""
    shape = dict(SHAPES[dataset])
    for k, v in (("n_users", n_users), ("n_items", n_items), ("n_cates", n_cates), ("mean_len", mean_len)):
        if v is not None:
            shape[k] = v
    rng = np.random.default_rng(seed)
    d = os.path.join(root, dataset)
    os.makedirs(d, exist_ok=True)
    NU, NI, NC = shape["n_users"], shape["n_items"], shape["n_cates"]
    borders = np.asarray(bar_border_list if dataset == "wechat" else takatak_bar_border_list_dict[10], np.float64).copy()
    borders[-1] = min(borders[-1], 5.0)            # cap the open last decile (keeps takatak away from bucket 10)
    if dataset == "takatak":
        borders = borders[:11]
    item_cate = rng.integers(1, NC + 1, size=NI + 1)
    item_dur = rng.integers(5, 61, size=NI + 1).astype(np.float64)
    vocab = lambda n, dflt: {dflt: 0, **{str(i): i for i in range(1, n + 1)}}
    for name, voc in (("user_vocab.pkl", vocab(NU, "default_uid")), ("item_vocab.pkl", vocab(NI, "default_mid")),
                      ("category_vocab.pkl", vocab(NC, "default_cat"))):
        with open(os.path.join(d, name), "wb") as f:
            pickle.dump(voc, f)
    with open(os.path.join(d, f"{dataset}_business_recommenders.csv"), "w") as f:
        for i in range(1, NI + 1):
            f.write(f"{i}\t{item_cate[i]}\t{item_dur[i]:.1f}\n")
    lens = np.clip(rng.lognormal(np.log(shape["mean_len"]) - 0.32, 0.8, size=NU), min_len, max_len).astype(np.int64)
    nb = len(borders) - 1
    join = lambda a, fmt="{}": ",".join(fmt.format(x) for x in a)
    with open(os.path.join(d, "train_data"), "w") as ftr, open(os.path.join(d, "valid_data"), "w") as fva, \
            open(os.path.join(d, "test_data"), "w") as fte:
        for u in range(1, NU + 1):
            L = int(lens[u - 1]) + 2 * eval_per_user
            items = _zipf(rng, shape["zipf"], L, NI)
            dec = rng.integers(0, nb, size=L)
            ratio = borders[dec] + rng.random(L) * (borders[dec + 1] - borders[dec])
            durs = item_dur[items]
            plays_ms = np.maximum((ratio * durs * 1000).astype(np.int64), 1)
            sats = (plays_ms / 1000.0 / durs >= 1.0).astype(np.int64)
            cates = item_cate[items]
            n_tr = L - 2 * eval_per_user
            ftr.write("\t".join([str(u), join(items[:n_tr]), join(cates[:n_tr]), join(durs[:n_tr], "{:.1f}"),
                                 join(sats[:n_tr]), join(plays_ms[:n_tr])]) + "\n")
            for k in range(2 * eval_per_user):
                pos = n_tr + k
                out = fva if k < eval_per_user else fte
                h0 = max(0, pos - 120)
                hist = "\t".join([join(items[h0:pos]), join(cates[h0:pos]), join(durs[h0:pos], "{:.1f}"), join(sats[h0:pos]),
                                  join(plays_ms[h0:pos])])
                out.write("\t".join([str(sats[pos]), str(plays_ms[pos]), str(u), str(items[pos]), str(cates[pos]),
                                     f"{durs[pos]:.1f}", hist]) + "\n")
                for _ in range(n_neg):
                    ni = int(rng.integers(1, NI + 1))
                    out.write("\t".join(["0", "0", str(u), str(ni), str(item_cate[ni]), f"{item_dur[ni]:.1f}", hist]) + "\n")
    return d


def array_batch(seed, B, T, n_users, n_items, n_cates, grouped=True, min_len=1, zipf_a=1.1, n_buckets=10):
    """Array-level synthetic feed in the layout of io/sequential_iterator.py:1111-1135 (live keys only), for
    benchmarks at sizes where the text path would dominate (cfg 3).  grouped=True gives the
    training layout: 5 consecutive rows share one history (io/sequential_iterator.py:645-684)."""
    rng = np.random.default_rng(seed)
    G = 5
    assert not grouped or B % G == 0
    n_hist = B // G if grouped else B
    lens = rng.integers(min_len, T + 1, size=n_hist)
    col = np.arange(T)[None, :]
    mask = (col < lens[:, None]).astype(np.int32)
    ih = (_zipf(rng, zipf_a, (n_hist, T), max(n_items - 1, 1)) * mask).astype(np.int32)
    ch = (_zipf(rng, zipf_a, (n_hist, T), max(n_cates - 1, 1)) * mask).astype(np.int32)
    bk = (rng.integers(0, n_buckets, size=(n_hist, T)) * mask).astype(np.float32)
    rep = G if grouped else 1
    return {
        "item_history": np.repeat(ih, rep, axis=0),
        "item_cate_history": np.repeat(ch, rep, axis=0),
        "item_loop_times_history": np.repeat(bk, rep, axis=0),
        "mask": np.repeat(mask, rep, axis=0).astype(np.float32),
        "users": np.repeat(rng.integers(1, n_users, size=n_hist), rep).astype(np.float32),
        "items": _zipf(rng, zipf_a, B, max(n_items - 1, 1)).astype(np.int32),
        "cates": _zipf(rng, zipf_a, B, max(n_cates - 1, 1)).astype(np.int32),
        "labels_satisfied": rng.integers(0, 2, size=(B, 1)).astype(np.float32),
        "labels_play": rng.integers(0, 2, size=(B, 1)).astype(np.float32),
        "plays": rng.integers(0, n_buckets, size=(B, 1)).astype(np.float32),
    }


class SyntheticVocab:
    """dict-like `token -> index` vocabulary of `n` entries ("default_*" -> 0, "i" -> i) that pickles in a few bytes:
    array-level benchmarks on 10 M-row tables need `len(vocab)` (sequential_base_model.py:565-567), not 10 M dict entries."""

    def __init__(self, n, default):
        self.n, self.default = int(n), default

    def __len__(self):
        return self.n

    def _index(self, key):
        if key == self.default:
            return 0
        try:
            i = int(key)
        except (TypeError, ValueError):
            return None
        return i if 0 < i < self.n else None

    def __contains__(self, key):
        return self._index(key) is not None

    def __getitem__(self, key):
        i = self._index(key)
        if i is None:
            raise KeyError(key)
        return i

    def get(self, key, default=None):
        i = self._index(key)
        return default if i is None else i


def write_vocab_only(root, dataset, n_users, n_items, n_cates):
    """Just the files a model needs to be constructed (vocab pickles + meta csv), for array-level benchmarks."""
    d = os.path.join(root, dataset)
    os.makedirs(d, exist_ok=True)
    vocab = lambda n, dflt: SyntheticVocab(n, dflt) if n > 1_000_000 else {dflt: 0, **{str(i): i for i in range(1, n)}}
    for name, voc in (("user_vocab.pkl", vocab(n_users, "default_uid")), ("item_vocab.pkl", vocab(n_items, "default_mid")),
                      ("category_vocab.pkl", vocab(n_cates, "default_cat"))):
        with open(os.path.join(d, name), "wb") as f:
            pickle.dump(voc, f)
    with open(os.path.join(d, f"{dataset}_business_recommenders.csv"), "w") as f:
        f.write("1\t1\t10.0\n")
    return d


def write_eval_file(path, n_impressions, n_neg, T, n_users, n_items, n_cates, seed=5, zipf_a=1.1):
    """An 11-column eval file (io/sequential_iterator.py:231-268) with one positive line followed by `n_neg` negative lines per
    impression, all sharing the impression's history of up to T items - the layout run_weighted_eval(num_ngs=n_neg) expects
    (sequential_base_model.py:437,456).  Token i is vocabulary index i (write_vocab_only / generate)."""
    rng = np.random.default_rng(seed)
    join = lambda a, fmt="{}": ",".join(fmt.format(x) for x in a)
    with open(path, "w") as f:
        for _ in range(n_impressions):
            L = int(rng.integers(max(1, T // 2), T + 1))
            items = _zipf(rng, zipf_a, L, max(n_items - 1, 1))
            cates = _zipf(rng, zipf_a, L, max(n_cates - 1, 1))
            durs = rng.integers(5, 61, size=L).astype(np.float64)
            plays_ms = np.maximum((rng.random(L) * 2.0 * durs * 1000).astype(np.int64), 1)
            sats = (plays_ms / 1000.0 >= durs).astype(np.int64)
            hist = "\t".join([join(items), join(cates), join(durs, "{:.1f}"), join(sats), join(plays_ms)])
            u = int(rng.integers(1, n_users))
            tgt = rng.integers(1, n_items, size=n_neg + 1)
            tc = rng.integers(1, n_cates, size=n_neg + 1)
            lines = []
            for k in range(n_neg + 1):
                lab, play = ("1", "15000") if k == 0 else ("0", "0")
                lines.append("\t".join([lab, play, str(u), str(tgt[k]), str(tc[k]), "20.0", hist]))
            f.write("\n".join(lines) + "\n")
    return path
