"""csrc/tokenizer.cu against the Python parser of the same module (itself bit-identical to the reference's parser_one_line,
tests/test_iterator_and_metrics.py): the flat columns must be EQUAL element for element, including every float64 bit, and any
file outside the strict subset must be refused (so that Python - and with it the reference's behaviour - takes over)."""
import os
import pickle
import random

import numpy as np
import pytest

from oracle import gen_golden as G
from pamrec_b200 import deeprec_utils as DU
from pamrec_b200 import sequential_iterator as IT


def _iterator(d, dataset="wechat", vocabs=None, **kw):
    vocabs = vocabs or ({"default_uid": 0, **{f"u{i}": i for i in range(1, 40)}}, {"default_mid": 0, **{str(i): i for i in range(1, 500)}},
                        {"default_cat": 0, **{f"c{i}": i for i in range(1, 9)}})
    for name, voc in zip(("user_vocab.pkl", "item_vocab.pkl", "category_vocab.pkl"), vocabs):
        with open(d / name, "wb") as f:
            pickle.dump(voc, f)
    (d / f"{dataset}_business_recommenders.csv").write_text("1\t1\t10.0\n")
    hp = DU.prepare_hparams(None, model_type="mmoe", dataset=dataset, bucket_num=10, batch_size=10, max_seq_length=20,
                            noise_train_hist=0, noise_train_listwise=0, user_vocab=str(d / "user_vocab.pkl"),
                            item_vocab=str(d / "item_vocab.pkl"), cate_vocab=str(d / "category_vocab.pkl"),
                            **{"noise_only_predict": 0, **kw})
    return IT.SequentialIterator(hp, None)


def _same_columns(it, path, train):
    nat = it._tokenize_native(str(path), train)
    assert nat is not None, "the native tokenizer refused a file of the strict subset"
    it.train = train
    py = it._flatten(it.parse_file(str(path)), train)
    assert list(nat) == list(py)
    for k in py:
        assert nat[k].dtype == py[k].dtype and nat[k].shape == py[k].shape, k
        assert nat[k].tobytes() == py[k].tobytes(), k            # bytes: -0.0 vs 0.0 and every last bit of a double count
    return nat


@pytest.mark.parametrize("case", list(G.CASES))
def test_native_tokenizer_equals_python_parser_on_the_synthetic_files(case, tmp_path):
    data_dir = G.synth_case(case, str(tmp_path))
    it = IT.SequentialIterator(G.hparams_for(case, data_dir), None)
    for name in ("train_data", "valid_data"):
        for threads in ("1", "3"):
            os.environ["PAMREC_TOKENIZER_THREADS"] = threads
            try:
                col = _same_columns(it, os.path.join(data_dir, name), name == "train_data")
            finally:
                del os.environ["PAMREC_TOKENIZER_THREADS"]
        assert len(col["user_ids"]) > 0 and len(col["items"]) > 0


def test_number_spellings_round_like_python(tmp_path):
    """Decimal spellings whose nearest double is hard (halfway cases, many digits, exponents, signed zeros), CRLF line ends, a last
    line without a newline, extra columns, blanks around a column."""
    d = tmp_path / "wechat"
    d.mkdir()
    it = _iterator(d)
    hard = ["0.1", "1e23", "8.41e21", "2.2250738585072014e-308", "9007199254740993", "9007199254740992.5", "0.30000000000000004",
            "1.7976931348623157e308", "5e-324".replace("5e-324", "4.9e-300"), "-0.0", "+3.5", ".5", "7.", "1E5", "1e+2", "1e-2", "000012.50",
            "123456789012345678901234567890", "0.000000000000000000000000000001", "2.5000000000000002220446049250313e-1",
            "1.00000000000000011102230246251565404236316680908203125", "1.00000000000000011102230246251565404236316680908203124"]
    n = len(hard)
    rng = random.Random(3)
    rows = []
    for u in range(1, 8):
        toks = hard[:]
        rng.shuffle(toks)
        plays = [str(rng.choice([7999, 8000, 12345, 1]))for _ in range(n)]
        rows.append("\t".join([f"u{u}", ",".join(str(rng.randrange(1, 520)) for _ in range(n)), ",".join(f"c{rng.randrange(1, 12)}" for _ in range(n)),
                               ",".join(toks), ",".join(rng.choice(["0", "1", "1.0", "0.0"]) for _ in range(n)), ",".join(plays)]))
    rows[1] = "  " + rows[1] + "  "                               # line.strip()
    rows[2] = rows[2].replace("\t", " \t ", 2)                    # column.strip() on the history columns (not on the user column)
    rows[3] = rows[3] + "\textra\tcolumns"
    (d / "train_data").write_bytes(("\r\n".join(rows)).encode())  # CRLF, no newline at the end
    col = _same_columns(it, d / "train_data", True)
    assert col["user_ids"].tolist() == [1, 2, 0, 4, 5, 6, 7]      # "u3 " is not in the vocabulary: only the history columns are stripped
    ev = []
    for r in rows[:4]:
        w = r.strip().split("\t")
        ev.append("\t".join(["1", "12345.678", w[0].strip(), "77", "c3", "0.1"] + w[1:6]))
    (d / "valid_data").write_text("\n".join(ev) + "\n")
    col = _same_columns(it, d / "valid_data", False)
    assert col["label_play"][0] == 12345.678 / 1000 and col["tgt_item"].tolist() == [77] * 4


def test_random_decimals_round_like_python(tmp_path):
    """60 000 random decimal spellings (1-25 digits, the point anywhere, optional sign / exponent) through both parsers: covers the
    exact short-decimal fast path (<= 15 digits) and std::from_chars beyond it, against float()."""
    d = tmp_path / "wechat"
    d.mkdir()
    it = _iterator(d)
    rng = random.Random(11)

    def spell():
        nd = rng.choice([1, 2, 3, 5, 8, 14, 15, 16, 17, 18, 21, 25])
        digits = "".join(rng.choice("0123456789") for _ in range(nd))
        cut = rng.randrange(0, nd + 1)
        s = digits[:cut] + ("." + digits[cut:] if rng.random() < 0.8 or cut == 0 else digits[cut:])
        if s.startswith("."):
            s = rng.choice(["", "0"]) + s
        if rng.random() < 0.15:
            s += rng.choice("eE") + rng.choice(["", "+", "-"]) + str(rng.randrange(0, 30))
        return rng.choice(["", "", "-", "+"]) + s
    n, rows = 200, []
    for u in range(100):
        rows.append("\t".join([f"u{u % 39 + 1}", ",".join("7" for _ in range(n)), ",".join("c1" for _ in range(n)), ",".join(spell() for _ in range(n)),
                               ",".join(spell() for _ in range(n)), ",".join(spell() for _ in range(n))]))
    (d / "train_data").write_text("\n".join(rows) + "\n")
    col = _same_columns(it, d / "train_data", True)
    assert len(col["durs"]) == 100 * n


@pytest.mark.parametrize("what", ["non_ascii", "lone_cr", "underscore", "inf", "nan", "hex", "inner_blank", "trailing_comma", "ragged",
                                  "missing_column", "blank_line", "overflow", "empty_column", "float_label"])
def test_everything_else_is_left_to_python(tmp_path, what):
    """Inputs Python's float()/int()/str.strip() treat in ways the strict parser does not reproduce are REFUSED; the iterator then
    behaves exactly like the Python parser (same batches, or the same exception)."""
    d = tmp_path / "wechat"
    d.mkdir()
    it = _iterator(d)
    n = 12
    base = lambda u: [f"u{u}", ",".join(str(i + 1) for i in range(n)), ",".join(["c1"] * n), ",".join(["10.0"] * n), ",".join(["1"] * n),
                      ",".join(["12000"] * n)]
    rows = [base(u) for u in range(1, 5)]
    train = True
    if what == "non_ascii":
        rows[1][0] = "ü2"
    elif what == "underscore":
        rows[2][3] = rows[2][3].replace("10.0", "1_0.0", 1)
    elif what in ("inf", "nan"):
        rows[2][3] = rows[2][3].replace("10.0", what, 1)
    elif what == "hex":
        rows[2][3] = rows[2][3].replace("10.0", "0x10", 1)       # strtod would take it, Python raises
    elif what == "inner_blank":
        rows[0][5] = rows[0][5].replace("12000,", "12000, ", 1)
    elif what == "trailing_comma":
        rows[3][4] += ","
    elif what == "ragged":
        rows[1][5] += ",9000"
    elif what == "missing_column":
        rows[2] = rows[2][:5]
    elif what == "overflow":
        rows[0][3] = rows[0][3].replace("10.0", "1e999", 1)      # float("1e999") = inf in Python
    elif what == "empty_column":
        rows[1][2] = ""
    text = "\n".join("\t".join(r) for r in rows) + "\n"
    if what == "lone_cr":
        text = text.replace("\n", "\r", 1)                       # universal newlines: still a line end for Python
    if what == "blank_line":
        text = text.replace("\n", "\n\n", 1)
    if what == "float_label":
        train = False
        text = "\n".join("\t".join(["1.0", "9000", r[0], "5", "c1", "10.0"] + r[1:]) for r in rows) + "\n"
    path = d / ("train_data" if train else "valid_data")
    path.write_bytes(text.encode())
    assert it._tokenize_native(str(path), train) is None

    def run(py):
        os.environ["PAMREC_PY_ITERATOR"] = "1" if py else "0"
        try:
            random.seed(4)
            fresh = _iterator(d)
            return [{k: v.tobytes() for k, v in b.items()} for b in fresh.load_data_from_file(str(path))], None
        except Exception as e:                                   # noqa: BLE001 - the point is that both raise the same thing
            return None, (type(e), str(e))
        finally:
            del os.environ["PAMREC_PY_ITERATOR"]
    assert run(False) == run(True)


def test_vocabularies_that_are_not_plain_str_to_int_dicts_keep_python(tmp_path):
    from pamrec_b200.synth import SyntheticVocab
    d = tmp_path / "wechat"
    d.mkdir()
    it = _iterator(d, vocabs=({"default_uid": 0, "u1": 1}, SyntheticVocab(2_000_000, "default_mid"), {"default_cat": 0, "c1": 1}))
    n = 12
    row = "\t".join(["u1", ",".join(str(1_000_000 + i) for i in range(n)), ",".join(["c1"] * n), ",".join(["10.0"] * n), ",".join(["1"] * n),
                     ",".join(["12000"] * n)])
    (d / "train_data").write_text(row + "\n")
    assert it._tokenize_native(str(d / "train_data"), True) is None
    random.seed(1)
    b = next(it.load_data_from_file(str(d / "train_data")))
    first = int(b["items"][0])                                               # parsed by Python (ids above one million), batched natively
    assert 1_000_001 <= first <= 1_000_005 and b["items"][:5].tolist() == [first + i for i in range(5)]
    it2 = _iterator(d, vocabs=({"default_uid": 0, "u1": 1}, {"default_mid": 0, 7: 7}, {"default_cat": 0}))
    assert it2._tokenize_native(str(d / "train_data"), True) is None
    with pytest.raises(FileNotFoundError):
        _iterator(d)._tokenize_native(str(d / "nope"), True)
    it3 = _iterator(d, noise_only_predict=0.5)                               # draws from np.random per element while parsing
    assert it3._tokenize_native(str(d / "train_data"), True) is None


def test_fuzz_native_accepts_only_what_it_parses_like_python(tmp_path):
    """400 random small files mixing well-formed lines with every kind of irregularity (odd spellings, blanks, missing / extra /
    ragged columns, CR, non-ASCII).  Whenever the native tokenizer ACCEPTS a file its columns must equal the Python parser's
    byte for byte (so Python did not raise either); refusing is always allowed, but well-formed files must be accepted."""
    d = tmp_path / "wechat"
    d.mkdir()
    it = _iterator(d)
    rng = random.Random(2024)
    num_ok = ["0", "1", "7", "12.5", "0.25", "3e2", "-1.5", "+2", "1E-3", "100000", ".5", "5.", "007", "1.0", "0.0"]
    num_odd = ["", " 1", "1 ", "1_0", "nan", "inf", "-inf", "0x10", "1e", "e5", "--1", "1.2.3", "١", "1e400", "1,5".replace(",", ";"), "\t"]
    ids = [str(i) for i in range(1, 30)] + ["zz", "", "c1", "c9", "u3", " 4", "ü"]
    accepted = refused = 0
    for case in range(400):
        train = rng.random() < 0.5
        clean = rng.random() < 0.35
        rows = []
        for _ in range(rng.randrange(1, 5)):
            n = rng.randrange(1, 6)

            def col(pool_ok, pool_odd, k=None):
                k = n if k is None else k
                return ",".join(rng.choice(pool_ok if clean or rng.random() < 0.93 else pool_odd) for _ in range(k))
            ragged = (lambda: n) if clean or rng.random() < 0.9 else (lambda: rng.randrange(1, 6))
            hist = [col(ids[:29], ids, ragged()), col(["c1", "c2", "c9"], ids, ragged()), col(num_ok, num_odd, ragged()),
                    col(["0", "1", "1.0", "0.0"], num_odd, ragged()), col(["12000", "500", "7999.5", "8000"], num_odd, ragged())]
            head = [f"u{rng.randrange(1, 45)}"] if train else [rng.choice(["0", "1"] if clean else ["0", "1", "1.0", " 1", "+1", "x"]),
                                                                rng.choice(num_ok if clean else num_ok + num_odd[:6]), f"u{rng.randrange(1, 45)}",
                                                                rng.choice(ids[:29]), "c2", rng.choice(num_ok)]
            cols = head + hist
            if not clean and rng.random() < 0.08:
                cols = cols[:rng.randrange(1, len(cols))]                      # missing columns
            if rng.random() < 0.1:
                cols = cols + ["extra"]
            line = "\t".join(cols)
            if not clean and rng.random() < 0.1:
                line = rng.choice([" ", "\t", ""]) + line + rng.choice([" ", "\t", "\x0b", ""])
            rows.append(line)
        sep = "\n" if clean else rng.choice(["\n", "\n", "\r\n", "\r", "\n\n"])
        text = sep.join(rows) + rng.choice(["\n", ""])
        path = d / ("train_data" if train else "valid_data")
        path.write_bytes(text.encode("utf-8"))
        nat = it._tokenize_native(str(path), train)
        if nat is None:
            assert not clean, text
            refused += 1
            continue
        accepted += 1
        it.train = train
        py = it._flatten(it.parse_file(str(path)), train)              # must not raise: the native parser accepted the file
        assert py is not None and list(nat) == list(py)
        for k in py:
            assert nat[k].dtype == py[k].dtype and nat[k].tobytes() == py[k].tobytes(), (case, k, text)
    assert accepted > 100 and refused > 100, (accepted, refused)
