"""pamrec_b200.prefetch.Prefetcher: same items, same order, exceptions and early exit, and the iterator's batches through it."""
import os
import random
import tempfile
import threading
import time

import numpy as np
import pytest

from oracle import gen_golden as G
from pamrec_b200 import sequential_iterator as IT
from pamrec_b200.prefetch import Prefetcher


def test_order_and_completion():
    assert list(Prefetcher(iter(range(1000)), depth=3)) == list(range(1000))
    assert list(Prefetcher(iter(()))) == []


def test_producer_runs_ahead_but_is_bounded():
    produced = []

    def gen():
        for i in range(50):
            produced.append(i)
            yield i
    p = Prefetcher(gen(), depth=2)
    assert next(p) == 0
    time.sleep(0.3)
    assert 2 <= len(produced) <= 5          # ran ahead of the consumer, but only by the queue depth (+1 in hand, +1 blocked)
    assert list(p) == list(range(1, 50))


def test_exception_reaches_the_consumer_after_the_items_before_it():
    def gen():
        yield 1
        yield 2
        raise ValueError("boom")
    p = Prefetcher(gen())
    assert next(p) == 1 and next(p) == 2
    with pytest.raises(ValueError, match="boom"):
        next(p)


def test_close_stops_the_producer():
    n = threading.active_count()
    p = Prefetcher(iter(range(10 ** 9)), depth=2)
    assert next(p) == 0
    p.close()
    time.sleep(0.1)
    assert threading.active_count() <= n


def test_iterator_batches_are_unchanged_through_the_prefetcher():
    case = list(G.CASES)[0]
    with tempfile.TemporaryDirectory() as tmp:
        data_dir = G.synth_case(case, tmp)
        hp = G.hparams_for(case, data_dir)
        train, valid = os.path.join(data_dir, "train_data"), os.path.join(data_dir, "valid_data")
        random.seed(8)
        it = IT.SequentialIterator(hp, None)
        direct = [list(it.load_data_from_file(train)), list(it.load_data_from_file(valid)), list(it.load_data_from_file(train))]
        random.seed(8)
        it = IT.SequentialIterator(hp, None)
        ahead = []
        # a scoring pass in the middle of a suspended training generator, as fit_step does at every eval_step
        p = Prefetcher(it.load_data_from_file(train))
        first = [next(p) for _ in range(2)]
        mid_eval = list(Prefetcher(it.load_data_from_file(valid)))
        ahead.append(first + list(p))
        ahead.append(mid_eval)
        ahead.append(list(Prefetcher(it.load_data_from_file(train))))
    for a, b in zip(direct, ahead):
        assert len(a) == len(b) > 0
        for x, y in zip(a, b):
            assert list(x) == list(y)
            for k in x:
                assert np.array_equal(x[k], y[k], equal_nan=True), k
