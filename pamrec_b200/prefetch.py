"""Background producer for the host loops: the next feed dict is built (tokenised file -> native batcher -> 19 arrays) while
the device runs the current step.  The reference builds each batch inline between two `sess.run` calls
(sequential_base_model.py:288-296); overlapping the two is invisible to the caller: same batches, same order, and the
`random` calls of an epoch (io/sequential_iterator.py:545,622) still happen before its first batch, in the producer.

Only one producer runs at a time per iterator and the consumer must not draw from Python's `random` while it runs (the model
loops do not).  `depth` bounds the host memory held in flight."""
import queue
import threading

_END = object()


class Prefetcher:
    def __init__(self, iterable, depth=2):
        self._q = queue.Queue(maxsize=max(int(depth), 1))
        self._stop = threading.Event()
        self._exc = None
        self._thread = threading.Thread(target=self._run, args=(iterable,), daemon=True)
        self._thread.start()

    def _put(self, item):
        while not self._stop.is_set():
            try:
                self._q.put(item, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def _run(self, iterable):
        try:
            for item in iterable:
                if not self._put(item):
                    return
        except BaseException as e:          # re-raised in the consumer
            self._exc = e
        finally:
            self._put(_END)

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is _END:
            self._thread.join()
            if self._exc is not None:
                exc, self._exc = self._exc, None
                raise exc
            raise StopIteration
        return item

    def close(self):
        """Stop early (e.g. early stopping breaks out of an epoch): the producer exits at its next put."""
        self._stop.set()
        try:
            while True:
                self._q.get_nowait()
        except queue.Empty:
            pass
        self._thread.join(timeout=5.0)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
