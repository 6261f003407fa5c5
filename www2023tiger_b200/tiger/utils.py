"""One-batch-ahead prefetching (reference: tiger/utils.py:33-57, used at train_self_supervised.py:135).

The collator of this package launches its kernels from the prefetch thread, so batch i+1's neighbor
search overlaps batch i's model step exactly as the reference's host-side sampler did."""
import queue
import threading


class BackgroundThreadGenerator:
    """Iterates `generator` in a daemon thread, keeping at most `max_prefetch` items ready."""

    _END = object()

    def __init__(self, generator, max_prefetch: int = 1):
        self.generator = generator
        self.queue = queue.Queue(max_prefetch)
        self.error = None
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        try:
            for item in self.generator:
                self.queue.put(item)
        except BaseException as exc:  # surfaced in the consumer thread
            self.error = exc
        self.queue.put(self._END)

    def __iter__(self):
        return self

    def __next__(self):
        item = self.queue.get()
        if item is self._END:
            if self.error is not None:
                raise self.error
            raise StopIteration
        return item

    def __len__(self):
        return len(self.generator)
