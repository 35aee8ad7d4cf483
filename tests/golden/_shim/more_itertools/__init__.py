"""Stand-in for `more_itertools` (absent here): tnmf/utils/signals.py imports `chunked` only.
Used ONLY by tests/golden/make_golden.py when importing the reference's signal generators."""
from itertools import islice


def chunked(iterable, n):
    it = iter(iterable)
    while True:
        chunk = list(islice(it, n))
        if not chunk:
            return
        yield chunk
