"""Minimal stand-in for the `bitarray` package, ONLY for tests/golden/make_iou_golden.py.

The reference's eval_script.py imports bitarray (not installed in the build container); its IoU code uses exactly:
bitarray(), bitarray('0' * k), append, insert(0, v), slice assignment a[i:] = other, len, &, |, count().  This list-of-ints
class provides those and nothing else.  Test infrastructure for generating golden vectors from the REAL reference; never
imported by the product or by the tests themselves.
"""


class bitarray:
    def __init__(self, init=None):
        if init is None:
            self.b = []
        elif isinstance(init, str):
            self.b = [1 if c == "1" else 0 for c in init]
        else:
            self.b = [1 if v else 0 for v in init]

    def append(self, v):
        self.b.append(1 if v else 0)

    def insert(self, i, v):
        self.b.insert(i, 1 if v else 0)

    def __len__(self):
        return len(self.b)

    def __setitem__(self, key, value):
        if isinstance(key, slice):
            self.b[key] = list(value.b)
        else:
            self.b[key] = 1 if value else 0

    def __getitem__(self, key):
        if isinstance(key, slice):
            return bitarray(self.b[key])
        return self.b[key]

    def _check(self, other):
        if len(self.b) != len(other.b):
            raise ValueError("bitarrays of equal length expected")

    def __and__(self, other):
        self._check(other)
        return bitarray([x & y for x, y in zip(self.b, other.b)])

    def __or__(self, other):
        self._check(other)
        return bitarray([x | y for x, y in zip(self.b, other.b)])

    def count(self, v=1):
        return sum(1 for x in self.b if x == (1 if v else 0))

    def tolist(self):
        return list(self.b)
