"""Run-structured sample store with the API of ``pypmc.tools.History`` (pypmc/tools/_history.py:7-116):
``append(n)`` opens a new run and returns a writable view of its ``(n, dim)`` block, ``h[i]`` / ``h[a:b]`` return
ONE array covering the selected runs (a view: writing to it changes the history), ``len(h)`` counts runs,
``clear()`` forgets everything.  Host bookkeeping only -- the samples a run produced are what the kernels read."""
import numpy as _np


class History(object):
    def __init__(self, dim, prealloc=1):
        self.dim = int(dim)
        assert self.dim == dim, "``dim`` must be an integer"
        self.prealloc = int(prealloc)
        assert self.prealloc == prealloc, "``prealloc`` must be an integer"
        self.clear()

    def clear(self):
        """Delete the history."""
        self._buf = _np.empty((max(self.prealloc, 0), self.dim))
        self._bounds = []          # (start, stop) row range of every run
        self._used = 0

    def __len__(self):
        return len(self._bounds)

    def append(self, new_points_len):
        """Open a run of ``new_points_len`` points and return a reference to its memory."""
        n = int(new_points_len)
        assert n >= 1, "Must at least append one point!"
        start, stop = self._used, self._used + n
        if stop > len(self._buf):       # grow geometrically; earlier views keep pointing at the old block
            grown = _np.empty((max(stop, 2 * len(self._buf)), self.dim))
            grown[:start] = self._buf[:start]
            self._buf = grown
        self._bounds.append((start, stop))
        self._used = stop
        return self._buf[start:stop]

    def __getitem__(self, item):
        picked = self._bounds[item]
        if not picked:
            return _np.array(())
        if isinstance(item, slice):
            if item.step is not None:
                raise NotImplementedError('strided slicing is not supported')
            return self._buf[picked[0][0]:picked[-1][1]]
        return self._buf[picked[0]:picked[1]]
