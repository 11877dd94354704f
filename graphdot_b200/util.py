"""Small host helpers used by the kernel front end (reference
graphdot/util/__init__.py:19-46 ``Timer``; graphdot/util/iterable.py
``flatten``/``fold_like``/``replace``)."""
import time
from collections import OrderedDict


def flatten(tree):
    """Depth-first leaves of nested lists/tuples: ((1, 2), 3) -> 1, 2, 3."""
    for item in tree:
        if isinstance(item, (list, tuple)):
            yield from flatten(item)
        else:
            yield item


def fold_like(flat, template):
    """Inverse of ``flatten``: shape ``flat`` like the nested ``template``."""
    flat = list(flat)

    def build(tpl, pos):
        out = []
        for item in tpl:
            if hasattr(item, '__iter__') and not isinstance(item, str):
                sub, pos = build(item, pos)
                out.append(sub)
            else:
                out.append(flat[pos])
                pos += 1
        return tuple(out), pos

    return build(template, 0)[0]


def replace(iterable, old, new):
    for item in iterable:
        yield new if (isinstance(item, type(old)) and item == old) else item


class Timer:
    """tic/toc stopwatch keyed by phase name."""

    _scales = {'s': 1.0, 'ms': 1e3, 'us': 1e6, 'ns': 1e9}

    def __init__(self):
        self.reset()

    def reset(self):
        self.t = OrderedDict()
        self.dt = OrderedDict()

    def tic(self, tag):
        self.t[tag] = time.perf_counter()

    def toc(self, tag):
        self.dt[tag] = time.perf_counter() - self.t.pop(tag)

    def report(self, unit='s'):
        if unit not in self._scales:
            raise ValueError(f'Unknown unit {unit}')
        for tag, dt in self.dt.items():
            print('%9.1f %s on %s' % (dt * self._scales[unit], unit, tag))
