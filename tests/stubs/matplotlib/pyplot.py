import json
import os
import threading
import time

_animations = []
_lines = []
_texts = []


class _Line:
    def __init__(self, x, y):
        self.x, self.y, self.updates = list(x), list(y), 0

    def set_data(self, x, y):
        self.x, self.y = list(x), list(y)
        self.updates += 1


class _Text:
    def __init__(self, s):
        self.s = s

    def set_text(self, s):
        self.s = s


class _Axes:
    def __init__(self):
        self.ylim = None

    def set_ylim(self, a, b):
        self.ylim = (float(a), float(b))


_axes = _Axes()


def subplots():
    return object(), _axes


def plot(x, y, *a, **kw):
    ln = _Line(x, y)
    _lines.append(ln)
    return [ln]


def text(x, y, s):
    t = _Text(s)
    _texts.append(t)
    return t


def show():
    """Drive every FuncAnimation until the other (data) threads are done, then once more for the last frame."""
    frames_seen, distinct = 0, 0
    last = None
    for an in _animations:
        if an.init_func:
            an.init_func()
    t0 = time.time()
    while time.time() - t0 < float(os.environ.get("SQ_MPL_STUB_TIMEOUT", "600")):
        others = [t for t in threading.enumerate() if t is not threading.current_thread() and t is not threading.main_thread()
                  and t.is_alive() and not t.daemon]
        for an in _animations:
            an.func(frames_seen)
        frames_seen += 1
        y = tuple(_lines[0].y) if _lines else None
        if y != last:
            distinct += 1
            last = y
        if not others:
            break
        time.sleep(0.001)
    rep = os.environ.get("SQ_MPL_STUB_REPORT")
    if rep:
        with open(rep, "w") as f:
            json.dump({"updates": frames_seen, "distinct_frames": distinct, "ylim": _axes.ylim,
                       "npoints": len(_lines[0].y) if _lines else 0, "last_y": [float(v) for v in (_lines[0].y if _lines else [])],
                       "text": _texts[0].s if _texts else None}, f)
