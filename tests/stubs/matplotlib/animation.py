from . import pyplot as _plt


class FuncAnimation:
    def __init__(self, fig, func, frames=None, init_func=None, blit=False, **kw):
        self.func, self.init_func, self.frames = func, init_func, frames
        _plt._animations.append(self)
