"""Stand-ins for the third-party packages the reference's SLAM layer imports and this image lacks (faiss, g2o, matplotlib,
skimage, colour_demosaicing), so that `slam/*.py` and `datasets/*.py` can be imported UNMODIFIED from /root/reference in the
build container.  Test infrastructure only."""
import sys
import types
from pathlib import Path

REFERENCE = Path('/root/reference')
NAMES = ('matplotlib', 'matplotlib.pyplot', 'matplotlib.cm', 'matplotlib.colors', 'colour_demosaicing', 'skimage',
         'skimage.transform', 'faiss', 'g2o')


class Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        child = Stub(f'{self.__name__}.{name}')
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return None

    def __mro_entries__(self, bases):          # `class X(g2o.SparseOptimizer)` in slam/pose_graph_optimization.py
        return (object,)


def install():
    """put /root/reference on sys.path and stub what cannot be imported; returns the list of stubbed module names"""
    if str(REFERENCE) not in sys.path:
        sys.path.insert(0, str(REFERENCE))
    added = []
    for name in NAMES:
        try:
            __import__(name)
        except Exception:
            stub = Stub(name)
            stub.__path__ = []          # importable as a package (`import matplotlib.cm`)
            sys.modules[name] = stub
            added.append(name)
    return added


def remove(added):
    for name in added:
        sys.modules.pop(name, None)
