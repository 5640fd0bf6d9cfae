"""Import hygiene for the reference's modules.

`pterotactyl/utility/utils.py:14-25` imports matplotlib, pytorch3d, and (through pretty_render.py:11-13) trimesh
and pyrender at module scope; the trainers add submitit (`vision/train.py:15`), the simulator pybullet.  None of
them is needed by the reconstruction hot path, but without them `import pterotactyl.utility.utils` fails before a
single kernel runs.  `install_import_stubs()` registers a meta-path finder that serves *import-only* placeholders
for whichever of those packages is genuinely absent:

  * `import matplotlib.pyplot as plt`, `from submitit.helpers import Checkpointable`, `import pyrender` all succeed;
  * a placeholder class can be subclassed (`class Engine(Checkpointable)`), because the trainers do that at
    import time;
  * *using* a placeholder (calling `plt.figure()`, instantiating `pyrender.Scene`) raises ImportError naming the
    package that is really missing -- nothing is silently faked.

A package that is installed is never shadowed: the finder is appended to `sys.meta_path`, so it only answers when
every regular finder has failed.
"""
import importlib.abc
import importlib.machinery
import sys
import types

# third-party packages the reference imports that the reconstruction path never calls
STUBBABLE = ("matplotlib", "trimesh", "pyrender", "pybullet", "pybullet_utils", "pybullet_data", "submitit", "rtree",
             "skimage")


class _Placeholder:
    """Base of every attribute a stub module hands out.  Subclassing is allowed; direct use raises."""
    _ptk_missing = "?"

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)

    def __new__(cls, *a, **k):
        if "_ptk_stub_itself" in cls.__dict__:
            raise ImportError(f"{cls._ptk_missing} is not installed (ptk_b200.import_stubs only lets it be imported)")
        return super().__new__(cls)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        ph = type(name, (_Placeholder,), {"_ptk_missing": full, "_ptk_stub_itself": True, "__module__": self.__name__})
        setattr(self, name, ph)
        return ph


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, names):
        self.names = set(names)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.names:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []  # a package: `import matplotlib.pyplot` resolves through this finder again
        mod.__ptk_stub__ = True
        return mod

    def exec_module(self, module):
        pass


_finder = None


def install_import_stubs(names=STUBBABLE):
    """Make the reference's non-hot-path third-party imports resolve.  Returns the names that are served by stubs
    (i.e. not installed).  Idempotent."""
    global _finder
    if _finder is None:
        _finder = _StubFinder(())
        sys.meta_path.append(_finder)  # last: real packages always win
    _finder.names |= set(names)
    import importlib.util
    stubbed = []
    for n in sorted(_finder.names):
        spec = importlib.util.find_spec(n)
        if spec is not None and spec.loader is _finder:
            stubbed.append(n)
    return stubbed


def is_stub(module):
    return bool(getattr(module, "__ptk_stub__", False))
