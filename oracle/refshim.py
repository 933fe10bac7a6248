"""
Import shim for the UNMODIFIED reference (mittagessen/kraken) in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (kraken_b200/) imports this.
/root/reference does not exist on the GPU box, so this module is only used by
`oracle/make_golden.py` (fixture generation) and by CPU tests that are skipped when
the reference tree is absent.

The reference cannot be imported directly in this image because eight third-party
roots are missing (coremltools, lightning, skimage, shapely, lxml, htrmopo,
torchmetrics, iso639).  They are only touched by CoreML
(de)serialisation, Fabric and polygon geometry - none of which is on the numeric
path - so we serve inert stub modules for exactly those roots.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('KRAKEN_REFERENCE_ROOT', '/root/reference')
_STUB_ROOTS = ('coremltools', 'lightning', 'skimage', 'shapely', 'lxml', 'htrmopo',
               'torchmetrics', 'iso639')


class _Stub(types.ModuleType):
    """A module that is also callable, subclassable and yields further stubs."""
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        full = f'{self.__name__}.{name}'
        mod = sys.modules.get(full)
        if mod is None:
            mod = _Stub(full)
            mod.__spec__ = importlib.machinery.ModuleSpec(full, _Finder(), is_package=True)
            mod.__path__ = []
            sys.modules[full] = mod
        setattr(self, name, mod)
        return mod

    def __call__(self, *a, **k):
        return self

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'kraken'))


def install():
    """Make `import kraken` resolve to the reference tree. Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    missing = []
    for root in _STUB_ROOTS:
        try:
            importlib.import_module(root)
        except Exception:
            missing.append(root)
    if missing:
        sys.meta_path.insert(0, _Finder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
