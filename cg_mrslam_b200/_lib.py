"""Loader for the C-ABI library (include/*.h). There is no Python or CPU fallback: if the CUDA
library has not been built the import of any compute entry point fails loudly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcgmrslam_b200.so")

_cache = {}


class BuildError(RuntimeError):
    pass


def load(path=None):
    """dlopen the product library (or, for the host-logic unit tests only, an explicit path)."""
    path = path or LIB_PATH
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise BuildError(
            "%s is missing: build it with `make -C cg_mrslam_b200/csrc` (or "
            "`python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    _cache[path] = lib
    return lib
