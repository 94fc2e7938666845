"""ctypes loader for libcald_b200.so (the C-ABI boundary, include/cald_b200*.h).

The product path has no CPU fallback: if the shared library is missing or cannot be
loaded this raises, it never degrades to a Python implementation.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcald_b200.so")
_lib = None


class CaldError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CaldError("%s not found - run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(or ./build.sh); there is no CPU fallback" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.cald_ops_last_error.restype = ctypes.c_char_p
    return _lib


def check_ops(rc):
    if rc != 0:
        raise CaldError(lib().cald_ops_last_error().decode("utf-8", "replace"))
