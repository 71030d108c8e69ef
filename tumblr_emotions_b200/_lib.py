"""ctypes binding of libdeepsent.so.

The prototypes are read from include/deepsent.h (the single source of truth for the C ABI), so a
signature change cannot silently desynchronise the Python side.  There is no CPU fallback: if the
library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "deepsent.h")
DEV_HEADER = os.path.join(HERE, "..", "include", "deepsent_dev.h")
LIB_PATH = os.path.join(HERE, "libdeepsent.so")
DEV_LIB_PATH = os.path.join(HERE, "libdeepsent_dev.so")      # product objects + launch-policy overrides + hardware probe

_CTYPE = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float,
}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """{name: (return type, [(c type, arg name), ...])} for every `ds_*` prototype."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int)\s+(ds_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                parsed.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, parsed)
    return protos


def _to_ctype(ctype: str):
    if "*" in ctype:
        return ctypes.c_void_p
    base = ctype.replace("const", "").strip()
    return _CTYPE[base]


class DeepSentLib:
    """Loads the shared library and exposes every `ds_*` entry point as a checked method."""

    def __init__(self, path: str = LIB_PATH, dev: bool = False):
        if not os.path.exists(path):
            raise RuntimeError(
                "libdeepsent.so not found at %s - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % path)
        import torch  # noqa: F401  (loads libcudart into the process before our library resolves it)
        self._dll = ctypes.CDLL(path)
        self.protos = parse_header()
        self.dev = dev
        if dev:
            self.protos.update(parse_header(DEV_HEADER))
        self._dll.ds_last_error.restype = ctypes.c_char_p
        self._dll.ds_last_error.argtypes = []
        for name, (ret, args) in self.protos.items():
            fn = getattr(self._dll, name)          # raises AttributeError if the symbol is not exported
            if name == "ds_last_error":
                continue
            fn.restype = ctypes.c_int
            fn.argtypes = [_to_ctype(t) for t, _ in args]
            if name in ("ds_version", "ds_sm_count", "ds_debug_get", "ds_launch_count", "ds_comm_nccl_version"):
                setattr(self, name[3:], fn)
            else:
                setattr(self, name[3:], self._checked(name, fn))

    def last_error(self) -> str:
        return self._dll.ds_last_error().decode()

    def _checked(self, name, fn):
        def call(*a):
            rc = fn(*a)
            if rc != 0:
                raise RuntimeError("%s failed (%d): %s" % (name, rc, self.last_error()))
        call.__name__ = name
        return call


_LIB = None
_PRODUCT = None
_DEV = None


def lib() -> DeepSentLib:
    """the library every wrapper in ops.py calls: the product build unless `use_dev(True)` switched to the development build"""
    global _LIB, _PRODUCT
    if _LIB is None:
        if os.environ.get("DS_DEV") == "1":
            return use_dev(True)
        _PRODUCT = _PRODUCT or DeepSentLib()
        _LIB = _PRODUCT
    return _LIB


def use_dev(on: bool = True) -> DeepSentLib:
    """Route the wrappers through libdeepsent_dev.so (adds ds_debug_set/get and the probe) or back to the product library.
    The two libraries have separate state: call ops.init(device) after switching (Engine.__init__ does)."""
    global _LIB, _PRODUCT, _DEV
    if on:
        _DEV = _DEV or DeepSentLib(DEV_LIB_PATH, dev=True)
        _LIB = _DEV
    else:
        _PRODUCT = _PRODUCT or DeepSentLib()
        _LIB = _PRODUCT
    return _LIB
