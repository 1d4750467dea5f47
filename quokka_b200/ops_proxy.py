"""Indirection so that the runtime picks up the same `ops` module the executors use (tests swap it for
the CPU shim; the product always resolves to quokka_b200.ops)."""
from . import executors as _x


class _Proxy:
    def __getattr__(self, name):
        return getattr(_x.ops, name)


ops = _Proxy()
