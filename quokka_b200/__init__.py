"""quokka_b200 -- B200-native execution backend for Quokka's columnar hot path.

Python mirrors the reference's operator protocols (Executor / input reader / partitioner,
QuokkaContext / DataStream); the work is done by hand-written sm_100a kernels in libqk.so
(include/qk.h).  There is no CPU fallback."""
__version__ = "0.1.0"
