"""kmg -- host-side Python layer over libkmg.so (B200-native Gram construction for DNA string kernels)."""
from ._cabi import KmgError, lib, last_error  # noqa: F401
