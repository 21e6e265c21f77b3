"""slow5tools_b200 -- B200-native BLOW5 per-record codec (svb-zd + zlib) behind the slow5lib API.

The product is the C-ABI library ``libslow5b200.so`` (``include/slow5b200.h``), built in-tree from
``slow5tools_b200/csrc`` for sm_100a.  This package is the thin Python host mirror used by the tests
and ``bench.py``: it binds the C-ABI with ctypes and hands it torch device memory and streams.
There is no CPU fallback anywhere in this package; importing works without a GPU (so the build and
symbol checks can run), every compute call needs one.
"""
from ._capi import lib, S5BError, ERR, strerror, library_path  # noqa: F401
from .codec import Codec, sig_layout, svb_slot_layout  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["lib", "S5BError", "ERR", "strerror", "library_path", "Codec", "sig_layout",
           "svb_slot_layout", "synth"]
