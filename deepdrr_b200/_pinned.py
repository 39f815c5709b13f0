"""Pool of page-locked host buffers for the images ``Projector.project`` returns.

The reference hands back whatever ``cupy.ndarray.get()`` allocates (pageable memory,
deepdrr/projector/projector.py:786-787).  A device-to-host copy into pageable memory is staged through
the driver's bounce buffers at a fraction of the PCIe rate, so the images are written into page-locked
blocks (``drr_host_alloc``) instead.  A block is wrapped in a NumPy array; when the last view of that
array dies the block returns to the pool, so a loop ``img = projector(*poses)`` cycles through two
blocks.  ``max_outstanding_bytes`` bounds what a caller who keeps every result alive can pin; beyond it
``take`` returns None and the caller falls back to a plain ``np.empty``.
"""
from __future__ import annotations

import ctypes
import threading
import weakref
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib


class PinnedPool:
    def __init__(self, max_outstanding_bytes: int = 2 << 30, keep_free: int = 3):
        self.max_outstanding_bytes = int(max_outstanding_bytes)
        self.keep_free = int(keep_free)
        self._free: Dict[int, List[int]] = {}
        self._outstanding = 0
        self._closed = False
        self._lock = threading.Lock()

    def take(self, shape: Tuple[int, ...], dtype=np.float32) -> Optional[np.ndarray]:
        """A C-contiguous array of ``shape`` in page-locked memory, or None when the pool is exhausted / closed."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if nbytes == 0:
            return None
        with self._lock:
            if self._closed or self._outstanding + nbytes > self.max_outstanding_bytes:
                return None
            blocks = self._free.get(nbytes)
            ptr = blocks.pop() if blocks else None
            self._outstanding += nbytes
        if ptr is None:
            p = ctypes.c_void_p()
            if _lib.load().drr_host_alloc(nbytes, ctypes.byref(p)) != _lib.OK or not p.value:
                with self._lock:
                    self._outstanding -= nbytes
                return None
            ptr = int(p.value)
        buf = (ctypes.c_byte * nbytes).from_address(ptr)
        weakref.finalize(buf, self._give_back, ptr, nbytes)  # runs when the last array viewing `buf` is collected
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def _give_back(self, ptr: int, nbytes: int) -> None:
        with self._lock:
            self._outstanding -= nbytes
            blocks = self._free.setdefault(nbytes, [])
            if not self._closed and len(blocks) < self.keep_free:
                blocks.append(ptr)
                return
        try:
            _lib.load().drr_host_free(ctypes.c_void_p(ptr))
        except Exception:  # interpreter shutdown
            pass

    def close(self) -> None:
        """Free the idle blocks; blocks still held by live arrays are freed when those arrays die."""
        with self._lock:
            self._closed = True
            free, self._free = self._free, {}
        for blocks in free.values():
            for ptr in blocks:
                try:
                    _lib.load().drr_host_free(ctypes.c_void_p(ptr))
                except Exception:
                    pass
