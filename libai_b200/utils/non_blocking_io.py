"""Asynchronous file writes: ``write`` calls are queued to one daemon thread per path.

Spec: reference libai/utils/non_blocking_io.py (``NonBlockingIOManager`` :58, ``NonBlockingIO``
:195, ``NonBlockingBufferedIO`` :291).  One writer thread per distinct path preserves write order
for that path; ``_join`` drains queues, ``_close_thread_pool`` stops workers.
"""
from __future__ import annotations

import io
import logging
import queue
import threading
from dataclasses import dataclass
from typing import IO, Callable, Dict, Optional, Union


@dataclass
class PathData:
    queue: "queue.Queue"
    thread: threading.Thread


class NonBlockingIOManager:
    def __init__(self, buffered: bool = False, executor=None):
        self._path_to_data: Dict[str, PathData] = {}
        self._buffered = buffered
        self._io_cls = NonBlockingBufferedIO if buffered else NonBlockingIO
        self._lock = threading.Lock()

    def _worker(self, q: "queue.Queue") -> None:
        while True:
            fn = q.get()
            try:
                if fn is None:
                    return
                fn()
            except Exception:  # keep the writer alive; surface the problem in the log
                logging.getLogger(__name__).exception("asynchronous IO task failed")
            finally:
                q.task_done()

    def get_non_blocking_io(self, path: str, io_obj, callback_after_file_close: Optional[Callable] = None, buffering: int = -1):
        with self._lock:
            if path not in self._path_to_data:
                q: "queue.Queue" = queue.Queue()
                t = threading.Thread(target=self._worker, args=(q,), daemon=True)
                t.start()
                self._path_to_data[path] = PathData(q, t)
            data = self._path_to_data[path]
        kwargs = {"buffering": buffering} if self._buffered else {}
        return self._io_cls(
            notify_manager=lambda fn: data.queue.put(fn),
            io_obj=io_obj,
            callback_after_file_close=callback_after_file_close,
            **kwargs,
        )

    def _join(self, path: Optional[str] = None) -> bool:
        if path and path not in self._path_to_data:
            raise ValueError(f"{path} has no async IO associated with it. Make sure `opena({path})` is called first.")
        for p in [path] if path else list(self._path_to_data):
            self._path_to_data[p].queue.join()
        return True

    def _close_thread_pool(self) -> bool:
        for data in self._path_to_data.values():
            data.queue.join()
            data.queue.put(None)
        for data in self._path_to_data.values():
            data.thread.join()
        self._path_to_data.clear()
        return True


class NonBlockingIO(io.IOBase):
    """File-like object whose ``write``/``seek``/``truncate``/``close`` are executed in order on
    the manager's writer thread."""

    def __init__(self, notify_manager: Callable, io_obj: Union[IO[str], IO[bytes]], callback_after_file_close: Optional[Callable] = None):
        super().__init__()
        self._notify = notify_manager
        self._io = io_obj
        self._callback = callback_after_file_close
        self._close_called = False

    def readable(self) -> bool:
        return False

    def writable(self) -> bool:
        return True

    def seekable(self) -> bool:
        return True

    def write(self, b) -> None:
        self._notify(lambda: self._io.write(b))

    def seek(self, offset: int, whence: int = 0) -> int:
        self._notify(lambda: self._io.seek(offset, whence))
        return 0

    def tell(self) -> int:
        raise ValueError("ioPath async writes does not support `tell` calls.")

    def truncate(self, size: int = None) -> int:
        self._notify(lambda: self._io.truncate(size) if size is not None else self._io.truncate())
        return 0

    def close(self) -> None:
        if self._close_called:
            return
        self._close_called = True

        def _finish():
            self._io.close()
            if self._callback:
                self._callback()

        self._notify(_finish)

    @property
    def closed(self):
        return self._close_called


class NonBlockingBufferedIO(NonBlockingIO):
    """Variant that accumulates writes in memory and hands them to the writer in chunks."""

    MAX_BUFFER_BYTES = 10 * 1024 * 1024

    def __init__(self, notify_manager, io_obj, callback_after_file_close=None, buffering: int = -1):
        super().__init__(notify_manager, io_obj, callback_after_file_close)
        self._limit = self.MAX_BUFFER_BYTES if buffering <= 0 else buffering
        self._chunks = []
        self._size = 0

    def write(self, b) -> None:
        self._chunks.append(b)
        self._size += len(b)
        if self._size >= self._limit:
            self.flush()

    def flush(self) -> None:
        if not self._chunks:
            return
        chunks, self._chunks, self._size = self._chunks, [], 0
        joined = (b"" if isinstance(chunks[0], (bytes, bytearray)) else "").join(chunks)
        self._notify(lambda: self._io.write(joined))

    def close(self) -> None:
        self.flush()
        super().close()
