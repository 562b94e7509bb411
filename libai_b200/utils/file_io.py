"""Path abstraction with pluggable handlers (local files, http(s) URLs, lazily-resolved paths).

Spec: reference libai/utils/file_io.py (an iopath-style ``PathManagerBase`` with
``NativePathHandler`` :504, ``HTTPURLHandler`` :718, ``OneDrivePathHandler`` :776, ``LazyPath``
:98, ``file_lock`` :70, ``PathManagerFactory`` :1277) and libai/utils/non_blocking_io.py
(asynchronous ``opena``).  This implementation keeps the same entry points with a much smaller
core: handlers are matched by prefix, the native handler is the fallback, and asynchronous
writes go through one writer thread per path (see :mod:`non_blocking_io`).
"""
from __future__ import annotations

import base64
import errno
import logging
import os
import shutil
import tempfile
import threading
from collections import OrderedDict
from contextlib import contextmanager
from typing import IO, Any, Callable, Dict, List, MutableMapping, Optional, Union
from urllib.parse import urlparse

__all__ = [
    "LazyPath",
    "PathHandler",
    "NativePathHandler",
    "HTTPURLHandler",
    "OneDrivePathHandler",
    "PathManagerBase",
    "PathManagerFactory",
    "PathManager",
    "get_cache_dir",
    "file_lock",
]


def get_cache_dir(cache_dir: Optional[str] = None) -> str:
    """Default ``$LIBAI_CACHE`` or ``~/.cache/libai_b200``; falls back to a temp dir if unwritable."""
    if cache_dir is None:
        cache_dir = os.path.expanduser(os.getenv("LIBAI_CACHE", "~/.cache/libai_b200"))
    try:
        os.makedirs(cache_dir, exist_ok=True)
        assert os.access(cache_dir, os.W_OK)
    except (OSError, AssertionError):
        cache_dir = os.path.join(tempfile.gettempdir(), "libai_b200_cache")
        os.makedirs(cache_dir, exist_ok=True)
    return cache_dir


@contextmanager
def file_lock(path: str):
    """Inter-process lock on ``path + '.lock'`` (fcntl based; portalocker is not required)."""
    dirname = os.path.dirname(path)
    if dirname:
        try:
            os.makedirs(dirname, exist_ok=True)
        except OSError:
            pass
    lock_path = path + ".lock"
    try:
        import fcntl

        with open(lock_path, "w") as fh:
            fcntl.flock(fh, fcntl.LOCK_EX)
            try:
                yield
            finally:
                fcntl.flock(fh, fcntl.LOCK_UN)
    except ImportError:  # pragma: no cover - non POSIX
        yield


class LazyPath(os.PathLike):
    """A path whose value is produced by ``func`` on first use (e.g. a download)."""

    def __init__(self, func: Callable[[], str]) -> None:
        self._func = func
        self._value: Optional[str] = None

    def _get(self) -> str:
        if self._value is None:
            self._value = self._func()
        return self._value

    def __fspath__(self) -> str:
        return self._get()

    def __getattr__(self, name: str):
        if name in ("_func", "_value"):
            raise AttributeError(name)
        if self._value is None:
            raise AttributeError(f"Uninitialized LazyPath has no attribute: {name}.")
        return getattr(self._value, name)

    def __getitem__(self, key):
        if self._value is None:
            raise TypeError("Uninitialized LazyPath is not subscriptable.")
        return self._value[key]

    def __repr__(self) -> str:
        return "LazyPath(value={})".format(self._value) if self._value else f"LazyPath(func={self._func})"

    def __str__(self) -> str:
        return self._get()

    def __eq__(self, other):
        return isinstance(other, (str, LazyPath)) and str(self) == str(other)

    def __hash__(self):
        return hash(self._get())


class PathHandler:
    """Operations for one family of URIs; subclasses override what they support."""

    _strict_kwargs_check = True

    def _check_kwargs(self, kwargs: Dict[str, Any]) -> None:
        if self._strict_kwargs_check:
            if kwargs:
                raise ValueError(f"Unused arguments: {kwargs}")
        elif kwargs:
            logging.getLogger(__name__).warning(f"[PathManager] unused arguments: {kwargs}")

    def _get_supported_prefixes(self) -> List[str]:
        raise NotImplementedError()

    def _get_local_path(self, path: str, force: bool = False, **kwargs) -> str:
        raise NotImplementedError()

    def _copy_from_local(self, local_path: str, dst_path: str, overwrite: bool = False, **kwargs) -> bool:
        raise NotImplementedError()

    def _open(self, path: str, mode: str = "r", buffering: int = -1, **kwargs):
        raise NotImplementedError()

    def _opena(self, path: str, mode: str = "r", buffering: int = -1, **kwargs):
        raise NotImplementedError()

    def _copy(self, src_path: str, dst_path: str, overwrite: bool = False, **kwargs) -> bool:
        raise NotImplementedError()

    def _mv(self, src_path: str, dst_path: str, **kwargs) -> bool:
        raise NotImplementedError()

    def _exists(self, path: str, **kwargs) -> bool:
        raise NotImplementedError()

    def _isfile(self, path: str, **kwargs) -> bool:
        raise NotImplementedError()

    def _isdir(self, path: str, **kwargs) -> bool:
        raise NotImplementedError()

    def _ls(self, path: str, **kwargs) -> List[str]:
        raise NotImplementedError()

    def _mkdirs(self, path: str, **kwargs) -> None:
        raise NotImplementedError()

    def _rm(self, path: str, **kwargs) -> None:
        raise NotImplementedError()

    def _symlink(self, src_path: str, dst_path: str, **kwargs) -> bool:
        raise NotImplementedError()

    def _set_cwd(self, path: Optional[str], **kwargs) -> bool:
        raise NotImplementedError()

    def _async_join(self, path: Optional[str] = None, **kwargs) -> bool:
        return True

    def _async_close(self, **kwargs) -> bool:
        return True


class NativePathHandler(PathHandler):
    """Local filesystem."""

    _cwd = None

    def __init__(self):
        self._io_manager = None

    def _abs(self, path) -> str:
        path = os.fspath(path)
        return os.path.normpath(path if not self._cwd else os.path.join(self._cwd, path))

    def _get_local_path(self, path, force=False, **kwargs):
        self._check_kwargs(kwargs)
        return os.fspath(path)

    def _copy_from_local(self, local_path, dst_path, overwrite=False, **kwargs):
        self._check_kwargs(kwargs)
        return self._copy(local_path, dst_path, overwrite)

    def _open(self, path, mode="r", buffering=-1, encoding=None, errors=None, newline=None, closefd=True, opener=None, **kwargs):
        self._check_kwargs(kwargs)
        return open(self._abs(path), mode, buffering=buffering, encoding=encoding, errors=errors, newline=newline, closefd=closefd, opener=opener)

    def _opena(self, path, mode="r", buffering=-1, callback_after_file_close=None, **kwargs):
        from .non_blocking_io import NonBlockingIOManager

        if self._io_manager is None:
            self._io_manager = NonBlockingIOManager(buffered=False)
        self._check_kwargs(kwargs)
        return self._io_manager.get_non_blocking_io(
            path=self._abs(path), io_obj=self._open(path, mode), callback_after_file_close=callback_after_file_close
        )

    def _async_join(self, path=None, **kwargs):
        if self._io_manager is None:
            return True
        return self._io_manager._join(self._abs(path) if path else None)

    def _async_close(self, **kwargs):
        if self._io_manager is None:
            return True
        return self._io_manager._close_thread_pool()

    def _copy(self, src_path, dst_path, overwrite=False, **kwargs):
        self._check_kwargs(kwargs)
        src, dst = self._abs(src_path), self._abs(dst_path)
        if os.path.exists(dst) and not overwrite:
            logging.getLogger(__name__).error(f"Destination file {dst} already exists.")
            return False
        try:
            shutil.copyfile(src, dst)
            return True
        except Exception as e:
            logging.getLogger(__name__).error(f"Error in file copy - {e}")
            return False

    def _mv(self, src_path, dst_path, **kwargs):
        self._check_kwargs(kwargs)
        src, dst = self._abs(src_path), self._abs(dst_path)
        if os.path.exists(dst):
            logging.getLogger(__name__).error(f"Destination file {dst} already exists.")
            return False
        try:
            shutil.move(src, dst)
            return True
        except Exception as e:
            logging.getLogger(__name__).error(f"Error in move operation - {e}")
            return False

    def _symlink(self, src_path, dst_path, **kwargs):
        self._check_kwargs(kwargs)
        src, dst = self._abs(src_path), self._abs(dst_path)
        if os.path.exists(dst):
            return False
        try:
            os.symlink(src, dst)
            return True
        except Exception:
            return False

    def _exists(self, path, **kwargs):
        self._check_kwargs(kwargs)
        return os.path.exists(self._abs(path))

    def _isfile(self, path, **kwargs):
        self._check_kwargs(kwargs)
        return os.path.isfile(self._abs(path))

    def _isdir(self, path, **kwargs):
        self._check_kwargs(kwargs)
        return os.path.isdir(self._abs(path))

    def _ls(self, path, **kwargs):
        self._check_kwargs(kwargs)
        return os.listdir(self._abs(path))

    def _mkdirs(self, path, **kwargs):
        self._check_kwargs(kwargs)
        try:
            os.makedirs(self._abs(path), exist_ok=True)
        except OSError as e:
            if e.errno != errno.EEXIST:
                raise

    def _rm(self, path, **kwargs):
        self._check_kwargs(kwargs)
        os.remove(self._abs(path))

    def _set_cwd(self, path, **kwargs):
        self._check_kwargs(kwargs)
        if path is None:
            self._cwd = None
            return True
        if not os.path.exists(path):
            raise ValueError(f"{path} is not a valid Unix path")
        self._cwd = path
        return True


class HTTPURLHandler(PathHandler):
    """Downloads http/https/ftp URLs into the cache directory and serves the local copy."""

    MAX_FILENAME_LEN = 250

    def __init__(self) -> None:
        self.cache_map: Dict[str, str] = {}

    def _get_supported_prefixes(self):
        return ["http://", "https://", "ftp://"]

    def _get_local_path(self, path, force=False, **kwargs):
        self._check_kwargs(kwargs)
        if force or path not in self.cache_map or not os.path.exists(self.cache_map[path]):
            from .download import download

            parsed = urlparse(path)
            dirname = os.path.join(get_cache_dir(), os.path.dirname(parsed.path.lstrip("/")))
            filename = path.split("/")[-1]
            if parsed.query:
                filename = filename.split("?").pop(0)
            if len(filename) > self.MAX_FILENAME_LEN:
                filename = filename[:100] + "_" + base64.urlsafe_b64encode(filename.encode()).decode()[:50]
            cached = os.path.join(dirname, filename)
            with file_lock(cached):
                if not os.path.isfile(cached):
                    logging.getLogger(__name__).info(f"Downloading {path} ...")
                    cached = download(path, dirname, filename=filename)
            self.cache_map[path] = cached
        return self.cache_map[path]

    def _open(self, path, mode="r", buffering=-1, **kwargs):
        self._check_kwargs(kwargs)
        assert mode in ("r", "rb"), f"{type(self).__name__} does not support open with {mode} mode"
        assert buffering == -1, f"{type(self).__name__} does not support the `buffering` argument"
        return open(self._get_local_path(path, force=False), mode)


class OneDrivePathHandler(HTTPURLHandler):
    """OneDrive share links → direct download URLs."""

    ONE_DRIVE_PREFIX = "https://1drv.ms/u/s!"

    def create_one_drive_direct_download(self, one_drive_url: str) -> str:
        data = base64.b64encode(bytes(one_drive_url, "utf-8")).decode("utf-8")
        data = data.replace("/", "_").replace("+", "-").rstrip("=")
        return f"https://api.onedrive.com/v1.0/shares/u!{data}/root/content"

    def _get_supported_prefixes(self):
        return [self.ONE_DRIVE_PREFIX]

    def _get_local_path(self, path, force=False, **kwargs):
        return super()._get_local_path(self.create_one_drive_direct_download(path), force=force, **kwargs)


class PathManagerBase:
    """Dispatches every operation to the handler whose prefix matches the path."""

    def __init__(self) -> None:
        self._path_handlers: MutableMapping[str, PathHandler] = OrderedDict()
        self._native_path_handler: PathHandler = NativePathHandler()
        self._cwd: Optional[str] = None
        self._async_handlers = set()

    def __get_path_handler(self, path: Union[str, os.PathLike]) -> PathHandler:
        path = os.fspath(path)
        for prefix, handler in self._path_handlers.items():
            if path.startswith(prefix):
                return handler
        return self._native_path_handler

    _handler = __get_path_handler

    def open(self, path: str, mode: str = "r", buffering: int = -1, **kwargs) -> IO:
        return self.__get_path_handler(path)._open(path, mode, buffering=buffering, **kwargs)

    def opent(self, path: str, mode: str = "r", buffering: int = 32, **kwargs):
        return self.open(path, mode, **kwargs)

    def opena(self, path: str, mode: str = "r", buffering: int = -1, callback_after_file_close=None, **kwargs):
        """Non-blocking open for *writing*; writes are queued to a background thread."""
        if "w" not in mode and "a" not in mode:
            raise ValueError("`opena` mode must be write or append")
        handler = self.__get_path_handler(path)
        fh = handler._opena(path, mode, buffering=buffering, callback_after_file_close=callback_after_file_close, **kwargs)
        self._async_handlers.add(handler)
        return fh

    def async_join(self, *paths: str, **kwargs) -> bool:
        ok = True
        if not paths:
            for h in list(self._async_handlers):
                ok &= h._async_join(**kwargs)
        else:
            for p in paths:
                ok &= self.__get_path_handler(p)._async_join(p, **kwargs)
        return ok

    def async_close(self, **kwargs) -> bool:
        ok = self.async_join(**kwargs)
        for h in list(self._async_handlers):
            ok &= h._async_close(**kwargs)
        self._async_handlers.clear()
        return ok

    def copy(self, src_path, dst_path, overwrite=False, **kwargs) -> bool:
        h = self.__get_path_handler(src_path)
        assert h == self.__get_path_handler(dst_path), "copy across different path handlers is unsupported"
        return h._copy(src_path, dst_path, overwrite, **kwargs)

    def mv(self, src_path, dst_path, **kwargs) -> bool:
        h = self.__get_path_handler(src_path)
        assert h == self.__get_path_handler(dst_path), "mv across different path handlers is unsupported"
        return h._mv(src_path, dst_path, **kwargs)

    def get_local_path(self, path, force=False, **kwargs) -> str:
        path = os.fspath(path)
        return self.__get_path_handler(path)._get_local_path(path, force=force, **kwargs)

    def copy_from_local(self, local_path, dst_path, overwrite=False, **kwargs) -> bool:
        assert os.path.exists(local_path), f"local_path = {local_path}"
        return self.__get_path_handler(dst_path)._copy_from_local(local_path, dst_path, overwrite=overwrite, **kwargs)

    def exists(self, path, **kwargs) -> bool:
        return self.__get_path_handler(path)._exists(path, **kwargs)

    def isfile(self, path, **kwargs) -> bool:
        return self.__get_path_handler(path)._isfile(path, **kwargs)

    def isdir(self, path, **kwargs) -> bool:
        return self.__get_path_handler(path)._isdir(path, **kwargs)

    def ls(self, path, **kwargs) -> List[str]:
        return self.__get_path_handler(path)._ls(path, **kwargs)

    def mkdirs(self, path, **kwargs) -> None:
        return self.__get_path_handler(path)._mkdirs(path, **kwargs)

    def rm(self, path, **kwargs) -> None:
        return self.__get_path_handler(path)._rm(path, **kwargs)

    def symlink(self, src_path, dst_path, **kwargs) -> bool:
        h = self.__get_path_handler(src_path)
        assert h == self.__get_path_handler(dst_path)
        return h._symlink(src_path, dst_path, **kwargs)

    def set_cwd(self, path: Optional[str], **kwargs) -> bool:
        if path is None and self._cwd is None:
            return True
        if self.__get_path_handler(path or self._cwd)._set_cwd(path, **kwargs):
            self._cwd = path
            return True
        return False

    def register_handler(self, handler: PathHandler, allow_override: bool = True) -> None:
        assert isinstance(handler, PathHandler), handler
        for prefix in handler._get_supported_prefixes():
            if prefix in self._path_handlers and not allow_override:
                raise KeyError(f"Prefix '{prefix}' already registered by {self._path_handlers[prefix]}!")
            self._path_handlers[prefix] = handler
        # longest prefix first so that more specific handlers win
        self._path_handlers = OrderedDict(sorted(self._path_handlers.items(), key=lambda t: t[0], reverse=True))

    def set_strict_kwargs_checking(self, enable: bool) -> None:
        self._native_path_handler._strict_kwargs_check = enable
        for h in self._path_handlers.values():
            h._strict_kwargs_check = enable


class PathManagerFactory:
    """Named ``PathManagerBase`` singletons."""

    GLOBAL_PATH_MANAGER = "global_path_manager"
    pm_list: Dict[str, PathManagerBase] = {}
    _lock = threading.Lock()

    @staticmethod
    def get(key: str = GLOBAL_PATH_MANAGER) -> PathManagerBase:
        with PathManagerFactory._lock:
            if key not in PathManagerFactory.pm_list:
                PathManagerFactory.pm_list[key] = PathManagerBase()
            return PathManagerFactory.pm_list[key]

    @staticmethod
    def remove(key: str) -> None:
        with PathManagerFactory._lock:
            PathManagerFactory.pm_list.pop(key, None)


PathManager = PathManagerFactory.get()
PathManager.register_handler(HTTPURLHandler())
PathManager.register_handler(OneDrivePathHandler())
