"""Logging setup: colored stdout on rank 0, a ``log.txt[.rankN]`` file on every rank, and
rate-limited helpers (spec: reference libai/utils/logger.py:56-214). No termcolor dependency:
ANSI codes are emitted directly."""
import atexit
import functools
import logging
import os
import sys
import time
from collections import Counter

_ANSI = {"red": "\033[31m", "green": "\033[32m", "blink": "\033[5m", "underline": "\033[4m", "end": "\033[0m"}


def colored(text, color=None, attrs=()):
    pre = _ANSI.get(color, "") + "".join(_ANSI.get(a, "") for a in attrs)
    return f"{pre}{text}{_ANSI['end']}" if pre else text


class _ColorfulFormatter(logging.Formatter):
    def __init__(self, *args, root_name="", abbrev_name="", **kwargs):
        self._root = root_name + "."
        self._abbrev = (abbrev_name + ".") if abbrev_name else ""
        super().__init__(*args, **kwargs)

    def formatMessage(self, record):
        record.name = record.name.replace(self._root, self._abbrev)
        msg = super().formatMessage(record)
        if record.levelno == logging.WARNING:
            return colored("WARNING", "red", ("blink",)) + " " + msg
        if record.levelno >= logging.ERROR:
            return colored("ERROR", "red", ("blink", "underline")) + " " + msg
        return msg


@functools.lru_cache(maxsize=None)
def _cached_log_stream(filename):
    stream = open(filename, "a", buffering=1)
    atexit.register(stream.close)
    return stream


@functools.lru_cache()
def setup_logger(output=None, distributed_rank=0, *, color=True, name="libai_b200", abbrev_name=None):
    """Configure and return the ``name`` logger.

    ``output`` may be a directory (→ ``<dir>/log.txt``) or a ``.txt/.log`` file; non-zero ranks
    append ``.rank{N}`` to the file name and do not log to stdout.
    """
    logger = logging.getLogger(name)
    logger.setLevel(logging.DEBUG)
    logger.propagate = False
    if abbrev_name is None:
        abbrev_name = "lb" if name == "libai_b200" else name
    plain = logging.Formatter("[%(asctime)s] %(name)s %(levelname)s: %(message)s", datefmt="%m/%d %H:%M:%S")
    if distributed_rank == 0:
        sh = logging.StreamHandler(stream=sys.stdout)
        sh.setLevel(logging.DEBUG)
        if color and sys.stdout.isatty():
            sh.setFormatter(
                _ColorfulFormatter(
                    colored("[%(asctime)s %(name)s]: ", "green") + "%(message)s",
                    datefmt="%m/%d %H:%M:%S",
                    root_name=name,
                    abbrev_name=str(abbrev_name),
                )
            )
        else:
            sh.setFormatter(plain)
        logger.addHandler(sh)
    if output is not None:
        filename = output if output.endswith((".txt", ".log")) else os.path.join(output, "log.txt")
        if distributed_rank > 0:
            filename += f".rank{distributed_rank}"
        os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
        fh = logging.StreamHandler(_cached_log_stream(filename))
        fh.setLevel(logging.DEBUG)
        fh.setFormatter(plain)
        logger.addHandler(fh)
    return logger


def _find_caller():
    frame = sys._getframe(2)
    while frame:
        code = frame.f_code
        if os.path.join("utils", "logger.") not in code.co_filename:
            mod = frame.f_globals.get("__name__", "libai_b200")
            if mod == "__main__":
                mod = "libai_b200"
            return mod, (code.co_filename, frame.f_lineno, code.co_name)
        frame = frame.f_back
    return "libai_b200", ("", 0, "")


_LOG_COUNTER = Counter()
_LOG_TIMER = {}


def log_first_n(lvl, msg, n=1, *, name=None, key="caller"):
    """Log only the first ``n`` times for a given ``key`` ("caller", "message" or both)."""
    if isinstance(key, str):
        key = (key,)
    assert len(key) > 0
    caller_module, caller_key = _find_caller()
    hash_key = ()
    if "caller" in key:
        hash_key += caller_key
    if "message" in key:
        hash_key += (msg,)
    _LOG_COUNTER[hash_key] += 1
    if _LOG_COUNTER[hash_key] <= n:
        logging.getLogger(name or caller_module).log(lvl, msg)


def log_every_n(lvl, msg, n=1, *, name=None):
    caller_module, key = _find_caller()
    _LOG_COUNTER[key] += 1
    if n == 1 or _LOG_COUNTER[key] % n == 1:
        logging.getLogger(name or caller_module).log(lvl, msg)


def log_every_n_seconds(lvl, msg, n=1, *, name=None):
    caller_module, key = _find_caller()
    last = _LOG_TIMER.get(key)
    now = time.time()
    if last is None or now - last >= n:
        logging.getLogger(name or caller_module).log(lvl, msg)
        _LOG_TIMER[key] = now
