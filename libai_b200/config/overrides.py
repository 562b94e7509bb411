"""Command-line override grammar (``key.path=value``), Hydra "basic override" subset.

The reference delegates to ``hydra.core.override_parser`` (reference: libai/config/lazy.py:390-400).
Hydra is unavailable here, so this is a small recursive-descent parser for the value grammar
exercised by the reference (tests/config/test_lazy_config.py:68-77 and the docs):

    value   := null | bool | int | float | quoted-string | list | dict | bare-string
    list    := '[' value (',' value)* ']'
    dict    := '{' key ':' value (',' key ':' value)* '}'

Prefixes ``+`` / ``++`` (add / force add) are accepted and behave like a plain set; ``~key``
(deletion) raises ``NotImplementedError`` exactly like the reference.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Any, List

_INT = re.compile(r"^[+-]?\d+(_\d+)*$")
_FLOAT = re.compile(
    r"^[+-]?((\d+\.\d*|\.\d+|\d+)([eE][+-]?\d+)?|inf|nan)$", re.IGNORECASE
)


@dataclass
class Override:
    key_or_group: str
    _value: Any
    prefix: str = ""

    def value(self):
        return self._value

    def is_delete(self) -> bool:
        return self.prefix == "~"

    def is_add(self) -> bool:
        return self.prefix in ("+", "++")


class OverrideSyntaxError(ValueError):
    pass


class _ValueParser:
    def __init__(self, text: str):
        self.s = text
        self.i = 0

    def _ws(self):
        while self.i < len(self.s) and self.s[self.i] in " \t":
            self.i += 1

    def parse(self):
        self._ws()
        if self.i >= len(self.s):
            return ""
        v = self._value(top=True)
        self._ws()
        if self.i != len(self.s):
            raise OverrideSyntaxError(f"Unexpected trailing text in override value: {self.s!r}")
        return v

    def _value(self, top=False):
        self._ws()
        if self.i >= len(self.s):
            return ""
        c = self.s[self.i]
        if c == "[":
            return self._list()
        if c == "{":
            return self._dict()
        if c in "\"'":
            return self._quoted()
        return self._bare(top)

    def _list(self):
        self.i += 1
        out = []
        self._ws()
        if self.i < len(self.s) and self.s[self.i] == "]":
            self.i += 1
            return out
        while True:
            out.append(self._value())
            self._ws()
            if self.i >= len(self.s):
                raise OverrideSyntaxError(f"Unterminated list in {self.s!r}")
            if self.s[self.i] == ",":
                self.i += 1
                continue
            if self.s[self.i] == "]":
                self.i += 1
                return out
            raise OverrideSyntaxError(f"Bad list syntax in {self.s!r}")

    def _dict(self):
        self.i += 1
        out = {}
        self._ws()
        if self.i < len(self.s) and self.s[self.i] == "}":
            self.i += 1
            return out
        while True:
            self._ws()
            j = self.i
            while self.i < len(self.s) and self.s[self.i] not in ":,}":
                self.i += 1
            key = self.s[j : self.i].strip().strip("\"'")
            if self.i >= len(self.s) or self.s[self.i] != ":":
                raise OverrideSyntaxError(f"Bad dict syntax in {self.s!r}")
            self.i += 1
            out[key] = self._value()
            self._ws()
            if self.i >= len(self.s):
                raise OverrideSyntaxError(f"Unterminated dict in {self.s!r}")
            if self.s[self.i] == ",":
                self.i += 1
                continue
            if self.s[self.i] == "}":
                self.i += 1
                return out
            raise OverrideSyntaxError(f"Bad dict syntax in {self.s!r}")

    def _quoted(self):
        q = self.s[self.i]
        self.i += 1
        buf = []
        while self.i < len(self.s):
            c = self.s[self.i]
            if c == "\\" and self.i + 1 < len(self.s) and self.s[self.i + 1] in (q, "\\"):
                buf.append(self.s[self.i + 1])
                self.i += 2
                continue
            if c == q:
                self.i += 1
                return "".join(buf)
            buf.append(c)
            self.i += 1
        raise OverrideSyntaxError(f"Unterminated quoted string in {self.s!r}")

    def _bare(self, top):
        j = self.i
        stop = "" if top else ",]}"
        while self.i < len(self.s) and self.s[self.i] not in stop:
            self.i += 1
        return _convert_scalar(self.s[j : self.i].strip())


def _convert_scalar(tok: str):
    low = tok.lower()
    if low in ("null", "~"):
        return None
    if low == "true":
        return True
    if low == "false":
        return False
    if _INT.match(tok):
        return int(tok.replace("_", ""))
    if _FLOAT.match(tok):
        return float(tok)
    return tok


def parse_value(text: str):
    return _ValueParser(text).parse()


def parse_override(text: str) -> Override:
    prefix = ""
    body = text
    if body.startswith("++"):
        prefix, body = "++", body[2:]
    elif body.startswith("+"):
        prefix, body = "+", body[1:]
    elif body.startswith("~"):
        prefix, body = "~", body[1:]
    if "=" not in body:
        if prefix == "~":
            return Override(body, None, prefix)
        raise OverrideSyntaxError(f"Override {text!r} is not of the form key=value")
    key, raw = body.split("=", 1)
    key = key.strip()
    if not key:
        raise OverrideSyntaxError(f"Override {text!r} has an empty key")
    return Override(key, parse_value(raw), prefix)


class OverridesParser:
    """Drop-in for ``hydra.core.override_parser.overrides_parser.OverridesParser``."""

    @classmethod
    def create(cls):
        return cls()

    def parse_overrides(self, overrides: List[str]) -> List[Override]:
        return [parse_override(o) for o in overrides]
