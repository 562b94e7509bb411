#!/usr/bin/env bash
# Lint like CI does (reference dev/linter.sh): isort, black, flake8, clang-format.  Tools that are not installed are skipped.
set -u
cd "$(dirname "$0")/.."
status=0
run() { if command -v "$1" >/dev/null 2>&1; then echo "== $*"; "$@" || status=1; else echo "== $1 not installed, skipped"; fi; }
run isort --check-only --diff libai_b200 configs projects tests tools bench.py
run black --check -l 120 libai_b200 configs projects tests tools bench.py
run flake8 libai_b200 configs projects tests tools bench.py
if command -v clang-format >/dev/null 2>&1; then
  find libai_b200 -name '*.cu' -o -name '*.cuh' -o -name '*.cpp' | xargs clang-format --dry-run -Werror || status=1
else
  echo "== clang-format not installed, skipped"
fi
exit $status
