#!/usr/bin/env bash
# CPU tier (reference dev/run_unittest.sh): everything not marked gpu, including the 2-process gloo tests.
set -e
cd "$(dirname "$0")/.."
python -m pytest tests/ -x -q -m "not gpu" "$@"
