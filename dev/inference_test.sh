#!/usr/bin/env bash
# Generation utilities, pipelines and the inference-side projects (reference dev/inference_test.sh).
set -e
cd "$(dirname "$0")/.."
python -m pytest tests/inference tests/projects -x -q "$@"
