#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests -m gpu -x -q -k "checkercylinder or mutated" 2>&1 | tail -2
timeout 300 tools/ab_bench.sh 2>&1 | grep -v generic
