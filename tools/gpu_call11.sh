#!/bin/bash
# full GPU suite on the in-tree library, then A/B of the bench headline: variants/*.so vs in-tree (same box)
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
tools/ab_bench.sh 2>&1 | grep -v generic_instantiation
