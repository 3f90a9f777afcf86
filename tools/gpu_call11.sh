#!/bin/bash
# full GPU suite on the in-tree library (bounded: a hung kernel must not take the box), then A/B of the bench headline
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 tools/ab_bench.sh 2>&1 | grep -v generic_instantiation
