#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q -k "oracle_same_samples or baseline_configs or edge_cases or mutated or row_chunking or full_size_properties" 2>&1 | tail -4
timeout 400 tools/ab_bench.sh 2>&1 | grep -v generic_instantiation
