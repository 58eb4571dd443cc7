#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_next_rows.py tests/test_gpu_projection.py -m gpu -q --timeout 300 2>&1 | grep -E "^E  |passed|failed|Error" | head -n 12 | cut -c1-250
timeout 200 python scripts/bench_pair_configs.py 2>&1 | tail -n 10
