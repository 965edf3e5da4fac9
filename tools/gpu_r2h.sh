#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batch_fold.py -m gpu -q 2>&1 | grep -E "^E |passed|failed" | head -30 | cut -c1-300
