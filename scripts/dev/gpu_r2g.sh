#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "stress_rows or extremal or cfg5_shard or chunked or batch_cases" 2>&1 | tail -15
python scripts/e2e_probe.py 2>&1 | tail -10
