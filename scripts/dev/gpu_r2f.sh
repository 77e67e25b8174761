#!/usr/bin/env bash
set -u
bash scripts/gpu_quick.sh r2f_occ32
TB_SCAN_FUSED_OCC=28 bash scripts/gpu_quick.sh r2f_occ28
