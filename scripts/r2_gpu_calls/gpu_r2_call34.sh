#!/bin/bash
# Round 2, GPU call 34 (1 GPU): checkpoint of the round's final code (scripts/gpu_checkpoint.sh r2j).
set -u
bash scripts/gpu_checkpoint.sh r2j
