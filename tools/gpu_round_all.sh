#!/bin/bash
bash tools/gpu_round.sh ${1:-r2a}
bash tools/gpu_round_b.sh ${2:-r2b}
