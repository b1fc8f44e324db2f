#!/bin/bash
# last check of the round: aggregation timings, then the full sequence of tools/gpu_check_all.sh
timeout 200 python tools/spmm_time.py
tools/gpu_check_all.sh
