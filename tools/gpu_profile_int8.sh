#!/bin/bash
# ncu per-launch metrics of one ResNet-152 INT8 batch-32 forward pass (tactics pinned by a cache file so that ncu does not
# perturb the load-time tuner).  Run under gpurun from the repo root; outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r2_int8}
export PROFILE_DEPTH=152 PROFILE_BATCH=32 PROFILE_PREC=int8 PROFILE_TUNE_STREAMS=8
export B2_TUNE_CACHE=$PWD/gpurun_out/tactic_cache_$TAG.txt
rm -f "$B2_TUNE_CACHE"
python tools/profile_forward.py 1 > /dev/null 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__waves_per_multiprocessor
timeout 900 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/ncu_raw_$TAG.csv python tools/profile_forward.py 1 > gpurun_out/ncu_$TAG.log 2>&1
python tools/condense_ncu.py gpurun_out/ncu_raw_$TAG.csv gpurun_out/launch_names.txt > gpurun_out/ncu_metrics_$TAG.csv
cp gpurun_out/launch_names.txt gpurun_out/launch_names_$TAG.txt
rm -f gpurun_out/ncu_raw_$TAG.csv
wc -l gpurun_out/ncu_metrics_$TAG.csv "$B2_TUNE_CACHE"
