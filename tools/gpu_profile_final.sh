#!/bin/bash
# Final-of-round evidence: tactic cache, ncu per-launch metrics of one forward pass (tactics pinned by the cache),
# phase timing.  Run under gpurun from the repo root; outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1h}
export B2_TUNE_CACHE=$PWD/gpurun_out/tactic_cache_$TAG.txt
rm -f "$B2_TUNE_CACHE"
python tools/profile_forward.py 1 > /dev/null 2>&1            # tunes on 4 streams, writes the cache
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__waves_per_multiprocessor
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/ncu_raw_$TAG.csv python tools/profile_forward.py 1 > gpurun_out/ncu_$TAG.log 2>&1
python tools/condense_ncu.py gpurun_out/ncu_raw_$TAG.csv gpurun_out/launch_names.txt > gpurun_out/ncu_metrics_$TAG.csv
cp gpurun_out/launch_names.txt gpurun_out/launch_names_$TAG.txt
rm -f gpurun_out/ncu_raw_$TAG.csv
python tools/gpu_phase_timing.py > gpurun_out/phase_timing_$TAG.txt 2>&1
wc -l gpurun_out/ncu_metrics_$TAG.csv gpurun_out/phase_timing_$TAG.txt "$B2_TUNE_CACHE"
