#!/bin/bash
# Round-2 final evidence run (1 GPU): GPU suite, sanitizer, the default bench line (+ reference arm), ncu launch list of the bench
# command, ncu DRAM-byte metrics of the memory-bound kernels in a full-128^2 step, CUPTI breakdowns.   Outputs -> gpurun_out/r2t_*
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2t_suite.log 2>&1
tail -3 $O/r2t_suite.log
SANITIZE_TIMEOUT=150 timeout 700 bash tools/sanitize.sh > $O/r2t_sanitize_summary.txt 2>&1
cat $O/r2t_sanitize_summary.txt
for t in memcheck racecheck synccheck; do gzip -f $O/sanitizer_$t.log; done
timeout 900 python bench.py > $O/r2t_bench.json 2> $O/r2t_bench.err
tail -c 600 $O/r2t_bench.json; echo
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2t_bench_reference.json 2> $O/r2t_bench_reference.err
tail -c 400 $O/r2t_bench_reference.json; echo
XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/kineto_step.py > $O/r2t_kineto_full.txt 2>&1
XUNET_NO_PDL=1 timeout 200 python tools/kineto_step.py > $O/r2t_kineto_small.txt 2>&1
XUNET_NO_PDL=1 XU_MODEL=full XU_B=4 XU_S=128 timeout 300 python tools/conv_step_profile.py > $O/r2t_conv_step_profile.txt 2>&1
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2t_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-full128 --skip-cpu-baseline --sampler-steps 0 --no-graph > $O/r2t_launches.log 2>&1
gzip -f $O/r2t_launches.csv
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,launch__registers_per_thread,launch__grid_size,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed"
XU_MODEL=full XU_B=4 XU_S=128 timeout 600 ncu --metrics $M --clock-control none \
  -k regex:'gn_|adam|weight_prep|copy_channels|resample|scale_add|emb_|logsnr|thin|cout3|loss|pose|sampler|diffusion' --csv --log-file $O/r2t_membound.csv python tools/run_step.py 1 > $O/r2t_membound.log 2>&1
gzip -f $O/r2t_membound.csv
ls -la $O/r2t_*
du -sh $O
