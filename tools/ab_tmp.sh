set -u
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_full_width.py -m gpu -q -x -p no:cacheprovider > $O/r3a_tests.log 2>&1
tail -2 $O/r3a_tests.log
SM="--workload small64 --no-full128 --steps 30 --warmup 5 --skip-cpu-baseline --sampler-steps 0"
for i in 1 2; do
timeout 200 python bench.py $SM > $O/r3a_small_new$i.json 2>/dev/null
XUNET_LIB=novel_view_synthesis_3d_b200/libxunet_b200_prev3.so timeout 200 python bench.py $SM > $O/r3a_small_prev$i.json 2>/dev/null
done
for f in $O/r3a_small_*.json; do echo $f $(grep -h -o '"ms_per_step": [0-9.]*' $f | head -1); done
XUNET_NO_PDL=1 timeout 200 python tools/kineto_step.py 2>/dev/null | sed -n 2,4p
