"""Render a batch-sweep jsonl (tools/batch_sweep.sh) as a markdown table."""
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith('{')]
print('| workload | GPUs | per-GPU batch | ms/step | images/s | step TFLOP/s per GPU | of sustained tensor peak | kernels/step |')
print('|---|---|---|---|---|---|---|---|')
for d in rows:
    c = d['config']
    k = d['kernels_per_step']
    print(f"| {c['workload']} | {d['n_gpus']} | {c['per_gpu_batch']} | {d['ms_per_step']:.3f} | {d['value']:.1f} | {d['step_tflops']:.1f} | "
          f"{d['step_tensor_frac_of_sustained']:.3f} | {k['forward']}+{k['backward']}+1 |")
