#!/bin/bash
# sweep of the wave-loop options on the bench workload; one compact line per run:  tools/sweep_cache.sh "<bench args>" ...
run() {
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ablation "$@" 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$*', '| sims/s %.2fM e2e %.2fM ms/step %.1f | kernel_us %.1f sims/launch %.0f launches/step %.0f hit %.3f frac %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], r['kernel_us'], r['sims_per_launch'], r['launches_per_step'], r['cache_hit_frac'], r['frac']))
" || tail -3 gpurun_out/sweep.err
}
for a in "$@"; do run $a; done
