#!/bin/bash
# tools/sweep_selfplay.sh "<selfplay_bench args>" ...
run() { python tools/selfplay_bench.py "$@" 2>gpurun_out/sp.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('$*', '| games/h %.2fM sims/s %.2fM sec %.2f waves %d hit %.3f term %.3f maxnodes %s' % (d['games_per_hour']/1e6, d['sims_per_sec']/1e6, d['seconds'], d['waves'], d['cache_hit_frac'], d['terminal_leaf_frac'], d['max_nodes_used']))
" || tail -3 gpurun_out/sp.err; }
for a in "$@"; do run $a; done
