#!/bin/bash
# kernel-tuning helper (GPU box): bench.py device numbers for the default library and for env/library variants given as "name|ENV=.. ENV=.."
for spec in "$@"; do
  name=${spec%%|*}; envs=${spec#*|}
  env $envs python bench.py --no-cpu --steps 200 --e2e-steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
p = d['roofline']['phase_ms_per_step']
print('$name', 'value %.4g ms/step %.4f e2e %.4g | integ %.4f rebuild %.4f contact %.4f reduce %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], p['ms_integrate'], p['ms_rebuild'], p['ms_contact'], p['ms_reduce']))
"
done
