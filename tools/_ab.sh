for cfg in "TE_WG_HALO=0 TE_WGRAD_WS_CACHE=0" ""; do
  env $cfg timeout 600 python bench.py --steps 16 --warmup 3 --fp32-steps 0 --no-cpu-baseline > gpurun_out/ab.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); print('$cfg', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])"
done
