while read -r v; do
  env $v timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps ${ES:-20} > gpurun_out/v.json 2> gpurun_out/v.err
  python -c "
import json; d=json.load(open('gpurun_out/v.json')); print('$v', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value'],2))"
  grep "zsb pipe" gpurun_out/v.err | tail -${TR:-0} | cut -c1-140
done
