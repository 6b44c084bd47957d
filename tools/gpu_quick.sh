mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2k/bench.json 2> gpurun_out/r2k/bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2k/bench.json')); print('cold %.3f warm %.3f keyed %.3f nocache %.3f e2e %.1fM'%(d['ms_per_step'], d['warm_keycache']['ms_per_step'], d['keyed']['ms_per_step'], d['no_keycache']['ms_per_step'], d['e2e']['value']/1e6)); print(d['roofline']['traffic'], d['roofline']['traffic_source'])"
