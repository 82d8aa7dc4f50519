timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
python -c "
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('value',d['value']/1e9,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e6, d['e2e'].get('host_d2h_ceiling_gbs'), d['e2e'].get('frac_of_host_ceiling'))
print('roofline',d['roofline']['frac'], 'match_ms',d['extra']['match_ms'],'compose',d['extra']['compose_ms'],'idx build',d['extra']['index_build_ms'], 'per render', d['extra']['value_per_render']/1e9)
print('scan', d['extra']['match_scan']['frac'], d['extra']['match_scan']['ms_per_launch'])
c2=d['extra']['c2_4to1']; print('c2', c2.get('value'), c2.get('match_ms'), c2.get('roofline',{}).get('frac'), c2.get('e2e',{}).get('value'))
print({k:(v.get('roofline',{}).get('frac') if isinstance(v,dict) else v) for k,v in d['extra'].items() if k.startswith('c3') or k.startswith('c5')})
print(d['extra'].get('resize_lanczos3'))
print('cpu',d['cpu_baseline'])
"
