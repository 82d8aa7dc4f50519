timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo bench N=$N rc=$?; tail -1 gpurun_out/bench_n$N.err
python -c "
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('N',d['n_gpus'],'value',d['value']/1e9,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e'].get('host_d2h_ceiling_gbs'), d['e2e'].get('frac_of_host_ceiling'), d['e2e']['host_ceiling']['d2h_one_rank_alone_gbs'])
print('match_ms',d['extra']['match_ms'],'compose',d['extra']['compose_ms'], 'photo', d['extra'].get('c4_photo_like_source',{}).get('value'))
c2=d['extra']['c2_4to1']; print('c2', c2.get('value'), c2.get('match_ms'), c2.get('roofline',{}).get('frac'), c2.get('e2e',{}).get('value'))
print('group', d['extra'].get('e2e_single_process_group'))
print('c3', d['extra'].get('c3_analysis_sharded',{}).get('ms'), 'weak', d['extra'].get('weak_scaling',{}).get('value'))
"
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench N=1 rc=$?
python -c "
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('N',d['n_gpus'],'value',d['value']/1e9,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e6, d['e2e'].get('frac_of_host_ceiling'), 'roof', d['roofline']['frac'], d['roofline']['traffic'])
print('c2', d['extra']['c2_4to1'].get('value'), d['extra']['c2_4to1'].get('roofline',{}).get('frac'), 'per render', d['extra']['value_per_render']/1e9, 'photo', d['extra'].get('c4_photo_like_source'))
"
