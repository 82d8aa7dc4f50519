nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1; bash tools/topo.sh > /dev/null 2>&1; cp gpurun_out/topo.txt gpurun_out/topo_n8_full.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -6
timeout 300 python tools/bench_hostcopy.py > gpurun_out/hostcopy_n8.txt 2>&1; tail -32 gpurun_out/hostcopy_n8.txt
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo bench N=$N rc=$?; tail -3 gpurun_out/bench_n$N.err
python -c "
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('N',d['n_gpus'],'value',d['value']/1e9,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e'].get('host_d2h_ceiling_gbs'), d['e2e'].get('frac_of_host_ceiling'), d['e2e']['host_ceiling']['d2h_one_rank_alone_gbs'])
print('match_ms',d['extra']['match_ms'],'compose',d['extra']['compose_ms'])
c2=d['extra']['c2_4to1']; print('c2', c2.get('value'), c2.get('match_ms'), c2.get('roofline',{}).get('frac'), c2.get('e2e',{}).get('value'))
print('group', d['extra'].get('e2e_single_process_group'))
print('c3', d['extra'].get('c3_analysis_sharded'))
print('weak', d['extra'].get('weak_scaling'))
"
done
