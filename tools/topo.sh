nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
nvidia-smi -L >> gpurun_out/topo.txt 2>&1
lscpu >> gpurun_out/topo.txt 2>&1
echo "--- numa" >> gpurun_out/topo.txt
ls /sys/devices/system/node/ >> gpurun_out/topo.txt 2>&1
for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) >> gpurun_out/topo.txt; grep MemTotal $n/meminfo >> gpurun_out/topo.txt; done
echo "--- pci" >> gpurun_out/topo.txt
nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv >> gpurun_out/topo.txt 2>&1
for d in /sys/bus/pci/devices/*; do v=$(cat $d/vendor 2>/dev/null); if [ "$v" = "0x10de" ]; then echo $d numa=$(cat $d/numa_node) local_cpus=$(cat $d/local_cpulist) >> gpurun_out/topo.txt; fi; done
echo "--- affinity" >> gpurun_out/topo.txt
python -c "import os;print(len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8])" >> gpurun_out/topo.txt 2>&1
free -g >> gpurun_out/topo.txt 2>&1
which numactl >> gpurun_out/topo.txt 2>&1
cat /proc/cpuinfo | grep "model name" | head -1 >> gpurun_out/topo.txt
systemd-detect-virt >> gpurun_out/topo.txt 2>&1
