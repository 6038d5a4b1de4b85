nvidia-smi topo -m 2>&1 | head -30
nproc; grep -i "allowed_list" /proc/self/status
ls /sys/devices/system/node/ | head; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ] && [ "$(cat $d/class)" = "0x030200" ]; then echo $d numa $(cat $d/numa_node); fi; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${1:-8} --master-addr 127.0.0.1 --master-port 29512 tools/probes/pcie_probe.py 2>&1 | grep "^rank" | sort
