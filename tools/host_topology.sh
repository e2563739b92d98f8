nvidia-smi topo -m 2>&1 | head -14
lscpu | grep -i "numa\|socket\|^CPU(s)"
nproc
for d in /sys/bus/pci/devices/*; do c=$(cat $d/class); if [ "$c" = "0x030200" ]; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done
free -g | head -2
