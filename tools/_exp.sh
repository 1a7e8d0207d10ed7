#!/bin/bash
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag
lscpu | grep -i "numa\|model name\|^CPU(s)"
nvidia-smi topo -m 2>&1 | head -6
for hp in 0 1 0 1; do
  R3D_HOST_HUGEPAGES=$hp timeout 300 python tools/e2e_probe.py 1500 2>&1 | tail -1
  grep -i "AnonHugePages" /proc/meminfo
done
