#!/bin/bash
# What the GPU box looks like from inside the container: CPU / NUMA / PCIe topology and memory policy.
# Output goes to gpurun_out/box_topology.txt (read by hand; feeds cones_perception_b200/placement.py).
out=${1:-gpurun_out/box_topology.txt}
mkdir -p "$(dirname "$out")"
{
  echo "== nproc / affinity"; nproc; grep -i -E 'cpus_allowed_list|mems_allowed_list' /proc/self/status
  echo "== lscpu"; lscpu | head -40
  echo "== numa nodes"; ls /sys/devices/system/node/ 2>&1
  for n in /sys/devices/system/node/node*; do echo "$n cpulist: $(cat $n/cpulist 2>&1)"; grep -E 'MemTotal|MemFree' $n/meminfo 2>&1; done
  echo "== nvidia-smi topo"; nvidia-smi topo -m 2>&1
  echo "== nvidia-smi -L"; nvidia-smi -L 2>&1
  echo "== pci numa_node of nvidia devices"
  for d in /sys/bus/pci/devices/*; do
    if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa_node=$(cat $d/numa_node 2>&1) local_cpulist=$(cat $d/local_cpulist 2>&1) class=$(cat $d/class 2>&1)"; fi
  done
  echo "== pcie link"; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv 2>&1
  echo "== free"; free -g
  echo "== hugepages"; grep -i huge /proc/meminfo
  echo "== cgroup cpuset"; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>&1
  echo "== numactl"; which numactl 2>&1; numactl -H 2>&1 | head -20
  echo "== virtualization"; systemd-detect-virt 2>&1; grep -m1 -i hypervisor /proc/cpuinfo
} > "$out" 2>&1
