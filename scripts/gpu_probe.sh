#!/bin/bash
# Toolchain / topology probe of the GPU box: is there a Fortran compiler, libhdf5 or h5py to pin
# the oracle to the real reference with?  How are the GPUs attached to the host (PCIe switches, NUMA)?
O=gpurun_out/toolchain_probe.txt
{
echo "== date"; date -u
echo "== compilers / hdf5 tools (command -v; empty = absent)"
for t in gfortran flang flang-new nvfortran ifort ifx lfortran f2c f77 f95 h5dump h5cc h5fc h5ls mpif90 numactl lspci; do
  printf '%-12s %s\n' "$t" "$(command -v $t || echo -)"; done
echo "== python modules"
for m in h5py tables netCDF4 scipy numpy; do python -c "import $m; print('$m', $m.__version__)" 2>&1 | tail -1; done
echo "== ldconfig: hdf5 / lapack / gfortran / blas"
ldconfig -p | grep -iE "hdf5|lapack|gfortran|openblas|libblas" || echo "(none)"
echo "== gcc language front ends"
ls "$(dirname "$(gcc -print-prog-name=cc1)")" 2>/dev/null
echo "== stray files"
find / \( -iname "libhdf5*" -o -iname "f951" -o -iname "*gfortran*" -o -iname "*.mod" \) -not -path "/proc/*" 2>/dev/null | head -20
echo "== stock-libhdf5-written files on the box"
find / \( -name "*.h5" -o -name "*.hdf5" -o -name "testhdf5*" \) -not -path "/proc/*" 2>/dev/null | head
echo "== host"
nproc; lscpu | grep -iE "model name|socket|numa|thread|core" ; free -g | head -2
echo "== nvidia-smi topo"
nvidia-smi topo -m 2>&1
echo "== nvidia-smi -q (pcie)"
nvidia-smi -q 2>/dev/null | grep -iE "Product Name|Bus Id|Link Width|PCIe Generation|Max|Current" | head -40
echo "== pci tree"
(lspci -tv 2>/dev/null || ls /sys/bus/pci/devices | head -50) | head -80
echo "== numa nodes"
ls /sys/devices/system/node 2>/dev/null; cat /sys/devices/system/node/node*/cpulist 2>/dev/null
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ]; then echo "$d numa=$(cat $d/numa_node) local_cpus=$(cat $d/local_cpulist) speed=$(cat $d/current_link_speed 2>/dev/null) width=$(cat $d/current_link_width 2>/dev/null)"; fi; done
} > $O 2>&1
cat $O | head -120
