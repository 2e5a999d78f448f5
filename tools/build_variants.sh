#!/bin/bash
# Kernel-parameter variants of libpn2gpu.so, cross-compiled here so that a sweep on the GPU box costs no compile time:
#   tools/build_variants.sh "name1:-DFLAG=..;name2:-DFLAG=.. -DFLAG2=.."   ->  photons-2.0_b200/variants/libpn2gpu_<name>.so
# (pn2_walk.cu and pn2_operators.cu are recompiled; select one with PN2GPU_LIB=<path>)
cd "$(dirname "$0")/../photons-2.0_b200/csrc" || exit 1
mkdir -p ../variants
IFS=';' read -ra V <<< "$1"
for v in "${V[@]}"; do
  name=${v%%:*}; flags=${v#*:}
  /usr/local/cuda/bin/nvcc $flags -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -fmad=false -c -o ../variants/pn2_walk_$name.o pn2_walk.cu || exit 1
  /usr/local/cuda/bin/nvcc $flags -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -c -o ../variants/pn2_operators_$name.o pn2_operators.cu || exit 1
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libpn2gpu_$name.so pn2_api.o ../variants/pn2_operators_$name.o pn2_p2p.o pn2_modeb.o pn2_tree.o ../variants/pn2_walk_$name.o pn2_let.o pn2_migrate.o pn2_integrate.o pn2_pm.o -ldl || exit 1
  rm -f ../variants/pn2_walk_$name.o ../variants/pn2_operators_$name.o
  echo "built variants/libpn2gpu_$name.so ($flags)"
done
