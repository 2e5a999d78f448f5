#!/bin/bash
# SASS opcode histogram of the hot kernels (runs without a GPU): cuobjdump -sass on the built object
# usage: tools/sass_histogram.sh > profiles/rNN_sass_histogram.txt
O=photons-2.0_b200/csrc/pn2_walk.o
for k in '_Z17walk_fused_kernelILi8ELb1EEv8WalkArgs8P2PConst' '_Z21walk_fused_f64_kernelILi8ELb1EEv8WalkArgs8P2PConst' '_Z20frontier_node_kernel8WalkArgs8P2PConst'; do
  echo "## $(echo $k | c++filt)   ($(git rev-parse --short HEAD), nvcc $(nvcc --version | grep -o 'V[0-9][0-9.]*' | head -1))"
  cuobjdump -sass -fun "$k" $O | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sort | uniq -c | sort -rn | head -40
  echo
done
echo "## TMA / tcgen05 mnemonics (UBLKCP, UTMALDG, UTCMMA): $(cuobjdump -sass $O | grep -cE 'UBLKCP|UTMALDG|UTCMMA|UTCHMMA') -- the path is CUDA-core work staged with LDGSTS (cp.async)"
