#!/bin/bash
# scratch: K1 variant sweep (built on the GPU box, one probe per variant)
mkdir -p gpurun_out
: > gpurun_out/k1_sweep2.jsonl
run() {
  R3D_NVCC_EXTRA="$1" python 3d_reconstruction_system_b200/build.py --force > gpurun_out/build_k1.log 2>&1 || { echo "build failed: $1"; tail -5 gpurun_out/build_k1.log; return; }
  R3D_NVCC_EXTRA="$1" timeout 300 python tools/k1_probe.py 4500 10 2>> gpurun_out/k1_sweep.err | tee -a gpurun_out/k1_sweep2.jsonl
}
run "-DK1V_GROUPS=2 -DK1V_STAGES=2 -DK1V_MINB=3"
run "-DK1V_GROUPS=2 -DK1V_STAGES=1 -DK1V_MINB=3"
run "-DK1V_GROUPS=2 -DK1V_STAGES=2 -DK1V_MINB=3 -DK1V_INTERLEAVE=1"
run "-DK1V_GROUPS=2 -DK1V_STAGES=2 -DK1V_MINB=2 -DK1V_OUTBUFS=3"
run "-DK1V_GROUPS=3 -DK1V_STAGES=2 -DK1V_MINB=2"
run "-DK1V_GROUPS=1 -DK1V_STAGES=2 -DK1V_MINB=3"
run "-DK1V_GROUPS=1 -DK1V_STAGES=2 -DK1V_MINB=4"
run "-DK1V_GROUPS=2 -DK1V_STAGES=2 -DK1V_MINB=3"
