#!/bin/bash
# bisect of the NCCL merge parity worker failure
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # label env...
  label=$1; shift
  env "$@" timeout 300 $TR --master-port 29511 tests/workers/nccl_octomap_worker.py 19 3 > gpurun_out/mg2c_$label.log 2>&1; echo "$label exit $?"; grep -h "AssertionError\|ok:" gpurun_out/mg2c_$label.log | head -3
}
run default A=1
run batch4 R3D_SCAN_BATCH=4
run chunk256 R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200_c256.so
run b4c256 R3D_SCAN_BATCH=4 R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200_c256.so
run malloc R3D_POOL_MALLOC=1
run nooverlap R3D_PIPE_OVERLAP=0
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -3
