#!/bin/bash
# Round-1 sessions zp-zr (N GPUs): partitioned Newton solve: check against the single-GPU solve, then 1 M cells over N
# ranks next to the single-GPU solve.  (Sessions zq / zr also ran a CUDA-graph mode of the Krylov loop that has since been
# removed, see DESIGN.md section 8 and profiles/r1zq_*, r1zr_newton55_n1.log.)
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -x -q > $OUT/pytest_solver_r1zr.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_solver_r1zr.log
timeout 600 $TR scripts/check_partitioned_newton.py > $OUT/check_partitioned_n${N}_r1zr.log 2>&1; echo "check rc=$?"; grep -v "^\*\|OMP_NUM\|^$" $OUT/check_partitioned_n${N}_r1zr.log | tail -6 | cut -c1-330
timeout 600 $TR scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --partition > $OUT/newton55_part_n${N}_r1zr.log 2>&1; echo "newton partition rc=$?"; tail -1 $OUT/newton55_part_n${N}_r1zr.log | cut -c1-1200
timeout 600 python scripts/bench_newton.py --grid 55 --steps 2 --forcing ew > $OUT/newton55_n1_r1zr.log 2>&1; echo "newton n1 rc=$?"; tail -1 $OUT/newton55_n1_r1zr.log | cut -c1-1200
