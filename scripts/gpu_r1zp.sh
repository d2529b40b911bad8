#!/bin/bash
# Round-1 session zr (N GPUs): partitioned Newton solve, Krylov iterations replayed from a CUDA graph cached across linear solves.
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -x -q > $OUT/pytest_solver_r1zr.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_solver_r1zr.log
timeout 600 $TR scripts/check_partitioned_newton.py > $OUT/check_partitioned_n${N}_r1zr.log 2>&1; echo "check rc=$?"; grep -v "^\*\|OMP_NUM\|^$" $OUT/check_partitioned_n${N}_r1zr.log | tail -6 | cut -c1-330
for extra in "" "--no-graph"; do
timeout 900 $TR scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --partition $extra > $OUT/newton55_part_n${N}_r1zr$extra.log 2>&1; echo "newton partition $extra rc=$?"; tail -1 $OUT/newton55_part_n${N}_r1zr$extra.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','solve_s','linear_solve_s','ms_per_krylov_iteration','cuda_graph_replays','cuda_graph_captures','cuda_graph_error','mean_sigma_xx')})"
timeout 900 python scripts/bench_newton.py --grid 55 --steps 2 --forcing ew $extra > $OUT/newton55_n1_r1zr$extra.log 2>&1; echo "newton n1 $extra rc=$?"; tail -1 $OUT/newton55_n1_r1zr$extra.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','solve_s','linear_solve_s','ms_per_krylov_iteration','cuda_graph_replays','cuda_graph_captures','cuda_graph_error','mean_sigma_xx')})"
done
