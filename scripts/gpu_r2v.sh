#!/bin/bash
# Round-2 session v (1 GPU): Eisenstat-Walker forcing bounds of the stand-in Newton driver (config 5, one GPU).
TAG=${1:-r2v}
OUT=gpurun_out; mkdir -p $OUT
for cfg in "1e-2 0.9" "3e-2 0.9" "1e-1 0.9" "3e-1 0.9" "1e-1 0.5" "5e-1 0.9"; do
  set -- $cfg
  timeout 200 python scripts/bench_newton.py --grid 55 --steps 2 --forcing ew --driver device --eta-max $1 --eta-gamma $2 > $OUT/newton55_eta_$1_$2_$TAG.log 2>&1; echo "rc=$? eta_max=$1 gamma=$2"
  tail -1 $OUT/newton55_eta_$1_$2_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('solve_s','linear_solve_s','residual_s','newton_iterations','mean_sigma_xx','plastic_fraction_final')}, [sum(k) for k in d['krylov_iterations']])"
done
