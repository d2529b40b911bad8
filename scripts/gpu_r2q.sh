#!/bin/bash
# Round-2 session q (1 GPU): whole GPU test suite + bench (all blocks) on the state with the faster Drucker-Prager
# kernels, the CTA-parallel Krylov reduction tails and the two-pairs-per-trip uniaxial kernel.
TAG=${1:-r2q}
OUT=gpurun_out; mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
echo "== pytest all"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
echo "== models"; timeout 900 python scripts/bench_models.py --out $OUT/models_$TAG.json > $OUT/models_$TAG.log 2>&1; echo "models rc=$?"; grep -o '"kernel": "[^"]*"\|"ms": [0-9.e-]*\|"frac_of_measured_hbm": [0-9.e-]*' $OUT/models_$TAG.log | paste - - - | cut -c1-150
