#!/bin/bash
# statistics-first scatter pass: tests, then the overlapping mixtures at the headline shape
python -m pytest tests/test_emgmm_gpu.py -m gpu -x -q -k "${KEXPR:-statistics_first or band_and_side or quadratic or config3}" > gpurun_out/pytest_gpu_gs.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_gs.log
tail -30 gpurun_out/pytest_gpu_gs.log
[ -n "$SKIP_OV" ] || SCS="${SCS:-1 0.5}" bash scripts/gpu_overlap.sh
