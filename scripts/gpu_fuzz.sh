#!/bin/bash
# random-shape stress of the CUDA path against the oracle with the round's new code paths on (defaults) and forced
timeout 400 python scripts/fuzz_parity.py ${NP:-300} ${SEEDP:-11} > gpurun_out/fuzz_parity.log 2>&1; echo "fuzz_parity exit $?"; tail -4 gpurun_out/fuzz_parity.log
MLF_B200_QF=2 timeout 300 python scripts/fuzz_parity.py ${NQ:-200} ${SEEDQ:-12} > gpurun_out/fuzz_parity_qf2.log 2>&1; echo "fuzz_parity (quadratic form forced) exit $?"; tail -4 gpurun_out/fuzz_parity_qf2.log
timeout 400 python scripts/fuzz_api.py ${NA:-150} ${SEEDA:-13} > gpurun_out/fuzz_api.log 2>&1; echo "fuzz_api exit $?"; tail -4 gpurun_out/fuzz_api.log
