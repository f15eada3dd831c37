#!/bin/bash
# k-means initialiser: tests, then its time at the headline shape with the first- and second-generation kernels
python -m pytest tests/test_emgmm_gpu.py -m gpu -x -q -k "kmeans or fresh_object" > gpurun_out/pytest_gpu_km.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_km.log
tail -25 gpurun_out/pytest_gpu_km.log
for v in 1 0; do
  echo "MLF_B200_KMEANS_V1=$v"; MLF_B200_DEBUG=1 MLF_B200_KMEANS_V1=$v python scripts/time_kmeans_init.py ${N:-1e8} 2>&1 | grep -E "k-means|init_seconds" | tee -a gpurun_out/kmeans_init.txt
done
