set -x
timeout 300 python tools/noise_cluster_ab.py > gpurun_out/noise_cluster_ab.log 2>&1
tail -20 gpurun_out/noise_cluster_ab.log
timeout 600 python -m pytest tests -m gpu -x -q -k "noise or sampling" 2>&1 | tail -5
