export POLARIS_SCENE_CACHE=/tmp/polaris_scenes
( POLARIS_CUDA_LIB=$PWD/ab_sortleaf.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_depth or golden or slots or deterministic or bounce_count or chains" 2>&1 ) | tail -60 > gpurun_out/sortleaf_fail.txt
bash tools/run_gpu_tests.sh r02m
