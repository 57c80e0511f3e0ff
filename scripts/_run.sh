cd $GRAFT_REPO_ROOT
EXPB200_DEPOSIT_VARIANT=4 timeout 300 python -m pytest tests/test_gpu_sphere.py tests/test_gpu_ref_parity.py tests/test_gpu_multistep.py tests/test_gpu_covariance.py -x -q 2>&1 | tail -5 > gpurun_out/r02l_gputest.txt
for v in 1 4; do EXPB200_DEPOSIT_VARIANT=$v timeout 200 python scripts/kbench.py 1e7 10 > gpurun_out/r02l_kbench_v$v.txt 2>&1; done
