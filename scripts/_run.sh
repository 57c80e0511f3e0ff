cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_ref_parity.py -x -q 2>&1 | tail -12 > gpurun_out/r02p_gputest.txt
