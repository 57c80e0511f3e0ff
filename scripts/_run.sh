cd $GRAFT_REPO_ROOT
for r in 2 3 4 5; do EXPB200_WS_REGS=$r timeout 200 python scripts/kbench.py 1e7 10 > gpurun_out/r02f_kbench_regs$r.txt 2>&1; done
EXPB200_FORCE_VARIANT=2 timeout 200 python scripts/kbench.py 1e7 10 > gpurun_out/r02f_kbench_v2.txt 2>&1
