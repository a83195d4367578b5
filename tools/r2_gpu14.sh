#!/bin/bash
# round-2 run 14: where did k_pf_tail's 22 us come from?  stage times with the current library, without the prominence block table in the PF
# translation units, and with the r2k peaks.cuh in them; ncu of k_pf_tail for the first two
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
for lib in "" "$PWD/variants/libepid_pfnotab.so" "$PWD/variants/libepid_pfoldpeaks.so"; do
  EPID_LIB=$lib timeout 200 python tools/r2_stages.py --iters 10 2>&1 | tail -1 | sed "s#^#lib=${lib##*/} #" | tee -a $O/r14_stages.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k k_pf_tail -s 1 -c 1 -o $O/r14_tail_new -f python tools/prof_pf.py 1 512 > /dev/null 2>&1
EPID_LIB=$PWD/variants/libepid_pfoldpeaks.so timeout 300 ncu --set full --clock-control none --import-source on -k k_pf_tail -s 1 -c 1 -o $O/r14_tail_old -f python tools/prof_pf.py 1 512 > /dev/null 2>&1
