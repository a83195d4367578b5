#!/bin/bash
# round-2 validation run 6: the new GPU tests (combine / zoom / profiles / hill / top / metrics / field / gamma / WL synthetic / PF DICOM)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r6_build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_combine.py tests/test_gpu_profiles_ext.py tests/test_gpu_profile.py tests/test_gpu_field.py -x -q -m gpu > gpurun_out/r6_new.log 2>&1
echo "new tests exit $?" >> gpurun_out/r6_new.log
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r6_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/r6_all.log
tail -5 gpurun_out/r6_new.log; tail -15 gpurun_out/r6_all.log
