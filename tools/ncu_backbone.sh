set -x
# the default stem configuration (928 threads x 64 registers) cannot run ncu's instrumented passes (LaunchFailed kills the target):
# the block / chain kernels get --set full, the stem the hardware-counter sections only
python tools/profile_target.py 96 4096 > gpurun_out/profile_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"blaze_block|blaze_chain" -s 8 -c 8 -o gpurun_out/prof_backbone_r02 python tools/profile_target.py 96 4096 > gpurun_out/ncu_full.log 2>&1
echo ncu exit $?
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --clock-control none -k regex:"stem_" -s 1 -c 1 -o gpurun_out/prof_stem_r02 python tools/profile_target.py 96 4096 > gpurun_out/ncu_stem.log 2>&1
echo ncu stem exit $?
ls -la gpurun_out/*.ncu-rep
