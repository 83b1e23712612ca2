set -x
python tools/profile_target.py 96 4096 > gpurun_out/profile_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"stem_|blaze_block|blaze_chain" -s 9 -c 9 -o gpurun_out/prof_backbone_r02 python tools/profile_target.py 96 4096 > gpurun_out/ncu_full.log 2>&1
echo ncu exit $?
ls -la gpurun_out/*.ncu-rep
