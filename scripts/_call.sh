set -x
A="python bench.py --graphs 0 --steps 3 --warmup 3 --no-also --no-cpu-baseline --e2e-steps 8"
B="python bench.py --graphs 0 --steps 3 --warmup 3 --no-also --no-cpu-baseline --e2e-steps 8 --batch 64 --size 64"
$A > gpurun_out/r02_plainA.json 2> gpurun_out/r02_plainA.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_configA_eager.csv $A > gpurun_out/r02_ncuA.log 2>&1
$B > gpurun_out/r02_plainB.json 2> gpurun_out/r02_plainB.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_configB_eager.csv $B > gpurun_out/r02_ncuB.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'corr_tc|prep_kernel|paste|shift_bwd|blend_s|recheck' -s 60 -c 10 -o gpurun_out/r02_full_B $B > gpurun_out/r02_ncu_fullB.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'corr_tc|prep_kernel|paste|shift_bwd|blend_s|recheck|build_exc' -s 60 -c 9 -o gpurun_out/r02_full_A $A > gpurun_out/r02_ncu_fullA.log 2>&1
ls -la gpurun_out/r02_*
