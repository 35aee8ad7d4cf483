# launch list + one full capture of each hot kernel of a bench step (run under gpurun, one GPU); TAG names the outputs,
# WORKLOAD (default cfg2) the bench workload, KREGEX the kernels of the full capture
mkdir -p gpurun_out
T=${TAG:-r02}
CMD="python bench.py --workload ${WORKLOAD:-cfg2} --steps 2 --warmup 3 --no-cpu-baseline --no-cfg3"
K=${KREGEX:-hupd_ts_kernel|recon_ts_kernel|gradw_ts_kernel}
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launch.log 2>&1
$CMD > gpurun_out/${T}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 8 -c 4 -o gpurun_out/prof_${T} -f $CMD > gpurun_out/${T}_ncu_full.log 2>&1
tail -3 gpurun_out/${T}_ncu_full.log
ls -la gpurun_out | grep ${T}
