set -x
python tools/profile_step.py 64 1 > gpurun_out/r02_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py 64 1 > gpurun_out/r02_ncu_a.log 2>&1
VGQA_CHAIN=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_chain.csv python tools/profile_step.py 64 1 > gpurun_out/r02_ncu_b.log 2>&1
VGQA_CHAIN=1 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 12 -c 2 -f -o gpurun_out/r02_chain python tools/profile_step.py 64 1 > gpurun_out/r02_ncu_c.log 2>&1
ncu --metrics sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum --clock-control none -s 468 -c 480 --csv --log-file gpurun_out/r02_dec_tensor.csv python tools/profile_step.py 64 1 > gpurun_out/r02_ncu_d.log 2>&1
ls -la gpurun_out/r02_*
