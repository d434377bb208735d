set -x
python bench.py --steps 2 --warmup 1 --min-warmup 1 --no-cpu-baseline --only c2,lloyd > gpurun_out/plain.json 2>/dev/null; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --min-warmup 1 --no-cpu-baseline --only c2,lloyd > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 4 -c 1 -o gpurun_out/r2_topp -f python tools/p2_probe.py > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 5 -c 1 -o gpurun_out/r2_collect -f python tools/p2_probe.py > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:p2_exact -s 3 -c 1 -o gpurun_out/r2_exact -f python tools/p2_probe.py > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:p2u_select -s 1 -c 1 -o gpurun_out/r2_p2u -f python tools/p2_probe.py > gpurun_out/ncu_d.log 2>&1; tail -1 gpurun_out/ncu_d.log
for f in r2_topp r2_collect r2_exact r2_p2u; do python tools/ncu_summary.py gpurun_out/$f.ncu-rep gpurun_out/${f}_ncu_full_summary.csv; done
ls -la gpurun_out | head -30
