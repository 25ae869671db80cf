cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-spinup --no-extra --repeats 2"
case "$1" in
 A) $CMD > gpurun_out/plain_A.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu_A.log 2>&1; echo rc=$?;;
 B) $CMD > gpurun_out/plain_B.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:amp_stream -s 30 -c 1 -o gpurun_out/r02_amp_stream $CMD > gpurun_out/ncu_B.log 2>&1; echo rc=$?;;
 C) python tools/time_search.py --steps 2 > gpurun_out/plain_C.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sim_gemm_kernel -s 5 -c 1 -o gpurun_out/r02_sim_gemm_filter python tools/time_search.py --steps 2 > gpurun_out/ncu_C.log 2>&1; echo rc=$?;;
 D) python tools/shard_stage_probe.py --G 8 --steps 1 > gpurun_out/plain_D.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"owner_finalize|tc_rescore|tc_sort_pack|tc_collect|surv_hist|tau_union" -s 24 -c 6 -o gpurun_out/r02_shard_tail python tools/shard_stage_probe.py --G 8 --steps 1 > gpurun_out/ncu_D.log 2>&1; echo rc=$?;;
esac
