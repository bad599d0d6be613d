set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/r02m_pytest.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/r02m_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 > gpurun_out/r02m_ncu1.log 2>&1
python scripts/bench_batched_closure.py --reps 3 > gpurun_out/r02m_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cluster_closure -s 2 -c 2 -o gpurun_out/r02_prof_cluster -f python scripts/bench_batched_closure.py --reps 3 > gpurun_out/r02m_ncu2.log 2>&1
python scripts/profile_em.py > gpurun_out/r02m_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:em_ -c 8 -o gpurun_out/r02_prof_em -f python scripts/profile_em.py > gpurun_out/r02m_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
# mid-size supports (configs[3]-shaped frames on the lock-step path): launch list and --set full of the stage kernels
python scripts/groupwise_c4.py --frames 16 --iters 1 --lockstep 1 --groups 1 > gpurun_out/r02af_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_c4_lockstep.csv python scripts/groupwise_c4.py --frames 16 --iters 1 --lockstep 1 --groups 1 > gpurun_out/r02af_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"small_adj_mid|small_rhs_step|small_mid_finish" -s 30 -c 3 -o gpurun_out/r02_prof_mid -f python scripts/groupwise_c4.py --frames 16 --iters 1 --lockstep 1 --groups 1 > gpurun_out/r02af_ncu2.log 2>&1
# summaries: python scripts/launch_summary.py X.csv ; ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_summary.py
