python -m pytest tests/test_gpu_batched.py -x -q > gpurun_out/r02af_pytest.log 2>&1; tail -2 gpurun_out/r02af_pytest.log
python scripts/groupwise_c4.py --frames 16 --iters 1 --lockstep 1 > gpurun_out/r02af_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_c4_lockstep.csv python scripts/groupwise_c4.py --frames 16 --iters 1 --lockstep 1 > gpurun_out/r02af_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_c4_lockstep.csv | head -14
ncu --set full --clock-control none --import-source on -k regex:"small_adj_mid|small_rhs_step|small_mid_finish" -s 30 -c 3 -o gpurun_out/r02_prof_mid -f python scripts/groupwise_c4.py --frames 16 --iters 1 --lockstep 1 > gpurun_out/r02af_ncu2.log 2>&1
ls -la gpurun_out/r02_prof_mid.ncu-rep
