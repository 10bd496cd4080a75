set -x
BOPT="--steps 1 --warmup 1 --no-seeds --no-extra --no-cpu --chains 1 --spots 0 --no-region"
python bench.py $BOPT > gpurun_out/bshort.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py $BOPT > gpurun_out/ncu_l.log 2>&1
python tools/diag_chain.py 100000 300 12 > gpurun_out/diag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:prepare_kernel --launch-skip 10 --launch-count 1 -o gpurun_out/r2_prepare -f python tools/diag_chain.py 100000 300 12 > gpurun_out/ncu_p.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fix_walk_kernel --launch-skip 10 --launch-count 1 -o gpurun_out/r2_fix_walk -f python tools/diag_chain.py 100000 300 12 > gpurun_out/ncu_w.log 2>&1
python tools/prof_binning.py 100000 3 > gpurun_out/binprof.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:bin_onepass --launch-skip 2 --launch-count 1 -o gpurun_out/r2_bin_onepass -f python tools/prof_binning.py 100000 3 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out
