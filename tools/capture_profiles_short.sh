T=r2k
set -x
mkdir -p /tmp/rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_bench_65536.csv \
    python bench.py --steps 2 --warmup 3 --no-small --no-mdp --no-tasks --no-e2e > gpurun_out/${T}_ncu_bench.log 2>&1; echo launches=$?
cap() {
  ncu --set full --clock-control none --import-source on -k regex:$4 -s 20 -c 2 -f -o /tmp/rep/$1 python tools/prof_step.py $2 30 $3 > gpurun_out/${T}_ncu_$1.log 2>&1; echo full_$1=$?
  ncu -i /tmp/rep/$1.ncu-rep --page raw --csv > gpurun_out/${T}_$1_raw.csv 2>/dev/null
  ncu -i /tmp/rep/$1.ncu-rep --page source --csv --kernel-id :::1 2>/dev/null | gzip > gpurun_out/${T}_$1_source.csv.gz
}
cap step65536 65536 walk zbot_step
python tools/ncu_profile_json.py gpurun_out/${T}_step65536_raw.csv 65536 zbot_step gpurun_out/${T}_step65536_source.csv.gz > gpurun_out/${T}_profjson65536.log 2>&1 && cp profiles/step_65536.json gpurun_out/
cap step4096 4096 walk zbot_step
python tools/ncu_profile_json.py gpurun_out/${T}_step4096_raw.csv 4096 zbot_step gpurun_out/${T}_step4096_source.csv.gz > gpurun_out/${T}_profjson4096.log 2>&1 && cp profiles/step_4096.json gpurun_out/
python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_with_profile.json 2> gpurun_out/${T}_bench_with_profile.err; echo bench_with_profile=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo ref=$?
