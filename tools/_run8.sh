TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2o_topo8.txt 2>&1
lscpu | grep -i "numa\|^CPU(s)\|socket\|model name" > gpurun_out/r2o_lscpu8.txt
for N in 4 8; do
$TR --nproc-per-node $N --master-port 2955$N tools/e2e_ranks.py --out gpurun_out/r2o_e2e_ranks_$N.json > gpurun_out/r2o_e2e$N.log 2>&1; echo e2e$N=$?
done
$TR --nproc-per-node 8 --master-port 29561 tools/e2e_ranks.py --bind --out gpurun_out/r2o_e2e_ranks_8_bind.json > gpurun_out/r2o_e2e8b.log 2>&1; echo e2e8b=$?
$TR --nproc-per-node 8 --master-port 29562 tools/bench_ppo_ranks.py --out gpurun_out/r2o_ppo_ranks_8.json > gpurun_out/r2o_ppo8.log 2>&1; echo ppo8=$?
for N in 2 4 8; do
$TR --nproc-per-node $N --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2o_bench$N.json 2> gpurun_out/r2o_bench$N.err; echo bench$N=$?
done
