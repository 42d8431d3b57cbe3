TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29541 tools/e2e_ranks.py --out gpurun_out/r2n_e2e_ranks_2.json > gpurun_out/r2n_e2e2.log 2>&1; echo e2e2=$?
$TR --nproc-per-node 2 --master-port 29542 tools/bench_ppo_ranks.py --out gpurun_out/r2n_ppo_ranks_2.json > gpurun_out/r2n_ppo2.log 2>&1; echo ppo2=$?
python tools/bench_ppo_ranks.py --out gpurun_out/r2n_ppo_ranks_1.json > gpurun_out/r2n_ppo1.log 2>&1; echo ppo1=$?
$TR --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2n_bench2.json 2> gpurun_out/r2n_bench2.err; echo bench2=$?
nvidia-smi topo -m > gpurun_out/r2n_topo2.txt 2>&1
