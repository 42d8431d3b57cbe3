# MDP-only kernel tuning sweep (inputs larger than L2, CUDA-graph replay): bash tools/mdp_sweep.sh
for tile in 128 112 104 96 80 72 64 56 48 40 32; do
echo "tile=$tile one-shot 65536"; ZBOT_MDP_TILE=$tile ZBOT_MDP_PIPE=0 python tools/bench_mdp.py 65536 20 graph-rotate
done
for tile in 112 104 72 56; do echo "tile=$tile one-shot 262144"; ZBOT_MDP_TILE=$tile ZBOT_MDP_PIPE=0 python tools/bench_mdp.py 262144 12 graph-rotate; done
