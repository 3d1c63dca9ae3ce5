mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/sharded_check.py 2>&1 | tail -8
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/scale_1.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/scale_2.json | cut -c1-400
